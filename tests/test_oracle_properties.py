"""CPU tier: the arithmetic identities the kernels lean on, checked against the oracle's plain formulas in numpy
fp32 (IEEE, round to nearest even, no flush to zero -- what the GPU build is compiled for).

  * max|floor(P/s)| over a group needs only the group's min P and max P;
  * Appendix A.7: max(|P| / s) == max|P| / s (division by a positive constant is monotone);
  * the threshold filter of BwdThresholdOp (csrc/lq_fq.cu): |dy| >= RN(|nz| * lam_hi) implies
    RN(|dy| / |nz|) >= lam, with lam_hi = fp32(lam * (1 + 2^-20)) -- so only elements that FAIL the filter need the
    exact division (NQ:64, NQ:81); including lam = 0 and non-finite operands.
"""
import numpy as np
import pytest

from oracle import lq_oracle as o

F32 = np.float32


def _ulp_step(x, k):
    """x moved by k units in the last place (x > 0, finite)."""
    return (x.view(np.int32) + np.int32(k)).view(np.float32)


@pytest.mark.parametrize("seed", range(4))
def test_max_code_from_group_extremes(seed):
    rng = np.random.default_rng(seed)
    P = rng.normal(0, 0.05, (64, 257)).astype(F32)
    P[3, :] = np.abs(P[3, :])                    # an all-positive and an all-negative group
    P[4, :] = -np.abs(P[4, :])
    s = rng.uniform(1e-5, 2e-2, (64, 1)).astype(F32)
    q, _ = o.fq_forward(P, s)
    direct = np.max(np.abs(q), axis=1)
    lo, hi = P.min(axis=1, keepdims=True), P.max(axis=1, keepdims=True)
    from_extremes = np.maximum(np.abs(np.floor(lo / s)), np.abs(np.floor(hi / s)))[:, 0]
    assert np.array_equal(direct, from_extremes)


@pytest.mark.parametrize("seed", range(4))
def test_maxbin_of_quotients_is_quotient_of_max(seed):
    rng = np.random.default_rng(100 + seed)
    P = rng.normal(0, 0.05, (32, 1000)).astype(F32)
    s = rng.uniform(1e-5, 2e-2, (32, 1)).astype(F32)
    assert np.array_equal(np.max(np.abs(P) / s, axis=1), np.max(np.abs(P), axis=1) / s[:, 0])


def _filter_says_above(dy, nz, lam):
    lam_hi = F32(np.float64(lam) * (1.0 + 2.0 ** -20))
    with np.errstate(all="ignore"):
        return np.abs(dy) >= (np.abs(nz) * lam_hi).astype(F32)


def _exact_above(dy, nz, lam):
    with np.errstate(all="ignore"):
        return (np.abs(dy) / np.abs(nz)).astype(F32) >= F32(lam)


@pytest.mark.parametrize("seed", range(3))
def test_threshold_filter_never_calls_a_below_element_above(seed):
    rng = np.random.default_rng(200 + seed)
    n = 400_000
    # lam inside the kernel's window [2^-60, 2^60]; nz where the filtered path can see it: eps or 2^-40 .. 2^81
    lam = np.exp2(rng.uniform(-60, 60, n)).astype(F32)
    nz = (np.exp2(rng.uniform(-40, 81, n)) * rng.choice([-1.0, 1.0], n)).astype(F32)
    nz[::50] = o.EPS_FLOAT32                       # y == 0 -> eps (NQ:63)
    # adversarial dy: the exact boundary nz * lam, then a few units in the last place either side
    boundary = (np.abs(nz).astype(np.float64) * lam.astype(np.float64)).astype(F32)
    for k in range(-6, 7):
        with np.errstate(all="ignore"):           # boundary can be inf (2^81 * 2^60): dy is then inf or NaN, both fine
            dy = _ulp_step(boundary, k) * rng.choice([-1.0, 1.0], n).astype(F32)
        above_fast, above_exact = _filter_says_above(dy, nz, lam), _exact_above(dy, nz, lam)
        assert not np.any(above_fast & ~above_exact), k
    # and unrelated magnitudes
    dy = (np.exp2(rng.uniform(-60, 60, n)) * rng.choice([-1.0, 1.0], n)).astype(F32)
    above_fast, above_exact = _filter_says_above(dy, nz, lam), _exact_above(dy, nz, lam)
    assert not np.any(above_fast & ~above_exact)
    # the filter is tight: away from the boundary it decides almost everything by itself
    assert np.mean(above_fast == above_exact) > 0.999


def test_threshold_filter_at_lambda_zero_and_with_special_gradients():
    """dy is never range-checked by the kernel, so every special value must be safe; nz is: the filtered path runs
    only for |s| in [2^-40, 2^40] and |P| in [2^-80, 2^80] (csrc/lq_common.cuh), where nz = q*s is eps (y == 0, NQ:63)
    or has 2^-40 <= |nz| <= ~2^81 -- never zero, denormal, infinite or NaN (those operands take the exact path:
    with a denormal nz the product underflows and a zero dy would wrongly pass, with an infinite one inf >= inf would)."""
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-40, 1e-38, 3e38, 1.0, -2.5e-3], F32)
    nz_vals = np.array([o.EPS_FLOAT32, 2.0 ** -40, -(2.0 ** -40), 1e-3, -0.3, 1.0, 2.0 ** 81, -(2.0 ** 81)], F32)
    dy, nz = (a.reshape(-1) for a in np.meshgrid(special, nz_vals))
    for lam in (0.0, -0.0, 2.0 ** -60, 1e-11, 1e-3, 0.05, 1000.0, 2.0 ** 60):
        above_fast, above_exact = _filter_says_above(dy, nz, F32(lam)), _exact_above(dy, nz, F32(lam))
        assert not np.any(above_fast & ~above_exact), lam
    # and the two counter-examples that the operand window exists for
    assert _filter_says_above(F32(0.0), F32(1e-45), F32(1e-11)) and not _exact_above(F32(0.0), F32(1e-45), F32(1e-11))
    assert _filter_says_above(F32(np.inf), F32(np.inf), F32(1e-3)) and not _exact_above(F32(np.inf), F32(np.inf), F32(1e-3))


def test_lambda_zero_gives_zero_scale_gradient():
    """NQ:77-79 with tanh(0) = 0: every finite ratio is above, the mean is (-)0, ds = 0 * max|q| = 0."""
    rng = np.random.default_rng(7)
    P = rng.normal(0, 0.05, (6, 300)).astype(F32)
    dy = rng.normal(0, 1e-3, (6, 300)).astype(F32)
    s = rng.uniform(5e-3, 2e-2, (6, 1)).astype(F32)
    q, y = o.fq_forward(P, s)
    _, ds = o.nested_backward(dy, q, y, s, 0.0)
    assert np.all(ds == 0.0)

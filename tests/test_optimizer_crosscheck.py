"""CPU tier: pin oracle/lq_oracle.py's Keras-2.11 optimizer restatement (adam_step / sgd_step, the oracle of SURVEY 8
f1) against two INDEPENDENT implementations of the same update rules, since Keras itself cannot be installed here:

  (1) torch.optim.SGD / torch.optim.Adam in float64, through exact algebraic mappings (constant learning rate):
        Keras SGD   m <- mu m - lr g ;  w <- w + mu m - lr g (nesterov) | w + m          (optimizer_experimental/sgd.py)
        torch SGD   b <- mu b + g    ;  w <- w - lr (g + mu b) (nesterov) | w - lr b
        identical under m = -lr b.  Keras' `weight_decay` is decoupled (w <- w - lr wd w before the update): applied
        by hand, since torch's is an L2 term on g.
        Keras Adam  w <- w - alpha_t m / (sqrt(v) + eps),  alpha_t = lr sqrt(1 - b2^t) / (1 - b1^t)      (adam.py)
        torch Adam  w <- w - lr (m / (1 - b1^t)) / (sqrt(v) / sqrt(1 - b2^t) + eps')
        identical when eps' = eps / sqrt(1 - b2^t): the test sets the group's eps before every step.
  (2) a float64 numpy evaluation of the published formulas written out of line below.

What this does NOT pin: that Keras 2.11 computes exactly these formulas (they are restated from its published source;
the reference's experiment.py:763-771 and CIFAR-10/paper_implementation/experiment.py:130-141 only construct the
optimizers).  DESIGN.md says so."""
import numpy as np
import pytest

from oracle import lq_oracle as o

torch = pytest.importorskip("torch")

STEPS = 25


def _data(seed, n=257):
    rng = np.random.default_rng(seed)
    p0 = rng.normal(0, 0.05, n).astype(np.float32)
    grads = [(rng.normal(0, 1e-2, n) * (1 + 0.1 * t)).astype(np.float32) for t in range(STEPS)]
    return p0, grads


@pytest.mark.parametrize("momentum,nesterov,wd", [(0.0, False, 0.0), (0.9, False, 0.0), (0.9, True, 0.0), (0.9, True, 5e-4)])
def test_sgd_restatement_equals_torch_sgd(momentum, nesterov, wd):
    lr = 0.1
    p0, grads = _data(1)
    p, m = p0.copy(), (np.zeros_like(p0) if momentum else None)
    w = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.SGD([w], lr=lr, momentum=momentum, nesterov=nesterov)
    for g in grads:
        p, m = o.sgd_step(p, g, m, lr, momentum=momentum, nesterov=nesterov, weight_decay=wd)
        with torch.no_grad():
            w.mul_(1.0 - lr * wd)                       # Keras' decoupled decay, before the update
        w.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
    want = w.detach().numpy()
    assert np.max(np.abs(p - want)) <= 2e-6 * np.max(np.abs(want))      # fp32 restatement against the fp64 run
    if momentum:
        buf = opt.state[w]["momentum_buffer"].numpy()
        assert np.max(np.abs(m - (-lr * buf))) <= 2e-6 * np.max(np.abs(buf)) * lr


@pytest.mark.parametrize("l2", [0.0, 1e-4])
def test_adam_restatement_equals_torch_adam(l2):
    lr, b1, b2, eps = 1e-4, 0.9, 0.999, 1e-7
    p0, grads = _data(2)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    w = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([w], lr=lr, betas=(b1, b2), eps=eps)
    for t, g in enumerate(grads, start=1):
        p, m, v = o.adam_step(p, g, m, v, lr, t, beta_1=b1, beta_2=b2, epsilon=eps, l2=l2)
        opt.param_groups[0]["eps"] = eps / np.sqrt(1.0 - b2 ** t)
        w.grad = torch.tensor(g, dtype=torch.float64) + 2.0 * l2 * w.detach()      # keras.regularizers.l2: l2 * sum(w^2)
        opt.step()
    want = w.detach().numpy()
    # Adam's first steps move every weight by ~lr whatever the gradient: compare the DISPLACEMENT.  The restatement keeps
    # the parameter in fp32 like Keras (half an ulp of |w| ~ 0.15 per step = 7e-9 against steps of 1e-4): 5e-5 of the
    # displacement after 25 steps; a wrong formula (epsilon placement, bias correction) is off by 1e-2 and more.
    assert np.max(np.abs((p - p0) - (want - p0))) <= 5e-5 * np.max(np.abs(want - p0))


def test_restatements_equal_the_formulas_in_float64():
    lr, b1, b2, eps, mu, wd = 1e-3, 0.9, 0.999, 1e-7, 0.9, 5e-4
    p0, grads = _data(3)
    pa, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    ps, ms = p0.copy(), np.zeros_like(p0)
    A, M, V = p0.astype(np.float64), np.zeros(p0.size), np.zeros(p0.size)
    S, MS = p0.astype(np.float64), np.zeros(p0.size)
    for t, g in enumerate(grads, start=1):
        pa, m, v = o.adam_step(pa, g, m, v, lr, t, min_value=-0.02)
        ps, ms = o.sgd_step(ps, g, ms, lr, momentum=mu, nesterov=True, weight_decay=wd, min_value=-0.02)
        G = g.astype(np.float64)
        M = b1 * M + (1 - b1) * G
        V = b2 * V + (1 - b2) * G * G
        A = np.maximum(A - lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t) * M / (np.sqrt(V) + eps), -0.02)   # MinValueConstraint NQ:43
        S = S * (1 - lr * wd)
        MS = mu * MS - lr * G
        S = np.maximum(S + mu * MS - lr * G, -0.02)
    assert np.max(np.abs((pa - p0) - (A - p0))) <= 5e-5 * np.max(np.abs(A - p0))
    assert np.max(np.abs((ps - p0) - (S - p0))) <= 5e-5 * np.max(np.abs(S - p0))

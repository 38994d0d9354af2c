"""CPU tier: the oracle against every committed output of the reference's own source."""
import numpy as np
import pytest

from conftest import bits_equal, rel_err
from oracle import lq_oracle as o

TOL = 1e-5   # BASELINE.json north_star: scale gradients and loss terms within 1e-5 relative


def test_appendix_b_known_answers():
    P = np.array([[0.30, -0.26, 0.04], [-0.05, 0.0, 0.11]], np.float32)
    s = np.array([[0.1], [0.02]], np.float32)
    dy = np.array([[0.003, -0.5, 0.0], [0.2, 0.1, -0.3]], np.float32)
    q, y = o.fq_forward(P, s)
    assert q.tolist() == [[3, -3, 0], [-3, 0, 5]]
    assert y.astype(np.float64).tolist() == [[0.30000001192092896, -0.30000001192092896, 0.0],
                                              [-0.05999999865889549, 0.0, 0.09999999403953552]]
    assert o.nested_backward(dy, q, y, s, 0.0)[1].ravel().tolist() == [0.0, 0.0]
    ds = o.nested_backward(dy, q, y, s, 0.05)[1].ravel().astype(np.float64)
    assert rel_err(ds, [-0.08993705362081528, -0.24979187548160553]) < 1e-6
    ds = o.nested_backward(dy, q, y, s, 1000.0)[1].ravel().astype(np.float64)
    assert rel_err(ds, [-3.0, -3.3333334922790527]) < 1e-6
    s1 = np.array([0.1], np.float32)
    q1, y1 = o.fq_forward(P, s1)
    assert q1.tolist() == [[3, -3, 0], [-1, 0, 1]]
    assert rel_err(o.nested_backward(dy, q1, y1, s1, 0.05)[1], [-0.04496852681040764]) < 1e-6
    L = [(P, s, np.array([0.05, -0.2], np.float32), np.array([0.1], np.float32))]
    assert float(o.maxbin_penalty(L)) == 3.6875
    assert rel_err(o.difference_penalty(L), 1.9362499713897705) < 1e-6
    assert float(o.inverse_penalty(L)) == 25.0


def test_scale_shapes_match_reference(golden):
    for c in golden.manifest["shapes"]:
        assert list(o.scale_shape(c["param"], c["orientation"])) == c["scale"], c
    with pytest.raises(ValueError) as e:
        o.scale_shape((3, 3), "diagonal")
    assert str(e.value) == golden.manifest["bad_orientation_raises"]


def test_forward_bit_exact_and_nested_ds(golden):
    n = 0
    for c, P, s, dy, q_ref, y_ref, ds_ref in golden.op_cases():
        q, y = o.fq_forward(P, s)
        assert bits_equal(q, q_ref), c
        assert bits_equal(y, y_ref), c
        for accum in ("fp32", "fp64"):
            dP, ds = o.nested_backward(dy, q, y, s, c["lam"], accum=accum)
            assert dP is dy or bits_equal(dP, dy)
            assert ds.shape == ds_ref.shape, c
            assert rel_err(ds, ds_ref) < TOL, (c, ds, ds_ref)
        n += 1
    assert n >= 200


def test_ste_variants(golden):
    for c in golden.manifest["ste"]:
        g = {k: golden[f"{c['id']}/{k}"] for k in ("P", "s", "dy", "q", "y", "ds")}
        q, y = o.fq_forward(g["P"], g["s"])
        assert bits_equal(q, g["q"]) and bits_equal(y, g["y"]), c
        dP, ds = o.ste_backward(g["dy"], g["s"])
        assert bits_equal(dP, g["dy"]) and bits_equal(ds, g["ds"]) and not ds.any(), c


def test_loss_terms_and_grads(golden):
    fns = {"compute_maxbin_penalty": (o.maxbin_penalty, o.maxbin_penalty_grads),
           "compute_difference_penalty": (o.difference_penalty, o.difference_penalty_grads),
           "compute_inverse_penalty": (o.inverse_penalty, o.inverse_penalty_grads)}
    for c in golden.manifest["loss"]:
        layers = golden.loss_layers(c)
        val, grads = fns[c["method"]]
        assert rel_err(val(layers), golden[f"{c['id']}/penalty"]) < TOL, c
        for accum in ("fp32", "fp64"):
            for got, ref in zip(grads(layers, 1.0, accum=accum), golden.loss_grads(c)):
                for name, a, b in zip(("K", "sK", "b", "sb"), got, ref):
                    assert a.shape == b.shape, (c, name)
                    scale = max(float(np.max(np.abs(b))), 1e-30)
                    assert float(np.max(np.abs(a.astype(np.float64) - b))) <= TOL * scale, (c, name)
        if not c.get("no_total"):
            # LF:47-73: per-sample scce + gamma*penalty
            yt, yp = golden[f"{c['id']}/y_true"], golden[f"{c['id']}/y_pred"]
            scce = -np.log(np.clip(yp[np.arange(len(yt)), yt], 1e-7, 1.0)).astype(np.float32)
            tot = scce + np.float32(c["gamma"]) * val(layers)
            assert rel_err(tot, golden[f"{c['id']}/total_loss"]) < TOL, c


def test_min_value_constraint(golden):
    out = o.min_value_constraint(golden["constraint/in"], golden.manifest["constraint"][0]["min_value"])
    assert bits_equal(out, golden["constraint/out"])
    assert float(o.SCALE_INIT) == 1.1920928955078125e-05


def test_group_layout_table():
    # SURVEY 8a unification table
    assert o.group_layout((24, 10), (1,)) == (1, 1, 240)
    assert o.group_layout((24, 10), (24, 1)) == (1, 24, 10)
    assert o.group_layout((24, 10), (1, 10)) == (24, 10, 1)
    assert o.group_layout((24, 10), (1, 1)) == (1, 1, 240)
    assert o.group_layout((3, 3, 8, 16), (3, 1, 1, 1)) == (1, 3, 3 * 8 * 16)
    assert o.group_layout((3, 3, 8, 16), (1, 3, 1, 1)) == (3, 3, 8 * 16)
    assert o.group_layout((3, 3, 8, 16), (1, 1, 8, 1)) == (9, 8, 16)
    assert o.group_layout((16,), (1,)) == (1, 1, 16)
    for shape in [(3, 3, 8, 16), (24, 10), (5, 3, 6, 7)]:
        for orient in ("rowwise", "columnwise", "channelwise", "scalar"):
            ss = o.scale_shape(shape, orient)
            outer, C, inner = o.group_layout(shape, ss)
            idx = np.arange(int(np.prod(shape)))
            g = (idx // inner) % C
            s = np.arange(C, dtype=np.float32).reshape(ss) if len(ss) > 1 else np.zeros(ss, np.float32)
            assert np.array_equal(np.broadcast_to(s, shape).reshape(-1), g.astype(np.float32))


def test_rne_clamp_extension():
    P = np.array([0.25, 0.35, -0.25, 1.0, -1.0, 0.05], np.float32)
    s = np.array([0.1], np.float32)
    q, y = o.fq_forward(P, s, round_mode=o.ROUND_RNE, qmin=-4, qmax=4)
    assert q.tolist() == [2.0, 4.0, -2.0, 4.0, -4.0, 0.0]   # 2.5->2, 3.5->4 (ties to even), clamp at +-4

"""CPU tier: the C restatement (oracle/lq_oracle.c) against the numpy oracle and the fixtures."""
import numpy as np
import pytest

from conftest import bits_equal, rel_err
from oracle import lq_oracle as o

oc = pytest.importorskip("oracle.lq_oracle_c")


@pytest.fixture(scope="module")
def co():
    try:
        return oc.COracle()
    except ImportError as e:
        pytest.skip(str(e))


def test_c_oracle_matches_golden_ops(co, golden):
    n = 0
    for c, P, s, dy, q_ref, y_ref, ds_ref in golden.op_cases():
        layout = o.group_layout(P.shape, s.shape)
        q, y, ds = co.fwd_bwd(P.reshape(-1).copy(), s.reshape(-1).copy(), dy.reshape(-1).copy(), layout, c["lam"])
        assert bits_equal(q.reshape(P.shape), q_ref) and bits_equal(y.reshape(P.shape), y_ref), c
        assert rel_err(ds.reshape(ds_ref.shape), ds_ref) < 1e-5, (c, ds, ds_ref)
        n += 1
    assert n >= 200


def test_c_oracle_penalties_match_golden(co, golden):
    kinds = {"compute_maxbin_penalty": "maxbin", "compute_difference_penalty": "difference",
             "compute_inverse_penalty": "inverse"}
    for c in golden.manifest["loss"]:
        tensors = []
        for K, sK, b, sb in golden.loss_layers(c):
            tensors.append((K.reshape(-1).copy(), sK.reshape(-1).copy(), o.group_layout(K.shape, sK.shape)))
            tensors.append((b.reshape(-1).copy(), sb.reshape(-1).copy(), o.group_layout(b.shape, sb.shape)))
        assert rel_err(co.penalty(kinds[c["method"]], tensors), golden[f"{c['id']}/penalty"]) < 1e-5, c


def test_c_oracle_matches_numpy_random(co):
    rng = np.random.default_rng(9)
    for layout in [(7, 5, 130), (300, 12, 1), (1, 1, 4097), (2, 64, 33)]:
        n = layout[0] * layout[1] * layout[2]
        P = rng.normal(0, 0.05, n).astype(np.float32)
        dy = rng.normal(0, 1e-3, n).astype(np.float32)
        s = rng.uniform(5e-3, 2e-2, layout[1]).astype(np.float32)
        q, y, ds = co.fwd_bwd(P, s, dy, layout, 1e-3)
        q_ref, y_ref = o.fq_forward(P.reshape(layout), s.reshape(1, -1, 1))
        _, ds_ref = o.nested_backward(dy.reshape(layout), q_ref, y_ref, s.reshape(1, -1, 1), 1e-3, accum="fp64")
        assert bits_equal(y.reshape(layout), y_ref)
        assert rel_err(ds, ds_ref.reshape(-1)) < 1e-6

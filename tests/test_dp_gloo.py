"""CPU tier, world_size 2 over gloo: the data-parallel gradient exchange (training/dp.py).

The CUDA kernels cannot run here, so the scale backward is replaced by the ORACLE's threshold rule
(test infrastructure) -- what is under test is the host logic: bucket layout, what is exchanged in
which order, and the semantic difference between the two modes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPES_Q = [(3, 3, 4, 8), (8,), (16, 5)]
SHAPES_S = [(1, 1, 4, 1), (1,), (1, 5)]
SHAPES_P = [(8,), (8,), (5, 2)]
LAM = 0.05


def _inputs(rank):
    rng = np.random.default_rng(7)
    Ps = [rng.normal(0, 0.05, s).astype(np.float32) for s in SHAPES_Q]
    ss = [rng.uniform(5e-3, 2e-2, s).astype(np.float32) for s in SHAPES_S]
    rng_r = np.random.default_rng(100 + rank)
    dys = [rng_r.normal(0, 1e-3, s).astype(np.float32) for s in SHAPES_Q]
    plain = [rng_r.normal(0, 1.0, s).astype(np.float32) for s in SHAPES_P]
    return Ps, ss, dys, plain


def _oracle_ds(Ps, ss, dys):
    from oracle import lq_oracle as o
    out = []
    for P, s, dy in zip(Ps, ss, dys):
        q, y = o.fq_forward(P, s)
        out.append(o.nested_backward(dy, q, y, s, LAM)[1])
    return out


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from learned_quantization_b200.training.dp import GradBucket, TorchDistComm, exchange_gradients
        Ps, ss, dys, plain = _inputs(rank)
        bucket = GradBucket(SHAPES_Q, SHAPES_P, SHAPES_S, "cpu")
        for v, a in zip(bucket.dy + bucket.plain, dys + plain):
            v.copy_(torch.from_numpy(a))

        def scale_backward():
            cur = [v.numpy().copy() for v in bucket.dy]
            for v, ds in zip(bucket.ds, _oracle_ds(Ps, ss, cur)):
                v.copy_(torch.from_numpy(np.asarray(ds, np.float32).reshape(v.shape)))

        exchange_gradients(bucket, TorchDistComm(), scale_backward, mode)
        q.put((rank, [v.numpy().copy() for v in bucket.dy], [v.numpy().copy() for v in bucket.plain],
               [v.numpy().copy() for v in bucket.ds]))
    finally:
        dist.destroy_process_group()


def _run(mode, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_bucket_layout_is_aligned_and_ordered():
    from learned_quantization_b200.training.dp import GradBucket
    b = GradBucket(SHAPES_Q, SHAPES_P, SHAPES_S, "cpu")
    ptrs = [v.data_ptr() for v in b.dy + b.plain + b.ds]
    assert ptrs == sorted(ptrs) and all((p - b.flat.data_ptr()) % 16 == 0 for p in ptrs)
    assert b.exact_part().numel() == b.ds[0].data_ptr() // 4 - b.flat.data_ptr() // 4
    assert [tuple(v.shape) for v in b.ds] == SHAPES_S
    b.flat.fill_(1.0)
    assert all(float(v.sum()) == v.numel() for v in b.dy + b.plain + b.ds)
    with pytest.raises(ValueError):
        from learned_quantization_b200.training.dp import exchange_gradients, SingleComm
        exchange_gradients(b, SingleComm(), lambda: None, "sometimes")


def test_exact_mode_equals_single_device_on_the_global_batch():
    (r0, dy0, pl0, ds0), (r1, dy1, pl1, ds1) = _run("exact", 29611)
    Ps, ss, a, pa = _inputs(0)
    _, _, b, pb = _inputs(1)
    mean_dy = [(x + y) * np.float32(0.5) for x, y in zip(a, b)]
    for got0, got1, want in zip(dy0, dy1, mean_dy):
        assert np.array_equal(got0, got1) and np.allclose(got0, want, rtol=0, atol=1e-9)
    for got0, got1, x, y in zip(pl0, pl1, pa, pb):
        assert np.array_equal(got0, got1) and np.allclose(got0, (x + y) * 0.5, rtol=1e-6)
    # ds is the reference rule evaluated on the averaged dy, bit-identical on both ranks
    want_ds = _oracle_ds(Ps, ss, [np.asarray(v, np.float32) for v in dy0])
    for got0, got1, want in zip(ds0, ds1, want_ds):
        assert np.array_equal(got0, got1)
        assert np.array_equal(got0.reshape(-1), np.asarray(want, np.float32).reshape(-1))


def test_local_mode_averages_the_per_rank_scale_gradients():
    (r0, dy0, pl0, ds0), (r1, dy1, pl1, ds1) = _run("local", 29612)
    Ps, ss, a, _ = _inputs(0)
    _, _, b, _ = _inputs(1)
    da, db = _oracle_ds(Ps, ss, a), _oracle_ds(Ps, ss, b)
    exact = _oracle_ds(Ps, ss, [(x + y) * np.float32(0.5) for x, y in zip(a, b)])
    differs = False
    for got0, got1, x, y, e in zip(ds0, ds1, da, db, exact):
        want = (np.asarray(x, np.float32) + np.asarray(y, np.float32)).reshape(-1) * np.float32(0.5)
        assert np.array_equal(got0, got1) and np.allclose(got0.reshape(-1), want, rtol=1e-6, atol=1e-12)
        differs |= not np.allclose(want, np.asarray(e, np.float32).reshape(-1), rtol=1e-3)
    assert differs, "the two modes are expected to differ: ds is not linear in dy"

"""GPU tier, ONE device: the fused peer-memory gradient exchange (lq_dpx_*, include/lq.h).

The ranks of a real run are processes on different GPUs (tests/test_gpu_dp.py, two devices).  Here the
same kernels are driven as several "ranks" inside one process on one device -- one arena per rank,
connected by raw pointers (lq_dpx_connect_ptrs), one stream per rank, few enough blocks to be resident
together -- so the exchange logic (ownership, rank-order mean, in-place stores, partials to every rank,
device-side epochs, the time-out) is checked on the single-GPU box the round-end tests run on.
Oracle: the mean is (dy_0 + dy_1 + ...) * (1/W) in rank order, bit for bit; ds must equal the
single-device lq_mt_fq_bwd on that mean bit for bit (same partial layout, same second stage)."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal


def ds_close(got, ref, tol=1e-5):
    """per-group relative error with a floor of 1e-6 of the largest group (as tests/test_gpu_parity.py)"""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = max(float(np.max(np.abs(ref))), 1e-30) if ref.size else 1.0
    return bool(np.all(np.abs(got - ref) <= tol * np.maximum(np.abs(ref), 1e-6 * scale)))

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

LAM = 1e-3
# ROW (inner >= 128), COL (interleaved, one column per thread), per-tensor, a bias, an odd-sized tensor
SHAPES = [((3, 3, 16, 256), (1, 1, 16, 1)), ((576, 64, 1), (1, 64, 1)), ((7, 7, 3, 64), (1,)), ((64,), (1,)),
          ((37, 5, 131), (1, 5, 1)), ((128, 10), (128, 1))]
N_PLAIN = 1234


@pytest.fixture(scope="module")
def lq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import learned_quantization_b200 as m
    torch.cuda.set_device(0)
    return m


class Rank:
    """One rank's arena with the bucket [dy_0 | dy_1 | ... | plain] and the partial workspace inside it."""

    def __init__(self, lq, rank, world, layouts, sizes, ws_bytes):
        from learned_quantization_b200._lib import check, lib
        self.lib, self.check = lib, check
        self.h = C.c_void_p()
        handle = C.create_string_buffer(64)
        base = C.c_void_p()
        self.offs, off = [], 0
        for n in sizes + [N_PLAIN]:
            self.offs.append(off)
            off += -(-n // 4) * 4
        self.n_floats = off
        self.ws_off = -(-off * 4 // 256) * 256
        nbytes = self.ws_off + max(ws_bytes, 256)
        check(lib.lq_dpx_create(rank, world, nbytes, C.byref(self.h), handle, C.byref(base)))
        from learned_quantization_b200.training.dp import _DeviceSpan
        self.base = base.value
        self.arena = torch.as_tensor(_DeviceSpan(base.value, -(-nbytes // 256) * 256, self), device="cuda:0")
        self.flat = self.arena[: off * 4].view(torch.float32)
        self.ws = self.arena[self.ws_off:]
        self.stream = torch.cuda.Stream()

    def view(self, i, n):
        return self.flat[self.offs[i]: self.offs[i] + n]

    def close(self):
        self.check(self.lib.lq_dpx_destroy(self.h))


def make_world(lq, world, rng):
    ops = lq.ops
    Ps = [torch.from_numpy(rng.normal(0, 0.05, size=sh).astype(np.float32)).cuda() for sh, _ in SHAPES]
    ss = [torch.from_numpy(rng.uniform(5e-3, 2e-2, size=ssh).astype(np.float32)).cuda() for _, ssh in SHAPES]
    layouts = [ops.tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
    sizes = [p.numel() for p in Ps]
    ws_bytes = ops.workspace_bytes([ops.make_desc(layout=l) for l in layouts], lq._lib.LQ_PHASE_BWD)
    ranks = [Rank(lq, r, world, layouts, sizes, ws_bytes) for r in range(world)]
    ptrs = (C.c_void_p * world)(*[r.base for r in ranks])
    for r in ranks:
        r.check(r.lib.lq_dpx_connect_ptrs(r.h, ptrs))
    return Ps, ss, layouts, sizes, ranks


def fill(ranks, sizes, rng):
    """fresh local gradients on every rank; returns them as numpy arrays [rank][tensor]"""
    out = []
    for r in ranks:
        mine = []
        for i, n in enumerate(sizes + [N_PLAIN]):
            a = rng.normal(0, 1e-3, size=n).astype(np.float32)
            r.view(i, n).copy_(torch.from_numpy(a))
            mine.append(a)
        out.append(mine)
    torch.cuda.synchronize()
    return out


def launch(lq, rk, Ps, ss, layouts, sizes, dss, mode, channel=0):
    ops = lq.ops
    descs = ops.desc_array([ops.make_desc(P=p, s=s, dy=rk.view(i, n), ds=ds, layout=l)
                            for i, (p, s, n, ds, l) in enumerate(zip(Ps, ss, sizes, dss, layouts))])
    plain = rk.view(len(sizes), N_PLAIN)
    with torch.cuda.stream(rk.stream):
        rk.check(rk.lib.lq_dpx_exchange(rk.h, descs, len(descs), plain.data_ptr(), N_PLAIN, LAM, mode, channel, 0,
                                        rk.ws.data_ptr(), rk.ws.numel(), rk.stream.cuda_stream))
    return descs


def mean_in_rank_order(parts):
    acc = parts[0].copy()
    for a in parts[1:]:
        acc = (acc + a).astype(np.float32)
    return (acc * np.float32(1.0 / len(parts))).astype(np.float32)


def reference_ds(lq, Ps, ss, layouts, means):
    ops = lq.ops
    dys = [torch.from_numpy(m).cuda().view(p.shape) for m, p in zip(means, Ps)]
    dss = [torch.empty_like(s) for s in ss]
    ops.mt_fq_bwd(dys, Ps, ss, dss, LAM, layouts)
    torch.cuda.synchronize()
    return [d.cpu().numpy() for d in dss]


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_fused_exchange_equals_mean_then_single_device_backward(lq, world):
    rng = np.random.default_rng(100 + world)
    Ps, ss, layouts, sizes, ranks = make_world(lq, world, rng)
    try:
        for rep in range(2):                    # twice: the device-side epochs must advance by themselves
            local = fill(ranks, sizes, rng)
            dss = [[torch.full_like(s, 7.0) for s in ss] for _ in ranks]
            keep = [launch(lq, rk, Ps, ss, layouts, sizes, dss[r], lq._lib.LQ_SGRAD_THRESHOLD) for r, rk in enumerate(ranks)]
            torch.cuda.synchronize()
            for rk in ranks:
                flag = C.c_int(0)
                rk.check(rk.lib.lq_dpx_error(rk.h, C.byref(flag)))
                assert flag.value == 0, "a rank timed out"
            means = [mean_in_rank_order([local[r][i] for r in range(world)]) for i in range(len(sizes) + 1)]
            want_ds = reference_ds(lq, Ps, ss, layouts, means[:-1])
            for r, rk in enumerate(ranks):
                for i, n in enumerate(sizes + [N_PLAIN]):
                    assert bits_equal(rk.view(i, n).cpu().numpy(), means[i]), (rep, r, i)
                for i in range(len(sizes)):
                    # the same scale gradient as the single-device backward of the mean (the two cut interleaved tensors into
                    # row chunks of different lengths: equal within rounding), and the same BITS on every rank
                    assert ds_close(dss[r][i].cpu().numpy().ravel(), want_ds[i].ravel()), (rep, r, i)
                    assert bits_equal(dss[r][i].cpu().numpy().ravel(), dss[0][i].cpu().numpy().ravel()), (rep, r, i)
            del keep
    finally:
        torch.cuda.synchronize()
        for rk in ranks:
            rk.close()


def test_ste_mode_only_averages_and_zeroes_ds(lq):
    rng = np.random.default_rng(7)
    Ps, ss, layouts, sizes, ranks = make_world(lq, 2, rng)
    try:
        local = fill(ranks, sizes, rng)
        dss = [[torch.full_like(s, 7.0) for s in ss] for _ in ranks]
        keep = [launch(lq, rk, Ps, ss, layouts, sizes, dss[r], lq._lib.LQ_SGRAD_NONE) for r, rk in enumerate(ranks)]
        torch.cuda.synchronize()
        for r, rk in enumerate(ranks):
            for i, n in enumerate(sizes + [N_PLAIN]):
                assert bits_equal(rk.view(i, n).cpu().numpy(), mean_in_rank_order([local[0][i], local[1][i]]))
            assert all(not d.cpu().numpy().any() for d in dss[r])
        del keep
    finally:
        torch.cuda.synchronize()
        for rk in ranks:
            rk.close()


def test_replayed_from_cuda_graphs(lq):
    """The launch is captured once per rank and replayed on new gradients: epochs live on the device."""
    rng = np.random.default_rng(8)
    Ps, ss, layouts, sizes, ranks = make_world(lq, 2, rng)
    try:
        dss = [[torch.empty_like(s) for s in ss] for _ in ranks]
        fill(ranks, sizes, rng)
        keep = [launch(lq, rk, Ps, ss, layouts, sizes, dss[r], lq._lib.LQ_SGRAD_THRESHOLD) for r, rk in enumerate(ranks)]   # warm
        torch.cuda.synchronize()
        graphs = []
        for r, rk in enumerate(ranks):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=rk.stream):
                keep.append(launch(lq, rk, Ps, ss, layouts, sizes, dss[r], lq._lib.LQ_SGRAD_THRESHOLD))
            graphs.append(g)
        for rep in range(3):
            local = fill(ranks, sizes, rng)
            for g, rk in zip(graphs, ranks):
                with torch.cuda.stream(rk.stream):
                    g.replay()
            torch.cuda.synchronize()
            means = [mean_in_rank_order([local[r][i] for r in range(2)]) for i in range(len(sizes) + 1)]
            want_ds = reference_ds(lq, Ps, ss, layouts, means[:-1])
            for r, rk in enumerate(ranks):
                flag = C.c_int(0)
                rk.check(rk.lib.lq_dpx_error(rk.h, C.byref(flag)))
                assert flag.value == 0
                for i, n in enumerate(sizes + [N_PLAIN]):
                    assert bits_equal(rk.view(i, n).cpu().numpy(), means[i]), (rep, r, i)
                for i in range(len(sizes)):
                    # the same scale gradient as the single-device backward of the mean (the two cut interleaved tensors into
                    # row chunks of different lengths: equal within rounding), and the same BITS on every rank
                    assert ds_close(dss[r][i].cpu().numpy().ravel(), want_ds[i].ravel()), (rep, r, i)
                    assert bits_equal(dss[r][i].cpu().numpy().ravel(), dss[0][i].cpu().numpy().ravel()), (rep, r, i)
    finally:
        torch.cuda.synchronize()
        for rk in ranks:
            rk.close()


def test_missing_peer_times_out_instead_of_hanging(lq):
    rng = np.random.default_rng(9)
    Ps, ss, layouts, sizes, ranks = make_world(lq, 2, rng)
    try:
        fill(ranks, sizes, rng)
        dss = [torch.empty_like(s) for s in ss]
        keep = launch(lq, ranks[0], Ps, ss, layouts, sizes, dss, lq._lib.LQ_SGRAD_THRESHOLD)    # rank 1 never shows up
        flag = C.c_int(0)
        ranks[0].check(ranks[0].lib.lq_dpx_error(ranks[0].h, C.byref(flag)))                     # synchronises (~2 s)
        assert flag.value == 1
        del keep
    finally:
        torch.cuda.synchronize()
        for rk in ranks:
            rk.close()


def test_arguments_are_checked(lq):
    from learned_quantization_b200._lib import LqError
    rng = np.random.default_rng(10)
    Ps, ss, layouts, sizes, ranks = make_world(lq, 1, rng)
    rk = ranks[0]
    ops = lq.ops
    try:
        outside = torch.zeros(sizes[0], device="cuda")
        ds = torch.empty_like(ss[0])
        d = ops.desc_array([ops.make_desc(P=Ps[0], s=ss[0], dy=outside, ds=ds, layout=layouts[0])])
        with pytest.raises(LqError, match="outside the arena"):
            rk.check(rk.lib.lq_dpx_exchange(rk.h, d, 1, None, 0, LAM, 1, 0, 0, rk.ws.data_ptr(), rk.ws.numel(), 0))
        d = ops.desc_array([ops.make_desc(P=Ps[0], s=ss[0], dy=rk.view(0, sizes[0]), dX=outside, ds=ds, layout=layouts[0])])
        with pytest.raises(LqError, match="in place"):
            rk.check(rk.lib.lq_dpx_exchange(rk.h, d, 1, None, 0, LAM, 1, 0, 0, rk.ws.data_ptr(), rk.ws.numel(), 0))
        d = ops.desc_array([ops.make_desc(P=Ps[0], s=ss[0], dy=rk.view(0, sizes[0]), ds=ds, layout=layouts[0])])
        with pytest.raises(LqError, match="workspace"):
            rk.check(rk.lib.lq_dpx_exchange(rk.h, d, 1, None, 0, LAM, 1, 0, 0, rk.ws.data_ptr(), 16, 0))
    finally:
        rk.close()


@pytest.mark.parametrize("world,n", [(1, 4096), (2, 2048 * 5 + 12), (3, 2048 * 40 + 4), (4, 1000), (8, 2048 * 19 + 2044)])
def test_copy_engine_allreduce_is_the_rank_order_mean(lq, world, n):
    """lq_dpx_allreduce: bulk asynchronous copies through a shared-memory ring; same bits as the register kernel."""
    from learned_quantization_b200._lib import check, lib
    from learned_quantization_b200.training.dp import _DeviceSpan
    rng = np.random.default_rng(200 + world)
    hs, bases, flats, streams = [], [], [], []
    try:
        for r in range(world):
            h, handle, base = C.c_void_p(), C.create_string_buffer(64), C.c_void_p()
            check(lib.lq_dpx_create(r, world, n * 4 + 256, C.byref(h), handle, C.byref(base)))
            hs.append(h); bases.append(base.value)
            flats.append(torch.as_tensor(_DeviceSpan(base.value, n * 4, None), device="cuda:0").view(torch.float32))
            streams.append(torch.cuda.Stream())
        ptrs = (C.c_void_p * world)(*bases)
        for h in hs:
            check(lib.lq_dpx_connect_ptrs(h, ptrs))
        for rep in range(3):                       # more tiles than stages on the second and third pass too
            local = [rng.normal(0, 1e-3, size=n).astype(np.float32) for _ in range(world)]
            for f, a in zip(flats, local):
                f.copy_(torch.from_numpy(a))
            torch.cuda.synchronize()
            for h, f, st in zip(hs, flats, streams):
                check(lib.lq_dpx_allreduce(h, f.data_ptr(), n, 2, 2, st.cuda_stream))     # channel 2, two blocks per rank
            torch.cuda.synchronize()
            want = mean_in_rank_order(local)
            for r, (h, f) in enumerate(zip(hs, flats)):
                flag = C.c_int(0)
                check(lib.lq_dpx_error(h, C.byref(flag)))
                assert flag.value == 0, "a rank timed out"
                assert bits_equal(f.cpu().numpy(), want), (rep, r)
    finally:
        torch.cuda.synchronize()
        for h in hs:
            check(lib.lq_dpx_destroy(h))


# ------------------------------------------------------------------ the sharded microbench's fused exchange + finish
def _p2p_world(world, C_):
    from learned_quantization_b200._lib import check, lib
    hs, bufs = [], []
    for r in range(world):
        h, handle, buf = C.c_void_p(), C.create_string_buffer(64), C.c_void_p()
        check(lib.lq_p2p_create(r, world, C_, C.byref(h), handle))
        check(lib.lq_p2p_buffer(h, C.byref(buf)))
        hs.append(h); bufs.append(buf.value)
    ptrs = (C.c_void_p * world)(*bufs)
    for h in hs:
        check(lib.lq_p2p_connect_ptrs(h, ptrs))
    return hs


@pytest.mark.parametrize("world", [2, 4])
def test_p2p_finish_equals_gather_then_finish(lq, world):
    """lq_fq_bwd_threshold_finish_p2p (statistics pushed into every rank's slots, flags, fold in rank order) against
    lq_fq_bwd_threshold_finish on the gathered statistics: bit for bit, on every rank, twice (epoch parity)."""
    from learned_quantization_b200._lib import check, lib
    ops = lq.ops
    C_ = 300
    rng = np.random.default_rng(40 + world)
    hs = _p2p_world(world, C_)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        for rep in range(3):
            stats = [torch.from_numpy(np.stack([rng.normal(0, 1, C_), rng.integers(0, 3, C_).astype(np.float64),
                                                rng.integers(0, 50, C_).astype(np.float64)], axis=1)).cuda().contiguous()
                     for _ in range(world)]
            want = ops.fq_bwd_threshold_finish(torch.stack(stats), 1e-3, 12345.0)
            outs = [torch.empty(C_, device="cuda") for _ in range(world)]
            torch.cuda.synchronize()
            for h, st, out, sm in zip(hs, stats, outs, streams):
                check(lib.lq_fq_bwd_threshold_finish_p2p(h, st.data_ptr(), C_, 1e-3, 12345.0, out.data_ptr(), sm.cuda_stream))
            torch.cuda.synchronize()
            for h, out in zip(hs, outs):
                flag = C.c_int(0)
                check(lib.lq_p2p_error(h, C.byref(flag)))
                assert flag.value == 0
                assert bits_equal(out.cpu().numpy(), want.cpu().numpy()), rep
    finally:
        torch.cuda.synchronize()
        for h in hs:
            check(lib.lq_p2p_destroy(h))


def test_p2p_finish_times_out_without_its_peer(lq):
    from learned_quantization_b200._lib import check, lib
    C_ = 64
    hs = _p2p_world(2, C_)
    try:
        st = torch.ones((C_, 3), dtype=torch.float64, device="cuda")
        out = torch.zeros(C_, device="cuda")
        check(lib.lq_fq_bwd_threshold_finish_p2p(hs[0], st.data_ptr(), C_, 1e-3, 10.0, out.data_ptr(), 0))   # rank 1 never launches
        flag = C.c_int(0)
        check(lib.lq_p2p_error(hs[0], C.byref(flag)))          # synchronises: ~2 s
        assert flag.value == 1 and bool(torch.isnan(out).all())
    finally:
        torch.cuda.synchronize()
        for h in hs:
            check(lib.lq_p2p_destroy(h))

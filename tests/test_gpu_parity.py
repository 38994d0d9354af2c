"""GPU tier (-m gpu): the CUDA path, called through the C-ABI (ctypes), against the oracle
and the committed reference-source fixtures.  Forward outputs and codes must be BIT-EXACT;
scale gradients / loss terms within 1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import bits_equal, rel_err
from oracle import lq_oracle as o

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def lq():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    import learned_quantization_b200 as pkg
    return pkg


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


WORST = {"ds": 0.0}      # largest per-group scale-gradient error seen by ds_close in this session (printed by the last test)


def ds_err(got, ref):
    """Per-group relative error of a scale gradient: |got - ref| / max(|ref|, 1e-6 * max|ref|), NaN == NaN."""
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    if not ref.size:
        return 0.0
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6 * scale)
    err = np.where(np.isnan(got) & np.isnan(ref), 0.0, err)
    return float(np.max(err))


def ds_close(got, ref, tol=TOL):
    e = ds_err(got, ref)
    if np.isfinite(e):
        WORST["ds"] = max(WORST["ds"], e)
    return e <= tol


# ------------------------------------------------------------------ golden fixtures
def test_golden_forward_bit_exact_and_ds(lq, golden):
    ops = lq.ops
    n = 0
    for c, P, s, dy, q_ref, y_ref, ds_ref in golden.op_cases():
        Pd, sd, dyd = dev(P), dev(s), dev(dy)
        y, codes = ops.fq_fwd(Pd, sd, codes_dtype=torch.float32)
        assert bits_equal(host(y), y_ref), c
        assert bits_equal(host(codes), q_ref), c
        dX, ds = ops.fq_bwd_threshold(dyd, Pd, sd, c["lam"])
        assert dX.data_ptr() == dyd.data_ptr()
        assert ds_close(host(ds).reshape(ds_ref.shape), ds_ref), (c, host(ds), ds_ref)
        dX2, ds2 = ops.fq_bwd_threshold(dyd, Pd, sd, c["lam"], write_dx=True)
        assert bits_equal(host(dX2), dy) and bits_equal(host(ds2), host(ds)), c
        n += 1
    assert n >= 200


def test_golden_autograd_modules(lq, golden):
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    from learned_quantization_b200.custom_loss_terms.custom_components import custom_layers as cl
    from learned_quantization_b200.paper_implementation.custom_components import custom_layers as pi
    for c, P, s, dy, q_ref, y_ref, ds_ref in list(golden.op_cases())[::7]:
        Pt, st = dev(P).requires_grad_(), dev(s).requires_grad_()
        y = nq.my_custom_gradient(Pt, st, c["lam"])
        y.backward(dev(dy))
        assert bits_equal(host(y), y_ref) and bits_equal(host(Pt.grad), dy), c
        assert st.grad.shape == st.shape and ds_close(host(st.grad), ds_ref), c
    for c in golden.manifest["ste"]:
        g = {k: golden[f"{c['id']}/{k}"] for k in ("P", "s", "dy", "q", "y", "ds")}
        Pt, st = dev(g["P"]).requires_grad_(), dev(g["s"]).requires_grad_()
        y = cl.my_custom_gradient(Pt, st) if c["module"].startswith("CL") else pi.my_custom_gradient(Pt, st, 1e-3)
        y.backward(dev(g["dy"]))
        assert bits_equal(host(y), g["y"]) and bits_equal(host(Pt.grad), g["dy"]), c
        assert bits_equal(host(st.grad), g["ds"]) and not host(st.grad).any(), c


def test_golden_scale_layer_shapes_and_errors(lq, golden):
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    for c in golden.manifest["shapes"]:
        layer = nq.CustomQuantizedScaleLayer(penalty_threshold=0.0, initializer=None, orientation=c["orientation"])
        layer.build(tuple(c["param"]), "cuda")
        assert list(layer.scale.shape) == c["scale"], c
        assert float(layer.scale.detach().flatten()[0]) == 1.1920928955078125e-05
    bad = nq.CustomQuantizedScaleLayer(penalty_threshold=0.0, initializer=None, orientation="diagonal")
    with pytest.raises(ValueError) as e:
        bad.build((3, 3), "cuda")
    assert str(e.value) == golden.manifest["bad_orientation_raises"]
    w = golden["constraint/in"]
    out = nq.MinValueConstraint(golden.manifest["constraint"][0]["min_value"])(dev(w))
    assert bits_equal(host(out), golden["constraint/out"])


def test_golden_loss_terms(lq, golden):
    from learned_quantization_b200 import _lib
    ops = lq.ops
    kinds = {"compute_maxbin_penalty": _lib.LQ_PEN_MAXBIN, "compute_difference_penalty": _lib.LQ_PEN_DIFFERENCE,
             "compute_inverse_penalty": _lib.LQ_PEN_INVERSE}
    for c in golden.manifest["loss"]:
        layers, grads = golden.loss_layers(c), golden.loss_grads(c)
        Ps, ss, gP, gs = [], [], [], []
        for (K, sK, b, sb), (gK, gsK, gb, gsb) in zip(layers, grads):
            Ps += [dev(K), dev(b)]; ss += [dev(sK), dev(sb)]; gP += [gK, gb]; gs += [gsK, gsb]
        kind = kinds[c["method"]]
        out, ws = ops.penalty_fwd(kind, Ps, ss)
        assert rel_err(host(out), golden[f"{c['id']}/penalty"]) < TOL, c
        up = torch.tensor(1.0, device="cuda")
        dXs, dss = ops.penalty_bwd(kind, Ps, ss, ws, 1.0, upstream_dev=up)
        for i in range(len(Ps)):
            if kind != _lib.LQ_PEN_INVERSE:
                scale = max(float(np.max(np.abs(gP[i]))), 1e-30)
                assert float(np.max(np.abs(host(dXs[i]).astype(np.float64) - gP[i]))) <= TOL * scale, (c, i)
            scale = max(float(np.max(np.abs(gs[i]))), 1e-30)
            assert float(np.max(np.abs(host(dss[i]).astype(np.float64).reshape(gs[i].shape) - gs[i]))) <= TOL * scale, (c, i)
        # accumulate flag and host-scalar upstream: g <- 1 + 2.5 * grad
        if kind != _lib.LQ_PEN_INVERSE:
            acc_dX = [torch.ones_like(p) for p in Ps]
            acc_ds = [torch.ones_like(s) for s in ss]
            ops.penalty_bwd(kind, Ps, ss, ws, 2.5, dXs=acc_dX, dss=acc_ds, accumulate=True)
            for i in range(len(Ps)):
                scale = max(float(np.max(np.abs(gP[i]))), 1e-30)
                assert float(np.max(np.abs(host(acc_dX[i]).astype(np.float64) - 1.0 - 2.5 * gP[i]))) <= 1e-6 + 3 * TOL * scale


def test_golden_loss_classes_autograd(lq, golden, tmp_path):
    import types
    from learned_quantization_b200.custom_loss_terms.custom_components import custom_loss_functions as lf
    for c in golden.manifest["loss"]:
        if c.get("no_total"):
            continue
        mnist = c["module"] == "LF_MNIST"
        layers, raw = [], []
        for K, sK, b, sb in golden.loss_layers(c):
            t = [dev(a).requires_grad_() for a in (K, sK, b, sb)]
            ns = types.SimpleNamespace
            if mnist:
                layers.append(ns(W=t[0], nested_q_w_layer=ns(scale=t[1]), b=t[2], nested_q_b_layer=ns(scale=t[3])))
            else:
                layers.append(ns(kernel=t[0], nested_q_k_layer=ns(scale=t[1]), b=t[2], nested_q_b_layer=ns(scale=t[3])))
            raw.append(t)
        obj = getattr(lf, c["class"])(layers=layers, penalty_rate=c["gamma"], log_dir=str(tmp_path))
        pen = getattr(obj, c["method"])()
        assert rel_err(host(pen), golden[f"{c['id']}/penalty"]) < TOL
        pen.backward()
        for t, ref in zip(raw, golden.loss_grads(c)):
            for tt, r in zip(t, ref):
                got = np.zeros_like(r) if tt.grad is None else host(tt.grad)
                assert float(np.max(np.abs(got.astype(np.float64) - r))) <= TOL * max(float(np.max(np.abs(r))), 1e-30)
        tot = obj.compute_total_loss(dev(golden[f"{c['id']}/y_true"]), dev(golden[f"{c['id']}/y_pred"]))
        assert rel_err(host(tot), golden[f"{c['id']}/total_loss"]) < TOL
        obj.flush_logs()
        assert (tmp_path / "custom_losses" / "total_loss.log").read_text().strip() != ""


def test_golden_host_layers(lq, golden):
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    from learned_quantization_b200.custom_loss_terms.custom_components import custom_layers as cl
    from learned_quantization_b200.paper_implementation.custom_components import custom_layers as pi
    mods = {"NQ": nq, "CL": cl, "PI": pi}
    ops_ = lq.ops
    # the fixtures are fp32 CPU convolutions; cuDNN would otherwise pick TF32 (conv is not our path)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for c in golden.manifest["layers"]:
        cid = c["id"]
        ctor = dict(c["ctor"])
        for k in ("kernel_size", "strides"):
            if k in ctor:
                ctor[k] = tuple(ctor[k])
        weights = [golden[f"{cid}/w0"]] + ([golden[f"{cid}/w1"]] if f"{cid}/w1" in golden.arrays else [])
        layer = getattr(mods[c["module"]], c["class"])(**ctor, initializer=None, regularizer=None, trained_weights=weights)
        x = dev(golden[f"{cid}/x"]).requires_grad_()
        layer.build(tuple(x.shape), "cuda")
        qn = [n for n in ("nested_q_w_layer", "nested_q_k_layer", "nested_q_b_layer") if hasattr(layer, n)]
        with torch.no_grad():
            for i, n in enumerate(qn):
                getattr(layer, n).scale.copy_(dev(golden[f"{cid}/s{i}"]))
        out = layer(x)
        ref = golden[f"{cid}/out"]
        assert out.shape == ref.shape, c
        assert float(np.max(np.abs(host(out) - ref))) <= 2e-5 * max(float(np.max(np.abs(ref))), 1.0), c
        (out * dev(golden[f"{cid}/G"])).sum().backward()
        params = [layer.W if hasattr(layer, "W") else layer.kernel] + ([layer.b] if hasattr(layer, "b") else [])
        for i, p in enumerate(params):
            r = golden[f"{cid}/dw{i}"]
            assert float(np.max(np.abs(host(p.grad) - r))) <= 1e-4 * max(float(np.max(np.abs(r))), 1e-30), (c, i)
        r = golden[f"{cid}/dx"]
        assert float(np.max(np.abs(host(x.grad) - r))) <= 1e-4 * max(float(np.max(np.abs(r))), 1e-30), c
        for i, n in enumerate(qn):
            r = golden[f"{cid}/ds{i}"]
            got = host(getattr(layer, n).scale.grad)
            assert got.shape == r.shape
            # end to end ds depends on dy only through thresholded ratios, and cuDNN's dy differs from the fixture's
            # in the last bits: a loose bound here ...
            assert float(np.max(np.abs(got - r))) <= 2e-3 * max(float(np.max(np.abs(r))), 1e-30), (c, n, got, r)
            # ... and the operator itself held to 1e-5 per group on the FIXTURE'S OWN dy (the gradient the reference's
            # layer handed to its custom_grad is the parameter gradient it returned: dP = dy, NQ:118)
            if c["module"] == "NQ":
                q = getattr(layer, n)
                _, ds_direct = ops_.fq_bwd_threshold(dev(golden[f"{cid}/dw{i}"]), params[i].detach(), q.scale.detach(),
                                                     float(q.penalty_threshold))
                assert ds_close(host(ds_direct).reshape(r.shape), r), (c, n, host(ds_direct), r)


# ------------------------------------------------------------------ randomized vs oracle
LAYOUTS = [
    # (shape, scale shape) -> exercises ROW/COL regimes, vector/scalar paths, segments, tails
    ((1, 1, 5000), (1,)), ((3, 3, 64, 128), (1, 1, 64, 1)), ((3, 3, 16, 32), (1, 1, 16, 1)),
    ((7, 7, 3, 64), (7, 1, 1, 1)), ((7, 7, 3, 64), (1, 7, 1, 1)), ((784, 128), (1, 128)), ((784, 128), (784, 1)),
    ((128, 10), (1, 10)), ((128, 10), (128, 1)), ((37, 5, 131), (1, 5, 1)), ((9, 6, 4100), (1, 6, 1)),
    ((2, 3, 10001), (1, 3, 1)), ((1000, 3), (1, 3)), ((513, 7, 2), (1, 7, 1)), ((5,), (1,)), ((1,), (1,)),
    ((64, 1024), (1, 1024)), ((300, 260, 4), (1, 260, 1)), ((16, 2048, 128), (1, 2048, 1)),
    # TMA-staged COL backward (outer >= 512, W % 4 == 0, >= 1 Mi elements): full / partial column tiles,
    # 32-, 64- and 128-column tiles, row tails, inner > 1
    ((4100, 260), (1, 260)), ((2100, 64, 8), (1, 64, 1)), ((27000, 10, 4), (1, 10, 1)), ((1030, 1024), (1, 1024)),
    ((16500, 64), (1, 64)), ((9000, 128), (1, 128)),
]


@pytest.mark.parametrize("shape,s_shape", LAYOUTS)
@pytest.mark.parametrize("lam", [0.0, 1e-10, 1e-3, 0.05])
def test_random_vs_oracle(lq, shape, s_shape, lam):
    ops = lq.ops
    rng = np.random.default_rng(abs(hash((shape, s_shape))) % (2 ** 31))
    P = rng.normal(0, 0.05, size=shape).astype(np.float32)
    P.reshape(-1)[:: 17] = 0.0
    s = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
    dy = rng.normal(0, 1e-3, size=shape).astype(np.float32)
    dy.reshape(-1)[:: 13] = 0.0
    q_ref, y_ref = o.fq_forward(P, s)
    _, ds_ref = o.nested_backward(dy, q_ref, y_ref, s, lam, accum="fp64")
    Pd, sd, dyd = dev(P), dev(s), dev(dy)
    y, codes = ops.fq_fwd(Pd, sd, codes_dtype=torch.int32)
    assert bits_equal(host(y), y_ref)
    assert np.array_equal(host(codes), q_ref.astype(np.int32))
    _, ds = ops.fq_bwd_threshold(dyd, Pd, sd, lam)
    assert ds_close(host(ds).reshape(ds_ref.shape), ds_ref), (host(ds), ds_ref)
    _, ds_b = ops.fq_bwd_threshold(dyd, Pd, sd, lam)
    assert bits_equal(host(ds), host(ds_b)), "second-stage reduction must be deterministic"


def _fuzz_shapes():
    rng = np.random.default_rng(20261018)
    shapes = []
    for _ in range(36):                                   # small: every traversal, tails, odd sizes
        outer = int(rng.integers(1, 700))
        C = int(rng.choice([1, 2, 3, 5, 8, 16, 33, 64, 130]))
        inner = int(rng.choice([1, 2, 3, 4, 7, 16, 63, 64, 127, 128, 129, 512, 1000, 4097]))
        if outer * C * inner <= 300_000:
            shapes.append((outer, C, inner))
    for _ in range(8):                                    # TMA-staged COL backward: >= 1 Mi elements, W % 4 == 0
        W = int(rng.choice([32, 36, 64, 100, 128, 132, 260, 1028]))
        inner = int(rng.choice([d for d in (1, 2, 4) if W % d == 0]))
        outer = int(rng.integers((1 << 20) // W + 1, (3 << 19) // W + 600))
        shapes.append((outer, W // inner, inner))
    return shapes


@pytest.mark.parametrize("layout", _fuzz_shapes())
def test_fuzzed_layouts_vs_oracle(lq, layout):
    ops = lq.ops
    outer, C, inner = layout
    rng = np.random.default_rng(outer * 1_000_003 + C * 1009 + inner)
    P = rng.normal(0, 0.05, size=layout).astype(np.float32)
    P.reshape(-1)[:: 11] = 0.0
    dy = rng.normal(0, 1e-3, size=layout).astype(np.float32)
    s = rng.uniform(5e-3, 2e-2, size=(1, C, 1)).astype(np.float32)
    lam = float(rng.choice([1e-3, 1e-10, 0.05]))
    q_ref, y_ref = o.fq_forward(P, s)
    _, ds_ref = o.nested_backward(dy, q_ref, y_ref, s, lam, accum="fp64")
    Pd, sd, dyd = dev(P), dev(s), dev(dy)
    y = ops.fq_fwd(Pd, sd, layout)
    assert bits_equal(host(y), y_ref)
    dX, ds = ops.fq_bwd_threshold(dyd, Pd, sd, lam, layout, write_dx=True)
    assert bits_equal(host(dX), dy)
    assert ds_close(host(ds).reshape(-1), ds_ref.reshape(-1)), (layout, lam)


def divcheck(lq, s_bits, a_sign_exp_bits):
    """(mismatches, pairs checked, first bad pair) of the kernels' hoisted-reciprocal division
    vs __fdiv_rn over all 2^23 dividend mantissas for every divisor in s_bits."""
    from learned_quantization_b200._lib import lib, check
    sb = torch.as_tensor(np.asarray(s_bits, np.uint32).view(np.int32)).cuda()
    out = torch.zeros(3, dtype=torch.int64, device="cuda")
    check(lib.lq_debug_divcheck(sb.data_ptr(), sb.numel(), int(a_sign_exp_bits), out.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
    return [int(v) for v in out.cpu()]


def test_hoisted_reciprocal_division_is_ieee_division(lq):
    """The forward is graded bit-exact against tf.truediv (NQ:56-58): the per-group reciprocal
    fast path must give the IEEE quotient for every dividend mantissa.  4096 divisors (random
    mantissas and exponents inside the fast window, plus the structured mantissas that are
    hard for reciprocal-based division) x 2^23 dividend mantissas x 4 dividend exponents/signs."""
    rng = np.random.default_rng(123)
    mant = rng.integers(0, 1 << 23, size=4096, dtype=np.uint32)
    special = np.array([0, 1, 2, 3, 0x7fffff, 0x7ffffe, 0x400000, 0x400001, 0x3fffff, 0x555555, 0x2aaaaa,
                        0x7ffff0, 0x000fff, 0x600000, 0x200000, 0x100000], np.uint32)
    mant[:special.size] = special
    exp = rng.integers(127 - 40, 127 + 40, size=4096, dtype=np.uint32)       # |s| in [2^-40, 2^40)
    exp[:special.size] = 127
    sign = (rng.integers(0, 2, size=4096, dtype=np.uint32) << 31)
    sign[:special.size] = 0
    s_bits = sign | (exp << 23) | mant
    total = 0
    for a_hi in (127 << 23, 126 << 23, (1 << 31) | (130 << 23), (127 - 75) << 23):
        bad, seen, first = divcheck(lq, s_bits, a_hi)
        assert bad == 0, f"{bad} mismatches, first a|s = {first:#018x}"
        total += seen
    assert total == 4 * 4096 * (1 << 23)
    # outside the window nothing takes the fast path (and so nothing is counted)
    bad, seen, _ = divcheck(lq, np.array([0x00000001, 0x7f000000, 0, 0x7f800000], np.uint32), 127 << 23)
    assert (bad, seen) == (0, 0)
    bad, seen, _ = divcheck(lq, s_bits[:16], 0)            # denormal / zero dividends
    assert (bad, seen) == (0, 0)


@pytest.mark.parametrize("shape,s_shape,nshards", [((8, 16, 256), (1, 16, 1), 4), ((6, 1, 9000), (1, 1, 1), 3),
                                                   ((640, 12, 1), (1, 12, 1), 5), ((8, 16, 256), (1, 16, 1), 1)])
def test_sharded_backward_equals_unsharded(lq, shape, s_shape, nshards):
    """a tensor split along `outer`; per-shard statistics + one gather + finish == the
    single-device backward on the whole tensor (and the oracle)."""
    ops = lq.ops
    rng = np.random.default_rng(21)
    P = rng.normal(0, 0.05, shape).astype(np.float32)
    dy = rng.normal(0, 1e-3, shape).astype(np.float32)
    s = rng.uniform(5e-3, 2e-2, s_shape).astype(np.float32)
    for lam in (1e-3, 0.0, 1e3):
        q_ref, y_ref = o.fq_forward(P, s)
        _, ds_ref = o.nested_backward(dy, q_ref, y_ref, s, lam, accum="fp64")
        _, ds_whole = ops.fq_bwd_threshold(dev(dy), dev(P), dev(s), lam)
        cuts = np.linspace(0, shape[0], nshards + 1).astype(int)
        stats = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            Pa, da = dev(P[a:b]), dev(dy[a:b])
            dX, st = ops.fq_bwd_threshold_stats(da, Pa, dev(s), lam, (int(b - a), shape[1], shape[2]), write_dx=True)
            assert bits_equal(host(dX), dy[a:b])
            stats.append(st)
        n_per_group = shape[0] * shape[2]
        ds = ops.fq_bwd_threshold_finish(torch.stack(stats).contiguous(), lam, n_per_group)
        assert ds_close(host(ds), host(ds_whole).reshape(-1), tol=1e-6)
        assert ds_close(host(ds), ds_ref.reshape(-1))


def test_p2p_finish_world1_equals_plain_finish(lq):
    """lq_fq_bwd_threshold_finish_p2p with a single rank: same kernel (push, flag, wait, fold), its own
    memory as the only peer; must equal lq_fq_bwd_threshold_finish, epoch after epoch."""
    from learned_quantization_b200.training.dp import P2PFinish
    ops = lq.ops
    rng = np.random.default_rng(33)
    shape = (64, 300, 16)
    P = rng.normal(0, 0.05, shape).astype(np.float32)
    s = rng.uniform(5e-3, 2e-2, (1, 300, 1)).astype(np.float32)
    p2p = P2PFinish(max_groups=512, rank=0, world=1)
    try:
        for it in range(5):
            dy = rng.normal(0, 1e-3, shape).astype(np.float32)
            _, st = ops.fq_bwd_threshold_stats(dev(dy), dev(P), dev(s), 1e-3, (64, 300, 16))
            want = ops.fq_bwd_threshold_finish(st.unsqueeze(0).contiguous(), 1e-3, 64 * 16)
            got = p2p.finish(st, 1e-3, 64 * 16, torch.empty(300, device="cuda"))
            assert bits_equal(host(got), host(want)), it
        assert not p2p.timed_out()
    finally:
        p2p.close()


def test_unaligned_pointers_take_scalar_path(lq):
    ops = lq.ops
    rng = np.random.default_rng(5)
    for shape, s_shape in [((6, 8, 256), (1, 8, 1)), ((100, 12), (1, 12)), ((1, 1, 9000), (1,))]:
        n = int(np.prod(shape))
        P = rng.normal(0, 0.05, size=n + 1).astype(np.float32)
        dy = rng.normal(0, 1e-3, size=n + 1).astype(np.float32)
        s = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
        Pd, dyd = dev(P)[1:].view(shape), dev(dy)[1:].view(shape)     # 4-byte aligned only
        assert Pd.data_ptr() % 16 != 0
        q_ref, y_ref = o.fq_forward(P[1:].reshape(shape), s)
        _, ds_ref = o.nested_backward(dy[1:].reshape(shape), q_ref, y_ref, s, 1e-3, accum="fp64")
        y = ops.fq_fwd(Pd, dev(s), out=torch.empty(n + 1, device="cuda")[1:].view(shape))
        assert bits_equal(host(y), y_ref)
        _, ds = ops.fq_bwd_threshold(dyd, Pd, dev(s), 1e-3)
        assert ds_close(host(ds).reshape(ds_ref.shape), ds_ref)


def test_special_values(lq):
    ops = lq.ops
    P = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-38, -1e-38, 3e38, 0.3, -0.26, 1e-45, -1e-45], np.float32)
    P = np.tile(P, 20)
    for sv in (0.1, 1.1920929e-05, 3.0):
        s = np.array([sv], np.float32)
        q_ref, y_ref = o.fq_forward(P, s)
        y, codes = ops.fq_fwd(dev(P), dev(s), codes_dtype=torch.float32)
        assert bits_equal(host(y), y_ref), sv
        assert bits_equal(host(codes), q_ref), sv
    # NaN ratio -> counts as "below" and poisons ds
    Pn = np.array([0.3, 0.2, -0.1, 0.25] * 40, np.float32)
    dy = np.zeros_like(Pn)
    dy[5] = np.nan
    s = np.array([0.1], np.float32)
    q_ref, y_ref = o.fq_forward(Pn, s)
    _, ds_ref = o.nested_backward(dy, q_ref, y_ref, s, 1e-3)
    _, ds = ops.fq_bwd_threshold(dev(dy), dev(Pn), dev(s), 1e-3)
    assert np.isnan(ds_ref).all() and np.isnan(host(ds)).all()


@pytest.mark.parametrize("lam", [0.0, -0.0, 1e-11, 1e-3])
@pytest.mark.parametrize("shape,s_shape", [((4, 6, 256), (1, 6, 1)), ((1, 1, 4100), (1,)), ((2048, 40, 1), (1, 40, 1)),
                                           ((16384, 64), (1, 64))])
def test_special_gradients_at_every_lambda(lq, shape, s_shape, lam):
    """lambda = 0 takes the filtered fast path (|dy| >= |nz| * 0): infinities, NaNs, zeros and denormals in dy,
    zeros / infinities in y, in every traversal (ROW, register COL, TMA-staged COL) must come out as the oracle's."""
    ops = lq.ops
    rng = np.random.default_rng(77)
    s = rng.uniform(5e-3, 2e-2, s_shape).astype(np.float32)
    special_dy = np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-40, 3e38], np.float32)
    for special_P in (np.array([0.0, -0.0, 1e-45, 1e-3], np.float32),                     # finite codes
                      np.array([0.0, 1e-3, np.inf, -np.inf, 3e38], np.float32)):          # infinite codes too
        P = rng.normal(0, 0.05, shape).astype(np.float32)
        dy = rng.normal(0, 1e-3, shape).astype(np.float32)
        flatP, flatd = P.reshape(-1), dy.reshape(-1)
        n = flatP.size
        idx = rng.choice(n, size=special_dy.size * special_P.size, replace=False)
        k = 0
        for a in special_dy:
            for b in special_P:
                flatd[idx[k]], flatP[idx[k]] = a, b
                k += 1
        for poison in (False, True):
            if poison:
                flatd[int(rng.integers(n))] = np.nan        # one NaN ratio: its group's ds becomes NaN, the others stay
            with np.errstate(all="ignore"):
                q_ref, y_ref = o.fq_forward(P, s)
                _, ds_ref = o.nested_backward(dy, q_ref, y_ref, s, lam, accum="fp64")
            dX, ds = ops.fq_bwd_threshold(dev(dy), dev(P), dev(s), lam, write_dx=True)
            assert bits_equal(host(dX), dy)
            got, want = host(ds).reshape(-1), ds_ref.reshape(-1)
            assert np.array_equal(np.isnan(got), np.isnan(want)), (lam, poison)
            assert np.array_equal(np.isinf(got), np.isinf(want)) and np.array_equal(np.sign(got[np.isinf(got)]),
                                                                                    np.sign(want[np.isinf(want)]))
            ok = np.isfinite(want)
            assert ds_close(got[ok], want[ok]), (lam, poison)


def test_rne_clamp_extension(lq):
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(3)
    P = rng.normal(0, 0.05, size=(4, 16, 256)).astype(np.float32)
    s = rng.uniform(5e-3, 2e-2, size=(1, 16, 1)).astype(np.float32)
    q_ref, y_ref = o.fq_forward(P, s, round_mode=o.ROUND_RNE, qmin=-8, qmax=7)
    y, codes = ops.fq_fwd(dev(P), dev(s), codes_dtype=torch.int8, round_mode=_lib.LQ_ROUND_RNE, qmin=-8.0, qmax=7.0)
    assert bits_equal(host(y), y_ref)
    assert np.array_equal(host(codes), q_ref.astype(np.int8))
    assert host(codes).min() == -8 and host(codes).max() == 7
    # the extension's backward: ordinary clamp bounds (filtered path), and a clamp that lifts every code to 1e30 so
    # that q*s overflows for the large scales below (exact path; inf / inf ratios poison those groups in the oracle too)
    dy = rng.normal(0, 1e-3, size=P.shape).astype(np.float32)
    dy.reshape(-1)[::97] = np.inf
    s_big = s.copy()
    s_big[0, ::2, 0] = 1e12
    for sc, lo, hi in ((s, -8.0, 7.0), (s, -3e38, 3e38), (s_big, 1e30, float("inf"))):
        with np.errstate(all="ignore"):
            q_ref, y_ref = o.fq_forward(P, sc, round_mode=o.ROUND_RNE, qmin=lo, qmax=hi)
            _, ds_ref = o.nested_backward(dy, q_ref, y_ref, sc, 1e-3, accum="fp64")
        _, ds = ops.fq_bwd_threshold(dev(dy), dev(P), dev(sc), 1e-3, round_mode=_lib.LQ_ROUND_RNE, qmin=lo, qmax=hi)
        got, want = host(ds).reshape(-1), ds_ref.reshape(-1)
        assert np.array_equal(np.isnan(got), np.isnan(want)), (lo, hi)
        assert np.array_equal(np.isinf(got), np.isinf(want)), (lo, hi)
        ok = np.isfinite(want)
        assert ds_close(got[ok], want[ok]), (lo, hi)


def test_int8_codes_wrap_like_reference_export(lq):
    ops = lq.ops
    rng = np.random.default_rng(4)
    P = rng.normal(0, 0.05, size=(3, 3, 8, 16)).astype(np.float32)
    s = np.full((1, 1, 8, 1), o.SCALE_INIT, np.float32)          # init-time codes reach +-2e4
    _, codes = ops.fq_fwd(dev(P), dev(s), codes_dtype=torch.int8)
    assert np.array_equal(host(codes), o.codes_int8(P, s))


@pytest.mark.parametrize("shape,s_shape", LAYOUTS[:10] + [((1 << 22,), (1,)), ((4096, 512), (1, 512))])
def test_forward_statistics_are_by_products_of_the_same_pass(lq, shape, s_shape):
    """lq_fq_fwd_stats: y bit-exact as ever; per group min P / max P exact, sum |P - P/s| within 1e-5 of
    the oracle's terms summed in fp64; and the quantities the callbacks / loss terms derive from them."""
    ops = lq.ops
    rng = np.random.default_rng(31)
    P = rng.normal(0, 0.05, size=shape).astype(np.float32)
    s = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
    y, st = ops.fq_fwd_stats(dev(P), dev(s))
    q_ref, y_ref = o.fq_forward(P, s)
    assert bits_equal(host(y), y_ref)
    st = host(st)
    outer, C, inner = o.group_layout(shape, s_shape)
    Pg = P.reshape(outer, C, inner)
    sg = np.broadcast_to(s, shape).reshape(outer, C, inner) if len(s_shape) == len(shape) else np.full((outer, C, inner), s.reshape(-1)[0], np.float32)
    assert np.array_equal(st[:, 0], Pg.min(axis=(0, 2))) and np.array_equal(st[:, 1], Pg.max(axis=(0, 2)))
    terms = np.abs(Pg - Pg / sg).astype(np.float32)
    assert rel_err(st[:, 2], terms.astype(np.float64).sum(axis=(0, 2))) < TOL
    # max |q| of custom_callbacks.py:84-129 and MaxBin's per-group term (LF:90-94) from the two extremes
    s1 = sg[0, :, 0]
    max_abs_q = np.maximum(np.abs(np.floor(st[:, 0] / s1)), np.abs(np.floor(st[:, 1] / s1)))
    assert np.array_equal(max_abs_q, np.abs(q_ref.reshape(outer, C, inner)).max(axis=(0, 2)))
    assert np.array_equal(np.maximum(np.abs(st[:, 0]), np.abs(st[:, 1])) / s1, (np.abs(Pg) / sg).max(axis=(0, 2)))


def test_multi_tensor_equals_single(lq):
    ops = lq.ops
    rng = np.random.default_rng(11)
    shapes = [((3, 3, 16, 32), (1, 1, 16, 1)), ((32,), (1,)), ((7, 7, 3, 64), (1, 1, 3, 1)), ((64,), (1,)),
              ((3, 3, 64, 128), (1, 1, 64, 1)), ((128,), (1,)), ((784, 128), (784, 1)), ((1, 1, 64, 128), (1, 1, 64, 1))]
    shapes = shapes * 8          # 64 tensors: more than one kernel-parameter table
    shapes.append(((3, 3, 1024, 512), (1,)))      # one group with 4608 partials: the two-level second stage of lq_mt_fq_bwd
    Ps = [dev(rng.normal(0, 0.05, size=sh).astype(np.float32)) for sh, _ in shapes]
    ss = [dev(rng.uniform(5e-3, 2e-2, size=ssh).astype(np.float32)) for _, ssh in shapes]
    dys = [dev(rng.normal(0, 1e-3, size=sh).astype(np.float32)) for sh, _ in shapes]
    ys = [torch.empty_like(p) for p in Ps]
    dss = [torch.empty_like(s) for s in ss]
    ops.mt_fq_fwd(Ps, ss, ys)
    ops.mt_fq_bwd(dys, Ps, ss, dss, 1e-3)
    for p, s, dy, y, ds in zip(Ps, ss, dys, ys, dss):
        y1 = ops.fq_fwd(p, s)
        _, ds1 = ops.fq_bwd_threshold(dy, p, s, 1e-3)
        # y: same bits.  ds: the multi-tensor launch cuts interleaved tensors into longer row chunks than a launch of the
        # tensor by itself would (a share of ONE grid instead of a grid each), so the partials are added in another order
        assert bits_equal(host(y), host(y1)) and ds_close(host(ds).reshape(-1), host(ds1).reshape(-1))
    dss_b = [torch.empty_like(s) for s in ss]
    ops.mt_fq_bwd(dys, Ps, ss, dss_b, 1e-3)
    assert all(bits_equal(host(a), host(b)) for a, b in zip(dss, dss_b)), "the multi-tensor backward must be deterministic"
    dss0 = [torch.ones_like(s) for s in ss]
    from learned_quantization_b200 import _lib
    ops.mt_fq_bwd(dys, Ps, ss, dss0, 1e-3, scale_grad_mode=_lib.LQ_SGRAD_NONE)
    assert all(not host(d).any() for d in dss0)


def test_penalties_random_vs_oracle(lq):
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(21)
    specs = [((3, 3, 64, 128), (1, 1, 64, 1)), ((7, 7, 3, 64), (7, 1, 1, 1)), ((784, 128), (1, 128)),
             ((3, 3, 128, 256), (1,)), ((37, 5, 131), (1, 5, 1))]
    layers = []
    for sh, ssh in specs:
        K = rng.normal(0, 0.05, size=sh).astype(np.float32)
        sK = rng.uniform(5e-3, 2e-2, size=ssh).astype(np.float32)
        b = rng.normal(0, 0.05, size=(sh[-1],)).astype(np.float32)
        sb = rng.uniform(5e-3, 2e-2, size=(1,)).astype(np.float32)
        layers.append((K, sK, b, sb))
    Ps, ss = [], []
    for K, sK, b, sb in layers:
        Ps += [dev(K), dev(b)]; ss += [dev(sK), dev(sb)]
    for kind, val, grad in ((_lib.LQ_PEN_MAXBIN, o.maxbin_penalty, o.maxbin_penalty_grads),
                            (_lib.LQ_PEN_DIFFERENCE, o.difference_penalty, o.difference_penalty_grads),
                            (_lib.LQ_PEN_INVERSE, o.inverse_penalty, o.inverse_penalty_grads)):
        out, ws = ops.penalty_fwd(kind, Ps, ss)
        assert rel_err(host(out), val(layers)) < TOL, kind
        dXs, dss = ops.penalty_bwd(kind, Ps, ss, ws, 0.5)
        ref = grad(layers, 0.5, accum="fp64")
        flat_ref = []
        for gK, gsK, gb, gsb in ref:
            flat_ref += [(gK, gsK), (gb, gsb)]
        for i, (gP, gs) in enumerate(flat_ref):
            if kind != _lib.LQ_PEN_INVERSE:
                assert float(np.max(np.abs(host(dXs[i]).astype(np.float64) - gP))) <= TOL * max(float(np.max(np.abs(gP))), 1e-30), (kind, i)
            assert float(np.max(np.abs(host(dss[i]).astype(np.float64).reshape(gs.shape) - gs))) <= TOL * max(float(np.max(np.abs(gs))), 1e-30), (kind, i)


@pytest.mark.parametrize("shape,s_shape", [((8, 4, 512), (1, 4, 1)), ((1, 1, 6000), (1,)), ((1000, 64), (1, 64)),
                                           ((520, 36), (1, 36)), ((300, 7), (1, 7))])
def test_maxbin_ties_inside_and_across_quads(lq, shape, s_shape):
    """reduce_max's gradient splits evenly among ties (LF:90-94 under TF autodiff).  The kernels count ties four
    elements at a time: put equal maxima next to each other (one ROW quad / one column's consecutive rows), far
    apart, and with opposite signs, in some groups only."""
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(5)
    K = rng.normal(0, 0.05, size=shape).astype(np.float32)
    flat = K.reshape(-1)
    C = s_shape[1] if len(s_shape) > 1 else 1
    inner = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    row = C * inner
    top = np.float32(0.75)
    if C == 1:
        flat[[0, 1, 2, 3, 4, 2048, 5999]] = [top, -top, top, top, top, -top, top]
    else:
        for g in range(0, C, 2):                                   # every other group gets ties
            base = g * inner
            if inner > 1:
                flat[base:base + 3] = [top, -top, top]             # one quad of the group's first row
                flat[base + row * (shape[0] - 1) + inner - 1] = top    # last element of its last row
            else:
                for r in (0, 1, 2, 3, 4, shape[0] // 2, shape[0] - 1):  # consecutive rows of one column, then far ones
                    flat[r * row + base] = top if r % 2 == 0 else -top
    sK = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
    b = np.array([0.3, -0.3, 0.3, 0.1], np.float32)
    sb = np.array([0.01], np.float32)
    layers = [(K, sK, b, sb)]
    Ps, ss = [dev(K), dev(b)], [dev(sK), dev(sb)]
    out, ws = ops.penalty_fwd(_lib.LQ_PEN_MAXBIN, Ps, ss)
    assert rel_err(host(out), o.maxbin_penalty(layers)) < TOL
    dXs, dss = ops.penalty_bwd(_lib.LQ_PEN_MAXBIN, Ps, ss, ws, 1.0)
    (gK, gsK, gb, gsb), = o.maxbin_penalty_grads(layers, 1.0, accum="fp64")
    for got, want in ((dXs[0], gK), (dXs[1], gb), (dss[0], gsK), (dss[1], gsb)):
        got = host(got).astype(np.float64).reshape(want.shape)
        assert np.array_equal(got != 0, want != 0)                 # exactly the tied elements carry gradient
        assert float(np.max(np.abs(got - want))) <= TOL * max(float(np.max(np.abs(want))), 1e-30)


@pytest.mark.parametrize("shape,s_shape", [((64, 8, 256), (1, 8, 1)), ((4096, 24), (1, 24)), ((1 << 21,), (1,)),
                                           ((40, 3, 2000), (1, 3, 1))])
@pytest.mark.parametrize("fill", ["constant", "zeros", "one_ulp_below", "tiny", "mixed_scales"])
def test_maxbin_tie_counting_corner_cases(lq, shape, s_shape, fill):
    """The MaxBin passes work on |P| and divide once per group; ties in the QUOTIENT are what TF's reduce_max
    gradient splits over (LF:90-94).  Cases the |P|-domain shortcut must get right: every element tied (more
    candidate units than the second stage's list holds), all-zero groups, elements one ulp below the maximum
    whose quotient still rounds to it, values below the fast window, and scales outside the fast set."""
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(17)
    K = rng.normal(0, 0.05, size=shape).astype(np.float32)
    sK = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
    if fill == "constant":
        K[...] = np.float32(0.25) * np.where(rng.random(shape) < 0.5, -1, 1).astype(np.float32)
    elif fill == "zeros":
        K[...] = 0.0
        K.reshape(-1)[::7] = -0.0
    elif fill == "one_ulp_below":
        top = np.float32(0.8125)
        flat = K.reshape(-1)
        idx = rng.choice(flat.size, size=min(flat.size // 3, 5000), replace=False)
        flat[idx[0::3]] = top
        flat[idx[1::3]] = np.nextafter(top, np.float32(0))
        flat[idx[2::3]] = -np.nextafter(np.nextafter(top, np.float32(0)), np.float32(0))
    elif fill == "tiny":
        K *= np.float32(1e-30)
    else:
        sK.reshape(-1)[0] = np.float32(1e-14)          # below the fast window: exact regime for that group
        if sK.size > 1:
            sK.reshape(-1)[1] = np.float32(3e13)
    b = np.array([0.3, -0.3, 0.3, 0.1], np.float32)
    sb = np.array([0.01], np.float32)
    layers = [(K, sK, b, sb)]
    Ps, ss = [dev(K), dev(b)], [dev(sK), dev(sb)]
    out, ws = ops.penalty_fwd(_lib.LQ_PEN_MAXBIN, Ps, ss)
    assert rel_err(host(out), o.maxbin_penalty(layers)) < TOL
    dXs, dss = ops.penalty_bwd(_lib.LQ_PEN_MAXBIN, Ps, ss, ws, 1.0)
    (gK, gsK, gb, gsb), = o.maxbin_penalty_grads(layers, 1.0, accum="fp64")
    for got, want in ((dXs[0], gK), (dXs[1], gb), (dss[0], gsK), (dss[1], gsb)):
        got = host(got).astype(np.float64).reshape(want.shape)
        assert np.array_equal(got != 0, want != 0)                 # exactly the tied elements carry gradient
        assert float(np.max(np.abs(got - want))) <= TOL * max(float(np.max(np.abs(want))), 1e-30)
    # Difference on the same tensors (fast and exact regimes side by side for "mixed_scales")
    out, ws = ops.penalty_fwd(_lib.LQ_PEN_DIFFERENCE, Ps, ss)
    assert rel_err(host(out), o.difference_penalty(layers)) < TOL
    dXs, dss = ops.penalty_bwd(_lib.LQ_PEN_DIFFERENCE, Ps, ss, ws, 1.0)
    (gK, gsK, gb, gsb), = o.difference_penalty_grads(layers, 1.0, accum="fp64")
    for got, want in ((dXs[0], gK), (dXs[1], gb), (dss[0], gsK), (dss[1], gsb)):
        got = host(got).astype(np.float64).reshape(want.shape)
        assert float(np.max(np.abs(got - want))) <= TOL * max(float(np.max(np.abs(want))), 1e-30)


@pytest.mark.parametrize("shape", [(20000, 128), (30011, 132)])
def test_penalties_interleaved_lean_kernels(lq, shape):
    """Interleaved tensors of this size take the lean COL kernels of lq_penalty.cu (stat_col_kernel forward,
    pen_bwd_col_kernel backward: rows per thread >= 8); the same data at an address that is not 16-byte aligned takes
    the generic traversal.  dX must have the same bits either way, penalty and ds must agree within TOL, and both must
    match the oracle -- with quads that are not ordinary (values beyond 2^80, re-done out of line), whole rows of zeros,
    a chunk tail (30011 rows), a nearly empty second column block (132 columns) whose scale is outside Difference's
    fast set (that block falls back to the generic traversal), plain and accumulate mode."""
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(shape[0])
    K = rng.normal(0, 0.05, size=shape).astype(np.float32)
    K[17, :] = 0.0; K[18, 5] = -0.0
    sK = rng.uniform(5e-3, 2e-2, size=(1, shape[1])).astype(np.float32)
    if shape[1] > 128:                                                  # (the first shape stays ordinary: a strict ds check)
        K[5, 3] = 1e25; K[100, 7] = -3e24; K[shape[0] - 2, 9] = 2e24   # > 2^80 = 1.2e24: their quads are not ordinary
        sK[0, 130] = 1.0                                                # not in Difference's fast set (1/2 < s < 2)
    b = rng.normal(0, 0.05, size=(shape[1],)).astype(np.float32)
    sb = np.array([0.01], np.float32)
    layers = [(K, sK, b, sb)]
    n = K.size
    buf = torch.empty(n + 1, device="cuda")
    buf[1:].copy_(dev(K).reshape(-1))
    K_unaligned = buf[1:].view(shape)
    assert K_unaligned.data_ptr() % 16 != 0
    for kind, val, grad in ((_lib.LQ_PEN_MAXBIN, o.maxbin_penalty, o.maxbin_penalty_grads),
                            (_lib.LQ_PEN_DIFFERENCE, o.difference_penalty, o.difference_penalty_grads)):
        ref_val = val(layers)
        (gK, gsK, gb, gsb), = grad(layers, 1.0, accum="fp64")
        results = []
        for Kt in (dev(K), K_unaligned):
            Ps, ss = [Kt, dev(b)], [dev(sK), dev(sb)]
            out, ws = ops.penalty_fwd(kind, Ps, ss)
            assert rel_err(host(out), ref_val) < TOL, kind
            dX0 = torch.empty(n + 1, device="cuda")[1:].view(shape) if Kt is K_unaligned else torch.empty_like(Kt)
            dXs, dss = ops.penalty_bwd(kind, Ps, ss, ws, 1.0, dXs=[dX0, torch.empty_like(Ps[1])],
                                       dss=[torch.empty_like(ss[0]), torch.empty_like(ss[1])])
            scale = max(float(np.max(np.abs(gK))), 1e-30)
            assert float(np.max(np.abs(host(dXs[0]).astype(np.float64) - gK))) <= TOL * scale, kind
            assert ds_err(host(dss[0]), gsK) <= TOL, kind
            acc_dX = [torch.ones(n + 1, device="cuda")[1:].view(shape) if Kt is K_unaligned else torch.ones_like(Kt),
                      torch.ones_like(Ps[1])]
            acc_ds = [torch.ones_like(s) for s in ss]
            ops.penalty_bwd(kind, Ps, ss, ws, 2.5, dXs=acc_dX, dss=acc_ds, accumulate=True)
            assert float(np.max(np.abs(host(acc_dX[0]).astype(np.float64) - 1.0 - 2.5 * gK))) <= 1e-6 + 3 * TOL * scale, kind
            results.append((host(out), host(dXs[0]).copy(), host(dss[0]).copy(), host(acc_dX[0]).copy()))
        (o1, d1, s1, a1), (o2, d2, s2, a2) = results
        assert bits_equal(d1, d2) and bits_equal(a1, a2), kind            # lean == generic, bit for bit
        assert rel_err(o1, o2) < TOL and ds_err(s1, s2) <= TOL, kind


@pytest.mark.parametrize("round_mode,clamp", [(0, None), (1, (-7.0, 7.0))])
def test_forward_with_loss_term_equals_the_two_passes(lq, round_mode, clamp):
    """lq_mt_fq_fwd_penalty (the penalty forward fused into K1): y has the bits lq_mt_fq_fwd writes, the
    term equals lq_penalty_fwd's, and the workspace it leaves drives lq_penalty_bwd to the same gradients -- ROW and
    interleaved tensors, odd sizes, an unaligned tensor, zeros and a non-fast scale, more tensors than one table holds."""
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(77)
    shapes = [((3, 3, 64, 128), (1, 1, 64, 1)), ((7, 7, 3, 64), (1, 1, 1, 64)), ((784, 128), (1, 128)), ((128,), (1,)),
              ((37, 5, 131), (1, 5, 1)), ((9, 6, 4100), (1, 6, 1)), ((5000, 132), (1, 132)), ((513, 7, 2), (1, 7, 1))]
    shapes += [((64 + i, 24), (1, 24)) for i in range(50)]                 # > kMaxMT tensors: two rounds
    Ps, ss = [], []
    for sh, ssh in shapes:
        P = rng.normal(0, 0.05, size=sh).astype(np.float32)
        P.reshape(-1)[::19] = 0.0
        s = rng.uniform(5e-3, 2e-2, size=ssh).astype(np.float32)
        Ps.append(dev(P)); ss.append(dev(s))
    ss[2][0, 5] = 1.0                                                      # outside Difference's fast set
    buf = torch.empty(Ps[4].numel() + 1, device="cuda")                     # a tensor that is not 16-byte aligned
    buf[1:].copy_(Ps[4].reshape(-1))
    Ps[4] = buf[1:].view(shapes[4][0])
    kw = dict(round_mode=round_mode)
    if clamp:
        kw.update(qmin=clamp[0], qmax=clamp[1])
    ys_ref = [torch.empty_like(p) for p in Ps]
    ops.mt_fq_fwd(Ps, ss, ys_ref, **kw)
    for kind in (_lib.LQ_PEN_MAXBIN, _lib.LQ_PEN_DIFFERENCE):
        out_ref, ws_ref = ops.penalty_fwd(kind, Ps, ss)
        ys = [torch.full_like(p, float("nan")) for p in Ps]
        out, ws = ops.mt_fq_fwd_penalty(kind, Ps, ss, ys, **kw)
        for y, yr in zip(ys, ys_ref):
            assert bits_equal(host(y), host(yr)), kind
        assert rel_err(host(out), host(out_ref)) < TOL, kind
        dXs, dss = ops.penalty_bwd(kind, Ps, ss, ws, 0.5)
        dXr, dsr = ops.penalty_bwd(kind, Ps, ss, ws_ref, 0.5)
        for a, b in zip(dXs, dXr):
            assert bits_equal(host(a), host(b)), kind
        for a, b in zip(dss, dsr):
            assert ds_err(host(a), host(b)) <= TOL, kind
    with pytest.raises(lq._lib.LqError):                                    # Inverse reads no tensor
        ops.mt_fq_fwd_penalty(_lib.LQ_PEN_INVERSE, Ps, ss, [torch.empty_like(p) for p in Ps])


def test_penalties_many_tensors_and_a_huge_group(lq):
    """More tensors than one kernel-parameter table holds (the scalar stage chains over launches) and a
    per-tensor scale on a 9.4 M-element kernel (more than 4096 partials: the two-level fold)."""
    from learned_quantization_b200 import _lib
    ops = lq.ops
    rng = np.random.default_rng(22)
    specs = [((3, 3, 8, 16), (1, 1, 8, 1))] * 14 + [((40, 12), (40, 1))] * 14 + [((3, 3, 1024, 1024), (1,))]
    layers = []
    for sh, ssh in specs:
        K = rng.normal(0, 0.05, size=sh).astype(np.float32)
        sK = rng.uniform(5e-3, 2e-2, size=ssh).astype(np.float32)
        b = rng.normal(0, 0.05, size=(sh[-1],)).astype(np.float32)
        sb = rng.uniform(5e-3, 2e-2, size=(1,)).astype(np.float32)
        layers.append((K, sK, b, sb))
    K = layers[-1][0]
    K.reshape(-1)[[5, 77777, 9_000_000]] = np.float32(0.9)          # a three-way tie for the global maximum
    Ps, ss = [], []
    for K, sK, b, sb in layers:
        Ps += [dev(K), dev(b)]; ss += [dev(sK), dev(sb)]
    assert len(Ps) == 58
    for kind, val, grad in ((_lib.LQ_PEN_MAXBIN, o.maxbin_penalty, o.maxbin_penalty_grads),
                            (_lib.LQ_PEN_DIFFERENCE, o.difference_penalty, o.difference_penalty_grads),
                            (_lib.LQ_PEN_INVERSE, o.inverse_penalty, o.inverse_penalty_grads)):
        out, ws = ops.penalty_fwd(kind, Ps, ss)
        assert rel_err(host(out), val(layers)) < TOL, kind
        dXs, dss = ops.penalty_bwd(kind, Ps, ss, ws, 2.0)
        ref = grad(layers, 2.0, accum="fp64")
        for i in (0, 1, 28, 29, 56, 57):                              # first, middle and the huge layer
            gK, gsK, gb, gsb = ref[i // 2]
            gP, gs = (gK, gsK) if i % 2 == 0 else (gb, gsb)
            if kind != _lib.LQ_PEN_INVERSE:
                assert float(np.max(np.abs(host(dXs[i]).astype(np.float64) - gP))) <= TOL * max(float(np.max(np.abs(gP))), 1e-30), (kind, i)
            assert float(np.max(np.abs(host(dss[i]).astype(np.float64).reshape(gs.shape) - gs))) <= TOL * max(float(np.max(np.abs(gs))), 1e-30), (kind, i)


def test_host_buffer_path_matches_device_path(lq):
    ops = lq.ops
    rng = np.random.default_rng(31)
    for layout in [(64, 512, 512), (1, 1, 9_000_001), (1, 6, 3_000_000), (70000, 64, 1), (3, 5, 7), (300000, 1, 40),
                   (40000, 16, 16), (9, 2_000_000, 1)]:
        n = layout[0] * layout[1] * layout[2]
        P = rng.normal(0, 0.05, size=n).astype(np.float32)
        dy = rng.normal(0, 1e-3, size=n).astype(np.float32)
        s = rng.uniform(5e-3, 2e-2, size=layout[1]).astype(np.float32)
        hP, hdy, hs = (torch.from_numpy(a).pin_memory() for a in (P, dy, s))
        hy = torch.empty(n).pin_memory()
        hds = torch.empty(layout[1]).pin_memory()
        ws = torch.empty(ops.host_workspace_bytes(layout), dtype=torch.uint8, device="cuda")
        ops.host_fq_fwd_bwd(hP, hs, hdy, 1e-3, hy, hds, layout, ws)
        Pd, sd, dyd = dev(P), dev(s), dev(dy)
        y = ops.fq_fwd(Pd, sd, layout)
        _, ds = ops.fq_bwd_threshold(dyd, Pd, sd, 1e-3, layout)
        assert bits_equal(hy.numpy(), host(y)), layout
        assert ds_close(hds.numpy(), host(ds)), layout


def test_errors_are_loud(lq):
    from learned_quantization_b200 import LqError, _lib
    ops = lq.ops
    P = torch.zeros(4, 4, device="cuda")
    s = torch.ones(1, device="cuda")
    with pytest.raises(RuntimeError):
        ops.fq_fwd(P.cpu(), s.cpu())                      # no CPU fallback
    with pytest.raises(LqError) as e:
        ops.fq_fwd(P, s, round_mode=7)
    assert e.value.status == -3
    with pytest.raises(LqError) as e:
        ops.fq_fwd(P, s, qmin=3.0, qmax=1.0)
    assert e.value.status == -3
    with pytest.raises(ValueError):
        ops.group_layout((4, 4), (2, 2))
    rc = _lib.lib.lq_fq_bwd_threshold(P.data_ptr(), P.data_ptr(), s.data_ptr(), 0.0, None, s.data_ptr(),
                                      1, 1, 16, 0, float("-inf"), float("inf"), None, 0, None)
    assert rc == -4 and b"workspace" in _lib.lib.lq_last_error()
    rc = _lib.lib.lq_fq_fwd(None, s.data_ptr(), P.data_ptr(), None, 0, 1, 1, 16, 0, float("-inf"), float("inf"), None)
    assert rc == -1


# ------------------------------------------------------------------ BASELINE full size
@pytest.mark.parametrize("layout", [(256, 512, 512), (1, 1, 1 << 26), (1 << 17, 512, 1)])
def test_full_size_64mi_vs_oracle(lq, layout):
    """The headline configuration (64 Mi elements) against the oracle evaluated on the whole
    tensor for the forward (bit-exact) and group by group for ds."""
    ops = lq.ops
    n = layout[0] * layout[1] * layout[2]
    g = torch.Generator(device="cuda").manual_seed(42)
    Pd = torch.randn(n, device="cuda", generator=g) * 0.05
    dyd = torch.randn(n, device="cuda", generator=g) * 1e-3
    sd = torch.rand(layout[1], device="cuda", generator=g) * 1.5e-2 + 5e-3
    y = ops.fq_fwd(Pd, sd, layout)
    _, ds = ops.fq_bwd_threshold(dyd, Pd, sd, 1e-3, layout)
    _, ds2 = ops.fq_bwd_threshold(dyd, Pd, sd, 1e-3, layout, write_dx=True)
    assert bits_equal(host(ds), host(ds2))
    shape = layout
    P, dy, s = host(Pd).reshape(shape), host(dyd).reshape(shape), host(sd).reshape(1, layout[1], 1)
    q_ref, y_ref = o.fq_forward(P, s)
    assert bits_equal(host(y).reshape(shape), y_ref)
    _, ds_ref = o.nested_backward(dy, q_ref, y_ref, s, 1e-3, accum="fp64")
    assert ds_close(host(ds).reshape(ds_ref.shape), ds_ref)


# ------------------------------------------------------------------ beyond 32-bit element indices
@pytest.mark.parametrize("layout", [(1, 1, (1 << 31) + 12345), (5, 1024, 1 << 19), (5 << 24, 32, 1)])
def test_more_than_2_pow_31_elements(lq, layout):
    """Maximum sizes: tensors whose element (and byte) offsets do not fit 32 bits, in each traversal regime.
    Size-independent properties: the forward of the whole tensor equals, bit for bit, the forward of
    independent pieces of it (each below 2^31 elements, the range the oracle tests cover) and the oracle on
    windows at the start, across the 2^31-element / 2^32-byte marks and at the ragged end; the scale gradient
    of the whole tensor equals statistics of the pieces + finish."""
    ops = lq.ops
    outer, C, inner = layout
    n = outer * C * inner
    free, _ = torch.cuda.mem_get_info()
    if free < 5 * 4 * n + (4 << 30):
        pytest.skip("not enough free device memory")
    lam = 1e-3
    g = torch.Generator(device="cuda").manual_seed(7)
    Pd = torch.empty(n, device="cuda")
    dyd = torch.empty(n, device="cuda")
    step = 1 << 28
    for a in range(0, n, step):                                   # generate piecewise: bounded temporaries
        b = min(n, a + step)
        Pd[a:b] = torch.randn(b - a, device="cuda", generator=g) * 0.05
        dyd[a:b] = torch.randn(b - a, device="cuda", generator=g) * 1e-3
    sd = torch.rand(C, device="cuda", generator=g) * 1.5e-2 + 5e-3
    y = ops.fq_fwd(Pd, sd, layout)
    dX, ds = ops.fq_bwd_threshold(dyd, Pd, sd, lam, layout, write_dx=True)
    torch.cuda.synchronize()

    # pieces: along `outer` when there is one, along the flat index for the per-tensor layout
    if C == 1:
        cuts = [0, (1 << 30) + 4, n]
        piece = lambda a, b: (1, 1, b - a)
        span = lambda a, b: (a, b)
    else:
        cuts = [0, outer // 2, outer]
        piece = lambda a, b: (b - a, C, inner)
        span = lambda a, b: (a * C * inner, b * C * inner)
    stats = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        lo, hi = span(a, b)
        assert hi - lo < (1 << 31)
        y_piece = ops.fq_fwd(Pd[lo:hi], sd, piece(a, b))
        assert torch.equal(y_piece.view(torch.int32), y[lo:hi].view(torch.int32)), (a, b)
        del y_piece
        _, st = ops.fq_bwd_threshold_stats(dyd[lo:hi], Pd[lo:hi], sd, lam, piece(a, b))
        stats.append(st)
    assert torch.equal(dX.view(torch.int32), dyd.view(torch.int32))
    ds_pieces = ops.fq_bwd_threshold_finish(torch.stack(stats).contiguous(), lam, n // C)
    assert ds_close(host(ds).reshape(-1), host(ds_pieces).reshape(-1), tol=1e-6)

    # oracle on windows (whole groups rows for the grouped layouts)
    row = C * inner if C > 1 and inner < 128 else 1
    marks = [0, (1 << 30), (1 << 31), n]
    w = 1 << 16
    for m in marks:
        lo = max(0, min(n - w, m - w // 2))
        lo -= lo % (row if row > 1 else 4)
        hi = min(n, lo + w) if m != n else n
        idx = torch.arange(lo, hi, device="cuda")
        grp = host(((idx // inner) % C).to(torch.int32))
        _, y_ref = o.fq_forward(host(Pd[lo:hi]), host(sd)[grp])
        assert bits_equal(host(y[lo:hi]), y_ref), m


def test_zz_worst_scale_gradient_error_of_the_session(lq):
    """Not a check of its own: prints the largest per-group relative ds error ds_close met (bound: 1e-5)."""
    print(f"\nworst per-group scale-gradient error this session: {WORST['ds']:.3e} (tolerance {TOL:.0e}, floor 1e-6 * max|ds|)")
    assert WORST["ds"] <= TOL

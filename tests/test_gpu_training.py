"""GPU tier: the step after / around the hot path -- fused optimizer + constraint,
the one-launch-per-phase QuantBank, the TrainStep, and the model builders' tensor inventories."""
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import lq_oracle as o

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lq():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    import learned_quantization_b200 as pkg
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return pkg


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


SIZES = [1, 3, 4, 5, 127, 4096, 4097, 10000, 70001]


def _tensors(rng):
    ps = [rng.normal(0, 0.05, n).astype(np.float32) for n in SIZES]
    gs = [rng.normal(0, 1e-2, n).astype(np.float32) for n in SIZES]
    l2 = [0.0 if i % 2 else 1e-4 for i in range(len(SIZES))]
    mins = [None if i % 3 else 0.01 for i in range(len(SIZES))]
    return ps, gs, l2, mins


def test_fused_adam_matches_keras_restatement(lq):
    ops = lq.ops
    rng = np.random.default_rng(5)
    ps, gs, l2, mins = _tensors(rng)
    P, G = [dev(p) for p in ps], [dev(g) for g in gs]
    table = ops.OptimizerTable(P, G, adam=True, l2=l2, min_value=mins, weight_decay=0.0)
    ms, vs = [np.zeros_like(p) for p in ps], [np.zeros_like(p) for p in ps]
    for step in (1, 2, 3):
        scale_dev = torch.tensor(0.5, device="cuda")
        ops.opt_adam(table, 1e-3, step, grad_scale=2.0, grad_scale_dev=scale_dev)      # net scale 1.0
        for i in range(len(ps)):
            ps[i], ms[i], vs[i] = o.adam_step(ps[i], gs[i], ms[i], vs[i], 1e-3, step, l2=l2[i], min_value=mins[i])
            assert rel_err(P[i].cpu().numpy(), ps[i], tiny=1e-12) < 2e-6, (step, i)
            assert rel_err(table.m[i].cpu().numpy(), ms[i], tiny=1e-12) < 2e-6
            assert rel_err(table.v[i].cpu().numpy(), vs[i], tiny=1e-12) < 2e-6
            if mins[i] is not None:
                assert float(P[i].min()) >= float(np.float32(mins[i]))
        for g, gh in zip(G, gs):                     # new gradients every step
            gh[:] = rng.normal(0, 1e-2, gh.shape).astype(np.float32)
            g.copy_(dev(gh))


@pytest.mark.parametrize("momentum,nesterov,wd", [(0.0, False, 0.0), (0.9, False, 0.0), (0.9, True, 5e-4)])
def test_fused_sgd_matches_keras_restatement(lq, momentum, nesterov, wd):
    ops = lq.ops
    rng = np.random.default_rng(6)
    ps, gs, l2, mins = _tensors(rng)
    P, G = [dev(p) for p in ps], [dev(g) for g in gs]
    table = ops.OptimizerTable(P, G, adam=False, momentum=(momentum != 0.0), l2=l2, min_value=mins, weight_decay=wd)
    ms = [np.zeros_like(p) for p in ps]
    for _ in range(3):
        ops.opt_sgd(table, 0.1, momentum, nesterov)
        for i in range(len(ps)):
            ps[i], m = o.sgd_step(ps[i], gs[i], ms[i], 0.1, momentum, nesterov, l2=l2[i], min_value=mins[i], weight_decay=wd)
            if momentum:
                ms[i] = m
                assert rel_err(table.m[i].cpu().numpy(), ms[i], tiny=1e-12) < 2e-6
            assert rel_err(P[i].cpu().numpy(), ps[i], tiny=1e-12) < 2e-6, i


def test_optimizer_errors(lq):
    from learned_quantization_b200 import _lib
    d = (_lib.OptDesc * 1)(_lib.OptDesc(p=None, g=None, m=None, v=None, n=4, l2=0, min_value=0, weight_decay=0, reserved=0))
    assert _lib.lib.lq_opt_adam(d, 1, 1e-3, 0.9, 0.999, 1e-7, 1, None, 1.0, None, None) == -1
    assert _lib.lib.lq_opt_adam(d, 1, 1e-3, 0.9, 0.999, 1e-7, 0, None, 1.0, None, None) == -3


class _TinyNet(torch.nn.Module):
    def __init__(self, L, lam, orientation):
        super().__init__()
        from learned_quantization_b200 import RandomNormal, regularizers
        init = RandomNormal(seed=3)
        self.c1 = L.CustomConv2DLayer(seed=3, penalty_threshold=lam, orientation=orientation, initializer=init, filters=8,
                                      kernel_size=(3, 3), strides=(1, 1), padding="same", name="a",
                                      regularizer=regularizers.l2(1e-4))
        self.c2 = L.CustomConv2DLayer(seed=3, penalty_threshold=lam, orientation=orientation, initializer=init, filters=16,
                                      kernel_size=(3, 3), strides=(2, 2), padding="same", name="b", regularizer=None)
        self.fc = L.CustomDenseLayer(seed=3, units=10, penalty_threshold=lam, orientation="columnwise", initializer=init,
                                     name="fc", regularizer=None)
        self.plain = torch.nn.Linear(10, 10)

    def forward(self, x):
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x)).mean(dim=(1, 2))
        return torch.softmax(self.plain(self.fc(x)), dim=-1)


@pytest.mark.parametrize("orientation", ["channelwise", "scalar", "rowwise"])
def test_train_step_equals_per_layer_path_plus_keras_adam(lq, orientation):
    """Bank path (1 launch per phase) == per-layer autograd path (one launch per tensor): same y,
    same dP, same ds; and the fused update equals the restated Keras Adam + l2 + constraint."""
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training.step import TrainStep, quantised_pairs
    torch.manual_seed(0)
    net = _TinyNet(nq, 1e-3, orientation).cuda()
    x = torch.rand(4, 12, 12, 3, device="cuda") * 255.0
    y = torch.randint(0, 10, (4,), device="cuda")
    net(x)                                             # builds the layers
    pairs = quantised_pairs(net)
    assert len(pairs) == 6
    with torch.no_grad():                              # "trained" scales so that codes are small
        for _, _, p, q in pairs:
            q.scale.uniform_(5e-3, 2e-2)
    # per-layer path
    for p in net.parameters():
        p.grad = None
    scce(y, net(x)).mean().backward()
    ref = {id(p): p.grad.clone() for p in net.parameters()}
    before = {id(p): p.detach().clone() for p in net.parameters()}
    # bank path
    step = TrainStep(net, scce, optimizer="adam", lr=1e-3)
    step(x, y)
    b = step.bucket
    def close(a, r, tol=1e-4):       # cuDNN may pick different algorithms for the two runs: compare in max-norm
        a, r = a.cpu().numpy().reshape(-1), r.cpu().numpy().reshape(-1)
        return float(np.max(np.abs(a - r))) <= tol * max(float(np.max(np.abs(r))), 1e-30)
    for v, p in zip(b.dy, step.bank.params):
        assert close(v, ref[id(p)])
    for v, s in zip(b.ds, step.bank.scales):
        r = ref[id(s)].cpu().numpy()
        # ds depends on dy only through thresholded ratios; the two dy differ in their last bits
        assert np.max(np.abs(v.cpu().numpy() - r)) <= 2e-3 * max(np.max(np.abs(r)), 1e-30) + 1e-12
    for v, p in zip(b.plain, step.plain):
        assert close(v, ref[id(p)])
    # update: Keras Adam step 1 on those gradients; l2 on c1's kernel AND bias; constraint on scales
    l2_of = {id(net.c1.kernel): 1e-4, id(net.c1.b): 1e-4}
    scale_ids = {id(s) for s in step.bank.scales}
    grads = {id(p): v for v, p in zip(b.dy + b.plain + b.ds, step.bank.params + step.plain + step.bank.scales)}
    for p in net.parameters():
        g = grads[id(p)].cpu().numpy().reshape(-1)
        p0 = before[id(p)].cpu().numpy().reshape(-1)
        want, _, _ = o.adam_step(p0, g, np.zeros_like(p0), np.zeros_like(p0), 1e-3, 1, l2=l2_of.get(id(p), 0.0),
                                 min_value=(o.SCALE_INIT if id(p) in scale_ids else None))
        assert rel_err(p.detach().cpu().numpy().reshape(-1), want, tiny=1e-12) < 2e-6
    # layers fall back to the per-layer path once the bank is released
    step.bank.release()
    assert all(q._banked is None for q in step.bank.qlayers)


def test_mnist_step_trains_and_keeps_scales_above_floor(lq):
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training import models
    from learned_quantization_b200.training.step import TrainStep
    torch.manual_seed(1)
    net = models.mnist_mlp(1e-10, seed=42, orientation="rowwise").cuda()
    x, y = models.synthetic_batch(32, (28, 28, 1), "cuda")
    x = x / 255.0
    net(x)
    step = TrainStep(net, scce, optimizer="adam", lr=1e-3)
    losses = [float(step(x, y)) for _ in range(30)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    for s in step.bank.scales:
        assert float(s.min()) >= float(o.SCALE_INIT)


@pytest.mark.parametrize("split", [False, True])
def test_graph_replay_equals_eager_steps(lq, split):
    """A captured step replayed N times == N eager steps (same parameters, Adam slots and t).
    split=True drives the data-parallel form (three graphs, collective between them) with a
    world-2 communicator whose allreduce is the identity."""
    from learned_quantization_b200.training.dp import Comm

    class _Loopback(Comm):
        world, rank = 2, 0

        def allreduce_mean_(self, flat):
            return None

    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training.step import TrainStep
    torch.backends.cudnn.deterministic = True
    x = torch.rand(4, 12, 12, 3, device="cuda") * 255.0
    y = torch.randint(0, 10, (4,), device="cuda")
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        net = _TinyNet(nq, 1e-3, "channelwise").cuda()
        net(x)
        nets.append(net)
    eager, graphed = (TrainStep(n, scce, optimizer="adam", lr=1e-3, comm=(_Loopback() if split else None)) for n in nets)
    graphed.capture(x, y, warmup=2)                    # 2 warm-up steps + 1 captured-while-recording (not executed)
    for _ in range(3):
        graphed.replay(x, y)
    for _ in range(5):
        eager(x, y)
    torch.cuda.synchronize()
    assert int(graphed._step_dev) == int(eager._step_dev) == 5
    for a, b in zip(nets[0].parameters(), nets[1].parameters()):
        assert rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy(), tiny=1e-9) < 1e-5
    torch.backends.cudnn.deterministic = False


def test_model_builders_have_the_reference_tensor_inventory(lq):
    """40 quantised tensors / 11 166 912 kernel elements (IMAGENETTE), 22 (paper)."""
    from learned_quantization_b200.training import models
    from learned_quantization_b200.training.step import quantised_pairs
    net = models.imagenette_resnet(1e-11, orientation="channelwise").cuda().eval()
    with torch.no_grad():
        out = net(torch.rand(2, 64, 64, 3, device="cuda") * 255)
    assert out.shape == (2, 10)
    pairs = quantised_pairs(net)
    kernels = [p for _, name, p, _ in pairs if name == "kernel"]
    assert len(pairs) == 40 and len(kernels) == 20
    assert sum(k.numel() for k in kernels) == 11_166_912
    assert sum(p.numel() for _, name, p, _ in pairs if name == "b") == 4_800
    assert sum(q.scale.numel() for _, name, _, q in pairs) == 3_843 + 20
    net2 = models.cifar_resnet18(0.0, orientation="scalar").cuda().eval()
    with torch.no_grad():
        out2 = net2(torch.rand(2, 32, 32, 3, device="cuda") * 255)
    assert out2.shape == (2, 10)
    assert len(quantised_pairs(net2)) == 22


# ------------------------------------------------------------------ f2: code statistics / int8 export
@pytest.mark.parametrize("shape,s_shape,s_range", [((3, 3, 16, 32), (1, 1, 16, 1), (5e-3, 2e-2)), ((784, 128), (784, 1), (5e-3, 2e-2)),
                                                   ((5000,), (1,), (1.2e-5, 1.2e-5)), ((2, 3, 10001), (1, 3, 1), (1e-3, 1e-2))])
def test_code_stats_equal_numpy_unique(lq, shape, s_shape, s_range):
    """custom_callbacks.py:84-129: np.unique(floor(P / s), return_counts=True) and max|q|."""
    ops = lq.ops
    rng = np.random.default_rng(11)
    P = rng.normal(0, 0.05, shape).astype(np.float32)
    s = rng.uniform(*s_range, s_shape).astype(np.float32)
    q = np.floor(P / s)
    vals, cnts = np.unique(q.reshape(-1), return_counts=True)
    st = ops.code_stats(dev(P), dev(s))
    assert st["n_unique"] == len(vals) and st["outside_window"] == 0
    assert np.array_equal(st["unique"].numpy(), vals.astype(np.float32)) and np.array_equal(st["counts"].numpy(), cnts)
    assert st["min"] == vals[0] and st["max"] == vals[-1] and st["max_abs"] == np.max(np.abs(q))
    assert st["int8_overflow"] == int(np.sum((q < -128) | (q > 127)))
    # a window smaller than the range: the rest is counted, not dropped
    st2 = ops.code_stats(dev(P), dev(s), max_bins=8)
    assert st2["outside_window"] == int(np.sum(q >= vals[0] + 8)) and int(st2["counts"].sum()) + st2["outside_window"] == q.size
    assert np.array_equal(ops.codes_int8(dev(P), dev(s)).cpu().numpy(), q.astype(np.int64).astype(np.int8))


def test_save_compress_parameters_writes_the_reference_artefacts(lq, tmp_path):
    """utils/log_scripts.py:61-97: weights.npy (dict of int8 codes), weights.zip, file_sizes.log."""
    import zipfile
    from learned_quantization_b200.training import models
    from learned_quantization_b200.utils import log_scripts
    net = models.mnist_mlp(1e-10, orientation="rowwise").cuda()
    net(torch.rand(2, 28, 28, 1, device="cuda"))
    with torch.no_grad():
        net.d1.nested_q_w_layer.scale.fill_(1e-3)
    log_scripts.save_compress_parameters(net, str(tmp_path))
    w = np.load(tmp_path / "weights.npy", allow_pickle=True).item()
    assert sorted(w) == ["custom_dense_layer/W", "custom_dense_layer/b", "custom_dense_layer_1/W", "custom_dense_layer_1/b"] \
        or len(w) == 4
    k = [n for n in w if n.endswith("/W")][0]
    W = net.d1.W.detach().cpu().numpy()
    want = np.floor(W / net.d1.nested_q_w_layer.scale.detach().cpu().numpy()).astype(np.int64).astype(np.int8)
    assert w[k].dtype == np.int8 and np.array_equal(w[k], want)
    with zipfile.ZipFile(tmp_path / "weights.zip") as z:
        assert z.namelist() == ["weights.npy"]
    lines = open(tmp_path / "file_sizes.log").read().splitlines()
    assert lines[0].startswith("Weights size: ") and lines[1].startswith("Compressed weights size: ")
    # the second dense layer still has the initial scale 1.19e-5: its codes overflow int8 and are reported
    assert "custom_dense_layer" in open(tmp_path / "int8_overflow.log").read()


# ------------------------------------------------------------------ the other two variants through TrainStep
class _TinyNetCL(torch.nn.Module):
    """custom_loss_terms variant: STE layers (ds = 0 from the operator), scale gradients from the loss term."""

    def __init__(self, L, orientation):
        super().__init__()
        from learned_quantization_b200 import RandomNormal
        init = RandomNormal(seed=4)
        self.c1 = L.CustomConv2DLayer(seed=4, penalty_rate=1e-3, orientation=orientation, initializer=init, filters=8,
                                      kernel_size=(3, 3), strides=(1, 1), padding="same", name="a", regularizer=None)
        self.fc = L.CustomDenseLayer(seed=4, units=10, penalty_rate=1e-3, orientation="rowwise", initializer=init,
                                     name="fc", regularizer=None)

    def forward(self, x):
        return torch.softmax(self.fc(torch.relu(self.c1(x)).mean(dim=(1, 2))), dim=-1)


@pytest.mark.parametrize("cls_name", ["SCCEMaxBin", "SCCEDifference", "SCCEInverse"])
def test_custom_loss_variant_step_equals_per_layer_autograd(lq, cls_name, tmp_path):
    """IMAGENETTE/custom_loss_terms/experiment.py:784-793: loss = scce + gamma * penalty(P, s).  Through the
    bank, dP = dy (STE) + penalty gradient, ds = penalty gradient; must equal the per-layer autograd path."""
    from learned_quantization_b200.custom_loss_terms.custom_components import custom_layers as cl
    from learned_quantization_b200.custom_loss_terms.custom_components import custom_loss_functions as lf
    from learned_quantization_b200.training.step import TrainStep
    torch.manual_seed(0)
    net = _TinyNetCL(cl, "channelwise").cuda()
    x = torch.rand(4, 10, 10, 3, device="cuda") * 255.0
    y = torch.randint(0, 10, (4,), device="cuda")
    net(x)
    with torch.no_grad():
        for q in (net.c1.nested_q_k_layer, net.c1.nested_q_b_layer, net.fc.nested_q_w_layer, net.fc.nested_q_b_layer):
            q.scale.uniform_(5e-3, 2e-2)
    loss_obj = getattr(lf, cls_name)(layers=[net.c1, net.fc], penalty_rate=0.5, log_dir=str(tmp_path))
    loss_obj.log_to_files = False
    for p in net.parameters():
        p.grad = None
    loss_obj.compute_total_loss(y, net(x)).mean().backward()
    ref = {id(p): (None if p.grad is None else p.grad.clone()) for p in net.parameters()}
    # fuse_loss: MaxBin / Difference come out of the step's forward pass as by-products (lq_mt_fq_fwd_penalty); the
    # stand-alone pass (fuse_loss=False, and always for Inverse) must give the same step
    for fuse in (True, False):
        step = TrainStep(net, loss_obj.compute_total_loss, optimizer="sgd", lr=0.0, fuse_loss=fuse)     # lr 0: parameters stay put
        assert step.fused_loss == (fuse and cls_name != "SCCEInverse")
        loss = step(x, y)
        assert loss_obj._fused is None, "the fused term must not outlive its step"
        b = step.bucket
        for v, p in list(zip(b.dy, step.bank.params)) + list(zip(b.ds, step.bank.scales)):
            r = ref[id(p)]
            r = torch.zeros_like(p) if r is None else r
            a, rr = v.cpu().numpy().reshape(-1), r.cpu().numpy().reshape(-1)
            assert float(np.max(np.abs(a - rr))) <= 1e-4 * max(float(np.max(np.abs(rr))), 1e-30), (cls_name, fuse, tuple(p.shape))
        assert any(float(v.abs().max()) > 0 for v in b.ds), "the loss term must produce scale gradients"
        if fuse:
            loss_fused = float(loss)
        else:
            assert abs(float(loss) - loss_fused) <= 1e-6 * abs(loss_fused)


def test_paper_variant_step_with_sgd_nesterov(lq):
    """CIFAR-10/paper_implementation: STE layers without conv bias, SGD(0.1, 0.9, nesterov, wd 5e-4)."""
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training import models
    from learned_quantization_b200.training.step import TrainStep
    torch.manual_seed(2)
    net = models.cifar_resnet18(0.0, orientation="scalar").cuda().train()
    x, y = models.synthetic_batch(8, (32, 32, 3), "cuda")
    x = x / 255.0
    net(x)
    loss_fn = lambda yt, logits: scce(yt, torch.softmax(logits, dim=-1))
    step = TrainStep(net, loss_fn, optimizer="sgd", lr=0.01, momentum=0.9, nesterov=True, weight_decay=5e-4)
    assert len(step.bank.params) == 22 and step.bank.sgrad_mode == 0
    before = [p.detach().clone() for p in step.bank.params[:3]]
    losses = [float(step(x, y)) for _ in range(8)]
    assert np.isfinite(losses).all()
    assert all(float(v.abs().max()) == 0.0 for v in step.bucket.ds)        # PI:62-64: scale gradient is zeros
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, step.bank.params[:3]))


def test_tracking_callbacks_write_the_reference_log_files(lq, tmp_path):
    """custom_callbacks.py:9-237: same files, same 'Epoch N' + value-per-line format, numbers from the device."""
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_callbacks as cb
    from learned_quantization_b200.training import models
    net = models.mnist_mlp(1e-10, orientation="rowwise").cuda()
    net(torch.rand(2, 28, 28, 1, device="cuda"))
    with torch.no_grad():
        net.d1.nested_q_w_layer.scale.uniform_(5e-3, 2e-2)
        net.d1.nested_q_b_layer.scale.fill_(1e-3)
        net.d1.b.normal_(0, 0.05)
    c = cb.NestedScaleTrackingCallback(net.d1, str(tmp_path))
    c.on_train_begin()
    c.on_epoch_end(0)
    c.on_epoch_end(1)
    c.on_train_end()
    W = net.d1.W.detach().cpu().numpy()
    q = np.floor(W / net.d1.nested_q_w_layer.scale.detach().cpu().numpy())
    lines = open(c.log_file_path_qk_epoch).read().split()
    assert lines == ["Epoch", "0", str(len(np.unique(q))), "Epoch", "1", str(len(np.unique(q)))]
    mx = open(c.log_file_path_qk_epoch_max).read().splitlines()
    assert mx[0] == "Epoch 0" and len(mx) == 2 * (1 + 784)
    assert np.allclose([float(v) for v in mx[1:785]], np.max(np.abs(q), axis=1))
    uniq = [l.split(", ") for l in open(c.log_file_path_qk_unique).read().splitlines()]
    vals, cnts = np.unique(q.reshape(-1), return_counts=True)
    assert [float(a) for a, _ in uniq] == list(vals.astype(float)) and [int(b) for _, b in uniq] == list(cnts)
    assert len(open(c.log_file_path_qk).read().splitlines()) == q.size
    assert os.path.basename(c.log_file_path_k_scale) == "Weights_(784, 128)_Rowwise-scaler_(784, 1).log"
    a = cb.AccuracyLossTrackingCallBack(str(tmp_path))
    a.on_epoch_end(0, {"val_accuracy": 0.5, "val_loss": 1.0, "accuracy": 0.25, "loss": 2.0})
    assert open(tmp_path / "accuracy" / "train_accuracy.log").read() == "Epoch 0\n0.25\n"


# ------------------------------------------------------------------ f3: kernels stored in cuDNN's order
@pytest.mark.parametrize("orientation", ["channelwise", "rowwise", "columnwise", "scalar"])
def test_ohwi_kernel_storage_is_the_same_operator(lq, orientation):
    """A kernel kept in (Cout, kh, kw, Cin) memory order is the same logical HWIO tensor with the same groups:
    fake-quant output, dP and ds equal the HWIO-stored layer's, per layer and through the bank."""
    from learned_quantization_b200 import RandomNormal
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training.step import TrainStep

    class Net(torch.nn.Module):
        def __init__(self, storage):
            super().__init__()
            init = RandomNormal(seed=5)
            self.c1 = nq.CustomConv2DLayer(seed=5, penalty_threshold=1e-3, orientation=orientation, initializer=init,
                                           filters=16, kernel_size=(3, 3), strides=(1, 1), padding="same", name="a",
                                           regularizer=None, kernel_storage=storage)
            self.c2 = nq.CustomConv2DLayer(seed=5, penalty_threshold=1e-3, orientation=orientation, initializer=init,
                                           filters=10, kernel_size=(3, 3), strides=(2, 2), padding="same", name="b",
                                           regularizer=None, kernel_storage=storage)
            self.c3 = nq.CustomConv2DLayer(seed=5, penalty_threshold=1e-3, orientation=orientation, initializer=init,
                                           filters=10, kernel_size=(1, 1), strides=(1, 1), padding="same", name="c",
                                           regularizer=None, kernel_storage=storage)     # size-1 dims: ambiguous strides

        def forward(self, x):
            return torch.softmax(self.c3(self.c2(torch.relu(self.c1(x)))).mean(dim=(1, 2)), dim=-1)

    x = torch.rand(4, 12, 12, 8, device="cuda")
    y = torch.randint(0, 10, (4,), device="cuda")
    nets = {}
    for storage in ("hwio", "ohwi"):
        net = Net(storage).cuda()
        net(x)
        nets[storage] = net
    a, b = nets["hwio"], nets["ohwi"]
    assert b.c1.kernel.shape == a.c1.kernel.shape and not b.c1.kernel.is_contiguous()
    assert b.c1.kernel.permute(3, 0, 1, 2).is_contiguous()
    with torch.no_grad():
        for la, lb in ((a.c1, b.c1), (a.c2, b.c2), (a.c3, b.c3)):
            lb.kernel.copy_(la.kernel); lb.b.copy_(la.b)
            for qa, qb in ((la.nested_q_k_layer, lb.nested_q_k_layer), (la.nested_q_b_layer, lb.nested_q_b_layer)):
                qa.scale.uniform_(5e-3, 2e-2); qb.scale.copy_(qa.scale)
    # the operator itself: bit-identical values at the same logical positions
    ya = a.c1.nested_q_k_layer(a.c1.kernel)
    yb = b.c1.nested_q_k_layer(b.c1.kernel)
    assert lq.ops.same_layout(yb, b.c1.kernel) and torch.equal(ya, yb)
    grads = {}
    for name, net in nets.items():
        for p in net.parameters():
            p.grad = None
        scce(y, net(x)).mean().backward()
        grads[name] = [p.grad.clone() for p in net.parameters()]
    for ga, gb in zip(grads["hwio"], grads["ohwi"]):
        assert float((ga - gb).abs().max()) <= 1e-4 * max(float(ga.abs().max()), 1e-30) + 1e-12
    # through the bank: gradients land in the bucket in OHWI order, parameters update identically
    steps = {k: TrainStep(n, scce, optimizer="adam", lr=1e-3) for k, n in nets.items()}
    for k in steps:
        steps[k](x, y)
    assert lq.ops.same_layout(steps["ohwi"].bucket.dy[0], b.c1.kernel)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert float((pa - pb).abs().max()) <= 1e-5 * max(float(pa.abs().max()), 1e-30) + 1e-9


@pytest.mark.parametrize("rows,cols", [(1, 1), (257, 3), (64 * 56 * 56, 64), (4097, 512), (1000, 10), (12345, 130)])
def test_colsum_is_the_bias_gradient(lq, rows, cols):
    x = torch.randn(rows, cols, device="cuda")
    got = lq.ops.colsum(x)
    want = x.double().sum(dim=0)
    assert float((got.double() - want).abs().max()) <= 1e-6 * max(float(want.abs().max()), 1.0) * max(1.0, rows ** 0.5 / 30)
    assert torch.equal(got, lq.ops.colsum(x)), "deterministic"


def test_conv_bias_gradient_matches_autograd(lq):
    from learned_quantization_b200.custom_components._impl import conv2d_nhwc
    torch.manual_seed(3)
    x = torch.randn(4, 14, 14, 8, device="cuda", requires_grad=True)
    k = torch.randn(3, 3, 8, 16, device="cuda", requires_grad=True)
    b = torch.randn(16, device="cuda", requires_grad=True)
    for strides, pad in (((1, 1), "same"), ((2, 2), "same"), ((1, 1), "valid")):
        for t in (x, k, b):
            t.grad = None
        conv2d_nhwc(x, k, strides, pad, bias=b).square().sum().backward()
        got = [t.grad.clone() for t in (x, k, b)]
        for t in (x, k, b):
            t.grad = None
        (conv2d_nhwc(x, k, strides, pad) + b).square().sum().backward()
        for g, t in zip(got, (x, k, b)):
            assert float((g - t.grad).abs().max()) <= 1e-4 * float(t.grad.abs().max())


# ------------------------------------------------------------------ BASELINE configs[0]: the MNIST model, step by step
def test_mnist_training_steps_follow_the_oracle(lq):
    """MNIST/nested_quantization_layer/experiment.py:121-173,303-305 restated in numpy on top of the oracle
    (fake-quant forward, nested backward, Keras Adam, MinValueConstraint) against TrainStep, three steps:
    the quantised forward is bit-exact, so the trajectories differ only by matmul / softmax rounding."""
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training import models
    from learned_quantization_b200.training.step import TrainStep
    lam, lr, B = 1e-10, 1e-3, 32
    torch.manual_seed(4)
    net = models.mnist_mlp(lam, seed=42, orientation="rowwise").cuda()
    x, y = models.synthetic_batch(B, (28, 28, 1), "cuda", seed=5)
    x = x / 255.0
    net(x)
    rng = np.random.default_rng(6)
    with torch.no_grad():
        for q in (net.d1.nested_q_w_layer, net.d1.nested_q_b_layer, net.d2.nested_q_w_layer, net.d2.nested_q_b_layer):
            q.scale.copy_(dev(rng.uniform(2e-3, 8e-3, tuple(q.scale.shape)).astype(np.float32)))
        net.d1.b.copy_(dev(rng.normal(0, 0.05, 128).astype(np.float32)))
        net.d2.b.copy_(dev(rng.normal(0, 0.05, 10).astype(np.float32)))
    names = ["W1", "b1", "W2", "b2"]
    P = dict(zip(names, (t.detach().cpu().numpy().copy() for t in (net.d1.W, net.d1.b, net.d2.W, net.d2.b))))
    S = dict(zip(names, (q.scale.detach().cpu().numpy().copy() for q in
                         (net.d1.nested_q_w_layer, net.d1.nested_q_b_layer, net.d2.nested_q_w_layer, net.d2.nested_q_b_layer))))
    from oracle.mnist_step import MnistOracle
    ref = MnistOracle(P, S, variant="nested", value=lam, lr=lr)
    P, S = ref.P, ref.S
    X = x.flatten(1).cpu().numpy().astype(np.float32)
    Y = y.cpu().numpy()
    step = TrainStep(net, scce, optimizer="adam", lr=lr)
    for it in range(1, 4):
        loss_ref = ref.step(X, Y)              # oracle/mnist_step.py: the reference's step in numpy
        # ---- our step
        loss = float(step(x, y))
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref), (it, loss, loss_ref)
        got_P = dict(zip(names, (t.detach().cpu().numpy() for t in (net.d1.W, net.d1.b, net.d2.W, net.d2.b))))
        got_S = dict(zip(names, (q.scale.detach().cpu().numpy() for q in
                                 (net.d1.nested_q_w_layer, net.d1.nested_q_b_layer, net.d2.nested_q_w_layer, net.d2.nested_q_b_layer))))
        for k in names:
            # Adam's first steps move every weight by ~lr whatever the gradient's size: compare on that scale
            assert float(np.max(np.abs(got_P[k] - P[k]))) <= 0.02 * lr, (it, k)
            assert float(np.max(np.abs(got_S[k] - S[k]))) <= 0.02 * lr, (it, k)


# ------------------------------------------------------------------ f4: the reference's own callbacks as the oracle
@pytest.mark.parametrize("kind", ["conv", "dense"])
def test_callbacks_write_what_the_reference_callbacks_write(lq, kind, tmp_path):
    """tests/golden/callback_logs.* = every file the reference's custom_callbacks.py (run unmodified on the TF shim,
    tests/golden/make_callback_golden.py) wrote for one conv / one dense layer over on_train_begin, two epochs and
    on_train_end.  The drop-in callbacks, fed the same parameters from device memory, must write the same tree:
    same file names, same bytes -- so whatever parses the reference's logs (utils/plot_scripts.py, checked on
    the fixture by tests/test_callback_logs.py) parses ours."""
    import json
    from learned_quantization_b200 import RandomNormal
    from learned_quantization_b200.custom_components import _impl
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_callbacks as cbs
    from learned_quantization_b200.nested_quantization_layer.custom_components import custom_layers as nq
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(gold, "callback_logs.json")) as f:
        fx = json.load(f)
    arr = np.load(os.path.join(gold, "callback_logs.npz"))
    _impl._UIDS.clear()                              # Keras auto-names start afresh, as in the fixture run
    init = RandomNormal(seed=1)
    if kind == "conv":
        layer = nq.CustomConv2DLayer(seed=1, penalty_threshold=1e-3, orientation="channelwise", initializer=init, filters=8,
                                     kernel_size=(3, 3), strides=(1, 1), padding="same", name="x", regularizer=None).cuda()
        layer(torch.zeros(1, 5, 5, 4, device="cuda"))
        param, qk = layer.kernel, layer.nested_q_k_layer
    else:
        layer = nq.CustomDenseLayer(seed=1, units=6, penalty_threshold=1e-3, orientation="rowwise", initializer=init,
                                    name="x", regularizer=None).cuda()
        layer(torch.zeros(1, 20, device="cuda"))
        param, qk = layer.W, layer.nested_q_w_layer
    assert layer.name == fx["layer_names"][kind]

    def stage(i):
        with torch.no_grad():
            param.copy_(dev(arr[f"{kind}/stage{i}/K"])); layer.b.copy_(dev(arr[f"{kind}/stage{i}/b"]))
            qk.scale.copy_(dev(arr[f"{kind}/stage{i}/sK"])); layer.nested_q_b_layer.scale.copy_(dev(arr[f"{kind}/stage{i}/sb"]))
    run = str(tmp_path / "run")
    stage(0)
    cb = cbs.NestedScaleTrackingCallback(layer, run)
    cb.on_train_begin()
    for epoch in (0, 1):
        stage(epoch + 1)
        cb.on_epoch_end(epoch)
    cb.on_train_end()
    acc = cbs.AccuracyLossTrackingCallBack(run)
    for epoch, logs in enumerate(({"val_accuracy": 0.5, "val_loss": 1.25, "accuracy": 0.4375, "loss": 1.5},
                                  {"val_accuracy": 0.625, "val_loss": 1.0, "accuracy": 0.5625, "loss": 1.125})):
        acc.on_epoch_end(epoch, logs)
    got = {}
    for root, _, fs in os.walk(run):
        for fn in fs:
            p = os.path.join(root, fn)
            with open(p) as fh:
                got[os.path.relpath(p, run)] = fh.read()
    want = fx["files"][kind]
    assert sorted(got) == sorted(want)
    for k in sorted(want):
        assert got[k] == want[k], k


# ------------------------------------------------------------------ f1 against an independent implementation
@pytest.mark.parametrize("kind", ["adam", "sgd_nesterov"])
def test_fused_optimizers_equal_torch_optim(lq, kind):
    """lq_opt_adam / lq_opt_sgd against torch.optim (float64) through the exact mappings of
    tests/test_optimizer_crosscheck.py (m = -lr * buf; eps' = eps / sqrt(1 - beta2^t); Keras' decoupled weight decay by
    hand): an implementation of the update rules that shares nothing with the oracle's restatement."""
    ops = lq.ops
    rng = np.random.default_rng(12)
    shapes = [(3, 3, 16, 32), (32,), (257,), (64, 10)]
    p0 = [rng.normal(0, 0.05, sh).astype(np.float32) for sh in shapes]
    ps = [dev(a) for a in p0]
    gs = [torch.zeros_like(p) for p in ps]
    ws = [torch.tensor(a, dtype=torch.float64, device="cuda", requires_grad=True) for a in p0]
    steps, b1, b2, eps = 12, 0.9, 0.999, 1e-7
    if kind == "adam":
        lr, wd = 1e-4, 0.0
        table = ops.OptimizerTable(ps, gs, adam=True)
        opt = torch.optim.Adam(ws, lr=lr, betas=(b1, b2), eps=eps)
    else:
        lr, wd = 0.1, 5e-4
        table = ops.OptimizerTable(ps, gs, adam=False, momentum=True, weight_decay=wd)
        opt = torch.optim.SGD(ws, lr=lr, momentum=0.9, nesterov=True)
    for t in range(1, steps + 1):
        for g, w in zip(gs, ws):
            g.copy_(dev((rng.normal(0, 1e-2, tuple(g.shape)) * (1 + 0.1 * t)).astype(np.float32)))
            w.grad = g.to(torch.float64)
        if kind == "adam":
            ops.opt_adam(table, lr, t, b1, b2, eps)
            opt.param_groups[0]["eps"] = eps / np.sqrt(1.0 - b2 ** t)
        else:
            ops.opt_sgd(table, lr, 0.9, True)
            with torch.no_grad():
                for w in ws:
                    w.mul_(1.0 - lr * wd)
        opt.step()
    for a, p, w in zip(p0, ps, ws):
        got, want = p.detach().cpu().numpy().astype(np.float64) - a, w.detach().cpu().numpy() - a
        assert np.max(np.abs(got - want)) <= 5e-5 * np.max(np.abs(want)), kind

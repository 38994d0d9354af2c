"""CPU tier: host-side layout logic (no kernels run): the [outer, C, inner] view of a tensor as it lies in
memory, for contiguous tensors and for kernels stored in cuDNN's order, against a brute-force group map."""
import itertools

import numpy as np
import pytest
import torch

from oracle import lq_oracle as o


@pytest.fixture(scope="module")
def ops():
    import learned_quantization_b200 as pkg
    return pkg.ops


def brute_force_groups(t, s_shape):
    """group id of every element in MEMORY order, from the logical broadcasting rule of NQ:56-58."""
    gid = torch.arange(int(np.prod(s_shape))).reshape(s_shape)
    logical = gid.expand(t.shape) if len(s_shape) == t.dim() else gid.reshape([1] * t.dim()).expand(t.shape)
    perm = sorted(range(t.dim()), key=lambda d: (-t.stride(d), d))
    return logical.permute(perm).reshape(-1).numpy()


@pytest.mark.parametrize("orientation", ["rowwise", "columnwise", "channelwise", "scalar"])
@pytest.mark.parametrize("storage", ["hwio", "ohwi", "dense2d", "bias"])
def test_tensor_layout_matches_the_broadcasting_rule(ops, orientation, storage):
    if storage == "hwio":
        t = torch.zeros(3, 2, 5, 4)
    elif storage == "ohwi":
        t = torch.zeros(4, 3, 2, 5).permute(1, 2, 3, 0)
        assert t.shape == (3, 2, 5, 4) and not t.is_contiguous()
    elif storage == "dense2d":
        t = torch.zeros(6, 7)
    else:
        t = torch.zeros(9)
    if t.dim() < 3 and orientation == "channelwise" and t.dim() < 2:
        pytest.skip("a bias only ever takes the scalar scale")
    try:
        s_shape = o.scale_shape(tuple(t.shape), orientation)
    except Exception:
        pytest.skip("orientation not defined for this rank")
    if len(s_shape) != 1 and len(s_shape) != t.dim():
        pytest.skip("not broadcastable")
    outer, C, inner = ops.tensor_layout(t, s_shape)
    assert outer * C * inner == t.numel()
    idx = np.arange(t.numel())
    got = (idx // inner) % C
    want = brute_force_groups(t, s_shape)
    # same partition of the elements into groups, with the same numbering
    assert np.array_equal(got, want), (storage, orientation, (outer, C, inner))
    if t.is_contiguous():
        assert (outer, C, inner) == o.group_layout(tuple(t.shape), tuple(s_shape))


def test_mem_perm_dense_and_match_strides(ops):
    a = torch.arange(24.0).reshape(2, 3, 4)
    assert ops.mem_perm(a) == (0, 1, 2)
    b = a.permute(2, 0, 1)
    assert ops.mem_perm(b) == (1, 2, 0) and ops.dense(b) is b
    c = a[:, ::2]                               # not a permutation of a contiguous block
    assert ops.mem_perm(c) is None and ops.dense(c).is_contiguous()
    d = ops.match_strides(a.permute(2, 0, 1).contiguous(), b)
    assert d.stride() == b.stride() and torch.equal(d, b)
    one = torch.zeros(5, 1, 1, 7).permute(1, 2, 3, 0)     # size-1 dims have arbitrary strides
    like = torch.zeros(1, 1, 7, 5)
    assert ops.same_layout(one.contiguous(), like)
    with pytest.raises(ValueError):
        ops.group_layout((3, 4), (3, 4))        # two non-1 axes
    with pytest.raises(ValueError):
        ops.scale_shape((3, 3), "diagonal")


def test_grad_bucket_mirrors_parameter_memory():
    from learned_quantization_b200.training.dp import GradBucket
    k = torch.arange(4 * 3 * 2 * 5, dtype=torch.float32).reshape(4, 3, 2, 5).permute(1, 2, 3, 0)    # logical HWIO
    b = GradBucket([k, (6,)], [(2, 2)], [(1, 1, 5, 1), (1,)], "cpu")
    assert b.dy[0].shape == k.shape and b.dy[0].stride() == k.stride()
    b.dy[0].copy_(k)
    flat = b.flat[: k.numel()]
    assert torch.equal(flat, torch.arange(float(k.numel())))   # memory order of the bucket == memory order of k
    offs = [v.data_ptr() - b.flat.data_ptr() for v in b.dy + b.plain + b.ds]
    assert offs == sorted(offs) and all(o_ % 16 == 0 for o_ in offs)
    assert b.exact_part().numel() == offs[3] // 4 and b.all().numel() >= b.exact_part().numel() + 5 + 1


def test_default_initializer_uses_keras_glorot_fans():
    """initializer=None falls back to what Keras' add_weight would do: glorot_uniform with Keras' fans --
    receptive_field * (Cin, Cout) for a kernel, n for a vector (biases are NOT zero-filled)."""
    from learned_quantization_b200.custom_components import _impl

    class L(_impl._HostLayer):
        initializer = None
    layer = L()
    torch.manual_seed(0)
    for shape, fans in (((3, 3, 16, 32), (9 * 16, 9 * 32)), ((784, 128), (784, 128)), ((64,), (64, 64))):
        limit = (6.0 / sum(fans)) ** 0.5
        w = layer._init_weight(shape, None, "cpu").detach()
        assert tuple(w.shape) == shape and float(w.abs().max()) <= limit
        assert float(w.abs().max()) > 0.8 * limit and abs(float(w.mean())) < 0.2 * limit


# ------------------------------------------------------------------ geometry of multi-tensor launches (host only)
def _mt_geometry(layouts, phase):
    import ctypes as C
    import learned_quantization_b200 as pkg
    from learned_quantization_b200 import _lib
    arr = pkg.ops.desc_array([pkg.ops.make_desc(layout=l) for l in layouts])
    out = (C.c_int64 * (4 * len(layouts)))()
    _lib.check(_lib.lib.lq_debug_mt_geometry(arr, len(layouts), phase, out))
    return [tuple(out[4 * i:4 * i + 4]) for i in range(len(layouts))]


IMAGENETTE_OHWI = [(64 * 49, 3, 1)] + [(64 * 9, 64, 1)] * 4 + [(128 * 9, 64, 1)] + [(128 * 9, 128, 1)] * 3 + [(128, 64, 1)] + \
                  [(256 * 9, 128, 1)] + [(256 * 9, 256, 1)] * 3 + [(256, 128, 1)] + [(512 * 9, 256, 1)] + \
                  [(512 * 9, 512, 1)] * 3 + [(512, 256, 1)]


@pytest.mark.parametrize("phase_name", ["LQ_PHASE_FWD", "LQ_PHASE_BWD"])
def test_interleaved_tensors_of_a_table_share_one_grid(phase_name):
    """lq_traverse.cuh mt_col_targets: by itself an interleaved tensor gets enough blocks to fill the GPU; 20 conv kernels
    (O,H,W,I storage, channelwise scales: all interleaved) in ONE launch must not get 20 times that.  Chunks only grow, so
    the partials never outgrow the workspace sized for the stand-alone geometry; a block never exceeds 64 Ki elements."""
    from learned_quantization_b200 import _lib
    phase = getattr(_lib, phase_name)
    alone = [_mt_geometry([l], phase)[0] for l in IMAGENETTE_OHWI]
    table = _mt_geometry(IMAGENETTE_OHWI, phase)
    assert all(a[0] == 1 and t[0] == 1 for a, t in zip(alone, table))            # all interleaved
    assert sum(t[2] for t in table) <= 148 * 8 < sum(a[2] for a in alone)        # one shared grid instead of twenty
    for l, a, t in zip(IMAGENETTE_OHWI, alone, table):
        if phase == _lib.LQ_PHASE_BWD:
            assert t[1] >= a[1] and t[3] <= a[3], (l, a, t)                      # longer chunks, fewer partials
        cols_per_block = min(l[1], 1024 if phase == _lib.LQ_PHASE_FWD else 256)
        assert t[1] * cols_per_block <= 65536 + cols_per_block * 64, (l, t)      # (one rounding unit of slack)
    # the biases (rows of one group) and a tensor that is the table's only interleaved one are untouched
    mixed = [(1, 1, 64), (131072, 512, 1), (1, 1, 512)]
    assert _mt_geometry(mixed, phase) == [_mt_geometry([l], phase)[0] for l in mixed]


def test_mt_geometry_rejects_bad_arguments():
    import learned_quantization_b200 as pkg
    from learned_quantization_b200 import _lib
    with pytest.raises(_lib.LqError):
        _mt_geometry([(4, 4, 4)], _lib.LQ_PHASE_PENALTY_FWD)

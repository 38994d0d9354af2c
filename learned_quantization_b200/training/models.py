"""The reference's model builders on the drop-in layers -- benchmark harness only (SURVEY 2:
experiment.py / resnet.py are out of scope as drivers; their ARCHITECTURES fix the shapes,
counts and order of the quantised tensors the hot path sees per step).

    mnist_mlp            MNIST/nested_quantization_layer/experiment.py:121-173
    imagenette_resnet    IMAGENETTE/nested_quantization_layer/experiment.py:627-760
    cifar_resnet18       CIFAR-10/paper_implementation/resnet.py:30-140

Activations are NHWC like the reference's; BatchNorm uses Keras defaults (momentum 0.99,
epsilon 1e-3) unless the builder says otherwise.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .. import regularizers
from ..initializers import RandomNormal


def _layers(variant: str):
    if variant == "nested":
        from ..nested_quantization_layer.custom_components import custom_layers as m
    elif variant == "custom_loss":
        from ..custom_loss_terms.custom_components import custom_layers as m
    elif variant == "paper":
        from ..paper_implementation.custom_components import custom_layers as m
    else:
        raise ValueError(f"unknown variant {variant!r}")
    return m


def _q_kwargs(variant: str, value: float):
    return {"penalty_rate": value} if variant == "custom_loss" else {"penalty_threshold": value}


class _BN(nn.Module):
    """BatchNormalization over the last (channel) axis of an NHWC tensor."""

    def __init__(self, channels, momentum=0.99, eps=1e-3):
        super().__init__()
        self.bn = nn.BatchNorm2d(channels, eps=eps, momentum=1.0 - momentum)

    def forward(self, x):
        return self.bn(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)


def _same_pool(x, k, s):
    """MaxPooling2D(pool_size=k, strides=s, padding='same') on NHWC."""
    h, w = x.shape[1], x.shape[2]
    ph = max((-(-h // s) - 1) * s + k - h, 0)
    pw = max((-(-w // s) - 1) * s + k - w, 0)
    xc = x.permute(0, 3, 1, 2)
    if ph // 2 == 0 and pw // 2 == 0:
        # all of TF's SAME padding falls on the bottom/right: windows clipped at the border
        # (ceil_mode) are the same thing as -inf padding there, without materialising it
        return F.max_pool2d(xc, k, s, ceil_mode=True).permute(0, 2, 3, 1)
    if ph % 2 == 0 and pw % 2 == 0:
        return F.max_pool2d(xc, k, s, padding=(ph // 2, pw // 2)).permute(0, 2, 3, 1)
    xc = F.pad(xc, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
    return F.max_pool2d(xc, k, s).permute(0, 2, 3, 1)


class MnistMLP(nn.Module):
    """Flatten -> CustomDense(128) -> relu -> CustomDense(10) -> softmax."""

    def __init__(self, value: float, seed: int, orientation: str, variant: str = "nested"):
        super().__init__()
        L = _layers(variant)
        init = RandomNormal(seed=seed)
        kw = _q_kwargs(variant, value)
        self.d1 = L.CustomDenseLayer(seed=seed, units=128, orientation=orientation, initializer=init,
                                     name="custom_dense_layer_1", regularizer=None, **kw)
        self.d2 = L.CustomDenseLayer(seed=seed, units=10, orientation=orientation, initializer=init,
                                     name="custom_dense_layer_2", regularizer=None, **kw)

    def forward(self, x):
        x = torch.relu(self.d1(x.flatten(1)))
        return torch.softmax(self.d2(x), dim=-1)


class _ResidualBlock(nn.Module):
    """experiment.py:627-704: conv-bn-relu, conv-bn, (1x1 conv-bn shortcut when strided), add, relu."""

    def __init__(self, L, kw, cin, filters, orientation, stride, seed, storage):
        super().__init__()
        init = RandomNormal(seed=seed)
        reg = regularizers.l2(1e-4)
        mk = lambda ks, st, nm: L.CustomConv2DLayer(seed=seed, orientation=orientation, initializer=init, filters=filters,
                                                    kernel_size=ks, strides=st, padding="same", name=nm,
                                                    regularizer=reg, kernel_storage=storage, **kw)
        self.c1, self.b1 = mk((3, 3), stride, "conv_0"), _BN(filters)
        self.c2, self.b2 = mk((3, 3), (1, 1), "conv_1"), _BN(filters)
        self.short = None
        if tuple(stride) != (1, 1):
            self.short, self.bs = mk((1, 1), stride, "custom_conv2d_layer_shortcut"), _BN(filters)

    def forward(self, x):
        sc = x
        x = torch.relu(self.b1(self.c1(x)))
        x = self.b2(self.c2(x))
        if self.short is not None:
            sc = self.bs(self.short(sc))
        return torch.relu(x + sc)


class ImagenetteResNet(nn.Module):
    """experiment.py:707-760: 7x7/2 conv, bn, relu, maxpool 3/2, dropout .2, 8 residual blocks
    (64,64,128,128,256,256,512,512), dropout .3, GAP, dropout .3, Dense(10, softmax).
    20 quantised convs with bias -> 40 quantised tensors per step."""

    def __init__(self, value: float, seed: int, orientation: str, variant: str = "nested", num_classes: int = 10,
                 kernel_storage: str = "hwio"):
        super().__init__()
        L = _layers(variant)
        kw = _q_kwargs(variant, value)
        init = RandomNormal(seed=seed)
        self.stem = L.CustomConv2DLayer(seed=seed, orientation=orientation, initializer=init, filters=64,
                                        kernel_size=(7, 7), strides=(2, 2), padding="same",
                                        name="custom_conv2d_layer_64_0", regularizer=None, kernel_storage=kernel_storage, **kw)
        self.bn = _BN(64)
        cfg = [(64, 64, (1, 1)), (64, 64, (1, 1)), (64, 128, (2, 2)), (128, 128, (1, 1)),
               (128, 256, (2, 2)), (256, 256, (1, 1)), (256, 512, (2, 2)), (512, 512, (1, 1))]
        self.blocks = nn.ModuleList(_ResidualBlock(L, kw, ci, co, orientation, st, seed, kernel_storage) for ci, co, st in cfg)
        self.fc = nn.Linear(512, num_classes)
        with torch.no_grad():
            self.fc.weight.copy_(init((num_classes, 512)))
            self.fc.bias.zero_()

    def forward(self, x):
        x = torch.relu(self.bn(self.stem(x)))
        x = F.dropout(_same_pool(x, 3, 2), 0.2, self.training)
        for blk in self.blocks:
            x = blk(x)
        x = F.dropout(x, 0.3, self.training).mean(dim=(1, 2))
        x = F.dropout(x, 0.3, self.training)
        return torch.softmax(self.fc(x), dim=-1)


class _KaimingNormal:
    """VarianceScaling(scale=2, mode='fan_out', distribution='untruncated_normal') (resnet.py:28)."""

    def __init__(self, seed):
        self._gen = torch.Generator().manual_seed(int(seed))

    def __call__(self, shape):
        fan_out = shape[-1] * (math.prod(shape[:-2]) if len(shape) > 2 else 1)
        return torch.randn(tuple(shape), generator=self._gen) * math.sqrt(2.0 / fan_out)


class _BasicBlock(nn.Module):
    def __init__(self, L, mk, cin, planes, stride):
        super().__init__()
        self.c1, self.b1 = mk(planes, (3, 3), stride), _BN(planes, 0.9, 1e-5)
        self.c2, self.b2 = mk(planes, (3, 3), 1), _BN(planes, 0.9, 1e-5)
        self.down = None
        if stride != 1 or cin != planes:
            self.down, self.bd = mk(planes, (1, 1), stride), _BN(planes, 0.9, 1e-5)

    @staticmethod
    def _pad(x, p):
        return F.pad(x, (0, 0, p, p, p, p))          # ZeroPadding2D on NHWC

    def forward(self, x):
        idn = x
        out = torch.relu(self.b1(self.c1(self._pad(x, 1))))
        out = self.b2(self.c2(self._pad(out, 1)))
        if self.down is not None:
            idn = self.bd(self.down(idn))
        return torch.relu(idn + out)


class CifarResNet18(nn.Module):
    """resnet.py:96-140 on CustomConv2DLayerNoBias + a quantised fc: 20 kernels + fc W, b -> 22 tensors."""

    def __init__(self, penalty_threshold: float, seed: int, orientation: str, num_classes: int = 10):
        super().__init__()
        L = _layers("paper")
        init = _KaimingNormal(seed)
        mk = lambda planes, ks, st: L.CustomConv2DLayerNoBias(
            seed=seed, penalty_threshold=penalty_threshold, orientation=orientation, initializer=init,
            filters=planes, kernel_size=ks, strides=st if isinstance(st, tuple) else (st, st), padding="valid",
            name="conv", regularizer=None, trained_weights=None)
        self.conv1, self.bn1 = mk(64, (7, 7), 2), _BN(64, 0.9, 1e-5)
        blocks, cin = [], 64
        for planes, stride in ((64, 1), (64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1)):
            blocks.append(_BasicBlock(L, mk, cin, planes, stride))
            cin = planes
        self.blocks = nn.ModuleList(blocks)
        g = torch.Generator().manual_seed(int(seed))
        bound = 1.0 / math.sqrt(512)
        uni = lambda shape: (torch.rand(tuple(shape), generator=g) * 2 - 1) * bound
        self.fc = L.CustomDenseLayer(seed=seed, units=num_classes, penalty_threshold=penalty_threshold,
                                     orientation=orientation, initializer=uni, name="fc", regularizer=None)

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(F.pad(x, (0, 0, 3, 3, 3, 3)))))
        x = F.max_pool2d(F.pad(x, (0, 0, 1, 1, 1, 1)).permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
        for blk in self.blocks:
            x = blk(x)
        return self.fc(x.mean(dim=(1, 2)))           # logits, as in the reference (resnet.py:126-137)


class SelfCheckNet(nn.Module):
    """A small BatchNorm-free model on the nested layers (two quantised convs of different storage orders and a
    quantised dense layer: interleaved (COL) and per-tensor layouts, an early and a late piece of the gradient exchange):
    without batch statistics, `world` ranks with a shard of a batch each must end a step where one device with
    the whole batch ends it.  Used by bench.py's pre-timing self-check and tests/test_gpu_dp.py."""

    def __init__(self, penalty_threshold: float = 1e-3, seed: int = 9):
        super().__init__()
        from ..nested_quantization_layer.custom_components import custom_layers as nq
        init = RandomNormal(seed=seed)
        mk = lambda f, nm, st: nq.CustomConv2DLayer(seed=seed, penalty_threshold=penalty_threshold, orientation="channelwise",
                                                    initializer=init, filters=f, kernel_size=(3, 3), strides=(1, 1),
                                                    padding="same", name=nm, regularizer=None, kernel_storage=st)
        self.c1, self.c2 = mk(8, "a", "ohwi"), mk(32, "b", "hwio")
        self.fc = nq.CustomDenseLayer(seed=seed, units=10, penalty_threshold=penalty_threshold, orientation="rowwise",
                                      initializer=init, name="fc", regularizer=None)

    def forward(self, x):
        x = torch.relu(self.c2(torch.relu(self.c1(x))))
        return torch.softmax(self.fc(x.mean(dim=(1, 2))), dim=-1)

    def quant_layers(self):
        return (self.c1.nested_q_k_layer, self.c1.nested_q_b_layer, self.c2.nested_q_k_layer, self.c2.nested_q_b_layer,
                self.fc.nested_q_w_layer, self.fc.nested_q_b_layer)


def mnist_mlp(value, seed=42, orientation="rowwise", variant="nested"):
    return MnistMLP(value, seed, orientation, variant)


def imagenette_resnet(value, seed=42, orientation="channelwise", variant="nested", kernel_storage="hwio"):
    return ImagenetteResNet(value, seed, orientation, variant, kernel_storage=kernel_storage)


def cifar_resnet18(penalty_threshold=0.0, seed=42, orientation="scalar"):
    return CifarResNet18(penalty_threshold, seed, orientation)


def synthetic_batch(batch: int, shape: Tuple[int, int, int], device, seed: int = 42):
    """images U[0,255] fp32 NHWC (the reference feeds un-normalised floats), labels randint(0,10)."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.rand((batch, *shape), generator=g, device=device) * 255.0
    y = torch.randint(0, 10, (batch,), generator=g, device=device)
    return x, y

"""tinygrad.nn stand-ins (Conv2d, Linear, GroupNorm, LayerNorm, Embedding) with the attribute names the reference's savers
read (python/save.py). Parameters are drawn from one seeded torch generator (set_seed) so a model build is reproducible."""
from __future__ import annotations

import math

import torch

from ..tensor import Tensor

_GEN = torch.Generator().manual_seed(0)


def set_seed(seed: int) -> None:
    _GEN.manual_seed(seed)


def _uniform(shape, bound, offset=0.0):
    return Tensor((torch.rand(tuple(shape), generator=_GEN, dtype=torch.float32) * 2.0 - 1.0) * bound + offset)


class Conv2d:
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        fan_in = in_channels // groups * self.kernel_size[0] * self.kernel_size[1]
        self.weight = _uniform((out_channels, in_channels // groups, *self.kernel_size), math.sqrt(3.0 / fan_in))
        self.bias = _uniform((out_channels,), 1.0 / math.sqrt(fan_in)) if bias else None

    def __call__(self, x):
        return x.conv2d(self.weight, self.bias, padding=self.padding, stride=self.stride, dilation=self.dilation,
                        groups=self.groups)


class Linear:
    def __init__(self, in_features, out_features, bias=True):
        self.weight = _uniform((out_features, in_features), math.sqrt(3.0 / in_features))
        self.bias = _uniform((out_features,), 1.0 / math.sqrt(in_features)) if bias else None

    def __call__(self, x):
        return x.linear(self.weight.transpose(), self.bias)


class GroupNorm:
    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True):
        self.num_groups, self.num_channels, self.eps = num_groups, num_channels, eps
        self.weight = _uniform((num_channels,), 0.1, 1.0) if affine else None
        self.bias = _uniform((num_channels,), 0.1) if affine else None

    def __call__(self, x):
        # reshape for layernorm to work as group norm; subtract mean and divide stddev
        x = x.reshape(x.shape[0], self.num_groups, -1).layernorm(eps=self.eps).reshape(x.shape)
        if self.weight is None or self.bias is None:
            return x
        ones = [1] * (len(x.shape) - 2)
        return x * self.weight.reshape(1, -1, *ones) + self.bias.reshape(1, -1, *ones)


class LayerNorm:
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        self.normalized_shape = (normalized_shape,) if isinstance(normalized_shape, int) else tuple(normalized_shape)
        self.axis, self.eps = tuple(-1 - i for i in range(len(self.normalized_shape))), eps
        self.weight = _uniform(self.normalized_shape, 0.1, 1.0) if elementwise_affine else None
        self.bias = _uniform(self.normalized_shape, 0.1) if elementwise_affine else None

    def __call__(self, x):
        x = x.layernorm(eps=self.eps, axis=self.axis)
        return x if self.weight is None else x * self.weight + self.bias


class Embedding:
    def __init__(self, vocab_size, embed_size):
        self.vocab_size, self.embed_size = vocab_size, embed_size
        self.weight = _uniform((vocab_size, embed_size), math.sqrt(3.0))

    def __call__(self, idx):
        return Tensor(self.weight.t[idx.t.long()])

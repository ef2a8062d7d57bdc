"""tinygrad.tensor.Tensor stand-in on torch CPU tensors. Each method restates the tinygrad 0.9.2 definition it replaces
(composition order kept where it matters numerically: softmax, layernorm, gelu)."""
from __future__ import annotations

import functools
import math

import numpy as np
import torch
import torch.nn.functional as F


def _unwrap(x):
    return x.t if isinstance(x, Tensor) else x


class Tensor:
    no_grad = False

    def __init__(self, data, dtype=None):
        if isinstance(data, Tensor):
            t = data.t
        elif isinstance(data, torch.Tensor):
            t = data
        elif isinstance(data, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(data))
        else:  # python scalars / lists: ints -> int32, floats -> float32 (tinygrad defaults)
            a = np.asarray(data)
            t = torch.from_numpy(a.astype(np.float32 if a.dtype.kind == "f" else np.int32))
        self.t = t

    # ---- construction
    @staticmethod
    def arange(start, stop=None, step=1):
        if stop is None:
            start, stop = 0, start
        isf = any(isinstance(v, float) for v in (start, stop, step))
        return Tensor(torch.arange(start, stop, step, dtype=torch.float32 if isf else torch.int32))

    @staticmethod
    def full(shape, fill_value):
        return Tensor(torch.full(tuple(shape), fill_value, dtype=torch.float32))

    @staticmethod
    def zeros(*shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        return Tensor(torch.zeros(tuple(shape), dtype=torch.float32))

    @staticmethod
    def ones(*shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        return Tensor(torch.ones(tuple(shape), dtype=torch.float32))

    @staticmethod
    def empty(*shape):
        return Tensor.zeros(*shape)

    # ---- inspection
    @property
    def shape(self):
        return tuple(self.t.shape)

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def realize(self):
        return self

    # ---- movement
    def reshape(self, *shape, **kw):
        if "shape" in kw:
            shape = kw["shape"]
        elif len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return Tensor(self.t.reshape(tuple(shape)))

    def permute(self, *order):
        if len(order) == 1 and isinstance(order[0], (tuple, list)):
            order = order[0]
        return Tensor(self.t.permute(tuple(order)))

    def transpose(self, ax1=1, ax2=0):
        return Tensor(self.t.transpose(ax1, ax2))

    def expand(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return Tensor(self.t.expand(tuple(shape)))

    def cat(self, *args, dim=0):
        return Tensor(torch.cat([self.t] + [_unwrap(a) for a in args], dim=dim))

    def chunk(self, num, dim=0):
        return [Tensor(c) for c in self.t.chunk(num, dim=dim)]

    def triu(self, k=0):
        return Tensor(self.t.triu(k))

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    # ---- elementwise
    def _bin(self, other, fn, reverse=False):
        o = _unwrap(other)
        return Tensor(fn(o, self.t) if reverse else fn(self.t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return Tensor(-self.t)
    def add(self, o): return self + o
    def sub(self, o): return self - o
    def mul(self, o): return self * o
    def div(self, o): return self / o
    def exp(self): return Tensor(self.t.exp())
    def cos(self): return Tensor(self.t.cos())
    def sin(self): return Tensor(self.t.sin())
    def tanh(self): return Tensor(self.t.tanh())
    def sigmoid(self): return Tensor(torch.sigmoid(self.t))
    def rsqrt(self): return Tensor(self.t.rsqrt())
    def sqrt(self): return Tensor(self.t.sqrt())

    def swish(self):  # x * x.sigmoid()
        return self * self.sigmoid()

    def silu(self):
        return self.swish()

    def gelu(self):  # tinygrad 0.9.2: the tanh approximation (NOT the erf form burn uses)
        return 0.5 * self * (1 + (self * 0.7978845608 * (1 + 0.044715 * self * self)).tanh())

    def quick_gelu(self):
        return self * (self * 1.702).sigmoid()

    # ---- reductions
    def mean(self, axis=None, keepdim=False):
        return Tensor(self.t.mean() if axis is None else self.t.mean(dim=axis, keepdim=keepdim))

    def sum(self, axis=None, keepdim=False):
        return Tensor(self.t.sum() if axis is None else self.t.sum(dim=axis, keepdim=keepdim))

    def max(self, axis=None, keepdim=False):
        return Tensor(self.t.max() if axis is None else self.t.amax(dim=axis, keepdim=keepdim))

    def softmax(self, axis=-1):  # m = x - max; e = exp(m); e / sum(e)
        m = self - self.max(axis=axis, keepdim=True)
        e = m.exp()
        return e.div(e.sum(axis=axis, keepdim=True))

    def layernorm(self, axis=-1, eps=1e-5):  # y = x - mean; y * rsqrt(mean(y*y) + eps)
        y = self - self.mean(axis, keepdim=True)
        return y.mul((y * y).mean(axis, keepdim=True).add(eps).rsqrt())

    # ---- contractions
    def dot(self, w):
        return Tensor(torch.matmul(self.t, _unwrap(w)))

    def matmul(self, w):
        return self.dot(w)

    def __matmul__(self, w):
        return self.dot(w)

    def linear(self, weight, bias=None):
        x = self.dot(weight)
        return x + bias if bias is not None else x

    def conv2d(self, weight, bias=None, groups=1, stride=1, dilation=1, padding=0):
        # tinygrad: int -> all sides; 2-tuple (ph, pw) -> [pw, pw, ph, ph]; 4-tuple = (left, right, top, bottom) as given
        if isinstance(padding, int):
            pad = [padding] * 4
        elif len(padding) == 4:
            pad = list(padding)
        else:
            pad = [p for p in padding for _ in range(2)][::-1]
        x = F.pad(self.t, pad)
        b = None if bias is None else _unwrap(bias)
        return Tensor(F.conv2d(x, _unwrap(weight), b, stride=stride, padding=0, dilation=dilation, groups=groups))

    def sequential(self, ll):
        return functools.reduce(lambda x, f: f(x), ll, self)

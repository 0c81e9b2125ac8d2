"""theano.tensor subset (see ../../README.md)."""
import numpy as np
import torch

import theano
from theano import Variable, as_variable, TensorType  # noqa: F401
from theano.tensor import nnet, sharedvar, shared_randomstreams  # noqa: F401


def _sym(ndim, default_dtype):
    def make(name=None, dtype=None):
        if name is not None and dtype is None and name in theano._TORCH_DTYPES:
            name, dtype = None, name
        return Variable(None, (), ndim=ndim, dtype=str(dtype or default_dtype), name=name or 'input')
    return make


scalar, vector, matrix, tensor3, tensor4 = (_sym(n, 'float32') for n in range(5))
iscalar, ivector, imatrix = (_sym(n, 'int32') for n in range(3))
fscalar, fvector, fmatrix = (_sym(n, 'float32') for n in range(3))


def constant(value, **kw):
    return as_variable(value)


def _unary(fn):
    def apply(x):
        x = as_variable(x)
        return Variable(fn, (x,), ndim=x.ndim, dtype=x.dtype)
    return apply


log = _unary(torch.log)
exp = _unary(torch.exp)
sqrt = _unary(torch.sqrt)
tanh = _unary(torch.tanh)
sqr = _unary(lambda a: a * a)
abs_ = _unary(torch.abs)


def clip(x, lo, hi):
    """T.clip: gradient 1 on the closed interval [lo, hi], 0 outside (Theano Clip.grad)."""
    x = as_variable(x)
    lo_f, hi_f = float(np.float32(lo)), float(np.float32(hi))      # python floats become floatX constants
    return Variable(lambda a: torch.clamp(a, lo_f, hi_f), (x,), ndim=x.ndim, dtype=x.dtype)


def sum(x, axis=None, keepdims=False, **kw):  # noqa: A001
    return as_variable(x).sum(axis=axis, keepdims=keepdims)


def mean(x, axis=None, **kw):
    return as_variable(x).mean(axis=axis)


def dot(a, b):
    a, b = as_variable(a), as_variable(b)
    return Variable(lambda x, y: x @ y, (a, b), ndim=a.ndim + b.ndim - 2, dtype='float32')


def take(a, indices, axis=None, mode=None):
    a, indices = as_variable(a), as_variable(indices)
    assert axis == 0
    return Variable(lambda x, i: x[i.long()], (a, indices), ndim=a.ndim - 1 + indices.ndim, dtype=a.dtype)


def and_(a, b):
    return as_variable(a)._bin(b, lambda x, y: x & y)


def switch(c, a, b):
    c, a, b = as_variable(c), as_variable(a), as_variable(b)
    return Variable(lambda x, y, z: torch.where(x.bool(), y, z), (c, a, b), ndim=max(a.ndim, b.ndim), dtype=a.dtype)

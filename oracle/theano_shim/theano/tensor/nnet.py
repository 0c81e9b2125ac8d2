import torch

from theano import Variable, as_variable


def softmax(x):
    """Row softmax of a matrix (max-subtracted, like Theano's Softmax op)."""
    x = as_variable(x)
    assert x.ndim == 2
    return Variable(lambda a: torch.softmax(a, dim=1), (x,), ndim=2, dtype=x.dtype)


def sigmoid(x):
    x = as_variable(x)
    return Variable(torch.sigmoid, (x,), ndim=x.ndim, dtype=x.dtype)


def categorical_crossentropy(coding_dist, true_dist):
    coding_dist, true_dist = as_variable(coding_dist), as_variable(true_dist)
    if true_dist.ndim == coding_dist.ndim:
        return -(true_dist * Variable(torch.log, (coding_dist,), ndim=coding_dist.ndim, dtype=coding_dist.dtype)) \
            .sum(axis=coding_dist.ndim - 1)
    raise NotImplementedError('integer targets are not used by sert/models.py')

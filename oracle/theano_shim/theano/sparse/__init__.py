import numpy as np
import torch

from theano import Variable
from theano.sparse import type  # noqa: F401,A004


def dense_from_sparse(x):
    return Variable(lambda a: torch.from_numpy(np.asarray(a.todense(), dtype=np.float32)), (x,), ndim=2, dtype='float32')

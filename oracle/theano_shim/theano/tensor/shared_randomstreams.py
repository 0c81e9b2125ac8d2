import numpy as np
import torch

from theano import Variable, as_variable

# every draw of every stream, in call order: (stream id, int64 array).  tests/golden/make_golden.py reads this
# so the same negatives can be fed to the oracle and to the CUDA path.
DRAWS = []
_STREAMS = []


class RandomStreams(object):
    def __init__(self, seed=None):
        self.rng = np.random.RandomState(seed)
        self.stream_id = len(_STREAMS)
        _STREAMS.append(self)

    def choice(self, size=1, a=2, replace=True, p=None, **kw):
        p = as_variable(p)
        size = tuple(int(s) for s in size)
        stream = self

        def op(pv):
            probs = pv.detach().numpy().astype(np.float64)
            probs = probs / probs.sum()
            out = stream.rng.choice(int(a), size=size, replace=True, p=probs).astype(np.int64)
            DRAWS.append((stream.stream_id, out.copy()))
            return torch.from_numpy(out)
        return Variable(op, (p,), ndim=len(size), dtype='int64', name='random_choice')

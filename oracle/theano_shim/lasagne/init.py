import numpy as np


class Initializer(object):
    def __call__(self, shape):
        return self.sample(shape)


class GlorotUniform(Initializer):
    """lasagne.init.GlorotUniform (gain 1): U(-a, a), a = sqrt(6 / (fan_in + fan_out)), from np.random."""

    def sample(self, shape):
        n1, n2 = shape[:2]
        receptive = int(np.prod(shape[2:]))
        a = np.sqrt(6.0 / ((n1 + n2) * receptive))
        return np.random.uniform(low=-a, high=a, size=shape).astype(np.float32)


class Constant(Initializer):
    def __init__(self, val=0.0):
        self.val = val

    def sample(self, shape):
        return (np.ones(shape) * self.val).astype(np.float32)

"""Normalised Shannon entropy (drop-in for the reference's sert/math_utils.py:5-25)."""
import numpy as np
import scipy.stats


def entropy(pk, *args, normalize=False, **kwargs):
    """scipy.stats.entropy(pk, ...); normalize=True divides by the entropy of the uniform distribution over
    np.size(pk) outcomes, log(n) in the requested base, so the result lies in [0, 1]."""
    value = scipy.stats.entropy(pk, *args, **kwargs)
    if not normalize:
        return value
    uniform = np.log(np.size(pk))          # numpy's log: the debug output is compared bit for bit with the reference's
    if kwargs.get('base'):
        uniform /= np.log(kwargs['base'])
    return value / uniform

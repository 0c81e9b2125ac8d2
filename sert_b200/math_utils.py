"""Normalised Shannon entropy (drop-in for the reference's sert/math_utils.py:5-25)."""
import numpy as np
import scipy.stats


def entropy(pk, *args, **kwargs):
    """scipy.stats.entropy with an optional normalize=True dividing by the maximum entropy log(n)."""
    normalize = kwargs.pop('normalize', False)

    e = scipy.stats.entropy(pk, *args, **kwargs)

    if normalize:
        maximum_entropy = np.log(np.size(pk))
        base = kwargs.get('base')
        if base:
            maximum_entropy /= np.log(base)

        e /= maximum_entropy

    return e

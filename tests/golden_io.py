"""Loaders for the committed golden fixtures (generated from the reference's own code by
tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import scipy.sparse as sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def csr(d, prefix):
    return sparse.csr_matrix((d[prefix + '_data'], d[prefix + '_indices'], d[prefix + '_indptr']),
                             shape=tuple(int(v) for v in d[prefix + '_shape']))


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def loglinear():
    d = load_npz('loglinear_ref.npz')
    d['training_set'] = (d['x_train'], csr(d, 'y_train'), d['w_train'])
    d['validation_set'] = (d['x_val'], csr(d, 'y_val'))
    return d


def vectorspace():
    d = load_npz('vectorspace_ref.npz')
    d['training_set'] = (d['x_train'], d['y_train'], d['w_train'])
    d['validation_set'] = (d['x_val'], d['y_val'])
    return d

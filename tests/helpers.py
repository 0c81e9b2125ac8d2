"""Shared builders for the parity tests: the same seeded inputs go to the CUDA path and the oracle."""
import numpy as np

from oracle import sert_oracle as O
from sert_b200 import synth


def close(got, ref, rtol=1e-4, atol_scale=1e-5, what=''):
    """|got-ref| <= rtol*|ref| + atol_scale*max|ref| (1e-4 relative fp32, BASELINE.json north_star)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    err = np.abs(got - ref)
    tol = rtol * np.abs(ref) + atol_scale * scale + 1e-30
    bad = err > tol
    assert not bad.any(), '%s: %d/%d out of tolerance, max err %.3e (scale %.3e)' % (
        what, int(bad.sum()), ref.size, float(err.max()), scale)


def vs_problem(seed, V, E, dw, de, W, B, k, n_batches, n_val_batches=1, gain=1.0, weights=False):
    rng = np.random.default_rng(seed)
    train, val = synth.vectorspace_corpus(seed, V, E, W, B * n_batches + 3, B * n_val_batches)
    if weights:
        train = (train[0], train[1], synth.make_weights(rng, train[0].shape[0]))
    R = synth.glorot(rng, (V, dw)) * np.float32(gain)
    Eemb = synth.glorot(rng, (E, de)) * np.float32(gain)
    Wp = synth.glorot(rng, (dw, de)) * np.float32(gain)
    bp = (rng.standard_normal(de) * 0.01).astype(np.float32)
    neg = rng.integers(0, E, size=(max(n_batches, n_val_batches), B, k)).astype(np.int32)
    return dict(train=train, val=val, R=R, Eemb=Eemb, Wp=Wp, bp=bp, neg=neg,
                V=V, E=E, dw=dw, de=de, W=W, B=B, k=k, n_batches=n_batches)


def ll_problem(seed, V, E, dw, W, B, n_batches, n_val_batches=1, gain=1.0):
    rng = np.random.default_rng(seed)
    train, val = synth.loglinear_corpus(seed, V, E, W, B * n_batches + 5, B * n_val_batches)
    R = synth.glorot(rng, (V, dw)) * np.float32(gain)
    Wd = synth.glorot(rng, (dw, E)) * np.float32(gain)
    bd = (rng.standard_normal(E) * 0.01).astype(np.float32)
    return dict(train=train, val=val, R=R, Wd=Wd, bd=bd, V=V, E=E, dw=dw, W=W, B=B, n_batches=n_batches)


def vs_oracle(p, lam):
    return O.VectorSpaceOracle(p['B'], p['R'], p['Wp'], p['bp'], p['Eemb'], lam, p['train'], p['val'])


def ll_oracle(p, lam):
    return O.LogLinearOracle(p['B'], p['R'], p['Wd'], p['bd'], lam, p['train'], p['val'])

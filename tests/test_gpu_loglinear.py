"""Parity of the CUDA log-linear path (through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def make_model(p, lam, **kw):
    from sert_b200 import models
    return models.LanguageModel(
        batch_size=p['B'], window_size=p['W'], representations_init=p['R'],
        output_layer_size=p['E'], regularization_lambda=lam,
        training_set=p['train'], validation_set=p['val'], dense_init=(p['Wd'], p['bd']), **kw)


def forward_host(model, split, b):
    from sert_b200 import _native as N
    nat = model._native
    B, W, E = nat.cfg.batch, nat.cfg.window, nat.cfg.entities
    z = np.empty((B * W, E), np.float32)
    s = np.empty((B, E), np.float32)
    ell = np.empty(B, np.float32)
    N.check(nat.lib.sert_ll_forward_host(nat.handle, split, b, N.host_ptr(z), N.host_ptr(s), N.host_ptr(ell)))
    return z, s, ell


@pytest.mark.parametrize('dims,gain', [
    (dict(V=5000, E=200, dw=64, W=10, B=64), 1.0),          # BASELINE.json configs[0] shapes, small batch
    (dict(V=700, E=715, dw=300, W=10, B=32), 1.0),          # W3C shapes: dw=300, E=715 (not a multiple of 4)
    (dict(V=300, E=37, dw=16, W=3, B=16), 60.0),            # clipped regime: p < 1e-7 on most entries
    (dict(V=70000, E=1500, dw=32, W=2, B=8), 8.0),
])
def test_forward_logits_match_oracle(dims, gain):
    from oracle import sert_oracle as O
    p = H.ll_problem(13, n_batches=2, gain=gain, **dims)
    model = make_model(p, 0.01)
    B = p['B']
    for b in range(2):
        z, s, ell = forward_host(model, 0, b)
        sl = slice(b * B, (b + 1) * B)
        f = O.loglinear_forward(p['R'], p['Wd'], p['bd'], p['train'][0][sl])
        H.close(z, f['z'].reshape(z.shape), what='per-word logits z')
        H.close(s, f['s'], what='joint logits s')
        ref_ell = O.loglinear_instance_losses(f['o'], O.dense_rows(p['train'][1], sl.start, sl.stop))
        H.close(ell, ref_ell, what='instance loss')


@pytest.mark.parametrize('gain', [1.0, 25.0])
def test_training_steps_match_oracle(gain):
    """4 Adadelta steps (general clipped backward when gain=25) + eval losses."""
    p = H.ll_problem(17, V=900, E=120, dw=32, W=4, B=64, n_batches=4, gain=gain)
    lam = 0.01
    model = make_model(p, lam)
    oracle = H.ll_oracle(p, lam)
    H.close(model.test_fn(1), oracle.eval_batch('train', 1), what='initial eval loss')
    H.close(model.validate_fn(0), oracle.eval_batch('val', 0), what='initial validation loss')
    for j, b in enumerate([2, 0, 3, 1]):
        H.close(model.train_fn(b), oracle.train_batch(b), what='train loss step %d' % j)
    Wd, bd = model.get_dense()
    H.close(model.get_representations(), oracle.R, rtol=2e-4, what='R')
    H.close(Wd, oracle.Wd, rtol=2e-4, what='Wd')
    H.close(bd, oracle.bd, rtol=2e-4, atol_scale=1e-4, what='bd')
    from sert_b200 import _native as N
    accu = model._native.get_tensor(N.PARAM_DENSE_W, oracle.Wd.shape, N.STATE_S1)
    delta = model._native.get_tensor(N.PARAM_WORD_REPR, oracle.R.shape, N.STATE_S2)
    H.close(accu, oracle.state['Wd'][0], rtol=5e-4, atol_scale=1e-4, what='Adadelta accu (Wd)')
    H.close(delta, oracle.state['R'][1], rtol=5e-4, atol_scale=1e-4, what='Adadelta delta (R)')
    H.close(model.test_fn(3), oracle.eval_batch('train', 3), rtol=2e-4, what='eval loss after training')


def test_epoch_protocol_and_host_batches():
    from sert_b200 import _native as N
    p = H.ll_problem(23, V=400, E=64, dw=16, W=3, B=32, n_batches=3)
    a, b = make_model(p, 0.01), make_model(p, 0.01)
    order = [1, 2, 0]
    n, mean = a.train(order=order)
    losses = [b.train_fn(bi) for bi in order]
    assert n == 3
    np.testing.assert_allclose(mean, np.mean(losses), rtol=1e-6)
    mean_e, std_e = a.train_error()
    ref = [b.test_fn(i) for i in range(3)]
    np.testing.assert_allclose([mean_e, std_e], [np.mean(ref), np.std(ref)], rtol=1e-5)
    # streamed host batch == resident batch
    c = make_model(p, 0.01)
    nat = c._native
    x, y, w = p['train']
    y = y.tocsr()
    indptr = np.ascontiguousarray(y.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(y.indices, dtype=np.int32)
    data = np.ascontiguousarray(y.data, dtype=np.float32)
    for j, bi in enumerate(order):
        sl = slice(bi * 32, (bi + 1) * 32)
        xb = np.ascontiguousarray(x[sl], dtype=np.int32)
        wb = np.ascontiguousarray(w[sl], dtype=np.float32)
        ip = np.ascontiguousarray(indptr[sl.start:sl.stop + 1])
        out = np.zeros(1, np.float32)
        N.check(nat.lib.sert_train_batch_host(nat.handle, N.host_ptr(xb), None, N.host_ptr(ip), N.host_ptr(indices),
                                              N.host_ptr(data), N.host_ptr(wb), None, N.host_ptr(out)))
        np.testing.assert_allclose(out[0], losses[j], rtol=1e-5)


def test_predict_fn_matches_oracle_and_pickles():
    import pickle
    from oracle import sert_oracle as O
    p = H.ll_problem(29, V=500, E=90, dw=24, W=5, B=16, n_batches=1)
    model = make_model(p, 0.01)
    state = model.get_state()
    assert len(state) == 2 and state[1].shape == p['R'].shape
    fn = pickle.loads(pickle.dumps(state[0]))
    rng = np.random.default_rng(1)
    batch = rng.integers(0, 500, size=(16, 5)).astype(np.uint16)
    mask = np.ones((16, 5), np.int8)
    got = fn(batch, mask)
    ref = O.loglinear_predict(p['R'], p['Wd'], p['bd'], batch)
    assert got.shape == (16, 5, 90) and got.dtype == np.float32
    H.close(got, ref, what='predict_fn distributions')
    np.testing.assert_allclose(got.sum(axis=2), 1.0, atol=1e-5)
    # fewer rows than the batch size (the batcher always sends full batches, but the ABI allows less)
    H.close(fn(batch[:3], mask[:3]), ref[:3], what='partial batch')


@pytest.mark.parametrize('dims,tensor', [
    (dict(V=3000, E=300, dw=64, W=10, B=512), 1),      # Z and dX on tcgen05 (40 tiles each), dWd on FMA tiles
    (dict(V=2000, E=4096, dw=256, W=4, B=64), 1),      # Z and dWd on tcgen05 (32 tiles each), dX on FMA tiles
    (dict(V=3000, E=300, dw=64, W=10, B=512), 0),      # same shapes, tensor cores off
    (dict(V=3000, E=8200, dw=64, W=8, B=512), 1),      # all three on tcgen05: fused backward tail (dZ written once, as
                                                       # the split operands; bias gradient through the ones row)
    (dict(V=20000, E=8200, dw=300, W=10, B=256), 1),   # configs[4]'s row width over a ragged entity axis: pair operands
                                                       # over 5 K blocks (projection, 20 x 33 tiles in clusters), 129
                                                       # (dX: split-K, second n-tile 44 columns wide) and 40 (gWd: dZ's
                                                       # rows as the N-major B operand, 301 rows with the ones row)
])
def test_tensor_core_projection_matches_oracle(dims, tensor):
    """The word x entity GEMMs through the tcgen05 bf16x3 path: logits within 1e-4 relative, 3 Adadelta steps."""
    from oracle import sert_oracle as O
    from sert_b200 import _native as N
    p = H.ll_problem(41, n_batches=3, gain=2.0, **dims)
    lam = 0.01
    model = make_model(p, lam)
    N.check(model._native.lib.sert_model_set_tensor_cores(model._native.handle, tensor))
    oracle = H.ll_oracle(p, lam)
    z, s, ell = forward_host(model, 0, 1)
    B = p['B']
    f = O.loglinear_forward(p['R'], p['Wd'], p['bd'], p['train'][0][B:2 * B])
    H.close(z, f['z'].reshape(z.shape), what='per-word logits z')
    H.close(s, f['s'], what='joint logits s')
    for j, b in enumerate([2, 0, 1]):
        H.close(model.train_fn(b), oracle.train_batch(b), what='train loss step %d' % j)
    Wd, bd = model.get_dense()
    H.close(model.get_representations(), oracle.R, rtol=2e-4, what='R')
    H.close(Wd, oracle.Wd, rtol=2e-4, what='Wd')
    H.close(bd, oracle.bd, rtol=2e-4, atol_scale=1e-4, what='bd')


def test_empty_validation_set_and_tail_drop():
    """0 validation instances: constructor reshapes x to (0, W) (sert/models.py:448-454); validation_error is
    (nan, nan) as in the reference (np.mean of no batches); an incomplete last training batch is ignored."""
    import warnings
    import scipy.sparse as sp
    p = H.ll_problem(51, V=200, E=30, dw=16, W=3, B=16, n_batches=2)
    x, y, w = p['train']
    empty = (np.zeros((0,), dtype=x.dtype), sp.csr_matrix((0, 30), dtype=np.float32))
    p2 = dict(p, val=empty)
    model = make_model(p2, 0.01)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        mean, std = model.validation_error()
    assert np.isnan(mean) and np.isnan(std)
    n, loss = model.train(order=[0, 1])
    assert n == 2 and np.isfinite(loss)          # 2*16+5 instances -> 2 batches, 5 dropped
    with pytest.raises(RuntimeError, match='out of range'):
        model.train_fn(2)


@pytest.mark.parametrize('gain', [25.0, 60.0])
def test_fused_backward_tail_in_the_clipped_regime(gain):
    """All three word x entity GEMMs on tensor cores with logits driven into the clips (log-domain clip masks of
    ll_joint / ll_racc_log / ll_dz_split against the oracle's p-domain clips)."""
    p = H.ll_problem(43, n_batches=2, gain=gain, V=1500, E=8192, dw=64, W=8, B=512)
    model = make_model(p, 0.01)
    oracle = H.ll_oracle(p, 0.01)
    for j in range(2):
        H.close(model.train_fn(j), oracle.train_batch(j), what='train loss step %d' % j)
    Wd, bd = model.get_dense()
    H.close(model.get_representations(), oracle.R, rtol=2e-4, what='R')
    H.close(Wd, oracle.Wd, rtol=2e-4, what='Wd')
    H.close(bd, oracle.bd, rtol=2e-4, atol_scale=1e-4, what='bd')

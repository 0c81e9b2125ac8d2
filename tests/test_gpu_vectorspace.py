"""Parity of the CUDA vector-space path (through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def make_model(p, lam, **kw):
    from sert_b200 import models
    return models.VectorSpaceLanguageModel(
        batch_size=p['B'], window_size=p['W'], num_negative_samples=p['k'],
        representations_init=p['R'], entity_representations_init=p['Eemb'],
        regularization_lambda=lam, training_set=p['train'], validation_set=p['val'],
        dense_init=(p['Wp'], p['bp']), **kw)


def forward_host(model, split, b, neg):
    import torch
    from sert_b200 import _native as N
    nat = model._native
    B, k, de = nat.cfg.batch, nat.cfg.num_negatives, nat.cfg.entity_dim
    scores = np.empty((B, k + 1), np.float32)
    proj = np.empty((B, de), np.float32)
    ell = np.empty(B, np.float32)
    negt = torch.from_numpy(np.ascontiguousarray(neg, dtype=np.int32)).to(nat.device)
    N.check(nat.lib.sert_vs_forward_host(nat.handle, split, b, N.dev_ptr(negt), N.host_ptr(scores),
                                         N.host_ptr(proj), N.host_ptr(ell)))
    return scores, proj, ell


@pytest.mark.parametrize('dims', [
    dict(V=500, E=200, dw=64, de=32, W=10, B=128, k=10),
    dict(V=3000, E=777, dw=300, de=128, W=4, B=256, k=10),     # product-search.sh shapes (dw=300, de=128, window 4)
    dict(V=1200, E=50, dw=128, de=256, W=3, B=64, k=3),
    dict(V=70000, E=2000, dw=8, de=520, W=33, B=32, k=37),     # uint32 indices, window > 32, ragged de
])
def test_forward_logits_match_oracle(dims):
    from oracle import sert_oracle as O
    p = H.vs_problem(11, n_batches=2, **dims)
    model = make_model(p, 0.01)
    for b in range(2):
        scores, proj, ell = forward_host(model, 0, b, p['neg'][b])
        sl = slice(b * p['B'], (b + 1) * p['B'])
        f = O.vectorspace_forward(p['R'], p['Wp'], p['bp'], p['Eemb'], p['train'][0][sl], p['train'][1][sl], p['neg'][b])
        ref = np.concatenate([f['score_pos'][:, None], f['score_neg']], axis=1)
        H.close(scores, ref, what='logits')          # logits within 1e-4 relative fp32
        H.close(proj, f['u'], what='projection')
        H.close(ell, f['ell'], what='instance loss')


@pytest.mark.parametrize('fused,overlap', [(1, 1), (1, 0), (2, 1), (0, 1), (0, 0)])
@pytest.mark.parametrize('gain,weights,dims', [
    (1.0, False, dict(dw=64, de=48, W=5, B=128, k=6)),
    (1.0, True, dict(dw=128, de=128, W=10, B=163, k=10)),        # BASELINE configs[1] dims, ragged last tile
    (30.0, True, dict(dw=128, de=128, W=7, B=72, k=15)),         # tile kernel into the clips, 16 scores per instance
    (1.0, False, dict(dw=128, de=128, W=32, B=40, k=1)),         # tile kernel, widest window / fewest negatives
    (1.0, True, dict(dw=128, de=128, W=3, B=64, k=16)),          # 17 scores: falls back to the warp kernel
    (40.0, True, dict(dw=64, de=48, W=5, B=128, k=6)),
    (1.0, True, dict(dw=300, de=128, W=4, B=67, k=10)),          # product-search dims: tile kernel, 3 chunks per word row
    (30.0, True, dict(dw=300, de=128, W=4, B=64, k=10)),         # ... into the clips
    (1.0, False, dict(dw=200, de=128, W=6, B=40, k=5)),          # tile kernel, 2 chunks, ragged second chunk
    (1.0, True, dict(dw=64, de=128, W=5, B=48, k=7)),            # tile kernel, word rows shorter than a chunk
    (3.0, True, dict(dw=32, de=32, W=40, B=96, k=3)),            # window > 32
])
def test_training_steps_match_oracle(gain, weights, dims, fused, overlap):
    """5 Adam steps + eval losses through the fused kernels (fused=1: tile kernel where d_e = 128 and d_w <= 384, else
    the warp kernel; fused=2: warp kernel) and the per-stage kernels (fused=0); gain >= 30 drives tanh / sigmoid into the
    1e-7 clips."""
    from sert_b200 import _native as N
    p = H.vs_problem(5, V=800, E=300, n_batches=5, gain=gain, weights=weights, **dims)
    lam = 0.01
    model = make_model(p, lam)
    N.check(model._native.lib.sert_model_set_fused(model._native.handle, fused))
    N.check(model._native.lib.sert_model_set_overlap(model._native.handle, overlap))
    oracle = H.vs_oracle(p, lam)
    order = [3, 0, 4, 1, 2]
    e0 = model.test_fn(1, p['neg'][1])
    H.close(e0, oracle.eval_batch('train', 1, p['neg'][1]), what='initial eval loss')
    v0 = model.validate_fn(0, p['neg'][0])
    H.close(v0, oracle.eval_batch('val', 0, p['neg'][0]), what='initial validation loss')
    for j, b in enumerate(order):
        got = model.train_fn(b, p['neg'][j])
        ref = oracle.train_batch(b, p['neg'][j])
        H.close(got, ref, what='train loss step %d' % j)
    R, Eemb = model.get_representations()
    Wp, bp = model.get_dense()
    H.close(R, oracle.R, rtol=2e-4, what='R')
    H.close(Eemb, oracle.Eemb, rtol=2e-4, what='Eemb')
    H.close(Wp, oracle.Wp, rtol=2e-4, what='Wp')
    H.close(bp, oracle.bp, rtol=2e-4, atol_scale=1e-4, what='bp')
    from sert_b200 import _native as N
    m = model._native.get_tensor(N.PARAM_ENTITY_REPR, oracle.Eemb.shape, N.STATE_S1)
    v = model._native.get_tensor(N.PARAM_WORD_REPR, oracle.R.shape, N.STATE_S2)
    H.close(m, oracle.state['Eemb'][0], rtol=2e-4, atol_scale=1e-4, what='Adam m (Eemb)')
    H.close(v, oracle.state['R'][1], rtol=2e-4, atol_scale=1e-4, what='Adam v (R)')
    H.close(model.test_fn(2, p['neg'][2]), oracle.eval_batch('train', 2, p['neg'][2]), rtol=2e-4,
            what='eval loss after training')


@pytest.mark.parametrize('hot', ['auto', 'none', 'forced'])
def test_hot_word_rows_are_result_neutral(hot):
    """The tile kernel spreads the gradient additions of very frequent word ids over private copies
    (sert_model_set_hot_words); whatever the hot set, the step must match the oracle."""
    p = H.vs_problem(17, V=300, E=200, dw=128, de=128, W=10, B=200, k=10, n_batches=4, weights=True)
    model = make_model(p, 0.01)
    if hot == 'auto':
        assert 1 <= model.hot_words.size <= 32          # V=300 Zipf: the top words exceed 64 occurrences per batch
    elif hot == 'none':
        model.set_hot_words([])
    else:
        model.set_hot_words(np.r_[np.arange(0, 300, 10), 7][:32])   # 31 ids, frequent and rare ones
    oracle = H.vs_oracle(p, 0.01)
    for j, b in enumerate([2, 0, 3, 1]):
        H.close(model.train_fn(b, p['neg'][j]), oracle.train_batch(b, p['neg'][j]), what='train loss step %d' % j)
    R, Eemb = model.get_representations()
    H.close(R, oracle.R, rtol=2e-4, what='R')
    H.close(Eemb, oracle.Eemb, rtol=2e-4, what='Eemb')
    with pytest.raises(RuntimeError, match='duplicate hot word'):
        model.set_hot_words([3, 3])
    with pytest.raises(RuntimeError, match='out of range'):
        model.set_hot_words([300])


def test_switching_step_variants_mid_training_keeps_parity():
    """The lazy tile step leaves a loss pending and marks the hot word rows; switching to the warp kernel, the per-stage
    kernels, the single-stream step and back must flush / unmark them: every loss and the final tables match the oracle."""
    from sert_b200 import _native as N
    p = H.vs_problem(41, V=400, E=150, dw=128, de=128, W=6, B=200, k=7, n_batches=10, weights=True)
    model = make_model(p, 0.01)
    assert model.hot_words.size >= 1
    oracle = H.vs_oracle(p, 0.01)
    lib, h = model._native.lib, model._native.handle
    plan = [(1, 1), (1, 1), (2, 1), (0, 1), (1, 1), (1, 0), (1, 1), (2, 0), (1, 1), (1, 1)]     # (fused, overlap)
    for j, (fused, overlap) in enumerate(plan):
        N.check(lib.sert_model_set_fused(h, fused))
        N.check(lib.sert_model_set_overlap(h, overlap))
        H.close(model.train_fn(j, p['neg'][j]), oracle.train_batch(j, p['neg'][j]), what='train loss step %d' % j)
        if j == 4:      # an eval between lazy steps shares the accumulators
            H.close(model.test_fn(0, p['neg'][0]), oracle.eval_batch('train', 0, p['neg'][0]), rtol=2e-4, what='eval')
    R, Eemb = model.get_representations()
    Wp, bp = model.get_dense()
    H.close(R, oracle.R, rtol=2e-4, what='R')
    H.close(Eemb, oracle.Eemb, rtol=2e-4, what='Eemb')
    H.close(Wp, oracle.Wp, rtol=2e-4, what='Wp')
    H.close(bp, oracle.bp, rtol=2e-4, atol_scale=1e-4, what='bp')


def test_epoch_api_and_host_batches():
    """train() over a whole epoch in one call == per-batch train_fn; streamed host batches == resident data."""
    from sert_b200 import _native as N
    p = H.vs_problem(9, V=600, E=150, dw=32, de=32, W=4, B=64, k=5, n_batches=4)
    a, b = make_model(p, 0.01), make_model(p, 0.01)
    order = [2, 0, 3, 1]
    n, mean = a.train(order=order, negatives=p['neg'][:4])
    assert n == 4
    losses = [b.train_fn(bi, p['neg'][j]) for j, bi in enumerate(order)]
    np.testing.assert_allclose(mean, np.mean(losses), rtol=1e-6)
    c = make_model(p, 0.01)
    nat = c._native
    x, y, w = p['train']
    for j, bi in enumerate(order):
        sl = slice(bi * 64, (bi + 1) * 64)
        xb = np.ascontiguousarray(x[sl], dtype=np.int32)
        yb = np.ascontiguousarray(y[sl], dtype=np.int32)
        wb = np.ascontiguousarray(w[sl], dtype=np.float32)
        nb = np.ascontiguousarray(p['neg'][j], dtype=np.int32)
        out = np.zeros(1, np.float32)
        N.check(nat.lib.sert_train_batch_host(nat.handle, N.host_ptr(xb), N.host_ptr(yb), None, None, None,
                                              N.host_ptr(wb), N.host_ptr(nb), N.host_ptr(out)))
        np.testing.assert_allclose(out[0], losses[j], rtol=1e-5)
    # errors over epochs: (mean, std) protocol, tail dropped
    mean_e, std_e = a.train_error(negatives=p['neg'][:4])
    assert np.isfinite(mean_e) and std_e >= 0


@pytest.mark.parametrize('dims', [dict(dw=128, de=128, W=10, B=96, k=10), dict(dw=32, de=48, W=4, B=64, k=5)])
def test_pipelined_host_batches_match_resident_training(dims):
    """sert_train_batch_host_async / sert_train_host_wait (train_stream): same losses and parameters as train_fn on the
    resident data set, for the lazy tile-kernel step and for the per-stage step; tickets waited late and in bulk."""
    from sert_b200 import _native as N
    p = H.vs_problem(23, V=700, E=180, n_batches=11, weights=True, **dims)
    B = p['B']
    a, b = make_model(p, 0.01), make_model(p, 0.01)
    order = [4, 0, 9, 2, 7, 1, 10, 3, 8, 5, 6]
    ref = [a.train_fn(bi, p['neg'][j]) for j, bi in enumerate(order)]
    x, y, w = p['train']

    def batches(lo, hi):
        for j in range(lo, hi):
            sl = slice(order[j] * B, (order[j] + 1) * B)
            yield x[sl], y[sl], w[sl], p['neg'][j]

    got = list(b.train_stream(batches(0, 5)))
    # resident-data steps in between share the accumulators and the pending loss with the pipelined ones
    got.append(b.train_fn(order[5], p['neg'][5]))
    got += list(b.train_stream(batches(6, 11), depth=4))
    np.testing.assert_allclose(got, ref, rtol=1e-5)
    Ra, Ea = a.get_representations()
    Rb, Eb = b.get_representations()
    # float atomics: the summation order of a gradient row differs from run to run, and Adam's m / sqrt(v) turns a
    # last-bit difference of a nearly cancelling sum into a visible one -- same tolerance as against the oracle
    H.close(Rb, Ra, rtol=2e-4, what='R, pipelined vs resident')
    H.close(Eb, Ea, rtol=2e-4, what='Eemb, pipelined vs resident')
    oracle = H.vs_oracle(p, 0.01)
    for j, bi in enumerate(order):
        oracle.train_batch(bi, p['neg'][j])
    H.close(Rb, oracle.R, rtol=2e-4, what='R, pipelined vs oracle')
    H.close(Eb, oracle.Eemb, rtol=2e-4, what='Eemb, pipelined vs oracle')
    # tickets: only the last 8 can be waited for; a NaN surfaces at the wait
    nat = b._native
    out = N.ctypes.c_float(0)
    assert nat.lib.sert_train_host_wait(nat.handle, 0, N.ctypes.byref(out)) != 0
    assert b'window' in N.load().sert_last_error()
    bad = p['Wp'].copy()
    bad[0, 0] = np.nan
    nat.set_tensor(N.PARAM_DENSE_W, bad)
    with pytest.raises(RuntimeError, match='NaN or infinity'):
        b.train_stream(batches(0, 2))


def test_device_sampled_negatives_and_nan_guard():
    p = H.vs_problem(3, V=400, E=100, dw=16, de=16, W=3, B=32, k=4, n_batches=3)
    model = make_model(p, 0.0, seed=1234)
    n, mean = model.train()
    assert n == 3 and np.isfinite(mean)
    # NaN parameters must surface as the reference's RuntimeError (sert/models.py:372-379)
    from sert_b200 import _native as N
    bad = p['Wp'].copy()
    bad[0, 0] = np.nan
    model._native.set_tensor(N.PARAM_DENSE_W, bad)
    with pytest.raises(RuntimeError, match='NaN or infinity'):
        model.train()


def test_predict_fn_matches_oracle_and_pickles():
    import pickle
    from oracle import sert_oracle as O
    p = H.vs_problem(21, V=300, E=64, dw=300, de=128, W=4, B=32, k=2, n_batches=1)
    model = make_model(p, 0.01)
    state = model.get_state()
    assert len(state) == 3 and state[1].shape == p['R'].shape and state[2].shape == p['Eemb'].shape
    fn = pickle.loads(pickle.dumps(state[0]))
    rng = np.random.default_rng(0)
    for _ in range(3):
        avg = p['R'][rng.integers(0, 300, 5)].mean(axis=0)
        out = fn(avg)
        assert out.shape == (1, 128)
        H.close(out[0], O.vectorspace_predict(p['Wp'], p['bp'], avg), what='predict_fn')


def test_checkpoint_resume_is_exact():
    """Train 2 steps, checkpoint (parameters + Adam state + step), train 2 more; a fresh model restored from the
    checkpoint must reproduce the last 2 steps (the reference cannot resume: no optimiser state)."""
    p = H.vs_problem(31, V=500, E=120, dw=32, de=32, W=4, B=64, k=4, n_batches=4)
    a = make_model(p, 0.01)
    for j in range(2):
        a.train_fn(j, p['neg'][j])
    ckpt = a.get_checkpoint()
    assert int(ckpt['step']) == 2
    tail_a = [a.train_fn(j, p['neg'][j]) for j in (2, 3)]
    b = make_model(p, 0.01)
    b.set_checkpoint(ckpt)
    tail_b = [b.train_fn(j, p['neg'][j]) for j in (2, 3)]
    np.testing.assert_allclose(tail_a, tail_b, rtol=1e-6)     # float atomics: summation order may differ
    Ra, Ea = a.get_representations()
    Rb, Eb = b.get_representations()
    np.testing.assert_allclose(Ra, Rb, rtol=1e-6, atol=1e-9)     # atomics order may differ in the last bit
    np.testing.assert_allclose(Ea, Eb, rtol=1e-6, atol=1e-9)


def test_bf16_optimizer_state_mode_tracks_the_float32_model():
    """dtype_mode 1 (BASELINE.json configs[1] "bf16" perf mode): Adam's m / v live in bfloat16 with stochastic
    rounding.  The forward pass is untouched (first loss identical), later losses and parameters follow the float32
    oracle within the bf16 rounding of the state (stated deviation: 2e-3 relative on losses, 3 % of the parameter
    change per step on parameters), the state reads back as bf16-representable values, and a checkpoint resumes."""
    p = H.vs_problem(21, V=4000, E=1500, dw=128, de=128, W=10, B=256, k=10, n_batches=8)
    model = make_model(p, 0.01, optimizer_state_dtype='bfloat16')
    oracle = H.vs_oracle(p, 0.01)
    R0 = p['R'].copy()
    got, ref = [], []
    for j in range(8):
        got.append(model.train_fn(j, p['neg'][j]))
        ref.append(oracle.train_batch(j, p['neg'][j]))
    assert abs(got[0] - ref[0]) <= 1e-5 * abs(ref[0])
    np.testing.assert_allclose(got, ref, rtol=2e-3)
    R, Eemb = model.get_representations()
    moved = np.abs(oracle.R - R0).mean()
    assert np.abs(R - oracle.R).mean() < 0.03 * 8 * moved / 8 + 1e-7, (np.abs(R - oracle.R).mean(), moved)
    ckpt = model.get_checkpoint()
    state = ckpt['entity_representations/state2'] if 'entity_representations/state2' in ckpt else None
    if state is not None:                     # what comes back is bf16-representable
        assert (state.view(np.uint32) & 0xffff == 0).all()
    other = make_model(p, 0.01, optimizer_state_dtype='bfloat16')
    other.set_checkpoint(ckpt)
    a, b = model.train_fn(0, p['neg'][0]), other.train_fn(0, p['neg'][0])
    assert abs(a - b) <= 1e-6 * abs(a)
    # the arena really is smaller: 16 instead of 24 bytes per parameter for (theta, m, v) + 4 for the gradient
    full = make_model(p, 0.01)
    assert model._native.arena.numel() < full._native.arena.numel() - 3 * (4000 + 1500) * 128


def test_checkpoint_resumes_device_sampled_negatives():
    """with negatives drawn on the device a resumed model must continue the Philox stream, not replay it."""
    p = H.vs_problem(33, V=500, E=120, dw=32, de=32, W=4, B=64, k=4, n_batches=4)
    a = make_model(p, 0.01, seed=1234)
    for j in range(2):
        a.train_fn(j)
    ckpt = a.get_checkpoint()
    assert int(ckpt['sampler_draws']) == 2 and int(ckpt['sampler_seed']) == 1234
    tail_a = [a.train_fn(j) for j in (2, 3)]
    b = make_model(p, 0.01, seed=99)
    b.set_checkpoint(ckpt)
    tail_b = [b.train_fn(j) for j in (2, 3)]
    np.testing.assert_allclose(tail_a, tail_b, rtol=1e-6)
    c = make_model(p, 0.01, seed=1234)          # same seed but a restarted stream: different negatives
    ckpt2 = dict(ckpt)
    ckpt2['sampler_draws'] = np.uint64(0)
    c.set_checkpoint(ckpt2)
    assert abs(c.train_fn(2) - tail_a[0]) > 1e-7

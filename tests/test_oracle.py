"""Self-validation of the CPU oracle: f64 autograd twin, the pooled==per-word identity,
optimiser known answers, tiny literal known-answer vectors."""
import numpy as np
import pytest
import torch

from oracle import sert_oracle as O

# the graph's clip constants are float32 (1-1e-7 rounds to 0.99999988); the f64 twin uses the same values
LO, HI = float(O.CLIP_LO), float(O.CLIP_HI)
TLO, THI = float(O.TANH_LO), float(O.TANH_HI)


def _ll_inputs(gain, seed=0, V=50, E=13, dw=8, W=3, B=6):
    rng = np.random.default_rng(seed)
    R = O.glorot_uniform(rng, (V, dw)) * np.float32(gain)
    Wd = O.glorot_uniform(rng, (dw, E)) * np.float32(gain)
    bd = (rng.standard_normal(E) * 0.1).astype(np.float32)
    x = rng.integers(0, V, (B, W))
    y = np.zeros((B, E), np.float32)
    for i in range(B):
        y[i, rng.choice(E, 2, replace=False)] = 0.5
    w = rng.uniform(0.5, 2, B).astype(np.float32)
    return R, Wd, bd, x, y, w


@pytest.mark.parametrize('gain', [1.0, 20.0])
def test_loglinear_grads_match_f64_autograd(gain):
    R, Wd, bd, x, y, w = _ll_inputs(gain)
    lam = 0.01
    B, W = x.shape
    g = O.loglinear_train_loss_and_grads(R, Wd, bd, x, y, w, lam)
    if gain > 1:
        assert np.mean((g['p'] < 1e-7) | (g['p'] > 1 - 1e-7)) > 0.2      # the clips are active
    Rt, Wt, bt = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (R, Wd, bd)]
    X = Rt[torch.tensor(x)].reshape(B * W, -1)
    p = torch.softmax(X @ Wt + bt, -1)
    s = torch.log(torch.clamp(p, LO, HI)).reshape(B, W, -1).sum(1)
    o = torch.softmax(s, -1)
    ell = -(torch.tensor(y, dtype=torch.float64) * torch.log(torch.clamp(o, LO, HI))).sum(1)
    L = (ell * torch.tensor(w, dtype=torch.float64)).mean() + lam * ((Wt ** 2).sum() + (Rt ** 2).sum()) / (2 * B)
    L.backward()
    np.testing.assert_allclose(g['loss'], L.item(), rtol=2e-6)
    for got, ref in ((g['gR'], Rt.grad), (g['gWd'], Wt.grad), (g['gbd'], bt.grad)):
        ref = ref.numpy()
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('gain', [1.0, 12.0])
def test_vectorspace_grads_match_f64_autograd(gain):
    rng = np.random.default_rng(1)
    V, E, dw, de, W, B, k = 40, 11, 8, 5, 3, 6, 4
    R = O.glorot_uniform(rng, (V, dw)) * np.float32(gain)
    Wp = O.glorot_uniform(rng, (dw, de)) * np.float32(gain)
    bp = (rng.standard_normal(de) * 0.1).astype(np.float32)
    Em = O.glorot_uniform(rng, (E, de)) * np.float32(gain)
    x = rng.integers(0, V, (B, W))
    y = rng.integers(0, E, B).astype(np.int32)
    neg = rng.integers(0, E, (B, k))
    w = rng.uniform(0.5, 2, B).astype(np.float32)
    lam = 0.01
    g = O.vectorspace_train_loss_and_grads(R, Wp, bp, Em, x, y, neg, w, lam)
    Rt, Wt, bt, Et = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (R, Wp, bp, Em)]
    h = Rt[torch.tensor(x)].mean(1)
    u = torch.clamp(torch.tanh(h @ Wt + bt), TLO, THI)
    pos = torch.clamp(torch.sigmoid((Et[torch.tensor(y.astype(np.int64))] * u).sum(-1)), LO, HI)
    ng = torch.clamp(torch.sigmoid((Et[torch.tensor(neg)] * u[:, None]).sum(-1)), LO, HI)
    ell = -(torch.log(pos) + torch.log(1 - ng).sum(1))
    L = (ell * torch.tensor(w, dtype=torch.float64)).mean() + \
        lam * ((Wt ** 2).sum() + (Rt ** 2).sum() + (Et ** 2).sum()) / (2 * B)
    L.backward()
    # near the clips 1-sigmoid loses bits in float32 (as it does in the reference's float32 graph), which an
    # f64 twin cannot reproduce: the saturated case checks the formulas at a looser tolerance
    tol = 3e-6 if gain == 1.0 else 2e-3
    np.testing.assert_allclose(g['loss'], L.item(), rtol=tol)
    for got, ref in ((g['gR'], Rt.grad), (g['gWp'], Wt.grad), (g['gbp'], bt.grad), (g['gE'], Et.grad)):
        ref = ref.numpy()
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max())


def test_pooled_identity_when_unclipped_and_not_when_clipped():
    """SURVEY.md note N1: softmax(sum_w log softmax(z_w)) == softmax((sum_w R[x_w]).W + W_win*b) iff no clip."""
    for gain, same in ((1.0, True), (40.0, False)):
        R, Wd, bd, x, y, w = _ll_inputs(gain)
        f = O.loglinear_forward(R, Wd, bd, x)
        pooled = R[x].sum(axis=1) @ Wd + x.shape[1] * bd
        o2 = O.softmax_rows(pooled.astype(np.float32))
        if same:
            np.testing.assert_allclose(f['o'], o2, rtol=2e-4, atol=1e-7)
        else:
            assert np.abs(f['o'] - o2).max() > 1e-3


def test_adadelta_and_adam_known_answers():
    # hand-computed scalars, Lasagne 0.1 update rules
    p, a, d = np.float32([1.0]), np.float32([0.0]), np.float32([0.0])
    g = np.float32([0.5])
    p1, a1, d1 = O.adadelta_update(p, g, a, d)
    accu = 0.05 * 0.25
    upd = 0.5 * np.sqrt(1e-6) / np.sqrt(accu + 1e-6)
    np.testing.assert_allclose([p1[0], a1[0], d1[0]], [1.0 - upd, accu, 0.05 * upd * upd], rtol=1e-5)
    p2, a2, d2 = O.adadelta_update(p1, g, a1, d1)
    accu2 = 0.95 * accu + 0.05 * 0.25
    upd2 = 0.5 * np.sqrt(0.05 * upd * upd + 1e-6) / np.sqrt(accu2 + 1e-6)
    np.testing.assert_allclose(p2[0], 1.0 - upd - upd2, rtol=1e-5)
    # Adam: first step moves by ~lr * sign(g) regardless of |g|
    m = v = np.float32([0.0])
    q1, m1, v1 = O.adam_update(p, g, m, v, 1)
    a_1 = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(q1[0], 1.0 - a_1 * (0.1 * 0.5) / (np.sqrt(0.001 * 0.25) + 1e-8), rtol=1e-5)
    np.testing.assert_allclose(1.0 - q1[0], 1e-3, rtol=1e-3)
    q2, m2, v2 = O.adam_update(q1, g, m1, v1, 2)
    a_2 = 1e-3 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    np.testing.assert_allclose(m2[0], 0.9 * 0.05 + 0.05, rtol=1e-6)
    np.testing.assert_allclose(q2[0], q1[0] - a_2 * m2[0] / (np.sqrt(v2[0]) + 1e-8), rtol=1e-6)


def test_tiny_known_answer_loglinear():
    """V=3,E=2,dw=1,W=2,B=1 worked by hand."""
    R = np.float32([[0.0], [1.0], [2.0]])
    Wd = np.float32([[1.0, -1.0]])
    bd = np.float32([0.0, 0.0])
    x = np.array([[1, 2]])
    f = O.loglinear_forward(R, Wd, bd, x)
    # z = [[1,-1],[2,-2]] -> log-softmax rows: [-log(1+e^-2), -2-log(1+e^-2)], [-log(1+e^-4), -4-log(1+e^-4)]
    s0 = -np.log1p(np.exp(-2.0)) - np.log1p(np.exp(-4.0))
    s1 = s0 - 6.0
    np.testing.assert_allclose(f['s'][0], [s0, s1], rtol=1e-6)
    np.testing.assert_allclose(f['o'][0], [1 / (1 + np.exp(-6.0)), 1 / (1 + np.exp(6.0))], rtol=1e-5)
    y = np.float32([[0.0, 1.0]])
    ell = O.loglinear_instance_losses(f['o'], y)
    np.testing.assert_allclose(ell[0], np.log1p(np.exp(6.0)), rtol=1e-5)


def test_tiny_known_answer_vectorspace():
    R = np.float32([[1.0, 0.0], [0.0, 1.0]])
    Wp = np.eye(2, dtype=np.float32)
    bp = np.zeros(2, np.float32)
    Em = np.float32([[2.0, 0.0], [0.0, -2.0]])
    x = np.array([[0, 1]])
    f = O.vectorspace_forward(R, Wp, bp, Em, x, np.int32([0]), np.array([[1]]))
    u = np.tanh(0.5)
    np.testing.assert_allclose(f['u'][0], [u, u], rtol=1e-6)
    np.testing.assert_allclose(f['score_pos'][0], 2 * u, rtol=1e-6)
    np.testing.assert_allclose(f['score_neg'][0, 0], -2 * u, rtol=1e-6)
    # ell = -(log sig(2u) + log(1 - sig(-2u))) = 2*log(1+exp(-2u))
    np.testing.assert_allclose(f['ell'][0], 2 * np.log1p(np.exp(-2 * u)), rtol=1e-5)
    np.testing.assert_allclose(O.vectorspace_predict(Wp, bp, np.float32([0.5, 0.5])), [u, u], rtol=1e-6)


def test_batch_protocol_drops_tail_and_reports_mean_std():
    from sert_b200 import synth
    train, val = synth.vectorspace_corpus(3, 50, 20, 3, 70, 33)
    rng = np.random.default_rng(0)
    orc = O.VectorSpaceOracle(32, O.glorot_uniform(rng, (50, 4)), O.glorot_uniform(rng, (4, 4)),
                              np.zeros(4, np.float32), O.glorot_uniform(rng, (20, 4)), 0.01, train, val)
    neg = rng.integers(0, 20, (2, 32, 3))
    mean, std = orc.error('train', neg)
    errs = [orc.eval_batch('train', b, neg[b]) for b in range(2)]       # 70 // 32 == 2 batches, 6 rows dropped
    np.testing.assert_allclose([mean, std], [np.mean(errs), np.std(errs)])
    n, m = orc.train_epoch([1, 0], neg)
    assert n == 2 and np.isfinite(m) and orc.t == 2

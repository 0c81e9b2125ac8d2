"""CPU oracle for the SERT training / scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sert_b200/`` (the product) may import
this module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline.

It is a numpy float32 restatement of the arithmetic that the reference's Theano
graph performs (all citations are relative to the reference checkout):

* ``sert/models.py:158-183``   SparseProjectionLayer (embedding gather)
* ``sert/models.py:186-212``   ProductTimestepsLayer (sum of clipped log-probs + softmax)
* ``sert/models.py:215-241``   MeanLayer / ClipLayer
* ``sert/models.py:244-292``   WeightedObjective / clipped_categorical_crossentropy
* ``sert/models.py:764-795``   L2 regularisation (which tensors, which scale)
* ``sert/models.py:804-878``   LanguageModel (log-linear) graph
* ``sert/models.py:893-1009``  sigmoid distance, negative sampling, entity gather
* ``sert/models.py:1024-1118`` VectorSpaceLanguageModel graph and its predict_fn
* ``sert/models.py:322-399,638-668`` batch slicing, tail drop, mean/std protocol
* ``bin/query.py:199-382``     ranking callbacks

Third-party arithmetic that is NOT vendored in the reference (Theano==0.8.2,
Lasagne==0.1, ``requirements.txt:3,11``) is restated from its published
definition: ``lasagne.updates.adadelta`` / ``lasagne.updates.adam`` (Lasagne 0.1
forms), ``T.nnet.softmax`` (max-subtracted), ``T.clip`` (gradient is 1 on the
closed interval, 0 outside), ``T.nnet.sigmoid``, ``lasagne.init.GlorotUniform``.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  The oracle is pinned (a) against the reference's own
``sert/models.py`` graph-building code executed under an eager numpy/torch shim
of Theano/Lasagne (``oracle/theano_shim`` + ``tests/golden/make_golden.py``) and
(b) against the unmodified ``sert/inference.py`` / ``bin/query.py`` callbacks run
in the build container; fixtures from both live in ``tests/golden/``.

All arrays are float32 unless noted; full reductions use float64 accumulators
like Theano's CPU ``Sum`` op (acc_dtype float64 for float32 inputs).
"""
from __future__ import annotations

import collections
import operator

import numpy as np

F32 = np.float32
CLIP_LO = F32(1e-7)                   # sert/models.py:200,290,900
CLIP_HI = F32(1.0 - 1e-7)             # rounds to 0.99999988 in f32
TANH_LO = F32(-1.0 + 1e-7)            # sert/models.py:1067
TANH_HI = F32(1.0 - 1e-7)             # sert/models.py:1068

ADADELTA = dict(learning_rate=1.0, rho=0.95, epsilon=1e-6)              # lasagne 0.1 defaults
ADAM = dict(learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8)   # lasagne 0.1 defaults


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------

def glorot_uniform(rng, shape):
    """lasagne.init.GlorotUniform().sample(shape) for a 2-D shape (bin/train.py:128,170)."""
    n1, n2 = shape
    a = np.sqrt(6.0 / (n1 + n2))
    return rng.uniform(-a, a, size=shape).astype(F32)


def softmax_rows(z):
    """T.nnet.softmax: exp(z - max) / sum, row-wise over the last axis, in f32."""
    z = np.asarray(z, dtype=F32)
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m, dtype=F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def sigmoid(x):
    x = np.asarray(x, dtype=F32)
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def _in_closed(v, lo, hi):
    return ((v >= lo) & (v <= hi)).astype(F32)


def dense_rows(y, start, stop, num_cols=None):
    """Dense float32 slice of the label matrix (S.dense_from_sparse, sert/models.py:75-89,497-500)."""
    if hasattr(y, 'tocsr'):
        return np.asarray(y[start:stop].todense(), dtype=F32)
    return np.asarray(y[start:stop], dtype=F32)


def sumsq64(a):
    return float(np.sum(np.square(a.astype(np.float64))))


# ----------------------------------------------------------------------------
# log-linear model  (sert/models.py:804-878)
# ----------------------------------------------------------------------------

def loglinear_forward(R, Wd, bd, x):
    """x (B,W) ints -> dict with z (B,W,E) logits, p per-word softmax, s joint logits, o (B,E)."""
    B, W = x.shape
    X = R[x.astype(np.int64)].reshape(B * W, -1)          # :180, :838
    z = (X @ Wd + bd).astype(F32)                         # :846-849
    p = softmax_rows(z)                                   # :841
    q = np.clip(p, CLIP_LO, CLIP_HI)                      # :200
    logq = np.log(q, dtype=F32).reshape(B, W, -1)
    s = logq.sum(axis=1, dtype=F32)                       # :201
    o = softmax_rows(s)                                   # :210
    return dict(X=X, z=z.reshape(B, W, -1), p=p.reshape(B, W, -1), s=s, o=o)


def loglinear_predict(R, Wd, bd, batch, mask=None):
    """predict_fn(batch, mask) -> (B,W,E) unclipped per-word softmax (sert/models.py:868,880-890).
    The mask input is accepted and ignored (on_unused_input='warn')."""
    return loglinear_forward(R, Wd, bd, batch)['p']


def loglinear_instance_losses(o, y_dense):
    c = np.clip(o, CLIP_LO, CLIP_HI)                      # :290
    return (-(y_dense * np.log(c, dtype=F32)).sum(axis=1, dtype=F32)).astype(F32)   # :292


def loglinear_eval_loss(R, Wd, bd, x, y_dense):
    """test_fn / validate_fn: mean of unweighted instance losses, no regulariser (:751-752)."""
    f = loglinear_forward(R, Wd, bd, x)
    return F32(loglinear_instance_losses(f['o'], y_dense).mean(dtype=F32))


def loglinear_train_loss_and_grads(R, Wd, bd, x, y_dense, w, lam):
    """Loss of train_fn and d(loss)/d[R, Wd, bd] in the general (clipped) regime."""
    B, W = x.shape
    f = loglinear_forward(R, Wd, bd, x)
    p, o, X = f['p'], f['o'], f['X']
    ell = loglinear_instance_losses(o, y_dense)
    data_loss = F32((ell * w).mean(dtype=F32))            # :279-282
    reg = F32(0.0)
    if lam > 0.0:                                         # :764-795
        reg = F32(lam * sumsq64(Wd) / (2.0 * B) + lam * sumsq64(R) / (2.0 * B))
    loss = F32(data_loss + reg)

    # backward
    c = np.clip(o, CLIP_LO, CLIP_HI)
    dell = (w / F32(B)).astype(F32)[:, None]
    do = (-dell * y_dense / c) * _in_closed(o, CLIP_LO, CLIP_HI)
    ds = o * (do - (do * o).sum(axis=1, keepdims=True, dtype=F32))
    qc = np.clip(p, CLIP_LO, CLIP_HI)
    dp = (ds[:, None, :] / qc) * _in_closed(p, CLIP_LO, CLIP_HI)
    dz = p * (dp - (dp * p).sum(axis=2, keepdims=True, dtype=F32))
    dz2 = dz.reshape(B * W, -1).astype(F32)
    scale = F32(lam / B) if lam > 0.0 else F32(0.0)
    gWd = (X.T @ dz2).astype(F32) + scale * Wd
    gbd = dz2.sum(axis=0, dtype=F32)
    dX = (dz2 @ Wd.T).astype(F32)
    gR = np.zeros_like(R)
    np.add.at(gR, x.astype(np.int64).reshape(-1), dX)
    gR += scale * R
    out = dict(f)
    out.update(loss=loss, data_loss=data_loss, ell=ell, ds=ds.astype(F32), dz=dz2,
               gR=gR.astype(F32), gWd=gWd.astype(F32), gbd=gbd.astype(F32))
    return out


# ----------------------------------------------------------------------------
# vector-space model  (sert/models.py:1024-1118)
# ----------------------------------------------------------------------------

def vectorspace_forward(R, Wp, bp, Eemb, x, y, neg):
    """x (B,W), y (B,) int32, neg (B,k) ints -> dict(h,t,u,score_pos,score_neg,pos,neg,ell)."""
    h = R[x.astype(np.int64)].mean(axis=1, dtype=F32)      # :1047-1051
    a = (h @ Wp + bp).astype(F32)
    t = np.tanh(a, dtype=F32)                              # :1055-1061
    u = np.clip(t, TANH_LO, TANH_HI)                       # :1065-1068
    epos = Eemb[y.astype(np.int64)]                        # :990
    eneg = Eemb[neg.astype(np.int64)]
    score_pos = (epos * u).sum(axis=1, dtype=F32)          # :896-898
    score_neg = (eneg * u[:, None, :]).sum(axis=2, dtype=F32)
    spos, sneg = sigmoid(score_pos), sigmoid(score_neg)
    pos = np.clip(spos, CLIP_LO, CLIP_HI)                  # :900
    ng = np.clip(sneg, CLIP_LO, CLIP_HI)
    ell = -(np.log(pos, dtype=F32) + np.log(F32(1.0) - ng, dtype=F32).sum(axis=1, dtype=F32))   # :1091-1098
    return dict(h=h, a=a, t=t, u=u, epos=epos, eneg=eneg, score_pos=score_pos, score_neg=score_neg,
                spos=spos, sneg=sneg, pos=pos, neg=ng, ell=ell.astype(F32))


def vectorspace_predict(Wp, bp, avg):
    """predict_fn(avg (dw,)) -> tanh(avg.Wp + bp), NO clip (sert/models.py:1107-1118)."""
    return np.tanh((np.asarray(avg, dtype=F32) @ Wp + bp).astype(F32), dtype=F32)


def vectorspace_eval_loss(R, Wp, bp, Eemb, x, y, neg):
    return F32(vectorspace_forward(R, Wp, bp, Eemb, x, y, neg)['ell'].mean(dtype=F32))


def vectorspace_train_loss_and_grads(R, Wp, bp, Eemb, x, y, neg, w, lam):
    B, W = x.shape
    f = vectorspace_forward(R, Wp, bp, Eemb, x, y, neg)
    data_loss = F32((f['ell'] * w).mean(dtype=F32))
    reg = F32(0.0)
    if lam > 0.0:                                          # :1100-1105, :764-795
        reg = F32(lam * sumsq64(Wp) / (2.0 * B) + lam * (sumsq64(R) + sumsq64(Eemb)) / (2.0 * B))
    loss = F32(data_loss + reg)

    c = (w / F32(B)).astype(F32)
    spos, sneg, pos, ng, u, t = f['spos'], f['sneg'], f['pos'], f['neg'], f['u'], f['t']
    # d ell / d pos = -1/pos ; through clip ; through sigmoid
    gpos = (-c / pos) * _in_closed(spos, CLIP_LO, CLIP_HI) * spos * (F32(1.0) - spos)
    gneg = (c[:, None] / (F32(1.0) - ng)) * _in_closed(sneg, CLIP_LO, CLIP_HI) * sneg * (F32(1.0) - sneg)
    gpos = gpos.astype(F32)
    gneg = gneg.astype(F32)
    du = gpos[:, None] * f['epos'] + (gneg[:, :, None] * f['eneg']).sum(axis=1, dtype=F32)
    gE = np.zeros_like(Eemb)
    np.add.at(gE, y.astype(np.int64), gpos[:, None] * u)
    np.add.at(gE, neg.astype(np.int64).reshape(-1),
              (gneg[:, :, None] * u[:, None, :]).reshape(-1, u.shape[1]))
    dt = du * _in_closed(t, TANH_LO, TANH_HI)
    da = (dt * (F32(1.0) - t * t)).astype(F32)
    scale = F32(lam / B) if lam > 0.0 else F32(0.0)
    gWp = (f['h'].T @ da).astype(F32) + scale * Wp
    gbp = da.sum(axis=0, dtype=F32)
    dh = (da @ Wp.T).astype(F32)
    gR = np.zeros_like(R)
    np.add.at(gR, x.astype(np.int64).reshape(-1), np.repeat(dh / F32(W), W, axis=0))
    gR += scale * R
    gE += scale * Eemb
    out = dict(f)
    out.update(loss=loss, data_loss=data_loss, gpos=gpos, gneg=gneg, du=du.astype(F32), da=da, dh=dh,
               gE=gE.astype(F32), gR=gR.astype(F32), gWp=gWp.astype(F32), gbp=gbp.astype(F32))
    return out


# ----------------------------------------------------------------------------
# optimisers (Lasagne 0.1; call sites sert/models.py:548-549,820,922)
# ----------------------------------------------------------------------------

def adadelta_update(param, grad, accu, delta_accu, learning_rate=1.0, rho=0.95, epsilon=1e-6):
    rho, eps, lr = F32(rho), F32(epsilon), F32(learning_rate)
    one_m = F32(1.0) - rho
    accu_new = rho * accu + one_m * grad * grad
    update = grad * np.sqrt(delta_accu + eps, dtype=F32) / np.sqrt(accu_new + eps, dtype=F32)
    param_new = param - lr * update
    delta_new = rho * delta_accu + one_m * update * update
    return param_new.astype(F32), accu_new.astype(F32), delta_new.astype(F32)


def adam_alpha(t, learning_rate=1e-3, beta1=0.9, beta2=0.999):
    """a_t = lr*sqrt(1-b2^t)/(1-b1^t) evaluated in float32 like the Theano graph (t is a floatX scalar)."""
    t = F32(t)
    b1, b2, lr, one = F32(beta1), F32(beta2), F32(learning_rate), F32(1.0)
    return F32(lr * np.sqrt(one - np.power(b2, t, dtype=F32), dtype=F32) / (one - np.power(b1, t, dtype=F32)))


def adam_update(param, grad, m, v, t, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """t is the step count AFTER increment (first call t=1)."""
    b1, b2, eps = F32(beta1), F32(beta2), F32(epsilon)
    a_t = adam_alpha(t, learning_rate, beta1, beta2)
    m_new = b1 * m + (F32(1.0) - b1) * grad
    v_new = b2 * v + (F32(1.0) - b2) * grad * grad
    step = a_t * m_new / (np.sqrt(v_new, dtype=F32) + eps)
    return (param - step).astype(F32), m_new.astype(F32), v_new.astype(F32)


# ----------------------------------------------------------------------------
# model objects with the batching protocol (sert/models.py:322-399,638-668)
# ----------------------------------------------------------------------------

class _OracleBase(object):

    def _num_batches(self, n):
        return n // self.B                                 # tail dropped, :355-359

    def _slice(self, b):
        return slice(b * self.B, (b + 1) * self.B)         # :322-326

    def train_epoch(self, order, negatives=None):
        """order: iterable of batch indices (the shuffled order, :363-367)."""
        losses = []
        for j, b in enumerate(order):
            losses.append(self.train_batch(int(b), None if negatives is None else negatives[j]))
            if not np.isfinite(losses[-1]):
                raise RuntimeError('Encountered NaN or infinity')
        return len(losses), np.mean(losses)

    def error(self, split, negatives=None):
        x, y = (self.x_train, self.y_train) if split == 'train' else (self.x_val, self.y_val)
        n = self._num_batches(x.shape[0])
        errs = [self.eval_batch(split, b, None if negatives is None else negatives[b]) for b in range(n)]
        return np.mean(errs), np.std(errs)


class LogLinearOracle(_OracleBase):
    """models.LanguageModel: params [R, Wd, bd], Adadelta, dense L2 (sert/models.py:804-878,820)."""

    def __init__(self, batch_size, R, Wd, bd, lam, training_set, validation_set):
        self.B = int(batch_size)
        self.R, self.Wd, self.bd = R.astype(F32).copy(), Wd.astype(F32).copy(), bd.astype(F32).copy()
        self.lam = float(lam)
        self.x_train, self.y_train, self.w_train = training_set
        self.x_val, self.y_val = validation_set
        self.state = {n: (np.zeros_like(p), np.zeros_like(p))
                      for n, p in (('R', self.R), ('Wd', self.Wd), ('bd', self.bd))}

    def train_batch(self, b, _neg=None):
        sl = self._slice(b)
        g = loglinear_train_loss_and_grads(self.R, self.Wd, self.bd, self.x_train[sl],
                                           dense_rows(self.y_train, sl.start, sl.stop),
                                           np.asarray(self.w_train[sl], dtype=F32), self.lam)
        for name, grad in (('R', g['gR']), ('Wd', g['gWd']), ('bd', g['gbd'])):
            a, d = self.state[name]
            p, a, d = adadelta_update(getattr(self, name), grad, a, d)
            setattr(self, name, p)
            self.state[name] = (a, d)
        return g['loss']

    def eval_batch(self, split, b, _neg=None):
        x, y = (self.x_train, self.y_train) if split == 'train' else (self.x_val, self.y_val)
        sl = self._slice(b)
        return loglinear_eval_loss(self.R, self.Wd, self.bd, x[sl], dense_rows(y, sl.start, sl.stop))


class VectorSpaceOracle(_OracleBase):
    """models.VectorSpaceLanguageModel: params [Eemb, R, Wp, bp], Adam (one shared t), dense L2."""

    def __init__(self, batch_size, R, Wp, bp, Eemb, lam, training_set, validation_set):
        self.B = int(batch_size)
        self.R, self.Wp, self.bp, self.Eemb = (a.astype(F32).copy() for a in (R, Wp, bp, Eemb))
        self.lam = float(lam)
        self.x_train, self.y_train, self.w_train = training_set
        self.x_val, self.y_val = validation_set
        self.t = 0
        self.state = {n: (np.zeros_like(getattr(self, n)), np.zeros_like(getattr(self, n)))
                      for n in ('Eemb', 'R', 'Wp', 'bp')}

    def train_batch(self, b, neg):
        sl = self._slice(b)
        g = vectorspace_train_loss_and_grads(self.R, self.Wp, self.bp, self.Eemb, self.x_train[sl],
                                             self.y_train[sl], neg, np.asarray(self.w_train[sl], dtype=F32),
                                             self.lam)
        self.t += 1
        for name, grad in (('Eemb', g['gE']), ('R', g['gR']), ('Wp', g['gWp']), ('bp', g['gbp'])):
            m, v = self.state[name]
            p, m, v = adam_update(getattr(self, name), grad, m, v, self.t)
            setattr(self, name, p)
            self.state[name] = (m, v)
        return g['loss']

    def eval_batch(self, split, b, neg):
        x, y = (self.x_train, self.y_train) if split == 'train' else (self.x_val, self.y_val)
        sl = self._slice(b)
        return vectorspace_eval_loss(self.R, self.Wp, self.bp, self.Eemb, x[sl], y[sl], neg)


# ----------------------------------------------------------------------------
# scoring (sert/inference.py:170-183, bin/query.py:199-382)
# ----------------------------------------------------------------------------

def aggregate_product(distribution):
    """inference.aggregate_distribution(mode='product', axis=0): exact zeros are skipped (:173-174)."""
    return np.exp(np.sum(np.ma.log(distribution).filled(0), axis=0))


def loglinear_rank(per_term_distribution):
    """LogLinearCallback.process (bin/query.py:204-233) -> (indices descending, values)."""
    distribution = aggregate_product(np.asarray(per_term_distribution))
    distribution = distribution / distribution.sum()
    ranked = np.argsort(distribution)[::-1]
    return ranked, distribution[ranked]


def normalise_rows(m):
    """bin/query.py:270-274 (entities) and :333-336 (query projection), float32."""
    m = np.array(m, dtype=F32, copy=True)
    m /= np.linalg.norm(m, axis=1)[:, np.newaxis]
    return m


def vectorspace_rank(entities_normalised, projection, top=None, algorithm='kd_tree'):
    """VectorSpaceCallback.query/process (bin/query.py:304-367) for ONE query projection.

    entities_normalised: (E,de) f32 rows already L2-normalised (callback constructor).
    Candidate set: exact Euclidean k-NN (sklearn tree search, distances in f64) or all E via cdist;
    relevance recomputed as (sum(e*q)+1)/2.  Returns (indices, values) sorted by value descending
    with Python's stable sort, as the callback does.
    """
    import scipy.spatial.distance
    q = normalise_rows(np.asarray(projection, dtype=F32).reshape(1, -1))
    E = entities_normalised
    if top is not None and top <= E.shape[0]:
        import sklearn.neighbors
        nn = sklearn.neighbors.NearestNeighbors(n_neighbors=top, algorithm=algorithm, metric='euclidean')
        nn.fit(E)
        _, indices = nn.kneighbors(q)
    else:
        pd = scipy.spatial.distance.cdist(q, E, metric='euclidean')
        indices = pd.argsort(axis=1)
    candidates = collections.defaultdict(float)
    for candidate in indices[0, :]:
        score = np.sum(E[candidate, :] * q[0, :])
        score = (score + 1.0) / 2.0
        candidates[candidate] += score
    idx, val = map(np.array, zip(*sorted(candidates.items(), reverse=True, key=operator.itemgetter(1))))
    return idx, val


def write_run_order(assessments):
    """trec_utils.write_run ordering (sub:trec_utils.py:560-561): sorted((relevance, id), reverse=True)."""
    return sorted(assessments, reverse=True)

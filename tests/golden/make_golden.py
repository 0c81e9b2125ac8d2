#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

What runs unmodified from /root/reference:
  * sert/models.py            (LanguageModel, VectorSpaceLanguageModel: graph construction, batching,
                               train/test/validate/predict callables) -- under the eager Theano/Lasagne shim
                               in oracle/theano_shim (Theano 0.8.2 / Lasagne 0.1 are not installable here)
  * sert/inference.py, sert/math_utils.py                       (pure numpy; run as they are)
  * bin/query.py  callbacks   (LogLinearCallback, VectorSpaceCallback, compute_normalised_entropy)
  * bin/train.py  sparse_to_one_hot_multiple
  * bin/prepare.py instances_and_labels_to_arrays (and the w_train expression of main(), :395-399)
  * cvangysel-common io_utils.windowed_translated_token_stream
  * cvangysel-common trec_utils.parse_query / parse_topics / write_run
Third-party modules the reference imports but this path never calls (bs4, nltk, gensim) are
stubbed with empty modules; cvangysel.sklearn_utils.neighbors_algorithm, which crashes on modern sklearn
(sklearn.neighbors.ball_tree is gone), is replaced by a function returning 'kd_tree' -- the exact tree
search that 'auto' resolved to for k < E/2 in the sklearn of the reference's time.

Outputs: *.npz / *.json next to this script (small; committed).  Nothing here runs on the GPU box.
"""
import importlib.util
import io
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'



class _Stub(types.ModuleType):
    """Empty stand-in for a third-party module the reference imports but this path never calls."""

    def __getattr__(self, item):
        if item.startswith('__'):
            raise AttributeError(item)
        return type(item, (object,), {})


for name in ('bs4', 'nltk', 'nltk.probability', 'nltk.corpus', 'gensim', 'sklearn.cross_validation'):
    sys.modules.setdefault(name, _Stub(name))
sys.modules['nltk'].probability = sys.modules['nltk.probability']
sys.modules['nltk'].corpus = sys.modules['nltk.corpus']
sys.path[:0] = [os.path.join(ROOT, 'oracle', 'theano_shim'), REF, os.path.join(REF, 'cvangysel-common', 'py')]
sys.path.append(ROOT)

import numpy as np  # noqa: E402
import scipy.sparse  # noqa: E402

import sert.models as ref_models  # noqa: E402  (the reference's)
import sert.inference as ref_inference  # noqa: E402
import sert.math_utils as ref_math  # noqa: E402
import cvangysel  # noqa: E402  (the reference's)
from cvangysel import sklearn_utils, trec_utils  # noqa: E402
from theano.tensor import shared_randomstreams as RS  # noqa: E402

assert ref_models.__file__.startswith(REF) and cvangysel.__file__.startswith(REF)
sklearn_utils.neighbors_algorithm = lambda metric: 'kd_tree'


def load_script(name):
    spec = importlib.util.spec_from_file_location('ref_bin_' + name, os.path.join(REF, 'bin', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_query = load_script('query')
ref_train = load_script('train')
ref_prepare = load_script('prepare')

from sert_b200 import synth  # noqa: E402


def shared_by_name(fn, name):
    return [sv for sv, _ in fn.updates if getattr(sv, 'name', None) == name][0]


def csr_parts(m, prefix):
    m = m.tocsr()
    return {prefix + '_indptr': m.indptr.astype(np.int64), prefix + '_indices': m.indices.astype(np.int32),
            prefix + '_data': m.data.astype(np.float32), prefix + '_shape': np.array(m.shape, dtype=np.int64)}


def gen_loglinear():
    np.random.seed(20160817)
    V, E, dw, W, B = 400, 64, 16, 5, 32
    train, val = synth.loglinear_corpus(101, V, E, W, B * 6 + 7, B * 2 + 3)
    R0 = synth.glorot(np.random.default_rng(1), (V, dw))
    model = ref_models.LanguageModel(batch_size=B, window_size=W, representations_init=R0.copy(),
                                     output_layer_size=E, regularization_lambda=0.01,
                                     training_set=train, validation_set=val)
    Rv, Wv, bv = (shared_by_name(model.train_fn, n) for n in ('Representations', 'W', 'b'))
    out = dict(V=V, E=E, dw=dw, W=W, B=B, lam=0.01, x_train=train[0], w_train=train[2], x_val=val[0],
               R0=Rv.get_value().copy(), Wd0=Wv.get_value().copy(), bd0=bv.get_value().copy())
    out.update(csr_parts(train[1], 'y_train'))
    out.update(csr_parts(val[1], 'y_val'))
    out['train_error0'] = np.array(model.train_error(), dtype=np.float64)
    out['validation_error0'] = np.array(model.validation_error(), dtype=np.float64)
    order = np.array([4, 1, 5, 0, 3, 2], dtype=np.int64)
    out['order'] = order
    out['train_losses'] = np.array([float(np.asarray(model.train_fn(int(b))).reshape(-1)[0]) for b in order],
                                   dtype=np.float32)
    out['R1'], out['Wd1'], out['bd1'] = Rv.get_value().copy(), Wv.get_value().copy(), bv.get_value().copy()
    out['train_error1'] = np.array(model.train_error(), dtype=np.float64)
    out['validation_error1'] = np.array(model.validation_error(), dtype=np.float64)
    batch = train[0][:B]
    out['predict_batch'] = batch
    out['predict_out'] = np.asarray(model.predict_fn(batch, np.ones((B, W), np.int8)), dtype=np.float32)
    state = model.get_state()
    # Reference quirk: ModelBase.get_state() tests hasattr(representations, '__iter__') (sert/models.py:674),
    # which is true for an ndarray, so the log-linear state is [predict_fn, R[0], R[1], ..., R[V-1]].
    assert len(state) == 1 + V and state[1].shape == (dw,)
    out['state_len'] = len(state)
    np.savez_compressed(os.path.join(HERE, 'loglinear_ref.npz'), **out)
    return model, out


def gen_vectorspace():
    np.random.seed(20160818)
    V, E, dw, de, W, B, k = 500, 96, 24, 16, 4, 32, 5
    train, val = synth.vectorspace_corpus(202, V, E, W, B * 6 + 5, B * 2 + 1)
    rng = np.random.default_rng(2)
    train = (train[0], train[1], synth.make_weights(rng, train[0].shape[0]))
    R0, E0 = synth.glorot(rng, (V, dw)), synth.glorot(rng, (E, de))
    model = ref_models.VectorSpaceLanguageModel(
        batch_size=B, window_size=W, num_negative_samples=k, representations_init=R0.copy(),
        entity_representations_init=E0.copy(), regularization_lambda=0.01,
        training_set=train, validation_set=val)
    g = lambda n: shared_by_name(model.train_fn, n)  # noqa: E731
    Rv, Wv, bv, Ev = g('Representations'), g('WordProjection.W'), g('WordProjection.b'), g('Class representations')
    out = dict(V=V, E=E, dw=dw, de=de, W=W, B=B, k=k, lam=0.01, x_train=train[0], y_train=train[1],
               w_train=train[2], x_val=val[0], y_val=val[1], R0=Rv.get_value().copy(), Wp0=Wv.get_value().copy(),
               bp0=bv.get_value().copy(), E0=Ev.get_value().copy())

    def with_draw(fn, b):
        n0 = len(RS.DRAWS)
        loss = float(np.asarray(fn(int(b))).reshape(-1)[0])
        assert len(RS.DRAWS) == n0 + 1
        return loss, RS.DRAWS[n0][1].astype(np.int32)

    losses, negs = zip(*[with_draw(model.test_fn, b) for b in range(6)])
    out['test_losses0'], out['test_negs0'] = np.array(losses, np.float32), np.stack(negs)
    losses, negs = zip(*[with_draw(model.validate_fn, b) for b in range(2)])
    out['val_losses0'], out['val_negs0'] = np.array(losses, np.float32), np.stack(negs)
    order = np.array([2, 5, 0, 3, 1, 4], dtype=np.int64)
    losses, negs = zip(*[with_draw(model.train_fn, b) for b in order])
    out['order'], out['train_losses'], out['train_negs'] = order, np.array(losses, np.float32), np.stack(negs)
    out['R1'], out['Wp1'], out['bp1'], out['E1'] = (v.get_value().copy() for v in (Rv, Wv, bv, Ev))
    losses, negs = zip(*[with_draw(model.test_fn, b) for b in range(6)])
    out['test_losses1'], out['test_negs1'] = np.array(losses, np.float32), np.stack(negs)
    avgs = np.stack([R0[rng.integers(0, V, 3)].mean(axis=0) for _ in range(4)]).astype(np.float32)
    out['predict_in'] = avgs
    out['predict_out'] = np.stack([np.asarray(model.predict_fn(a), dtype=np.float32) for a in avgs])
    state = model.get_state()
    assert len(state) == 3 and state[2].shape == (E, de)
    np.savez_compressed(os.path.join(HERE, 'vectorspace_ref.npz'), **out)
    return out


def gen_inference():
    """Reference WordBatcher / aggregate_distribution / entropy on ragged queries."""
    rng = np.random.default_rng(303)
    B, W, V, E = 4, 3, 50, 7
    table = rng.random((V, E)).astype(np.float32)
    table /= table.sum(axis=1, keepdims=True)

    def predict_fn(batch, mask):
        return table[batch.astype(np.int64)]             # (B, W, E): a fixed per-word distribution

    calls = []

    class Recorder(object):
        def __call__(self, payload, result, **kwargs):
            calls.append({'payload': [int(t) for t in payload], 'result': np.asarray(result).tolist(),
                          'kwargs': kwargs})

        def should_average_input(self):
            return False

    batcher = ref_inference.create(predict_fn, None, B, W, V, Recorder())
    queries = [[1], [2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12, 13, 14, 15], [16, 17], [18, 19, 20, 21, 22, 23],
               [24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35]]
    for i, q in enumerate(queries):
        batcher.submit(list(q), topic_id='T%d' % i)
    batcher.process()
    too_long = False
    try:
        batcher.submit(list(range(13)), topic_id='X')
    except RuntimeError:
        too_long = True
    dist = rng.random((5, 9)).astype(np.float32)
    dist[1, 3] = 0.0
    dist[4, 0] = 0.0
    agg = {mode: np.asarray(ref_inference.aggregate_distribution(dist, mode, 0)).tolist()
           for mode in ('sum', 'product', 'last', 'max', 'identity')}
    pk = rng.random(11)
    ent = {'plain': float(ref_math.entropy(pk)), 'base2_norm': float(ref_math.entropy(pk, base=2, normalize=True)),
           'norm': float(ref_math.entropy(pk, normalize=True))}
    json.dump({'B': B, 'W': W, 'V': V, 'E': E, 'table': table.tolist(), 'queries': queries, 'calls': calls,
               'too_long_raises': too_long, 'dist': dist.tolist(), 'aggregate': agg, 'pk': pk.tolist(),
               'entropy': ent, 'instance_dtype': str(batcher.batch.dtype)},
              open(os.path.join(HERE, 'inference_ref.json'), 'w'))


def gen_trec():
    queries = ['Semantic (entity) retrieval/ranking-2016', '  ontology engineering  ', 'C++ & Java (J2EE) 2.0',
               'naïve Ünïcode café', 'web-services/SOAP (xml (nested)) end', 'tab\tseparated\nlines', '',
               'e-mail <tag> a/b', '123 456']
    parsed = [trec_utils.parse_query(q) for q in queries]
    topics_text = 'EX1;first topic\n\nEX2;second; with delimiter\nEX1;first topic again\n'
    topics = trec_utils.parse_topics([io.StringIO(topics_text)])
    data = {'t1': [(np.float32(0.5), 'a'), (np.float32(0.5), 'b'), (np.float32(0.75), 'c'), (np.float32(0.25), 'd')],
            't2': [(0.125, 'x10'), (0.125, 'x9'), (0.5, 'y')], 't3': []}
    buf = io.StringIO()
    trec_utils.write_run('model_1.bin', data, buf)
    json.dump({'queries': queries, 'parsed': parsed, 'topics_text': topics_text, 'topics': list(topics.items()),
               'run_text': buf.getvalue()}, open(os.path.join(HERE, 'trec_ref.json'), 'w'))


def gen_one_hot():
    rng = np.random.default_rng(404)
    y = synth.make_csr_labels(rng, 40, 13)
    x = rng.integers(0, 99, (40, 3)).astype(np.uint16)
    w = rng.random(40).astype(np.float32)
    new_y, (new_x, new_w) = ref_train.sparse_to_one_hot_multiple(y, x, w)
    out = dict(x=x, w=w, new_y=new_y, new_x=new_x, new_w=new_w)
    out.update(csr_parts(y, 'y'))
    np.savez_compressed(os.path.join(HERE, 'one_hot_ref.npz'), **out)


def gen_prepare():
    """bin/prepare.py:543-599 on a seeded list of (doc_id, window, label) instances, shuffled and not."""
    rng = np.random.default_rng(20160816 + 31)
    n, W, V, n_ent = 240, 5, 70000, 14           # V > 65536: uint32 instances like BASELINE configs[1]
    entity_ids = ['ent-%02d' % i for i in range(n_ent)]
    class_mapping = {e: i for i, e in enumerate(rng.permutation(entity_ids).tolist())}
    docs = ['doc%03d' % i for i in range(40)]
    instances = []
    for i in range(n):
        k = int(rng.choice([1, 2, 3, 5], p=[0.6, 0.2, 0.15, 0.05]))
        ents = rng.choice(n_ent, k, replace=False)
        label = {entity_ids[int(e)]: 1.0 / k for e in ents}
        instances.append((docs[int(rng.integers(0, len(docs)))], tuple(int(v) for v in rng.integers(0, V, W)), label))
    per_doc = {}
    for doc_id, _, _ in instances:
        per_doc[doc_id] = per_doc.get(doc_id, 0) + 1
    max_len = max(per_doc.values())
    out = {'window_size': W, 'num_words': V, 'class_mapping': class_mapping, 'max_document_length': max_len,
           'instances': [[d, list(w), sorted(l.items())] for d, w, l in instances], 'cases': {}}
    for shuffle in (False, True):
        inst = list(instances)
        np.random.seed(4711)
        x, y = ref_prepare.instances_and_labels_to_arrays(inst, W, class_mapping, np.min_scalar_type(V - 1), shuffle)
        w = np.fromiter((float(max_len) / per_doc[doc_id] for doc_id, _, _ in inst), np.float32, len(inst))
        out['cases']['shuffle' if shuffle else 'ordered'] = {
            'x': x.tolist(), 'x_dtype': str(x.dtype), 'indptr': y.indptr.tolist(), 'indices': y.indices.tolist(),
            'data': [float(v) for v in y.data], 'y_dtype': str(y.dtype), 'indices_dtype': str(y.indices.dtype),
            'shape': list(y.shape), 'w': [float(v) for v in w]}
    # cvangysel io_utils.windowed_translated_token_stream (io_utils.py:151-211) on seeded token streams: out-of-vocabulary
    # tokens, end-of-sentence tokens, every (window, stride) pair up to 6, with and without padding
    import collections
    import random
    from cvangysel import io_utils as ref_io
    Word = collections.namedtuple('Word', ['id', 'count'])
    vocab = ['</s>', '<pad>'] + ['w%d' % i for i in range(25)]
    words = {t: Word(i, 1) for i, t in enumerate(vocab)}
    rnd = random.Random(20160816)
    cases = []
    for trial in range(120):
        n = rnd.choice([0, 1, 2, 3, 5, 8, 13, 21, 34])
        toks = ['</s>' if (trial % 3 == 0 and rnd.random() < 0.15) else rnd.choice(vocab[2:] + ['oov-a', 'oov-b'])
                for _ in range(n)]
        W = rnd.randint(1, 6)
        stride = rnd.randint(1, W)
        pad = rnd.choice([None, '<pad>'])
        windows = list(ref_io.windowed_translated_token_stream(iter(toks), W, words, eos_chars=[], stride=stride,
                                                                padding_token=pad))
        cases.append({'tokens': toks, 'window_size': W, 'stride': stride, 'padding_token': pad,
                      'windows': [list(w) for w in windows]})
    out['window_vocab'] = vocab
    out['window_cases'] = cases
    with open(os.path.join(HERE, 'prepare_ref.json'), 'w') as f:
        json.dump(out, f)


def gen_query_callbacks():
    """The reference's ranking callbacks on synthetic predict_fn outputs."""
    rng = np.random.default_rng(505)

    class Args(object):
        top = 10

    class ModelArgs(object):
        entity_representation_size = 12

    tokens = ['w%d' % i for i in range(30)]
    # ---- log-linear: per-term distributions (T, E) ----
    E = 25
    ll_cases, ll_out = [], []
    for T in (1, 3, 6):
        d = rng.random((T, E)).astype(np.float32) ** 4
        d /= d.sum(axis=1, keepdims=True)
        ll_cases.append(d)
    ranked = {}
    cb = ref_query.LogLinearCallback(Args(), ModelArgs(), tokens, io.StringIO(),
                                     lambda topic_id, idx, val: ranked.__setitem__(topic_id, (idx, val)))
    for i, d in enumerate(ll_cases):
        cb(list(range(d.shape[0])), d.copy(), topic_id='L%d' % i)
        ll_out.append(ranked['L%d' % i])
    debug_text = cb.f_debug_out.getvalue()
    # ---- vector space: entity matrix (E2, de) and query projections ----
    E2, de = 300, 12
    ents = rng.standard_normal((E2, de)).astype(np.float32) * rng.uniform(0.5, 3.0, (E2, 1)).astype(np.float32)
    projections = np.tanh(rng.standard_normal((8, de))).astype(np.float32)
    vs = {}
    for name, top in (('top10', 10), ('all', None)):
        Args.top = top
        ranked = {}
        cbv = ref_query.VectorSpaceCallback(ents.copy(), Args(), ModelArgs(), tokens, io.StringIO(),
                                            lambda topic_id, idx, val: ranked.__setitem__(topic_id, (idx, val)))
        for i, p in enumerate(projections):
            cbv([1, 2], p.copy().reshape(1, -1) if i % 2 else p.copy(), topic_id='V%d' % i)
        vs[name + '_idx'] = np.stack([np.asarray(ranked['V%d' % i][0], dtype=np.int64) for i in range(8)])
        vs[name + '_val'] = np.stack([np.asarray(ranked['V%d' % i][1]) for i in range(8)])
        vs[name + '_val_dtype'] = np.array(str(np.asarray(ranked['V0'][1]).dtype))
    out = dict(entities=ents, projections=projections, ll_debug=np.array(debug_text), **vs)
    for i, (d, (idx, val)) in enumerate(zip(ll_cases, ll_out)):
        out['ll_dist%d' % i], out['ll_idx%d' % i], out['ll_val%d' % i] = d, np.asarray(idx), np.asarray(val)
    np.savez_compressed(os.path.join(HERE, 'query_ref.npz'), **out)


if __name__ == '__main__':
    gen_loglinear()
    gen_vectorspace()
    gen_inference()
    gen_trec()
    gen_one_hot()
    gen_prepare()
    gen_query_callbacks()
    print('golden fixtures written to', HERE)
    for f in sorted(os.listdir(HERE)):
        print('  %-24s %8d bytes' % (f, os.path.getsize(os.path.join(HERE, f))))

"""Pins the oracle and the host-side mirrors to fixtures produced by the reference's OWN code
(sert/models.py under the Theano/Lasagne shim, sert/inference.py, bin/query.py callbacks, bin/train.py,
cvangysel trec_utils) -- see tests/golden/make_golden.py."""
import io

import numpy as np
import pytest

from oracle import sert_oracle as O
from tests import golden_io as G
from tests.helpers import close


def assert_mean_std(got, ref, rtol=5e-6):
    """(mean, std) over batches: the std of nearly equal float32 losses only carries absolute precision."""
    np.testing.assert_allclose(got[0], ref[0], rtol=rtol)
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-3, atol=2e-6 * abs(ref[0]))


class NumpyScorer(object):
    """Test-side stand-in for sert_b200.scoring.EntityScorer (exact float32 inner products on the host) so the
    host logic of VectorSpaceCallback can be checked without a GPU."""

    def __init__(self, entities, normalise=False, max_queries=1024, max_k=128, row_begin=0):
        self.E = np.asarray(entities, dtype=np.float32)

    def scores(self, queries):
        return (queries.astype(np.float32) @ self.E.T).astype(np.float32)

    def topk(self, queries, k):
        s = self.scores(queries)
        idx = np.argsort(-s, axis=1, kind='stable')[:, :k]
        return idx.astype(np.int32), np.take_along_axis(s, idx, axis=1)


def test_oracle_matches_reference_loglinear_model():
    d = G.loglinear()
    orc = O.LogLinearOracle(int(d['B']), d['R0'], d['Wd0'], d['bd0'], float(d['lam']), d['training_set'],
                            d['validation_set'])
    assert_mean_std(orc.error('train'), d['train_error0'])
    assert_mean_std(orc.error('val'), d['validation_error0'])
    losses = [orc.train_batch(int(b)) for b in d['order']]
    np.testing.assert_allclose(losses, d['train_losses'], rtol=2e-6)
    close(orc.R, d['R1'], rtol=1e-5, what='R after 6 Adadelta steps')
    close(orc.Wd, d['Wd1'], rtol=1e-5, what='Wd')
    close(orc.bd, d['bd1'], rtol=1e-5, atol_scale=1e-5, what='bd')
    assert_mean_std(orc.error('train'), d['train_error1'])
    assert_mean_std(orc.error('val'), d['validation_error1'])
    close(O.loglinear_predict(orc.R, orc.Wd, orc.bd, d['predict_batch']), d['predict_out'], rtol=1e-5,
          what='predict_fn')


def test_oracle_matches_reference_vectorspace_model():
    d = G.vectorspace()
    orc = O.VectorSpaceOracle(int(d['B']), d['R0'], d['Wp0'], d['bp0'], d['E0'], float(d['lam']),
                              d['training_set'], d['validation_set'])
    got = [orc.eval_batch('train', b, d['test_negs0'][b]) for b in range(6)]
    np.testing.assert_allclose(got, d['test_losses0'], rtol=2e-6)
    got = [orc.eval_batch('val', b, d['val_negs0'][b]) for b in range(2)]
    np.testing.assert_allclose(got, d['val_losses0'], rtol=2e-6)
    losses = [orc.train_batch(int(b), d['train_negs'][j]) for j, b in enumerate(d['order'])]
    np.testing.assert_allclose(losses, d['train_losses'], rtol=2e-6)
    for name, ref in (('R', 'R1'), ('Wp', 'Wp1'), ('Eemb', 'E1')):
        close(getattr(orc, name), d[ref], rtol=1e-5, what=name)
    close(orc.bp, d['bp1'], rtol=1e-5, atol_scale=1e-5, what='bp')
    got = [orc.eval_batch('train', b, d['test_negs1'][b]) for b in range(6)]
    np.testing.assert_allclose(got, d['test_losses1'], rtol=5e-6)
    for avg, ref in zip(d['predict_in'], d['predict_out']):
        assert ref.shape == (1, int(d['de']))            # DenseLayer adds b.dimshuffle('x', 0)
        close(O.vectorspace_predict(d['Wp1'], d['bp1'], avg), ref[0], rtol=1e-5, what='predict_fn')


def test_inference_mirror_matches_reference_batcher():
    from sert_b200 import inference, math_utils
    d = G.load_json('inference_ref.json')
    table = np.asarray(d['table'], dtype=np.float32)
    calls = []

    class Recorder(object):
        def __call__(self, payload, result, **kwargs):
            calls.append((list(payload), np.asarray(result), kwargs))

        def should_average_input(self):
            return False

    batcher = inference.create(lambda batch, mask: table[batch.astype(np.int64)], None, d['B'], d['W'], d['V'],
                               Recorder())
    assert str(batcher.batch.dtype) == d['instance_dtype'] and batcher.mask.dtype == np.int8
    for i, q in enumerate(d['queries']):
        batcher.submit(list(q), topic_id='T%d' % i)
    batcher.process()
    assert len(calls) == len(d['calls'])
    for (payload, result, kwargs), ref in zip(calls, d['calls']):
        assert payload == ref['payload'] and kwargs == ref['kwargs']
        np.testing.assert_array_equal(result, np.asarray(ref['result'], dtype=np.float32))
    with pytest.raises(RuntimeError):
        batcher.submit(list(range(13)), topic_id='X')
    dist = np.asarray(d['dist'], dtype=np.float32)
    for mode, ref in d['aggregate'].items():
        np.testing.assert_array_equal(np.asarray(inference.aggregate_distribution(dist, mode, 0)),
                                      np.asarray(ref, dtype=np.asarray(inference.aggregate_distribution(dist, mode, 0)).dtype))
    pk = np.asarray(d['pk'])
    assert math_utils.entropy(pk) == d['entropy']['plain']
    assert math_utils.entropy(pk, base=2, normalize=True) == d['entropy']['base2_norm']
    assert math_utils.entropy(pk, normalize=True) == d['entropy']['norm']


def test_embedding_mapper_batches_but_keeps_order():
    from sert_b200 import inference
    rng = np.random.default_rng(0)
    R = rng.standard_normal((20, 6)).astype(np.float32)
    Wp = rng.standard_normal((6, 4)).astype(np.float32)
    seen = []

    class Cb(object):
        def __call__(self, payload, result, **kw):
            seen.append((payload, result, kw['topic_id']))

        def should_average_input(self):
            return True

    mapper = inference.create(lambda avg: np.tanh(avg @ Wp), R, 8, 3, 20, Cb())
    qs = [[1, 2, 3], [4], [5, 6]]
    for i, q in enumerate(qs):
        mapper.submit(q, topic_id=i)
    mapper.process()
    assert [s[2] for s in seen] == [0, 1, 2]
    for (payload, result, _), q in zip(seen, qs):
        np.testing.assert_allclose(result, np.tanh(R[q].mean(axis=0) @ Wp), rtol=1e-6)


def test_trec_shim_matches_reference():
    from cvangysel import trec_utils
    d = G.load_json('trec_ref.json')
    for q, ref in zip(d['queries'], d['parsed']):
        assert trec_utils.parse_query(q) == ref, q
    topics = trec_utils.parse_topics([io.StringIO(d['topics_text'])])
    assert [list(kv) for kv in topics.items()] == d['topics']
    data = {'t1': [(np.float32(0.5), 'a'), (np.float32(0.5), 'b'), (np.float32(0.75), 'c'), (np.float32(0.25), 'd')],
            't2': [(0.125, 'x10'), (0.125, 'x9'), (0.5, 'y')], 't3': []}
    buf = io.StringIO()
    trec_utils.write_run('model_1.bin', data, buf)
    assert buf.getvalue() == d['run_text']
    # ties break on the id string, descending (sub:trec_utils.py:560-561)
    assert O.write_run_order([(0.5, 'a'), (0.5, 'b')]) == [(0.5, 'b'), (0.5, 'a')]


def test_one_hot_expansion_matches_reference():
    from sert_b200.synth import sparse_to_one_hot_multiple
    d = G.load_npz('one_hot_ref.npz')
    y = G.csr(d, 'y')
    new_y, (new_x, new_w) = sparse_to_one_hot_multiple(y, d['x'], d['w'])
    np.testing.assert_array_equal(new_y, d['new_y'])
    assert new_y.dtype == np.int32 and new_x.dtype == d['new_x'].dtype
    np.testing.assert_array_equal(new_x, d['new_x'])
    np.testing.assert_array_equal(new_w, d['new_w'])
    bad = y.tolil()
    bad[3, :] = 0
    with pytest.raises(RuntimeError):
        sparse_to_one_hot_multiple(bad.tocsr(), d['x'], d['w'])


def _run_callbacks(scorer_factory):
    from sert_b200 import ranking
    d = G.load_npz('query_ref.npz')

    class Args(object):
        top = 10

    class ModelArgs(object):
        entity_representation_size = 12

    tokens = ['w%d' % i for i in range(30)]
    ranked = {}
    sink = lambda topic_id, idx, val: ranked.__setitem__(topic_id, (np.asarray(idx), np.asarray(val)))  # noqa: E731
    debug = io.StringIO()
    cb = ranking.LogLinearCallback(Args(), ModelArgs(), tokens, debug, sink)
    for i in range(3):
        dist = d['ll_dist%d' % i]
        cb(list(range(dist.shape[0])), dist.copy(), topic_id='L%d' % i)
        np.testing.assert_array_equal(ranked['L%d' % i][0], d['ll_idx%d' % i])
        np.testing.assert_array_equal(ranked['L%d' % i][1], d['ll_val%d' % i])
        idx, val = O.loglinear_rank(dist)
        np.testing.assert_array_equal(idx, d['ll_idx%d' % i])
    assert debug.getvalue().split(': <zip')[0] == str(d['ll_debug']).split(': <zip')[0]
    for name, top in (('top10', 10), ('all', None)):
        Args.top = top
        ranked.clear()
        cbv = ranking.VectorSpaceCallback(d['entities'].copy(), Args(), ModelArgs(), tokens, io.StringIO(), sink,
                                          scorer_factory=scorer_factory)
        # half through the reference's per-query path, half through the batched path
        for i in range(4):
            p = d['projections'][i]
            cbv([1, 2], p.copy().reshape(1, -1) if i % 2 else p.copy(), topic_id='V%d' % i)
        cbv.process_many([[1, 2]] * 4, d['projections'][4:].copy(), [{'topic_id': 'V%d' % i} for i in range(4, 8)])
        for i in range(8):
            np.testing.assert_array_equal(ranked['V%d' % i][0], d[name + '_idx'][i])      # ranked lists identical
            np.testing.assert_array_equal(ranked['V%d' % i][1], d[name + '_val'][i])
            assert str(ranked['V%d' % i][1].dtype) == str(d[name + '_val_dtype'])
        E = O.normalise_rows(d['entities'])
        idx, val = O.vectorspace_rank(E, d['projections'][0], top=top)
        np.testing.assert_array_equal(idx, d[name + '_idx'][0])


def test_ranking_callbacks_match_reference_on_host():
    _run_callbacks(NumpyScorer)


def test_embedding_utils_round_trip(tmp_path):
    """word2vec-binary reader used by bin/train.py --representation_initializer (sub:embedding_utils.py:23-88)."""
    from cvangysel import embedding_utils
    rng = np.random.default_rng(0)
    items = [('Alpha', rng.standard_normal(5).astype(np.float32)), ('beta', rng.standard_normal(5).astype(np.float32)),
             ('gamma', rng.standard_normal(5).astype(np.float32))]
    path = str(tmp_path / 'vectors.bin')
    embedding_utils.write_binary_representations(path, items)
    assert embedding_utils.get_binary_representations_info(path) == (3, 5)
    loaded = dict(embedding_utils.load_binary_representations(path))
    assert sorted(loaded) == ['alpha', 'beta', 'gamma']            # words are lower-cased on load
    np.testing.assert_allclose(loaded['alpha'], items[0][1], rtol=0, atol=0)
    only = dict(embedding_utils.load_binary_representations(path, ['beta']))
    assert list(only) == ['beta']


def test_argparse_validators_match_reference_semantics():
    import argparse
    from cvangysel import argparse_utils as A
    assert A.positive_int('0') == 0 and A.positive_int('7') == 7         # the reference's "positive" accepts zero
    assert A.ratio('0.01') == 0.01 and A.positive_float('2.5') == 2.5
    for fn, bad in ((A.positive_int, '-1'), (A.positive_int, 'x'), (A.ratio, '1.5'), (A.positive_float, '0')):
        with pytest.raises(argparse.ArgumentTypeError):
            fn(bad)
    with pytest.raises(argparse.ArgumentTypeError):
        A.existing_file_path('/no/such/file')
    with pytest.raises(argparse.ArgumentTypeError):
        A.nonexisting_file_path(__file__)


def test_write_run_arrays_equals_write_run():
    """the array form of write_run emits the same bytes as the (reference-pinned) dict form,
    ties in relevance broken by descending object id, float32 and float64 relevance values, ragged rows."""
    import collections
    from cvangysel import trec_utils
    rng = np.random.default_rng(7)
    Q, k = 23, 17
    pool = np.array(['ent-%d' % i for i in range(40)] + ['e%s' % ('x' * i) for i in range(1, 6)] + ['Z', 'a', 'B-1'])
    for dtype in (np.float32, np.float64):
        ids = np.stack([rng.choice(pool, k, replace=False) for _ in range(Q)])
        rel = np.round(rng.random((Q, k)), 1).astype(dtype)                 # many ties
        rel[0, :3] = [1e-5, 123456.789, 3.0000001e-12]
        rel[1, :2] = [-0.0, 0.0]
        counts = rng.integers(0, k + 1, Q)
        counts[:3] = [k, k, 0]
        subjects = ['T%02d' % q for q in range(Q)]
        subjects[4] = b'bytes-topic'
        for limit in (10 ** 9, 5):
            data = collections.OrderedDict(
                (subjects[q], [(rel[q, j], ids[q, j].item()) for j in range(counts[q])]) for q in range(Q))
            a, b = io.StringIO(), io.StringIO()
            trec_utils.write_run('model_3.bin', data, a, max_objects_per_query=limit)
            trec_utils.write_run_arrays('model_3.bin', subjects, ids, rel, b, counts=counts,
                                        max_objects_per_query=limit)
            assert a.getvalue() == b.getvalue() and a.getvalue().count('\n') > 50


def test_write_topk_run_from_device_style_arrays():
    import collections
    from cvangysel import trec_utils
    from sert_b200 import ranking
    rng = np.random.default_rng(11)
    E, Q, k = 60, 9, 12
    inv = {i: 'entity/%03d' % (E - i) for i in range(E)}
    idx = np.stack([rng.choice(E, k, replace=False) for _ in range(Q)]).astype(np.int32)
    rel = np.sort(rng.random((Q, k)).astype(np.float32), axis=1)[:, ::-1].copy()
    idx[2, 7:] = -1                                       # a shard shorter than k
    idx[5, :] = -1
    topics = ['t%d' % q for q in range(Q)]
    data = collections.OrderedDict(
        (topics[q], [(rel[q, j], inv[int(idx[q, j])]) for j in range(k) if idx[q, j] >= 0]) for q in range(Q))
    a, b = io.StringIO(), io.StringIO()
    trec_utils.write_run('m.bin', data, a)
    ranking.write_topk_run('m.bin', topics, idx, rel, inv, b)
    assert a.getvalue() == b.getvalue() and a.getvalue().count('\n') == (idx >= 0).sum()


def test_candidate_set_is_verified_and_grown_on_near_ties():
    """more near-ties around the k-th place than the first request's candidate margin.  200 entities lie
    within float32 noise of one another along the query direction, in an order the float32 inner product cannot
    resolve; the k nearest by float64 distance (what the reference's tree search returns) must still come out, which
    needs the second, larger request."""
    import io
    from sert_b200 import ranking
    rng = np.random.default_rng(8)
    d, n, k = 16, 3000, 10
    q = rng.standard_normal(d).astype(np.float32)
    q /= np.linalg.norm(q)
    E = rng.standard_normal((n, d)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1)[:, None]
    ties = rng.choice(n, 200, replace=False)
    for j, row in enumerate(ties):                       # q plus a perturbation at the float32 rounding level
        v = q.astype(np.float64) * (1.0 + 1e-8 * j)
        v[j % d] += 3e-8 * ((j * 7919) % 13 - 6)
        E[row] = v.astype(np.float32)

    class Args(object):
        top = k

    class ModelArgs(object):
        entity_representation_size = d

    calls = []

    class CountingScorer(NumpyScorer):
        def topk(self, queries, kk):
            calls.append(kk)
            return NumpyScorer.topk(self, queries, kk)

    cb = ranking.VectorSpaceCallback(E.copy(), Args(), ModelArgs(), ['w'], io.StringIO(), lambda *a: None,
                                     scorer_factory=CountingScorer)
    En = cb.entity_representations
    dist, idx = cb.query(q.reshape(1, -1).copy())
    diff = En.astype(np.float64) - q.astype(np.float64)
    ref = np.sqrt(np.einsum('ij,ij->i', diff, diff))
    ref_order = np.argsort(ref, kind='stable')[:k]
    np.testing.assert_array_equal(np.sort(ref[idx[0]]), np.sort(ref[ref_order]))     # the same k distances
    assert set(idx[0].tolist()) <= set(ties.tolist())
    assert len(calls) >= 2 and calls[1] > calls[0], calls                              # it had to ask again


def test_run_collector_writes_the_reference_run_files():
    """bin/query.py's ranker_callback + two write_run calls (bin/query.py:83-92,149-156) against RunCollector: same
    bytes for the entity-profiling and the entity-finding run, including relevance ties (broken by descending id),
    float32 and float64 relevances, topics of different lengths and entities first seen late."""
    import collections
    import io
    from cvangysel import trec_utils
    from sert_b200.ranking import RunCollector
    rng = np.random.default_rng(3)
    inv = {i: 'ent-%03d' % ((i * 37) % 50) for i in range(50)}
    calls = []
    for t in range(12):
        n = int(rng.integers(1, 30))
        idx = rng.choice(50, n, replace=False)
        val = (rng.integers(0, 6, n) / 5.0 + (rng.random(n) < 0.5) * rng.random(n) * 1e-3)
        val = val.astype(np.float32) if t % 2 else val.astype(np.float64)
        order = np.argsort(-val, kind='stable')
        calls.append(('topic-%d' % (97 - 7 * t), idx[order], val[order]))
    topics_per_entity, entities_per_topic = collections.defaultdict(list), collections.defaultdict(list)
    collector = RunCollector(inv)
    for topic_id, idx, val in calls:
        collector(topic_id, idx, val)
        for entity_internal_id, relevance in zip(idx, val):          # the reference's ranker_callback
            entity_id = inv[entity_internal_id]
            topics_per_entity[entity_id].append((relevance, topic_id))
            entities_per_topic[topic_id].append((relevance, entity_id))
    ref_ep, ref_ef, got_ep, got_ef = io.StringIO(), io.StringIO(), io.StringIO(), io.StringIO()
    trec_utils.write_run('model', topics_per_entity, ref_ep)
    trec_utils.write_run('model', entities_per_topic, ref_ef)
    collector.write('model', got_ep, got_ef)
    assert got_ef.getvalue() == ref_ef.getvalue()
    assert got_ep.getvalue() == ref_ep.getvalue()

"""The CUDA path (through the C-ABI) against fixtures produced by the reference's own code."""
import numpy as np
import pytest

from tests import golden_io as G
from tests.helpers import close
from tests.test_golden_cpu import assert_mean_std, _run_callbacks

pytestmark = pytest.mark.gpu


def test_loglinear_model_matches_reference_fixture():
    from sert_b200 import models
    d = G.loglinear()
    model = models.LanguageModel(batch_size=int(d['B']), window_size=int(d['W']), representations_init=d['R0'],
                                 output_layer_size=int(d['E']), regularization_lambda=float(d['lam']),
                                 training_set=d['training_set'], validation_set=d['validation_set'],
                                 dense_init=(d['Wd0'], d['bd0']))
    assert_mean_std(model.train_error(), d['train_error0'], rtol=1e-4)
    assert_mean_std(model.validation_error(), d['validation_error0'], rtol=1e-4)
    losses = [model.train_fn(int(b)) for b in d['order']]
    np.testing.assert_allclose(losses, d['train_losses'], rtol=1e-4)
    Wd, bd = model.get_dense()
    close(model.get_representations(), d['R1'], rtol=2e-4, what='R')
    close(Wd, d['Wd1'], rtol=2e-4, what='Wd')
    close(bd, d['bd1'], rtol=2e-4, atol_scale=1e-4, what='bd')
    assert_mean_std(model.train_error(), d['train_error1'], rtol=1e-4)
    got = model.predict_fn(d['predict_batch'], np.ones(d['predict_batch'].shape, np.int8))
    close(got, d['predict_out'], rtol=2e-4, what='predict_fn')


def test_vectorspace_model_matches_reference_fixture():
    from sert_b200 import models
    d = G.vectorspace()
    model = models.VectorSpaceLanguageModel(
        batch_size=int(d['B']), window_size=int(d['W']), num_negative_samples=int(d['k']),
        representations_init=d['R0'], entity_representations_init=d['E0'], regularization_lambda=float(d['lam']),
        training_set=d['training_set'], validation_set=d['validation_set'], dense_init=(d['Wp0'], d['bp0']))
    got = [model.test_fn(b, d['test_negs0'][b]) for b in range(6)]
    np.testing.assert_allclose(got, d['test_losses0'], rtol=1e-4)
    got = [model.validate_fn(b, d['val_negs0'][b]) for b in range(2)]
    np.testing.assert_allclose(got, d['val_losses0'], rtol=1e-4)
    losses = [model.train_fn(int(b), d['train_negs'][j]) for j, b in enumerate(d['order'])]
    np.testing.assert_allclose(losses, d['train_losses'], rtol=1e-4)
    R, Eemb = model.get_representations()
    Wp, bp = model.get_dense()
    close(R, d['R1'], rtol=2e-4, what='R')
    close(Eemb, d['E1'], rtol=2e-4, what='Eemb')
    close(Wp, d['Wp1'], rtol=2e-4, what='Wp')
    close(bp, d['bp1'], rtol=2e-4, atol_scale=1e-4, what='bp')
    got = [model.test_fn(b, d['test_negs1'][b]) for b in range(6)]
    np.testing.assert_allclose(got, d['test_losses1'], rtol=1e-4)
    fn = model.predict_fn
    for avg, ref in zip(d['predict_in'], d['predict_out']):
        out = fn(avg)
        assert out.shape == ref.shape
        close(out, ref, rtol=2e-4, what='predict_fn')


def test_ranking_callbacks_with_device_scorer_match_reference_lists():
    """Ranked entity lists identical to the reference's VectorSpaceCallback (sklearn kd-tree) / LogLinearCallback."""
    from sert_b200.scoring import EntityScorer
    _run_callbacks(EntityScorer)


def test_loglinear_device_ranking_matches_reference_callback():
    """LogLinearCallback through the device ranking (sert_ll_rank_distributions) on the reference's own fixture:
    ranked lists identical to bin/query.py's callback, relevances equal up to the last float32 digits (the reference's
    numpy exp/log are not correctly rounded, so bit equality with any other implementation is not defined)."""
    import io
    from sert_b200 import ranking
    d = G.load_npz('query_ref.npz')
    dists = [d['ll_dist%d' % i] for i in range(3)]
    ranked = {}
    debug = io.StringIO()
    cb = ranking.LogLinearCallback(None, None, ['w%d' % i for i in range(30)], debug,
                                   lambda topic_id, idx, val: ranked.__setitem__(topic_id, (idx, val)))
    for i, result in enumerate(ranking.rank_distributions(dists)):
        cb.process_ranked(list(range(dists[i].shape[0])), *result, topic_id='L%d' % i)
    for i in range(3):
        idx, val = ranked['L%d' % i]
        np.testing.assert_array_equal(idx, d['ll_idx%d' % i])
        np.testing.assert_allclose(val, d['ll_val%d' % i], rtol=3e-6, atol=0)
    # the debug line carries the normalised entropies: same text up to float formatting
    ref_lines = str(d['ll_debug']).strip().split('\n')
    got_lines = debug.getvalue().strip().split('\n')
    for ref, got in zip(ref_lines, got_lines):
        assert ref.split()[:2] == got.split()[:2]
        np.testing.assert_allclose(float(got.split()[2].rstrip(':')), float(ref.split()[2].rstrip(':')), rtol=1e-5)


def test_loglinear_rank_queries_equals_host_callback_on_predict_fn_output():
    """sert_ll_rank_queries (tokens in, ranking out) against the reference-style host path on the SAME device model:
    predict_fn -> (rows, W, E) -> WordBatcher slices -> LogLinearCallback.process.  Includes multi-row queries, an
    E that is no power of two and one above a sort tile, and exact zeros (saturated logits)."""
    import io
    from sert_b200 import inference, models, ranking
    rng = np.random.default_rng(77)
    for V, E, dw, W, gain in [(300, 715, 32, 4, 1.0), (200, 5000, 16, 3, 1.0), (150, 257, 16, 5, 200.0)]:
        R = (rng.standard_normal((V, dw)) * 0.3 * gain).astype(np.float32)
        Wd = (rng.standard_normal((dw, E)) * 0.3).astype(np.float32)
        bd = (rng.standard_normal(E) * 0.1).astype(np.float32)
        fn = models.LogLinearPredictFn(R, Wd, bd, batch_size=16, window_size=W)
        queries = [list(rng.integers(0, V, size=int(n))) for n in rng.integers(1, 3 * W, size=9)]

        class HostOnly(object):                         # hides process_ranked: the reference-style host path
            def __init__(self, cb):
                self.cb = cb

            def __call__(self, payload, result, **kwargs):
                self.cb(payload, result, **kwargs)

        out_host, out_dev = {}, {}
        for kind, out in (('host', out_host), ('dev', out_dev)):
            cb = ranking.LogLinearCallback(
                None, None, ['w%d' % i for i in range(V)], io.StringIO(),
                lambda topic_id, idx, val, out=out: out.__setitem__(topic_id, (np.asarray(idx), np.asarray(val))))
            if kind == 'host':
                cb = HostOnly(cb)
            batcher = inference.WordBatcher(fn, 16, W, np.uint16, cb)
            for i, q in enumerate(queries):
                batcher.submit(q, topic_id='T%d' % i)
            batcher.process()
        assert sorted(out_host) == sorted(out_dev) and len(out_dev) == len(queries)
        for t in out_host:
            hi, hv = out_host[t]
            di, dv = out_dev[t]
            np.testing.assert_allclose(dv, hv, rtol=2e-5, atol=1e-30, err_msg=str((E, t)))
            sep = np.ones(E, bool)
            sep[1:] &= np.abs(np.diff(hv)) > 1e-5 * hv[1:]
            sep[:-1] &= np.abs(np.diff(hv)) > 1e-5 * hv[1:]
            assert (di[sep] == hi[sep]).all(), (E, t)
            assert sorted(di.tolist()) == list(range(E))

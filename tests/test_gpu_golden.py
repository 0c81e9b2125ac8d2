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

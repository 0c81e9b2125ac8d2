"""Parity at the BENCHMARKED shapes (BASELINE.json configs[1], configs[2], configs[3]) rather than at toy sizes:
the oracle on the full-size problem for a few steps, and size-independent properties where the oracle would take
minutes ("no parity check at the benchmarked shapes")."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_configs1_training_steps_match_oracle_at_full_size():
    """V=100k E=50k d=128 W=10 B=4096 k=10: three Adam steps, losses within 1e-4 relative, tables within 2e-4."""
    import bench
    from oracle import sert_oracle as O
    from sert_b200 import models
    cfg = bench.CFG2
    p = bench.make_problem(0, 3)
    model = models.VectorSpaceLanguageModel(
        batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'], representations_init=p['R'],
        entity_representations_init=p['Eemb'], regularization_lambda=cfg['lam'], training_set=p['train'],
        validation_set=p['val'], dense_init=(p['Wp'], p['bp']), loss_slots=64)
    orc = O.VectorSpaceOracle(cfg['B'], p['R'], p['Wp'], p['bp'], p['Eemb'], cfg['lam'], p['train'], p['val'])
    for j in range(3):
        H.close(model.train_fn(j, p['neg'][j]), orc.train_batch(j, p['neg'][j]), what='loss of step %d' % j)
    R, Eemb = model.get_representations()
    H.close(R, orc.R, rtol=2e-4, what='R')
    H.close(Eemb, orc.Eemb, rtol=2e-4, what='Eemb')


def _unit_rows(rng, n, d):
    a = rng.standard_normal((n, d), dtype=np.float32)
    a /= np.linalg.norm(a, axis=1)[:, None]
    return a


@pytest.mark.parametrize('E,d', [(50000, 128), (1000000, 256)])
def test_scoring_at_full_size_is_the_exact_ranking(E, d):
    """configs[2] / configs[3] shapes with all 10 000 queries in one call: (1) for a sample of queries the list equals
    the float64 brute-force ranking; (2) for EVERY query the list is sorted, has k distinct rows, and its scores equal
    the float32 inner products of the rows it names; (3) the one-launch seeded sweep answered (no fall-back)."""
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(20160816 + E)
    ent, qs = _unit_rows(rng, E, d), _unit_rows(rng, 10000, d)
    k = 100
    sc = EntityScorer(ent, max_queries=10000, max_k=128)
    idx, score = sc.topk(qs, k)
    assert sc.stats() == (1, 0)
    sample = rng.choice(10000, 24, replace=False)
    exact = qs[sample].astype(np.float64) @ ent.astype(np.float64).T
    order = np.argsort(-exact, axis=1, kind='stable')[:, :k]
    ex = np.take_along_axis(exact, order, axis=1)
    np.testing.assert_allclose(score[sample], ex, rtol=0, atol=2e-6)
    sep = np.abs(np.diff(ex, axis=1)).min(axis=1) > 1e-6
    assert sep.sum() >= 8 and (idx[sample][sep] == order[sep]).all()
    assert (np.diff(score, axis=1) <= 0).all()
    assert (np.sort(idx, axis=1)[:, 1:] != np.sort(idx, axis=1)[:, :-1]).all()
    rows = rng.choice(10000, 512, replace=False)
    recomputed = np.einsum('qkd,qd->qk', ent[idx[rows]], qs[rows])
    np.testing.assert_allclose(score[rows], recomputed, rtol=0, atol=2e-6)
    sc.close()

"""Device top-k scoring vs exact CPU ranking (ranked entity lists identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def exact_topk(E, q, k):
    s = (q.astype(np.float64) @ E.astype(np.float64).T)
    order = np.lexsort((np.arange(E.shape[0])[None, :].repeat(q.shape[0], 0), -s), axis=1)[:, :k]
    return order, np.take_along_axis(s, order, axis=1)


@pytest.mark.parametrize('mode', ['tensor', 'fma'])
@pytest.mark.parametrize('rows,d,Q,k', [(5000, 128, 37, 100), (12345, 256, 130, 10), (300, 64, 5, 100),
                                        (64, 32, 3, 64), (9000, 20, 65, 128), (20000, 300, 257, 100)])
def test_topk_matches_exact(rows, d, Q, k, mode):
    from oracle import sert_oracle as O
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(rows + d)
    E = O.normalise_rows(rng.standard_normal((rows, d)))
    q = O.normalise_rows(rng.standard_normal((Q, d)))
    sc = EntityScorer(E, normalise=False, max_queries=64, max_k=128)
    sc.set_mode(mode)
    idx, score = sc.topk(q, k)
    ref_idx, ref_score = exact_topk(E, q, k)
    np.testing.assert_allclose(score, ref_score, rtol=0, atol=2e-6)
    # identical ranked lists wherever the exact scores are separated by more than fp32 noise
    gaps = np.abs(np.diff(ref_score, axis=1)).min(axis=1) > 1e-6
    assert gaps.mean() > 0.5
    assert (idx[gaps] == ref_idx[gaps]).all()
    assert all(len(set(r.tolist())) == k for r in idx)


def test_normalisation_on_device_and_row_offset():
    from oracle import sert_oracle as O
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(5)
    E = rng.standard_normal((4100, 48)).astype(np.float32) * 3
    q = rng.standard_normal((9, 48)).astype(np.float32) * 7
    sc = EntityScorer(E, normalise=True, max_queries=16, max_k=16, row_begin=1000)
    idx, score = sc.topk(q, 16, normalise_queries=True)
    ref_idx, ref_score = exact_topk(O.normalise_rows(E), O.normalise_rows(q), 16)
    np.testing.assert_allclose(score, ref_score, atol=3e-6)
    assert (idx == ref_idx + 1000).mean() > 0.99


def test_ties_break_on_lower_row_id_and_short_shards():
    from sert_b200.scoring import EntityScorer
    E = np.zeros((10, 8), np.float32)
    E[:, 0] = 1.0                      # all rows identical: every score ties
    q = np.zeros((2, 8), np.float32)
    q[:, 0] = 1.0
    sc = EntityScorer(E, max_queries=4, max_k=16)
    idx, score = sc.topk(q, 16)
    assert (idx[:, :10] == np.arange(10)).all() and (idx[:, 10:] == -1).all()
    assert np.isneginf(score[:, 10:]).all()


def test_merge_of_shards_equals_single_device():
    import torch
    from oracle import sert_oracle as O
    from sert_b200 import _native as N
    from sert_b200.scoring import EntityScorer, shard_bounds
    rng = np.random.default_rng(77)
    E = O.normalise_rows(rng.standard_normal((7001, 64)))
    q = O.normalise_rows(rng.standard_normal((33, 64)))
    k, parts = 50, 4
    full_idx, full_score = EntityScorer(E, max_queries=64, max_k=64).topk(q, k)
    idxs, scores = [], []
    for r in range(parts):
        b, e = shard_bounds(E.shape[0], parts, r)
        i, s = EntityScorer(E[b:e], max_queries=64, max_k=64, row_begin=b).topk(q, k)
        idxs.append(i)
        scores.append(s)
    gi = torch.from_numpy(np.stack(idxs)).cuda()
    gs = torch.from_numpy(np.stack(scores)).cuda()
    oi = torch.empty((33, k), dtype=torch.int32, device='cuda')
    os_ = torch.empty((33, k), dtype=torch.float32, device='cuda')
    N.check(N.load().sert_topk_merge_dev(N.dev_ptr(gi), N.dev_ptr(gs), parts, 33, k, N.dev_ptr(oi), N.dev_ptr(os_),
                                         N.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert (oi.cpu().numpy() == full_idx).all()
    np.testing.assert_array_equal(os_.cpu().numpy(), full_score)


@pytest.mark.parametrize('mode', ['tensor', 'fma'])
def test_adversarial_ascending_scores_take_the_exact_fallback(mode):
    """Scores that grow with the row id defeat the optimistic threshold (every later row survives): the
    candidate lists overflow and the sweep must fall back to the conservative, overflow-free pass."""
    from sert_b200.scoring import EntityScorer
    rows, d, k = 30000, 16, 10
    rng = np.random.default_rng(3)
    v = rng.standard_normal(d).astype(np.float32)
    v /= np.linalg.norm(v)
    E = (np.linspace(0.1, 1.0, rows, dtype=np.float32)[:, None] * v[None, :]).astype(np.float32)
    q = np.stack([v, -v, 2 * v]).astype(np.float32)
    sc = EntityScorer(E, max_queries=8, max_k=16)
    sc.set_mode(mode)
    idx, score = sc.topk(q, k)
    ref_idx, ref_score = exact_topk(E, q, k)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_allclose(score, ref_score, rtol=1e-6, atol=1e-6)


def test_randomised_shapes_match_exact():
    """Random (rows, d, Q, k) including rows < k, d that is no multiple of 4 / 64, Q above max_queries (chunked
    calls) and k at the scorer's max_k; both arithmetic modes."""
    from oracle import sert_oracle as O
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(2026)
    for trial in range(10):
        rows = int(rng.choice([1, 7, 63, 200, 1500, 9000, 40000]))
        d = int(rng.choice([3, 10, 50, 64, 100, 128, 257]))
        Q = int(rng.integers(1, 90))
        k = int(rng.choice([1, 5, 32, 100]))
        E = O.normalise_rows(rng.standard_normal((rows, d)))
        q = O.normalise_rows(rng.standard_normal((Q, d)))
        sc = EntityScorer(E, max_queries=32, max_k=100)
        sc.set_mode('tensor' if trial % 2 == 0 else 'fma')
        idx, score = sc.topk(q, k)
        kk = min(k, rows)
        ref_idx, ref_score = exact_topk(E, q, kk)
        np.testing.assert_allclose(score[:, :kk], ref_score, rtol=0, atol=3e-6, err_msg=str((rows, d, Q, k)))
        assert (idx[:, kk:] == -1).all()
        gaps = np.abs(np.diff(ref_score, axis=1)).min(axis=1) > 2e-6 if kk > 1 else np.ones(Q, bool)
        assert (idx[gaps][:, :kk] == ref_idx[gaps]).all(), (rows, d, Q, k)
        sc.close()

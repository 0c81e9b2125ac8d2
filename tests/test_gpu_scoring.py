"""Device top-k scoring vs exact CPU ranking (ranked entity lists identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def exact_topk(E, q, k):
    s = (q.astype(np.float64) @ E.astype(np.float64).T)
    order = np.lexsort((np.arange(E.shape[0])[None, :].repeat(q.shape[0], 0), -s), axis=1)[:, :k]
    return order, np.take_along_axis(s, order, axis=1)


@pytest.mark.parametrize('mode', ['tensor', 'tensor_chunked', 'tensor3', 'fma'])
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


@pytest.mark.parametrize('mode', ['tensor', 'tensor_chunked', 'tensor3', 'fma'])
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


def test_coarse_sweep_keeps_the_exact_top_k_among_near_ties():
    """Rows whose exact scores differ by far less than the bf16 rounding error of the coarse sweep (1e-5 apart around
    the k-th place): the margin must keep all of them for the fp32 re-scoring.  A second matrix has 6000 rows 1e-6
    apart, all inside the margin band and more than a list holds: the coarse sweep overflows and the bf16x3 sweep
    answers."""
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(12)
    d, k = 64, 20
    for band, step in ((300, 1e-5), (6000, 1e-6)):
        rows = 20000
        q = rng.standard_normal((4, d)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1)[:, None]
        E = rng.standard_normal((rows, d)).astype(np.float32)
        E /= np.linalg.norm(E, axis=1)[:, None] * 2.0            # background: scores in (-0.5, 0.5)
        # a band of rows almost parallel to q[0], scores 0.9 + j * 1e-5 in shuffled row positions
        pos = rng.choice(rows, band, replace=False)
        noise = rng.standard_normal((band, d)).astype(np.float32)
        noise -= (noise @ q[0])[:, None] * q[0][None, :]
        noise /= np.linalg.norm(noise, axis=1)[:, None]
        s = (0.9 + step * np.arange(band)).astype(np.float32)
        E[pos] = s[:, None] * q[0][None, :] + np.sqrt(np.maximum(0, 0.98 - s * s))[:, None] * noise
        sc = EntityScorer(E, max_queries=8, max_k=32)
        idx, score = sc.topk(q, k)
        ref_idx, ref_score = exact_topk(E, q, k)
        np.testing.assert_allclose(score, ref_score, rtol=0, atol=2e-6)
        gaps = np.abs(np.diff(ref_score, axis=1)).min(axis=1) > 1e-6
        assert (idx[gaps] == ref_idx[gaps]).all()
        assert set(idx[0].tolist()) <= set(pos.tolist())
        # whatever the order inside fp32 noise, the returned SET is the exact top k up to score ties below 2e-6
        kth = ref_score[0, -1]
        exact0 = (q[0].astype(np.float64) @ E.astype(np.float64).T)
        assert (exact0[idx[0]] >= kth - 2e-6).all()
        sc.close()


def test_randomised_shapes_match_exact():
    """Random (rows, d, Q, k) including rows < k, d that is no multiple of 4 / 64, Q above max_queries (chunked
    calls) and k at the scorer's max_k; both arithmetic modes."""
    from oracle import sert_oracle as O
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(2026)
    for trial in range(16):
        rows = int(rng.choice([1, 7, 63, 200, 1500, 9000, 40000]))
        d = int(rng.choice([3, 10, 50, 64, 100, 128, 257]))
        Q = int(rng.integers(1, 90))
        k = int(rng.choice([1, 5, 32, 100]))
        E = O.normalise_rows(rng.standard_normal((rows, d)))
        q = O.normalise_rows(rng.standard_normal((Q, d)))
        sc = EntityScorer(E, max_queries=32, max_k=100)
        sc.set_mode(['tensor', 'fma', 'tensor3', 'tensor_chunked'][trial % 4])
        idx, score = sc.topk(q, k)
        kk = min(k, rows)
        ref_idx, ref_score = exact_topk(E, q, kk)
        np.testing.assert_allclose(score[:, :kk], ref_score, rtol=0, atol=3e-6, err_msg=str((rows, d, Q, k)))
        assert (idx[:, kk:] == -1).all()
        gaps = np.abs(np.diff(ref_score, axis=1)).min(axis=1) > 2e-6 if kk > 1 else np.ones(Q, bool)
        assert (idx[gaps][:, :kk] == ref_idx[gaps]).all(), (rows, d, Q, k)
        sc.close()


def test_seeded_sweep_answers_ordinary_shards_and_agrees_with_the_chunked_sweep():
    """The one-launch sweep (strided sample -> seeded thresholds -> one GEMM -> finalize) answers random shards
    without falling back, and returns bit-identical lists and scores to the chunked coarse sweep."""
    from oracle import sert_oracle as O
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(99)
    for rows, d, Q, k in [(50000, 128, 300, 100), (6250, 128, 200, 100), (130000, 256, 64, 100), (20000, 64, 50, 10)]:
        E = O.normalise_rows(rng.standard_normal((rows, d)))
        q = O.normalise_rows(rng.standard_normal((Q, d)))
        sc = EntityScorer(E, max_queries=512, max_k=128)
        idx, score = sc.topk(q, k)
        assert sc.stats() == (1, 0), (rows, d, sc.stats())
        sc.set_mode('tensor_chunked')
        idx2, score2 = sc.topk(q, k)
        np.testing.assert_array_equal(idx, idx2)
        np.testing.assert_array_equal(score, score2)
        ref_idx, ref_score = exact_topk(E, q, k)
        np.testing.assert_allclose(score, ref_score, rtol=0, atol=2e-6)
        gaps = np.abs(np.diff(ref_score, axis=1)).min(axis=1) > 1e-6
        assert (idx[gaps] == ref_idx[gaps]).all()
        sc.close()


def test_seeded_sweep_falls_back_when_the_sample_misleads():
    """50 rows almost parallel to the query sit exactly in the sampled tiles, one per sample group: the seeded
    threshold lands just below their score, fewer than k rows survive the sweep, and the call must fall back to the
    chunked sweep and still return the exact top k."""
    from oracle import sert_oracle as O
    from sert_b200.scoring import EntityScorer
    rng = np.random.default_rng(7)
    rows, d, k = 20000, 64, 100
    E = O.normalise_rows(rng.standard_normal((rows, d))) * np.float32(0.5)
    q = O.normalise_rows(rng.standard_normal((3, d)))
    sc = EntityScorer(E, max_queries=8, max_k=128)
    plan = sc.plan(k)
    assert plan['group_rows'] > 0 and plan['rank'] < 50 < k
    per_tile = 256 // plan['group_rows']
    for i in range(50):                      # one such row in each of 50 sample groups
        tile, group = divmod(i, per_tile)
        E[tile * plan['tile_stride'] * 256 + group * plan['group_rows']] = q[0] * np.float32(0.9)
    sc.close()
    sc = EntityScorer(E, max_queries=8, max_k=128)
    idx, score = sc.topk(q, k)
    assert sc.stats() == (0, 1), sc.stats()
    ref_idx, ref_score = exact_topk(E, q, k)
    np.testing.assert_allclose(score, ref_score, rtol=0, atol=2e-6)
    gaps = np.abs(np.diff(ref_score, axis=1)).min(axis=1) > 1e-6
    assert (idx[gaps] == ref_idx[gaps]).all()


def test_coarse_margin_covers_bf16_rounding_midpoints():
    """operands next to bf16 rounding midpoints make the coarse (hi.hi) score err by the full unit
    roundoff 2^-8 per operand.  Family A rows (and the query) sit just below a midpoint and round DOWN: exact score
    1.0073, coarse 1.0.  Family B rows alternate an element that rounds UP with an exactly representable one:
    exact 1.0038, coarse 1.00195.  The coarse order is the reverse of the exact one, so without a margin of the full
    error bound the coarse top k would be all B; the exact top k is all A."""
    from sert_b200.scoring import EntityScorer
    d, k, rows = 64, 20, 12000
    rng = np.random.default_rng(4)
    base = np.float32(1.0 / 8.0)
    down = np.float32(1.0 + 2.0 ** -8 - 2.0 ** -12)
    q = np.full((1, d), down * base, np.float32)
    E = (rng.standard_normal((rows, d)) * 0.01).astype(np.float32)     # background, scores near 0
    pos = rng.choice(rows, 60, replace=False)
    E[pos[:30]] = down * base
    fam_b = np.empty(d, np.float32)
    fam_b[0::2] = np.float32(1.0 + 2.0 ** -8 + 2.0 ** -12) * base
    fam_b[1::2] = np.float32(1.0 - 2.0 ** -8) * base
    E[pos[30:]] = fam_b
    sc = EntityScorer(E, max_queries=4, max_k=32)
    for mode in ('tensor', 'tensor_chunked'):
        sc.set_mode(mode)
        idx, score = sc.topk(q, k)
        ref_idx, ref_score = exact_topk(E, q, k)
        np.testing.assert_allclose(score, ref_score, rtol=0, atol=2e-6)
        assert set(idx[0].tolist()) <= set(pos[:30].tolist()), mode

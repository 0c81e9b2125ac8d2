"""tcgen05 / TMEM / TMA GEMM (csrc/gemm_tc.cu) against float64 numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(a, b, terms, bias=None):
    from sert_b200 import _native as N
    lib = N.load()
    m, k = a.shape
    n = b.shape[0]
    c = np.empty((m, n), np.float32)
    N.check(lib.sert_debug_gemm_tc(N.host_ptr(a), N.host_ptr(b), m, n, k, terms,
                                   N.host_ptr(bias) if bias is not None else None, N.host_ptr(c)))
    return c


@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (128, 256, 128), (256, 512, 64), (100, 300, 70), (1, 9, 5),
                                   (1000, 2000, 128), (333, 715, 300), (4096, 777, 256)])
def test_split3_matches_float64(m, n, k):
    rng = np.random.default_rng(m * 31 + n * 7 + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((n, k)).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    got = run(a, b, 3, bias)
    scale = np.sqrt(k)
    err = np.abs(got - ref).max()
    assert err < 3e-5 * scale, err            # bf16x3: ~2^-16 relative per product
    got1 = run(a, b, 1)
    err1 = np.abs(got1 - (ref - bias)).max()
    assert err1 < 2e-2 * scale, err1          # plain bf16: ~2^-8 relative per product


@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (256, 512, 128), (100, 300, 70), (1, 9, 5), (1000, 2000, 128),
                                   (333, 715, 300), (4096, 777, 256), (2048, 1000, 192), (1152, 513, 1000),
                                   (8300, 257, 128), (1280, 2000, 320)])
def test_pair_operands_match_float64_and_the_three_block_layout(m, n, k):
    """Pair operands [hi | mid] (launch_gemm_tc_pair: two ring stages per 64-column block, hi.hi + hi.mid + mid.hi
    formed by the MMA warp) compute the same three products as the [hi|hi|mid] x [hi|mid|hi] layout: same error
    against float64, and the two agree to summation-order noise; single-CTA tiles and clusters (m-tiles >= 8), ragged
    M / N / K, K of 1 to 16 blocks."""
    rng = np.random.default_rng(m * 13 + n * 5 + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((n, k)).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    got2 = run(a, b, 2, bias)
    got3 = run(a, b, 3, bias)
    scale = np.sqrt(k)
    assert np.abs(got2 - ref).max() < 3e-5 * scale
    assert np.abs(got2 - got3).max() < 3e-5 * scale      # (a missing cross term would show as ~2e-3 * scale)


def test_pair_operands_identity_layout():
    """Exact check of the pair path's operand placement (hi stage / mid stage, clusters): values whose bf16 split has
    a non-zero mid term, one non-zero per row, so every output is a single product hi.hi + hi.mid + mid.hi."""
    for m in (256, 1152):
        n, k = 768, 192
        a = np.zeros((m, k), np.float32)
        b = np.zeros((n, k), np.float32)
        for i in range(m):
            a[i, i % k] = np.float32(1.0 + (i % 97) / 1024.0 + 2.0 ** -12)
        for j in range(n):
            b[j, (j * 5) % k] = np.float32(0.5 + (j % 89) / 512.0 + 2.0 ** -11)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        got = run(a, b, 2)
        np.testing.assert_allclose(got, ref, rtol=3e-5, atol=0)
        assert np.array_equal(got == 0, ref == 0)


def run_bn(a, bt):
    from sert_b200 import _native as N
    lib = N.load()
    m, k = a.shape
    n = bt.shape[1]
    c = np.empty((m, n), np.float32)
    N.check(lib.sert_debug_gemm_tc_bn(N.host_ptr(a), N.host_ptr(bt), m, n, k, N.host_ptr(c)))
    return c


@pytest.mark.parametrize('m,n,k', [(128, 256, 64), (301, 2000, 1024), (100, 300, 70), (1, 9, 5), (301, 715, 10240),
                                   (64, 777, 256), (301, 64, 128), (129, 513, 192)])
def test_n_major_b_operand_matches_float64(m, n, k):
    """B given as its transpose (k, n): the kernel loads 64 x 64 boxes of the row-major (K, N) matrix and multiplies
    through an MN-major shared-memory descriptor (launch_gemm_tc_pair_bn -- how gWd = X^T . dZ reads dZ's rows without a
    transposed copy).  Same bound as the K-major pair path; ragged M, N (last tile narrower than 256) and K."""
    rng = np.random.default_rng(m * 3 + n * 11 + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    bt = rng.standard_normal((k, n)).astype(np.float32)
    ref = a.astype(np.float64) @ bt.astype(np.float64)
    got = run_bn(a, bt)
    # The second term: tcgen05 adds into its fp32 accumulator by TRUNCATION, so a sum drifts toward zero by up to half
    # an ulp of the accumulator per MMA step -- at K = 10240 (1 920 steps on sums of magnitude 100-200) every output
    # comes out 1e-3 to 1.7e-2 smaller in magnitude than the float64 product, the same through the K-major path.
    assert np.abs(got - ref).max() < 5e-5 * np.sqrt(k) + 2e-6 * k
    got_k = run(a, np.ascontiguousarray(bt.T), 2)
    assert np.abs(got - got_k).max() < 1e-5 * np.sqrt(k)      # same products in the same order: only the operand path differs


def test_n_major_b_identity_layout():
    """One non-zero per row of A and per column of B^T with a non-zero mid term: every output is a single product, so
    a box landing at the wrong offset or a wrong descriptor stride shows up exactly."""
    m, n, k = 256, 768, 192
    a = np.zeros((m, k), np.float32)
    bt = np.zeros((k, n), np.float32)
    for i in range(m):
        a[i, i % k] = np.float32(1.0 + (i % 97) / 1024.0 + 2.0 ** -12)
    for j in range(n):
        bt[(j * 5) % k, j] = np.float32(0.5 + (j % 89) / 512.0 + 2.0 ** -11)
    ref = a.astype(np.float64) @ bt.astype(np.float64)
    got = run_bn(a, bt)
    np.testing.assert_allclose(got, ref, rtol=3e-5, atol=0)
    assert np.array_equal(got == 0, ref == 0)


def test_identity_layout():
    """Each output column picks one B row: catches swizzle / descriptor / TMEM lane mistakes exactly."""
    m, n, k = 256, 512, 128
    a = np.zeros((m, k), np.float32)
    b = np.zeros((n, k), np.float32)
    for i in range(m):
        a[i, i % k] = 1.0 + i
    for j in range(n):
        b[j, (j * 3) % k] = 0.5 + j
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    got = run(a, b, 3)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-3)


def test_cluster_multicast_layout_is_exact():
    """Two-CTA clusters with a multicast B tile (m-tiles >= 8): every output column picks one B row and every output row
    one A element, so a half-tile landing at the wrong offset, in the wrong CTA or with the wrong swizzle shows up
    exactly; an odd number of m-tiles leaves the last cluster's second CTA without rows."""
    for m in (1024, 1152, 4224):
        n, k = 768, 128
        a = np.zeros((m, k), np.float32)
        b = np.zeros((n, k), np.float32)
        for i in range(m):
            a[i, i % k] = 1.0 + (i % 97)
        for j in range(n):
            b[j, (j * 5) % k] = 0.5 + (j % 89)
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        got = run(a, b, 3)
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('m,n,k', [(2048, 1000, 192), (4200, 300, 640), (1100, 513, 1000), (1024, 256, 64),
                                   (8300, 257, 128), (1280, 2000, 320)])
def test_cluster_multicast_equals_single_cta_tiles(m, n, k):
    """The cluster schedules (pairs by default, quads with SERT_GEMM_CLUSTER=4) change where operand tiles come
    from, not the arithmetic: bit-identical outputs with SERT_GEMM_CLUSTER=0 (16 / 33 / 9 / 8 / 65 / 10 m-tiles; ragged
    n; K of 1 to 16 blocks)."""
    import os
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((n, k)).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    got = run(a, b, 3, bias)
    try:
        os.environ['SERT_GEMM_CLUSTER'] = '0'
        plain = run(a, b, 3, bias)
        os.environ['SERT_GEMM_CLUSTER'] = '4'
        quads = run(a, b, 3, bias)
    finally:
        del os.environ['SERT_GEMM_CLUSTER']
    np.testing.assert_array_equal(got, plain)
    np.testing.assert_array_equal(quads, plain)
    ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
    assert np.abs(got - ref).max() < 6e-5 * np.sqrt(k)

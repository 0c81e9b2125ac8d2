"""Synthetic (word-window, entity) corpora shaped like the output of the reference's bin/prepare.py.

Used by tests and bench.py (there is no network / corpus in the build environment).  The arrays
follow the data.npz contract (bin/prepare.py:383-416, 543-599): ``x (N,W)`` in
``np.min_scalar_type(V-1)``, ``y`` CSR float32 ``(N,E)`` whose rows sum to 1
(bin/prepare.py:519-523, 593-597), ``w (N,)`` float32 instance weights (bin/prepare.py:395-403).
"""
import numpy as np
import scipy.sparse as sparse


def zipf_probs(n, s=1.07, shift=2.7):
    r = np.arange(n, dtype=np.float64)
    p = 1.0 / np.power(r + shift, s)
    return p / p.sum()


def sample_zipf(rng, n, size, s=1.07, shift=2.7):
    cdf = np.cumsum(zipf_probs(n, s, shift))
    cdf[-1] = 1.0
    return np.searchsorted(cdf, rng.random(size), side='right').astype(np.int64)


def make_windows(rng, num_instances, window, vocab):
    dtype = np.min_scalar_type(vocab - 1)
    return sample_zipf(rng, vocab, (num_instances, window)).astype(dtype)


def make_csr_labels(rng, num_instances, entities):
    """1-3 labels per row with probabilities .8/.15/.05, Zipf(1.0) entities, mass 1/n each."""
    counts = rng.choice([1, 2, 3], size=num_instances, p=[0.8, 0.15, 0.05])
    indptr = np.zeros(num_instances + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    cols = sample_zipf(rng, entities, nnz, s=1.0, shift=1.0).astype(np.int32)
    # distinct, sorted columns within a row (scipy canonical form)
    for i in np.nonzero(counts > 1)[0]:
        lo, hi = indptr[i], indptr[i + 1]
        seg = cols[lo:hi]
        while len(set(seg.tolist())) < len(seg):
            seg = sample_zipf(rng, entities, len(seg), s=1.0, shift=1.0).astype(np.int32)
        cols[lo:hi] = np.sort(seg)
    data = (1.0 / np.repeat(counts, counts)).astype(np.float32)
    idx_dtype = np.int32
    return sparse.csr_matrix((data, cols.astype(idx_dtype), indptr.astype(idx_dtype if nnz < 2**31 else np.int64)),
                             shape=(num_instances, entities))


def make_weights(rng, num_instances):
    """bin/prepare.py:395-399: w = max_len / len_doc (f32), document lengths ~ U{1..50}."""
    lens = rng.integers(1, 51, size=num_instances)
    return (50.0 / lens).astype(np.float32)


def sparse_to_one_hot_multiple(y, *matrices):
    """Vectorised equivalent of bin/train.py:186-245: one row per non-zero of y, copying the rows of
    every extra matrix.  Raises like the reference if a row has no non-zero."""
    assert sparse.issparse(y), 'Matrix y should be sparse.'
    num_instances, num_classes = y.shape
    assert num_classes < (1 << 31), \
        'Number of classes should be encodable in 32-bit signed integer.'
    cx = y.tocoo()
    order = np.lexsort((cx.col, cx.row)) if not _coo_sorted(cx) else slice(None)
    rows, cols = cx.row[order], cx.col[order]
    if num_instances and (len(rows) == 0 or len(np.unique(rows)) != num_instances):
        raise RuntimeError('Every truth value should have at least one non-zero index.')
    new_y = cols.astype(np.int32)
    new_matrices = []
    for matrix in matrices:
        assert isinstance(matrix, np.ndarray), 'Matrix {0} should be dense.'.format(repr(matrix))
        assert matrix.shape[0] == num_instances
        new_matrices.append(np.ascontiguousarray(matrix[rows]))
    return new_y, new_matrices


def _coo_sorted(cx):
    if len(cx.row) < 2:
        return True
    return bool(np.all(np.diff(cx.row) >= 0))


def loglinear_corpus(seed, V, E, W, n_train, n_val):
    rng = np.random.default_rng(seed)
    x_train = make_windows(rng, n_train, W, V)
    y_train = make_csr_labels(rng, n_train, E)
    w_train = make_weights(rng, n_train)
    x_val = make_windows(rng, n_val, W, V)
    y_val = make_csr_labels(rng, n_val, E)
    return (x_train, y_train, w_train), (x_val, y_val)


def vectorspace_corpus(seed, V, E, W, n_train, n_val):
    rng = np.random.default_rng(seed)
    x_train = make_windows(rng, n_train, W, V)
    y_train = sample_zipf(rng, E, n_train, s=1.0, shift=1.0).astype(np.int32)
    w_train = np.ones(n_train, dtype=np.float32)
    x_val = make_windows(rng, n_val, W, V)
    y_val = sample_zipf(rng, E, n_val, s=1.0, shift=1.0).astype(np.int32)
    return (x_train, y_train, w_train), (x_val, y_val)


def glorot(rng, shape):
    a = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def write_corpus_files(directory, kind, seed, V, E, W, n_train, n_val, num_topics=12):
    """Writes `data.npz`, `meta` and `topics` in the reference's on-disk formats (bin/prepare.py:373-416):
    meta = 5 sequential pickles (args, words{str->Word(id,count)}, tokens[str], entity_indices_inv{int->str},
    documents_per_entity); data.npz = x_train, y_train (CSR pickled as a 0-d object array), w_train,
    x_validate, y_validate; topics = `topic_id;terms` lines.  Returns the paths."""
    import argparse
    import os
    import pickle
    from cvangysel.io_utils import Word
    rng = np.random.default_rng(seed)
    train, val = loglinear_corpus(seed, V, E, W, n_train, n_val)
    def alpha(i):                       # letters only: parse_query drops digits (io_utils.py:70-88)
        out = ''
        while True:
            out = chr(ord('a') + i % 26) + out
            i //= 26
            if i == 0:
                return 'tok' + out
    tokens = [alpha(i) for i in range(V)]
    words = {tok: Word(id=i, count=int(V - i)) for i, tok in enumerate(tokens)}
    entity_indices_inv = {i: 'entity-%05d' % i for i in range(E)}
    documents_per_entity = {name: ['doc-%d' % i] for i, name in entity_indices_inv.items()}
    data_args = argparse.Namespace(window_size=W, kind=kind)
    os.makedirs(directory, exist_ok=True)
    meta_path = os.path.join(directory, 'meta')
    with open(meta_path, 'wb') as f:
        for obj in (data_args, words, tokens, entity_indices_inv, documents_per_entity):
            pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
    data_path = os.path.join(directory, 'data.npz')
    with open(data_path, 'wb') as f:
        np.savez(f, x_train=train[0], y_train=train[1], w_train=train[2], x_validate=val[0], y_validate=val[1])
    topics_path = os.path.join(directory, 'topics')
    with open(topics_path, 'w') as f:
        for t in range(num_topics):
            n_terms = int(rng.integers(1, 2 * W + 3))
            terms = [tokens[int(i)] for i in sample_zipf(rng, V, n_terms)]
            if t % 5 == 0:
                terms.append('zzzunknownzzz')                     # an out-of-vocabulary term
            f.write('T%03d;%s\n' % (t, ' '.join(terms)))
        f.write('T999;onlyunknownterms here\n')                   # skipped with a warning (bin/query.py:139-142)
    return data_path, meta_path, topics_path

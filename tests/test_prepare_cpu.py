"""sert_b200/prepare.py (the packing step that feeds the hot path) against the fixture made
by running the reference's own bin/prepare.py::instances_and_labels_to_arrays (tests/golden/make_golden.py)."""
import json
import os
import pickle

import numpy as np
import pytest
from scipy import sparse

from sert_b200 import prepare

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def ref():
    with open(os.path.join(HERE, 'golden', 'prepare_ref.json')) as f:
        d = json.load(f)
    d['instances'] = [(doc, tuple(w), dict((e, m) for e, m in label)) for doc, w, label in d['instances']]
    return d


@pytest.mark.parametrize('case', ['ordered', 'shuffle'])
def test_instances_and_labels_to_arrays_matches_reference(ref, case):
    inst = list(ref['instances'])
    np.random.seed(4711)                      # the reference shuffles with the global numpy RNG
    dtype = np.min_scalar_type(ref['num_words'] - 1)
    x, y = prepare.instances_and_labels_to_arrays(inst, ref['window_size'], ref['class_mapping'], dtype,
                                                  case == 'shuffle')
    c = ref['cases'][case]
    assert str(x.dtype) == c['x_dtype'] == 'uint32'
    np.testing.assert_array_equal(x, np.array(c['x'], dtype=dtype))
    assert sparse.isspmatrix_csr(y) and list(y.shape) == c['shape']
    assert str(y.dtype) == c['y_dtype'] and str(y.indices.dtype) == c['indices_dtype']
    np.testing.assert_array_equal(y.indptr, c['indptr'])
    np.testing.assert_array_equal(y.indices, c['indices'])
    np.testing.assert_array_equal(y.data, np.array(c['data'], dtype=np.float32))
    per_doc = {}
    for doc_id, _, _ in inst:
        per_doc[doc_id] = per_doc.get(doc_id, 0) + 1
    w = prepare.instance_weights(inst, per_doc, ref['max_document_length'])
    assert w.dtype == np.float32
    np.testing.assert_array_equal(w, np.array(c['w'], dtype=np.float32))


def test_edge_cases_and_file_round_trip(ref, tmp_path):
    # empty instance list: (0, W) / (0, E) like the reference's fromiter / csr_matrix on empty inputs
    x, y = prepare.instances_and_labels_to_arrays([], 3, {'a': 0, 'b': 1}, np.uint16, shuffle=False)
    assert x.shape == (0, 3) and x.dtype == np.uint16 and y.shape == (0, 2) and y.nnz == 0
    # uint32 above 65 536 words / uint16 up to there; entity indexing drops entities without instances, in order
    per_entity = {'ghost-a': 0}
    for _, _, label in ref['instances'][:60]:
        for e in label:
            per_entity[e] = per_entity.get(e, 0) + 1
    per_entity['ghost-b'] = 0
    packed = prepare.pack(list(ref['instances'][:50]), list(ref['instances'][50:60]), ref['window_size'],
                          ref['num_words'], per_entity, shuffle=False)
    assert packed['x_train'].dtype == np.uint32 and packed['x_train'].shape == (50, ref['window_size'])
    kept = [e for e in per_entity if per_entity[e]]
    assert [packed['entity_indices_inv'][i] for i in range(len(kept))] == kept
    assert packed['y_train'].shape == (50, len(kept)) and packed['y_validate'].shape == (10, len(kept))
    small = prepare.pack([('d', (1, 2), {'ent-01': 1.0})], [('d', (3, 4), {'ent-02': 1.0})], 2, 65536,
                         {'ent-01': 1, 'ent-02': 1}, instances_per_document={'d': 2}, max_document_length=4)
    assert small['x_train'].dtype == np.uint16 and small['w_train'].tolist() == [2.0]
    # data.npz / meta round trip in the formats bin/train.py and bin/query.py read
    c = ref['cases']['ordered']
    y = sparse.csr_matrix((np.array(c['data'], np.float32), np.array(c['indices']), np.array(c['indptr'])),
                          shape=c['shape'])
    xs = np.array(c['x'], dtype=np.uint32)
    w = np.array(c['w'], dtype=np.float32)
    data_path, meta_path = str(tmp_path / 'data.npz'), str(tmp_path / 'meta')
    prepare.write_data(data_path, xs, y, xs[:7], y[:7], w_train=w)
    loaded = np.load(data_path, allow_pickle=True)
    assert list(loaded.keys()) == ['x_train', 'y_train', 'w_train', 'x_validate', 'y_validate']
    np.testing.assert_array_equal(loaded['x_train'], xs)
    assert (loaded['y_train'].item() != y).nnz == 0 and loaded['y_validate'].item().shape == (7, c['shape'][1])
    prepare.write_meta(meta_path, {'window_size': 5}, {'w': (0, 3)}, ['w'], {0: 'ent-00'}, {'ent-00': ['doc000']})
    with open(meta_path, 'rb') as f:
        objs = [pickle.load(f) for _ in range(5)]
    assert objs[3] == {0: 'ent-00'} and prepare.read_meta(meta_path)[2] == ['w']


def test_document_windows_match_the_reference_generator(ref):
    """cvangysel io_utils.windowed_translated_token_stream (io_utils.py:151-211) as strided arrays: 120 seeded token
    streams with OOV tokens, end-of-sentence tokens, strides 1..window and optional padding (fixture made by running
    the reference's generator)."""
    import collections
    Word = collections.namedtuple('Word', ['id', 'count'])
    words = {t: Word(i, 1) for i, t in enumerate(ref['window_vocab'])}
    seen_padded = seen_eos = 0
    for case in ref['window_cases']:
        got = prepare.document_windows(case['tokens'], words, case['window_size'], case['stride'],
                                       case['padding_token'])
        assert got.shape == (len(case['windows']), case['window_size']) and got.dtype == np.int64
        assert got.tolist() == case['windows'], case
        seen_padded += any(words['<pad>'].id in w for w in case['windows'])
        seen_eos += '</s>' in case['tokens']
    assert seen_padded > 10 and seen_eos > 10
    # one run of ids, the building block: 7 tokens, window 3, stride 2 -> [0:3], [2:5], [4:7]; 8 tokens add a padded tail
    assert prepare.windows_from_ids(np.arange(7), 3, 2, padding_id=99).tolist() == [[0, 1, 2], [2, 3, 4], [4, 5, 6]]
    assert prepare.windows_from_ids(np.arange(8), 3, 2, padding_id=99).tolist()[-1] == [6, 7, 99]
    assert prepare.windows_from_ids(np.arange(8), 3, 2).shape == (3, 3)
    assert prepare.windows_from_ids(np.arange(2, dtype=np.uint16), 4, 1, padding_id=9).tolist() == [[0, 1, 9, 9]]


@pytest.mark.parametrize('shuffle', [False, True])
def test_pack_document_windows_equals_the_tuple_path(shuffle):
    """Array-native packing == instances_and_labels_to_arrays (pinned to the reference above) on the instances the same
    documents expand to, including the in-place shuffle driven by the global numpy RNG, and the w_train values."""
    rng = np.random.default_rng(8)
    W, V, n_ent = 4, 500, 9
    entity_ids = ['e%d' % i for i in range(n_ent)]
    class_mapping = {e: i for i, e in enumerate(rng.permutation(entity_ids).tolist())}
    doc_windows, doc_entities, instances, per_doc = [], [], [], {}
    for d in range(30):
        n = int(rng.integers(0, 7))                      # documents without windows occur
        win = rng.integers(0, V, (n, W))
        ents = rng.choice(entity_ids, int(rng.integers(1, 4)), replace=False).tolist()
        doc_windows.append(win)
        doc_entities.append(ents)
        label = {e: 1.0 / len(ents) for e in ents}
        per_doc['d%d' % d] = n
        instances += [('d%d' % d, tuple(row), label) for row in win.tolist()]
    max_len = max(per_doc.values())
    np.random.seed(99)
    inst = list(instances)
    x_ref, y_ref = prepare.instances_and_labels_to_arrays(inst, W, class_mapping, np.uint16, shuffle)
    w_ref = prepare.instance_weights(inst, per_doc, max_len)
    np.random.seed(99)
    x, y, w = prepare.pack_document_windows(doc_windows, doc_entities, class_mapping, np.uint16, shuffle,
                                            max_document_length=max_len)
    assert x.dtype == np.uint16 and y.indices.dtype == np.int32
    np.testing.assert_array_equal(x, x_ref)
    assert y.shape == y_ref.shape and (y != y_ref).nnz == 0
    np.testing.assert_array_equal(y.indptr, y_ref.indptr)
    np.testing.assert_array_equal(y.indices, y_ref.indices)
    np.testing.assert_array_equal(y.data, y_ref.data)
    np.testing.assert_array_equal(w, w_ref)

"""Host logic of the training driver (sert_b200/training.py behind bin/train.py): command line, data.npz loading,
one-hot expansion, pre-trained initialisation and the epoch protocol of bin/train.py:262-348, on a fake model."""
import collections
import os
import pickle

import numpy as np
import pytest
from scipy import sparse

from sert_b200 import models, prepare, training


class FakeModel(models.ModelInterface):
    def __init__(self, train_errors):
        super(FakeModel, self).__init__(batch_size=4)
        self.train_errors = list(train_errors)
        self.calls = []

    def train(self):
        self.calls.append('train')
        return 3, 0.5

    def train_error(self):
        self.calls.append('train_error')
        return self.train_errors.pop(0), 0.1

    def validation_error(self):
        self.calls.append('validation_error')
        return 1.0, 0.2

    def get_state(self):
        return ['predict_fn', np.arange(3)]


def test_command_line_matches_the_reference_flags(tmp_path):
    data = tmp_path / 'data.npz'
    data.write_bytes(b'x')
    args = training.parse_args(['--data', str(data), '--meta', str(data), '--type', 'vectorspace',
                                '--model_output', 'out', '--batch_size', '4096', '--num_negative_samples', '10',
                                '--word_representation_size', '128', '--one_hot_classes'])
    assert args.type is models.VectorSpaceLanguageModel and args.entity_representation_size == 128
    assert args.regularization_lambda == 0.01 and args.iterations == 1 and args.one_hot_classes
    with pytest.raises(SystemExit):
        training.parse_args(['--data', str(data), '--meta', str(data), '--type', 'cnn', '--model_output', 'o'])
    with pytest.raises(SystemExit):
        training.parse_args(['--data', str(tmp_path / 'missing'), '--meta', str(data), '--type', 'loglinear',
                             '--model_output', 'o'])


def test_data_loading_weights_and_one_hot(tmp_path):
    y = sparse.csr_matrix(np.array([[0.5, 0.5, 0], [0, 0, 1.0], [0, 1.0, 0]], dtype=np.float32))
    x = np.arange(6, dtype=np.uint16).reshape(3, 2)
    w = np.array([1.0, 2.0, 3.0], dtype=np.float32)
    path = str(tmp_path / 'data.npz')
    prepare.write_data(path, x, y, x[:1], y[:1], w_train=w)
    (xt, yt, wt), (xv, yv) = training.load_data_sets(path)
    np.testing.assert_array_equal(wt, w)
    assert sparse.issparse(yt) and yv.shape == (1, 3)
    (_, _, wt2), _ = training.load_data_sets(path, ignore_weights=True)
    np.testing.assert_array_equal(wt2, np.ones(3, np.float32))
    (x1, y1, w1), (xv1, yv1) = training.to_one_hot((xt, yt, wt), (xv, yv))
    assert x1.shape == (4, 2) and sorted(y1.tolist()) == [0, 1, 1, 2] and w1.shape == (4,)
    assert yv1.tolist() == [0, 1] and xv1.shape == (2, 2)
    with pytest.raises(RuntimeError, match='expects sparse truth values'):
        training.to_one_hot((xt, yt.toarray(), wt), (xv, yv))


def test_pretrained_rows_overwrite_the_glorot_table(tmp_path):
    Word = collections.namedtuple('Word', ['id', 'count'])
    words = {'Alpha': Word(0, 5), 'beta': Word(1, 3), 'gamma': Word(2, 1)}
    path = str(tmp_path / 'w2v.bin')
    vectors = {'alpha': np.full(4, 0.25, np.float32), 'beta': np.full(4, -0.5, np.float32)}
    with open(path, 'wb') as f:                          # word2vec binary: "<n> <d>\n" then "<word> " + d float32 + "\n"
        f.write(b'2 4\n')
        for word, vec in vectors.items():
            f.write(word.encode() + b' ' + vec.tobytes() + b'\n')
    np.random.seed(0)
    table = training.word_representations(4, words, ['alpha', 'beta', 'gamma'], path)
    assert table.shape == (3, 4) and table.dtype == np.float32
    np.testing.assert_array_equal(table[0], vectors['alpha'])      # looked up lower-cased
    np.testing.assert_array_equal(table[1], vectors['beta'])
    limit = np.sqrt(6.0 / (3 + 4))
    assert np.all(np.abs(table[2]) <= limit) and np.any(table[2] != 0)


def test_epoch_protocol_dumps_every_epoch_and_stops_when_learning_stalls(tmp_path):
    out = str(tmp_path / 'model')
    model = FakeModel([2.0, 1.5, 1.2, 1.2 + 1e-7, 0.3])
    training.train(model, 10, out, abort_threshold=1e-5, additional_args=[{'window_size': 5}])
    # errors(0), dump(0), then train -> errors -> dump per epoch; the 3rd epoch moves the error by 1e-7: stop
    assert model.calls == ['train_error', 'validation_error'] + ['train', 'train_error', 'validation_error'] * 3
    assert sorted(os.listdir(str(tmp_path))) == ['model_0.bin', 'model_1.bin', 'model_2.bin', 'model_3.bin']
    with open(out + '_2.bin', 'rb') as f:
        objs = [pickle.load(f) for _ in range(3)]
    assert objs[0] == {'window_size': 5} and objs[1] == 'predict_fn' and objs[2].tolist() == [0, 1, 2]
    assert training.EpochLoop.delta([2.0, 1.0]) == (-1.0, -0.5) and training.EpochLoop.delta([2.0]) == (0.0, 0.0)
    with pytest.raises(AssertionError):
        training.train(FakeModel([float('nan'), float('nan')]), 1, str(tmp_path / 'bad'))


def test_hot_word_selection():
    rng = np.random.default_rng(3)
    x = rng.integers(10, 1000, (4000, 10))
    x[:, 0] = 7                      # once in every window: 100 per batch of 100
    x[::2, 1] = 3                    # every other window: 50 per batch -> below the threshold of 64
    x[:, 2] = np.where(np.arange(4000) % 4 < 3, 5, x[:, 2])     # 75 per batch
    ids = models.hot_word_ids(x, 1000, 100)
    assert ids.dtype == np.int32 and ids.tolist() == [7, 5]
    assert models.hot_word_ids(x, 1000, 200).tolist()[:3] == [7, 5, 3]          # 200 / 150 / 100 per batch
    assert models.hot_word_ids(np.zeros((0, 10), np.int64), 1000, 100).size == 0
    dense = np.tile(np.arange(40), (50, 1))                                     # 40 words, each once per window
    assert models.hot_word_ids(dense, 40, 128).tolist() == list(range(32))      # capped at 32, lower ids first


def test_synthetic_corpus_files_feed_the_training_driver(tmp_path):
    """The files the CLI tests train from (synth.write_corpus_files) are in the reference's on-disk formats: the driver's
    loader, the meta reader and the topic parser accept them (the GPU-side continuation is tests/test_gpu_cli.py)."""
    from cvangysel import trec_utils
    from scipy import sparse as sp
    from sert_b200 import synth
    data_path, meta_path, topics_path = synth.write_corpus_files(str(tmp_path), 'loglinear', 5, V=300, E=40, W=4,
                                                                 n_train=96, n_val=32)
    (x, y, w), (xv, yv) = training.load_data_sets(data_path)
    assert x.shape == (96, 4) and x.dtype == np.uint16 and sp.issparse(y) and y.shape == (96, 40)
    assert w.shape == (96,) and w.dtype == np.float32 and xv.shape == (32, 4) and yv.shape == (32, 40)
    np.testing.assert_allclose(np.asarray(y.sum(axis=1)).ravel(), 1.0, rtol=1e-6)       # label masses sum to one
    data_args, words, tokens, entity_indices_inv, documents_per_entity = prepare.read_meta(meta_path)
    assert data_args.window_size == 4 and len(words) == len(tokens) == 300 and len(entity_indices_inv) == 40
    assert int(x.max()) < len(words) and words[tokens[17]].id == 17
    with open(topics_path) as f:
        topics = trec_utils.parse_topics(f)
    assert 'T999' in topics and len(topics) == 13
    known = [t for t in trec_utils.parse_query(topics['T000']) if t in words]
    assert known and 'zzzunknownzzz' not in words
    (x1, y1, w1), _ = training.to_one_hot((x, y, w), (xv, yv))
    assert x1.shape[0] == y.nnz and y1.dtype == np.int32 and int(y1.max()) < 40

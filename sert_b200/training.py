"""Training driver behind bin/train.py: the reference's command line (bin/train.py:28-63), data / meta loading
(:79-151) and epoch protocol (:262-348) on the B200-native models.

Fixes of crashes at the reference's pinned commit: `--ignore_weights` exists (the reference
reads args.ignore_weights without defining the flag, bin/train.py:81) and data.npz is loaded with allow_pickle (it
holds a pickled CSR matrix).
"""
import argparse
import logging
import os
import pickle

import numpy as np
import scipy
import scipy.sparse

from cvangysel import argparse_utils, embedding_utils, logging_utils
from sert_b200 import models
from sert_b200.synth import sparse_to_one_hot_multiple

MODEL_TYPES = {'loglinear': models.LanguageModel, 'vectorspace': models.VectorSpaceLanguageModel}

_FILE = argparse_utils.existing_file_path
_COUNT = argparse_utils.positive_int
# (flag, keyword arguments of add_argument), in the reference's order
CLI_FLAGS = (
    ('--loglevel', dict(type=str, default='INFO')),
    ('--data', dict(type=_FILE, required=True)),
    ('--meta', dict(type=_FILE, required=True)),
    ('--type', dict(choices=MODEL_TYPES, required=True)),
    ('--iterations', dict(type=_COUNT, default=1)),
    ('--batch_size', dict(type=_COUNT, default=1024)),
    ('--word_representation_size', dict(type=_COUNT, default=300)),
    ('--representation_initializer', dict(type=_FILE, default=None)),
    ('--entity_representation_size', dict(type=_COUNT, default=None)),        # vector space only
    ('--num_negative_samples', dict(type=_COUNT, default=None)),              # vector space only
    ('--one_hot_classes', dict(action='store_true', default=False)),
    ('--regularization_lambda', dict(type=argparse_utils.ratio, default=0.01)),
    ('--ignore_weights', dict(action='store_true', default=False)),
    ('--model_output', dict(type=str, required=True)),
)


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    for flag, options in CLI_FLAGS:
        parser.add_argument(flag, **options)
    args = parser.parse_args(argv)
    if args.entity_representation_size is None:
        args.entity_representation_size = args.word_representation_size
    args.type = MODEL_TYPES[args.type]           # the pickled Namespace carries the class (bin/query.py dispatches on it)
    return args


def _describe(arrays):
    return ' '.join('%s (%s)' % (a.shape, a.dtype) for a in arrays)


def load_data_sets(path, ignore_weights=False):
    """data.npz -> ((x_train, y_train, w_train), (x_validate, y_validate)); bin/train.py:79-98."""
    logging.info('Loading data from %s.', path)
    archive = np.load(path, allow_pickle=True)
    x_train, y_train = archive['x_train'], archive['y_train'][()]
    if ignore_weights or 'w_train' not in archive:
        logging.warning('No weights found in data set; assuming uniform instance weighting.')
        w_train = np.ones(x_train.shape[0], dtype=np.float32)
    else:
        w_train = archive['w_train']
    training = (x_train, y_train, w_train)
    validation = (archive['x_validate'], archive['y_validate'][()])
    logging.info('Training instances: %s', _describe(training))
    logging.info('Validation instances: %s', _describe(validation))
    return training, validation


def to_one_hot(training, validation):
    """--one_hot_classes: one instance per (instance, entity) pair (bin/train.py:106-120, 186-245)."""
    logging.info('Transforming y-values to one-hot values.')
    if not (scipy.sparse.issparse(training[1]) and scipy.sparse.issparse(validation[1])):
        raise RuntimeError('Argument --one_hot_classes expects sparse truth values.')
    y_train, (x_train, w_train) = sparse_to_one_hot_multiple(training[1], training[0], training[2])
    y_validate, (x_validate,) = sparse_to_one_hot_multiple(validation[1], validation[0])
    return (x_train, y_train, w_train), (x_validate, y_validate)


def word_representations(size, words, tokens, initializer_path=None):
    """Glorot-uniform table, rows overwritten from a word2vec binary where the (lower-cased) word occurs in it
    (bin/train.py:128-151)."""
    table = models.glorot_uniform((len(words), size))
    if not initializer_path:
        return table
    pretrained = dict(embedding_utils.load_binary_representations(initializer_path, tokens))   # duplicates: last wins
    found = [(meta.id, pretrained[word.lower()]) for word, meta in words.items() if word.lower() in pretrained]
    for row, vector in found:
        table[row] = vector
    logging.info('Initialized representations from pre-learned collection for %d words (%.2f%%).',
                 len(found), 100.0 * len(found) / float(len(words)))
    return table


def build_model(args, num_entities, training, validation, window_size, representations):
    common = dict(batch_size=args.batch_size, window_size=window_size, representations_init=representations,
                  regularization_lambda=args.regularization_lambda, training_set=training, validation_set=validation)
    if args.type is models.LanguageModel:
        return models.LanguageModel(output_layer_size=num_entities, **common)
    return models.VectorSpaceLanguageModel(
        entity_representations_init=models.glorot_uniform((num_entities, args.entity_representation_size)),
        num_negative_samples=args.num_negative_samples, **common)


class EpochLoop(object):
    """bin/train.py:262-348: errors(0) and dump(0), then per epoch train -> errors -> dump; stops when the mean
    training error moves by less than `abort_threshold` (or, with early_stopping, when validation gets worse)."""

    def __init__(self, model, output_path, pickled_prefix=(), abort_threshold=1e-5, early_stopping=False):
        assert isinstance(model, models.ModelInterface) and isinstance(abort_threshold, float)
        self.model, self.output_path, self.prefix = model, output_path, list(pickled_prefix)
        self.abort_threshold, self.early_stopping = abort_threshold, early_stopping
        self.errors = {'Training': [], 'Validation': []}          # lists of (mean, std)

    def measure(self):
        self.errors['Training'].append(tuple(self.model.train_error()))
        self.errors['Validation'].append(tuple(self.model.validation_error()))

    def dump(self, epoch):
        path = '{0}_{1}.bin'.format(self.output_path, epoch)
        with open(path, 'wb') as f:
            for obj in self.prefix + list(self.model.get_state()):
                pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
        logging.info('Saved model "%s" (%d megabyte).', path, os.path.getsize(path) / 1024 / 1024)

    @staticmethod
    def delta(means):
        if len(means) < 2:
            return 0.0, 0.0
        step = means[-1] - means[-2]
        return step, step / float(means[-2])

    def should_stop(self):
        train = [m for m, _ in self.errors['Training']]
        valid = [m for m, _ in self.errors['Validation']]
        assert np.all(np.isfinite(train[-1]))
        if self.early_stopping:
            assert np.all(np.isfinite(valid[-1]))
            if valid[-1] > valid[-2]:
                logging.info('Validation error stopped decreasing; aborting.')
                return True
        if len(train) > 1 and abs(train[-1] - train[-2]) < self.abort_threshold:
            logging.error('No learning was performed during the last iteration; aborting.')
            return True
        return False

    def run(self, num_epochs):
        self.measure()
        self.dump(0)
        for epoch in range(1, num_epochs + 1):
            logging.info('Epoch %d.', epoch)
            num_batches, mean_cost = self.model.train()
            logging.info('Epoch %d: processed %d batches; average error=%f.', epoch, num_batches, mean_cost)
            logging.info('Epoch %d: measuring training/validation error.', epoch)
            self.measure()
            for label, history in self.errors.items():
                logging.info('%s errors: %s; delta=%s', label, history, self.delta([m for m, _ in history]))
            self.dump(epoch)
            if self.should_stop():
                break


def train(model, num_epochs, output_path, abort_threshold=1e-5, early_stopping=False, additional_args=()):
    """Signature of the reference's bin/train.py::train."""
    EpochLoop(model, output_path, additional_args, abort_threshold, early_stopping).run(num_epochs)


def main(argv=None):
    args = parse_args(argv)
    try:
        logging_utils.configure_logging(args)
    except IOError:
        return -1
    logging_utils.log_module_info(np, scipy)

    training, validation = load_data_sets(args.data, args.ignore_weights)
    num_entities = training[1].shape[1]
    assert num_entities > 1
    if args.one_hot_classes:
        training, validation = to_one_hot(training, validation)

    logging.info('Loading meta-data from %s.', args.meta)
    with open(args.meta, 'rb') as f:
        data_args, words, tokens = (pickle.load(f) for _ in range(3))     # the rest of meta is for bin/query.py
    representations = word_representations(args.word_representation_size, words, tokens,
                                           args.representation_initializer)
    del words, tokens

    model = build_model(args, num_entities, training, validation, data_args.window_size, representations)
    train(model, args.iterations, args.model_output, abort_threshold=1e-5, early_stopping=False,
          additional_args=[args])
    return 0

#!/usr/bin/env python
"""Training driver with the reference's command line (bin/train.py:28-63) on the B200-native models.

Differences from the reference, all fixes of crashes at the pinned commit:
`--ignore_weights` exists (the reference reads args.ignore_weights without defining the flag,
bin/train.py:81) and data.npz is loaded with allow_pickle (it holds a pickled CSR matrix).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from cvangysel import argparse_utils, embedding_utils, logging_utils  # noqa: E402
from sert import models  # noqa: E402
from sert_b200.synth import sparse_to_one_hot_multiple  # noqa: E402,F401

import argparse  # noqa: E402
import logging  # noqa: E402
import numpy as np  # noqa: E402
import pickle  # noqa: E402
import scipy  # noqa: E402
import scipy.sparse  # noqa: E402

MODELS = {
    'loglinear': models.LanguageModel,
    'vectorspace': models.VectorSpaceLanguageModel,
}


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--loglevel', type=str, default='INFO')
    parser.add_argument('--data', type=argparse_utils.existing_file_path, required=True)
    parser.add_argument('--meta', type=argparse_utils.existing_file_path, required=True)
    parser.add_argument('--type', choices=MODELS, required=True)
    parser.add_argument('--iterations', type=argparse_utils.positive_int, default=1)
    parser.add_argument('--batch_size', type=argparse_utils.positive_int, default=1024)
    parser.add_argument('--word_representation_size', type=argparse_utils.positive_int, default=300)
    parser.add_argument('--representation_initializer', type=argparse_utils.existing_file_path, default=None)
    # Specific to VectorSpaceLanguageModel.
    parser.add_argument('--entity_representation_size', type=argparse_utils.positive_int, default=None)
    parser.add_argument('--num_negative_samples', type=argparse_utils.positive_int, default=None)
    parser.add_argument('--one_hot_classes', action='store_true', default=False)
    parser.add_argument('--regularization_lambda', type=argparse_utils.ratio, default=0.01)
    parser.add_argument('--ignore_weights', action='store_true', default=False)
    parser.add_argument('--model_output', type=str, required=True)
    return parser


def load_data(args):
    logging.info('Loading data from %s.', args.data)
    data_sets = np.load(args.data, allow_pickle=True)

    if 'w_train' in data_sets and not args.ignore_weights:
        w_train = data_sets['w_train']
    else:
        logging.warning('No weights found in data set; assuming uniform instance weighting.')
        w_train = np.ones(data_sets['x_train'].shape[0], dtype=np.float32)

    training_set = (data_sets['x_train'], data_sets['y_train'][()], w_train)
    validation_set = (data_sets['x_validate'], data_sets['y_validate'][()])

    logging.info('Training instances: %s (%s) %s (%s) %s (%s)',
                 training_set[0].shape, training_set[0].dtype, training_set[1].shape, training_set[1].dtype,
                 training_set[2].shape, training_set[2].dtype)
    logging.info('Validation instances: %s (%s) %s (%s)',
                 validation_set[0].shape, validation_set[0].dtype, validation_set[1].shape, validation_set[1].dtype)
    return training_set, validation_set


def initial_word_representations(args, words, tokens):
    vocabulary_size = len(words)
    representations = models.glorot_uniform((vocabulary_size, args.word_representation_size))

    if args.representation_initializer:
        # Duplicate words in the initializer are ignored by construction of the dictionary.
        lookup = dict(embedding_utils.load_binary_representations(args.representation_initializer, tokens))
        hits = 0
        for word, meta in words.items():
            vector = lookup.get(word.lower())
            if vector is not None:
                representations[meta.id] = vector
                hits += 1
        logging.info('Initialized representations from pre-learned collection for %d words (%.2f%%).',
                     hits, (hits / float(len(words))) * 100.0)
    return representations


def main(argv=None):
    args = build_parser().parse_args(argv)

    if args.entity_representation_size is None:
        args.entity_representation_size = args.word_representation_size

    args.type = MODELS[args.type]

    try:
        logging_utils.configure_logging(args)
    except IOError:
        return -1

    logging_utils.log_module_info(np, scipy)

    training_set, validation_set = load_data(args)

    num_entities = training_set[1].shape[1]
    assert num_entities > 1

    if args.one_hot_classes:
        logging.info('Transforming y-values to one-hot values.')

        if not scipy.sparse.issparse(training_set[1]) or not scipy.sparse.issparse(validation_set[1]):
            raise RuntimeError('Argument --one_hot_classes expects sparse truth values.')

        y_train, (x_train, w_train) = sparse_to_one_hot_multiple(
            training_set[1], training_set[0], training_set[2])
        training_set = (x_train, y_train, w_train)

        y_validate, (x_validate,) = sparse_to_one_hot_multiple(validation_set[1], validation_set[0])
        validation_set = (x_validate, y_validate)

    logging.info('Loading meta-data from %s.', args.meta)
    with open(args.meta, 'rb') as f:
        # The remainder of the meta file is not needed for training.
        data_args, words, tokens = (pickle.load(f) for _ in range(3))

    representations = initial_word_representations(args, words, tokens)
    del words
    del tokens

    model_options = {
        'batch_size': args.batch_size,
        'window_size': data_args.window_size,
        'representations_init': representations,
        'regularization_lambda': args.regularization_lambda,
        'training_set': training_set,
        'validation_set': validation_set,
    }

    if args.type == models.LanguageModel:
        model_options.update(output_layer_size=num_entities)
    elif args.type == models.VectorSpaceLanguageModel:
        model_options.update(
            entity_representations_init=models.glorot_uniform((num_entities, args.entity_representation_size)),
            num_negative_samples=args.num_negative_samples)

    model = args.type(**model_options)

    train(model, args.iterations, args.model_output,
          abort_threshold=1e-5, early_stopping=False, additional_args=[args])


def error_delta(error):
    if len(error) <= 1:
        return 0.0, 0.0
    absolute = error[-1] - error[-2]
    return absolute, absolute / float(error[-2])


def train(model, num_epochs, output_path, abort_threshold=1e-5, early_stopping=False, additional_args=[]):
    """Epoch protocol of bin/train.py:262-348: errors(0), dump(0), then per epoch train -> errors -> dump,
    stopping when the mean training error moves by less than abort_threshold."""
    assert isinstance(model, models.ModelInterface)
    assert isinstance(abort_threshold, float)

    history = {'training': ([], []), 'validation': ([], [])}

    def compute_errors():
        for name, (mean, std) in (('training', model.train_error()), ('validation', model.validation_error())):
            history[name][0].append(mean)
            history[name][1].append(std)

    def dump_model(epoch):
        filename = '{0}_{1}.bin'.format(output_path, epoch)
        with open(filename, 'wb') as f:
            for obj in additional_args + list(model.get_state()):
                pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
        logging.info('Saved model "%s" (%d megabyte).', filename, os.path.getsize(filename) / 1024 / 1024)

    compute_errors()
    dump_model(0)

    for epoch in range(1, num_epochs + 1):
        logging.info('Epoch %d.', epoch)

        num_batches, mean_cost = model.train()
        logging.info('Epoch %d: processed %d batches; average error=%f.', epoch, num_batches, mean_cost)

        logging.info('Epoch %d: measuring training/validation error.', epoch)
        compute_errors()

        for name, label in (('training', 'Training'), ('validation', 'Validation')):
            means, stds = history[name]
            logging.info('%s errors: %s; delta=%s', label, list(zip(means, stds)), error_delta(means))

        dump_model(epoch=epoch)

        train_means, validation_means = history['training'][0], history['validation'][0]
        assert np.all(np.isfinite(train_means[-1]))

        if early_stopping:
            assert np.all(np.isfinite(validation_means[-1]))
            if validation_means[-1] > validation_means[-2]:
                logging.info('Validation error stopped decreasing; aborting.')
                return

        if len(train_means) > 1 and abs(train_means[-1] - train_means[-2]) < abort_threshold:
            logging.error('No learning was performed during the last iteration; aborting.')
            return


if __name__ == "__main__":
    sys.exit(main())

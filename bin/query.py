#!/usr/bin/env python
"""Query driver with the reference's command line (bin/query.py:25-158) on the B200-native ranker."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from cvangysel import argparse_utils, logging_utils, trec_utils  # noqa: E402
from sert import inference, models  # noqa: E402
from sert_b200.ranking import (  # noqa: E402,F401
    Callback, LogLinearCallback, RunCollector, VectorSpaceCallback, compute_normalised_entropy)

import argparse  # noqa: E402
import io  # noqa: E402
import logging  # noqa: E402
import pickle  # noqa: E402


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--loglevel', type=str, default='INFO')
    parser.add_argument('--meta', type=argparse_utils.existing_file_path, required=True)
    parser.add_argument('--model', type=argparse_utils.existing_file_path, required=True)
    parser.add_argument('--topics', type=argparse_utils.existing_file_path, nargs='+')
    parser.add_argument('--top', type=argparse_utils.positive_int, default=None)
    parser.add_argument('--run_out', type=argparse_utils.nonexisting_file_path, required=True)
    return parser


def load_model(path):
    """[train args, predict_fn, word representations(, entity representations)] (bin/train.py:293-295)."""
    with open(path, 'rb') as f:
        model_args, predict_fn = (pickle.load(f) for _ in range(2))
        word_representations = pickle.load(f)
        try:
            entity_representations = pickle.load(f)
        except EOFError:
            entity_representations = None
    return model_args, predict_fn, word_representations, entity_representations


def main(argv=None):
    args = build_parser().parse_args(argv)

    try:
        logging_utils.configure_logging(args)
    except IOError:
        return -1

    model_args, predict_fn, word_representations, entity_representations = load_model(args.model)

    with open(args.meta, 'rb') as f:
        data_args, words, tokens, entity_indices_inv, entity_assocs = (pickle.load(f) for _ in range(5))

    topic_files = [open(filename, 'r') for filename in args.topics]
    topics = trec_utils.parse_topics(topic_files)
    for topic_file in topic_files:
        topic_file.close()

    model_name = os.path.basename(args.model)

    # entity profiling / entity finding rankings, kept as arrays (sert_b200.ranking.RunCollector) instead of the
    # reference's two dictionaries of (relevance, id) tuples (bin/query.py:80-92)
    ranker_callback = RunCollector(entity_indices_inv)

    with open('{0}_debug'.format(args.run_out), 'w') as f_debug_out:
        if model_args.type == models.LanguageModel:
            result_callback = LogLinearCallback(args, model_args, tokens, f_debug_out, ranker_callback)
        elif model_args.type == models.VectorSpaceLanguageModel:
            result_callback = VectorSpaceCallback(entity_representations, args, model_args, tokens,
                                                  f_debug_out, ranker_callback)
        else:
            raise RuntimeError('Unknown model type %s.' % model_args.type)

        batcher = inference.create(predict_fn, word_representations, model_args.batch_size,
                                   data_args.window_size, len(words), result_callback)

        logging.info('Batching queries using %s.', batcher)

        for q_id, (topic_id, terms) in enumerate(topics.items()):
            # Numeric tokens in queries are not replaced.
            query_terms = trec_utils.parse_query(terms)

            logging.debug('Query (%d/%d) %s: %s (%s)', q_id + 1, len(topics), topic_id, query_terms, terms)

            query_tokens = []
            for term in query_terms:
                if term not in words:
                    logging.debug('Term "%s" is OOV.', term)
                    continue
                query_tokens.append(words[term].id)

            if not query_tokens:
                logging.warning('Skipping query with terms "%s".', terms)
                continue

            batcher.submit(query_tokens, topic_id=topic_id)

        batcher.process()

    with io.open('{0}_ep'.format(args.run_out), 'w', encoding='utf8') as out_ep_run, \
            io.open('{0}_ef'.format(args.run_out), 'w', encoding='utf8') as out_ef_run:
        ranker_callback.write(model_name, out_ep_run, out_ef_run)

    logging.info('Saved run to %s.', args.run_out)


if __name__ == "__main__":
    sys.exit(main())

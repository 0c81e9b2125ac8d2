"""Query-side batching: drop-in for the reference's ``sert/inference.py`` (same names and contracts).

* ``WordBatcher`` (log-linear): packs query tokens into fixed ``(batch_size, window_size)`` index + mask
  arrays, a query of T tokens occupying ceil(T/W) left-aligned rows; flushes through ``predict_fn`` when
  the next query does not fit; hands each query the first T per-term rows of its slice
  (sert/inference.py:28-143).
* ``EmbeddingMapper`` (vector space): averages the word representations of a query and projects the
  average with ``predict_fn`` (sert/inference.py:146-167).  Here the projections of all submitted queries
  are computed in ONE device call when ``process()`` runs (the reference projects and ranks one query per
  submit); callbacks still fire in submission order, so the observable result is the same.
* ``aggregate_distribution`` (sert/inference.py:170-183).
"""
import logging

import numpy as np


def create(predict_fn, word_representations,
           batch_size, window_size, vocabulary_size,
           result_callback):
    assert result_callback is not None

    instance_dtype = np.min_scalar_type(vocabulary_size - 1)
    logging.info('Instance elements will be stored using %s.', instance_dtype)

    if result_callback.should_average_input():
        return EmbeddingMapper(predict_fn, word_representations, result_callback)

    return WordBatcher(predict_fn, batch_size, window_size, instance_dtype, result_callback)


class WordBatcher(object):

    OVERFLOW, TRUNCATE = range(5, 7)

    def __init__(self, predict_fn, batch_size, window_size, instance_dtype,
                 result_callback=None, overflow_mode=OVERFLOW):
        assert overflow_mode in (WordBatcher.OVERFLOW, WordBatcher.TRUNCATE)
        if result_callback is not None:
            assert hasattr(result_callback, '__call__')

        self.predict_fn = predict_fn
        self.batch_size, self.window_size = batch_size, window_size
        self.overflow_mode = overflow_mode
        self.callback = result_callback

        logging.debug('Using overflow mode "%s" for queries.',
                      'truncate' if overflow_mode == WordBatcher.TRUNCATE else 'overflow')

        self.batch = np.zeros((batch_size, window_size), dtype=instance_dtype)
        self.mask = np.zeros((batch_size, window_size), dtype=np.int8)
        self._empty_batch()

    def _empty_batch(self):
        self.batch.fill(0)
        self.mask.fill(0)
        self.num_used_instances = 0
        self.requests = []

    def _rows_for(self, num_tokens):
        return -(-num_tokens // self.window_size)

    def submit(self, query_tokens, **kwargs):
        assert len(query_tokens) > 0

        if len(query_tokens) > self.window_size and self.overflow_mode == WordBatcher.TRUNCATE:
            logging.error('Truncated query "%s" as it exceeded the window size.', query_tokens)
            query_tokens = query_tokens[:self.window_size]

        num_instances = self._rows_for(len(query_tokens))
        logging.debug('Payload %s requires %d instances (batch size=%d, current batch=%d).',
                      query_tokens, num_instances, self.batch_size, self.num_used_instances)

        if num_instances > self.batch_size:
            raise RuntimeError()
        if num_instances > self.batch_size - self.num_used_instances:
            self.process()

        self.requests.append((num_instances, query_tokens, kwargs))

        # lay the tokens out row-major over the rows of this request; the last row may be partial
        tokens = np.asarray(query_tokens, dtype=self.batch.dtype)
        first = self.num_used_instances
        flat_batch = self.batch[first:first + num_instances].reshape(-1)
        flat_mask = self.mask[first:first + num_instances].reshape(-1)
        flat_batch[:tokens.size] = tokens
        flat_mask[:tokens.size] = 1
        self.num_used_instances += num_instances

    def process(self):
        if not self.requests:
            return

        logging.debug('Processing batch (batch size=%d, current batch=%d).',
                      self.batch_size, self.num_used_instances)

        if hasattr(self.predict_fn, 'rank') and hasattr(self.callback, 'process_ranked'):
            # device path: per-term softmax, product of experts, renormalisation and the full ranking stay on the
            # device; the callback receives (entity ids, relevances) instead of a (T, E) distribution
            spans, row = [], 0
            for num_instances, payload, kwargs in self.requests:
                spans.append((row, len(payload)))
                row += num_instances
            ranked = self.predict_fn.rank(self.batch[:self.num_used_instances], spans)
            for (num_instances, payload, kwargs), result in zip(self.requests, ranked):
                self.callback.process_ranked(payload, *result, **kwargs)
            self._empty_batch()
            return

        results = self.predict_fn(self.batch, self.mask)
        logging.debug('Retrieved batch results %s.', results.shape)

        row = 0
        for num_instances, payload, kwargs in self.requests:
            per_term = results[row:row + num_instances].reshape((-1, results.shape[-1]))[:len(payload)]
            assert per_term.ndim == 2 and per_term.shape[0] == len(payload)
            self.callback(payload, per_term, **kwargs)
            row += num_instances

        self._empty_batch()


class EmbeddingMapper(object):

    def __init__(self, predict_fn, word_representations, result_callback):
        if result_callback is not None:
            assert hasattr(result_callback, '__call__')
        self.predict_fn = predict_fn
        self.word_representations = word_representations
        self.callback = result_callback
        self._pending = []

    def submit(self, query_tokens, **kwargs):
        avg_word_embedding = self.word_representations[query_tokens, :].mean(axis=0)
        self._pending.append((query_tokens, avg_word_embedding, kwargs))

    def process(self):
        if not self._pending:
            return
        pending, self._pending = self._pending, []
        averages = np.stack([avg for _, avg, _ in pending]).astype(np.float32, copy=False)
        if hasattr(self.predict_fn, 'project'):
            projections = self.predict_fn.project(averages)          # one device call for all queries
        else:
            projections = np.stack([self.predict_fn(avg) for avg in averages])
        if hasattr(self.callback, 'process_many'):
            self.callback.process_many([tokens for tokens, _, _ in pending], projections,
                                       [kw for _, _, kw in pending])
        else:
            for (tokens, _, kwargs), projection in zip(pending, projections):
                self.callback(tokens, projection, **kwargs)


def aggregate_distribution(distribution, mode, axis):
    if mode == 'sum':
        return np.mean(distribution, axis=axis)
    if mode == 'product':
        # exact zeros are masked out of the log and contribute a factor of 1 (sert/inference.py:173-174)
        return np.exp(np.sum(np.ma.log(distribution).filled(0), axis=axis))
    if mode == 'last':
        return np.take(distribution, axis=axis, indices=distribution.shape[axis] - 1)
    if mode == 'max':
        return np.max(distribution, axis=axis)
    if mode == 'identity':
        return distribution
    raise NotImplementedError()

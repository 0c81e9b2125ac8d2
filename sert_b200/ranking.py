"""Ranking callbacks: drop-in for the callbacks of the reference's ``bin/query.py:165-382``.

``LogLinearCallback.process`` is the reference's host arithmetic over per-term distributions (product of experts,
renormalise, rank ALL entities; bin/query.py:204-233), kept for callers that hand it a (T, E) array;
``process_ranked`` takes the same ranking computed on the device (``sert_ll_rank_queries``: nothing of size
(rows, W, E) crosses PCIe), which is what ``WordBatcher`` uses when ``predict_fn`` offers ``rank``.

``VectorSpaceCallback`` keeps the reference's observable behaviour (bin/query.py:241-367) but replaces
the sklearn k-NN / scipy cdist search by the device scorer (``sert_b200.scoring``): the device returns the
top-(k + margin) entities by float32 inner product of the L2-normalised vectors (== Euclidean k-NN
order), the host then
  1. re-selects the k nearest of those by exact float64 Euclidean distance (what sklearn's tree search
     computes), in ascending-distance order,
  2. recomputes every candidate's relevance exactly like the reference, ``(sum(e * q) + 1) / 2`` in
     float32 with numpy's summation order (bin/query.py:351-357), and
  3. orders them with Python's stable descending sort (bin/query.py:361-365).
``process_many`` does this for all submitted queries with ONE scorer call.
"""
import collections
import logging
import operator
import sys

import numpy as np

from cvangysel import trec_utils
from sert_b200 import inference, math_utils

# Extra device candidates fetched with the first request and re-ranked on the host.  NOT a correctness bound: query()
# verifies every row's candidate set against a float32 rounding bound and asks again with more candidates (finally all
# entities) until the set provably contains the k nearest entities by float64 distance.
CANDIDATE_MARGIN = 28


class Callback(object):

    def __init__(self, args, model_args, tokens, f_debug_out, rank_callback):
        self.args = args
        self.model_args = model_args

        self.tokens = tokens

        self.f_debug_out = f_debug_out

        self.rank_callback = rank_callback

        self.topic_projections = {}

    def _register(self, result, topic_id):
        assert topic_id not in self.topic_projections
        self.topic_projections[topic_id] = result.ravel()

    def __call__(self, payload, result, topic_id):
        self._register(result, topic_id)

        logging.debug('Result of shape %s for topic "%s".', result.shape, topic_id)

        self.process(payload, result, topic_id)

    def process(self, payload, distribution, topic_id):
        raise NotImplementedError()

    def should_average_input(self):
        raise NotImplementedError()


class LogLinearCallback(Callback):

    def process(self, payload, distribution, topic_id):
        terms = [self.tokens[token_id] for token_id in payload]
        term_entropies = compute_normalised_entropy(distribution, base=2)

        distribution = inference.aggregate_distribution(distribution, mode='product', axis=0)
        assert distribution.ndim == 1

        distribution /= distribution.sum()

        if not np.isclose(distribution.sum(), 1.0):
            logging.error('Encountered non-normalized distribution for topic "%s" (mass=%.10f).',
                          topic_id, distribution.sum())

        self.f_debug_out.write('Topic {0} {1}: {2}\n'.format(
            topic_id, math_utils.entropy(distribution, base=2, normalize=True), zip(terms, term_entropies)))

        top_ranked_indices = np.argsort(distribution)[::-1]

        self.rank_callback(topic_id, top_ranked_indices, distribution[top_ranked_indices])

    def process_ranked(self, payload, top_ranked_indices, top_ranked_values, term_entropies, entropy, mass,
                       topic_id):
        """The same observable behaviour as ``process`` from a device ranking (``LogLinearPredictFn.rank`` /
        ``rank_distributions``): the (T, E) per-term distributions never reach the host."""
        assert topic_id not in self.topic_projections
        self.topic_projections[topic_id] = None           # the reference keeps the raw (T, E) result; nothing reads it
        terms = [self.tokens[token_id] for token_id in payload]

        if not np.isclose(mass, 1.0):
            logging.error('Encountered non-normalized distribution for topic "%s" (mass=%.10f).', topic_id, mass)

        self.f_debug_out.write('Topic {0} {1}: {2}\n'.format(
            topic_id, entropy, zip(terms, [float(v) for v in term_entropies])))

        self.rank_callback(topic_id, np.asarray(top_ranked_indices, dtype=np.int64), np.asarray(top_ranked_values))

    def should_average_input(self):
        return False


def rank_distributions(distributions, top=None):
    """Device ranking of queries given their per-term distributions [(T_j, E) float32, ...] -- the input contract
    of ``LogLinearCallback.process`` (bin/query.py:204).  Returns what ``LogLinearPredictFn.rank`` returns."""
    from sert_b200 import _native as N
    E = distributions[0].shape[1]
    top = E if top is None else min(int(top), E)
    terms = np.array([d.shape[0] for d in distributions], np.int32)
    first = (np.cumsum(terms) - terms).astype(np.int32)
    stacked = np.ascontiguousarray(np.concatenate(distributions, axis=0), dtype=np.float32)
    nq = len(distributions)
    idx = np.empty((nq, top), np.int32)
    rel = np.empty((nq, top), np.float32)
    term_entropy = np.empty(int(terms.sum()), np.float32)
    entropy, mass = np.empty(nq, np.float32), np.empty(nq, np.float32)
    N.check(N.load().sert_ll_rank_distributions(
        N.host_ptr(stacked), stacked.shape[0], E, N.host_ptr(first), N.host_ptr(terms), nq, top, N.host_ptr(idx),
        N.host_ptr(rel), N.host_ptr(term_entropy), N.host_ptr(entropy), N.host_ptr(mass)))
    return [(idx[j], rel[j], term_entropy[first[j]:first[j] + terms[j]], float(entropy[j]), float(mass[j]))
            for j in range(nq)]


class VectorSpaceCallback(Callback):

    def __init__(self, entity_representations, *args, **kwargs):
        scorer_factory = kwargs.pop('scorer_factory', None)
        super(VectorSpaceCallback, self).__init__(*args, **kwargs)

        logging.info('Initializing k-NN for entity representations of shape %s.', entity_representations.shape)

        num_entities = entity_representations.shape[0]
        n_neighbors = self.args.top

        if n_neighbors is None:
            logging.warning('Parameter k not set; defaulting to all entities (k=%d).', num_entities)
        elif n_neighbors > num_entities:
            logging.warning('Parameter k exceeds number of entities; defaulting to all entities (k=%d).',
                            num_entities)
            n_neighbors = None

        # cosine similarity == Euclidean distance between L2-normalised vectors (bin/query.py:262-274)
        self.entity_representation_distance = 'euclidean'
        self.normalize_representations = True

        entity_representations /= np.linalg.norm(entity_representations, axis=1)[:, np.newaxis]
        logging.debug('Term projections will be normalized.')

        self.entity_representations = entity_representations
        self.n_neighbors = n_neighbors

        if n_neighbors:
            self.num_candidates = min(num_entities, n_neighbors + CANDIDATE_MARGIN)
            self.entity_avg = entity_representations.mean(axis=1)
            logging.info('Using %s as distance metric in entity space with the device top-k scorer '
                         '(k=%d, %d candidates re-ranked on the host).',
                         self.entity_representation_distance, n_neighbors, self.num_candidates)
        else:
            self.num_candidates = None
            logging.info('Using %s as distance metric in entity space.', self.entity_representation_distance)

        if scorer_factory is None:
            from sert_b200.scoring import EntityScorer
            scorer_factory = EntityScorer
        # room to ask for more candidates when near-ties at the k-th place demand it (query())
        self.max_candidates = min(num_entities, max(4 * (self.num_candidates or 1), 1024)) if n_neighbors else 1
        self.scorer = scorer_factory(entity_representations, normalise=False, max_queries=1024,
                                     max_k=max(self.max_candidates, 1))
        self.entity_neighbors = self.scorer if n_neighbors else None

    # -- candidate search -------------------------------------------------------------------------
    def query(self, centroids):
        """(distances, indices) with the reference's meaning: per row, the k nearest entities by Euclidean
        distance in ascending order (k = --top), or all entities when --top is unset (bin/query.py:304-318).

        The device ranks by float32 inner product; the reference's tree search ranks by float64 Euclidean distance
        between the float32 vectors.  A candidate set of m rows (device order) provably contains the k nearest when
        no row OUTSIDE it can be closer than the k-th nearest inside it: an outside row scores at most the m-th
        device score s_m, so its squared distance is at least |e|^2_min + |q|^2 - 2 (s_m + eps) with eps the float32
        summation error bound (d + 2) 2^-24 |q| |e|_max.  Rows for which that bound does not clear the k-th distance
        (masses of near-ties around the k-th place) are asked again with four times the candidates, finally with the
        dense scores of all entities."""
        centroids = np.ascontiguousarray(centroids, dtype=np.float32)
        E = self.entity_representations
        num_rows = centroids.shape[0]
        k = self.n_neighbors or E.shape[0]
        distances = np.full((num_rows, k), np.inf, dtype=np.float64)
        indices = np.full((num_rows, k), -1, dtype=np.int64)
        if not self.n_neighbors:
            scores = self.scorer.scores(centroids)
            cand_idx = np.argsort(-scores, axis=1, kind='stable')
            for row in range(num_rows):
                self._select(E, centroids[row], cand_idx[row], k, distances, indices, row)
            return distances, indices
        if not hasattr(self, '_norm_sq'):
            sq = np.einsum('ij,ij->i', E.astype(np.float64), E.astype(np.float64))
            self._norm_sq = (float(sq.min()), float(np.sqrt(sq.max())))
        e_sq_min, e_norm_max = self._norm_sq
        pending = np.arange(num_rows)
        m = self.num_candidates
        while pending.size:
            if m >= E.shape[0]:                                # every entity is a candidate: nothing left to prove
                scores = self.scorer.scores(centroids[pending])
                order = np.argsort(-scores, axis=1, kind='stable')
                for j, row in enumerate(pending):
                    self._select(E, centroids[row], order[j], k, distances, indices, row)
                break
            cand_idx, cand_score = self.scorer.topk(centroids[pending], m)
            unresolved = []
            for j, row in enumerate(pending):
                cands = cand_idx[j]
                present = cands >= 0
                kth = self._select(E, centroids[row], cands[present], k, distances, indices, row)
                if not present.all():                          # the scorer ran out of rows: all of them are candidates
                    continue
                q64 = centroids[row].astype(np.float64)
                q_sq = float(q64 @ q64)
                eps = (E.shape[1] + 2) * 2.0 ** -24 * np.sqrt(q_sq) * e_norm_max
                outside_sq = e_sq_min + q_sq - 2.0 * (float(cand_score[j, -1]) + eps)
                if not outside_sq > kth * kth:
                    unresolved.append(row)
            pending = np.asarray(unresolved, dtype=np.int64)
            m = min(4 * m, self.max_candidates) if m < self.max_candidates else E.shape[0]
        return distances, indices

    @staticmethod
    def _select(E, centroid, cands, k, distances, indices, row):
        """The k nearest of `cands` by float64 distance, ascending, into row `row`; returns the k-th distance."""
        diff = E[cands].astype(np.float64) - centroid.astype(np.float64)
        dist = np.sqrt(np.einsum('ij,ij->i', diff, diff))
        order = np.argsort(dist, kind='stable')[:k]
        assert len(order) == k, 'the scorer returned fewer candidates than neighbours were asked for'
        indices[row, :] = cands[order]
        distances[row, :] = dist[order]
        return float(dist[order[-1]])

    # -- ranking ----------------------------------------------------------------------------------
    def _normalise(self, term_projections):
        if term_projections.ndim == 1:
            term_projections = term_projections.reshape(1, -1)

        _, entity_representation_size = term_projections.shape
        assert entity_representation_size == self.model_args.entity_representation_size

        term_projections_l2_norm = np.linalg.norm(term_projections, axis=1)[:, np.newaxis]
        term_projections /= term_projections_l2_norm
        return term_projections

    def _rank(self, term_projection, term_indices):
        """Relevance of every candidate exactly as bin/query.py:344-365 computes it."""
        E = self.entity_representations
        matching_scores = (E[term_indices, :] * term_projection).sum(axis=1)      # == np.sum(e * q) per row
        matching_scores = (matching_scores + 1.0) / 2.0
        candidates = collections.defaultdict(float)
        for candidate, matching_score in zip(term_indices.tolist(), matching_scores):
            candidates[candidate] += matching_score
        top_ranked_indices, top_ranked_values = map(np.array, zip(
            *sorted(candidates.items(), reverse=True, key=operator.itemgetter(1))))
        return top_ranked_indices, top_ranked_values

    def process(self, payload, result, topic_id):
        terms = [self.tokens[token_id] for token_id in payload]

        term_projections = self._normalise(inference.aggregate_distribution(result, mode='identity', axis=0))

        logging.debug('Querying kneighbors for %s.', terms)

        distances, indices = self.query(term_projections)

        assert indices.shape[0] == term_projections.shape[0]
        assert indices.shape[0] == 1

        top_ranked_indices, top_ranked_values = self._rank(term_projections[0, :], indices[0, :])

        self.rank_callback(topic_id, top_ranked_indices, top_ranked_values)

    def process_many(self, payloads, projections, kwargs_list):
        """All submitted queries with one device scorer call (EmbeddingMapper.process)."""
        projections = np.array(projections, dtype=np.float32, copy=True)
        for projection, kwargs in zip(projections, kwargs_list):
            self._register(projection, kwargs['topic_id'])
        projections = self._normalise(projections)
        _, indices = self.query(projections)
        for row, kwargs in enumerate(kwargs_list):
            top_ranked_indices, top_ranked_values = self._rank(projections[row, :], indices[row, :])
            self.rank_callback(kwargs['topic_id'], top_ranked_indices, top_ranked_values)

    def should_average_input(self):
        return True


def compute_normalised_entropy(distribution, base=2):
    assert distribution.ndim == 2

    assert np.allclose(distribution.sum(axis=1), 1.0)

    return [math_utils.entropy(distribution[i, :], base=base, normalize=True)
            for i in range(distribution.shape[0])]


def _utf8_blob(strings):
    """(concatenated UTF-8 bytes as uint8 array, int64 offsets of len(strings) + 1)."""
    encoded = [s if isinstance(s, bytes) else str(s).encode('utf8') for s in strings]
    offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in encoded], out=offsets[1:])
    blob = np.frombuffer(b''.join(encoded) or b'\0', dtype=np.uint8)
    return blob, offsets


def write_topk_run(model_name, topic_ids, top_indices, relevances, entity_indices_inv, out_f,
                   max_objects_per_query=sys.maxsize):
    """Entity-finding run file (`<run_out>_ef`, bin/query.py:149-156) straight from a batched device ranking:
    `top_indices` (Q, k) internal entity rows as returned by scoring.EntityScorer.topk (-1 = no row, shards shorter
    than k) and `relevances` (Q, k).  Same bytes as feeding ranker_callback / trec_utils.write_run topic by topic:
    one lexsort orders every topic by descending (relevance, entity id) (trec_utils.py:560-561) and the library's
    host-side formatter (include/sert_b200.h: sert_format_run) prints the lines, relevance as Python's repr(float)."""
    from sert_b200 import _native as N
    top_indices = np.asarray(top_indices)
    rel = np.asarray(relevances)
    num_topics, k = top_indices.shape
    present = top_indices >= 0
    assert (present[:, :-1] >= present[:, 1:]).all(), 'missing rows must trail the list'
    counts = present.sum(axis=1)
    for q in np.flatnonzero(counts == 0):
        logging.warning('Received empty ranking for %s; ignoring.', topic_ids[q])
    names = [entity_indices_inv[i] for i in range(len(entity_indices_inv))]
    lookup = np.array(names, dtype=str)
    safe = np.where(present, top_indices, 0)
    # descending (relevance, entity id); unused slots behind every real assessment
    order = np.lexsort((np.where(present, lookup[safe], ''), np.where(present, rel, -np.inf)), axis=-1)[:, ::-1]
    keep = np.arange(k)[None, :] < np.minimum(counts, max_objects_per_query)[:, None]
    line_subject = np.broadcast_to(np.arange(num_topics, dtype=np.int32)[:, None], (num_topics, k))[keep]
    line_object = np.take_along_axis(safe, order, axis=1)[keep].astype(np.int32)
    line_rank = np.broadcast_to(np.arange(1, k + 1, dtype=np.int32)[None, :], (num_topics, k))[keep]
    line_rel = np.take_along_axis(rel, order, axis=1)[keep].astype(np.float64)      # '{0}'.format widens to double
    subject_blob, subject_off = _utf8_blob(topic_ids)
    object_blob, object_off = _utf8_blob(names)
    model = str(model_name).encode('utf8')
    n = int(line_rel.size)
    longest = int(np.diff(subject_off).max(initial=0) + np.diff(object_off).max(initial=0)) + len(model) + 64
    out = np.empty(max(1, n * longest), dtype=np.uint8)
    written = N.load().sert_format_run(
        N.host_ptr(subject_blob), N.host_ptr(subject_off), N.host_ptr(object_blob), N.host_ptr(object_off),
        N.host_ptr(np.ascontiguousarray(line_subject)), N.host_ptr(np.ascontiguousarray(line_object)),
        N.host_ptr(np.ascontiguousarray(line_rank)), N.host_ptr(np.ascontiguousarray(line_rel)), n, model,
        N.host_ptr(out), out.size)
    if written < 0:
        N.check(int(written))
    out_f.write(out[:written].tobytes().decode('utf8'))


class RunCollector(object):
    """``ranker_callback`` of bin/query.py:83-92 without per-assessment tuples: keeps, per topic, the arrays the
    ranking callback hands over, and writes both run files of bin/query.py:149-156 from flat arrays --
    ``<run_out>_ef`` (entity finding: topics rank entities) and ``<run_out>_ep`` (entity profiling: entities rank
    topics).  Same bytes as feeding ``trec_utils.write_run`` the two dictionaries the reference builds: subjects in
    first-insertion order, assessments by descending (relevance, object id) (trec_utils.py:560-561), relevance
    printed as '{0}'.format(value) (the float64 repr of the value, also for numpy.float32).  One lexsort per file and
    the library's host-side formatter instead of Q x E tuples, two tuple sorts per subject and a str.format per line."""

    def __init__(self, entity_indices_inv):
        self.entity_indices_inv = entity_indices_inv
        self.topic_ids, self.indices, self.values = [], [], []

    def __call__(self, topic_id, top_ranked_indices, top_ranked_values):
        self.topic_ids.append(topic_id)
        self.indices.append(np.asarray(top_ranked_indices, dtype=np.int64).ravel())
        self.values.append(np.asarray(top_ranked_values).ravel())

    def _flat(self):
        lengths = np.array([len(i) for i in self.indices], dtype=np.int64)
        topic_of = np.repeat(np.arange(len(self.indices), dtype=np.int64), lengths)
        entity = np.concatenate(self.indices) if self.indices else np.zeros(0, np.int64)
        value = (np.concatenate([v.astype(np.float64) for v in self.values]) if self.values
                 else np.zeros(0, np.float64))
        return topic_of, entity, value

    @staticmethod
    def _string_rank(strings):
        """Rank of every string in sorted order (ties share a rank): a numeric stand-in for string comparison."""
        _, inverse = np.unique(np.asarray(strings, dtype=str), return_inverse=True)
        return inverse.astype(np.int64)

    def write(self, model_name, out_ep, out_ef, max_objects_per_query=sys.maxsize):
        from sert_b200 import _native as N
        topic_of, entity, value = self._flat()
        if entity.size == 0:
            return
        entity_ids = sorted(set(entity.tolist()))
        entity_pos = {e: i for i, e in enumerate(entity_ids)}
        entity_names = [self.entity_indices_inv[e] for e in entity_ids]
        entity_of = np.fromiter((entity_pos[e] for e in entity.tolist()), dtype=np.int64, count=entity.size)
        topic_rank = self._string_rank(self.topic_ids)
        entity_rank = self._string_rank(entity_names)
        topic_blob, topic_off = _utf8_blob(self.topic_ids)
        entity_blob, entity_off = _utf8_blob(entity_names)
        model = str(model_name).encode('utf8')
        lib = N.load()

        def emit(subject_of, subject_order_key, object_of, object_rank, subject_blob, subject_off, object_blob,
                 object_off, out_f):
            # lines ordered by subject (insertion order), then descending (relevance, object id)
            order = np.lexsort((-object_rank[object_of], -value, subject_order_key[subject_of]))
            subj, obj, val = subject_of[order], object_of[order], value[order]
            starts = np.flatnonzero(np.r_[True, subj[1:] != subj[:-1]])
            rank = np.arange(subj.size) - np.repeat(starts, np.diff(np.r_[starts, subj.size])) + 1
            keep = rank <= max_objects_per_query
            subj, obj, val, rank = subj[keep], obj[keep], val[keep], rank[keep]
            n = int(subj.size)
            longest = int(np.diff(subject_off).max(initial=0) + np.diff(object_off).max(initial=0)) + len(model) + 64
            out = np.empty(max(1, n * longest), dtype=np.uint8)
            written = lib.sert_format_run(
                N.host_ptr(subject_blob), N.host_ptr(subject_off), N.host_ptr(object_blob), N.host_ptr(object_off),
                N.host_ptr(np.ascontiguousarray(subj, dtype=np.int32)), N.host_ptr(np.ascontiguousarray(obj, dtype=np.int32)),
                N.host_ptr(np.ascontiguousarray(rank, dtype=np.int32)), N.host_ptr(np.ascontiguousarray(val)), n, model,
                N.host_ptr(out), out.size)
            if written < 0:
                N.check(int(written))
            out_f.write(out[:written].tobytes().decode('utf8'))

        # entity profiling: subjects = entities in order of first appearance (dict insertion order of the reference)
        first_seen = np.full(len(entity_ids), entity.size, dtype=np.int64)
        np.minimum.at(first_seen, entity_of, np.arange(entity.size))
        entity_order = np.empty(len(entity_ids), dtype=np.int64)
        entity_order[np.argsort(first_seen, kind='stable')] = np.arange(len(entity_ids))
        emit(entity_of, entity_order, topic_of, topic_rank, entity_blob, entity_off, topic_blob, topic_off, out_ep)
        # entity finding: subjects = topics in call order
        emit(topic_of, np.arange(len(self.topic_ids), dtype=np.int64), entity_of, entity_rank, topic_blob, topic_off,
             entity_blob, entity_off, out_ef)

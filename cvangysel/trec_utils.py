"""TREC topic parsing, query tokenisation and run writing
(reference: cvangysel-common/py/cvangysel/trec_utils.py:189-204, 403-440, 531-580).

write_run defines the observable ranked-list output of bin/query.py: assessments are ordered by the
tuple (relevance, object id) DESCENDING, i.e. ties in relevance break on the id string, descending
(trec_utils.py:560-561), ranks start at 1.
"""
import collections
import io
import logging
import re
import sys

from cvangysel import io_utils

remove_parentheses_re = re.compile(r'\((.*)\)')


def parse_query(unsplitted_terms):
    """trec_utils.py:192-204: drop one level of parentheses, turn '/' and '-' into spaces, then the
    alphanumeric -> latin -> lower-case -> whitespace-token pipeline of io_utils."""
    assert isinstance(unsplitted_terms, str)
    text = remove_parentheses_re.sub(r'\1', unsplitted_terms.strip())
    text = text.replace('/', ' ').replace('-', ' ')
    return list(io_utils.token_stream(
        io_utils.lowercased_stream(
            io_utils.filter_non_latin_stream(
                io_utils.filter_non_alphanumeric_stream(iter(text)))),
        eos_chars=[]))


def parse_topics(file_or_files, max_topics=sys.maxsize, delimiter=';'):
    """trec_utils.py:403-440: lines `topic_id;terms`; later duplicates overwrite earlier ones."""
    assert max_topics >= 0 or max_topics is None
    topics = collections.OrderedDict()
    if not isinstance(file_or_files, (list, tuple)):
        file_or_files = list(file_or_files) if hasattr(file_or_files, '__iter__') and \
            not isinstance(file_or_files, io.IOBase) else [file_or_files]
    for f in file_or_files:
        assert isinstance(f, io.IOBase)
        for line in f:
            assert isinstance(line, str)
            line = line.strip()
            if not line:
                continue
            topic_id, terms = line.split(delimiter, 1)
            if topic_id in topics and topics[topic_id] != terms:
                logging.error('Duplicate topic "%s" (%s vs. %s).', topic_id, topics[topic_id], terms)
            topics[topic_id] = terms
            if max_topics > 0 and len(topics) >= max_topics:
                break
    return topics


def write_run(model_name, data, out_f, max_objects_per_query=sys.maxsize, skip_sorting=False):
    """trec_utils.py:531-580.  data: {subject_id: iterable of (relevance, object_id)}."""
    for subject_id, object_assesments in data.items():
        if not object_assesments:
            logging.warning('Received empty ranking for %s; ignoring.', subject_id)
            continue
        assert isinstance(object_assesments[0][1], (str, bytes))
        if not skip_sorting:
            object_assesments = sorted(object_assesments, reverse=True)
        if max_objects_per_query < sys.maxsize:
            object_assesments = object_assesments[:max_objects_per_query]
        if isinstance(subject_id, bytes):
            subject_id = subject_id.decode('utf8')
        lines = []
        for rank, (relevance, object_id) in enumerate(object_assesments, start=1):
            if isinstance(object_id, bytes):
                object_id = object_id.decode('utf8')
            lines.append('{0} Q0 {1} {2} {3} {4}\n'.format(subject_id, object_id, rank, relevance, model_name))
        out_f.write(''.join(lines))


def write_run_arrays(model_name, subject_ids, object_ids, relevances, out_f, counts=None,
                     max_objects_per_query=sys.maxsize):
    """write_run for rankings that already are arrays (the output of a batched top-k scorer): row q of `object_ids`
    (Q, k) str and `relevances` (Q, k) float holds the `counts[q]` (default k) assessed objects of `subject_ids[q]`.
    Produces byte for byte what write_run writes for {subject: [(relevance, object_id), ...]} -- the descending
    (relevance, object id) order of trec_utils.py:560-561 and '{0}'.format of the relevance values -- with one lexsort
    for all subjects instead of a tuple sort per subject, and no per-assessment tuples to build."""
    import numpy as np
    rel = np.asarray(relevances)
    ids = np.asarray(object_ids, dtype=str)
    assert rel.ndim == 2 and ids.shape == rel.shape and len(subject_ids) == rel.shape[0]
    num_subjects, k = rel.shape
    counts = np.full(num_subjects, k, dtype=np.int64) if counts is None else np.asarray(counts, dtype=np.int64)
    valid = np.arange(k)[None, :] < counts[:, None]
    for q in np.flatnonzero(counts == 0):
        logging.warning('Received empty ranking for %s; ignoring.', subject_ids[q])
    if num_subjects == 0 or k == 0:
        return
    # unused slots sort behind every real assessment
    rel_key = np.where(valid, rel, -np.inf)
    ids_key = np.where(valid, ids, '')
    order = np.lexsort((ids_key, rel_key), axis=-1)[:, ::-1]          # descending (relevance, object id)
    rel_sorted = np.take_along_axis(rel, order, axis=1)
    ids_sorted = np.take_along_axis(ids, order, axis=1)
    # Python floats / str from here: repr(float) is what '{0}'.format(numpy scalar) emits (float.__format__ of the
    # value widened to double), and one f-string per line over plain lists beats numpy's fixed-width string arrays
    rel_rows = rel_sorted.astype(np.float64).tolist()
    ids_rows = ids_sorted.tolist()
    limits = np.minimum(counts, max_objects_per_query).tolist()
    tail = ' {0}\n'.format(model_name)
    chunks = []
    for q, subject in enumerate(subject_ids):
        if isinstance(subject, bytes):
            subject = subject.decode('utf8')
        head = subject + ' Q0 '
        n = limits[q]
        chunks.append(''.join([f'{head}{o} {r} {v}{tail}'
                               for r, (o, v) in enumerate(zip(ids_rows[q][:n], rel_rows[q][:n]), start=1)]))
    out_f.write(''.join(chunks))

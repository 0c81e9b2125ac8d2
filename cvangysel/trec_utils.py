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

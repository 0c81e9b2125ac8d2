"""The step that feeds the hot path: packing windowed instances and their entity labels
into the arrays of `data.npz`, and the `data.npz` / `meta` writers.

Mirrors bin/prepare.py:373-416 (writers, instance weights) and :543-599 (`instances_and_labels_to_arrays`), same
signatures and outputs; the per-non-zero Python loops of the reference (one `sorted(...)`, three `list.extend` per
instance) are replaced by one pass that fills flat numpy arrays and a single lexsort.  The windowing of a document's
token stream (cvangysel io_utils.py:151-211, driven from bin/prepare.py:499-512) is here as strided array views;
corpus reading and the character-level tokeniser stay out of scope.
"""
import logging
import pickle

import numpy as np
from scipy import sparse


def windows_from_ids(ids, window_size, stride=1, padding_id=None):
    """All windows the reference's generator yields for ONE eos-free run of in-vocabulary word ids
    (cvangysel io_utils.py:151-211): full windows at offsets 0, stride, 2*stride, ...; when tokens remain behind the
    last full window (or the run is shorter than a window) and `padding_id` is given, one more window padded to
    `window_size`.  Returns an (n, window_size) array of ids' dtype; a strided view is copied once."""
    ids = np.asarray(ids)
    assert ids.ndim == 1 and 1 <= stride <= window_size
    n = ids.shape[0]
    if n >= window_size:
        full = np.lib.stride_tricks.sliding_window_view(ids, window_size)[::stride]
        consumed = (full.shape[0] - 1) * stride + window_size       # tokens read when the last full window was yielded
    else:
        full = np.empty((0, window_size), dtype=ids.dtype)
        consumed = 0
    new_tokens = n - consumed
    if padding_id is None or new_tokens <= 0:
        return np.ascontiguousarray(full)
    # the generator keeps window_size - stride old tokens, reads the new ones, then pads
    start = full.shape[0] * stride if full.shape[0] else 0
    tail = np.full(window_size, padding_id, dtype=ids.dtype)
    tail[:n - start] = ids[start:]
    return np.concatenate([full, tail[None, :]], axis=0)


def document_windows(tokens, words, window_size, stride=1, padding_token=None, eos_token='</s>', eos_chars=()):
    """windowed_translated_token_stream (io_utils.py:151-211) as arrays: out-of-vocabulary tokens are dropped, an
    end-of-sentence token clears the window (only full windows survive from the runs before it), the run that ends
    the stream gets the padded tail.  `words`: {token: Word(id, count)}.  Returns (n, window_size) int64 ids."""
    assert eos_token in words and (padding_token is None or padding_token in words)
    runs, current = [], []
    for token in tokens:
        if token in eos_chars or token == eos_token:
            runs.append(current)
            current = []
        elif token in words:
            current.append(words[token].id)
    padding_id = None if padding_token is None else words[padding_token].id
    parts = [windows_from_ids(np.asarray(run, dtype=np.int64), window_size, stride) for run in runs]
    parts.append(windows_from_ids(np.asarray(current, dtype=np.int64), window_size, stride, padding_id))
    return np.concatenate(parts, axis=0) if parts else np.empty((0, window_size), dtype=np.int64)


def instances_and_labels_to_arrays(instances, window_size, class_mapping, instance_dtype, shuffle):
    """bin/prepare.py:543-599.  instances: list of (doc_id, window (sequence of word ids), {entity_id: mass});
    returns x (N, window_size) `instance_dtype` and y CSR float32 (N, len(class_mapping)), columns sorted per row."""
    assert isinstance(instances, list)
    num_classes = len(class_mapping)
    if shuffle:
        logging.info('Shuffling instance and label pairs.')
        np.random.shuffle(instances)              # same in-place shuffle, same RNG consumption as the reference
    else:
        logging.info('Instances are not shuffled.')
    num_instances = len(instances)

    logging.info('Constructing dense instance matrix.')
    x = np.fromiter((element for _, instance, _ in instances for element in instance),
                    dtype=instance_dtype, count=num_instances * window_size)
    x = x.reshape((num_instances, window_size))

    logging.info('Constructing sparse label matrix.')
    counts = np.fromiter((len(label) for _, _, label in instances), dtype=np.int64, count=num_instances)
    assert num_instances == 0 or counts.min() > 0      # the reference's zip(*sorted(...)) fails on an empty label
    nnz = int(counts.sum())
    cols = np.fromiter((class_mapping[entity_id] for _, _, label in instances for entity_id in label),
                       dtype=np.int64, count=nnz)
    data = np.fromiter((mass for _, _, label in instances for mass in label.values()),
                       dtype=np.float32, count=nnz)
    rows = np.repeat(np.arange(num_instances, dtype=np.int64), counts)
    order = np.lexsort((cols, rows))                   # (class index) ascending inside every row, like sorted()
    indptr = np.zeros(num_instances + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    y = sparse.csr_matrix((data[order], cols[order], indptr), shape=(num_instances, num_classes))
    y.sum_duplicates()                                 # csr_matrix((data, (row, col))) sums duplicates too
    y.indices = y.indices.astype(np.int32, copy=False)
    y.indptr = y.indptr.astype(np.int32, copy=False)
    return x, y


def pack_document_windows(doc_windows, doc_entities, class_mapping, instance_dtype, shuffle,
                          max_document_length=None):
    """instances_and_labels_to_arrays (+ the w_train expression of bin/prepare.py:395-399) for instances that never
    become Python tuples: `doc_windows[d]` is the (n_d, W) window array of document d (document_windows above),
    `doc_entities[d]` the entity ids associated with it; every window of a document carries the candidate-centric
    label of bin/prepare.py:436-439,516-523 (mass 1/len(entities) on each).  Documents are taken in the given order
    and, with `shuffle`, permuted exactly like np.random.shuffle(instances) permutes the reference's list.
    Returns x, y (CSR float32) and, when `max_document_length` is given, w = max_document_length / n_d per window."""
    num_classes = len(class_mapping)
    counts = np.array([w.shape[0] for w in doc_windows], dtype=np.int64)
    num_instances = int(counts.sum())
    window_size = doc_windows[0].shape[1] if len(doc_windows) else 0
    x = (np.concatenate(doc_windows, axis=0) if num_instances else np.empty((0, window_size))).astype(instance_dtype)
    label_cols = [np.sort(np.fromiter((class_mapping[e] for e in ents), dtype=np.int64, count=len(ents)))
                  for ents in doc_entities]
    label_sizes = np.array([c.shape[0] for c in label_cols], dtype=np.int64)
    assert len(label_cols) == len(doc_windows) and (num_instances == 0 or label_sizes[counts > 0].min() > 0)
    # row r of document d: the document's sorted columns, mass 1/|label|
    row_sizes = np.repeat(label_sizes, counts)
    indptr = np.zeros(num_instances + 1, dtype=np.int64)
    np.cumsum(row_sizes, out=indptr[1:])
    indices = (np.concatenate([np.tile(c, n) for c, n in zip(label_cols, counts)])
               if num_instances else np.empty(0, dtype=np.int64))
    data = np.repeat((1.0 / np.maximum(row_sizes, 1)).astype(np.float32), row_sizes)
    y = sparse.csr_matrix((data, indices, indptr), shape=(num_instances, num_classes))
    w = None
    if max_document_length is not None:
        w = np.repeat((float(max_document_length) / np.maximum(counts, 1)).astype(np.float32), counts)
    if shuffle:
        logging.info('Shuffling instance and label pairs.')
        order = np.arange(num_instances)
        np.random.shuffle(order)              # the permutation np.random.shuffle applies to a list of this length
        x, y = x[order], y[order]
        w = None if w is None else w[order]
    y.indices = y.indices.astype(np.int32, copy=False)
    y.indptr = y.indptr.astype(np.int32, copy=False)
    return x, y, w


def instance_weights(instances, instances_per_document, max_document_length):
    """bin/prepare.py:395-399: w = max_document_length / (#instances of the instance's document), float32."""
    return np.fromiter((float(max_document_length) / instances_per_document[doc_id] for doc_id, _, _ in instances),
                       np.float32, len(instances))


def retained_entities(instances_per_entity):
    """bin/prepare.py:351-364: entities with at least one instance, indexed in iteration order."""
    entity_indices, entity_indices_inv = {}, {}
    for entity_id, num_instances in instances_per_entity.items():
        if not num_instances:
            continue
        entity_index = len(entity_indices)
        entity_indices[entity_id] = entity_index
        entity_indices_inv[entity_index] = entity_id
    return entity_indices, entity_indices_inv


def write_meta(meta_output, args, words, tokens, entity_indices_inv, documents_per_entity):
    """bin/prepare.py:373-376: five sequential pickles."""
    with open(meta_output, 'wb') as f:
        for obj in (args, words, tokens, entity_indices_inv, documents_per_entity):
            pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)


def read_meta(meta_path):
    with open(meta_path, 'rb') as f:
        return tuple(pickle.load(f) for _ in range(5))


def write_data(data_output, x_train, y_train, x_validate, y_validate, w_train=None):
    """bin/prepare.py:383-416: np.savez with keys x_train, y_train (0-d object array holding the CSR matrix),
    [w_train], x_validate, y_validate -- what bin/train.py:79-90 reads back."""
    data = {'x_train': x_train, 'y_train': y_train}
    if w_train is not None:
        assert w_train.shape == (x_train.shape[0],)
        data['w_train'] = w_train
    data['x_validate'] = x_validate
    data['y_validate'] = y_validate
    with open(data_output, 'wb') as f:
        np.savez(f, **data)


def pack(training_instances, validation_instances, window_size, num_words, instances_per_entity,
         instances_per_document=None, max_document_length=None, shuffle=True):
    """The tail of bin/prepare.py's main() (:351-413) as one call: entity indexing, both array pairs, weights."""
    entity_indices, entity_indices_inv = retained_entities(instances_per_entity)
    instance_dtype = np.min_scalar_type(num_words - 1)          # bin/prepare.py:380
    x_train, y_train = instances_and_labels_to_arrays(training_instances, window_size, entity_indices,
                                                      instance_dtype, shuffle)
    w_train = None
    if instances_per_document is not None:
        w_train = instance_weights(training_instances, instances_per_document, max_document_length)
    x_validate, y_validate = instances_and_labels_to_arrays(validation_instances, window_size, entity_indices,
                                                            instance_dtype, shuffle)
    return dict(x_train=x_train, y_train=y_train, w_train=w_train, x_validate=x_validate, y_validate=y_validate,
                entity_indices_inv=entity_indices_inv)

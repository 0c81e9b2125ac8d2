"""The step that feeds the hot path: packing windowed instances and their entity labels
into the arrays of `data.npz`, and the `data.npz` / `meta` writers.

Mirrors bin/prepare.py:373-416 (writers, instance weights) and :543-599 (`instances_and_labels_to_arrays`), same
signatures and outputs; the per-non-zero Python loops of the reference (one `sorted(...)`, three `list.extend` per
instance) are replaced by one pass that fills flat numpy arrays and a single lexsort.  Corpus reading, tokenising and
windowing (bin/prepare.py:474-533, cvangysel io_utils) stay out of scope.
"""
import logging
import pickle

import numpy as np
from scipy import sparse


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

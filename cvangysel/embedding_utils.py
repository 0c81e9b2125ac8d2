"""word2vec binary reader/writer (reference: cvangysel-common/py/cvangysel/embedding_utils.py:15-144),
used by bin/train.py --representation_initializer (bin/train.py:131-151)."""
import logging
import os
import struct

import numpy as np


def get_binary_representations_info(filename):
    with open(filename, 'rb') as f:
        vocabulary_size, vector_size = (int(v) for v in f.readline().strip().split())
    return vocabulary_size, vector_size


def load_binary_representations(filename, vocabulary=None):
    """Yields (lower-cased word, float vector) pairs; words outside `vocabulary` (if given) are skipped.
    A word is the bytes up to the next space; a vector is vector_size little-endian float32 values."""
    keep = None if vocabulary is None else set(vocabulary)
    with open(filename, 'rb') as f:
        vocabulary_size, vector_size = (int(v) for v in f.readline().strip().split())
        file_size = os.fstat(f.fileno()).st_size
        vector_bytes = struct.calcsize('f' * vector_size)
        reported = 0
        while f.tell() < file_size:
            progress = int(100 * float(f.tell()) / file_size)
            if progress % 10 == 0 and progress > reported:
                logging.info('Reading file %s with %d words (%d-dimensional): %d%% done.',
                             filename, vocabulary_size, vector_size, progress)
                reported = progress
            chars = []
            while True:
                char = f.read(1).decode()
                if char == ' ' or not char:
                    break
                chars.append(char)
            word = ''.join(chars).lower().strip()
            if not word and f.tell() == file_size:
                return                      # dangling whitespace at the end of the file
            buf = f.read(vector_bytes)
            if len(buf) < vector_bytes:
                logging.error('Encountered end-of-file before reading representation '
                              '(expected %d bytes, encountered %d bytes).', vector_bytes, len(buf))
                return
            if keep is not None and word not in keep:
                continue
            yield word, np.array(struct.unpack('f' * vector_size, buf))


def write_binary_representations(filename, words_and_representations):
    items = [(w, np.asarray(r, dtype=np.float32)) for w, r in words_and_representations]
    vector_size = items[0][1].size if items else 0
    with open(filename, 'wb') as f:
        f.write('{0} {1}\n'.format(len(items), vector_size).encode())
        for word, rep in items:
            assert rep.size == vector_size
            f.write(word.encode() + b' ' + rep.tobytes())

"""Host-side timing of the array-packing step (sert_b200/prepare.py) next to the reference's own functions when the
reference checkout is present (/root/reference, build container only):  python tools/prepare_bench.py [instances]"""
import collections
import importlib.util
import os
import sys
import time
import types
import warnings

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from sert_b200 import prepare  # noqa: E402

REF = '/root/reference'
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
W, V, n_ent = 10, 100000, 5000
rng = np.random.default_rng(1)
entity_ids = ['ent-%05d' % i for i in range(n_ent)]
class_mapping = {e: i for i, e in enumerate(entity_ids)}
x = rng.integers(0, V, (n, W))
ks = rng.choice([1, 2, 3], n, p=[0.8, 0.15, 0.05])
instances = []
for i in range(n):
    ents = rng.choice(n_ent, ks[i], replace=False)
    instances.append(('doc%d' % (i // 50), tuple(x[i].tolist()), {entity_ids[int(e)]: 1.0 / ks[i] for e in ents}))
t0 = time.perf_counter()
xa, ya = prepare.instances_and_labels_to_arrays(list(instances), W, class_mapping, np.uint32, False)
t_mine = time.perf_counter() - t0
print('instances_and_labels_to_arrays: %d instances, %d labels: %.2f s (this repo)' % (n, ya.nnz, t_mine))

# the same instances as per-document window arrays (50 consecutive instances share a document and its label)
doc_windows = [x[i:i + 50] for i in range(0, n, 50)]
doc_entities = [list(instances[i][2]) for i in range(0, n, 50)]
t0 = time.perf_counter()
xb, yb, wb = prepare.pack_document_windows(doc_windows, doc_entities, class_mapping, np.uint32, False, 50)
t_arr = time.perf_counter() - t0
print('pack_document_windows: %d documents -> %d instances: %.3f s (this repo, array-native)' % (
    len(doc_windows), xb.shape[0], t_arr))

Word = collections.namedtuple('Word', ['id', 'count'])
vocab = ['</s>', '<pad>'] + ['w%d' % i for i in range(5000)]
words = {t: Word(i, 1) for i, t in enumerate(vocab)}
docs = [[vocab[int(j)] for j in rng.integers(2, len(vocab), 400)] for _ in range(500)]
t0 = time.perf_counter()
total = sum(prepare.document_windows(d, words, W, 1, '<pad>').shape[0] for d in docs)
t_win = time.perf_counter() - t0
print('document_windows: 500 documents x 400 tokens -> %d windows: %.3f s (this repo)' % (total, t_win))

if os.path.isdir(REF):
    warnings.simplefilter('ignore')

    class _Stub(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith('__'):
                raise AttributeError(item)
            return type(item, (object,), {})
    for name in ('bs4', 'nltk', 'nltk.probability', 'nltk.corpus', 'gensim', 'sklearn.cross_validation'):
        sys.modules.setdefault(name, _Stub(name))
    sys.modules['nltk'].probability = sys.modules['nltk.probability']
    sys.modules['nltk'].corpus = sys.modules['nltk.corpus']
    for mod in [m for m in sys.modules if m == 'cvangysel' or m.startswith('cvangysel.')]:
        del sys.modules[mod]                                    # the reference's package, not this repo's shim
    sys.path[:0] = [os.path.join(REF, 'cvangysel-common', 'py')]
    spec = importlib.util.spec_from_file_location('ref_prepare', os.path.join(REF, 'bin', 'prepare.py'))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    t0 = time.perf_counter()
    xr, yr = ref.instances_and_labels_to_arrays(list(instances), W, class_mapping, np.uint32, False)
    t_ref = time.perf_counter() - t0
    same = (xr == xa).all() and (yr != ya).nnz == 0
    print('instances_and_labels_to_arrays: %.2f s (reference), identical output: %s, speed-up %.1fx' % (
        t_ref, same, t_ref / t_mine))
    from cvangysel import io_utils as ref_io
    t0 = time.perf_counter()
    total_ref = sum(len(list(ref_io.windowed_translated_token_stream(iter(d), W, words, eos_chars=[], stride=1,
                                                                       padding_token='<pad>'))) for d in docs)
    t_ref_win = time.perf_counter() - t0
    print('windowed_translated_token_stream: %.3f s (reference), same count: %s, speed-up %.1fx' % (
        t_ref_win, total_ref == total, t_ref_win / t_win))

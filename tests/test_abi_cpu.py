"""The C-ABI library loads without a GPU, exports every symbol the header declares, and fails loudly
(never falls back) when asked to compute without a device."""
import ctypes
import os
import re
import subprocess

import numpy as np

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'sert_b200.h')).read()
    return sorted(set(re.findall(r'SERT_API[^;(]*?\b(sert_\w+)\s*\(', text)))


def test_header_declares_and_library_exports_every_symbol(native_lib):
    from sert_b200 import _native as N
    declared = header_symbols()
    assert len(declared) >= 25
    assert sorted(N.SIGNATURES) == declared, 'ctypes table and header disagree'
    out = subprocess.run(['nm', '-D', '--defined-only', N.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    exported = set(re.findall(r' T (sert_\w+)', out))
    assert set(declared) <= exported
    assert native_lib.sert_abi_version() == 1


def test_arena_size_and_config_validation(native_lib):
    from sert_b200 import _native as N
    cfg = N.SertConfig(kind=N.KIND_VECTORSPACE, batch=4096, window=10, num_negatives=10, vocab=100000,
                       entities=50000, word_dim=128, entity_dim=128, lambda_=0.01, loss_slots=1024, seed=1,
                       inference_only=0, dtype_mode=0, reserved1=0)
    nbytes = N.c_size_t(0)
    N.check(native_lib.sert_model_arena_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)))
    params = 100000 * 128 + 50000 * 128 + 128 * 128 + 128
    assert 16 * params <= nbytes.value <= 16 * params + (64 << 20)    # theta, m, v, grad + workspaces
    cfg.inference_only = 1                                             # predict_fn models: no optimiser state / grads
    small = N.c_size_t(0)
    N.check(native_lib.sert_model_arena_bytes(ctypes.byref(cfg), ctypes.byref(small)))
    assert 4 * params <= small.value <= 4 * params + (64 << 20)
    cfg.inference_only = 0
    cfg.word_dim = 130                                                 # not a multiple of 4
    with pytest.raises(RuntimeError, match='multiple of 4'):
        N.check(native_lib.sert_model_arena_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)))
    cfg.word_dim, cfg.batch = 128, 0
    with pytest.raises(RuntimeError, match='batch_size'):
        N.check(native_lib.sert_model_arena_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)))
    sbytes = N.c_size_t(0)
    N.check(native_lib.sert_scorer_arena_bytes(1000000, 256, 10000, 128, ctypes.byref(sbytes)))
    assert sbytes.value >= 1000000 * 256 * 4


@pytest.mark.parametrize('V,E,dw,de,B', [(100000, 50000, 128, 128, 4096), (1500, 700, 128, 128, 256),
                                         (801, 333, 64, 48, 128), (40, 9000, 300, 128, 64), (5, 3, 4, 4, 8)])
def test_table_shard_plan_partitions_the_tables(native_lib, V, E, dw, de, B):
    """sert_table_shard_plan (host only): the pieces of sert_model_set_table_shard_comm tile the two tables without
    gaps, end on rows whose index is a multiple of 4, are nearly equal in floats, agree between the row view and the
    float view, and the instance bounds cut the batch into whole tiles of 8."""
    from sert_b200 import _native as N
    cfg = N.SertConfig(kind=N.KIND_VECTORSPACE, batch=B, window=5, num_negatives=4, vocab=V, entities=E, word_dim=dw,
                       entity_dim=de, lambda_=0.01, loss_slots=64, seed=1, inference_only=0, dtype_mode=0, reserved1=0)
    for world in range(1, 9):
        e = np.zeros(world + 1, np.int64)
        r = np.zeros(world + 1, np.int64)
        f = np.zeros(world + 1, np.int64)
        inst = np.zeros(world + 1, np.int32)
        N.check(native_lib.sert_table_shard_plan(ctypes.byref(cfg), world, N.host_ptr(e), N.host_ptr(r), N.host_ptr(f),
                                                 N.host_ptr(inst)))
        assert e[0] == 0 and r[0] == 0 and f[0] == 0 and e[-1] == E and r[-1] == V
        assert np.all(np.diff(e) >= 0) and np.all(np.diff(r) >= 0) and np.all(np.diff(f) >= 0)
        assert np.all((e % 4 == 0) | (e == E)) and np.all((r % 4 == 0) | (r == V))      # inside a table: whole 16-byte chunks
        owned = np.diff(e) * de + np.diff(r) * dw                     # floats of table rows per rank
        assert owned.sum() == E * de + V * dw
        # the float view covers the same rows (the arena pads each table to 64 floats: the pads ride with a neighbour)
        assert np.all(np.abs(np.diff(f) - owned) <= 128)
        ideal = (E * de + V * dw) / world
        assert np.all(np.abs(owned - ideal) <= 4 * max(dw, de))
        # a word-row piece starts only once the entity table is exhausted
        assert np.all((r[:-1] == 0) | (e[:-1] == E))
        assert inst[0] == 0 and inst[-1] == B and np.all(np.diff(inst) >= 0) and np.all(inst[:-1] % 8 == 0)
    with pytest.raises(RuntimeError, match='at most 8 ranks'):
        N.check(native_lib.sert_table_shard_plan(ctypes.byref(cfg), 9, N.host_ptr(np.zeros(10, np.int64)),
                                                 N.host_ptr(np.zeros(10, np.int64)), N.host_ptr(np.zeros(10, np.int64)),
                                                 N.host_ptr(np.zeros(10, np.int32))))


def test_no_cpu_fallback_without_device(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from sert_b200 import models
    import numpy as np
    with pytest.raises(RuntimeError, match='CUDA device'):
        models._NativeModel(0, 8, 2, 16, 4, 4)
    from sert_b200.scoring import EntityScorer
    with pytest.raises(RuntimeError, match='CUDA device'):
        EntityScorer(np.zeros((4, 4), np.float32))


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'sert_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.replace('the CPU oracle', ''), f


def test_run_formatter_prints_python_float_repr():
    """sert_format_run (host-only code of the library) must print relevance values exactly like '{0}'.format(float)
    -- Python's repr -- for every double: shortest round-trip digits, fixed notation for 1e-4 <= |v| < 1e16."""
    import numpy as np
    from sert_b200 import _native as N
    lib = N.load()
    rng = np.random.default_rng(99)
    values = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e-4, 1e-5, 9.999e-5, 1e15, 1e16, 9999999999999998.0, 1e17, 1e22,
              1.5e-5, 123456.789, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, float('inf'),
              float('-inf'), float('nan'), 0.30000000000000004, 100.0, 12345678901234567890.0, 2.5e-7, 3.0e10]
    values += (rng.random(2000) * 10.0 ** rng.integers(-30, 30, 2000) * rng.choice([-1, 1], 2000)).tolist()
    values += rng.random(2000).astype(np.float32).astype(np.float64).tolist()          # widened float32 relevances
    values += (rng.integers(-10 ** 6, 10 ** 6, 500) / 8.0).tolist()
    v = np.array(values, dtype=np.float64)
    n = v.size
    subjects, objects = ['topic-é'.encode('utf8')], [b'entity/007']
    s_blob = np.frombuffer(b''.join(subjects), dtype=np.uint8)
    o_blob = np.frombuffer(b''.join(objects), dtype=np.uint8)
    s_off = np.array([0, len(subjects[0])], dtype=np.int64)
    o_off = np.array([0, len(objects[0])], dtype=np.int64)
    zeros = np.zeros(n, dtype=np.int32)
    ranks = np.arange(1, n + 1, dtype=np.int32)
    out = np.empty(n * 128, dtype=np.uint8)
    written = lib.sert_format_run(N.host_ptr(s_blob), N.host_ptr(s_off), N.host_ptr(o_blob), N.host_ptr(o_off),
                                  N.host_ptr(zeros), N.host_ptr(zeros), N.host_ptr(ranks), N.host_ptr(v), n,
                                  b'model_7.bin', N.host_ptr(out), out.size)
    assert written > 0
    got = out[:written].tobytes().decode('utf8').splitlines()
    assert len(got) == n
    for i, line in enumerate(got):
        assert line == 'topic-é Q0 entity/007 {0} {1} model_7.bin'.format(i + 1, float(v[i])), (i, v[i], line)
    # too small an output buffer is an error, not an overrun
    assert lib.sert_format_run(N.host_ptr(s_blob), N.host_ptr(s_off), N.host_ptr(o_blob), N.host_ptr(o_off),
                               N.host_ptr(zeros), N.host_ptr(zeros), N.host_ptr(ranks), N.host_ptr(v), n,
                               b'm', N.host_ptr(out), 100) == -2

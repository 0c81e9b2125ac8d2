"""ctypes binding of libsert_b200.so (the C-ABI declared in include/sert_b200.h).

There is no CPU fallback: importing this module without the built library, or creating a
model without a CUDA device, raises.  PyTorch is used by callers only to allocate HBM and
to move arrays; every computation goes through the symbols bound here.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SERT_B200_LIB: another build of the same library (A/B measurements of two kernel variants in one GPU session)
LIB_PATH = os.environ.get('SERT_B200_LIB') or os.path.join(_HERE, 'libsert_b200.so')

KIND_LOGLINEAR, KIND_VECTORSPACE = 0, 1
SPLIT_TRAIN, SPLIT_VALIDATE = 0, 1
PARAM_WORD_REPR, PARAM_DENSE_W, PARAM_DENSE_B, PARAM_ENTITY_REPR = 0, 1, 2, 3
STATE_PARAM, STATE_S1, STATE_S2 = 0, 1, 2

c_void_p, c_int, c_int32, c_int64, c_size_t, c_float = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float)


class SertConfig(ctypes.Structure):
    _fields_ = [
        ('kind', c_int32), ('batch', c_int32), ('window', c_int32), ('num_negatives', c_int32),
        ('vocab', c_int64), ('entities', c_int64),
        ('word_dim', c_int32), ('entity_dim', c_int32),
        ('lambda_', c_float), ('loss_slots', c_int32),
        ('seed', ctypes.c_uint64),
        ('inference_only', c_int32), ('dtype_mode', c_int32), ('reserved1', c_int64),
    ]


# include/sert_b200.h sert_exchange_fn: int (*)(void *ctx, int32_t op, float *buf_dev, size_t count)
EXCHANGE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int32, c_void_p, c_size_t)

# name -> (restype, argtypes); must list every symbol declared in include/sert_b200.h
SIGNATURES = {
    'sert_abi_version': (c_int, []),
    'sert_last_error': (ctypes.c_char_p, []),
    'sert_launch_count': (ctypes.c_uint64, []),
    'sert_model_arena_bytes': (c_int, [ctypes.POINTER(SertConfig), ctypes.POINTER(c_size_t)]),
    'sert_model_create': (c_int, [ctypes.POINTER(SertConfig), c_void_p, c_size_t, c_void_p,
                                  ctypes.POINTER(c_void_p)]),
    'sert_model_destroy': (c_int, [c_void_p]),
    'sert_model_set_tensor': (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t]),
    'sert_model_get_tensor': (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t]),
    'sert_model_set_step': (c_int, [c_void_p, c_int64]),
    'sert_model_get_step': (c_int, [c_void_p, ctypes.POINTER(c_int64)]),
    'sert_model_get_sampler': (c_int, [c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    'sert_model_set_sampler': (c_int, [c_void_p, ctypes.c_uint64, ctypes.c_uint64]),
    'sert_model_set_entity_shard': (c_int, [c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    'sert_comm_unique_id': (c_int, [c_void_p, c_size_t]),
    'sert_comm_init': (c_int, [c_int32, c_int32, c_void_p, ctypes.POINTER(c_void_p)]),
    'sert_comm_destroy': (c_int, [c_void_p]),
    'sert_comm_info': (c_int, [c_void_p, ctypes.POINTER(c_int32), ctypes.POINTER(c_int32), ctypes.POINTER(c_int32),
                               ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    'sert_model_set_entity_shard_comm': (c_int, [c_void_p, c_void_p, c_int64, c_int64]),
    'sert_model_set_table_shard_comm': (c_int, [c_void_p, c_void_p, c_int32]),
    'sert_model_gather_table_state': (c_int, [c_void_p]),
    'sert_table_shard_plan': (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sert_model_table_shard_info': (c_int, [c_void_p, ctypes.POINTER(c_int32), ctypes.POINTER(c_int64),
                                            ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    'sert_scorer_set_comm': (c_int, [c_void_p, c_void_p]),
    'sert_model_profile': (c_int, [c_void_p, c_int]),
    'sert_model_set_fused': (c_int, [c_void_p, c_int]),
    'sert_model_set_overlap': (c_int, [c_void_p, c_int]),
    'sert_train_batch_host_async': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sert_train_host_wait': (c_int, [c_void_p, c_int64, c_void_p]),
    'sert_model_set_hot_words': (c_int, [c_void_p, c_void_p, c_int32]),
    'sert_model_set_tensor_cores': (c_int, [c_void_p, c_int]),
    'sert_debug_gemm_tc_bench': (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float)]),
    'sert_model_profile_read': (c_int, [c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int64),
                                        ctypes.POINTER(ctypes.c_double)]),
    'sert_model_attach_dataset': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    'sert_train_batches': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int32]),
    'sert_eval_batches': (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int32]),
    'sert_losses_fetch': (c_int, [c_void_p, c_int32, c_int64, c_void_p]),
    'sert_format_run': (c_int64, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, ctypes.c_char_p, c_void_p, c_int64]),
    'sert_train_batch_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    'sert_vs_forward_host': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sert_ll_forward_host': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    'sert_predict_loglinear': (c_int, [c_void_p, c_void_p, c_int32, c_void_p]),
    'sert_project_queries': (c_int, [c_void_p, c_void_p, c_int32, c_void_p]),
    'sert_ll_rank_queries': (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    'sert_ll_rank_distributions': (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    'sert_scorer_arena_bytes': (c_int, [c_int64, c_int32, c_int32, c_int32, ctypes.POINTER(c_size_t)]),
    'sert_scorer_create': (c_int, [c_void_p, c_int64, c_int32, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                   c_size_t, c_void_p, ctypes.POINTER(c_void_p)]),
    'sert_scorer_destroy': (c_int, [c_void_p]),
    'sert_scorer_set_mode': (c_int, [c_void_p, c_int32]),
    'sert_scorer_plan': (c_int, [c_void_p, c_int32, ctypes.POINTER(c_int32), ctypes.POINTER(c_int32),
                                 ctypes.POINTER(c_int32), ctypes.POINTER(c_int64), ctypes.POINTER(ctypes.c_double)]),
    'sert_scorer_stats': (c_int, [c_void_p, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    'sert_scorer_topk_host': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'sert_scorer_scores_host': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    'sert_scorer_topk_dev': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'sert_debug_gemm_tc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'sert_debug_gemm_tc_bn': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'sert_topk_merge_dev': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Loads the shared library once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libsert_b200.so is missing (%s): build it with `python -m sert_b200.build`; '
            'there is no CPU fallback for the SERT hot path.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sert_abi_version() != 1:
        raise RuntimeError('libsert_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().sert_last_error()
        raise RuntimeError(msg.decode('utf8', 'replace') if msg else 'libsert_b200 call failed')


def host_ptr(a):
    """Pointer to a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(c_void_p)


def dev_ptr(t):
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return c_void_p(t.data_ptr())


def launch_count():
    return int(load().sert_launch_count())

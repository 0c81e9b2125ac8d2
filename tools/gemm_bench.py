"""Raw throughput of the tcgen05 GEMM kernel: python tools/gemm_bench.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: E402,F401  (initialises CUDA the same way the product does)
from sert_b200 import _native as N  # noqa: E402

lib = N.load()
fn = lib.sert_debug_gemm_tc_bench
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_float)]
for (m, n, kt) in [(10112, 65536, 768), (10112, 65536, 384), (10112, 65536, 128), (8192, 8192, 8192)]:
    for mode in (0, 1):
        ms = ctypes.c_float(0)
        N.check(fn(m, n, kt, 5, mode, ctypes.byref(ms)))
        print('m=%d n=%d kt=%d mode=%s: %.3f ms  %.1f TFLOP/s' % (
            m, n, kt, 'store' if mode == 0 else 'topk', ms.value, 2.0 * m * n * kt / ms.value / 1e9))

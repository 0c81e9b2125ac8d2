#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_loglinear.py tests/test_gpu_loglinear_sharded.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_scoring.py -q 2>&1 | tail -3
run() { env "$@" timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "ms/step" | cut -c1-150 | tr '\n' ' '; echo " [$*]"; }
run A=transposed_stores
run SERT_B200_LIB=$PWD/tools/ab/libsert_b200_ts0.so
run A=transposed_stores
run SERT_B200_LIB=$PWD/tools/ab/libsert_b200_ts0.so

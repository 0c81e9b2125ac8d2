#!/bin/bash
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -x 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_loglinear.py tests/test_gpu_loglinear_sharded.py tests/test_gpu_golden.py tests/test_gpu_full_size.py -q 2>&1 | tail -4
run() { env "$@" timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "ms/step" | cut -c1-120 | tr '\n' ' '; echo " [$*]"; }
run SERT_LL_BN=1
run SERT_LL_BN=0
run SERT_GEMM_NFAST=0
run SERT_LL_BN=1

#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_loglinear.py tests/test_gpu_loglinear_sharded.py tests/test_gpu_golden.py -q 2>&1 | tail -3
run() { env "$@" timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "ms/step" | cut -c1-150 | tr '\n' ' '; echo " [$*]"; }
run A=equal_tiles
run SERT_GEMM_EQUAL_TILES=0
run A=equal_tiles
run SERT_GEMM_EQUAL_TILES=0 SERT_GEMM_NFAST=0

#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_loglinear.py tests/test_gpu_loglinear_sharded.py tests/test_gpu_golden.py -q 2>&1 | tail -3
run() { env "$@" timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "ms/step" | cut -c1-110 | tr '\n' ' '; echo " [$*]"; }
run A=new_epilogue_cg2
run SERT_B200_LIB=$PWD/tools/ab/libsert_b200_prev.so
run SERT_GEMM_CG2=0
run A=new_epilogue_cg2
run SERT_B200_LIB=$PWD/tools/ab/libsert_b200_prev.so

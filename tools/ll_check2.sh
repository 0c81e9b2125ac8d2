#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_loglinear.py tests/test_gpu_loglinear_sharded.py tests/test_gpu_golden.py tests/test_gpu_full_size.py tests/test_gpu_cli.py -q 2>&1 | tail -5
for T in 2 2; do
  SERT_LL_TERMS=$T timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "arena|ms/step" | tr '\n' ' '; echo " [terms $T]"
done
bash tools/ll_profile.sh 40 24

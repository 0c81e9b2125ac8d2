#!/bin/bash
# two-SM MMA variant (SERT_GEMM_CG2=1): GEMM tests, then the configs[4] step A/B
SERT_GEMM_CG2=1 timeout 150 python -m pytest tests/test_gpu_gemm_tc.py -q -x -k "cluster or pair or split3" 2>&1 | tail -6; echo "rc=$?"

#!/bin/bash
# compute-sanitizer memcheck over the kernels added in the second half of round 2 (small shapes)
S="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0"
timeout 900 $S python -m pytest tests/test_gpu_gemm_tc.py -q -x -k "pair or n_major or identity" 2>&1 | tail -4; echo "gemm rc=$?"
timeout 900 $S python -m pytest tests/test_gpu_loglinear.py -q -x -k "tensor_core or fused or forward" 2>&1 | tail -4; echo "ll rc=$?"
timeout 900 $S python -m pytest tests/test_gpu_table_shards.py -q -x -k "world_of_one" 2>&1 | tail -4; echo "shards rc=$?"

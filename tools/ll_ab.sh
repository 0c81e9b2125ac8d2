#!/bin/bash
run() { env "$@" timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "ms/step" | cut -c1-110 | tr '\n' ' '; echo " [$*]"; }
run A=default
run SERT_GEMM_CLUSTER=4
run A=default
run SERT_GEMM_CLUSTER=0

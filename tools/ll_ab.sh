#!/bin/bash
# A/B of the configs[4] log-linear step on one box: bash tools/ll_ab.sh "VAR=value" ["VAR2=value" ...]
# Every argument is one environment assignment timed against the default, alternating.
run() { env "$@" timeout 300 python tools/loglinear_bench.py 500000 200000 300 1024 6 1 2>&1 | grep -E "ms/step" | cut -c1-110 | tr '\n' ' '; echo " [$*]"; }
for v in "$@"; do
  run A=default
  run "$v"
done
run A=default

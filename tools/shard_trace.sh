#!/bin/bash
# Phase trace of the sharded step: bash tools/shard_trace.sh N [extra env assignments]  (run under gpurun --gpus N)
N=${1:-2}; shift
O=gpurun_out
env SERT_TABLE_SHARD_TRACE=1 "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29655 tools/table_shard_bench.py 100 > $O/trace_n$N.json 2> $O/trace_n$N.err; echo "rc=$?"
grep "table shard trace" $O/trace_n$N.err
grep -E "ms_per_step" $O/trace_n$N.json

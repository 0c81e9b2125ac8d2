#!/bin/bash
# Table-shard tests + step bench on an N-GPU box: bash tools/shard_check.sh N (run under gpurun --gpus N)
N=${1:-2}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_table_shards.py tests/test_gpu_vectorspace.py -q -x > $O/pytest_shards.log 2>&1; echo "pytest rc=$?"
tail -n 15 $O/pytest_shards.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
  tools/table_shard_bench.py 100 > $O/table_shard_bench_n$N.json 2> $O/table_shard_bench_n$N.err; echo "bench rc=$?"
tail -n 5 $O/table_shard_bench_n$N.err
grep -E "ms_per_step|speedup|nvlink_bytes_sent_per_step|identical|loss_rel" $O/table_shard_bench_n$N.json

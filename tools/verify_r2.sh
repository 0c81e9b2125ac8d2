#!/bin/bash
# Whole-state check on a 2-GPU box: every GPU test (multi-GPU legs included), smoke, bench at N=1 and N=2 and the
# table-shard bench (run under gpurun --gpus 2; outputs land in gpurun_out/).
O=gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"
tail -n 5 $O/pytest_gpu.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$? t=$((SECONDS-T0))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$? t=$((SECONDS-T0))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$? t=$((SECONDS-T0))"
head -c 1500 $O/bench_n1.json; echo; head -c 600 $O/bench_n2.json; echo

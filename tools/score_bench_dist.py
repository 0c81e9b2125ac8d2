"""Sharded scoring micro-benchmark: torchrun ... tools/score_bench_dist.py Q E d k reps"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from sert_b200.scoring import ShardedScorer, shard_bounds  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
Q, E, d, k, reps = (int(v) for v in (sys.argv[1:6] + ['10000', '50000', '128', '100', '10'][len(sys.argv) - 1:]))
b, e = shard_bounds(E, world, rank)
rng = np.random.default_rng(1 + rank)
ent = rng.standard_normal((e - b, d)).astype(np.float32)
ent /= np.linalg.norm(ent, axis=1)[:, None]
qs = np.random.default_rng(7).standard_normal((Q, d)).astype(np.float32)
qs /= np.linalg.norm(qs, axis=1)[:, None]
sc = ShardedScorer(ent, E, is_shard=True, max_queries=Q, max_k=128)
qd = torch.from_numpy(qs).cuda()
for _ in range(3):
    sc.topk_dev(qd, k)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sc.topk_dev(qd, k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
if rank == 0:
    print('world=%d Q=%d E=%d d=%d k=%d: %.3f ms per call' % (world, Q, E, d, k, ms))
dist.destroy_process_group()

"""Scoring-only micro-benchmark: python tools/score_bench.py [Q E d k reps mode]."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from sert_b200.scoring import EntityScorer  # noqa: E402

Q, E, d, k, reps = (int(v) for v in (sys.argv[1:6] + ['10000', '50000', '128', '100', '5'][len(sys.argv) - 1:]))
mode = sys.argv[6] if len(sys.argv) > 6 else 'tensor'
rng = np.random.default_rng(1)
ent = rng.standard_normal((E, d)).astype(np.float32)
ent /= np.linalg.norm(ent, axis=1)[:, None]
qs = rng.standard_normal((Q, d)).astype(np.float32)
qs /= np.linalg.norm(qs, axis=1)[:, None]
sc = EntityScorer(ent, max_queries=Q, max_k=128)
sc.set_mode(mode)
qd = torch.from_numpy(qs).cuda()
for _ in range(2):
    sc.topk_dev(qd, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sc.topk_dev(qd, k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('stats (seeded, fallback):', sc.stats())
print('mode=%s Q=%d E=%d d=%d k=%d: %.3f ms  %.3e entities/s  %.1f TFLOP/s (x1 flops)' %
      (mode, Q, E, d, k, ms, Q * E / ms * 1e3, 2.0 * Q * E * d / ms / 1e9))

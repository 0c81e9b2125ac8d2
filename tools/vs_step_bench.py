"""Vector-space training step timing at BASELINE.json configs[1] for each fused-kernel variant:
python tools/vs_step_bench.py [steps] [variants, e.g. 1,2,0]
(1 = tile kernel, 2 = warp kernel, 0 = per-stage kernels; see sert_model_set_fused in include/sert_b200.h)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402
from sert_b200 import _native as N, models  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else '1,2').split(',')]
cfg = bench.CFG2
warm = 5
nb = steps + warm
p = bench.make_problem(0, nb)
neg_dev = torch.from_numpy(p['neg']).cuda()
order = np.arange(nb, dtype=np.int64)
for variant in variants:
    model = models.VectorSpaceLanguageModel(
        batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'],
        representations_init=p['R'], entity_representations_init=p['Eemb'], regularization_lambda=cfg['lam'],
        training_set=p['train'], validation_set=p['val'], dense_init=(p['Wp'], p['bp']), loss_slots=max(1024, nb),
        optimizer_state_dtype='bfloat16' if os.environ.get('SERT_BENCH_BF16_STATE') else 'float32')
    nat = model._native
    N.check(nat.lib.sert_model_set_fused(nat.handle, variant))
    best = None
    for rep in range(3):
        N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[:warm]), warm, N.c_void_p(neg_dev.data_ptr()), 0))
        torch.cuda.synchronize()
        l0 = N.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[warm:]), steps,
                                           N.c_void_p(neg_dev.data_ptr() + warm * cfg['B'] * cfg['k'] * 4), warm))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = ms if best is None else min(best, ms)
        launches = (N.launch_count() - l0) / steps
    losses = np.empty(nb, np.float32)
    N.check(nat.lib.sert_losses_fetch(nat.handle, 0, nb, N.host_ptr(losses)))
    print('fused=%d: %.4f ms/step (best of 3)  %.3e pairs/s  %.1f launches/step  loss[-1]=%.6f' % (
        variant, best, cfg['B'] / best * 1e3, launches, losses[-1]), flush=True)
    nat.close()
    del model

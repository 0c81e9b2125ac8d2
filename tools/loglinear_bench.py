"""Log-linear training step timing: python tools/loglinear_bench.py V E dw B steps tensor(0|1)
Default = BASELINE.json configs[4] on one GPU: V=500k E=200k d=300 B=1024 window=10 (full-softmax stress)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from sert_b200 import _native as N, models, synth  # noqa: E402

args = sys.argv[1:] + ['500000', '200000', '300', '1024', '5', '1'][len(sys.argv) - 1:]
V, E, dw, B, steps, tensor = (int(v) for v in args[:6])
W = 10
nb = steps + 2
rng = np.random.default_rng(20160816 + 5)
t0 = time.time()
train, val = synth.loglinear_corpus(20160821, V, E, W, B * nb, B)
R, Wd, bd = synth.glorot(rng, (V, dw)), synth.glorot(rng, (dw, E)), np.zeros(E, np.float32)
print('data generated in %.1fs' % (time.time() - t0), flush=True)
model = models.LanguageModel(batch_size=B, window_size=W, representations_init=R, output_layer_size=E,
                             regularization_lambda=0.01, training_set=train, validation_set=val, dense_init=(Wd, bd),
                             loss_slots=64)
nat = model._native
print('arena %.1f GB' % (nat.arena.numel() / 1e9), flush=True)
N.check(nat.lib.sert_model_set_tensor_cores(nat.handle, tensor))
order = np.arange(nb, dtype=np.int64)
N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[:2]), 2, None, 0))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[2:]), steps, None, 2))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
losses = np.empty(nb, np.float32)
N.check(nat.lib.sert_losses_fetch(nat.handle, 0, nb, N.host_ptr(losses)))
flops = 6.0 * B * W * dw * E
print('V=%d E=%d dw=%d B=%d tensor=%d: %.2f ms/step  %.3e pairs/s  %.1f TFLOP/s (algorithmic fp32 flops)  losses %s' % (
    V, E, dw, B, tensor, ms, B / ms * 1e3, flops / ms / 1e9, losses[:4]))

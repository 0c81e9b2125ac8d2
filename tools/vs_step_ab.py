"""Vector-space step timing at two shapes (A/B through SERT_B200_LIB): python tools/vs_step_ab.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from sert_b200 import _native as N, models, synth  # noqa: E402

for (dw, de, W) in [(128, 128, 10), (300, 128, 4)]:
    V, E, B, k, nb = 100000, 50000, 4096, 10, 30
    rng = np.random.default_rng(3)
    train, val = synth.vectorspace_corpus(3, V, E, W, B * nb, B)
    model = models.VectorSpaceLanguageModel(
        batch_size=B, window_size=W, num_negative_samples=k, representations_init=synth.glorot(rng, (V, dw)),
        entity_representations_init=synth.glorot(rng, (E, de)), regularization_lambda=0.01, training_set=train,
        validation_set=val, loss_slots=1024)
    nat = model._native
    neg = torch.from_numpy(rng.integers(0, E, size=(nb, B, k)).astype(np.int32)).cuda()
    order = np.arange(nb, dtype=np.int64)
    run = lambda: N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order), nb, N.c_void_p(neg.data_ptr()), 0))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print('dw=%d de=%d W=%d: %.4f ms/step' % (dw, de, W, e0.elapsed_time(e1) / (20 * nb)))
    nat.close()

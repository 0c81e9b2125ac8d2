"""Table-sharded vector-space step next to the single-GPU step, BASELINE.json configs[1] sizes:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
      tools/table_shard_bench.py [steps]
Prints bench.py's table_shards object (rank 0)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from sert_b200 import _native as N, models
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    warm, cfg = 5, bench.CFG2
    n_batches = steps + warm
    p = bench.make_problem(0, n_batches)
    model = models.VectorSpaceLanguageModel(
        batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'],
        representations_init=p['R'], entity_representations_init=p['Eemb'], regularization_lambda=cfg['lam'],
        training_set=p['train'], validation_set=p['val'], dense_init=(p['Wp'], p['bp']), loss_slots=max(1024, n_batches))
    nat, lib = model._native, model._native.lib
    neg_dev = torch.from_numpy(p['neg']).cuda()
    order = np.arange(n_batches, dtype=np.int64)

    def train(lo, hi):
        N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order[lo:hi]), hi - lo,
                                       N.c_void_p(neg_dev.data_ptr() + lo * cfg['B'] * cfg['k'] * 4), lo))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    train(0, n_batches)
    first = np.empty(n_batches, np.float32)
    N.check(lib.sert_losses_fetch(nat.handle, 0, n_batches, N.host_ptr(first)))
    repeats = 10
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(repeats):
        train(warm, n_batches)
    ev1.record()
    barrier()
    single_ms = ev0.elapsed_time(ev1) / (repeats * steps)
    nat.close()
    del model
    out = bench.run_table_shards(cfg, steps, warm, repeats, single_ms, first, rank, world, barrier)
    if rank == 0:
        print(json.dumps(out, indent=1))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

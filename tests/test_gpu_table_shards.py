"""Table-sharded vector-space training (include/sert_b200.h: sert_model_set_table_shard_comm).
Mode 3 ("instance shards") additionally splits the batch's instances over the ranks: gradient rows are added into
their owners' arenas over NVLink and the step carries no NCCL call (two barrier kernels over peer memory).

One model at the global batch: every rank computes the step's gradient, each rank streams the Adam update over its
own piece of the two tables, the new parameters reach the other ranks either by grouped NCCL broadcasts (mode 1) or
by the update kernels' own NVLink stores into the next of two parameter buffers (mode 2).  Checked against the CPU
oracle exactly like the single-GPU step.  World size 1 runs on any box and covers the out-of-place update and the
buffer swap; the multi-process leg needs >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_model(p, lam, **kw):
    from sert_b200 import models
    return models.VectorSpaceLanguageModel(
        batch_size=p['B'], window_size=p['W'], num_negative_samples=p['k'],
        representations_init=p['R'], entity_representations_init=p['Eemb'],
        regularization_lambda=lam, training_set=p['train'], validation_set=p['val'],
        dense_init=(p['Wp'], p['bp']), **kw)


MODES = {False: 1, True: 2, 'instances': 3}


@pytest.mark.parametrize('peer_stores,fused,overlap', [(ps, f, o) for ps in (True, False) for f, o in ((1, 1), (1, 0), (0, 1))] +
                         [('instances', 1, 1)])
def test_world_of_one_matches_oracle(peer_stores, fused, overlap):
    from sert_b200 import _native as N
    from sert_b200.comm import Communicator
    p = H.vs_problem(23, V=900, E=300, dw=128, de=128, W=6, B=96, k=8, n_batches=5, weights=True)
    lam = 0.01
    model = make_model(p, lam, table_shard=Communicator(0, 1, None), table_shard_peer_stores=peer_stores)
    N.check(model._native.lib.sert_model_set_fused(model._native.handle, fused))
    N.check(model._native.lib.sert_model_set_overlap(model._native.handle, overlap))
    mode, b, e, n = model.table_shard_info()
    assert mode == MODES[peer_stores] and (b, e) == (0, n) and n == (900 + 300) * 128
    oracle = H.vs_oracle(p, lam)
    for j, bi in enumerate([3, 0, 4, 1, 2]):
        H.close(model.train_fn(bi, p['neg'][j]), oracle.train_batch(bi, p['neg'][j]), what='train loss step %d' % j)
    R, Eemb = model.get_representations()
    Wp, bp = model.get_dense()
    H.close(R, oracle.R, rtol=2e-4, atol_scale=5e-5, what='R')     # Adam turns ulp-level gradient noise of near-zero elements into ~1e-6 steps
    H.close(Eemb, oracle.Eemb, rtol=2e-4, atol_scale=5e-5, what='Eemb')
    H.close(Wp, oracle.Wp, rtol=2e-4, what='Wp')
    H.close(bp, oracle.bp, rtol=2e-4, atol_scale=1e-4, what='bp')
    H.close(model.test_fn(2, p['neg'][2]), oracle.eval_batch('train', 2, p['neg'][2]), rtol=2e-4,
            what='eval loss after training')


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from tests import helpers as H
from sert_b200 import models, _native as N
from sert_b200.comm import Communicator
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
comm = Communicator.from_torch_distributed()
lam = 0.01
for case, dims in enumerate([dict(V=1500, E=700, dw=128, de=128, W=8, B=256, k=10),      # tile kernel, hot rows
                             dict(V=801, E=333, dw=64, de=48, W=5, B=128, k=6)]):         # warp kernel, ragged pieces
    p = H.vs_problem(29 + case, n_batches=6, weights=True, **dims)
    oracle = H.vs_oracle(p, lam)
    order = [3, 0, 5, 1, 2, 4]
    ref = [oracle.train_batch(b, p['neg'][j]) for j, b in enumerate(order)]
    kw = dict(batch_size=p['B'], window_size=p['W'], num_negative_samples=p['k'], representations_init=p['R'],
              entity_representations_init=p['Eemb'], regularization_lambda=lam, training_set=p['train'],
              validation_set=p['val'], dense_init=(p['Wp'], p['bp']))
    # instance shards (each rank runs its own slice of the batch) serve the tile kernel's shapes
    for peer_stores in (('instances', True, False) if case == 0 else (True, False)):
        # ranks other than 0 start from DIFFERENT parameters: the attach call must bring rank 0's everywhere
        init = dict(kw)
        if rank != 0:
            init['representations_init'] = p['R'] * 0.5
        model = models.VectorSpaceLanguageModel(table_shard=comm, table_shard_peer_stores=peer_stores, **init)
        mode, b, e, n = model.table_shard_info()
        assert mode == {False: 1, True: 2, 'instances': 3}[peer_stores] and 0 <= b < e <= n, (mode, b, e, n)
        spans = [None] * world
        dist.all_gather_object(spans, (b, e))
        assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1)), spans
        # four steps in ONE call (look-ahead: with peer stores only the rows the next batch reads are sent; the call's
        # last step sends everything), then single-step calls
        got = list(model._native.run_batches('train', order[:4], np.stack([p['neg'][j] for j in range(4)])))
        got += [model.train_fn(bi, p['neg'][j]) for j, bi in enumerate(order) if j >= 4]
        H.close(got, ref, what='train losses (rank %%d, peer_stores=%%s)' %% (rank, peer_stores))
        R, Eemb = model.get_representations()
        Wp, bp = model.get_dense()
        H.close(R, oracle.R, rtol=2e-4, atol_scale=5e-5, what='R')     # Adam turns ulp-level gradient noise of near-zero elements into ~1e-6 steps
        H.close(Eemb, oracle.Eemb, rtol=2e-4, atol_scale=5e-5, what='Eemb')
        H.close(Wp, oracle.Wp, rtol=2e-4, what='Wp')
        H.close(bp, oracle.bp, rtol=2e-4, atol_scale=1e-4, what='bp')
        H.close(model.test_fn(2, p['neg'][2]), oracle.eval_batch('train', 2, p['neg'][2]), rtol=2e-4, what='eval loss')
        # ONE model: every rank holds bit-identical parameters
        mine = torch.from_numpy(np.concatenate([R.ravel(), Eemb.ravel(), Wp.ravel(), bp.ravel()])).cuda()
        ref0 = mine.clone()
        dist.broadcast(ref0, 0)
        assert torch.equal(mine, ref0), 'rank %%d differs from rank 0' %% rank
        # checkpoints carry the whole optimiser state on every rank
        ckpt = model.get_checkpoint()
        H.close(ckpt['entity_representations/state1'], oracle.state['Eemb'][0], rtol=2e-4, atol_scale=1e-4, what='Adam m (Eemb)')
        H.close(ckpt['representations/state2'], oracle.state['R'][1], rtol=2e-4, atol_scale=1e-4, what='Adam v (R)')
        H.close(ckpt['dense_w/state1'], oracle.state['Wp'][0], rtol=2e-4, atol_scale=1e-4, what='Adam m (Wp)')
        # negatives drawn on the device (one step ahead under look-ahead): every rank must draw the same ones
        sampled = model._native.run_batches('train', [1, 4, 0, 2, 5], None)
        assert np.all(np.isfinite(sampled))
        R2, E2 = model.get_representations()
        mine = torch.from_numpy(np.concatenate([sampled, R2.ravel(), E2.ravel()])).cuda()
        ref0 = mine.clone()
        dist.broadcast(ref0, 0)
        assert torch.equal(mine, ref0), 'rank %%d differs from rank 0 after device-sampled steps' %% rank
        model._native.close()
        dist.barrier()
        oracle = H.vs_oracle(p, lam)
        ref = [oracle.train_batch(b, p['neg'][j]) for j, b in enumerate(order)]
info = comm.info()
if rank == 0:
    print('TABLE_SHARDS_OK world=%%d collectives=%%d' %% (world, info['collectives']))
dist.destroy_process_group()
'''


def test_table_shards_multi_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', '29633', str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=420)
    text = out.stdout.decode()
    assert out.returncode == 0 and 'TABLE_SHARDS_OK' in text, text[-4000:]

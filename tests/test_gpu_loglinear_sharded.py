"""Entity-sharded (column-parallel) log-linear training, SURVEY.md 8(e).

Single-GPU leg: all shards live on cuda:0, one host thread per shard, collectives by sharding.LocalExchange -- the
kernels, the exchange points and the loss completion are exactly those of the multi-GPU run, so the sharded
arithmetic is checked against the CPU oracle on any box.  Multi-GPU leg (>= 2 GPUs): the same through NCCL."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_shards(world, body):
    """Runs body(rank, exchange) on `world` threads; returns the per-rank results, re-raising the first error."""
    from sert_b200 import sharding
    hub = sharding.LocalExchange(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = body(r, hub.shard(r))
        except BaseException as e:
            err[r] = e
            hub.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out


def make_model(p, lam, exchange):
    from sert_b200 import models
    return models.LanguageModel(
        batch_size=p['B'], window_size=p['W'], representations_init=p['R'],
        output_layer_size=p['E'], regularization_lambda=lam,
        training_set=p['train'], validation_set=p['val'], dense_init=(p['Wd'], p['bd']), entity_shard=exchange)


@pytest.mark.parametrize('world,gain,dims', [
    (2, 1.0, dict(V=900, E=120, dw=32, W=4, B=64)),
    (3, 25.0, dict(V=900, E=121, dw=32, W=4, B=64)),          # uneven shards, clipped regime
    (4, 1.0, dict(V=5000, E=200, dw=64, W=10, B=64)),         # BASELINE.json configs[0] shapes
])
def test_sharded_training_matches_oracle(world, gain, dims):
    p = H.ll_problem(17, n_batches=4, gain=gain, **dims)
    lam = 0.01
    oracle = H.ll_oracle(p, lam)
    ref = dict(e0=oracle.eval_batch('train', 1), v0=oracle.eval_batch('val', 0))
    order = [2, 0, 3, 1]
    ref['train'] = [oracle.train_batch(b) for b in order]
    ref['e1'] = oracle.eval_batch('train', 3)

    def body(rank, exchange):
        model = make_model(p, lam, exchange)
        got = dict(e0=model.test_fn(1), v0=model.validate_fn(0))
        got['train'] = [model.train_fn(b) for b in order]
        got['e1'] = model.test_fn(3)
        got['dense'] = model.get_dense()
        got['R'] = model.get_representations()
        got['ckpt'] = model.get_checkpoint()
        got['calls'] = exchange.calls
        return got

    results = run_shards(world, body)
    for rank, got in enumerate(results):
        H.close(got['e0'], ref['e0'], what='initial eval loss (rank %d)' % rank)
        H.close(got['v0'], ref['v0'], what='initial validation loss')
        H.close(got['train'], ref['train'], what='train losses')
        H.close(got['e1'], ref['e1'], rtol=2e-4, what='eval loss after training')
        H.close(got['R'], oracle.R, rtol=2e-4, what='R')
        H.close(got['dense'][0], oracle.Wd, rtol=2e-4, what='Wd (gathered)')
        H.close(got['dense'][1], oracle.bd, rtol=2e-4, atol_scale=1e-4, what='bd (gathered)')
        H.close(got['ckpt']['dense_w/state1'], oracle.state['Wd'][0], rtol=5e-4, atol_scale=1e-4, what='accu (Wd)')
        H.close(got['ckpt']['representations/state2'], oracle.state['R'][1], rtol=5e-4, atol_scale=1e-4,
                what='delta (R)')
        # every rank reports the same numbers; the replicated word table agrees up to the order of the float
        # atomics of the scatter-add (the same ulp-level run-to-run variation a single device has)
        assert got['train'][0] == results[0]['train'][0] and got['e0'] == results[0]['e0']
        H.close(got['R'], results[0]['R'], rtol=1e-5, what='replicated R across ranks')
    # 2 eval batches x 2 gathers + 1 loss reduce each, 4 train steps x 5 exchanges + 1 loss reduce each, 1 eval
    assert results[0]['calls'] == 3 * 3 + 4 * 6


def test_sharded_forward_hook_and_checkpoint_roundtrip():
    from oracle import sert_oracle as O
    from sert_b200 import _native as N
    p = H.ll_problem(31, V=600, E=90, dw=16, W=3, B=32, n_batches=2)
    f = O.loglinear_forward(p['R'], p['Wd'], p['bd'], p['train'][0][:32])
    ref_ell = O.loglinear_instance_losses(f['o'], O.dense_rows(p['train'][1], 0, 32))

    def body(rank, exchange):
        model = make_model(p, 0.01, exchange)
        nat = model._native
        B, W, E = nat.cfg.batch, nat.cfg.window, nat.cfg.entities
        z, s, ell = np.empty((B * W, E), np.float32), np.empty((B, E), np.float32), np.empty(B, np.float32)
        nat._check(nat.lib.sert_ll_forward_host(nat.handle, 0, 0, N.host_ptr(z), N.host_ptr(s), N.host_ptr(ell)))
        model.train_fn(0)
        ckpt = model.get_checkpoint()
        other = make_model(p, 0.01, exchange)
        other.set_checkpoint(ckpt)
        return dict(span=model._shard_span, z=z, s=s, ell=ell, a=model.train_fn(1), b=other.train_fn(1))

    for got in run_shards(2, body):
        b, e, _ = got['span']
        H.close(got['z'], f['z'].reshape(-1, 90)[:, b:e], what='local columns of z')
        H.close(got['s'], f['s'][:, b:e], what='local columns of s')
        H.close(got['ell'], ref_ell, what='instance losses completed over the shards')
        H.close(got['a'], got['b'], rtol=1e-6, what='resumed model continues')


def test_sharded_callback_error_surfaces():
    """A Python exception inside the exchange callback becomes the error of the training call."""
    from sert_b200 import sharding
    p = H.ll_problem(5, V=200, E=40, dw=16, W=2, B=16, n_batches=1)

    class Broken(sharding.Exchange):
        def all_gather(self, view, count):
            raise ValueError('exchange exploded')

        def all_reduce_sum(self, view):
            raise ValueError('exchange exploded')

    model = make_model(p, 0.01, Broken())
    with pytest.raises(ValueError, match='exchange exploded'):
        model.train_fn(0)


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from tests import helpers as H
from sert_b200 import models, sharding
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
p = H.ll_problem(41, V=3000, E=1003, dw=64, W=5, B=128, n_batches=4)
kw = dict(batch_size=p['B'], window_size=p['W'], representations_init=p['R'], output_layer_size=p['E'],
          regularization_lambda=0.01, training_set=p['train'], validation_set=p['val'],
          dense_init=(p['Wd'], p['bd']))
sharded = models.LanguageModel(entity_shard=sharding.CommExchange(), **kw)   # NCCL inside the library
single = models.LanguageModel(**kw)
order = [3, 1, 0, 2]
n, mean = sharded.train(order=order)
n1, mean1 = single.train(order=order)
assert n == n1 == 4
H.close(mean, mean1, what='epoch mean loss')
H.close(sharded.validation_error()[0], single.validation_error()[0], rtol=2e-4, what='validation error')
H.close(sharded.get_representations(), single.get_representations(), rtol=2e-4, what='R')
H.close(sharded.get_dense()[0], single.get_dense()[0], rtol=2e-4, what='Wd')
info = sharded._native.exchange.comm.info()
assert info['collectives'] > 0 and info['world'] == world, info
# the callback form (torch.distributed collectives ordered on the model's stream) gives the same numbers
legacy = models.LanguageModel(entity_shard=sharding.DistExchange(), **kw)
n2, mean2 = legacy.train(order=order)
H.close(mean2, mean, rtol=1e-5, what='callback exchange vs library exchange')
dist.barrier()
if rank == 0:
    print('LL_SHARDED_OK world=%%d nccl=%%d collectives=%%d' %% (world, info['nccl_version'], info['collectives']))
dist.destroy_process_group()
'''


def test_sharded_training_nccl(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', '29612', str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    text = out.stdout.decode()
    assert out.returncode == 0 and 'LL_SHARDED_OK' in text, text[-3000:]

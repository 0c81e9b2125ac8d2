"""Row-sharded scoring over NCCL (needs >= 2 GPUs; skipped otherwise): the merged lists equal single-GPU lists."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from sert_b200.scoring import EntityScorer, ShardedScorer
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
rng = np.random.default_rng(5)
E = rng.standard_normal((30011, 96)).astype(np.float32)
E /= np.linalg.norm(E, axis=1)[:, None]
E[20000] = E[17]                                   # exact tie across shards: lower row id must win
q = rng.standard_normal((123, 96)).astype(np.float32)
q /= np.linalg.norm(q, axis=1)[:, None]
k = 100
sharded = ShardedScorer(E, E.shape[0], max_queries=128, max_k=128)
idx, score = sharded.topk(q, k)
full = EntityScorer(E, max_queries=128, max_k=128)
ref_idx, ref_score = full.topk(q, k)
assert (idx == ref_idx).all(), (rank, np.argwhere(idx != ref_idx)[:5])
np.testing.assert_array_equal(score, ref_score)
info = sharded.comm.info()
assert info['collectives'] >= 1 and info['world'] == world, info       # the all-gather ran inside the library
qd = torch.from_numpy(q).cuda()
di, ds = sharded.topk_dev(qd, k)
torch.cuda.synchronize()
assert (di.cpu().numpy() == ref_idx).all()
np.testing.assert_array_equal(ds.cpu().numpy(), ref_score)
dist.barrier()
if rank == 0:
    print('SHARDED_OK world=%%d nccl=%%d' %% (world, info['nccl_version']))
dist.destroy_process_group()
'''


def test_sharded_scoring_matches_single_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', '29611', str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    text = out.stdout.decode()
    assert out.returncode == 0 and 'SHARDED_OK' in text, text[-3000:]


WORKER_SKEW = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from sert_b200.scoring import EntityScorer, ShardedScorer
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
rng = np.random.default_rng(9)
E = rng.standard_normal((40000, 64)).astype(np.float32)
E /= np.linalg.norm(E, axis=1)[:, None] * 2.0
q = rng.standard_normal((50, 64)).astype(np.float32)
q /= np.linalg.norm(q, axis=1)[:, None]
# every row of query 0's top 100 sits in the FIRST shard: the short per-shard lists of the first attempt cannot hold
# them, the merge must notice and the call must repeat with full-length lists
E[100:400] = (q[0][None, :] * np.linspace(0.6, 0.9, 300, dtype=np.float32)[:, None])
k = 100
sharded = ShardedScorer(E, E.shape[0], max_queries=64, max_k=128)
idx, score = sharded.topk(q, k)
full = EntityScorer(E, max_queries=64, max_k=128)
ref_idx, ref_score = full.topk(q, k)
assert (idx == ref_idx).all(), (rank, np.argwhere(idx != ref_idx)[:5])
np.testing.assert_array_equal(score, ref_score)
assert set(idx[0].tolist()) <= set(range(100, 400))
dist.barrier()
if rank == 0:
    print('SKEW_OK world=%%d' %% world)
dist.destroy_process_group()
'''


def test_sharded_scoring_skewed_shards_repeat_with_full_lists(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    script = tmp_path / 'worker_skew.py'
    script.write_text(WORKER_SKEW % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', '29613', str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    text = out.stdout.decode()
    assert out.returncode == 0 and 'SKEW_OK' in text, text[-3000:]

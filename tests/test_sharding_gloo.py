"""World-size-2 gloo test (CPU) of the host-side logic of row-sharded scoring: shard bounds, the packing of the
ONE all-gather, rank order of the gathered lists and the merge rule.  The per-shard top-k and the merge are
computed by test-side numpy stand-ins for the device kernels (which need a GPU: tests/test_gpu_scoring.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def local_topk(E, q, k, row_begin):
    s = (q @ E.T).astype(np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(E.shape[0]), s.shape), -s), axis=1)[:, :k]
    idx = np.full((q.shape[0], k), -1, np.int32)
    sc = np.full((q.shape[0], k), -np.inf, np.float32)
    n = order.shape[1]
    idx[:, :n] = order + row_begin
    sc[:, :n] = np.take_along_axis(s, order, axis=1)
    return idx, sc


def merge(g_idx, g_score, k):
    """Same rule as csrc/score.cu::merge_kernel: score descending, ties on lower global row id."""
    world, Q, kk = g_idx.shape
    idx = g_idx.transpose(1, 0, 2).reshape(Q, -1)
    sc = g_score.transpose(1, 0, 2).reshape(Q, -1)
    key_id = np.where(idx >= 0, idx, np.iinfo(np.int32).max)
    order = np.lexsort((key_id, -sc), axis=1)[:, :k]
    return np.take_along_axis(idx, order, axis=1), np.take_along_axis(sc, order, axis=1)


def worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sert_b200.scoring import all_gather_lists, shard_bounds
    rng = np.random.default_rng(11)
    E = rng.standard_normal((1001, 16)).astype(np.float32)
    E[500] = E[3]                                   # an exact tie across shards
    q = rng.standard_normal((7, 16)).astype(np.float32)
    k = 20
    b, e = shard_bounds(E.shape[0], world, rank)
    idx, sc = local_topk(E[b:e], q, k, b)
    g_idx, g_sc = all_gather_lists(torch.from_numpy(idx), torch.from_numpy(sc))
    assert g_idx.shape == (world, 7, k) and g_idx.dtype == torch.int32 and g_sc.dtype == torch.float32
    # rank r's slice must be rank r's list, bit for bit
    np.testing.assert_array_equal(g_idx[rank].numpy(), idx)
    np.testing.assert_array_equal(g_sc[rank].numpy(), sc)
    m_idx, m_sc = merge(g_idx.numpy(), g_sc.numpy(), k)
    ref_idx, ref_sc = local_topk(E, q, k, 0)
    np.testing.assert_array_equal(m_idx, ref_idx)
    np.testing.assert_array_equal(m_sc, ref_sc)
    np.save(os.path.join(out_dir, 'ok_%d.npy' % rank), m_idx)
    dist.destroy_process_group()


def test_shard_bounds_partition():
    from sert_b200.scoring import shard_bounds
    for rows in (0, 1, 7, 8, 50000, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_all_gather_and_merge_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / 'ok_0.npy'), np.load(tmp_path / 'ok_1.npy')
    np.testing.assert_array_equal(a, b)             # every rank ends with the same merged list


# ---- entity-sharded log-linear step: host-side exchange logic (sert_b200/sharding.py) over gloo ----------------
def ll_worker(rank, world, port, out_dir):
    """Each rank plays one column shard: the device kernels are replaced by numpy stand-ins on the shard's columns,
    the exchanges go through the REAL callback (pointer -> arena view -> collective) that libsert_b200 would call."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import sert_oracle as O
    from sert_b200 import sharding
    from tests import helpers as H
    p = H.ll_problem(3, V=300, E=45, dw=8, W=3, B=16, n_batches=1)
    B, W, E = p['B'], p['W'], p['E']
    x = p['train'][0][:B]
    f = O.loglinear_forward(p['R'], p['Wd'], p['bd'], x)
    ref_ell = O.loglinear_instance_losses(f['o'], O.dense_rows(p['train'][1], 0, B))

    ex = sharding.DistExchange()
    b, e = sharding.shard_bounds(E, world, rank)
    arena = torch.zeros(1 << 16, dtype=torch.uint8)
    cb = ex.bind(arena)
    fl = arena.view(torch.float32).numpy()            # float view of the fake arena
    base = arena.data_ptr()

    def exchange(op, off, count):
        assert cb(None, op, base + 4 * off, count) == 0, ex._error

    def gathered_stats(local, rows):
        """local (rows, E_loc) -> global (max, sum) through the [shard][2][rows] all-gather layout."""
        m = local.max(axis=1)
        fl[rank * 2 * rows: rank * 2 * rows + rows] = m
        fl[rank * 2 * rows + rows: (rank + 1) * 2 * rows] = np.exp(local - m[:, None]).sum(axis=1)
        exchange(sharding.XCHG_ALLGATHER, 0, 2 * rows)
        parts = fl[:world * 2 * rows].reshape(world, 2, rows).copy()
        gm = parts[:, 0].max(axis=0)
        return gm, (parts[:, 1] * np.exp(parts[:, 0] - gm)).sum(axis=0)

    z = f['z'].reshape(B * W, E)[:, b:e]
    gm, gs = gathered_stats(z, B * W)
    pw = np.exp(z - gm[:, None]) / gs[:, None]
    s = np.log(np.clip(pw, 1e-7, np.float32(1 - 1e-7))).reshape(B, W, e - b).sum(axis=1)
    sm, ss = gathered_stats(s, B)
    o = np.exp(s - sm[:, None]) / ss[:, None]
    y = O.dense_rows(p['train'][1], 0, B)[:, b:e]
    fl[4096:4096 + B] = -(y * np.log(np.clip(o, 1e-7, np.float32(1 - 1e-7)))).sum(axis=1)
    exchange(sharding.XCHG_ALLREDUCE_SUM, 4096, B)
    H.close(fl[4096:4096 + B], ref_ell, what='instance losses over %d shards' % world)
    full = ex.gather_columns(p['Wd'][:, b:e], E)
    np.testing.assert_array_equal(full, p['Wd'])
    assert ex.calls == 3
    np.save(os.path.join(out_dir, 'll_ok_%d.npy' % rank), fl[4096:4096 + B])
    dist.destroy_process_group()


def test_loglinear_exchange_world2(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(ll_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / 'll_ok_0.npy'), np.load(tmp_path / 'll_ok_1.npy')
    np.testing.assert_array_equal(a, b)

"""Entity scoring on the device: query x all-entities inner products with a fused running top-k.

Replaces the sklearn NearestNeighbors / cdist + argsort of VectorSpaceCallback.query
(bin/query.py:280-318).  ``ShardedScorer`` row-shards the entity matrix over one process per GPU; the ONE
all-gather of per-shard top-k lists and the merge run inside the library (NCCL, csrc/comm.cu).
"""
import numpy as np

from sert_b200 import _native as N


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('sert_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return torch


class EntityScorer(object):

    def __init__(self, entities, normalise=False, max_queries=1024, max_k=128, row_begin=0, device=None):
        torch = _torch()
        self.lib = N.load()
        entities = np.ascontiguousarray(entities, dtype=np.float32)
        assert entities.ndim == 2
        self.rows, self.d = entities.shape
        self.max_queries, self.max_k = int(max_queries), int(max_k)
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        nbytes = N.c_size_t(0)
        N.check(self.lib.sert_scorer_arena_bytes(self.rows, self.d, self.max_queries, self.max_k,
                                                 N.ctypes.byref(nbytes)))
        with torch.cuda.device(self.device):
            self.arena = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            self.stream = torch.cuda.current_stream(self.device)
            handle = N.c_void_p()
            N.check(self.lib.sert_scorer_create(N.host_ptr(entities), self.rows, self.d, int(row_begin),
                                                int(bool(normalise)), self.max_queries, self.max_k,
                                                N.dev_ptr(self.arena), nbytes.value,
                                                N.c_void_p(self.stream.cuda_stream), N.ctypes.byref(handle)))
        self.handle = handle

    def set_mode(self, mode):
        """'tensor' (default: one coarse bf16 tcgen05 GEMM launch with seeded thresholds and a rigorous score margin,
        exact fp32 re-scoring of the survivors, automatic fall-back on short lists / too many near-ties),
        'tensor_chunked' (the same arithmetic in growing chunks, no threshold seed), 'tensor3' (bf16x3 GEMM + fp32
        re-scoring) or 'fma' (fp32 CUDA-core tiles)."""
        N.check(self.lib.sert_scorer_set_mode(self.handle, {'fma': 0, 'tensor': 1, 'tensor3': 2, 'tensor_chunked': 3}[mode]))

    def plan(self, k):
        """Threshold-seeding plan of a top-k call: dict(group_rows, groups, rank, tile_stride, expected_survivors)."""
        g, G, j = N.c_int32(0), N.c_int32(0), N.c_int32(0)
        stride, T = N.c_int64(0), N.ctypes.c_double(0)
        N.check(self.lib.sert_scorer_plan(self.handle, int(k), N.ctypes.byref(g), N.ctypes.byref(G), N.ctypes.byref(j),
                                          N.ctypes.byref(stride), N.ctypes.byref(T)))
        return dict(group_rows=g.value, groups=G.value, rank=j.value, tile_stride=stride.value,
                    expected_survivors=T.value)

    def stats(self):
        """(calls answered by the seeded one-launch sweep, calls that fell back to the chunked sweeps)."""
        a, b = N.c_int64(0), N.c_int64(0)
        N.check(self.lib.sert_scorer_stats(self.handle, N.ctypes.byref(a), N.ctypes.byref(b)))
        return int(a.value), int(b.value)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.sert_scorer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def topk(self, queries, k, normalise_queries=False, out=None):
        """queries (Q,d) host float32 -> (idx (Q,k) int32 global row ids, score (Q,k) float32), best first.
        ``out`` = (idx, score) arrays to fill; with page-locked ``queries`` / ``out`` (e.g. numpy views of pinned torch
        tensors) the two copies run at PCIe speed instead of through the driver's staging buffers."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        assert queries.ndim == 2 and queries.shape[1] == self.d
        Q = queries.shape[0]
        if out is None:
            idx = np.empty((Q, k), dtype=np.int32)
            score = np.empty((Q, k), dtype=np.float32)
        else:
            idx, score = out
            assert idx.shape == (Q, k) and idx.dtype == np.int32 and idx.flags['C_CONTIGUOUS']
            assert score.shape == (Q, k) and score.dtype == np.float32 and score.flags['C_CONTIGUOUS']
        done = 0
        while done < Q:
            n = min(self.max_queries, Q - done)
            N.check(self.lib.sert_scorer_topk_host(self.handle, N.host_ptr(queries[done:done + n]), n,
                                                   int(bool(normalise_queries)), k,
                                                   N.host_ptr(idx[done:done + n]), N.host_ptr(score[done:done + n])))
            done += n
        return idx, score

    def scores(self, queries, normalise_queries=False):
        """Dense (Q, rows) float32 inner products (the "rank all entities" mode)."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        assert queries.ndim == 2 and queries.shape[1] == self.d
        out = np.empty((queries.shape[0], self.rows), dtype=np.float32)
        N.check(self.lib.sert_scorer_scores_host(self.handle, N.host_ptr(queries), queries.shape[0],
                                                 int(bool(normalise_queries)), N.host_ptr(out)))
        return out

    def topk_dev(self, queries_dev, k, normalise_queries=False):
        """Device in / device out, asynchronous on the scorer's stream."""
        torch = _torch()
        Q = queries_dev.shape[0]
        assert Q <= self.max_queries
        idx = torch.empty((Q, k), dtype=torch.int32, device=self.device)
        score = torch.empty((Q, k), dtype=torch.float32, device=self.device)
        N.check(self.lib.sert_scorer_topk_dev(self.handle, N.dev_ptr(queries_dev), Q, int(bool(normalise_queries)), k,
                                              N.dev_ptr(idx), N.dev_ptr(score)))
        return idx, score


def pack_lists(idx, score):
    """(Q,k) int32 row ids + (Q,k) float32 scores -> one int32 tensor (Q,k,2) so ONE collective moves both."""
    import torch
    return torch.stack([idx, score.view(torch.int32)], dim=-1).contiguous()


def unpack_lists(gathered):
    """(world,Q,k,2) int32 -> ((world,Q,k) int32 ids, (world,Q,k) float32 scores)."""
    import torch
    return gathered[..., 0].contiguous(), gathered[..., 1].contiguous().view(torch.float32)


def all_gather_lists(idx, score, group=None):
    """The exchange step of sharded scoring restated over torch.distributed (gloo in the CPU tests of the host-side
    merge logic; the product path does this all-gather inside the library): every rank receives every rank's
    (ids, scores)[Q,k]."""
    import torch
    import torch.distributed as dist
    packed = pack_lists(idx, score)
    world = dist.get_world_size(group)
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=torch.int32, device=packed.device)
    if packed.is_cuda:
        dist.all_gather_into_tensor(gathered, packed, group=group)         # NCCL: one fused all-gather
    else:
        dist.all_gather(list(gathered.unbind(0)), packed, group=group)     # gloo (CPU tests)
    return unpack_lists(gathered)


def shard_bounds(num_rows, world_size, rank):
    """Contiguous row shards of (almost) equal size: SURVEY.md 8(e)."""
    base, rem = divmod(num_rows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class ShardedScorer(object):
    """Row-sharded scoring, one process per GPU: local GEMM + top-k, ONE all-gather of the per-shard
    (idx, score)[Q,k] lists, k-way merge on every rank -- all three inside libsert_b200 on the scorer's stream
    (``sert_scorer_set_comm``; the all-gather is an ncclAllGather the library issues itself).  ``comm`` is a
    ``sert_b200.comm.Communicator``; by default one is built over the initialised torch.distributed group, which is
    used for nothing but shipping the 128-byte NCCL unique id."""

    def __init__(self, entities_full_or_shard, num_rows_total, group=None, is_shard=False, normalise=False,
                 max_queries=1024, max_k=128, comm=None):
        if comm is None:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size(group) > 1:
                from sert_b200.comm import Communicator
                comm = Communicator.from_torch_distributed(group)
        self.comm = comm
        self.world = comm.world if comm is not None else 1
        self.rank = comm.rank if comm is not None else 0
        begin, end = shard_bounds(num_rows_total, self.world, self.rank)
        shard = entities_full_or_shard if is_shard else entities_full_or_shard[begin:end]
        assert shard.shape[0] == end - begin
        self.local = EntityScorer(shard, normalise=normalise, max_queries=max_queries, max_k=max_k, row_begin=begin)
        if self.world > 1:
            N.check(self.local.lib.sert_scorer_set_comm(self.local.handle, comm.handle))

    def topk_dev(self, queries_dev, k, normalise_queries=False):
        """(idx, score) of the GLOBAL matrix on every rank; asynchronous on the scorer's stream."""
        return self.local.topk_dev(queries_dev, k, normalise_queries)

    def topk(self, queries, k, normalise_queries=False, out=None):
        """Host in / host out through sert_scorer_topk_host (H2D of the queries, D2H of the merged lists)."""
        return self.local.topk(queries, k, normalise_queries, out=out)

    def close(self):
        self.local.close()

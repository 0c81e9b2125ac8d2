"""Entity (column) sharding of the log-linear training step.

The reference trains on one device.  Here the dense layer ``W (dw,E), b (E,)`` of ``sert.models.LanguageModel``
(sert/models.py:846-849) and every (.,E) activation is split into contiguous column shards, one per rank
(one process per GPU); the word table is replicated.  libsert_b200 runs the local kernels and calls back at the
five exchange points of a step (include/sert_b200.h ``sert_exchange_fn``); the classes below implement that
callback:

* ``CommExchange``  -- no callback at all: the library issues the NCCL collectives itself (csrc/comm.cu); the
  product path on GPUs;
* ``DistExchange``  -- ``torch.distributed`` through the callback (gloo in the CPU tests);
* ``LocalExchange`` -- all shards in ONE process on ONE device, one host thread per shard (the C-ABI's threading
  contract), used to test the sharded arithmetic on a single GPU.

Both operate in place on views of the model's HBM arena; nothing here computes anything but the collectives.
"""
import ctypes
import threading

from sert_b200 import _native as N

XCHG_ALLREDUCE_SUM, XCHG_ALLGATHER = 0, 1


def shard_bounds(total, world, rank):
    """Contiguous, balanced split of ``total`` columns (same rule as scoring.shard_bounds)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class Exchange(object):
    """Base: turns (device pointer, count) into a float32 view of the arena and dispatches on ``op``."""

    rank, world = 0, 1

    _error = None
    calls = 0
    floats = 0

    def bind(self, arena, stream=None):
        """Returns the C callback for a model whose HBM arena is ``arena`` (one exchange may serve several
        models of the same rank, e.g. a resumed copy).  ``stream``: the model's torch stream -- the collectives are
        issued under it, whatever stream is current when the library calls back (the library's kernels run on the
        stream the model was created with; a collective ordered on another stream would race with them)."""
        import contextlib
        import torch
        base, nbytes = arena.data_ptr(), arena.numel()

        def trampoline(ctx, op, buf, count):
            try:
                n = count * (self.world if op == XCHG_ALLGATHER else 1)
                off = int(buf) - base
                assert 0 <= off and off + 4 * n <= nbytes, 'exchange buffer outside the arena'
                view = arena[off:off + 4 * n].view(torch.float32)
                scope = torch.cuda.stream(stream) if (stream is not None and arena.is_cuda) else contextlib.nullcontext()
                with scope:
                    if op == XCHG_ALLREDUCE_SUM:
                        self.all_reduce_sum(view)
                    elif op == XCHG_ALLGATHER:
                        self.all_gather(view, int(count))
                    else:
                        raise ValueError('unknown exchange op %d' % op)
                self.calls += 1
                self.floats += int(count)
                return 0
            except BaseException as e:          # never unwind through the C frames
                self._error = e
                return 1

        callback = N.EXCHANGE_FN(trampoline)
        self._callbacks = getattr(self, '_callbacks', []) + [callback]     # keep alive
        return callback

    def raise_pending(self):
        if self._error is not None:
            e, self._error = self._error, None
            raise e

    # host-side helpers used by the model wrapper ---------------------------------------------------------
    def gather_columns(self, local, total):
        """Concatenates per-rank numpy arrays along the last axis (get_dense / predict_fn / checkpoints)."""
        raise NotImplementedError()


class DistExchange(Exchange):
    """torch.distributed collectives on the current stream (the model's stream)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        assert dist.is_initialized(), 'torch.distributed is not initialised'
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, view):
        import torch.distributed as dist
        dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def all_gather(self, view, count):
        import torch.distributed as dist
        mine = view[self.rank * count:(self.rank + 1) * count].clone()
        dist.all_gather_into_tensor(view, mine, group=self.group)

    def gather_columns(self, local, total):
        import numpy as np
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, local, group=self.group)
        out = np.concatenate(parts, axis=-1)
        assert out.shape[-1] == total
        return out


class CommExchange(DistExchange):
    """The five exchanges issued by libsert_b200 itself as NCCL collectives on the model's stream
    (``sert_model_set_entity_shard_comm``): no callback, no torch collective on the step.  torch.distributed is
    used to ship the unique id (``Communicator.from_torch_distributed``) and, off the hot path, to gather column
    shards on the host for checkpoints / predict_fn (``gather_columns``)."""

    def __init__(self, comm=None, group=None):
        DistExchange.__init__(self, group)
        if comm is None:
            from sert_b200.comm import Communicator
            comm = Communicator.from_torch_distributed(group)
        assert (comm.rank, comm.world) == (self.rank, self.world)
        self.comm = comm


class LocalExchange(object):
    """Factory of per-shard exchanges that meet at a thread barrier; all models live on one device and share
    its current stream, so stream order is the enqueue order established by the barrier."""

    def __init__(self, world):
        self.world = int(world)
        self.barrier = threading.Barrier(self.world)
        self.views = [None] * self.world
        self.host = [None] * self.world

    def shard(self, rank):
        return _LocalShard(self, rank)


class _LocalShard(Exchange):
    def __init__(self, hub, rank):
        self.hub, self.rank, self.world = hub, int(rank), hub.world

    def _meet(self, view, combine):
        hub = self.hub
        hub.views[self.rank] = view
        if hub.barrier.wait(timeout=120) == 0:
            combine(hub.views)
        hub.barrier.wait(timeout=120)

    def all_reduce_sum(self, view):
        def combine(views):
            total = views[0].clone()
            for v in views[1:]:
                total += v
            for v in views:
                v.copy_(total)
        self._meet(view, combine)

    def all_gather(self, view, count):
        def combine(views):
            for r, src in enumerate(views):
                block = src[r * count:(r + 1) * count]
                for d, dst in enumerate(views):
                    if d != r:
                        dst[r * count:(r + 1) * count].copy_(block)
        self._meet(view, combine)

    def gather_columns(self, local, total):
        import numpy as np
        hub = self.hub
        hub.host[self.rank] = local
        hub.barrier.wait(timeout=120)
        out = np.concatenate(hub.host, axis=-1)
        hub.barrier.wait(timeout=120)
        assert out.shape[-1] == total
        return out


def attach(native_model, exchange, entity_begin, entities_total):
    """Registers ``exchange`` with a log-linear _NativeModel created over the shard's column count."""
    if isinstance(exchange, CommExchange):
        N.check(native_model.lib.sert_model_set_entity_shard_comm(
            native_model.handle, exchange.comm.handle, int(entity_begin), int(entities_total)))
        native_model.exchange = exchange
        return
    cb = exchange.bind(native_model.arena, getattr(native_model, 'stream', None))
    N.check(native_model.lib.sert_model_set_entity_shard(
        native_model.handle, exchange.rank, exchange.world, int(entity_begin), int(entities_total),
        ctypes.cast(cb, N.c_void_p), None))
    native_model.exchange = exchange

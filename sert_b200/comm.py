"""Library communicator (include/sert_b200.h ``sert_comm_*``): one NCCL communicator per process / GPU.

The collectives of the sharded paths run inside libsert_b200 (csrc/comm.cu); the host's only job is to get the
128-byte NCCL unique id from rank 0 to the other ranks.  ``from_torch_distributed`` ships it through an initialised
``torch.distributed`` group (any backend: the object broadcast goes through the group's store), ``from_env`` through a
file for launchers without torch.distributed.
"""
import ctypes
import os
import time

from sert_b200 import _native as N

ID_BYTES = 128


class Communicator(object):

    def __init__(self, rank, world, unique_id):
        self.lib = N.load()
        self.rank, self.world = int(rank), int(world)
        handle = N.c_void_p()
        buf = ctypes.create_string_buffer(bytes(unique_id), ID_BYTES) if unique_id is not None else None
        N.check(self.lib.sert_comm_init(self.rank, self.world, buf, ctypes.byref(handle)))
        self.handle = handle

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(ID_BYTES)
        N.check(N.load().sert_comm_unique_id(buf, ID_BYTES))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, group=None):
        """Rank 0 creates the id, torch.distributed broadcasts the 128 bytes; the current CUDA device joins."""
        import torch.distributed as dist
        assert dist.is_initialized(), 'torch.distributed is not initialised'
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 and world > 1 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(rank, world, box[0])

    @classmethod
    def from_file(cls, rank, world, path, timeout=120.0):
        """Rendezvous through a file rank 0 writes (atomic rename)."""
        if world > 1 and rank == 0:
            with open(path + '.tmp', 'wb') as f:
                f.write(cls.unique_id())
            os.replace(path + '.tmp', path)
        uid = None
        if world > 1:
            deadline = time.time() + timeout
            while not os.path.exists(path):
                if time.time() > deadline:
                    raise RuntimeError('no NCCL unique id at %s after %.0f s' % (path, timeout))
                time.sleep(0.01)
            with open(path, 'rb') as f:
                uid = f.read()
        return cls(rank, world, uid)

    def info(self):
        r, w, v = N.c_int32(0), N.c_int32(0), N.c_int32(0)
        n, b = N.c_int64(0), N.c_int64(0)
        N.check(self.lib.sert_comm_info(self.handle, ctypes.byref(r), ctypes.byref(w), ctypes.byref(v),
                                        ctypes.byref(n), ctypes.byref(b)))
        return {'rank': r.value, 'world': w.value, 'nccl_version': v.value, 'collectives': n.value, 'bytes': b.value}

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.sert_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Host-buffer (numpy in, numpy out) form of the sharded Sobol iteration: a thin ctypes driver of the
C-ABI entry points shgo_b200_ctx_shard_* / shgo_b200_iterate_sobol_host_shard_* (include/shgo_b200.h).

One object per rank.  torch.distributed is used for the two things the C ABI leaves to the caller:
exchanging the CUDA-IPC handles once, and the barrier between `begin` and `finish` of an iteration.
Reference path replaced: SHGO.iterate_delaunay's sampling + process_pools + minimizers
(shgo/_shgo.py:1052-1157) minus qhull, for one shard of the points.
"""
import ctypes

import numpy as np
import torch.distributed as dist

from . import _lib, _sobol


class ShardedHostIteration:
    def __init__(self, device, nshards, my_shard, rows_per_shard, group=None):
        self.group, self.nshards, self.rows = group, nshards, rows_per_shard
        self.ctx = ctypes.c_void_p()
        _lib.call("shgo_b200_ctx_create", device, ctypes.byref(self.ctx))
        mine = (ctypes.c_ubyte * 128)()
        _lib.call("shgo_b200_ctx_shard_init", self.ctx, nshards, my_shard, rows_per_shard, mine)
        handles = [None] * nshards
        if nshards > 1:
            dist.all_gather_object(handles, bytes(mine), group=group)
        else:
            handles[0] = bytes(mine)
        allh = (ctypes.c_ubyte * (128 * nshards)).from_buffer_copy(b"".join(handles))
        _lib.call("shgo_b200_ctx_shard_connect", self.ctx, allh)
        self._sv = {}
        self._cnt, self._nfev = ctypes.c_uint64(), ctypes.c_uint64()
        self._pool = None

    def iterate(self, functor, gset, d, bounds, row_ptr, col, remote_at_ends=False, pool_cap=None, pool_out=None):
        """row_ptr: int64 [rows + 1] GLOBAL offsets of this rank's rows, col: int32 matching slice with
        GLOBAL vertex ids (numpy, or pinned torch tensors exposing data_ptr()).  -> (pool ids, nfev)."""
        b = np.asarray(bounds, np.float64)
        lb, ub = np.ascontiguousarray(b[:, 0]), np.ascontiguousarray(b[:, 1])
        sv = self._sv.get(d)
        if sv is None:
            sv = self._sv[d] = np.ascontiguousarray(_sobol.direction_numbers(d))
        cap = self.rows if pool_cap is None else pool_cap
        if pool_out is None:
            if self._pool is None or len(self._pool) < cap:
                self._pool = np.empty(max(cap, 1), np.int64)
            pool_out = self._pool
        ptr = lambda a: a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data
        _lib.call("shgo_b200_iterate_sobol_host_shard_begin", self.ctx, functor, gset, d, sv.ctypes.data,
                  lb.ctypes.data, ub.ctypes.data, ptr(row_ptr), ptr(col), 1 if remote_at_ends else 0)
        if self.nshards > 1:
            dist.barrier(group=self.group)     # every shard of f is written
        _lib.call("shgo_b200_iterate_sobol_host_shard_finish", self.ctx, ptr(pool_out), cap,
                  ctypes.byref(self._cnt), ctypes.byref(self._nfev))
        k = min(self._cnt.value, cap)
        out = pool_out[:k] if not hasattr(pool_out, "data_ptr") else pool_out[:k].numpy()
        return out, self._nfev.value

    @property
    def pool_count(self):
        return self._cnt.value

    def close(self):
        if self.ctx:
            if self.nshards > 1:
                dist.barrier(group=self.group)   # nobody reads a peer's buffer any more
            _lib.lib().shgo_b200_ctx_destroy(self.ctx)
            self.ctx = None

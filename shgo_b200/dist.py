"""Multi-GPU hot path: one process per GPU (torch.distributed), vertices sharded by index range.

SURVEY.md section 8(e): rank r of G owns Sobol indices / graph rows [r n/G, (r+1) n/G).  Point generation
has a closed-form skip-ahead, so every shard generates and evaluates locally with no communication.
The star test needs neighbours' f across shards.  Two exchanges are implemented:

  'p2p'        (default on GPUs) the f shards are allocated with cudaMalloc, exported with CUDA IPC and
               peer-mapped into every rank over NVLink.  Phase 1 classifies every row against its
               LOCAL neighbours (no communication); phase 2 finishes the few rows that survived --
               their remote neighbours' f values are read in place over NVLink by the kernel itself,
               8 bytes per remote neighbour of a candidate.  Collectives are used for plumbing only
               (one stream-ordered barrier per iteration, pool gather).
  'fused'      (needs remote neighbours at the row ends) the same data path in ONE kernel: every rank
               publishes "my shard is written" into a flag word of every peer right after its
               evaluation, and the star kernel fetches the remote neighbours of its candidates in
               place over NVLink as soon as the owner's flag is up -- cp.async straight into a ring in
               shared memory, consumed three tiles later, so no register waits for the round trip.
               What was met before a peer was ready is finished by the phase-2 kernel after the
               barrier, as in 'p2p'.  Measured on 8 x B200 it is SLOWER than 'p2p' (DESIGN.md section 7:
               the candidate bookkeeping costs the memory-bound local phase more than the hidden NVLink
               time gives back), so it is an option, not the default.
  'allgather'  one all_gather of the f vector (NCCL on GPUs; gloo in the CPU tests), then the
               single-GPU star kernel on the rank's rows.  8 n (G-1)/G bytes received per GPU.

torch.distributed is plumbing here; the kernels are the ones of libshgo_b200.so.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import hotpath as hp


class _RawCudaBuffer:
    """cudaMalloc'ed memory exposed to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, nelem, typestr):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


class PeerMappedVector:
    """A vector of `rows` elements per rank (float64, or int32 words with typestr '<i4'), readable
    and writable by every rank's kernels (cudaMalloc + CUDA IPC, mapped over NVLink)."""

    def __init__(self, rows, group=None, typestr="<f8"):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows = rows
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.call("shgo_b200_ipc_alloc", max(rows * int(typestr[2:]), 256), ctypes.byref(ptr), handle)
        self.own_ptr = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_ptrs = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peer_ptrs.append(self.own_ptr)
            else:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                _lib.call("shgo_b200_ipc_open", buf, ctypes.byref(p))
                self.peer_ptrs.append(p.value)
                self._opened.append(p.value)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.table = torch.tensor(self.peer_ptrs, dtype=torch.int64, device=dev)
        self.local = torch.as_tensor(_RawCudaBuffer(self.own_ptr, rows, typestr), device=dev)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            _lib.call("shgo_b200_ipc_close", p)
        self._opened = []
        dist.barrier(group=self.group)
        if self.own_ptr:
            _lib.call("shgo_b200_ipc_free", self.own_ptr)
            self.own_ptr = None


def _all_gather(out, mine, group):
    try:
        dist.all_gather_into_tensor(out, mine, group=group)
    except (RuntimeError, NotImplementedError):      # backends without the flat variant
        parts = list(out.chunk(dist.get_world_size(group)))
        dist.all_gather(parts, mine.contiguous(), group=group)


def shard_range(n, world, rank):
    """rows [row0, row0 + rows) of `rank`; n must divide evenly (power-of-two problem sizes)."""
    if n % world:
        raise ValueError("n = %d is not divisible by the number of ranks %d" % (n, world))
    rows = n // world
    return rank * rows, rows


class ShardedSobolIteration:
    """One shgo iteration of the Sobol hot path over all ranks of `group`.

    row_ptr_local / col_local: this rank's CSR rows; row_ptr_local holds GLOBAL offsets starting at
    row_ptr[row0] (i.e. a slice of the global row_ptr), col_local the matching slice of col with
    GLOBAL vertex ids."""

    def __init__(self, functor, gset, tab, n, row_ptr_local, col_local, group=None, exchange="p2p",
                 pool_cap=None, remote_at_ends=False, chunks=1, phase1_warps=0, filter_remote=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.functor, self.gset, self.tab, self.n = functor, gset, tab, n
        self.row0, self.rows = shard_range(n, self.world, self.rank)
        self.exchange = exchange if self.world > 1 else "none"
        self.row_ptr, self.col = row_ptr_local, col_local
        dev = row_ptr_local.device
        self.dev = dev
        self.nnz_local = col_local.numel()
        self.feas = torch.empty(self.rows, dtype=torch.uint8, device=dev)
        self.is_min = torch.empty(self.rows, dtype=torch.uint8, device=dev)
        self.nfev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.work = hp.PoolWorkspace(self.rows, self.rows if pool_cap is None else pool_cap, dev)
        self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.peer = None
        self.off0 = int(row_ptr_local[0].item())
        # promise: in every row the neighbours owned by other ranks form a prefix and/or suffix
        # (true for ascending rows, i.e. for every CSR built by csr_from_simplices)
        self.remote_at_ends = bool(remote_at_ends)
        self._k = 0
        if self.exchange == "fused" and not self.remote_at_ends:
            self.exchange = "p2p"      # the fused kernel needs the remote_at_ends promise
        if self.exchange in ("p2p", "fused"):
            # two peer-mapped buffers used alternately: iteration k+1 writes the other buffer while
            # slow peers may still be reading buffer k, so ONE barrier per iteration is enough
            # (a rank passes barrier k+1 only after every peer has finished its reads of step k)
            self.peers = [PeerMappedVector(self.rows, group), PeerMappedVector(self.rows, group)]
            self.peer = self.peers[0]
            self.f_my = self.peer.local
            self.f_all = None
            ws = _lib.lib().shgo_b200_star_remote_workspace(self.rows, self.rows)
            self._remote_tmp = torch.empty(ws, dtype=torch.uint8, device=dev)
            # phase 1 lists its candidates here, phase 2 consumes the list (no compaction pass)
            self._cand_ws = torch.empty(_lib.lib().shgo_b200_star_local_workspace(self.rows), dtype=torch.uint8,
                                        device=dev)
            self._comm = torch.cuda.Stream(device=dev)
            # Conservative filter of the remote phase: every rank also produces the minimum of each
            # 32 consecutive f values (1/32 of f) and PUSHES that small vector into every peer's
            # copy with copy-engine transfers while phase 1 runs; a candidate whose f is below the
            # block minimum of a remote neighbour needs no 8-byte NVLink read for it (about 45 % of
            # them).  Two alternating sets of copies, like f.
            self.filter_remote = (bool(filter_remote) and self.exchange == "p2p" and self.rows % 32 == 0
                                  and self.world > 1)
            self._plog = hp.blockmin_layout(functor, gset)
            self.filter_remote = self.filter_remote and self._plog >= 0 and self.rows % (256 << self._plog) == 0
            if self.filter_remote:
                self._bmin = [PeerMappedVector(n // 32, group), PeerMappedVector(n // 32, group)]
            # 'p2p' with chunks > 1: the rows are classified in chunks; phase 2 of chunk c (bound by the
            # rate of small NVLink reads) runs on a second stream underneath phase 1 of chunk c+1
            # (bound by HBM), which keeps at most `phase1_warps` warps per SM resident so that the
            # phase-2 blocks find room.  One candidate list per chunk (phase 2 of chunk c reads its list
            # while phase 1 of chunk c+1 writes the next).
            self.chunks = max(1, int(chunks)) if self.exchange == "p2p" else 1
            self.phase1_warps = int(phase1_warps) if self.chunks > 1 else 0
            per = (self.rows // self.chunks) & ~31
            self._chunk_rows = [per if c < self.chunks - 1 else self.rows - c * per for c in range(self.chunks)]
            if self.chunks > 1:
                self._remote = torch.cuda.Stream(device=dev)
                self._cand_ws = [torch.empty(_lib.lib().shgo_b200_star_local_workspace(nr), dtype=torch.uint8,
                                             device=dev) for nr in self._chunk_rows]

            if self.exchange == "fused":
                # ready[q] on this rank = last iteration whose shard rank q has completely written
                self._ready = PeerMappedVector(max(self.world, 64), group, typestr="<i4")
                self._ready.local.zero_()
                torch.cuda.synchronize()
                dist.barrier(group=group)
        else:
            self.f_all = torch.empty(n if self.world > 1 else self.rows, dtype=torch.float64, device=dev)
            self.f_my = self.f_all[self.row0:self.row0 + self.rows] if self.world > 1 else self.f_all

    def _barrier_async(self):
        """Stream-ordered barrier "every rank has finished writing its f shard", issued on a side
        stream right after the evaluation so that its latency (and the skew between ranks) hides
        behind phase 1 of the star test; returns nothing, _barrier_wait() joins it."""
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self._token, group=self.group)

    def _barrier_wait(self):
        torch.cuda.current_stream().wait_stream(self._comm)

    def step(self, events=None):
        """One iteration; free of host synchronisation (capturable in a CUDA graph)."""
        if self.peer is not None:
            self.peer = self.peers[self._k & 1]
            self.f_my = self.peer.local
            self._k += 1
        bmin = None
        if self.peer is not None and getattr(self, "filter_remote", False):
            bm = self._bmin[(self._k - 1) & 1]
            lo, cnt = self.row0 // 32, self.rows // 32
            bmin = bm.local
            hp.eval_sobol_blockmin(self.functor, self.gset, self.tab, self.row0, self.rows, self.f_my, self.feas,
                                   bm.local)
        else:
            hp.eval_sobol(self.functor, self.gset, self.tab, self.row0, self.rows, f=self.f_my, feasible=self.feas)
        if events:
            events[0].record()
        if bmin is not None:
            # push this rank's block minima into every peer's copy (copy engines, side stream); the
            # barrier that follows on the same stream then also covers them
            self._comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm):
                src = bm.own_ptr + lo * 8
                for q, base in enumerate(bm.peer_ptrs):
                    if q != self.rank:
                        _lib.call("shgo_b200_copy_async", base + lo * 8, src, cnt * 8, self._comm.cuda_stream)
        if self.exchange == "fused":
            # tell every peer that this shard is written, then classify: candidates read their
            # remote neighbours in place inside the kernel once the owner's word is up.  The barrier
            # (for what the kernel met too early, and for the reuse of the f buffers) runs on the
            # side stream underneath.
            hp.publish_ready(self._ready.table, self.world, self.rank, self._k)
            self._barrier_async()
            hp.star_csr_fused(self.row_ptr, self.col, self.row0, self.peer.table, self.f_my, self.feas, self.world,
                              self.rows, self.rank, self._ready.local, self._k, self.is_min, self._cand_ws,
                              self.nfev, self.off0)
            self._barrier_wait()
            hp.star_csr_remote(self.row_ptr, self.col, self.row0, self.peer.table, self.peer.own_ptr, self.world,
                               self.rows, self.rank, self.is_min, None, self.off0, remote_at_ends=True,
                               cand_ws=self._cand_ws)
        elif self.exchange == "p2p" and self.chunks > 1:
            self._barrier_async()
            main = torch.cuda.current_stream()
            self._remote.wait_stream(self._comm)
            r0 = self.row0
            for c, nr in enumerate(self._chunk_rows):
                hp.star_csr_local(self.row_ptr, self.col, self.row0, self.f_my, self.feas, self.is_min,
                                  self.nfev, self.off0, r0, nr, self._cand_ws[c], self.remote_at_ends,
                                  max_warps=self.phase1_warps, accumulate_nfev=c > 0)
                self._remote.wait_stream(main)
                with torch.cuda.stream(self._remote):
                    hp.star_csr_remote(self.row_ptr, self.col, self.row0, self.peer.table, self.peer.own_ptr,
                                       self.world, self.rows, self.rank, self.is_min, None, self.off0, r0, nr,
                                       self.remote_at_ends, self._cand_ws[c])
                r0 += nr
            main.wait_stream(self._remote)
        elif self.exchange == "p2p":
            # Phase 1 needs nobody else; the barrier (every shard of f complete before anyone reads
            # it remotely) runs concurrently with it on a side stream.
            self._barrier_async()
            hp.star_csr_local(self.row_ptr, self.col, self.row0, self.f_my, self.feas, self.is_min, self.nfev,
                              self.off0, cand_ws=self._cand_ws, remote_at_ends=self.remote_at_ends)
            self._barrier_wait()
            hp.star_csr_remote(self.row_ptr, self.col, self.row0, self.peer.table, self.peer.own_ptr, self.world,
                               self.rows, self.rank, self.is_min, self._remote_tmp, self.off0,
                               remote_at_ends=self.remote_at_ends, cand_ws=self._cand_ws, block_min=bmin,
                               layout_plog=getattr(self, "_plog", 0))
        else:
            if self.exchange == "allgather":
                _all_gather(self.f_all, self.f_my, self.group)
            f_all = self.f_all
            if self.world == 1:        # single rank: f is the shard itself, addressed globally
                hp.star_csr(self.row_ptr, self.col, f_all, is_min=self.is_min, feasible=self.feas,
                            nfev=self.nfev)
            else:
                hp.star_csr_shard(self.row_ptr, self.col, self.row0, f_all, self.feas, self.is_min,
                                  self.nfev, self.off0)
        if events:
            events[1].record()
        hp.compact_pool(self.is_min, base=self.row0, work=self.work)

    def local_pool(self):
        c = int(self.work.count[0].item())
        return self.work.idx[:min(c, self.work.cap)]

    def global_pool(self):
        """ascending pool indices of the whole problem, on every rank (pools are small in real use)"""
        mine = self.local_pool().cpu().numpy()
        if self.world == 1:
            return mine
        parts = [None] * self.world
        dist.all_gather_object(parts, mine, group=self.group)
        return np.concatenate(parts)

    def global_nfev(self):
        t = self.nfev.clone()
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self):
        if self.peer is not None:
            for p in self.peers:
                p.close()
            if self.exchange == "fused":
                self._ready.close()
            if getattr(self, "filter_remote", False):
                for b in self._bmin:
                    b.close()
            self.peer = None
            self.peers = []


def split_range(n, world, rank):
    """[a, b) of `rank` when n items are dealt out as evenly as contiguous ranges allow"""
    return rank * n // world, (rank + 1) * n // world


class ShardedLatticeIteration:
    """One whole generation of the simplicial hypercube complex over all ranks of `group`
    (lattice path).

    The adjacency is implicit and a star reaches across the whole id range, so every rank keeps a
    full-size copy of f.  Ownership is granule-cyclic: the id range is cut into granules of
    hp.LATTICE_GRANULE consecutive vertices and rank r evaluates / classifies granules r, r+G, ...
    -- the cost of a star test is wildly uneven (a rejected vertex: a handful of reads; a true
    minimiser of the 10-D complex: 59 048) and clusters in id space, contiguous slabs measured
    0.21 .. 0.88 ms per rank at G = 8.  Exchanges:

      'push'       (GPUs) the copies of f are cudaMalloc'ed, exported with CUDA IPC and mapped into
                   every rank; the evaluation kernel stores each value into all copies at once over
                   NVLink (shgo_b200_lattice_eval_push), so there is no separate all-gather pass.
                   One stream-ordered barrier per generation; two alternating sets of copies make a
                   second barrier unnecessary (see ShardedSobolIteration).
      'allreduce'  plain collective baseline (NCCL on GPUs, gloo in the CPU tests): every rank
                   evaluates its granules into a zero-filled vector, one all_reduce(SUM) completes
                   it (x + 0 is exact; infeasible vertices are +inf on exactly one rank).
    """

    def __init__(self, functor, gset, d, gen, tab, group=None, exchange="push", granule=None):
        from . import lattice as lat
        self.granule = hp.LATTICE_GRANULE if granule is None else int(granule)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.functor, self.gset, self.d, self.gen, self.tab = functor, gset, d, gen, tab
        self.nv, self.ngrid = lat.counts(d, gen)
        self.exchange = exchange if self.world > 1 else "none"
        dev = tab.device
        self.dev = dev
        self.peers = []
        if self.exchange == "push":
            self.peers = [PeerMappedVector(self.nv, group), PeerMappedVector(self.nv, group)]
            self._comm = torch.cuda.Stream(device=dev)
            self._token = torch.zeros(1, dtype=torch.int32, device=dev)
            self.f = self.peers[0].local
        else:
            self.f = torch.zeros(self.nv, dtype=torch.float64, device=dev)
        self._k = 0
        self.feas = torch.zeros(self.nv, dtype=torch.uint8, device=dev)   # stays 0 outside own granules
        self.is_min = torch.empty(self.nv, dtype=torch.uint8, device=dev)
        self.work = hp.PoolWorkspace(self.nv, self.nv, dev)
        ws = _lib.lib().shgo_b200_lattice_star_workspace(self.nv)
        self._star_tmp = torch.empty(ws, dtype=torch.uint8, device=dev)

    def owned(self):
        """boolean mask (host) of the vertices this rank owns"""
        g = np.arange(self.nv) // self.granule
        return (g % self.world) == self.rank

    def step(self, events=None):
        d, gen, nv = self.d, self.gen, self.nv
        if self.exchange == "push":
            pv = self.peers[self._k & 1]
            self._k += 1
            self.f = pv.local
            others = [p for r, p in enumerate(pv.peer_ptrs) if r != self.rank]
            hp.lattice_eval_push(self.functor, self.gset, d, gen, self.tab, 0, nv, self.f, self.feas,
                                 others, self.rank, self.world, self.granule)
            # "every copy of f is complete": stream-ordered, joined before the star test
            main = torch.cuda.current_stream()
            self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(self._token, group=self.group)
            main.wait_stream(self._comm)
        else:
            if self.exchange == "allreduce":
                self.f.zero_()
            hp.lattice_eval_push(self.functor, self.gset, d, gen, self.tab, 0, nv, self.f, self.feas, [],
                                 self.rank, self.world, self.granule)
            if self.exchange == "allreduce":
                dist.all_reduce(self.f, group=self.group)
        if events:
            events[0].record()
        hp.lattice_star(d, gen, self.f, 0, nv, is_min=self.is_min, tmp=self._star_tmp, part=self.rank,
                        nparts=self.world, granule=self.granule)
        if events:
            events[1].record()
        hp.compact_pool(self.is_min, base=0, work=self.work)

    def local_pool(self):
        c = int(self.work.count[0].item())
        return self.work.idx[:min(c, self.work.cap)]

    def global_pool(self):
        """ascending pool (vertex ids) of the whole complex, on every rank"""
        mine = self.local_pool().cpu().numpy()
        if self.world > 1:
            parts = [None] * self.world
            dist.all_gather_object(parts, mine, group=self.group)
            mine = np.concatenate(parts)
        return np.sort(mine)

    def global_nfev(self):
        t = self.feas.sum(dtype=torch.int64).reshape(1).clone()
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self):
        for p in self.peers:
            p.close()
        self.peers = []

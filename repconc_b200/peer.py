"""NVLink peer-memory all-reduce of the Sinkhorn row sums (rc_peer_allreduce_f64).

torch.distributed's symmetric memory provides the peer-mapped buffers (plumbing); the exchange itself --
publish, system-scope flags, rank-ordered sum from peer memory -- is one kernel of librepconc_b200.so.
If symmetric memory cannot be set up (single process, no P2P, RC_PEER_ALLREDUCE=0) the caller keeps
using NCCL; both give every rank the same bits.
"""
import ctypes
import logging
import os

import torch
import torch.distributed as dist

from . import _lib

logger = logging.getLogger(__name__)
FLAG_PEER_TIMEOUT = 16


class PeerAllReduce:
    _cache = {}
    _disabled_reason = None

    @classmethod
    def get(cls, n, device, group=None):
        """Cached instance for vectors of n doubles on `device`, or None when unavailable."""
        if os.environ.get("RC_PEER_ALLREDUCE", "1") in ("0", "false", "False"):
            return None
        if cls._disabled_reason is not None:
            return None
        group = group if group is not None else dist.group.WORLD
        key = (int(n), str(device), id(group))
        inst = cls._cache.get(key)
        if inst is None:
            try:
                inst = cls(int(n), device, group)
            except Exception as e:  # symmetric memory not available on this system / build
                cls._disabled_reason = f"{type(e).__name__}: {e}"
                logger.warning("repconc_b200: peer all-reduce unavailable (%s); using NCCL", cls._disabled_reason)
                return None
            cls._cache[key] = inst
        return inst

    def __init__(self, n, device, group):
        import torch.distributed._symmetric_memory as symm
        lib = _lib.load()
        self.lib, self.n, self.group = lib, n, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = lib.rc_peer_allreduce_buffer_bytes(n)
        with torch.cuda.device(device):
            self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.hdl = symm.rendezvous(self.buf, group)
            ptrs = list(self.hdl.buffer_ptrs)
            assert len(ptrs) == self.world
            self.ptrs = (ctypes.c_uint64 * self.world)(*ptrs)
            self.hdl.barrier()          # every rank's flags are zero before the first exchange
        self.seq = 0

    def all_reduce(self, vec, flags):
        """in-place SUM over ranks of the fp64 CUDA tensor `vec` (n elements); `flags` is the int32 flag word"""
        assert vec.dtype == torch.float64 and vec.numel() == self.n and vec.is_contiguous()
        self.seq += 1
        _lib.check(self.lib.rc_peer_allreduce_f64(self.ptrs, self.rank, self.world, self.n, self.seq,
                                                  vec.data_ptr(), flags.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream),
                   "rc_peer_allreduce_f64")


class PeerSolve:
    """Symmetric buffers + sequence counter for rc_sinkhorn_solve_peer (the persistent Sinkhorn kernel with the
    row-sum exchange fused in).  One instance per (M, K, device, group); None when symmetric memory is unavailable."""
    _cache = {}

    @classmethod
    def get(cls, M, K, device, group=None):
        if os.environ.get("RC_PEER_ALLREDUCE", "1") in ("0", "false", "False"):
            return None
        if PeerAllReduce._disabled_reason is not None:
            return None
        group = group if group is not None else dist.group.WORLD
        if dist.get_world_size(group) > 16:
            return None
        key = (int(M), int(K), str(device), id(group))
        inst = cls._cache.get(key)
        if inst is None:
            try:
                inst = cls(int(M), int(K), device, group)
            except Exception as e:  # symmetric memory not available on this system / build
                PeerAllReduce._disabled_reason = f"{type(e).__name__}: {e}"
                logger.warning("repconc_b200: peer exchange unavailable (%s); using NCCL", PeerAllReduce._disabled_reason)
                return None
            cls._cache[key] = inst
        return inst

    def __init__(self, M, K, device, group):
        import torch.distributed._symmetric_memory as symm
        lib = _lib.load()
        self.lib, self.group = lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = lib.rc_sinkhorn_peer_buffer_bytes(M, K)
        with torch.cuda.device(device):
            self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.hdl = symm.rendezvous(self.buf, group)
            ptrs = list(self.hdl.buffer_ptrs)
            assert len(ptrs) == self.world
            self.ptrs = (ctypes.c_uint64 * self.world)(*ptrs)
            self.hdl.barrier()          # every rank's flags are zero before the first exchange
        self.seq = 0

    def take_seq(self, iters):
        """sequence numbers of one call: returns seq_base and advances by the number of exchanges (iters, min 1)"""
        base = self.seq
        self.seq += max(int(iters), 1)
        return base

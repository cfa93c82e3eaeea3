"""Host-side operators over the C ABI (include/repconc_b200.h).

PyTorch is plumbing here: device memory, the current CUDA stream and torch.distributed for the
three all-reduces of the constrained assignment.  All arithmetic of the path runs in
librepconc_b200.so.  CPU tensors are rejected -- there is no CPU implementation.
"""
import logging

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

logger = logging.getLogger(__name__)

FLAG_NONFINITE = 1
FLAG_AMPLITUDE = 2
FLAG_BADCODE = 4
FLAG_SPARSE_UNSAFE = 8
FLAG_PEER_TIMEOUT = 16
N_FLAG_BITS = 5


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.RepconcLibraryError(
            f"repconc_b200: `{name}` must be a CUDA tensor (got {type(t).__name__}"
            f"{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}); "
            "the B200 path has no CPU implementation")


def _rows_f32(x, name):
    """fp32, unit inner stride; row stride may exceed the width (views are fine)."""
    _require_cuda(x, name)
    if x.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tensor, got shape {tuple(x.shape)}")
    if x.dtype != torch.float32:
        x = x.float()                      # fp16/bf16 under autocast promote exactly like the reference
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def _centroids_f32(c):
    _require_cuda(c, "centroids")
    if c.dim() != 3:
        raise ValueError(f"centroids: expected (M, K, dsub), got {tuple(c.shape)}")
    c = c.detach()
    if c.dtype != torch.float32:
        c = c.float()
    return c.contiguous()


def _ld(x):
    return x.stride(0) if x.shape[0] > 1 else x.shape[1]


def _check_shapes(x, c):
    M, K, ds = c.shape
    if x.shape[1] != M * ds:
        raise ValueError(f"embedding width {x.shape[1]} != M*dsub = {M}*{ds}")
    return M, K, ds


# ------------------------------------------------------------------------------------------
# a2  NN assign  (modeling_repconc.py:47-52,66)
# ------------------------------------------------------------------------------------------
def nn_assign(x, centroids, uint8=False):
    """codes[b,m] = argmin_k ||x[b,m,:]-c[m,k,:]||^2.
    Returns the reference's (B,M) int64 view of an (M,B) buffer, or with uint8=True a
    contiguous (B,M) uint8 tensor (evaluate_repconc.py:69 fused)."""
    lib = _lib.load()
    x = _rows_f32(x, "continuous_embeds")
    c = _centroids_f32(centroids)
    M, K, ds = _check_shapes(x, c)
    B = x.shape[0]
    with torch.cuda.device(x.device):
        if uint8:
            out = torch.empty((B, M), dtype=torch.uint8, device=x.device)
            mb, u8 = None, out.data_ptr()
        else:
            out = torch.empty((M, B), dtype=torch.int64, device=x.device)
            mb, u8 = out.data_ptr(), None
        if B:
            _lib.check(lib.rc_nn_assign(x.data_ptr(), _ld(x), c.data_ptr(), B, M, K, ds, mb, u8, _stream()),
                       "rc_nn_assign")
    return out if uint8 else out.t()


# ------------------------------------------------------------------------------------------
# f3  fused encode epilogue: rotation + optional normalisation + NN assign (+ uint8 pack)
#     (modeling_repconc.py:98-103, evaluate_repconc.py:64-70)
# ------------------------------------------------------------------------------------------
ENCODE_DS = (4, 8, 12, 16, 24, 32)       # sub-vector sizes rc_encode_assign is compiled for


def encode_assign(pooled, rotation, centroids, normalize=False, uint8=False, return_rotated=True, out=None):
    """RepCONC.forward after the encoder for use_constraint = False, in one kernel:
        rotated = pooled @ rotation.T ; [per-sub-vector L2 normalise] ; codes = argmin_k ||rotated[:, m] - c[m, k]||^2
    Returns (rotated (B, D) fp32 or None, codes): codes as the reference's (B, M) int64 view, or with uint8=True a
    contiguous (B, M) uint8 tensor -- `out` (a (B, M) uint8 CUDA view, e.g. the tail of an index's code storage)
    receives them in place."""
    lib = _lib.load()
    x = _rows_f32(pooled, "dense_embed")
    c = _centroids_f32(centroids)
    M, K, ds = _check_shapes(x, c)
    _require_cuda(rotation, "rotation")
    r = rotation.detach().float().contiguous()
    D = M * ds
    if tuple(r.shape) != (D, D):
        raise ValueError(f"rotation {tuple(r.shape)} != ({D}, {D})")
    if ds not in ENCODE_DS:
        raise NotImplementedError(f"encode_assign: sub-vector dimension {ds} not in {ENCODE_DS}")
    B = x.shape[0]
    with torch.cuda.device(x.device):
        rotated = torch.empty((B, D), dtype=torch.float32, device=x.device) if return_rotated else None
        if uint8:
            if out is None:
                out = torch.empty((B, M), dtype=torch.uint8, device=x.device)
            if out.dtype != torch.uint8 or tuple(out.shape) != (B, M) or not out.is_contiguous() or not out.is_cuda:
                raise ValueError("encode_assign: `out` must be a contiguous (B, M) uint8 CUDA tensor")
            mb, u8 = None, out.data_ptr()
        else:
            out = torch.empty((M, B), dtype=torch.int64, device=x.device)
            mb, u8 = out.data_ptr(), None
        if B:
            _lib.check(lib.rc_encode_assign(x.data_ptr(), _ld(x), r.data_ptr(), c.data_ptr(), B, M, K, ds,
                                            1 if normalize else 0, rotated.data_ptr() if return_rotated else None,
                                            D, mb, u8, _stream()), "rc_encode_assign")
    return rotated, (out if uint8 else out.t())


# ------------------------------------------------------------------------------------------
# a1+a3  table + extrema  (modeling_repconc.py:50,76-77)
# ------------------------------------------------------------------------------------------
def dist_table(x, centroids):
    """-> (table (M,B,K) fp32, minmax (2,M) fp32 [max; min], flags int32[1])."""
    lib = _lib.load()
    x = _rows_f32(x, "continuous_embeds")
    c = _centroids_f32(centroids)
    M, K, ds = _check_shapes(x, c)
    B = x.shape[0]
    if B < 1:
        raise ValueError("dist_table: empty batch")
    with torch.cuda.device(x.device):
        table = torch.empty((M, B, K), dtype=torch.float32, device=x.device)
        minmax = torch.empty((2, M), dtype=torch.float32, device=x.device)
        flags = torch.zeros(1, dtype=torch.int32, device=x.device)
        st = _stream()
        _lib.check(lib.rc_minmax_init(minmax.data_ptr(), M, st), "rc_minmax_init")
        _lib.check(lib.rc_dist_table(x.data_ptr(), _ld(x), c.data_ptr(), B, M, K, ds, table.data_ptr(),
                                     minmax.data_ptr(), flags.data_ptr(), st), "rc_dist_table")
    return table, minmax, flags


class CudaAssignKernels:
    """The kernel side of the constrained assignment, as `constrained_assign_driver` needs it.
    (Tests drive the same host sequence with an emulated kernel set under gloo.)"""

    def __init__(self, x, centroids):
        self.lib = _lib.load()
        self.x = _rows_f32(x, "continuous_embeds")
        self.c = _centroids_f32(centroids)
        self.M, self.K, self.ds = _check_shapes(self.x, self.c)
        self.B = self.x.shape[0]
        self.device = self.x.device

    def table(self):
        self.tab, self.minmax, self.flags = dist_table(self.x, self.c)
        return self.minmax

    def _alloc_state(self):
        nbytes = self.lib.rc_sinkhorn_state_bytes(self.B, self.M, self.K)
        self.state = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        p = self.lib.rc_sinkhorn_rowsum_ptr(self.state.data_ptr(), self.B, self.M, self.K)
        off = p - self.state.data_ptr()
        self.P = self.state[off: off + self.M * self.K * 8].view(torch.float64).view(self.M, self.K)
        self.nsteps = 0

    def solve(self, eps, iters, uint8=False, dense=False):
        """single-rank Sinkhorn + argmax in one library call (rc_sinkhorn_solve); needs table() first"""
        self._alloc_state()
        if uint8:
            out = torch.empty((self.B, self.M), dtype=torch.uint8, device=self.device)
            mb, u8 = None, out.data_ptr()
        else:
            out = torch.empty((self.M, self.B), dtype=torch.int64, device=self.device)
            mb, u8 = out.data_ptr(), None
        _lib.check(self.lib.rc_sinkhorn_solve(self.tab.data_ptr(), self.minmax.data_ptr(), self.B, self.M, self.K,
                                              float(eps), int(iters), 1 if dense else 0, self.state.data_ptr(), mb,
                                              u8, self.flags.data_ptr(), _stream()), "rc_sinkhorn_solve")
        return out if uint8 else out.t()

    def solve_fused(self, eps, iters, B_global, group=None, uint8=False, dense=False):
        """W ranks: Sinkhorn + argmax with the per-iteration exchange of the row sums inside the persistent kernel
        (rc_sinkhorn_solve_peer).  Returns None when that path does not apply (dense re-run, K != 256, no peer
        memory) -- the caller then runs the step-wise sequence with all-reduces in between.  Needs table() and the
        all-reduced extrema first."""
        if dense or self.K != 256 or iters < 1:
            return None
        from .peer import PeerSolve
        ps = PeerSolve.get(self.M, self.K, self.device, group)
        if ps is None:
            return None
        self._alloc_state()
        if uint8:
            out = torch.empty((self.B, self.M), dtype=torch.uint8, device=self.device)
            mb, u8 = None, out.data_ptr()
        else:
            out = torch.empty((self.M, self.B), dtype=torch.int64, device=self.device)
            mb, u8 = out.data_ptr(), None
        _lib.check(self.lib.rc_sinkhorn_solve_peer(self.tab.data_ptr(), self.minmax.data_ptr(), self.B, int(B_global),
                                                   self.M, self.K, float(eps), int(iters), self.state.data_ptr(),
                                                   ps.ptrs, ps.rank, ps.world, ps.take_seq(iters), mb, u8,
                                                   self.flags.data_ptr(), _stream()), "rc_sinkhorn_solve_peer")
        return out if uint8 else out.t()

    def begin(self, eps):
        self._alloc_state()
        _lib.check(self.lib.rc_sinkhorn_begin(self.tab.data_ptr(), self.minmax.data_ptr(), self.B, self.M, self.K,
                                              float(eps), self.state.data_ptr(), self.flags.data_ptr(), _stream()),
                   "rc_sinkhorn_begin")
        return self.P

    def step(self, eps, B_global, dense=False):
        _lib.check(self.lib.rc_sinkhorn_step(self.tab.data_ptr(), self.B, int(B_global), self.M, self.K, float(eps),
                                             self.nsteps, 1 if dense else 0, self.state.data_ptr(),
                                             self.flags.data_ptr(), _stream()),
                   "rc_sinkhorn_step")
        self.nsteps += 1
        return self.P

    def finish(self, eps, apply_rowsum, uint8=False, B_global=None, dense=False):
        if uint8:
            out = torch.empty((self.B, self.M), dtype=torch.uint8, device=self.device)
            mb, u8 = None, out.data_ptr()
        else:
            out = torch.empty((self.M, self.B), dtype=torch.int64, device=self.device)
            mb, u8 = out.data_ptr(), None
        _lib.check(self.lib.rc_sinkhorn_finish(self.tab.data_ptr(), self.B,
                                               int(B_global) if B_global is not None else self.B, self.M,
                                               self.K, float(eps), 1 if apply_rowsum else 0, self.nsteps,
                                               1 if dense else 0, self.state.data_ptr(), mb, u8,
                                               self.flags.data_ptr(), _stream()), "rc_sinkhorn_finish")
        return out if uint8 else out.t()

    def read_flags(self):
        return int(self.flags.item())

    def clear_flags(self):
        self.flags.zero_()

    def set_dense(self, dense):
        """process-wide DEFAULT of the pass selection (tests / A-B runs); the product passes `dense` per call.
        Returns the previous default."""
        return bool(self.lib.rc_sinkhorn_set_dense(1 if dense else 0))


def reduce_flags(flags, distributed, group=None):
    """The flag word of an assignment as EVERY rank must see it: the bits are OR-ed over the ranks (a bit such as
    SPARSE_UNSAFE can be raised from rank-local data -- pool exhaustion -- and the ranks must take the same
    decision or their collective sequences diverge).  `flags`: int32[1] tensor, or an int (emulated kernels)."""
    if not distributed:
        return int(flags.item()) if isinstance(flags, torch.Tensor) else int(flags)
    if isinstance(flags, torch.Tensor):
        bits = ((flags.reshape(1) >> torch.arange(N_FLAG_BITS, device=flags.device)) & 1).to(torch.int32)
    else:
        bits = torch.tensor([(int(flags) >> i) & 1 for i in range(N_FLAG_BITS)], dtype=torch.int32)
    dist.all_reduce(bits, dist.ReduceOp.MAX, group=group)      # OR, bit by bit (NCCL has no BOR)
    return sum(int(b) << i for i, b in enumerate(bits.tolist()))


def constrained_assign_driver(kern, eps, iters, distributed, group=None, uint8=False, dense=False):
    """Host sequence of RepCONC.quantize with use_constraint=True (modeling_repconc.py:53-66):
    which kernel runs when, and where the reference's three all-reduces go
      * MAX / MIN of the per-sub-vector extrema       (:78-80)
      * SUM of the row sums, once per iteration       (:157; the total-sum all-reduce of :151 cancels
        in the first row normalisation and is not needed)
      * B *= world_size                               (:150)
    `kern` provides table/begin/step/finish/read_flags (CudaAssignKernels in the product).
    `dense`: run the dense Sinkhorn pass (the re-run after RC_FLAG_SPARSE_UNSAFE); it is an argument of every
    kernel call, not process state."""
    world = dist.get_world_size(group) if distributed else 1
    iters = max(int(iters), 0)
    minmax = kern.table()
    if distributed:
        # one collective for both extrema: MAX over [max, -min]
        minmax[1].neg_()
        dist.all_reduce(minmax, dist.ReduceOp.MAX, group=group)
        minmax[1].neg_()
    B_global = kern.B * world
    if not distributed and hasattr(kern, "solve") and not _stepwise():
        codes = kern.solve(eps, iters, uint8=uint8, dense=dense)
        return _check_assign_flags(kern, codes, eps, iters, distributed, group, uint8, dense)
    if distributed and hasattr(kern, "solve_fused") and not _stepwise():
        # W ranks, one persistent kernel per rank: the row sums are exchanged over NVLink peer memory inside it
        codes = kern.solve_fused(eps, iters, B_global, group, uint8=uint8, dense=dense)
        if codes is not None:
            return _check_assign_flags(kern, codes, eps, iters, distributed, group, uint8, dense)
    P = kern.begin(eps)
    # row-sum exchange: one-shot NVLink peer-memory all-reduce when available, NCCL otherwise
    reducer = None
    if distributed and isinstance(P, torch.Tensor) and P.is_cuda:
        from .peer import PeerAllReduce
        reducer = PeerAllReduce.get(P.numel(), P.device, group)

    def sum_rows(P):
        if reducer is not None:
            reducer.all_reduce(P, kern.flags)
        else:
            dist.all_reduce(P, dist.ReduceOp.SUM, group=group)

    for _ in range(max(iters - 1, 0)):
        if distributed:
            sum_rows(P)
        P = kern.step(eps, B_global, dense=dense)
    if distributed and iters >= 1:
        sum_rows(P)
    codes = kern.finish(eps, iters >= 1, uint8=uint8, B_global=B_global, dense=dense)
    return _check_assign_flags(kern, codes, eps, iters, distributed, group, uint8, dense)


def _stepwise():
    """RC_SINKHORN_STEPWISE=1: assignments go through begin/step/finish with the row sums exchanged between the
    calls, instead of the single persistent kernel (tests, debugging)"""
    import os
    return os.environ.get("RC_SINKHORN_STEPWISE", "0") not in ("", "0")


def _check_assign_flags(kern, codes, eps, iters, distributed, group, uint8, dense):
    flags = reduce_flags(kern.flags if hasattr(kern, "flags") else kern.read_flags(), distributed, group)
    if flags & FLAG_PEER_TIMEOUT:
        raise _lib.RepconcLibraryError("peer all-reduce timed out: a rank never reached the exchange "
                                       "(RC_PEER_TIMEOUT_MS bounds the wait)")
    if flags & FLAG_SPARSE_UNSAFE and not dense:
        # the sparse pass's error bound did not hold for this input (a centroid kept < 2^-8/K of the mass, or the
        # survivor pool ran out -- the latter depends on the rank's own rows): redo the whole assignment with
        # the dense pass.  `flags` is OR-ed over the ranks, so every rank takes this branch together.
        if hasattr(kern, "clear_flags"):
            kern.clear_flags()
        return constrained_assign_driver(kern, eps, iters, distributed, group, uint8, dense=True)
    if flags & FLAG_AMPLITUDE:
        raise AssertionError("amplitude > 0 (center_distance_for_constraint)")
    if flags & FLAG_NONFINITE:
        logger.warning("Sinkhorn Algorithm returns nan/inf values.")
    return codes


def constrained_assign(x, centroids, eps, iters, distributed=None, group=None, uint8=False):
    """RepCONC.quantize, use_constraint=True.  `distributed=None` follows the reference:
    the collective path is on iff torch.distributed is initialised (modeling_repconc.py:61)."""
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized()
    kern = CudaAssignKernels(x, centroids)
    if kern.B < 1:
        raise ValueError("constrained_assign: empty batch")
    with torch.cuda.device(kern.device):
        return constrained_assign_driver(kern, eps, iters, distributed, group, uint8)


# ------------------------------------------------------------------------------------------
# a6  decode + its backward  (modeling_repconc.py:168-184)
# ------------------------------------------------------------------------------------------
def _codes_i64(codes, device):
    if codes.dtype != torch.int64:
        codes = codes.long()
    if codes.device != device:
        codes = codes.to(device)
    return codes


def decode_forward(codes, centroids):
    lib = _lib.load()
    c = _centroids_f32(centroids)
    M, K, ds = c.shape
    if codes.dim() != 2 or codes.shape[1] != M:
        raise ValueError(f"decode: codes must be (bs, {M}), got {tuple(codes.shape)}")
    B = codes.shape[0]
    with torch.cuda.device(c.device):
        out = torch.empty((B, M * ds), dtype=torch.float32, device=c.device)
        if B == 0:
            return out
        flags = torch.zeros(1, dtype=torch.int32, device=c.device)
        if codes.dtype == torch.uint8 and codes.is_contiguous() and codes.device == c.device:
            rc = lib.rc_decode(None, 0, 0, codes.data_ptr(), c.data_ptr(), B, M, K, ds, out.data_ptr(),
                               flags.data_ptr(), _stream())
        else:
            codes = _codes_i64(codes, c.device)
            rc = lib.rc_decode(codes.data_ptr(), codes.stride(0), codes.stride(1), None, c.data_ptr(), B, M, K, ds,
                               out.data_ptr(), flags.data_ptr(), _stream())
        _lib.check(rc, "rc_decode")
        if _check_codes() and int(flags.item()) & FLAG_BADCODE:
            # the reference's `centroids[first_indices, second_indices]` raises on an out-of-range code
            raise IndexError(f"decode: a code is outside [0, {K})")
    return out


def _check_codes():
    """RC_CHECK_CODES=0 skips the read-back of the bad-code flag after rc_decode (one small device -> host copy per
    call; out-of-range codes are then clamped silently instead of raising like the reference's indexing)."""
    import os
    return os.environ.get("RC_CHECK_CODES", "1") not in ("0", "false", "False")


def decode_backward(codes, grad_q, centroid_shape):
    """grad_centroids[m,k,:] = sum_{b: codes[b,m]==k} grad_q[b,m,:]  (deterministic)."""
    lib = _lib.load()
    M, K, ds = centroid_shape
    grad_q = _rows_f32(grad_q, "grad_quantized")
    codes = _codes_i64(codes, grad_q.device)
    B = codes.shape[0]
    with torch.cuda.device(grad_q.device):
        grad_c = torch.empty((M, K, ds), dtype=torch.float32, device=grad_q.device)
        nws = lib.rc_decode_bwd_workspace_bytes(B, M, K, ds)
        ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=grad_q.device)
        _lib.check(lib.rc_decode_bwd(codes.data_ptr(), codes.stride(0), codes.stride(1), grad_q.data_ptr(),
                                     _ld(grad_q), B, M, K, ds, grad_c.data_ptr(), ws.data_ptr(), _stream()),
                   "rc_decode_bwd")
    return grad_c


class _Decode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, codes, centroids):
        ctx.save_for_backward(codes)
        ctx.cshape = tuple(centroids.shape)
        ctx.cdtype = centroids.dtype
        return decode_forward(codes, centroids)

    @staticmethod
    def backward(ctx, grad_out):
        (codes,) = ctx.saved_tensors
        grad_c = decode_backward(codes, grad_out, ctx.cshape)
        return None, grad_c.to(ctx.cdtype)


def decode(codes, centroids):
    """Drop-in for modeling_repconc.decode: torch codes (any int dtype / device) -> differentiable
    w.r.t. `centroids`; numpy codes -> numpy output (the reference's numpy branch)."""
    if isinstance(codes, torch.Tensor):
        assert isinstance(centroids, torch.Tensor)
        _require_cuda(centroids, "centroids")
        return _Decode.apply(codes, centroids)
    elif isinstance(codes, np.ndarray):
        if isinstance(centroids, torch.Tensor):
            _require_cuda(centroids, "centroids")
            c = centroids.detach()
        else:
            c = torch.from_numpy(np.ascontiguousarray(centroids, dtype=np.float32)).cuda()
        ct = torch.from_numpy(np.ascontiguousarray(codes).astype(np.int64, copy=False)).to(c.device)
        return decode_forward(ct, c).cpu().numpy()
    else:
        raise NotImplementedError()


# ------------------------------------------------------------------------------------------
# a8  quantisation loss  (finetune_repconc.py:367-374,389-396)
# ------------------------------------------------------------------------------------------
class _QuantLoss(torch.autograd.Function):
    """(mse_loss, surrogate) = f(x, centroids; codes, cached_grads, w) with
    mse_loss = ((q - x)**2).sum(-1).mean() * w, surrogate = <g,x> + <g,q>, q = decode(codes)."""

    @staticmethod
    def forward(ctx, x, centroids, codes, cached_grads, weight):
        lib = _lib.load()
        x32 = _rows_f32(x, "continuous_embeds")
        c = _centroids_f32(centroids)
        M, K, ds = _check_shapes(x32, c)
        n = x32.shape[0]
        codes = _codes_i64(codes, x32.device)
        g = _rows_f32(cached_grads, "cached_grads") if cached_grads is not None else None
        with torch.cuda.device(x32.device):
            out2 = torch.empty(2, dtype=torch.float32, device=x32.device)
            ws = torch.empty(lib.rc_mse_workspace_bytes(n, M, K, ds), dtype=torch.uint8, device=x32.device)
            _lib.check(lib.rc_mse_fwd(x32.data_ptr(), _ld(x32), None, 0, g.data_ptr() if g is not None else None,
                                      _ld(g) if g is not None else 0, codes.data_ptr(), codes.stride(0),
                                      codes.stride(1), c.data_ptr(), n, M, K, ds, float(weight), out2.data_ptr(),
                                      ws.data_ptr(), _stream()), "rc_mse_fwd")
        ctx.save_for_backward(x32, c, codes, g if g is not None else torch.empty(0, device=x32.device))
        ctx.has_g = g is not None
        ctx.weight = float(weight)
        ctx.x_dtype, ctx.c_dtype = x.dtype, centroids.dtype
        return out2[0], out2[1]

    @staticmethod
    def backward(ctx, grad_mse, grad_sur):
        lib = _lib.load()
        x32, c, codes, g = ctx.saved_tensors
        g = g if ctx.has_g else None
        M, K, ds = c.shape
        n = x32.shape[0]
        # upstream scalars (the AMP loss scale arrives through grad_mse); one host sync, as .backward() has
        gm = float(grad_mse) if grad_mse is not None else 0.0
        gs = float(grad_sur) if grad_sur is not None else 0.0
        need_x, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with torch.cuda.device(x32.device):
            grad_x = torch.empty_like(x32, memory_format=torch.contiguous_format) if need_x else None
            grad_c = torch.empty_like(c) if need_c else None
            ws = torch.empty(lib.rc_mse_workspace_bytes(n, M, K, ds), dtype=torch.uint8, device=x32.device)
            _lib.check(lib.rc_mse_bwd(x32.data_ptr(), _ld(x32), None, 0, g.data_ptr() if g is not None else None,
                                      _ld(g) if g is not None else 0, codes.data_ptr(), codes.stride(0),
                                      codes.stride(1), c.data_ptr(), n, M, K, ds, ctx.weight, gm, gs,
                                      grad_x.data_ptr() if need_x else None, None,
                                      grad_c.data_ptr() if need_c else None, ws.data_ptr(), _stream()),
                       "rc_mse_bwd")
        if grad_x is not None and grad_x.dtype != ctx.x_dtype:
            grad_x = grad_x.to(ctx.x_dtype)
        if grad_c is not None and grad_c.dtype != ctx.c_dtype:
            grad_c = grad_c.to(ctx.c_dtype)
        return grad_x, grad_c, None, None, None


def quantization_loss(continuous_embeds, centroids, codes, cached_grads, mse_loss_weight):
    """Fused replacement of finetune_repconc.py:367-374 for a document chunk:
        surrogate = dot(g, x) + dot(g, decode(codes));  mse_loss = ((q-x)**2).sum(-1).mean() * w
    Returns (mse_loss, surrogate), both differentiable w.r.t. continuous_embeds and centroids, so
    `(scaler.scale(mse_loss) + surrogate).backward()` (:390/:396) works unchanged."""
    _require_cuda(continuous_embeds, "continuous_embeds")
    _require_cuda(centroids, "centroids")
    return _QuantLoss.apply(continuous_embeds, centroids, codes, cached_grads, mse_loss_weight)

"""B200-native drop-in for `repconc.models.repconc.modeling_repconc`.

Same public surface as the reference module (src/repconc/models/repconc/modeling_repconc.py):
`RepCONC(nn.Module)` with the same constructor, attributes, methods and state-dict keys
(`centroids`, `rotation`, `dense_encoder.*`), `QuantizeOutput`, `sinkhorn_algorithm`, `decode`.
The arithmetic of quantize / decode runs in librepconc_b200.so (hand-written sm_100a kernels);
nothing here falls back to PyTorch ops or the CPU.
"""
import logging
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops
from .ops import decode  # noqa: F401  (module-level `decode`, modeling_repconc.py:168-184)

try:  # HF container type, exactly as the reference uses it (:13,:21-25)
    from transformers.modeling_outputs import ModelOutput
except Exception:  # pragma: no cover - transformers is part of the image
    ModelOutput = object

logger = logging.getLogger(__name__)


@dataclass
class QuantizeOutput(ModelOutput):
    continuous_embeds: Optional[torch.FloatTensor] = None
    quantized_embeds: Optional[torch.FloatTensor] = None
    discrete_codes: Optional[torch.LongTensor] = None


class RepCONC(nn.Module):
    """modeling_repconc.py:28-134.  `dense_encoder` is any module returning (bs, hidden) embeddings
    and carrying `.config.hidden_size` (examples/ance/modeling_ance.py:34-43 plug-in boundary)."""

    def __init__(self, config, dense_encoder, use_constraint: bool, sk_epsilon: float, sk_iters: int):
        super().__init__()
        self.config = config
        self.dense_encoder = dense_encoder
        # so we can use the rotation matrix of OPQ
        self.register_buffer('rotation', torch.eye(dense_encoder.config.hidden_size))
        self.centroids = nn.Parameter(
            torch.randn((config.MCQ_M, config.MCQ_K, config.hidden_size // config.MCQ_M)))
        if self.config.similarity_metric == "METRIC_CENTROID_COS":
            self.normalize_centrodis()
        self.centroids.requires_grad = True
        self.use_constraint, self.sk_epsilon, self.sk_iters = use_constraint, sk_epsilon, sk_iters

    @torch.no_grad()
    def quantize(self, continuous_embeds):
        """(bs, D) -> (bs, M) int64 codes, a transposed view of an (M, bs) buffer like the
        reference's `codes.t()` (:66).  Constraint off: argmin (:51-52).  Constraint on: centred
        table + Sinkhorn + argmax (:53-65); the collective path is taken iff torch.distributed is
        initialised (:61,:78), and NaN/Inf in Q only warns (:64-65)."""
        if not self.use_constraint:
            return ops.nn_assign(continuous_embeds, self.centroids)
        return ops.constrained_assign(continuous_embeds, self.centroids, self.sk_epsilon, self.sk_iters,
                                      distributed=dist.is_available() and dist.is_initialized())

    def decode(self, codes):
        # codes: bs, M
        return decode(codes, self.centroids)

    @staticmethod
    def center_distance_for_constraint(distances):
        """(M, bs, K) fp32 table -> centred table, modeling_repconc.py:73-85.  Kept for API parity;
        quantize() itself centres in place inside rc_sinkhorn_begin."""
        return center_distance_for_constraint(distances)

    def forward(self, input_ids, attention_mask, discrete_codes=None, return_code=False,
                return_quantized_embedding=False):
        dense_embed = self.dense_encoder(input_ids=input_ids, attention_mask=attention_mask)
        cos = self.config.similarity_metric == "METRIC_CENTROID_COS"
        if (discrete_codes is None and (return_code or return_quantized_embedding) and not self.use_constraint
                and not torch.is_grad_enabled() and dense_embed.is_cuda
                and self.centroids.shape[-1] in ops.ENCODE_DS):
            # corpus encoding (evaluate_repconc.py:64-70 runs under no_grad with use_constraint = False): rotation,
            # optional normalisation and NN assign in one kernel
            rotated_embed, discrete_codes = ops.encode_assign(dense_embed, self.rotation, self.centroids, normalize=cos)
            quantized_embeds = self.decode(discrete_codes) if return_quantized_embedding else None
            return QuantizeOutput(continuous_embeds=rotated_embed, quantized_embeds=quantized_embeds,
                                  discrete_codes=discrete_codes)
        rotated_embed = dense_embed @ self.rotation.T
        if self.config.similarity_metric == "METRIC_CENTROID_COS":
            rotated_embed = F.normalize(
                rotated_embed.reshape(len(rotated_embed), self.config.MCQ_M, -1), p=2, dim=-1
            ).reshape_as(rotated_embed)
        if discrete_codes is None and (return_code or return_quantized_embedding):
            discrete_codes = self.quantize(rotated_embed)
        quantized_embeds = self.decode(discrete_codes) if return_quantized_embedding else None
        return QuantizeOutput(
            continuous_embeds=rotated_embed,
            quantized_embeds=quantized_embeds,
            discrete_codes=discrete_codes,
        )

    @torch.no_grad()
    def normalize_centrodis(self):
        centroids = self.centroids.data.clone()
        centroids = F.normalize(centroids, dim=-1, p=2)
        self.centroids.data.copy_(centroids)

    def save_pretrained(self, output_dir):
        state_dict = self.state_dict()
        torch.save(state_dict, os.path.join(output_dir, "pytorch_model.bin"))
        self.config.save_pretrained(output_dir)
        self.dense_encoder.save_pretrained(os.path.join(output_dir, 'dense_encoder'))

    @classmethod
    def from_pretrained(cls, load_dir, use_constraint, sk_epsilon, sk_iters, encoder_loader=None):
        """modeling_repconc.py:124-134.  The dense encoder is not part of this package: by default
        it is loaded with the reference's own `AutoDense` (repconc.models.dense); pass
        `encoder_loader(path) -> nn.Module` to plug in any other encoder class."""
        if encoder_loader is None:
            try:
                from repconc.models.dense.modeling_dense import AutoDense
            except ImportError as e:  # pragma: no cover
                raise ImportError(
                    "RepCONC.from_pretrained needs an encoder loader: install the reference package "
                    "(repconc.models.dense.AutoDense) or pass encoder_loader=...") from e
            encoder_loader = AutoDense.from_pretrained
        dense_encoder = encoder_loader(os.path.join(load_dir, 'dense_encoder'))
        repconc = cls(dense_encoder.config, dense_encoder, use_constraint=use_constraint,
                      sk_epsilon=sk_epsilon, sk_iters=sk_iters)
        repconc.load_state_dict(torch.load(os.path.join(load_dir, "pytorch_model.bin"), map_location="cpu"))
        return repconc


@torch.no_grad()
def center_distance_for_constraint(distances):
    """Device implementation of the static method: runs the centring step of rc_sinkhorn_begin on a
    copy of `distances` (M, bs, K) and returns the centred fp32 table."""
    ops._require_cuda(distances, "distances")
    lib = ops._lib.load()
    d = distances.float().contiguous().clone()
    M, B, K = d.shape
    mx = d.amax(dim=(1, 2))
    mn = d.amin(dim=(1, 2))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(mx, dist.ReduceOp.MAX)
        dist.all_reduce(mn, dist.ReduceOp.MIN)
    minmax = torch.stack([mx, mn]).contiguous()
    with torch.cuda.device(d.device):
        state = torch.empty(lib.rc_sinkhorn_state_bytes(B, M, K), dtype=torch.uint8, device=d.device)
        flags = torch.zeros(1, dtype=torch.int32, device=d.device)
        # eps only scales the row sums this call discards
        ops._lib.check(lib.rc_sinkhorn_begin(d.data_ptr(), minmax.data_ptr(), B, M, K, 1.0, state.data_ptr(),
                                             flags.data_ptr(), ops._stream()), "rc_sinkhorn_begin")
    assert not (int(flags.item()) & ops.FLAG_AMPLITUDE), "amplitude > 0"
    return d


@torch.no_grad()
def sinkhorn_algorithm(out: Tensor, epsilon: float, sinkhorn_iterations: int, use_distrib_train: bool):
    """modeling_repconc.py:137-165 as a function: `out` (M, K, B) = -centred distances (values in
    [-1, 1], fp64 or fp32); returns Q (M, K, B) fp64 with columns summing to 1.

    Contract on `out`: the kernels iterate on an fp32 table, so `out` is cast to fp32.  That is EXACT for what
    RepCONC.quantize passes (`-center_distance_for_constraint(...)`: fp32 values widened by `.double()`, :54-56)
    and for any fp32 input; an fp64 `out` carrying more than 24 significant bits is rounded to fp32 first,
    i.e. the result is the reference's for `out.float().double()`.

    The training path (RepCONC.quantize) never materialises Q; this entry point exists for callers
    that want the transport plan itself.  It runs the same kernels on the (M,B,K) fp32 table rebuilt
    from `out` and expands Q from the row scaling with rc_sinkhorn_expand."""
    ops._require_cuda(out, "out")
    lib = ops._lib.load()
    M, K, B = out.shape
    world = dist.get_world_size() if use_distrib_train else 1
    _identity_centring_check()
    with torch.cuda.device(out.device):
        state = torch.empty(lib.rc_sinkhorn_state_bytes(B, M, K), dtype=torch.uint8, device=out.device)
        # identity centring: middle = 0 and amplitude = 1 in fp32, so rc_sinkhorn_begin leaves the table as is
        minmax = torch.empty((2, M), dtype=torch.float32, device=out.device)
        minmax[0] = 1.0 - 1e-5
        minmax[1] = -(1.0 - 1e-5)
        sp = ops._stream()
        base = state.data_ptr()
        off = lib.rc_sinkhorn_rowsum_ptr(base, B, M, K) - base
        P = state[off: off + M * K * 8].view(torch.float64).view(M, K)
        Q = torch.empty((M, K, B), dtype=torch.float64, device=out.device)
        if sinkhorn_iterations <= 0:
            # no iteration: the reference returns exp(out/eps) / sum(Q) * B, whose columns are NOT normalised
            # (:141-152,164).  Not on any training path; three plain torch ops in fp64 keep the exact values.
            Q = torch.exp(out.double() / epsilon)
            sum_Q = Q.sum(-1, keepdim=True).sum(-2, keepdim=True)
            if use_distrib_train:
                dist.all_reduce(sum_Q)
            Q /= sum_Q
            Q *= B * world
            return Q
        for dense in (0, 1):
            # the kernels consume the (M,B,K) fp32 centred table; `out` is its negated transpose
            table = (-out).transpose(1, 2).float().contiguous()
            flags = torch.zeros(1, dtype=torch.int32, device=out.device)
            ops._lib.check(lib.rc_sinkhorn_begin(table.data_ptr(), minmax.data_ptr(), B, M, K, float(epsilon),
                                                 base, flags.data_ptr(), sp), "rc_sinkhorn_begin")
            for it in range(sinkhorn_iterations - 1):
                if use_distrib_train:
                    dist.all_reduce(P)
                ops._lib.check(lib.rc_sinkhorn_step(table.data_ptr(), B, B * world, M, K, float(epsilon), it,
                                                    dense, base, flags.data_ptr(), sp), "rc_sinkhorn_step")
            if use_distrib_train:
                dist.all_reduce(P)
            ops._lib.check(lib.rc_sinkhorn_expand(table.data_ptr(), B, B * world, M, K, float(epsilon), 1, base,
                                                  Q.data_ptr(), flags.data_ptr(), sp), "rc_sinkhorn_expand")
            fl = ops.reduce_flags(flags, use_distrib_train)
            if not (fl & ops.FLAG_SPARSE_UNSAFE):
                break                                     # otherwise: once more with the dense pass, on every rank
    return Q


def _identity_centring_check():
    # middle = (max + min) / 2 == 0 and amplitude = max - middle + 1e-5 == 1 in fp32
    mx = np.float32(1.0) - np.float32(1e-5)
    assert np.float32(mx + np.float32(1e-5)) == np.float32(1.0)

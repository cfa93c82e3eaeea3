"""OPQ / PQ warm-up on the GPU -- the `warmup_from_embeds` of src/repconc/train/run_warmup.py:85-132.

The reference hands this step to Faiss (`index_factory("OPQ{M},PQ{M}x8")`, `index.train`, `index.add`).
Here it runs on the path's own kernels: the NN assignment of every Lloyd iteration is `rc_nn_assign`, the
cluster sums are the deterministic scatter-add of `rc_decode_bwd`, the cluster sizes `rc_code_histogram`,
the quantisation error `rc_mse_fwd`.  PyTorch supplies the rotation GEMMs and the d x d SVD of the
Procrustes step (library calls on 768 x 768 matrices, not the hot path).

Algorithm (Faiss 1.7.1 semantics, restated; oracle/oracle_np.py holds the numpy restatement the tests
compare against):
  * k-means per sub-vector (`ProductQuantizer::train` -> `Clustering::train`): initial centroids = K
    training points drawn by a seeded permutation (the same permutation for every sub-vector), `niter`
    Lloyd iterations (assign, mean), empty clusters re-seeded by splitting the largest cluster with the
    +-1/1024 perturbation of `Clustering::split_clusters` (largest cluster chosen deterministically).
  * OPQ (`OPQMatrix::train`): random orthogonal start, `niter` = 50 alternations of {rotate, train PQ
    (40 Lloyd iterations the first time, then 4 with a hot start), encode / decode, orthogonal Procrustes
    via SVD of recons^T x}; at most 65 536 training points.
  * final PQ on the rotated vectors (25 iterations, 256 points per centroid), then every corpus vector is
    encoded (`index.add`).
PARITY NOTE: Faiss itself is not installable here, its random streams differ from
torch's, and its BLAS summation order is not reproducible, so the result is not bit-comparable with a
Faiss run; the tests pin every Lloyd / Procrustes step against the numpy oracle from identical states and
check the properties the algorithm guarantees (monotone objective, orthogonal rotation).
"""
import logging

import numpy as np
import torch

from . import _lib, ops
from .faiss_compat import METRIC_INNER_PRODUCT, IndexPQ, copy_array_to_vector

logger = logging.getLogger(__name__)

SPLIT_EPS = 1.0 / 1024.0          # Clustering.cpp: EPS of split_clusters
OPQ_MAX_TRAIN_POINTS = 256 * 256  # OPQMatrix::max_train_points
PQ_MAX_POINTS_PER_CENTROID = 256  # ClusteringParameters::max_points_per_centroid


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _subsample(x, nmax, seed):
    """fvecs_maybe_subsample: a seeded random subset of nmax rows (order of the permutation)"""
    n = x.shape[0]
    if n <= nmax:
        return x
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    perm = torch.randperm(n, generator=g)[:nmax].to(x.device)
    return x.index_select(0, perm).contiguous()


def initial_centroids(x, M, K, seed):
    """K training points per sub-vector, the same seeded permutation for every sub-vector"""
    n, D = x.shape
    if n < K:
        raise ValueError(f"k-means needs at least K={K} training points, got {n}")
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    perm = torch.randperm(n, generator=g)[:K].to(x.device)
    ds = D // M
    return x.index_select(0, perm).view(K, M, ds).permute(1, 0, 2).contiguous()


def lloyd_step(x, centroids):
    """One Lloyd iteration on the device.
    -> (new centroids (M,K,ds) fp32, objective of the INPUT centroids = mean_b ||x_b - q_b||^2,
        codes (B,M) int64 view, counts (M,K) int32 before the empty-cluster split)"""
    lib = _lib.load()
    M, K, ds = centroids.shape
    B = x.shape[0]
    codes = ops.nn_assign(x, centroids)                                  # rc_nn_assign
    with torch.cuda.device(x.device):
        # objective: rc_mse_fwd with weight 1 (decode fused)
        out2 = torch.empty(2, dtype=torch.float32, device=x.device)
        ws = torch.empty(lib.rc_mse_workspace_bytes(B, M, K, ds), dtype=torch.uint8, device=x.device)
        _lib.check(lib.rc_mse_fwd(x.data_ptr(), x.stride(0), None, 0, None, 0, codes.data_ptr(), codes.stride(0),
                                  codes.stride(1), centroids.data_ptr(), B, M, K, ds, 1.0, out2.data_ptr(),
                                  ws.data_ptr(), _stream()), "rc_mse_fwd")
        sums = ops.decode_backward(codes, x, (M, K, ds))                  # rc_decode_bwd: deterministic sums
        counts = code_histogram(codes, K)
    new_c = torch.where(counts.unsqueeze(-1) > 0, sums / counts.clamp(min=1).unsqueeze(-1).float(), centroids)
    new_c = split_empty_clusters(new_c, counts)
    return new_c, out2[0], codes, counts


def code_histogram(codes, K):
    """counts[m,k] = #{b: codes[b,m]==k}  (rc_code_histogram); codes (B,M) int64 (any strides) or uint8"""
    lib = _lib.load()
    if not codes.is_cuda:
        raise _lib.RepconcLibraryError("code_histogram: `codes` must be a CUDA tensor")
    B, M = codes.shape
    if B == 0:
        return torch.zeros((M, K), dtype=torch.int32, device=codes.device)
    with torch.cuda.device(codes.device):
        counts = torch.empty((M, K), dtype=torch.int32, device=codes.device)
        flags = torch.zeros(1, dtype=torch.int32, device=codes.device)
        if codes.dtype == torch.uint8 and codes.is_contiguous():
            rc = lib.rc_code_histogram(None, 0, 0, codes.data_ptr(), B, M, K, counts.data_ptr(), flags.data_ptr(),
                                       _stream())
        else:
            c64 = codes if codes.dtype == torch.int64 else codes.long()
            rc = lib.rc_code_histogram(c64.data_ptr(), c64.stride(0), c64.stride(1), None, B, M, K,
                                       counts.data_ptr(), flags.data_ptr(), _stream())
        _lib.check(rc, "rc_code_histogram")
    return counts


def split_empty_clusters(centroids, counts):
    """Clustering::split_clusters with a deterministic donor: every empty cluster (ascending k) takes a
    copy of the currently largest cluster of its sub-vector (first maximum); even coordinates of the copy
    are scaled by 1+eps and of the donor by 1-eps, odd coordinates the other way round; the donor's size
    is halved between the two.  Host loop over the (rare) empty clusters only."""
    empty = (counts == 0).nonzero()
    if empty.numel() == 0:
        return centroids
    c = centroids.clone()
    cnt = counts.clone().cpu().numpy().astype(np.int64)
    ds = c.shape[2]
    sign = torch.ones(ds, device=c.device)
    sign[1::2] = -1.0
    for m, k in empty.cpu().numpy().tolist():           # nonzero() is row-major: m ascending, k ascending
        j = int(np.argmax(cnt[m]))
        if cnt[m, j] < 2:
            continue
        src = c[m, j].clone()
        c[m, k] = src * (1.0 + SPLIT_EPS * sign)
        c[m, j] = src * (1.0 - SPLIT_EPS * sign)
        cnt[m, k] = cnt[m, j] // 2
        cnt[m, j] -= cnt[m, k]
    return c


def train_pq(x, M, K=256, niter=25, seed=1234, init=None, max_points_per_centroid=PQ_MAX_POINTS_PER_CENTROID):
    """ProductQuantizer::train on the device.  x (n, D) CUDA fp32 -> (centroids (M,K,D/M), [objective per
    iteration])."""
    ops._require_cuda(x, "x")
    x = x.float().contiguous()
    if x.shape[1] % M:
        raise ValueError(f"dimension {x.shape[1]} is not a multiple of M={M}")
    x = _subsample(x, K * max_points_per_centroid, seed)
    c = initial_centroids(x, M, K, seed) if init is None else init.detach().float().contiguous()
    objs = []
    for _ in range(int(niter)):
        c, obj, _, _ = lloyd_step(x, c)
        objs.append(obj)
    return c, [float(o) for o in torch.stack(objs).cpu()] if objs else []


def random_rotation(d, seed, device):
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    a = torch.randn((d, d), generator=g, dtype=torch.float64)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r)).unsqueeze(0)     # unique factorisation: R with a positive diagonal
    return q.float().to(device)


def procrustes(x, recons):
    """orthogonal A (d,d) minimising ||x A^T - recons||_F:  U S V^T = svd(recons^T x), A = U V^T"""
    s = (recons.double().t() @ x.double())
    u, _, vt = torch.linalg.svd(s)
    return (u @ vt).float()


def train_opq(x, M, K=256, niter=50, niter_pq=4, niter_pq_0=40, seed=1234, init_rotation=None):
    """OPQMatrix::train on the device: -> (rotation (D,D), centroids of the inner PQ, [pq error per
    alternation])."""
    ops._require_cuda(x, "x")
    x = _subsample(x.float().contiguous(), OPQ_MAX_TRAIN_POINTS, seed)
    d = x.shape[1]
    A = random_rotation(d, seed, x.device) if init_rotation is None else init_rotation.float().to(x.device)
    c, errs = None, []
    for it in range(int(niter)):
        xproj = (x @ A.t()).contiguous()
        c, _ = train_pq(xproj, M, K, niter_pq_0 if it == 0 else niter_pq, seed, init=c, max_points_per_centroid=1000)
        codes = ops.nn_assign(xproj, c)
        recons = ops.decode_forward(codes, c)
        errs.append(((recons - xproj) ** 2).sum(-1).mean())
        A = procrustes(x, recons)
    return A, c, [float(e) for e in torch.stack(errs).cpu()] if errs else []


@torch.no_grad()
def warmup_from_embeds(corpus_embeds: np.ndarray, repconc, opq_niter=50, pq_niter=25, seed=1234, chunk=1 << 20):
    """Drop-in for run_warmup.warmup_from_embeds (run_warmup.py:85-132): learns `rotation` (OPQ) and
    `centroids` (PQ on the rotated embeddings), stores them in the module (`:122-126`), normalises the
    centroids for METRIC_CENTROID_COS (`:128-129`) and returns (repconc, index) where `index` holds the PQ
    codes of every corpus vector -- what the reference writes with faiss.write_index (`:187`)."""
    MCQ_M, MCQ_K = repconc.config.MCQ_M, repconc.config.MCQ_K
    assert MCQ_K == 256, "256 is a standard setting for K. "
    dev = repconc.centroids.device
    if dev.type != "cuda":
        raise _lib.RepconcLibraryError("warmup_from_embeds: move the module to a CUDA device first "
                                       "(the B200 path has no CPU implementation)")
    n, D = corpus_embeds.shape
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    idx = torch.randperm(n, generator=g)[:OPQ_MAX_TRAIN_POINTS].numpy() if n > OPQ_MAX_TRAIN_POINTS else np.arange(n)
    xtrain = torch.from_numpy(np.ascontiguousarray(corpus_embeds[np.sort(idx)], dtype=np.float32)).to(dev)
    rotation, _, errs = train_opq(xtrain, MCQ_M, MCQ_K, niter=opq_niter, seed=seed)
    logger.info("OPQ: pq error %.6g -> %.6g over %d alternations", errs[0] if errs else float("nan"),
                errs[-1] if errs else float("nan"), len(errs))
    xrot = (xtrain @ rotation.t()).contiguous()
    centroids, objs = train_pq(xrot, MCQ_M, MCQ_K, niter=pq_niter, seed=seed)
    logger.info("PQ: objective %.6g -> %.6g over %d iterations", objs[0] if objs else float("nan"),
                objs[-1] if objs else float("nan"), len(objs))
    repconc.rotation.copy_(rotation)
    repconc.centroids.data.copy_(centroids)
    if repconc.config.similarity_metric == "METRIC_CENTROID_COS":
        repconc.normalize_centrodis()
    # index.add(corpus_embeds): rotate + NN-assign every vector, uint8 codes straight from the kernel
    index = IndexPQ(D, MCQ_M, 8, METRIC_INNER_PRODUCT)
    index.is_trained = True
    copy_array_to_vector(repconc.centroids.data.detach().cpu().numpy().ravel(), index.pq.centroids)
    for lo in range(0, n, chunk):
        xb = torch.from_numpy(np.ascontiguousarray(corpus_embeds[lo:lo + chunk], dtype=np.float32)).to(dev)
        codes = ops.nn_assign((xb @ rotation.t()).contiguous(), centroids, uint8=True)
        index.add_codes(codes.cpu().numpy())
    return repconc, index

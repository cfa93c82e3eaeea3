"""numpy restatement of the OPQ / PQ warm-up (src/repconc/train/run_warmup.py:85-132 -> Faiss 1.7.1
`OPQMatrix::train`, `ProductQuantizer::train`, `Clustering::train` / `split_clusters`).

TEST INFRASTRUCTURE ONLY (imported by tests/ alone).  PARITY UNPINNED w.r.t. Faiss: faiss-gpu == 1.7.1
(setup.py:21) is not installable in this image, the reference holds no fixture for this step, and Faiss'
own random streams / BLAS summation order cannot be reproduced -- this file restates the published
algorithm step by step so the CUDA path can be checked state by state:
  * Clustering::train        : K initial centroids = training points picked by a seeded permutation; per
                               iteration assign (L2 argmin, first minimum), centroid = mean of its points
                               (an empty cluster keeps its position, then is re-seeded by split_clusters)
  * split_clusters           : empty cluster <- copy of a large cluster, coordinates scaled by 1 +- 1/1024
                               alternately, donor scaled the other way, donor size halved.  (Faiss draws the
                               donor at random in proportion to size; the restatement takes the largest.)
  * OPQMatrix::train         : rotate, train PQ (hot start after the first alternation), encode + decode,
                               A <- U V^T with U S V^T = svd(recons^T x)
The NN assignment uses the C oracle (oracle/repconc_oracle.c: nn_assign, ATen summation order), so codes
are comparable bit for bit with rc_nn_assign.
"""
import numpy as np

from . import oracle as O

SPLIT_EPS = 1.0 / 1024.0


def lloyd_step(x, c):
    """-> (new centroids fp32 (M,K,ds), objective of the input centroids (fp64), codes (B,M), counts (M,K))"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    M, K, ds = c.shape
    B = x.shape[0]
    codes = O.nn_assign(x, c)                                   # (B, M) int64
    xs = x.reshape(B, M, ds)
    q = c[np.arange(M)[None, :], codes]                          # (B, M, ds)
    obj = float((((q.astype(np.float64) - xs) ** 2).sum((1, 2))).mean())
    sums = np.zeros((M, K, ds), dtype=np.float64)
    counts = np.zeros((M, K), dtype=np.int64)
    for m in range(M):
        np.add.at(sums[m], codes[:, m], xs[:, m, :].astype(np.float64))
        counts[m] = np.bincount(codes[:, m], minlength=K)
    new_c = c.copy()
    nz = counts > 0
    new_c[nz] = (sums[nz] / counts[nz][:, None]).astype(np.float32)
    return split_empty_clusters(new_c, counts), obj, codes, counts


def split_empty_clusters(c, counts):
    c = c.copy()
    cnt = counts.astype(np.int64).copy()
    M, K, ds = c.shape
    sign = np.ones(ds, dtype=np.float32)
    sign[1::2] = -1.0
    for m in range(M):
        for k in range(K):
            if counts[m, k] != 0:
                continue
            j = int(np.argmax(cnt[m]))
            if cnt[m, j] < 2:
                continue
            src = c[m, j].copy()
            c[m, k] = (src * (np.float32(1.0) + np.float32(SPLIT_EPS) * sign)).astype(np.float32)
            c[m, j] = (src * (np.float32(1.0) - np.float32(SPLIT_EPS) * sign)).astype(np.float32)
            cnt[m, k] = cnt[m, j] // 2
            cnt[m, j] -= cnt[m, k]
    return c


def train_pq(x, c0, niter):
    c = np.ascontiguousarray(c0, dtype=np.float32)
    objs = []
    for _ in range(niter):
        c, obj, _, _ = lloyd_step(x, c)
        objs.append(obj)
    return c, objs


def decode(codes, c):
    M = c.shape[0]
    return c[np.arange(M)[None, :], codes].reshape(codes.shape[0], -1)


def procrustes(x, recons):
    s = recons.astype(np.float64).T @ x.astype(np.float64)
    u, _, vt = np.linalg.svd(s)
    return (u @ vt).astype(np.float32)


def opq_alternation(x, A, c, niter_pq):
    """one OPQ alternation from (rotation A, centroids c): -> (A', c', pq error)"""
    xproj = np.ascontiguousarray((x.astype(np.float32) @ A.T.astype(np.float32)), dtype=np.float32)
    c, _ = train_pq(xproj, c, niter_pq)
    codes = O.nn_assign(xproj, c)
    recons = decode(codes, c)
    err = float(((recons.astype(np.float64) - xproj) ** 2).sum(-1).mean())
    return procrustes(x, recons), c, err

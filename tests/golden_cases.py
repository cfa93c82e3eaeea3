"""Seeded synthetic inputs shared by tools/gen_golden.py (which runs the REFERENCE on them,
in the build container) and by the parity tests (which re-create them anywhere).

Inputs are re-generated from the seed, never stored; every fixture carries a sha256 of the
generated inputs so a drifting RNG is reported as such instead of as a parity failure.
"""
import hashlib

import numpy as np

# name -> dict(D, M, K, B, eps, iters, seed, centroids)
#   centroids: "randn" (RepCONC default init, modeling_repconc.py:41) or "sample"
#   (trained-like: sub-vectors of the batch itself)
ASSIGN_CASES = {
    "ds16_b512":   dict(D=128, M=8,  K=256, B=512,  eps=0.003, iters=50,  seed=11, centroids="randn"),
    "ds24_b384":   dict(D=192, M=8,  K=256, B=384,  eps=0.003, iters=100, seed=12, centroids="randn"),
    "ds8_b256":    dict(D=64,  M=8,  K=256, B=256,  eps=0.003, iters=50,  seed=13, centroids="sample"),
    "ds12_b300":   dict(D=96,  M=8,  K=256, B=300,  eps=0.05,  iters=10,  seed=14, centroids="randn"),
    "ds5_k64":     dict(D=20,  M=4,  K=64,  B=128,  eps=0.01,  iters=20,  seed=15, centroids="randn"),
    "ds48_b128":   dict(D=96,  M=2,  K=256, B=128,  eps=0.003, iters=30,  seed=16, centroids="randn"),
    "m48_b1024":   dict(D=768, M=48, K=256, B=1024, eps=0.003, iters=50,  seed=17, centroids="randn"),
    "m96_b512":    dict(D=768, M=96, K=256, B=512,  eps=0.003, iters=50,  seed=18, centroids="sample"),
    "m32_b512":    dict(D=768, M=32, K=256, B=512,  eps=0.003, iters=50,  seed=19, centroids="randn"),
    "m64_b256":    dict(D=768, M=64, K=256, B=256,  eps=0.003, iters=100, seed=20, centroids="randn"),
    "b1":          dict(D=128, M=8,  K=256, B=1,    eps=0.003, iters=5,   seed=21, centroids="randn"),
}
# BASELINE-size batches (configs[2] and the per-rank slab of configs[4]): the reference needs ~20-60 s and
# 13-26 GB for each in the build container; only codes, extrema, table hashes and the top-1/top-2 gap (fp16)
# are stored
ASSIGN_BIG_CASES = {
    "m48_b8192":   dict(D=768, M=48, K=256, B=8192, eps=0.003, iters=50, seed=61, centroids="randn"),
    "m96_b8192":   dict(D=768, M=96, K=256, B=8192, eps=0.003, iters=50, seed=62, centroids="randn"),
}
# two-rank run of the reference itself (gloo): each rank holds B/2 rows
DIST_CASES = {
    "dist2_ds16":  dict(D=128, M=8, K=256, B=512, eps=0.003, iters=50, seed=31, centroids="randn", world=2),
}
MSE_CASES = {
    "mse_n64":     dict(D=128, M=8,  K=256, n=64,  w=1e-4, scale=1.0,     seed=41),
    "mse_amp":     dict(D=768, M=48, K=256, n=64,  w=1e-4, scale=65536.0, seed=42),
}
# forward() after the encoder: rotation + (COS) normalisation + NN assign   (modeling_repconc.py:98-103)
ENCODE_CASES = {
    "enc_ip_eye":   dict(D=768, M=48, K=256, B=512, metric="METRIC_IP",           rotation="eye",    seed=61),
    "enc_ip_rot":   dict(D=768, M=48, K=256, B=512, metric="METRIC_IP",           rotation="random", seed=62),
    "enc_cos_rot":  dict(D=768, M=48, K=256, B=512, metric="METRIC_CENTROID_COS", rotation="random", seed=63),
    "enc_cos_m32":  dict(D=768, M=32, K=256, B=300, metric="METRIC_CENTROID_COS", rotation="random", seed=64),
    "enc_ip_m96":   dict(D=768, M=96, K=256, B=300, metric="METRIC_IP",           rotation="random", seed=65),
}
ADC_CASES = {
    "adc_m8":      dict(D=128, M=8,  K=256, N=10000, nq=64, ks=(10, 1000), seed=51),
    "adc_m48":     dict(D=768, M=48, K=256, N=4000,  nq=16, ks=(10, 100),  seed=52),
}


def _rng(seed):
    return np.random.default_rng(seed)


def assign_inputs(case):
    r = _rng(case["seed"])
    D, M, K, B = case["D"], case["M"], case["K"], case["B"]
    ds = D // M
    x = r.standard_normal((B, D), dtype=np.float32)
    if case["centroids"] == "randn":
        c = r.standard_normal((M, K, ds), dtype=np.float32)
    else:
        pool = r.standard_normal((K, D), dtype=np.float32)
        pool[: min(B, K // 2)] = x[: min(B, K // 2)]
        c = np.ascontiguousarray(pool.reshape(K, M, ds).transpose(1, 0, 2))
    return x, c


def encode_inputs(case):
    """pooled encoder outputs, rotation (orthogonal, or identity) and centroids (normalised rows for the COS metric,
    as normalize_centrodis keeps them)"""
    r = _rng(case["seed"])
    D, M, K, B = case["D"], case["M"], case["K"], case["B"]
    ds = D // M
    x = r.standard_normal((B, D), dtype=np.float32)
    c = r.standard_normal((M, K, ds), dtype=np.float32)
    if case["metric"] == "METRIC_CENTROID_COS":
        c = (c / np.linalg.norm(c, axis=-1, keepdims=True)).astype(np.float32)
    if case["rotation"] == "eye":
        rot = np.eye(D, dtype=np.float32)
    else:
        rot = np.linalg.qr(r.standard_normal((D, D)))[0].astype(np.float32)
    return x, np.ascontiguousarray(rot), np.ascontiguousarray(c)


def mse_inputs(case):
    r = _rng(case["seed"])
    D, M, K, n = case["D"], case["M"], case["K"], case["n"]
    ds = D // M
    x = r.standard_normal((n, D), dtype=np.float32)
    c = r.standard_normal((M, K, ds), dtype=np.float32)
    g = (r.standard_normal((n, D), dtype=np.float32) / np.float32(n)).astype(np.float32)
    codes = r.integers(0, K, size=(n, M), dtype=np.int64)
    codes[1] = codes[0]  # force collisions in the scatter-add
    return x, c, g, codes


def adc_inputs(case):
    r = _rng(case["seed"])
    D, M, K, N, nq = case["D"], case["M"], case["K"], case["N"], case["nq"]
    ds = D // M
    c = r.standard_normal((M, K, ds), dtype=np.float32)
    codes = r.integers(0, K, size=(N, M), dtype=np.uint8)
    codes[N // 2] = codes[N // 3]  # an exact duplicate document -> exact score tie
    q = r.standard_normal((nq, D), dtype=np.float32)
    return q, c, codes


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()

"""ctypes front-end of oracle/repconc_oracle.c (the C restatement of the reference).

TEST INFRASTRUCTURE ONLY (see repconc_oracle.c header).  Each wrapper names the
reference lines it follows; arrays are numpy, C-contiguous.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librepconc_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "repconc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_sum_torch_order.restype = ctypes.c_float
        _lib.orc_sum_torch_order.argtypes = [_f32p, ctypes.c_int64]
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def sum_torch_order(v):
    v = _f32(v)
    return np.float32(lib().orc_sum_torch_order(_p(v, _f32p), ctypes.c_int64(v.size)))


def dist_table(x, c):
    """modeling_repconc.py:47-50 -> (M,B,K) fp32, bit-identical to the reference on CPU."""
    x, c = _f32(x), _f32(c)
    M, K, ds = c.shape
    B = x.shape[0]
    assert x.shape[1] == M * ds
    out = np.empty((M, B, K), np.float32)
    lib().orc_dist_table(_p(x, _f32p), _p(c, _f32p), ctypes.c_int64(B), M, K, ds, _p(out, _f32p))
    return out


def nn_assign(x, c):
    """modeling_repconc.py:50-52,66 (use_constraint=False) -> (B,M) int64."""
    x, c = _f32(x), _f32(c)
    M, K, ds = c.shape
    B = x.shape[0]
    codes = np.empty((B, M), np.int64)
    lib().orc_nn_assign(_p(x, _f32p), _p(c, _f32p), ctypes.c_int64(B), M, K, ds, _p(codes, _i64p))
    return codes


def table_minmax(table):
    """modeling_repconc.py:76-77."""
    table = _f32(table)
    M, B, K = table.shape
    mx, mn = np.empty(M, np.float32), np.empty(M, np.float32)
    lib().orc_table_minmax(_p(table, _f32p), M, ctypes.c_int64(B), K, _p(mx, _f32p), _p(mn, _f32p))
    return mx, mn


def center_table(table, mx=None, mn=None):
    """RepCONC.center_distance_for_constraint, modeling_repconc.py:73-85."""
    table = _f32(table).copy()
    M, B, K = table.shape
    if mx is None:
        mx, mn = table_minmax(table)
    rc = lib().orc_center_table(_p(table, _f32p), M, ctypes.c_int64(B), K, _p(_f32(mx), _f32p),
                                _p(_f32(mn), _f32p))
    if rc != 0:
        raise AssertionError("amplitude > 0 (modeling_repconc.py:83)")
    return table


def sinkhorn(out, eps, iters):
    """sinkhorn_algorithm, modeling_repconc.py:137-165.  out: (M,K,B) fp64 -> Q (M,K,B) fp64.
    Distributed runs are emulated by passing the concatenated global batch."""
    Q = np.array(out, dtype=np.float64, order="C", copy=True)
    M, K, B = Q.shape
    lib().orc_sinkhorn(_p(Q, _f64p), M, K, ctypes.c_int64(B), ctypes.c_double(eps), int(iters))
    return Q


def constrained_assign(x, c, eps, iters, ext_max=None, ext_min=None, return_q=False):
    """RepCONC.quantize with use_constraint=True, modeling_repconc.py:47-66.
    Returns dict(codes (B,M) int64, nonfinite bool, max (M,), min (M,), [Q (M,K,B)])."""
    x, c = _f32(x), _f32(c)
    M, K, ds = c.shape
    B = x.shape[0]
    codes = np.empty((B, M), np.int64)
    bad = ctypes.c_int(0)
    mx, mn = np.empty(M, np.float32), np.empty(M, np.float32)
    Q = np.empty((M, K, B), np.float64) if return_q else None
    em = _f32(ext_max) if ext_max is not None else None
    en = _f32(ext_min) if ext_min is not None else None
    rc = lib().orc_constrained_assign(
        _p(x, _f32p), _p(c, _f32p), ctypes.c_int64(B), M, K, ds, ctypes.c_double(eps), int(iters),
        _p(em, _f32p), _p(en, _f32p), _p(codes, _i64p), ctypes.byref(bad), _p(mx, _f32p),
        _p(mn, _f32p), _p(Q, _f64p))
    if rc == -1:
        raise AssertionError("amplitude > 0 (modeling_repconc.py:83)")
    if rc != 0:
        raise MemoryError("oracle out of memory")
    res = dict(codes=codes, nonfinite=bool(bad.value), max=mx, min=mn)
    if return_q:
        res["Q"] = Q
    return res


def decode(codes, c):
    """decode(), modeling_repconc.py:168-184 -> (B, D) fp32."""
    c = _f32(c)
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    M, K, ds = c.shape
    B = codes.shape[0]
    out = np.empty((B, M * ds), np.float32)
    lib().orc_decode(_p(codes, _i64p), _p(c, _f32p), ctypes.c_int64(B), M, K, ds, _p(out, _f32p))
    return out


def mse_surrogate(x, q, g, codes, K, w, scale=1.0):
    """finetune_repconc.py:367-374,389-396 closed form.
    Returns dict(mse, surrogate, grad_x, grad_q, grad_c)."""
    x, q, g = _f32(x), _f32(q), _f32(g)
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    n, D = x.shape
    M = codes.shape[1]
    ds = D // M
    mse, sur = ctypes.c_float(0), ctypes.c_float(0)
    gx, gq = np.empty_like(x), np.empty_like(x)
    gc = np.empty((M, K, ds), np.float32)
    lib().orc_mse_surrogate(_p(x, _f32p), _p(q, _f32p), _p(g, _f32p), _p(codes, _i64p),
                            ctypes.c_int64(n), M, K, ds, ctypes.c_float(w), ctypes.c_float(scale),
                            ctypes.byref(mse), ctypes.byref(sur), _p(gx, _f32p), _p(gq, _f32p),
                            _p(gc, _f32p))
    return dict(mse=np.float32(mse.value), surrogate=np.float32(sur.value), grad_x=gx, grad_q=gq,
                grad_c=gc)


def adc_lut(queries, c):
    """Faiss IndexPQ inner-product table as used at evaluate_repconc.py:81-85,182 -> (nq,M,K)."""
    queries, c = _f32(queries), _f32(c)
    M, K, ds = c.shape
    nq = queries.shape[0]
    lut = np.empty((nq, M, K), np.float32)
    lib().orc_adc_lut(_p(queries, _f32p), _p(c, _f32p), ctypes.c_int64(nq), M, K, ds,
                      _p(lut, _f32p))
    return lut


def adc_search(queries, c, codes, k, id_offset=0):
    """index.search(query_embeds, topk), evaluate_repconc.py:182 -> (scores (nq,k) f32, ids i64)."""
    queries, c = _f32(queries), _f32(c)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    M, K, ds = c.shape
    nq, N = queries.shape[0], codes.shape[0]
    assert codes.shape[1] == M
    scores = np.empty((nq, k), np.float32)
    ids = np.empty((nq, k), np.int64)
    rc = lib().orc_adc_search(_p(queries, _f32p), _p(c, _f32p), _p(codes, _u8p),
                              ctypes.c_int64(nq), ctypes.c_int64(N), M, K, ds, ctypes.c_int64(k),
                              ctypes.c_int64(id_offset), _p(scores, _f32p), _p(ids, _i64p))
    if rc != 0:
        raise MemoryError("oracle out of memory")
    return scores, ids


def topk_merge(scores_in, ids_in):
    """k-way merge of per-shard sorted top-k lists: (W,nq,k) -> (nq,k)."""
    scores_in = _f32(scores_in)
    ids_in = np.ascontiguousarray(ids_in, dtype=np.int64)
    W, nq, k = scores_in.shape
    scores = np.empty((nq, k), np.float32)
    ids = np.empty((nq, k), np.int64)
    lib().orc_topk_merge(_p(scores_in, _f32p), _p(ids_in, _i64p), W, ctypes.c_int64(nq),
                         ctypes.c_int64(k), _p(scores, _f32p), _p(ids, _i64p))
    return scores, ids


def mrr_at_k(run_ids, rel_ids, k=10):
    """MRR@k as eval_utils.py:136-141,182-190 computes it for one relevant doc per query:
    truncate the ranking to k, reciprocal rank of the first relevant hit, mean, round 5 dp.
    run_ids: (nq, >=k) ranked doc ids; rel_ids: (nq,) the relevant doc of each query."""
    run_ids = np.asarray(run_ids)[:, :k]
    hit = run_ids == np.asarray(rel_ids)[:, None]
    rank = np.where(hit.any(1), hit.argmax(1) + 1, 0)
    rr = np.where(rank > 0, 1.0 / np.maximum(rank, 1), 0.0)
    return round(float(rr.mean()), 5)

"""numpy restatement of the same path (cross-check of the C oracle on small cases).

TEST INFRASTRUCTURE ONLY.  Slow and literal on purpose.
"""
import numpy as np

_V, _ILP, _LEVELS = 8, 4, 4


def _f32(a):
    return np.asarray(a, dtype=np.float32)


def _multi_row_sum(rows):
    """ATen SumKernel.cpp multi_row_sum; rows: list of [ILP arrays]."""
    size = len(rows)
    ceil_log2 = 0 if size <= 1 else int(np.ceil(np.log2(size)))
    level_power = max(4, ceil_log2 // _LEVELS)
    level_step = 1 << level_power
    level_mask = level_step - 1
    zero = np.zeros_like(rows[0][0])
    acc = [[zero.copy() for _ in range(_ILP)] for _ in range(_LEVELS)]
    i = 0
    while i + level_step <= size:
        for _ in range(level_step):
            for k in range(_ILP):
                acc[0][k] = _f32(acc[0][k] + rows[i][k])
            i += 1
        for j in range(1, _LEVELS):
            for k in range(_ILP):
                acc[j][k] = _f32(acc[j][k] + acc[j - 1][k])
                acc[j - 1][k] = zero.copy()
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size:
        for k in range(_ILP):
            acc[0][k] = _f32(acc[0][k] + rows[i][k])
        i += 1
    for j in range(1, _LEVELS):
        for k in range(_ILP):
            acc[0][k] = _f32(acc[0][k] + acc[j][k])
    return acc[0]


def _row_sum(items):
    size = len(items)
    size_ilp = size // _ILP
    zero = np.zeros_like(items[0])
    if size_ilp:
        ps = _multi_row_sum([[items[i * _ILP + k] for k in range(_ILP)] for i in range(size_ilp)])
    else:
        ps = [zero.copy() for _ in range(_ILP)]
    for i in range(size_ilp * _ILP, size):
        ps[0] = _f32(ps[0] + items[i])
    for k in range(1, _ILP):
        ps[0] = _f32(ps[0] + ps[k])
    return ps[0]


def sum_last_torch_order(a):
    """fp32 sum over the last (contiguous) axis in ATen-CPU order (SumKernel.cpp)."""
    a = _f32(a)
    n = a.shape[-1]
    if n >= _V:
        nvec = n // _V
        vacc = _row_sum([a[..., i * _V:(i + 1) * _V] for i in range(nvec)])
        fin = np.zeros(a.shape[:-1], np.float32)
        for k in range(nvec * _V, n):
            fin = _f32(fin + a[..., k])
        for k in range(_V):
            fin = _f32(fin + vacc[..., k])
        return fin
    return _row_sum([a[..., k] for k in range(n)])


def dist_table(x, c):
    """modeling_repconc.py:50 -> (M,B,K)."""
    x, c = _f32(x), _f32(c)
    M, K, ds = c.shape
    B = x.shape[0]
    diff = _f32(x.reshape(B, M, 1, ds).transpose(1, 0, 2, 3) - c[:, None, :, :])
    return sum_last_torch_order(_f32(diff * diff))


def center(table, mx=None, mn=None):
    """modeling_repconc.py:73-85."""
    table = _f32(table)
    if mx is None:
        mx = table.max(-1).max(-1)
        mn = table.min(-1).min(-1)
    middle = _f32((mx + mn) / np.float32(2))
    amplitude = _f32(_f32(mx - middle) + np.float32(1e-5))
    assert np.all(amplitude > 0)
    return _f32(_f32(table - middle[:, None, None]) / amplitude[:, None, None])


def sinkhorn(out, eps, iters):
    """modeling_repconc.py:137-165, literal."""
    Q = np.exp(np.asarray(out, np.float64) / eps)
    M, K, B = Q.shape
    Q /= Q.sum(-1, keepdims=True).sum(-2, keepdims=True)
    for _ in range(iters):
        Q /= Q.sum(2, keepdims=True)
        Q /= K
        Q /= Q.sum(1, keepdims=True)
        Q /= B
    Q *= B
    return Q


def constrained_assign(x, c, eps, iters):
    """modeling_repconc.py:47-66 with use_constraint=True -> (B,M) int64."""
    d = center(dist_table(x, c)).astype(np.float64)
    Q = sinkhorn(-d.transpose(0, 2, 1), eps, iters).transpose(0, 2, 1)
    return np.argmax(Q, -1).T.astype(np.int64)


def nn_assign(x, c):
    return np.argmin(dist_table(x, c), -1).T.astype(np.int64)


def decode(codes, c):
    """modeling_repconc.py:168-184."""
    M = codes.shape[1]
    first = np.tile(np.arange(M), len(codes))
    return np.asarray(c)[first, np.asarray(codes).reshape(-1)].reshape(len(codes), -1)


def adc_scores(queries, c, codes):
    """sum_m <q_m, c[m, code[n,m]]>, fp32, m ascending, j ascending, no FMA -> (nq,N)."""
    queries, c = _f32(queries), _f32(c)
    M, K, ds = c.shape
    nq = queries.shape[0]
    qv = queries.reshape(nq, M, 1, ds)
    lut = np.zeros((nq, M, K), np.float32)
    for j in range(ds):
        lut = _f32(lut + _f32(qv[..., j] * c[None, :, :, j]))
    s = np.zeros((nq, codes.shape[0]), np.float32)
    for m in range(M):
        s = _f32(s + lut[:, m, :][:, codes[:, m]])
    return s


def encode_assign(pooled, rotation, c, normalize, nn=nn_assign):
    """RepCONC.forward after the encoder with use_constraint = False (modeling_repconc.py:98-103):
    rotated = pooled @ rotation.T (fp32), optional per-sub-vector L2 normalisation with F.normalize's eps = 1e-12
    (:99-100), codes = argmin of the squared distances (:51-52).  `nn` = the NN-assign restatement to use
    (this module's numpy one, or the C oracle's).  Returns (rotated (B, D) fp32, codes (B, M) int64)."""
    pooled, rotation, c = _f32(pooled), _f32(rotation), _f32(c)
    M = c.shape[0]
    y = pooled @ rotation.T
    if normalize:
        ys = y.reshape(len(y), M, -1)
        n = np.sqrt((ys * ys).sum(-1, keepdims=True, dtype=np.float32))
        y = (ys / np.maximum(n, np.float32(1e-12))).reshape(len(y), -1).astype(np.float32)
    return y, nn(np.ascontiguousarray(y), c)

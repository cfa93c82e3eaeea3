"""Probe: ATen CPU fp32 sum order model vs torch (run in the build container).
Prints mismatch counts (all 0 expected) for reduced sizes 1..768; the model is restated in
oracle/repconc_oracle.c:orc_sum_torch_order and in repconc_b200/csrc (table kernels)."""
import torch, numpy as np
def f32(a): return a.astype(np.float32)
def multi_row_sum(rows, nrows):
    # rows: list of "row" items; each row is list of nrows arrays. Returns list of nrows accumulators.
    size = len(rows)
    num_levels = 4
    ceil_log2 = 0 if size <= 1 else int(np.ceil(np.log2(size)))
    level_power = max(4, ceil_log2 // num_levels)
    level_step = 1 << level_power
    level_mask = level_step - 1
    z = np.zeros_like(rows[0][0]) if size else None
    acc = [[None]*nrows for _ in range(num_levels)]
    def add(a, b):
        if a is None: return f32(np.float32(0) + b)
        return f32(a + b)
    i = 0
    while i + level_step <= size:
        for j in range(level_step):
            for k in range(nrows): acc[0][k] = add(acc[0][k], rows[i][k])
            i += 1
        for j in range(1, num_levels):
            for k in range(nrows):
                acc[j][k] = add(acc[j][k], acc[j-1][k] if acc[j-1][k] is not None else np.float32(0))
                acc[j-1][k] = None
            mask = level_mask << (j * level_power)
            if (i & mask) != 0: break
    while i < size:
        for k in range(nrows): acc[0][k] = add(acc[0][k], rows[i][k])
        i += 1
    for j in range(1, num_levels):
        for k in range(nrows):
            if acc[j][k] is not None: acc[0][k] = add(acc[0][k], acc[j][k])
    return acc[0]
def row_sum(items):
    # items: list of arrays (vectors or scalars), torch row_sum with ilp 4
    size = len(items); ilp = 4; size_ilp = size // ilp
    rows = [[items[i*ilp+k] for k in range(ilp)] for i in range(size_ilp)]
    ps = multi_row_sum(rows, ilp) if size_ilp else [None]*ilp
    for i in range(size_ilp*ilp, size):
        ps[0] = f32(ps[0] + items[i]) if ps[0] is not None else f32(items[i])
    for k in range(1, ilp):
        if ps[k] is not None: ps[0] = f32(ps[0] + ps[k]) if ps[0] is not None else ps[k]
    return ps[0]
def torch_inner_sum(sq):
    ds = sq.shape[-1]; V = 8
    if ds >= V:
        nvec = ds // V
        vecs = [sq[..., i*V:(i+1)*V] for i in range(nvec)]
        vacc = row_sum(vecs)  # (..., 8)
        fin = np.zeros(sq.shape[:-1], np.float32)
        for k in range(nvec*V, ds): fin = f32(fin + sq[..., k])
        for k in range(V): fin = f32(fin + vacc[..., k])
        return fin
    else:
        r = row_sum([sq[..., k] for k in range(ds)])
        return r
torch.manual_seed(0)
for ds in (1,2,3,4,6,7,8,12,16,24,32,40,48,64,96,192,384,768):
    for shape in ((3,5,7),(2,33,256),(1,1,1)):
        sq = torch.randn(*shape, ds)**2
        ref = sq.sum(-1).numpy()
        got = torch_inner_sum(sq.numpy())
        print(ds, shape, 'mismatch', int((ref!=got).sum()), 'of', ref.size)

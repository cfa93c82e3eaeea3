"""B200-native drop-in for the PQ-index / search half of
`repconc.models.repconc.evaluate_repconc` (evaluate_repconc.py:78-135,180-206).

Same function names, argument meaning and return values as the reference; the index objects are
the Faiss-free ones of `repconc_b200.faiss_compat`, the search is `rc_adc_search` (hand-written
sm_100a kernels).  The HF-Trainer based encoders of the same reference file (`encode_corpus`,
`encode_query`, `RepCONCEvaluater`) are callers of this path and stay in the reference.
"""
import math
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, faiss_compat as faiss, ops

try:
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it


def initialize_index(model):
    """evaluate_repconc.py:78-86: empty IndexPQ(D, M, 8, IP) carrying the model's centroids."""
    D, M = model.config.hidden_size, model.config.MCQ_M
    assert model.config.MCQ_K == 256
    index = faiss.IndexPQ(D, M, 8, faiss.METRIC_INNER_PRODUCT)
    index.is_trained = True
    # set centroid values
    centroids = model.centroids.data.detach().cpu().numpy()
    faiss.copy_array_to_vector(centroids.ravel(), index.pq.centroids)
    return index


def add_docs(index, new_codes: np.ndarray):
    """evaluate_repconc.py:89-98: append (n, M) uint8 codes.  (The reference copies the whole code
    vector on every call; here it is one amortised append.)"""
    M = index.pq.code_size
    new_n = len(new_codes)
    assert new_codes.shape == (new_n, M)
    index.add_codes(new_codes)


def from_pq_to_ivfpq(indexpq):
    """evaluate_repconc.py:101-118: view the PQ index as a one-list IVFPQ.  With nlist = 1 and a zero
    coarse centroid the IVFPQ inner-product search is the plain PQ scan, so this shares the data."""
    ivfpq = faiss.IndexIVFPQ(indexpq.pq.d, indexpq.pq.M, indexpq.pq.nbits, indexpq.metric_type)
    ivfpq.pq = indexpq.pq
    ivfpq.is_trained = True
    ivfpq.codes = indexpq.codes
    ivfpq.ntotal = indexpq.ntotal
    return ivfpq


def load_index_to_gpu(index, single_gpu_id: Optional[int] = None):
    """evaluate_repconc.py:121-135: make the index device-resident (one H2D copy of N*M bytes).
    `single_gpu_id` given: that device holds the whole index (reference :123-128).
    `single_gpu_id` None: the reference clones the index onto ALL visible GPUs from this one process
    (`index_cpu_to_all_gpus`, :130-134 -- replicas, queries split between them).  Here the corpus is SHARDED over
    the visible devices instead (each device scans 1/n of the codes for every query, one host thread per device,
    per-shard top-k merged on the first device): same results, n times the scan rate, and it is reached by the
    unchanged evaluator, which searches from the main process only (run_repconc_eval.py:93-100,155).
    RC_ADC_DEVICES=n limits the number of devices used."""
    if isinstance(index, (faiss.GpuIndexPQ, faiss.MultiGpuIndexPQ)):
        return index
    if single_gpu_id is None:
        import os
        n = torch.cuda.device_count()
        n = min(n, int(os.environ.get("RC_ADC_DEVICES", n)))
        if n > 1 and index.ntotal >= n:
            return faiss.MultiGpuIndexPQ.from_host(index, list(range(n)))
        return faiss.GpuIndexPQ.from_host(index, torch.device("cuda", torch.cuda.current_device()))
    return faiss.GpuIndexPQ.from_host(index, torch.device("cuda", single_gpu_id))


def search(query_ids: np.ndarray, query_embeds: np.ndarray, corpus_ids: np.ndarray, index, topk: int):
    """evaluate_repconc.py:180-185."""
    if (isinstance(index, _DEVICE_INDEXES) and isinstance(query_embeds, np.ndarray)
            and isinstance(corpus_ids, np.ndarray) and corpus_ids.dtype == np.int64):
        # position -> corpus id on the device, before the copy back (same values as :183)
        topk_scores, topk_ids = index.search(query_embeds, topk, corpus_ids=corpus_ids)
        if isinstance(index, (ShardedSearcher, ReplicatedSearcher)) and index.rank != 0:
            return topk_scores, topk_ids                     # empty on the ranks that do not receive the result
    else:
        topk_scores, topk_idx = index.search(query_embeds, topk)
        topk_ids = corpus_ids[topk_idx]      # == np.vstack([corpus_ids[x] for x in topk_idx]) (:183)
    assert len(query_ids) == len(topk_scores) == len(topk_ids)
    return topk_scores, topk_ids


def batch_search(query_ids: np.ndarray, query_embeds: np.ndarray, corpus_ids: np.ndarray, index, topk: int,
                 batch_size: int):
    """evaluate_repconc.py:188-206."""
    all_topk_scores, all_topk_ids = [], []
    iterations = math.ceil(len(query_ids) / batch_size)
    if (iterations > 1 and isinstance(index, _DEVICE_INDEXES) and isinstance(query_embeds, np.ndarray)
            and isinstance(corpus_ids, np.ndarray) and corpus_ids.dtype == np.int64):
        # same batches (np.array_split, :193-197), searched back to back with the copy-back of one batch
        # overlapping the scan of the next
        scores, ids = index.search_batches(np.array_split(query_embeds, iterations), topk, corpus_ids=corpus_ids)
        # (the multi-process searchers deliver the result on rank 0 only; the other ranks get empty arrays)
        assert len(scores) == len(ids) and len(scores) in (len(query_ids), 0)
        return scores, ids
    for query_id_iter, query_embeds_iter in tqdm(zip(
        np.array_split(query_ids, iterations),
        np.array_split(query_embeds, iterations),
    ), total=iterations, desc="Batch search"):
        topk_scores, topk_ids = search(query_id_iter, query_embeds_iter, corpus_ids, index, topk)
        all_topk_scores.append(topk_scores)
        all_topk_ids.append(topk_ids)
    if len(all_topk_scores) == 1:            # one batch: concatenate would only copy
        return all_topk_scores[0], all_topk_ids[0]
    all_topk_scores = np.concatenate(all_topk_scores, axis=0)
    all_topk_ids = np.concatenate(all_topk_ids, axis=0)
    return all_topk_scores, all_topk_ids


# ----------------------------------------------------------------------------------------------
# multi-GPU: corpus-sharded scan.  The reference only replicates the index
# (`co.shard = False`, evaluate_repconc.py:131-134); sharding is what the 8-GPU config needs.
# ----------------------------------------------------------------------------------------------
def shard_bounds(ntotal: int, rank: int, world: int):
    """contiguous, balanced row ranges: the first (ntotal % world) shards hold one extra row."""
    base, extra = divmod(int(ntotal), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def load_index_shard_to_gpu(index, rank: Optional[int] = None, world: Optional[int] = None, device=None):
    """This rank's shard of the corpus as a resident index whose ids are global row positions."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(index.ntotal, rank, world)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return faiss.GpuIndexPQ.from_host(index, device, lo, hi)


def merge_shard_results(all_scores: torch.Tensor, all_ids: torch.Tensor):
    """(W, nq, k) per-shard sorted lists -> (nq, k) global top-k (score desc, id asc), on the GPU."""
    lib = _lib.load()
    ops._require_cuda(all_scores, "all_scores")
    W, nq, k = all_scores.shape
    all_scores = all_scores.float().contiguous()
    all_ids = all_ids.long().contiguous()
    with torch.cuda.device(all_scores.device):
        scores = torch.empty((nq, k), dtype=torch.float32, device=all_scores.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=all_scores.device)
        _lib.check(lib.rc_topk_merge(all_scores.data_ptr(), all_ids.data_ptr(), W, nq, k, scores.data_ptr(),
                                     ids.data_ptr(), ops._stream()), "rc_topk_merge")
    return scores, ids


class ShardedSearcher:
    """Multi-PROCESS corpus-sharded search (one process per GPU, torch.distributed): wraps this rank's shard
    (a GpuIndexPQ whose ids are global row positions) and runs the exchange after the local scan.

    Every rank scans its shard for ALL queries (no collective touches the scan).  Then, instead of all-gathering
    W full (nq, k) lists onto every rank and merging all of them everywhere, the queries are dealt out: rank r
    OWNS the r-th block of ceil(nq / W) queries, one all_to_all brings it the W per-shard lists of its block
    (12 B x nq x k x (W-1)/W sent and received per rank), it merges them (rc_topk_merge on nq / W queries), and
    one all_gather of the merged blocks gives every rank the final (nq, k) result (12 B x nq x k received).
    `search` takes and returns CUDA tensors; `search_batches` is the host-array pipeline behind batch_search
    (only rank 0 maps ids and copies results back; the other ranks return empty arrays)."""

    def __init__(self, shard_index, group=None):
        self.shard = shard_index
        self.group = group
        self.device = shard_index.device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        n = torch.tensor([shard_index.ntotal], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(n, group=group)
        self.ntotal = int(n.item())
        self.last_stats = None

    def search(self, x, k, corpus_ids=None):
        """CUDA tensor in -> CUDA tensors out on every rank; numpy in -> numpy out (rank 0; see search_batches)."""
        k = int(k)
        if not isinstance(x, torch.Tensor):
            return self.search_batches([np.ascontiguousarray(x, dtype=np.float32)], k, corpus_ids=corpus_ids)
        x = x.to(self.device)
        s, i = self.shard.search_tensor(x, k)
        self.last_stats = self.shard.last_stats
        W = self.world
        if W == 1:
            return s, i
        nq = s.shape[0]
        nb = (nq + W - 1) // W                       # queries per owner
        if nb * W != nq:                             # pad with empty lists
            ps = torch.full((nb * W, k), -3.4028234663852886e38, dtype=s.dtype, device=s.device)
            pi = torch.full((nb * W, k), -1, dtype=i.dtype, device=i.device)
            ps[:nq], pi[:nq] = s, i
            s, i = ps, pi
        rs, ri = torch.empty_like(s), torch.empty_like(i)
        dist.all_to_all_single(rs, s, group=self.group)          # (W, nb, k): every shard's list of my block
        dist.all_to_all_single(ri, i, group=self.group)
        ms, mi = merge_shard_results(rs.view(W, nb, k), ri.view(W, nb, k))
        fs, fi = torch.empty_like(s), torch.empty_like(i)
        dist.all_gather_into_tensor(fs, ms, group=self.group)
        dist.all_gather_into_tensor(fi, mi, group=self.group)
        return fs[:nq], fi[:nq]

    def search_tensor(self, x, k):
        return self.search(x, k)

    def search_batches(self, batches, k, corpus_ids=None):
        return self.shard.search_batches(batches, k, corpus_ids=corpus_ids, search_fn=self.search,
                                         copy_back=self.rank == 0)


class ReplicatedSearcher:
    """Multi-PROCESS search over index REPLICAS (one process per GPU, every rank holds the whole index): the
    queries are split, rank r scans the r-th block of ceil(nq / W) queries over the full corpus and one all_gather
    of the (nq / W, k) blocks gives every rank the result.  This is the reference's own multi-GPU mode --
    `index_cpu_to_all_gpus` with `shard = False` replicates the index and splits the queries
    (evaluate_repconc.py:130-134) -- and the right one whenever the coded corpus fits one GPU (8.8 M x 48 B =
    424 MB): every per-query cost (tables, threshold sample, re-score, sort) shrinks with W, and the exchange is
    12 B x nq x k per rank instead of the per-shard lists.  Corpora beyond one GPU use ShardedSearcher."""

    def __init__(self, index, group=None):
        self.index = index
        self.group = group
        self.device = index.device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ntotal = index.ntotal
        self.last_stats = None

    def search(self, x, k, corpus_ids=None):
        """CUDA tensor in (the same queries on every rank) -> CUDA tensors out on every rank; numpy in -> numpy out
        (rank 0; see search_batches)."""
        k = int(k)
        if not isinstance(x, torch.Tensor):
            return self.search_batches([np.ascontiguousarray(x, dtype=np.float32)], k, corpus_ids=corpus_ids)
        x = x.to(self.device)
        W = self.world
        if W == 1:
            out = self.index.search_tensor(x, k)
            self.last_stats = self.index.last_stats
            return out
        nq = x.shape[0]
        nb = (nq + W - 1) // W
        lo = min(self.rank * nb, nq)
        hi = min(lo + nb, nq)
        s, i = self.index.search_tensor(x[lo:hi], k)
        if hi > lo:
            self.last_stats = self.index.last_stats
        # one collective for both arrays: (nb, 3k) int32 rows = [score bits | id as two words]
        return self._gather_blocks(s, i, nb, nq, k)

    def search_tensor(self, x, k):
        return self.search(x, k)

    def _gather_blocks(self, s, i, nb, nq, k):
        """(<= nb, k) block results of every rank -> (nq, k) on every rank: one all_gather of (nb, 3k) int32 rows"""
        W = self.world
        n = s.shape[0]
        if n != nb:                                  # last blocks: pad with empty lists
            ps = torch.full((nb, k), -3.4028234663852886e38, dtype=s.dtype, device=s.device)
            pi = torch.full((nb, k), -1, dtype=i.dtype, device=i.device)
            ps[:n], pi[:n] = s, i
            s, i = ps, pi
        mine = torch.cat((s.contiguous().view(torch.int32), i.contiguous().view(torch.int32)), dim=1)
        allr = torch.empty((W * nb, 3 * k), dtype=torch.int32, device=s.device)
        dist.all_gather_into_tensor(allr, mine, group=self.group)
        return allr[:nq, :k].contiguous().view(torch.float32), allr[:nq, k:].contiguous().view(torch.int64)

    def search_batches(self, batches, k, corpus_ids=None):
        """host arrays in: every rank stages and uploads only ITS block of each batch (1 / W of the queries)"""
        k = int(k)
        W, r = self.world, self.rank
        if W == 1:
            return self.index.search_batches(batches, k, corpus_ids=corpus_ids)
        full = [int(len(b)) for b in batches]
        nbs = [(n + W - 1) // W for n in full]
        blocks = [np.ascontiguousarray(b[min(r * nb, n): min((r + 1) * nb, n)], dtype=np.float32)
                  for b, nb, n in zip(batches, nbs, full)]
        todo = [j for j, n in enumerate(full) if n > 0]      # search_batches skips empty batches
        state = {"j": 0}

        def search_block(xd, kk):
            j = todo[state["j"]]
            state["j"] += 1
            s, i = self.index.search_tensor(xd, kk)
            if xd.shape[0]:
                self.last_stats = self.index.last_stats
            return self._gather_blocks(s, i, nbs[j], full[j], kk)

        return self.index.search_batches(blocks, k, corpus_ids=corpus_ids, search_fn=search_block,
                                         copy_back=r == 0, result_sizes=full)


_DEVICE_INDEXES = (faiss.GpuIndexPQ, faiss.MultiGpuIndexPQ, ShardedSearcher, ReplicatedSearcher)


def sharded_search(shard_index, query_embeds, topk: int, group=None):
    """One corpus-sharded search over torch.distributed ranks (see ShardedSearcher).  Returns CUDA tensors."""
    return ShardedSearcher(shard_index, group).search(query_embeds, topk)

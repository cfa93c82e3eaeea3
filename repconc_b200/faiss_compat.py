"""Faiss-free stand-ins for the handful of Faiss objects the reference's PQ-search path touches
(evaluate_repconc.py:78-135,180-206; run_repconc_eval.py:39-58,123-127; finetune_jpq.py:157-162,
176,208-214; run_warmup.py:181-189).  A maintainer swaps `import faiss` for

    from repconc_b200 import faiss_compat as faiss

in those files; the index objects keep the attribute names the call sites use (`index.pq.M`,
`index.pq.centroids`, `index.codes`, `index.ntotal`, `index.search(x, k)`), the search itself runs
in librepconc_b200.so on the GPU (there is no CPU search path: `--cpu_search` also lands here).

Only IndexPQ with nbits == 8 and METRIC_INNER_PRODUCT is modelled -- that is all the reference
builds (evaluate_repconc.py:80-81).
"""
import struct

import numpy as np
import torch

from . import _lib, ops

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


# ----------------------------------------------------------------------------------------------
# std::vector look-alikes (faiss.copy_array_to_vector / vector_to_array call sites)
# ----------------------------------------------------------------------------------------------
class _Vector:
    dtype = np.uint8

    def __init__(self, n=0):
        self._a = np.zeros(n, dtype=self.dtype)
        self._owner = None

    def size(self):
        return int(self._a.size)

    def resize(self, n):
        n = int(n)
        if n != self._a.size:
            new = np.zeros(n, dtype=self.dtype)
            m = min(n, self._a.size)
            new[:m] = self._a[:m]
            self._a = new
            self._touch()

    def _touch(self):
        if self._owner is not None:
            self._owner._version += 1

    def __len__(self):
        return self.size()


class FloatVector(_Vector):
    dtype = np.float32


class ByteVector(_Vector):
    dtype = np.uint8


def copy_array_to_vector(a, v):
    a = np.ascontiguousarray(a).ravel()
    if a.dtype != v.dtype:
        raise TypeError(f"copy_array_to_vector: expected {np.dtype(v.dtype)}, got {a.dtype}")
    v._a = a.copy()
    v._touch()


def vector_to_array(v):
    return v._a.copy()


def omp_set_num_threads(n):  # run_repconc_eval.py:149 -- no CPU threads are used by the search
    return None


def get_num_gpus():
    return torch.cuda.device_count()


# ----------------------------------------------------------------------------------------------
# ProductQuantizer / IndexPQ
# ----------------------------------------------------------------------------------------------
class ProductQuantizer:
    def __init__(self, d, M, nbits, owner=None):
        if nbits != 8:
            raise NotImplementedError("only 8-bit PQ codes (K = 256) are supported")
        if d % M != 0:
            raise ValueError(f"d={d} is not a multiple of M={M}")
        self.d, self.M, self.nbits = int(d), int(M), int(nbits)
        self.ksub = 1 << nbits
        self.dsub = self.d // self.M
        self.code_size = self.M
        self.centroids = FloatVector(self.M * self.ksub * self.dsub)   # layout [m][k][dsub]
        self.centroids._owner = owner

    def centroid_array(self):
        return self.centroids._a.reshape(self.M, self.ksub, self.dsub)


class IndexPQ:
    """faiss.IndexPQ(d, M, 8, METRIC_INNER_PRODUCT) as the reference uses it (evaluate_repconc.py:81)."""

    def __init__(self, d, M, nbits=8, metric=METRIC_INNER_PRODUCT):
        if metric != METRIC_INNER_PRODUCT:
            raise NotImplementedError("only METRIC_INNER_PRODUCT is supported")
        self._version = 0
        self.d = int(d)
        self.metric_type = metric
        self.is_trained = False
        self.ntotal = 0
        self.pq = ProductQuantizer(d, M, nbits, owner=self)
        self.codes = ByteVector(0)                                      # layout [n][m]
        self.codes._owner = self
        self._gpu = None
        self._gpu_version = -1

    def code_array(self):
        return self.codes._a[: self.ntotal * self.pq.M].reshape(self.ntotal, self.pq.M)

    def add_codes(self, new_codes):
        """append (n, M) uint8 codes (what add_docs does through the Faiss vector API)."""
        new_codes = np.ascontiguousarray(new_codes, dtype=np.uint8)
        assert new_codes.ndim == 2 and new_codes.shape[1] == self.pq.M
        self.codes._a = np.concatenate([self.codes._a[: self.ntotal * self.pq.M], new_codes.ravel()])
        self.ntotal += len(new_codes)
        self._version += 1

    def _resident(self, device=None):
        if self._gpu is None or self._gpu_version != self._version or self._gpu.ntotal != self.ntotal:
            self._gpu = GpuIndexPQ.from_host(self, device)
            self._gpu_version = self._version
        return self._gpu

    def search(self, x, k):
        """Faiss `index.search(x, k)` -> (D, I).  Runs on the current CUDA device; the codes are
        uploaded on first use and kept resident until the index is modified."""
        return self._resident().search(x, k)


class IndexIVFPQ(IndexPQ):
    """What `from_pq_to_ivfpq` returns: the same PQ index seen as a one-list IVFPQ (nlist = 1, zero
    coarse centroid, ids = arange) -- numerically the identical search (evaluate_repconc.py:101-118)."""
    nlist = 1


_PINNED = {}     # (device index, name) -> page-locked staging buffer, see GpuIndexPQ._pinned


class GpuIndexPQ:
    """Device-resident PQ index: codes (N, M) uint8 and centroids (M, 256, dsub) fp32 in HBM.
    `search` accepts numpy arrays (returns numpy) or CUDA tensors (returns CUDA tensors, as
    finetune_jpq.py:176 needs).  `set_centroids` refreshes the centroids in place -- the resident
    replacement of JPQ's per-step `synchronize_model_index` re-clone (finetune_jpq.py:208-214)."""

    def __init__(self, codes, centroids, id_offset=0):
        ops._require_cuda(codes, "codes")
        ops._require_cuda(centroids, "centroids")
        assert codes.dtype == torch.uint8 and codes.dim() == 2 and codes.is_contiguous()
        self.codes = codes
        self.centroids = centroids.detach().float().contiguous()
        self.M, self.ksub, self.dsub = self.centroids.shape
        assert codes.shape[1] == self.M and self.ksub == 256
        self.d = self.M * self.dsub
        self.id_offset = int(id_offset)
        self.metric_type = METRIC_INNER_PRODUCT
        self.is_trained = True
        self._ws = None
        self._ws_bytes = {}
        self._stats_buf = (_lib.c_i64 * 4)()
        self._stage = {}
        self._ids_cache = {}
        self.last_stats = None

    @classmethod
    def from_host(cls, index, device=None, lo=0, hi=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        hi = index.ntotal if hi is None else hi
        codes = torch.from_numpy(np.ascontiguousarray(index.code_array()[lo:hi])).to(device)
        cent = torch.from_numpy(index.pq.centroid_array().copy()).to(device)
        return cls(codes, cent, id_offset=lo)

    @property
    def ntotal(self):
        return int(self.codes.shape[0])

    @property
    def device(self):
        return self.codes.device

    def add(self, new_codes):
        """Append (n, M) uint8 codes that are already on the device (the GPU-side `add_docs`: corpus
        encoding with `ops.nn_assign(..., uint8=True)` never has to leave HBM).  Amortised growth."""
        ops._require_cuda(new_codes, "new_codes")
        assert new_codes.dtype == torch.uint8 and new_codes.dim() == 2 and new_codes.shape[1] == self.M
        self._reserve(int(new_codes.shape[0])).copy_(new_codes.to(self.device))

    def _reserve(self, n_new):
        """grow the code storage by n_new rows (amortised) and return the (n_new, M) uint8 view of the new tail"""
        n_old = self.ntotal
        cap = getattr(self, "_cap", None)
        if cap is None or cap.shape[0] < n_old + n_new or cap.data_ptr() != self.codes.data_ptr():
            cap = torch.empty((max(2 * n_old, n_old + n_new), self.M), dtype=torch.uint8, device=self.device)
            cap[:n_old] = self.codes
            self._cap = cap
        self.codes = cap[: n_old + n_new]
        return cap[n_old:n_old + n_new]

    def add_encoded(self, pooled, rotation, normalize=False):
        """Corpus encoding straight into the index: `pooled` (n, d) CUDA fp32 encoder outputs are
        rotated, optionally normalised per sub-vector (METRIC_CENTROID_COS), NN-assigned against this index's
        centroids and written as uint8 rows at the tail of the code storage by ONE kernel (rc_encode_assign) --
        the device-side `model(...return_code=True)` + `.cpu().numpy().astype(np.uint8)` + `add_docs` of
        evaluate_repconc.py:64-70,89-98.  Returns the (n, M) view of the appended codes."""
        ops._require_cuda(pooled, "pooled")
        n_new = int(pooled.shape[0])
        tail = self._reserve(n_new)
        if n_new:
            ops.encode_assign(pooled, rotation, self.centroids, normalize=normalize, uint8=True, return_rotated=False,
                              out=tail)
        return tail

    def set_centroids(self, centroids):
        with torch.no_grad():
            self.centroids.copy_(centroids.detach().reshape(self.centroids.shape))

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def search_tensor(self, x, k):
        """x (nq, d) CUDA fp32 -> (scores (nq,k) fp32, ids (nq,k) int64), both CUDA."""
        lib = _lib.load()
        x = ops._rows_f32(x, "query_embeds")
        if x.device != self.device:
            x = x.to(self.device)
        if x.shape[1] != self.d:
            raise ValueError(f"query width {x.shape[1]} != index dimension {self.d}")
        nq, k = x.shape[0], int(k)
        # (a search of a rank's share of a split batch lasts ~3 ms: the host side of a call is kept short -- no device
        #  context switch when this index's device is already current, workspace size cached per shape)
        switch = torch.cuda.current_device() != self.device.index
        if switch:
            prev = torch.cuda.current_device()
            torch.cuda.set_device(self.device)
        try:
            scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            if nq == 0:
                return scores, ids
            key = (nq, self.ntotal, k)
            nws = self._ws_bytes.get(key)
            if nws is None:
                nws = self._ws_bytes[key] = lib.rc_adc_search_workspace_bytes(nq, self.ntotal, self.M, 256, k)
            ws = self._workspace(nws)
            _lib.check(lib.rc_adc_search(x.data_ptr(), ops._ld(x), self.centroids.data_ptr(), self.codes.data_ptr(),
                                         nq, self.ntotal, self.M, 256, self.dsub, k, self.id_offset,
                                         scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(),
                                         ops._stream()), "rc_adc_search")
            st = self._stats_buf
            lib.rc_adc_last_stats(st)
            self.last_stats = dict(filtered=st[0], dense=st[1], max_candidates=st[2], sample=st[3])
        finally:
            if switch:
                torch.cuda.set_device(prev)
        return scores, ids

    def _pinned(self, name, shape, dtype):
        """persistent page-locked staging buffers (allocated once, grown on demand): host arrays go through
        them so that both directions are single asynchronous DMA copies"""
        n = int(np.prod(shape))
        key = (self.device.index, name)
        buf = _PINNED.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            # page-locking is slow (tens of ms, and it synchronises the device): the buffers are shared by all the
            # index objects of a process on this device and only ever grow
            buf = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
            _PINNED[key] = buf
        return buf[:n].view(shape)

    def _resident_ids(self, corpus_ids):
        """device copy of an external id table.  The cache is validated against a PRIVATE host copy of the table
        with a full comparison (one memcmp-speed pass, ~10 ms for 8.8 M ids, once per search / batch_search call):
        any in-place edit of the caller's array, wherever it is, refreshes the device copy."""
        a = np.ascontiguousarray(corpus_ids)
        if a.dtype != np.int64:
            return None
        hit = self._ids_cache
        if not hit or hit["host"].shape != a.shape or not self._equal_parallel(hit["host"], a):
            self._ids_cache = {"host": a.copy(), "dev": torch.from_numpy(a).to(self.device)}
        return self._ids_cache["dev"]

    def _equal_parallel(self, a, b):
        """a == b everywhere (same shape, contiguous), the comparison split over four threads (numpy releases the GIL)"""
        if getattr(self, "_cmp_pool", None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._cmp_pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix="repconc-cmp")
        a, b = a.reshape(-1), b.reshape(-1)
        n = a.shape[0]
        if n < (1 << 20):
            return bool(np.array_equal(a, b))
        cuts = [n * i // 4 for i in range(5)]
        return all(self._cmp_pool.map(lambda j: bool(np.array_equal(a[cuts[j]:cuts[j + 1]], b[cuts[j]:cuts[j + 1]])),
                                      range(4)))

    def search(self, x, k, corpus_ids=None):
        """Faiss `index.search(x, k)`.  With `corpus_ids` (int64 array) the returned ids are
        `corpus_ids[position]`, mapped on the device before the copy back (evaluate_repconc.py:183)."""
        if isinstance(x, torch.Tensor):
            if not x.is_cuda:
                x = x.to(self.device)
            return self.search_tensor(x, k)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError(f"query_embeds: expected (nq, d), got {x.shape}")
        nq, k = x.shape[0], int(k)
        if nq == 0:
            return np.empty((0, k), np.float32), np.empty((0, k), np.int64)
        with torch.cuda.device(self.device):
            hx = self._pinned("x", x.shape, torch.float32)
            hx.copy_(torch.from_numpy(x))
            s, i = self.search_tensor(hx.to(self.device, non_blocking=True), k)
            ids_dev = self._resident_ids(corpus_ids) if corpus_ids is not None else None
            if ids_dev is not None:
                mapped = torch.empty_like(i)
                _lib.check(_lib.load().rc_map_ids(i.data_ptr(), ids_dev.data_ptr(), ids_dev.numel(), i.numel(),
                                                  mapped.data_ptr(), ops._stream()), "rc_map_ids")
                i = mapped
            hs = self._pinned("s", (nq, k), torch.float32)
            hi = self._pinned("i", (nq, k), torch.int64)
            hs.copy_(s, non_blocking=True)
            hi.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return hs.numpy().copy(), hi.numpy().copy()


    def search_batches(self, batches, k, corpus_ids=None, search_fn=None, copy_back=True, result_sizes=None):
        """Pipelined `search` over a list of host query arrays (the loop of evaluate_repconc.batch_search,
        `:193-203`): while the GPU scans batch i+1, the results of batch i travel device -> pinned host on a copy
        stream and are written into their slice of the preallocated output.  Returns (scores (n,k) fp32,
        ids (n,k) int64) for the concatenated batches -- value for value what per-batch `search` calls give.
        `search_fn(x_dev, k) -> (scores, ids)` replaces this index's own scan (multi-device / multi-process
        searchers pass theirs); `copy_back=False` runs the searches but skips the id mapping and the copy-back
        (ranks other than 0 of a sharded search) and returns empty arrays."""
        k = int(k)
        search_fn = search_fn or self.search_tensor
        in_sizes = [int(len(b)) for b in batches]            # rows staged / uploaded per batch
        # rows RETURNED per batch: the input sizes, unless the search function gathers more than it is given (a rank
        # of a query-split search uploads only its block of each batch and receives the whole batch's results)
        sizes = [int(v) for v in result_sizes] if result_sizes is not None else in_sizes
        assert len(sizes) == len(in_sizes)
        n = sum(sizes)
        out_s = np.empty((n if copy_back else 0, k), np.float32)
        out_i = np.empty((n if copy_back else 0, k), np.int64)
        if n == 0:
            return out_s, out_i
        with torch.cuda.device(self.device):
            compute = torch.cuda.current_stream()
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
                self._up_stream = torch.cuda.Stream(device=self.device)
            # rc_adc_search returns only when the scan has finished (it reads the survivor counters back), so the
            # copy-back of batch i is drained by a helper thread while the caller's thread sits in the scan of
            # batch i+1 (ctypes and numpy's block copies both release the GIL).  A second helper validates the
            # cached device copy of the id table and stages / uploads the queries of the later batches while the
            # first batch is scanned.
            if getattr(self, "_drain_pool", None) is None:
                from concurrent.futures import ThreadPoolExecutor
                self._drain_pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="repconc-drain")
                self._stage_pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="repconc-stage")
            dev_index = self.device.index
            # external id table: the cached device copy is used at once (optimistically) while the helper validates
            # it against the caller's array -- a full comparison, ~10 ms per 100 MB spread over four threads --; a
            # stale cache is detected before anything is returned and the call is simply repeated with the fresh copy
            hit = self._ids_cache
            a_ids = np.ascontiguousarray(corpus_ids) if (corpus_ids is not None and copy_back) else None
            optimistic = (hit["dev"] if (a_ids is not None and hit and a_ids.dtype == np.int64
                                         and hit["host"].shape == a_ids.shape) else None)

            def resident_ids():
                with torch.cuda.device(dev_index):
                    if a_ids.dtype != np.int64:
                        return None, False
                    cur = self._ids_cache
                    if cur and cur["host"].shape == a_ids.shape and self._equal_parallel(cur["host"], a_ids):
                        return cur["dev"], False
                    self._ids_cache = {"host": a_ids.copy(), "dev": torch.from_numpy(a_ids).to(self.device)}
                    return self._ids_cache["dev"], True

            busy = [None, None]     # per staging slot: future of the drain that still reads it
            lo = 0

            def drain(plo, phi, slot, ev, _keep):
                # two workers, one per staging slot: with short scans (a rank's share of a split batch) the host copy
                # of a batch (12 B x nq x k) would otherwise take longer than the scan of the next one
                ev.synchronize()
                if prefault.get(plo) is not None:
                    prefault[plo].result()                    # this batch's output pages exist (touched while scans ran)
                hs = self._pinned(f"s{slot}", (phi - plo, k), torch.float32).numpy()
                hi_ = self._pinned(f"i{slot}", (phi - plo, k), torch.int64).numpy()

                def copy_ids():
                    out_i[plo:phi] = hi_
                # the ids (2/3 of the bytes) on a staging helper (idle once the queries are uploaded), the scores here:
                # the drain of the LAST batch is not overlapped by any scan
                f = self._stage_pool.submit(copy_ids)
                out_s[plo:phi] = hs
                f.result()

            trace = getattr(self, "_trace", None)      # debugging: list that receives (label, seconds) marks
            import time as _time

            def mark(label):
                if trace is not None:
                    trace.append((label, _time.perf_counter()))

            # queries go to the device in groups: the first batch alone (so that the scan starts at once), then the
            # rest in groups of <= 256 MB, each staged into pinned memory and uploaded on its own stream by the
            # staging helper while earlier batches are scanned.  Two pinned / device buffers alternate.
            group_rows = max(1, (256 << 20) // (self.d * 4))
            groups = []                                   # (first batch, end batch, first row, end row)
            bj, row = 0, 0
            while bj < len(batches):
                rows, b0 = 0, bj
                while bj < len(batches) and (rows == 0 or (b0 > 0 and rows + in_sizes[bj] <= group_rows)):
                    rows += in_sizes[bj]
                    bj += 1
                groups.append((b0, bj, row, row + rows))
                row += rows

            def load_group(gi):
                b0, b1, r0, r1 = groups[gi]
                with torch.cuda.device(dev_index):
                    hx = self._pinned(f"xg{gi & 1}", (r1 - r0, self.d), torch.float32)
                    r = 0
                    for b in batches[b0:b1]:
                        b = np.ascontiguousarray(b, dtype=np.float32)
                        if b.ndim != 2 or b.shape[1] != self.d:
                            raise ValueError(f"query_embeds: expected (nq, {self.d}), got {b.shape}")
                        hx[r:r + len(b)].copy_(torch.from_numpy(b))
                        r += len(b)
                    with torch.cuda.stream(self._up_stream):
                        xd = hx.to(self.device, non_blocking=True)
                        up = torch.cuda.Event()
                        up.record(self._up_stream)
                    return xd, up

            def submit_group(gi):
                # the pinned buffer of this slot was last read by the upload of group gi - 2, which the caller waited
                # for (future + event) before searching that group: free by construction
                return self._stage_pool.submit(load_group, gi) if gi < len(groups) else None

            # the freshly allocated output arrays are touched once by a helper while the first scans run: the page
            # faults of their first write (~0.3 us per 4 KB page, 1-2 ms for the last batch) leave the drains.
            # Order of the helper's work: first batch's queries, first batch's output pages, id table, the rest.
            prefault = {}
            touch_rows = []
            if copy_back and n * k >= (1 << 20):
                row = 0
                for nb_ in sizes:                              # in batch order: a drain waits for its own slice only
                    if nb_:
                        touch_rows.append((row, row + nb_))
                    row += nb_

            def touch(plo, phi):
                out_s[plo:phi].fill(0)
                out_i[plo:phi].fill(0)

            touched = [0]

            def touch_ahead(upto):
                # rolling: the output pages of the next two batches, not a backlog of every batch in front of the
                # helpers' other work (query staging, the id halves of the drains)
                while touched[0] < min(upto, len(touch_rows)):
                    plo_, phi_ = touch_rows[touched[0]]
                    prefault[plo_] = self._stage_pool.submit(touch, plo_, phi_)
                    touched[0] += 1

            first = submit_group(0)
            touch_ahead(1)
            ids_future = self._stage_pool.submit(resident_ids) if (corpus_ids is not None and copy_back) else None
            group_future = [first, submit_group(1)]
            touch_ahead(2)
            n_done = 0
            pos_keep = []       # (first row, rows, positions) of the batches mapped with an unvalidated id table
            cur_group, xd_group, g_lo, g_hi = -1, None, 0, 0
            ids_dev = None

            ilo = 0                                       # rows consumed from the staged inputs (lo: rows of results)
            for bi, xb in enumerate(batches):
                nb, nb_in = sizes[bi], in_sizes[bi]
                if nb == 0:
                    continue
                slot = bi & 1
                mark("begin")
                n_done += 1
                touch_ahead(n_done + 2)
                if nb_in > 0 and ilo >= g_hi:
                    cur_group += 1
                    xd_group, up = group_future[cur_group & 1].result()
                    compute.wait_event(up)
                    xd_group.record_stream(compute)
                    _, _, g_lo, g_hi = groups[cur_group]
                    # the other slot belonged to the previous group, whose searches have all returned (rc_adc_search
                    # synchronises): its pinned buffer can be refilled
                    if cur_group >= 1:
                        group_future[(cur_group + 1) & 1] = submit_group(cur_group + 1)
                mark("staged")
                if nb_in > 0:
                    xd = xd_group[ilo - g_lo: ilo - g_lo + nb_in]
                else:                                     # (a rank whose block of this batch is empty still takes part)
                    xd = torch.empty((0, self.d), dtype=torch.float32, device=self.device)
                ilo += nb_in
                mark("h2d")
                s, i = search_fn(xd, k)
                mark("search")
                if not copy_back:
                    lo += nb
                    continue
                if ids_future is not None:
                    if ids_dev is None:
                        ids_dev = optimistic if optimistic is not None else ids_future.result()[0]
                        if ids_dev is None:
                            raise TypeError("search_batches: corpus_ids must be an int64 array")
                    mapped = torch.empty_like(i)
                    _lib.check(_lib.load().rc_map_ids(i.data_ptr(), ids_dev.data_ptr(), ids_dev.numel(), i.numel(),
                                                      mapped.data_ptr(), ops._stream()), "rc_map_ids")
                    if optimistic is not None:
                        pos_keep.append((lo, nb, i))         # positions, in case the cached table turns out stale
                    i = mapped
                done = torch.cuda.Event()
                done.record(compute)
                mark("mapped")
                if busy[slot] is not None:                # the staging buffers of this slot are free again
                    busy[slot].result()
                mark("slotfree")
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(done)
                    self._pinned(f"s{slot}", (nb, k), torch.float32).copy_(s, non_blocking=True)
                    self._pinned(f"i{slot}", (nb, k), torch.int64).copy_(i, non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(self._copy_stream)
                busy[slot] = self._drain_pool.submit(drain, lo, lo + nb, slot, copied, (s, i))
                lo += nb
                mark("submitted")
            for f in busy:
                if f is not None:
                    f.result()
            mark("drained")
            stale = ids_future is not None and optimistic is not None and ids_future.result()[1]
            mark("validated")
            if stale:
                # the caller's id table had changed since it was cached: the kept positions are mapped again with the
                # refreshed device copy (no search is repeated, no other rank is involved)
                fresh = ids_future.result()[0]
                for plo, pnb, pos in pos_keep:
                    mapped = torch.empty_like(pos)
                    _lib.check(_lib.load().rc_map_ids(pos.data_ptr(), fresh.data_ptr(), fresh.numel(), pos.numel(),
                                                      mapped.data_ptr(), ops._stream()), "rc_map_ids")
                    out_i[plo:plo + pnb] = mapped.cpu().numpy()
        return out_s, out_i


class MultiGpuIndexPQ:
    """One process, several devices: the corpus is sharded contiguously over `shards` (GpuIndexPQ objects, one per
    device, ids = global row positions).  A search sends the queries to every device (peer copies), one host
    thread per device runs that shard's scan (ctypes releases the GIL, so the scans run concurrently), the
    per-shard (nq, k) lists are copied to the first device and merged there by rc_topk_merge.
    This is what `load_index_to_gpu(index)` returns when several GPUs are visible -- the counterpart of the
    reference's `index_cpu_to_all_gpus` (evaluate_repconc.py:130-134), sharding instead of replicating."""

    def __init__(self, shards):
        assert len(shards) >= 1
        self.shards = list(shards)
        self.device = self.shards[0].device
        self.M, self.dsub, self.d = self.shards[0].M, self.shards[0].dsub, self.shards[0].d
        self.metric_type = METRIC_INNER_PRODUCT
        self.is_trained = True
        self.last_stats = None
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=len(self.shards), thread_name_prefix="repconc-shard")

    @classmethod
    def from_host(cls, index, devices):
        n, W = index.ntotal, len(devices)
        shards = []
        for r, dev in enumerate(devices):
            base, extra = divmod(n, W)
            lo = r * base + min(r, extra)
            hi = lo + base + (1 if r < extra else 0)
            shards.append(GpuIndexPQ.from_host(index, torch.device("cuda", dev), lo, hi))
        return cls(shards)

    @property
    def ntotal(self):
        return sum(s.ntotal for s in self.shards)

    def set_centroids(self, centroids):
        for s in self.shards:
            s.set_centroids(centroids.to(s.device))

    def search_tensor(self, x, k):
        k = int(k)
        x = ops._rows_f32(x if x.is_cuda else x.to(self.device), "query_embeds")
        torch.cuda.current_stream(x.device).synchronize()          # the queries are complete before other devices read them

        def run(shard):
            with torch.cuda.device(shard.device):
                s, i = shard.search_tensor(x.to(shard.device), k)
                if shard.device != self.device:
                    s, i = s.to(self.device), i.to(self.device)
                torch.cuda.current_stream(shard.device).synchronize()
                return s, i, shard.last_stats
        res = list(self._pool.map(run, self.shards))
        self.last_stats = {key: sum(r[2][key] for r in res) if key in ("filtered", "dense") else
                           max(r[2][key] for r in res) for key in res[0][2]}
        if len(res) == 1:
            return res[0][0], res[0][1]
        lib = _lib.load()
        with torch.cuda.device(self.device):
            ss = torch.stack([r[0] for r in res]).contiguous()
            ii = torch.stack([r[1] for r in res]).contiguous()
            nq = ss.shape[1]
            scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            _lib.check(lib.rc_topk_merge(ss.data_ptr(), ii.data_ptr(), len(res), nq, k, scores.data_ptr(),
                                         ids.data_ptr(), ops._stream()), "rc_topk_merge")
        return scores, ids

    def search(self, x, k, corpus_ids=None):
        if isinstance(x, torch.Tensor):
            return self.search_tensor(x, k)
        return self.search_batches([np.ascontiguousarray(x, dtype=np.float32)], k, corpus_ids=corpus_ids)

    def search_batches(self, batches, k, corpus_ids=None):
        return self.shards[0].search_batches(batches, k, corpus_ids=corpus_ids, search_fn=self.search_tensor)


# ----------------------------------------------------------------------------------------------
# Faiss index file (`faiss.write_index` / `read_index`) for IndexPQ: fourcc "IxPq"
# (faiss/impl/index_write.cpp of the pinned 1.7.1: index header, ProductQuantizer, codes vector,
#  search_type, encode_signs, polysemous_ht).  Faiss is absent in this image, so the layout is
#  restated from its published source and NOT verified against a Faiss-written file.
# ----------------------------------------------------------------------------------------------
_FOURCC_IXPQ = struct.unpack("<I", b"IxPq")[0]


def write_index(index, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", _FOURCC_IXPQ))
        f.write(struct.pack("<iqqq", index.d, index.ntotal, 1 << 20, 1 << 20))
        f.write(struct.pack("<?i", bool(index.is_trained), index.metric_type))
        f.write(struct.pack("<QQQ", index.pq.d, index.pq.M, index.pq.nbits))
        cent = np.ascontiguousarray(index.pq.centroids._a, dtype=np.float32)
        f.write(struct.pack("<Q", cent.size))
        f.write(cent.tobytes())
        codes = np.ascontiguousarray(index.codes._a[: index.ntotal * index.pq.M], dtype=np.uint8)
        f.write(struct.pack("<Q", codes.size))
        f.write(codes.tobytes())
        f.write(struct.pack("<i?i", 0, False, index.pq.nbits * index.pq.M + 1))  # ST_PQ, no signs, ht


def read_index(path):
    with open(path, "rb") as f:
        (h,) = struct.unpack("<I", f.read(4))
        if h != _FOURCC_IXPQ:
            raise NotImplementedError(f"read_index: only IndexPQ ('IxPq') files are supported, got fourcc {h:#x}")
        d, ntotal, _, _ = struct.unpack("<iqqq", f.read(28))
        is_trained, metric = struct.unpack("<?i", f.read(5))
        if metric > 1:
            f.read(4)
        pd, pM, pnbits = struct.unpack("<QQQ", f.read(24))
        index = IndexPQ(pd, pM, pnbits, metric)
        (n,) = struct.unpack("<Q", f.read(8))
        copy_array_to_vector(np.frombuffer(f.read(4 * n), dtype=np.float32), index.pq.centroids)
        (n,) = struct.unpack("<Q", f.read(8))
        copy_array_to_vector(np.frombuffer(f.read(n), dtype=np.uint8), index.codes)
        index.ntotal = ntotal
        index.is_trained = is_trained
        assert index.d == d and n == ntotal * pM
    return index

#!/usr/bin/env python
"""bench.py -- headline benchmark of the RepCONC constrained-PQ hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--only adc,assign,adc_c4,assign_c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).

Primary metric (BASELINE.json configs[1]): PQ asymmetric-distance top-1000 search, 8,841,823 synthetic documents,
768-d, M=48, K=256; a "step" is one `search_batch` of 1200 queries (evaluate_repconc.py:38) against the whole
corpus.  N > 1: the same corpus REPLICATED on every GPU with the queries split N ways -- the reference's own multi-GPU
mode (index_cpu_to_all_gpus, evaluate_repconc.py:130-134); "strong" at every N, N = 1 included.  The corpus-sharded
search of the same corpus is the secondary record `adc_sharded`.
  value         queries/s with queries and corpus resident in HBM (CUDA events, max over ranks)
  e2e           the same through the reference-facing `batch_search` with HOST numpy queries / results
  roofline      the filtered-scan kernel against the limit that binds it (shared-memory wavefronts per clock per SM)
                and, as `hbm_algorithmic_frac`, SURVEY 8d's algorithmic bytes / duration / MEASURED_PEAKS.json
  cpu_baseline  the oracle's C/OpenMP restatement of the Faiss IndexPQ scan on a bounded sample of the SAME corpus,
                with bit-exact parity and MRR@10 equality on those queries
Secondary records in the same line, each with its own roofline / e2e / (N = 1) cpu_baseline:
  assign      BASELINE configs[2]: training-step loop, 8192 embeddings per GPU, M=48, Sinkhorn 50 iterations + decode
              + MSE forward/backward (weak scaling, the reference's all-reduces between ranks)
  assign_c5   BASELINE configs[4]: the same with M=96 (global batch 65,536 on 8 GPUs; 197 KB of row sums exchanged
              per iteration)
  adc_c4      BASELINE configs[3]: corpus-sharded ADC, 8,000,000 documents PER GPU (64 M on 8 GPUs), M=32, top-1000

Synthetic data follows SURVEY 8(d): documents x ~ N(0, I_768) from per-block seeds, centroids = sub-vectors of the
first 256 documents ("trained-like"), corpus codes = NN assign of the documents BY THE PATH ITSELF (rc_nn_assign, as
encode_corpus does), queries = document r(i) + 0.5 N(0, I), qrels {i: r(i)}; MRR@10 per eval_utils.py:136-141,182-190.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, K, TOPK = 768, 256, 1000
SEARCH_BATCH = 1200         # evaluate_repconc.py:38
SK_EPS, SK_ITERS, MSE_W = 0.003, 50, 1e-4
ASSIGN_B = 8192
C2_DOCS, C2_M = 8_841_823, 48          # MS MARCO passage count
C4_DOCS_PER_GPU, C4_M = 8_000_000, 32
C5_M = 96
GEN_BLOCK = 1 << 18                     # documents per generator block

ADC_WORKLOAD = ("ADC top-1000, 8,841,823 docs x 768-d, M=48 K=256, 1200-query search batches "
                "(BASELINE configs[1])")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed ncu capture of THIS
    round's build (written by tools/ncu_traffic.py into profiles/r02_traffic.json); None when not captured."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            import atexit
            atexit.register(self.__exit__)              # never leave the sampler behind if the bench dies
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None and self.proc.poll() is None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# SURVEY 8(d) synthetic evaluation set
# ----------------------------------------------------------------------------------------------------------
def make_eval_set(torch, ops, dev, M, n_docs, lo, hi, n_queries, seed):
    """Documents [0, n_docs) are N(0, I) rows generated block by block from per-block seeds (any rank can regenerate
    any block); centroids are the sub-vectors of the first 256 documents; rows [lo, hi) are coded by the path's own
    NN assign.  Queries: document r(i) + 0.5 noise, r(i) seeded.  Returns (codes (hi-lo, M) uint8 CUDA,
    centroids (M, 256, D/M) CUDA, queries (n_queries, D) CUDA, rel (n_queries,) int64 numpy)."""
    ds = D // M
    rel = np.random.default_rng(seed).integers(0, n_docs, size=n_queries)
    rel_t = torch.from_numpy(rel).to(dev)
    order = torch.argsort(rel_t)
    rel_sorted = rel_t[order]
    queries = torch.empty((n_queries, D), device=dev)
    codes = torch.empty((hi - lo, M), dtype=torch.uint8, device=dev)
    cent = None
    n_blocks = (n_docs + GEN_BLOCK - 1) // GEN_BLOCK
    bounds = torch.searchsorted(rel_sorted, torch.arange(n_blocks + 1, device=dev) * GEN_BLOCK).tolist()
    for b in range(n_blocks):
        b_lo, b_hi = b * GEN_BLOCK, min(n_docs, (b + 1) * GEN_BLOCK)
        need_q = bounds[b + 1] > bounds[b]
        need_c = b_lo < hi and b_hi > lo
        if not (need_q or need_c or b == 0):
            continue
        g = torch.Generator(device=dev).manual_seed(seed * 100003 + b)
        docs = torch.randn((GEN_BLOCK, D), generator=g, device=dev)[: b_hi - b_lo]
        if b == 0:
            cent = docs[:K].reshape(K, M, ds).transpose(0, 1).contiguous()
        if need_c:
            s, e = max(lo, b_lo), min(hi, b_hi)
            codes[s - lo:e - lo] = ops.nn_assign(docs[s - b_lo:e - b_lo], cent, uint8=True)
        if need_q:
            sel = order[bounds[b]:bounds[b + 1]]
            queries[sel] = docs[rel_t[sel] - b_lo]
    g = torch.Generator(device=dev).manual_seed(seed * 100003 + 99991)
    queries += 0.5 * torch.randn((n_queries, D), generator=g, device=dev)
    return codes, cent, queries, rel


def mrr_at_10(ids, rel):
    """eval_utils.py:136-141,182-190 for one relevant document per query (top-10 truncation, mean RR, 5 dp)."""
    ids = np.asarray(ids)[:, :10]
    hit = ids == np.asarray(rel)[:, None]
    rank = np.where(hit.any(1), hit.argmax(1) + 1, 0)
    return round(float(np.where(rank > 0, 1.0 / np.maximum(rank, 1), 0.0).mean()), 5)


# ----------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path on the box's host cores
# ----------------------------------------------------------------------------------------------------------
def _import_reference_module():
    """the UNMODIFIED reference package (pip-installed into baseline/_ref) or None"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "repconc")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        from repconc.models.repconc import modeling_repconc
        return modeling_repconc
    except Exception:
        return None


def reference_quantize_seconds(B, M, threads, reps=1, seed=7):
    """Time the imported reference's RepCONC.quantize (modeling_repconc.py:47-67, constraint on) on CPU tensors.
    Returns (best seconds, codes (B,M) int64 numpy, x, c) or None if the reference cannot be imported."""
    mod = _import_reference_module()
    if mod is None:
        return None
    import torch
    from transformers import PretrainedConfig
    torch.set_num_threads(threads)
    cfg = PretrainedConfig(hidden_size=D)
    cfg.MCQ_M, cfg.MCQ_K, cfg.similarity_metric = M, K, "METRIC_IP"

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.config = cfg
    r = np.random.default_rng(seed)
    x = r.standard_normal((B, D), dtype=np.float32)
    c = r.standard_normal((M, K, D // M), dtype=np.float32)
    model = mod.RepCONC(cfg, Dummy(), True, SK_EPS, SK_ITERS)
    with torch.no_grad():
        model.centroids.copy_(torch.from_numpy(c))
    best, codes = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        codes = model.quantize(torch.from_numpy(x))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, codes.contiguous().numpy(), x, c


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores.
      * ADC (the line's metric): Faiss is not installable here, so the scan is the oracle's C/OpenMP
        restatement of IndexPQ search (kind "port") on all host threads; each step scans a bounded sample of the
        1200-query batch against all 8,841,823 documents.  The codes are uniform random bytes and the queries
        N(0, I): the CPU scan's cost does not depend on the code distribution, and building the 8(d) corpus needs
        the path's own NN assign (minutes on the CPU, or our GPU kernels, which must not run in this arm).
      * assign: the UNMODIFIED reference module imported from baseline/_ref -- RepCONC.quantize on CPU tensors,
        config C1 (10,000 x 768, M=48, T=50), kind "reference"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    O.set_num_threads(cores)
    r = np.random.default_rng(7)
    nq = max(4 * cores, 32)
    M = C2_M
    codes = r.integers(0, 256, size=(C2_DOCS, M), dtype=np.uint8)
    c = r.standard_normal((M, K, D // M), dtype=np.float32)
    times = []
    for it in range(args.warmup + args.steps):
        q = r.standard_normal((nq, D), dtype=np.float32)
        t0 = time.perf_counter()
        O.adc_search(q, c, codes, TOPK)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
        if sum(times) > 150.0:                       # keep the whole run within a few minutes
            break
    del codes
    total = sum(times)
    value = nq * len(times) / total
    sample = (f"{nq} of the {SEARCH_BATCH} queries per step x {C2_DOCS} docs, k={TOPK}, {len(times)} timed steps; "
              "uniform random codes")
    assign = None
    try:
        import psutil
        big = psutil.virtual_memory().available > 48 * (1 << 30)
    except Exception:
        big = False
    Bc1 = 10_000 if big else 2048
    res = reference_quantize_seconds(Bc1, 48, cores, reps=1)
    if res is not None:
        dt = res[0]
        assign = {"metric": "constrained_assign_embeddings_per_sec", "value": Bc1 / dt, "unit": "embeddings/s",
                  "config": {"workload": "CPU reference: synthetic 768-d embeddings, M=48 K=256, one Sinkhorn assign "
                                         "(BASELINE configs[0])", "batch": Bc1, "sk_iters": SK_ITERS,
                             "sk_epsilon": SK_EPS},
                  "cpu_baseline": {"value": Bc1 / dt, "unit": "embeddings/s", "cores": cores, "kind": "reference",
                                   "sample": f"RepCONC.quantize imported from baseline/_ref, {Bc1} x {D}, M=48, "
                                             f"T={SK_ITERS}, {dt:.2f} s, torch CPU threads = {cores}"}}
    print(json.dumps({
        "impl": "reference", "metric": "adc_queries_per_sec", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ADC_WORKLOAD},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "assign": assign,
    }))


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--only", default="adc,adc_sharded,assign,adc_c4,assign_c5",
                    help="comma-separated workloads to run (the primary `adc` always runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    only = set(args.only.split(",")) | {"adc"}

    import torch
    import torch.distributed as dist
    from repconc_b200 import _lib, ops
    from repconc_b200 import evaluate_repconc as E
    from repconc_b200.faiss_compat import GpuIndexPQ

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    lib = _lib.load()
    lib.rc_adc_enable_timing(1)
    W, Ksteps = max(args.warmup, 3), args.steps
    hbm_peak, peak_src = peaks()
    cores = os.cpu_count() or 1
    do_cpu = rank == 0 and world == 1 and not args.skip_cpu
    if do_cpu:
        from oracle import oracle as O
        O.build()
        O.set_num_threads(cores)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # clocks / throttle reasons are sampled from the first GPU-timed region to the last: every sample is under load
    clk = ClockSampler(local)
    clk_started = [False]

    def clocks_on():
        if not clk_started[0]:
            clk.__enter__()
            clk_started[0] = True

    # ------------------------------------------------------------------ ADC workloads
    def run_adc(M, n_docs, lo, hi, seed, tag, replicated=False):
        """One ADC workload: documents [lo, hi) of an n_docs corpus on this rank.  Corpus-sharded: all queries on
        every rank, per-shard lists exchanged and merged.  `replicated` (lo, hi = whole corpus on every rank): the
        queries are split over the ranks, the reference's own multi-GPU mode (evaluate_repconc.py:130-134)."""
        n_q = (W + Ksteps) * SEARCH_BATCH
        codes, cent, q_all, rel = make_eval_set(torch, ops, dev, M, n_docs, lo, hi, n_q, seed)
        index = GpuIndexPQ(codes, cent, id_offset=lo)
        q_dev = q_all.view(W + Ksteps, SEARCH_BATCH, D)
        q_host = q_dev.cpu().numpy()
        rel = rel.reshape(W + Ksteps, SEARCH_BATCH)
        corpus_ids = np.arange(n_docs, dtype=np.int64)
        sharded = None
        if world > 1:
            sharded = E.ReplicatedSearcher(index) if replicated else E.ShardedSearcher(index)

        def step(i):
            if world > 1:
                return sharded.search(q_dev[i], TOPK)
            return index.search_tensor(q_dev[i], TOPK)

        for i in range(W):
            step(i)
        barrier()
        clocks_on()
        scan_ms, scan_launches, wavefronts = 0.0, 0, 0.0
        fallbacks = 0
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ids_dev = []
        e0.record()
        for i in range(W, W + Ksteps):
            out = step(i)
            ids_dev.append(out[1][:, :10])
            scan_ms += lib.rc_adc_last_scan_ms()
            scan_launches += lib.rc_adc_last_scan_launches()
            wavefronts += lib.rc_adc_last_scan_wavefronts()
            fallbacks += index.last_stats["dense"] if index.last_stats else 0
        e1.record()
        barrier()
        launches = _lib.launch_count() - l0
        ms = max_over_ranks(e0.elapsed_time(e1))
        qps = Ksteps * SEARCH_BATCH / (ms / 1e3)
        stats = dict(index.last_stats or {}, fallback_queries_in_timed_region=fallbacks)
        # MRR@10 of the timed queries (N > 1: the merged result is complete on rank 0)
        mrr = mrr_at_10(torch.cat(ids_dev).cpu().numpy(), rel[W:].reshape(-1))
        # e2e: host numpy queries -> batch_search -> host numpy results.  ONE batch_search call over the K steps'
        # queries with batch_size 1200, as the reference's evaluator issues it (evaluate_repconc.py:188-206): every
        # step's queries go host -> device and every step's results device -> host inside the timed region, the
        # copy-back of one batch overlapping the scan of the next.  N > 1: the same call on a sharded index (every
        # rank scans its shard, rank 0 receives the merged lists and is the only one that copies back).
        q_e2e = np.ascontiguousarray(q_host[W:W + Ksteps].reshape(-1, D))
        qids = np.arange(len(q_e2e))
        target = sharded if world > 1 else index
        # warm-up: the same call shape (staging buffers, workspaces and the device copy of the id table are steady state)
        E.batch_search(qids, q_e2e, corpus_ids, target, TOPK, SEARCH_BATCH)
        barrier()
        t0 = time.perf_counter()
        out = E.batch_search(qids, q_e2e, corpus_ids, target, TOPK, SEARCH_BATCH)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            assert out[0].shape == (Ksteps * SEARCH_BATCH, TOPK)
            assert mrr_at_10(out[1], rel[W:].reshape(-1)) == mrr
        e2e_qps = Ksteps * SEARCH_BATCH / e2e_s
        # roofline of the filtered scan (this rank's shard)
        n_shard = hi - lo
        alg_per_query = n_shard * M + 4 * M * K + 12 * TOPK
        q_per_rank = SEARCH_BATCH / world if (replicated and world > 1) else SEARCH_BATCH
        alg_per_launch = alg_per_query * (Ksteps * q_per_rank) / max(scan_launches, 1)
        scan_avg_s = scan_ms / 1e3 / max(scan_launches, 1)
        hbm_gbs = alg_per_launch / scan_avg_s / 1e9 if scan_avg_s > 0 else 0.0
        kname = lib.rc_adc_last_scan_kernel().decode()
        rec = {"qps": qps, "ms": ms, "e2e_qps": e2e_qps, "launches": launches, "stats": stats, "mrr": mrr,
               "n_shard": n_shard, "scan_ms": scan_ms, "scan_launches": scan_launches, "wavefronts": wavefronts,
               "alg_per_launch": alg_per_launch, "scan_avg_s": scan_avg_s, "hbm_gbs": hbm_gbs, "kernel": kname,
               "q_per_rank": q_per_rank,
               "cpu": None}
        if do_cpu:
            nq_cpu = max(8 * cores, 32)
            codes_h, c_h, qh = codes.cpu().numpy(), cent.cpu().numpy(), q_host[W][:nq_cpu]
            t0 = time.perf_counter()
            so, io = O.adc_search(qh, c_h, codes_h, TOPK)
            dt = time.perf_counter() - t0
            sg, ig = index.search(qh, TOPK)
            rec["cpu"] = {"value": nq_cpu / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                          "sample": f"{nq_cpu} queries x {n_docs} docs, M={M}, k={TOPK} (oracle C/OpenMP scan + heap, "
                                    "same corpus as the GPU leg)",
                          "gpu_matches_cpu_bit_exact": bool(np.array_equal(io, ig) and np.array_equal(so, sg)),
                          "mrr_at_10_gpu": mrr_at_10(ig, rel[W][:nq_cpu]),
                          "mrr_at_10_cpu": mrr_at_10(io, rel[W][:nq_cpu])}
            rec["cpu"]["mrr_at_10_identical"] = rec["cpu"]["mrr_at_10_gpu"] == rec["cpu"]["mrr_at_10_cpu"]
            del codes_h
        del index, codes, sharded
        torch.cuda.empty_cache()
        return rec

    def adc_roofline(rec, sm_mhz):
        """The filtered scan is bound by shared-memory bandwidth: one 128-byte wavefront per clock per SM is the
        LSU data-pipe ceiling (tools/microbench.cu measures it with conflict-free LDS.128).  `achieved` counts the
        wavefronts the kernel must issue (analytic, from the library: gathers + exchange + code bytes, the same
        count ncu reports as l1tex__data_pipe_lsu_wavefronts) per clock per SM at the sampled SM clock."""
        clk_hz = (sm_mhz or 1965.0) * 1e6
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        wf = rec["wavefronts"] / max(rec["scan_launches"], 1)
        ach = wf / (rec["scan_avg_s"] * clk_hz * sms) if rec["scan_avg_s"] > 0 else 0.0
        return {"bound": "smem_lsu", "kernel": rec["kernel"], "achieved": ach, "peak": 1.0,
                "unit": "wavefronts/clk/SM", "frac": ach,
                "peak_source": "LSU data pipe: 1 wavefront (128 B) per clock per SM; conflict-free LDS.128 "
                               "microbenchmark profiles/r02_microbench_pipes.txt",
                # (the committed captures are launches over a full 1200-query batch)
                "traffic": ncu_traffic(rec["kernel"] + f"@{rec['n_shard']}") if rec["q_per_rank"] == SEARCH_BATCH else None,
                "wavefronts_per_launch": wf, "launch_ms": rec["scan_avg_s"] * 1e3,
                "share_of_step": rec["scan_ms"] / rec["ms"],
                "hbm_algorithmic_frac": rec["hbm_gbs"] / hbm_peak, "hbm_algorithmic_gbs": rec["hbm_gbs"],
                "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peak_src,
                "algorithmic_bytes_per_launch": rec["alg_per_launch"],
                "note": "hbm_algorithmic_frac = SURVEY 8d bytes (every query streams every code byte once) / "
                        "duration / measured HBM peak; it exceeds 1 because 8-16 queries share every code byte a "
                        "CTA reads and a split's codes stay L2-resident (see `traffic` for the DRAM bytes)"}

    # primary metric at N > 1: index replicas + query split (the corpus fits every GPU; the reference's multi-GPU
    # mode); the corpus-sharded search of the same corpus is reported beside it, and configs[3] below is sharded
    adc = run_adc(C2_M, C2_DOCS, 0, C2_DOCS, seed=1234, tag="adc", replicated=True)
    adc_sharded = None
    if world > 1 and "adc_sharded" in only:
        lo, hi = E.shard_bounds(C2_DOCS, rank, world)
        adc_sharded = run_adc(C2_M, C2_DOCS, lo, hi, seed=1234, tag="adc_sharded")
    adc_c4 = None
    if "adc_c4" in only:
        n4 = C4_DOCS_PER_GPU * world
        adc_c4 = run_adc(C4_M, n4, rank * C4_DOCS_PER_GPU, (rank + 1) * C4_DOCS_PER_GPU, seed=4321, tag="adc_c4")

    # ------------------------------------------------------------------ constrained-assign workloads
    def run_assign(M, seed, tag):
        ds = D // M
        ga = torch.Generator(device=dev).manual_seed(seed + rank)
        gc = torch.Generator(device=dev).manual_seed(seed + 5000)
        cen = torch.randn((M, K, ds), generator=gc, device=dev).requires_grad_(True)
        xs = torch.randn((W + Ksteps, ASSIGN_B, D), generator=ga, device=dev)
        gs = torch.randn((ASSIGN_B, D), generator=ga, device=dev) / ASSIGN_B
        xs_host = xs.cpu().pin_memory()

        def assign_step(x):
            x = x.detach().requires_grad_(True)
            codes_ = ops.constrained_assign(x, cen, SK_EPS, SK_ITERS)           # table + Sinkhorn + argmax
            mse, sur = ops.quantization_loss(x, cen, codes_, gs, MSE_W)         # decode + loss
            (mse + sur).backward()                                             # grad_x, grad_centroids
            return codes_, mse

        for i in range(W):
            assign_step(xs[i])
        barrier()
        clocks_on()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(W, W + Ksteps):
            codes_last, _ = assign_step(xs[i])
        e1.record()
        barrier()
        launches = _lib.launch_count() - l0
        ms = max_over_ranks(e0.elapsed_time(e1))
        eps_ = Ksteps * ASSIGN_B * world / (ms / 1e3)
        # e2e: every step's embeddings come from pinned host memory and its codes + loss go back to the host;
        # the H2D of step i+1 is issued on a copy stream while step i computes (double buffering)
        copy_stream = torch.cuda.Stream(device=dev)

        def fetch(i):
            with torch.cuda.stream(copy_stream):
                xd = xs_host[i].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return xd, ev

        # results: the codes (B x M int64) and the loss of EVERY step travel to pinned host memory on a second copy
        # stream and are read by the host one step later, while the next step computes -- a training loop that logs
        # step i's loss after launching step i + 1; a blocking read per step would idle the GPU for the ~100 kernel
        # launches of the following step
        back_stream = torch.cuda.Stream(device=dev)
        h_codes = [torch.empty((ASSIGN_B, M), dtype=torch.int64, pin_memory=True) for _ in range(2)]
        h_loss = [torch.empty((), dtype=torch.float32, pin_memory=True) for _ in range(2)]

        def e2e_loop(first, last):
            nxt = fetch(first)
            pending, res = None, None
            for i in range(first, last):
                xd, ev = nxt
                if i + 1 < last:
                    nxt = fetch(i + 1)
                torch.cuda.current_stream().wait_event(ev)
                c_, mse_ = assign_step(xd)
                done = torch.cuda.Event()
                done.record()
                slot = i & 1
                with torch.cuda.stream(back_stream):
                    back_stream.wait_event(done)
                    h_codes[slot].copy_(c_, non_blocking=True)       # D2H of the codes and the loss
                    h_loss[slot].copy_(mse_.detach(), non_blocking=True)
                    got = torch.cuda.Event()
                    got.record(back_stream)
                c_.record_stream(back_stream)
                if pending is not None:
                    pending[1].synchronize()
                    res = h_codes[pending[0]].numpy(), float(h_loss[pending[0]])
                pending = (slot, got)
            pending[1].synchronize()
            res = h_codes[pending[0]].numpy(), float(h_loss[pending[0]])
            return res

        e2e_loop(0, W)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(W, W + Ksteps)
        barrier()
        e2e = Ksteps * ASSIGN_B * world / max_over_ranks(time.perf_counter() - t0)
        # the dominant kernel: the Sinkhorn iteration loop.  Its per-iteration time is the difference of two complete
        # assignments (50 and 10 iterations) through the product entry point, CUDA events on the launching stream.
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def assign_ms(T, reps=5):
            ts = []
            for _ in range(reps):
                barrier()
                p0.record()
                ops.constrained_assign(xs[0], cen, SK_EPS, T)
                p1.record()
                torch.cuda.synchronize()
                ts.append(max_over_ranks(p0.elapsed_time(p1)))
            return sorted(ts)[len(ts) // 2]
        assign_ms(SK_ITERS, 2)
        t_full, t_short = assign_ms(SK_ITERS), assign_ms(10)
        it_s = (t_full - t_short) / 1e3 / (SK_ITERS - 10)
        alg_it = 2 * 4 * M * K * ASSIGN_B           # the table is read once per half-iteration
        alg_step = ASSIGN_B * (4 * D + 4 * M * K * (2 * SK_ITERS + 1) + 8 * M)
        rec = {"metric": "constrained_assign_embeddings_per_sec", "value": eps_, "unit": "embeddings/s",
               "ms_per_step": ms / Ksteps, "scaling": "weak",
               "config": {"workload": f"training-step loop ({tag})", "batch_per_gpu": ASSIGN_B,
                          "global_batch": ASSIGN_B * world, "M": M, "K": K, "sk_iters": SK_ITERS,
                          "sk_epsilon": SK_EPS, "step": "table + centring + Sinkhorn + argmax + decode + MSE fwd/bwd",
                          "row_sum_exchange_bytes_per_iteration": M * K * 8 if world > 1 else 0},
               "e2e": {"value": e2e, "unit": "embeddings/s", "h2d_bytes_per_step": ASSIGN_B * D * 4,
                       "d2h_bytes_per_step": ASSIGN_B * M * 8 + 4},
               "gpu_launches": launches, "assignment_ms": t_full,
               "roofline": {"bound": "hbm", "kernel": "one Sinkhorn iteration of rc_sinkhorn_solve (survivor-list or "
                                                       "selection pass + row-sum reduce / update"
                                                       + (" + peer exchange)" if world > 1 else ")"),
                            "achieved": alg_it / it_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg_it / it_s / 1e9 / hbm_peak, "peak_source": peak_src,
                            "traffic": ncu_traffic(f"sinkhorn_iteration@M{M}"),
                            # what the list-pass iteration physically moves (ncu DRAM bytes of one launch) over the mean
                            # iteration time and the measured peak: the fraction to improve (the pass is latency-bound)
                            "hbm_physical_frac": (ncu_traffic(f"sinkhorn_iteration@M{M}") / it_s / 1e9 / hbm_peak
                                                  if ncu_traffic(f"sinkhorn_iteration@M{M}") and it_s > 0 else None),
                            "kernel_name": "sinkhorn_step_kernel + sinkhorn_reduce_update_kernel (launch chain)",
                            "algorithmic_bytes_per_launch": alg_it, "launch_ms": it_s * 1e3,
                            "step_hbm_algorithmic_frac": alg_step / (ms / Ksteps / 1e3) / 1e9 / hbm_peak,
                            "note": "algorithmic bytes = the fp32 table read once per half-iteration (SURVEY 8d); the "
                                    "iteration works on fp64 survivor lists (`traffic` = its DRAM bytes from ncu) and "
                                    "is bound by instruction issue, so frac > 1 is possible; launch_ms = mean "
                                    "iteration time, (assignment T=50 - assignment T=10) / 40"}}
        if do_cpu:
            # parity on the TIMED batch: the last timed step's codes against the oracle on the same 8192 rows
            xh = xs[W + Ksteps - 1].cpu().numpy()
            ch = cen.detach().cpu().numpy()
            t0 = time.perf_counter()
            ro = O.constrained_assign(xh, ch, SK_EPS, SK_ITERS)
            dt = time.perf_counter() - t0
            got = codes_last.cpu().numpy()
            rec["cpu_baseline"] = {"value": ASSIGN_B / dt, "unit": "embeddings/s", "cores": cores, "kind": "port",
                                   "sample": f"the timed batch: {ASSIGN_B} embeddings, M={M}, T={SK_ITERS} "
                                             "(oracle C/OpenMP)",
                                   "gpu_matches_cpu_bit_exact": bool(np.array_equal(got, ro["codes"])),
                                   "code_mismatches": int((got != ro["codes"]).sum())}
            if M == 48:
                # the imported reference module beside it, on a smaller batch so the run stays short
                res = reference_quantize_seconds(2048, M, cores)
                if res is not None:
                    dt_ref, codes_ref, x_ref, c_ref = res
                    got_ref = ops.constrained_assign(torch.from_numpy(x_ref).to(dev), torch.from_numpy(c_ref).to(dev),
                                                     SK_EPS, SK_ITERS).cpu().numpy()
                    rec["cpu_reference_module"] = {
                        "value": 2048 / dt_ref, "unit": "embeddings/s", "cores": cores, "kind": "reference",
                        "sample": f"RepCONC.quantize imported from baseline/_ref, 2048 x {D}, M={M}, T={SK_ITERS}",
                        "gpu_matches_reference_bit_exact": bool(np.array_equal(got_ref, codes_ref))}
        del xs, xs_host
        torch.cuda.empty_cache()
        return rec

    assign = run_assign(C2_M, 1000, "BASELINE configs[2]") if "assign" in only else None
    assign_c5 = run_assign(C5_M, 2000, "BASELINE configs[4]: global batch 8192 x n_gpus, M=96") \
        if "assign_c5" in only else None

    if clk_started[0]:
        clk.__exit__(None, None, None)
    clocks = clk.summary()

    def adc_record(rec, workload, scaling, n_docs):
        return {"metric": "adc_queries_per_sec", "value": rec["qps"], "unit": "queries/s",
                "ms_per_step": rec["ms"] / Ksteps, "scaling": scaling,
                "config": {"workload": workload, "docs_total": n_docs, "docs_per_gpu": rec["n_shard"],
                           "queries_per_step": SEARCH_BATCH, "topk": TOPK},
                "e2e": {"value": rec["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": SEARCH_BATCH * D * 4,
                        "d2h_bytes_per_step": SEARCH_BATCH * TOPK * 12},
                "gpu_launches": rec["launches"], "roofline": adc_roofline(rec, clocks["sm_mhz"]),
                "cpu_baseline": rec["cpu"], "search_stats": rec["stats"], "mrr_at_10": rec["mrr"]}

    if rank == 0:
        line = {
            "metric": "adc_queries_per_sec", "value": adc["qps"], "unit": "queries/s", "n_gpus": world,
            "steps": Ksteps, "warmup": W, "ms_per_step": adc["ms"] / Ksteps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ADC_WORKLOAD},
            "details": {"docs_per_gpu": adc["n_shard"], "queries_per_step": SEARCH_BATCH, "topk": TOPK,
                        "parallelism": (f"index replicated on {world} GPUs, queries split (the reference's "
                                        "index_cpu_to_all_gpus mode, evaluate_repconc.py:130-134); corpus-sharded "
                                        "search of the same corpus in `adc_sharded`") if world > 1 else "single GPU",
                        "corpus": "SURVEY 8(d): N(0,I) documents coded by rc_nn_assign, queries = document + 0.5 noise",
                        "l2_policy": "inputs larger than L2 (corpus 424 MB vs 126 MB L2); fresh queries every step"},
            "e2e": {"value": adc["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": SEARCH_BATCH * D * 4,
                    "d2h_bytes_per_step": SEARCH_BATCH * TOPK * 12},
            "gpu_launches": adc["launches"], "roofline": adc_roofline(adc, clocks["sm_mhz"]),
            "cpu_baseline": adc["cpu"], "clocks": clocks, "search_stats": adc["stats"], "mrr_at_10": adc["mrr"],
            "assign": assign, "assign_c5": assign_c5,
            "adc_sharded": adc_record(adc_sharded, ADC_WORKLOAD + f", corpus sharded over {world} GPUs (all queries "
                                      "on every rank, all_to_all + owner merge)", "strong", C2_DOCS)
            if adc_sharded else None,
            "adc_c4": adc_record(adc_c4, "corpus-sharded ADC top-1000, 8,000,000 docs per GPU x 768-d, M=32 K=256 "
                                         "(BASELINE configs[3]: 64 M docs on 8 GPUs)", "weak",
                                 C4_DOCS_PER_GPU * world) if adc_c4 else None,
        }
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

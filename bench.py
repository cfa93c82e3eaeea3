#!/usr/bin/env python
"""bench.py -- headline benchmark of the RepCONC constrained-PQ hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  Primary metric (BASELINE.json configs[1]): PQ asymmetric-distance
top-1000 search, 8,841,823 synthetic documents, 768-d, M=48, K=256; a "step" is one
`search_batch` of 1200 queries (evaluate_repconc.py:38) against the whole corpus.
  value     queries/s with queries and corpus resident in HBM (CUDA events, max over ranks)
  e2e       the same through the reference-facing `batch_search` with HOST numpy queries/results
  roofline  the filtered-scan kernel: algorithmic bytes (N*M + 4*M*K + 12*k per query)
            / CUDA-event duration of that kernel, against MEASURED_PEAKS.json
  cpu_baseline  the oracle's C/OpenMP restatement of the Faiss IndexPQ scan on a bounded sample
Secondary (`assign`, BASELINE.json configs[2]): constrained-assignment training step, 8192
embeddings per GPU, M=48, Sinkhorn 50 iterations + decode + MSE forward/backward, embeddings/s.
N > 1: the corpus is sharded N ways (strong scaling of the same 8.84M-doc search, all_gather +
merge of the per-shard top-k); the assignment batch is 8192 per rank with the reference's
all-reduce of the row sums every iteration (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (NCCL prints its version banner there)

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DOCS = 8_841_823          # MS MARCO passage count
D, M, K, TOPK = 768, 48, 256, 1000
SEARCH_BATCH = 1200         # evaluate_repconc.py:38
ASSIGN_B, SK_EPS, SK_ITERS, MSE_W = 8192, 0.003, 50, 1e-4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def make_corpus(torch, device, lo, hi, seed=1234):
    """uniform 8-bit codes (what the equal-cluster constraint produces); rows [lo, hi) of a corpus whose
    content depends only on the global row index block, so any sharding sees the same documents"""
    blk = 1 << 20
    out = torch.empty((hi - lo, M), dtype=torch.uint8, device=device)
    b0 = lo // blk
    for b in range(b0, (hi + blk - 1) // blk):
        g = torch.Generator(device=device).manual_seed(seed + b)
        rows = torch.randint(0, 256, (blk, M), generator=g, device=device, dtype=torch.uint8)
        s, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        out[s - lo:e - lo] = rows[s - b * blk:e - b * blk]
    return out


def run_reference(args):
    """The reference's CPU path on the box's host cores: the oracle's C/OpenMP restatement of the Faiss
    IndexPQ inner-product scan + heap top-k (Faiss itself is not installable here), on the
    same corpus / query distribution, each step a bounded sample of the 1200-query batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    O.set_num_threads(cores)
    r = np.random.default_rng(7)
    nq = max(4 * cores, 32)
    codes = r.integers(0, 256, size=(N_DOCS, M), dtype=np.uint8)
    c = r.standard_normal((M, K, D // M), dtype=np.float32)
    times = []
    for it in range(args.warmup + args.steps):
        q = r.standard_normal((nq, D), dtype=np.float32)
        t0 = time.perf_counter()
        O.adc_search(q, c, codes, TOPK)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = nq * len(times) / total
    sample = f"{nq} of the {SEARCH_BATCH} queries per step x {N_DOCS} docs, k={TOPK}"
    print(json.dumps({
        "impl": "reference", "metric": "adc_queries_per_sec", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ADC top-1000, 8,841,823 docs x 768-d, M=48 K=256 (BASELINE configs[1])",
                   "search_batch": SEARCH_BATCH, "sampled_queries_per_step": nq},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-assign", action="store_true", help="skip the secondary (assign) workload")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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

    # ------------------------------------------------------------------ ADC (primary)
    lo, hi = E.shard_bounds(N_DOCS, rank, world)
    codes = make_corpus(torch, dev, lo, hi)
    gq = torch.Generator(device=dev).manual_seed(99)
    cent = torch.randn((M, K, D // M), generator=gq, device=dev)
    index = GpuIndexPQ(codes, cent, id_offset=lo)
    q_dev = torch.randn((W + Ksteps, SEARCH_BATCH, D), generator=gq, device=dev)   # same on every rank
    q_host = q_dev.cpu().numpy()
    corpus_ids = np.arange(N_DOCS, dtype=np.int64)
    qids = np.arange(SEARCH_BATCH)

    def adc_step(i):
        if world > 1:
            return E.sharded_search(index, q_dev[i], TOPK)
        return index.search_tensor(q_dev[i], TOPK)

    def adc_step_e2e(i):
        if world > 1:
            s, ids = E.sharded_search(index, q_host[i], TOPK)      # host queries in, merged on the GPU
            return s.cpu().numpy(), ids.cpu().numpy()
        return E.batch_search(qids, q_host[i], corpus_ids, index, TOPK, SEARCH_BATCH)

    for i in range(W):
        adc_step(i)
    barrier()
    scan_ms, scan_launches = 0.0, 0
    l0 = _lib.launch_count()
    # clocks / throttle reasons are sampled from here to the end of the last GPU-timed region (ADC, ADC e2e,
    # assign, assign e2e): every sample is taken under load
    clk = ClockSampler(local)
    clk.__enter__()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + Ksteps):
        adc_step(i)
        scan_ms += lib.rc_adc_last_scan_ms()
        scan_launches += lib.rc_adc_last_scan_launches()
    e1.record()
    barrier()
    adc_launches = _lib.launch_count() - l0
    adc_ms = max_over_ranks(e0.elapsed_time(e1))
    adc_qps = Ksteps * SEARCH_BATCH / (adc_ms / 1e3)
    stats = index.last_stats
    # e2e: host numpy queries -> batch_search -> host numpy results.  N = 1: ONE batch_search call over the K
    # steps' queries with batch_size 1200, as the reference's evaluator issues it (evaluate_repconc.py:188-206):
    # every step's queries go host -> device and every step's results device -> host inside the timed region,
    # the copy-back of one batch overlapping the scan of the next.  N > 1: one sharded search per step.
    for i in range(min(W, 2)):
        adc_step_e2e(i)
    if world == 1:
        q_all = np.ascontiguousarray(q_host[W:W + Ksteps].reshape(-1, D))
        qids_all = np.arange(len(q_all))
        E.batch_search(qids_all, q_all, corpus_ids, index, TOPK, SEARCH_BATCH)      # warm-up (staging buffers)
        barrier()
        t0 = time.perf_counter()
        out = E.batch_search(qids_all, q_all, corpus_ids, index, TOPK, SEARCH_BATCH)
        barrier()
        e2e_s = time.perf_counter() - t0
        assert out[0].shape == (Ksteps * SEARCH_BATCH, TOPK)
    else:
        barrier()
        t0 = time.perf_counter()
        for i in range(W, W + Ksteps):
            out = adc_step_e2e(i)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        assert out[0].shape == (SEARCH_BATCH, TOPK)
    adc_e2e_qps = Ksteps * SEARCH_BATCH / e2e_s
    # roofline of the filtered scan (per-GPU shard)
    n_shard = hi - lo
    alg_per_query = n_shard * M + 4 * M * K + 12 * TOPK
    alg_per_launch = alg_per_query * (Ksteps * SEARCH_BATCH) / max(scan_launches, 1)
    scan_avg_s = scan_ms / 1e3 / max(scan_launches, 1)
    achieved = alg_per_launch / scan_avg_s / 1e9 if scan_avg_s > 0 else 0.0
    # dram__bytes_read + dram__bytes_write of ONE launch (1200 queries x 8,841,823 docs) from the committed
    # `ncu --set full` capture profiles/r01_adc_scan_cf_ncu.txt; scaled to this run's shard size
    ncu_traffic = (520.220672e6 + 19.589632e6) * (n_shard / N_DOCS) if scan_launches == Ksteps else None
    roofline = {"bound": "hbm", "kernel": "adc_scan_cf_kernel<48>", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": scan_avg_s * 1e3,
                "share_of_step": scan_ms / adc_ms,
                "note": "algorithmic bytes = every query streams every code byte once (SURVEY 8d); queries are "
                        "tiled 8 per CTA and a split's codes stay L2-resident, so DRAM traffic is ~900x below "
                        "this; the kernel is bound by shared-memory bandwidth (ncu: LSU wavefronts 98% of peak "
                        "with a bank-conflict-free table layout), not by HBM"}

    # ------------------------------------------------------------------ constrained assign (secondary)
    assign = None
    if not args.skip_assign:
        ga = torch.Generator(device=dev).manual_seed(1000 + rank)
        gc = torch.Generator(device=dev).manual_seed(5)
        cen = torch.randn((M, K, D // M), generator=gc, device=dev).requires_grad_(True)
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
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(W, W + Ksteps):
            assign_step(xs[i])
        e1.record()
        barrier()
        as_launches = _lib.launch_count() - l0
        as_ms = max_over_ranks(e0.elapsed_time(e1))
        as_eps = Ksteps * ASSIGN_B * world / (as_ms / 1e3)
        # e2e: every step's embeddings come from pinned host memory and its codes + loss go back to the host;
        # the H2D of step i+1 is issued on a copy stream while step i computes (double buffering)
        copy_stream = torch.cuda.Stream(device=dev)

        def fetch(i):
            with torch.cuda.stream(copy_stream):
                xd = xs_host[i].to(dev, non_blocking=True)        # H2D of the step's embeddings (pinned)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return xd, ev

        def assign_e2e_loop(first, last):
            nxt = fetch(first)
            for i in range(first, last):
                xd, ev = nxt
                if i + 1 < last:
                    nxt = fetch(i + 1)
                torch.cuda.current_stream().wait_event(ev)
                c_, mse_ = assign_step(xd)
                res = c_.cpu(), float(mse_.detach())              # D2H of the codes and the loss
            return res

        assign_e2e_loop(0, W)                                     # e2e warm-up (allocator reaches steady state)
        barrier()
        t0 = time.perf_counter()
        c_host, mse_host = assign_e2e_loop(W, W + Ksteps)
        barrier()
        as_e2e = Ksteps * ASSIGN_B * world / max_over_ranks(time.perf_counter() - t0)
        # the dominant kernels: one Sinkhorn iteration = survivor-list pass (or re-selection) + reduce/update.
        # N = 1: the product path is rc_sinkhorn_solve, so the per-iteration time is the difference of two solves
        # (50 and 10 iterations) on the launching stream; N > 1: the step-wise entry point (the exchange of the
        # row sums between steps is timed by the step as a whole, not here).
        kern = ops.CudaAssignKernels(xs[0], cen)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world == 1:
            def solve_ms(T, reps=5):
                ts = []
                for _ in range(reps):
                    kern.table()
                    torch.cuda.synchronize()
                    p0.record()
                    kern.solve(SK_EPS, T)
                    p1.record()
                    torch.cuda.synchronize()
                    ts.append(p0.elapsed_time(p1))
                return sorted(ts)[len(ts) // 2]
            solve_ms(SK_ITERS, 2)
            it_s = (solve_ms(SK_ITERS) - solve_ms(10)) / 1e3 / (SK_ITERS - 10)
        else:
            kern.table()
            kern.begin(SK_EPS)
            kern.step(SK_EPS, ASSIGN_B * world)
            kern.step(SK_EPS, ASSIGN_B * world)
            torch.cuda.synchronize()
            n_it = 20
            p0.record()
            for _ in range(n_it):
                kern.step(SK_EPS, ASSIGN_B * world)
            p1.record()
            torch.cuda.synchronize()
            it_s = p0.elapsed_time(p1) / 1e3 / n_it
        alg_it = 2 * 4 * M * K * ASSIGN_B           # the table is read once per half-iteration
        assign = {"metric": "constrained_assign_embeddings_per_sec", "value": as_eps, "unit": "embeddings/s",
                  "ms_per_step": as_ms / Ksteps, "scaling": "weak",
                  "config": {"workload": "training-step loop (BASELINE configs[2])", "batch_per_gpu": ASSIGN_B,
                             "M": M, "K": K, "sk_iters": SK_ITERS, "sk_epsilon": SK_EPS,
                             "step": "table + centring + Sinkhorn + argmax + decode + MSE fwd/bwd"},
                  "e2e": {"value": as_e2e, "unit": "embeddings/s", "h2d_bytes_per_step": ASSIGN_B * D * 4,
                          "d2h_bytes_per_step": ASSIGN_B * M * 8 + 4},
                  "gpu_launches": as_launches,
                  "roofline": {"bound": "hbm", "kernel": "sinkhorn iteration (sinkhorn_step_list_kernel / "
                                                          "sinkhorn_step_sparse_kernel + sinkhorn_reduce_update_kernel)",
                               "achieved": alg_it / it_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": alg_it / it_s / 1e9 / hbm_peak,
                               # dram__bytes_read + dram__bytes_write of ONE sinkhorn_step_list_kernel launch at
                               # this shape, from the committed `ncu --set full` capture
                               # profiles/r01_sinkhorn_list_ncu.txt
                               "traffic": 250.099968e6 + 4.836608e6,
                               "algorithmic_bytes_per_launch": alg_it, "launch_ms": it_s * 1e3,
                               "note": "algorithmic bytes = the fp32 table read once per half-iteration (SURVEY 8d); "
                                       "the iteration works on fp64 survivor lists (~190 MB per pass, ncu) and is "
                                       "bound by instruction issue, not HBM; launch_ms = mean over list and "
                                       "re-selection iterations incl. the reduce/update kernel"}}

    clk.__exit__(None, None, None)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N == 1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        O.set_num_threads(cores)
        nq_cpu = max(8 * cores, 32)
        codes_h = codes.cpu().numpy()
        c_h = cent.cpu().numpy()
        qh = q_host[W][:nq_cpu]
        t0 = time.perf_counter()
        so, io = O.adc_search(qh, c_h, codes_h, TOPK)
        dt = time.perf_counter() - t0
        sg, ig = index.search(qh, TOPK)
        parity = bool(np.array_equal(io, ig) and np.array_equal(so, sg))
        cpu_baseline = {"value": nq_cpu / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                        "sample": f"{nq_cpu} queries x {N_DOCS} docs, k={TOPK} (oracle C/OpenMP scan + heap)",
                        "gpu_matches_cpu_bit_exact": parity}
        if assign is not None:
            nb = 1024
            xh = xs[W][:nb].cpu().numpy()
            t0 = time.perf_counter()
            ro = O.constrained_assign(xh, cen.detach().cpu().numpy(), SK_EPS, SK_ITERS)
            dt = time.perf_counter() - t0
            got = ops.constrained_assign(xs[W][:nb], cen, SK_EPS, SK_ITERS).cpu().numpy()
            assign["cpu_baseline"] = {"value": nb / dt, "unit": "embeddings/s", "cores": cores, "kind": "port",
                                      "sample": f"{nb} embeddings, M={M}, T={SK_ITERS} (oracle C/OpenMP)",
                                      "gpu_matches_cpu_bit_exact": bool(np.array_equal(got, ro["codes"]))}

    clocks = clk.summary()
    if rank == 0:
        line = {
            "metric": "adc_queries_per_sec", "value": adc_qps, "unit": "queries/s", "n_gpus": world,
            "steps": Ksteps, "warmup": W, "ms_per_step": adc_ms / Ksteps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ADC top-1000, 8,841,823 docs x 768-d, M=48 K=256, 1200-query search batches "
                                   "(BASELINE configs[1])",
                       "docs_per_gpu": n_shard, "queries_per_step": SEARCH_BATCH, "topk": TOPK,
                       "parallelism": f"corpus-sharded x{world}" if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (corpus 424 MB vs 126 MB L2); fresh queries every step"},
            "e2e": {"value": adc_e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": SEARCH_BATCH * D * 4,
                    "d2h_bytes_per_step": SEARCH_BATCH * TOPK * 12},
            "gpu_launches": adc_launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "search_stats": stats, "assign": assign,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key counters per profiled launch -> text.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu.txt"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# ncu summary of {rep}  (ncu --set full --clock-control none; cold-cache, serialised launches)")
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print(f"\nkernel: {d.get('Kernel Name')}")
    for k in KEYS:
        if d.get(k) not in (None, ''):
            print(f"  {k:75s} {d[k]} {u.get(k, '')}")
    st = []
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k:
            try:
                st.append((float(d[k]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print("  top warp stall reasons (warps per issue-active cycle): " +
          ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))

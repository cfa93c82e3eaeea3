#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernels, from `ncu --set full` captures of THIS build -> profiles/r02_traffic.json
(read by bench.py as roofline.traffic).  Run here (no GPU needed) after the captures came back from the GPU box:

    gpurun -- 'ncu --set full --clock-control none --import-source on -k regex:adc_scan_u8_kernel -c 2 -f \
               -o gpurun_out/cap/adc_m48 python tools/prof_adc.py'                 (PM / PN select M and the corpus size)
    gpurun -- 'ncu --set full ... -k regex:sinkhorn_step_kernel -s 3 -c 1 -f -o gpurun_out/cap/list_m48 \
               python tools/prof_assign.py 12 8192 48'
    python tools/ncu_traffic.py adc:gpurun_out/cap/adc_m48.ncu-rep:8841823 list:gpurun_out/cap/list_m48.ncu-rep:48 ...
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_traffic.json")


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, rows[1]))


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v) * scale


def main():
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for spec in sys.argv[1:]:
        kind, rep, arg = spec.split(":")
        ls, units = launches(rep)
        # the longest launch of the capture is the kernel of interest (the other one is the threshold-sample scan)
        d = max(ls, key=lambda r: float(r["gpu__time_duration.sum"]))
        traffic = to_bytes(d["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + \
            to_bytes(d["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
        if kind == "adc":
            name = d["Kernel Name"]
            q = "16 queries/entry" if ", 16," in name else "8 queries/entry"
            key = f"adc_scan_u8_kernel<{q}>@{arg}"
        else:
            key = f"sinkhorn_iteration@M{arg}"
        out[key] = traffic
        print(key, f"{traffic / 1e6:.1f} MB  ({d['Kernel Name'][:60]}, {d['gpu__time_duration.sum']} {units['gpu__time_duration.sum']})")
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into small, committed text summaries under profiles/.
    python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.txt
    python tools/ncu_summary.py report   gpurun_out/r01_gemm.ncu-rep profiles/r01_gemm_ncu.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    total = 0.0
    for r in rows[hdr_i + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        t = float(r[vi].replace(",", "")) / 1e3  # ns -> us
        name = r[ki].split("(")[0][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n# source: {src}\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:72s} {n:8d} {t:12.1f} {t / n:10.1f} {t / total:7.3f}\n")
        f.write(f"{'TOTAL':72s} {sum(a[0] for a in agg.values()):8d} {total:12.1f}\n")
    print(open(dst).read())


def report(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source: {src}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[name_i][:110]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:70s} {r[i]:>16s} {units[i]}\n")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            try:
                f.write(f"  {'traffic = dram read + write':70s} {float(r[rd]) + float(r[wr]):16.3f} {units[rd]}\n")
            except ValueError:
                pass
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])

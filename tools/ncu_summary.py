#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_bench_launches.md
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep  profiles/r01_layer_tc_full.md [kernel-substring]

`launches`: the `--metrics gpu__time_duration.sum --clock-control none --csv` launch list of one
bench.py command -> per-kernel count / total / share table (times are cold-cache, serialised;
only the SHARES are meaningful).  `report`: key metrics of one `--set full` capture.
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def short(name):
    name = name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if name.startswith("at::"):
        return name[:70]
    name = name.split("(")[0]
    return name.split("<")[0].split("::")[-1][:70] + ("<...>" if "<" in name else "")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + float(r[vi].replace(",", "")))
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % src)
        f.write("Per-launch times under ncu are cold-cache and serialised: only each kernel's SHARE is meaningful.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, c, t / 1e3, 100 * t / tot))
        f.write("\ntotal: %d launches, %.2f ms\n" % (sum(c for c, _ in agg.values()), tot / 1e6))


def report(src, dst, match=""):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (%s)\n\n" % src)
        for r in rows[2:]:
            if match and match not in r[ki]:
                continue
            f.write("## %s  (grid %s, block %s)\n\n| metric | value | unit |\n|---|---:|---|\n"
                    % (short(r[ki]), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, r[i], units[i]))
            f.write("\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")

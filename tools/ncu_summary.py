#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep  profiles/X_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    # atomics: shared-memory (tile histograms / ranks), global ATOM (returning) and RED (non-returning), L2 side
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum",
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}) — gpu__time_duration.sum, --clock-control none (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {name} | {c} | {t:.1f} | {t / c:.1f} | {t / total:.3f} |\n")
        f.write(f"\ntotal {total:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full ({src}), --clock-control none\n\n")
        for r in rows[2:]:
            f.write(f"## {r[h.index('Kernel Name')].split('(')[0]}  (launch id {r[h.index('ID')]})\n\n")
            for m in METRICS:
                if m in h:
                    f.write(f"- {m} = {r[h.index(m)]} {units[h.index(m)]}\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + per-opcode and hottest-instruction tables.

usage: tools/ncu_summary.py gpurun_out/prof_raster.ncu-rep > profiles/<name>.txt
"""
import csv
import os
import io
import subprocess
import sys
from collections import defaultdict


def select_kernel(rows):
    """multi-kernel reports: keep the first source-page section of the kernel named by NCU_KERNEL"""
    k = os.environ.get("NCU_KERNEL")
    if not k:
        return rows
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for n, i in enumerate(starts):
        if k in rows[i][1]:
            end = starts[n + 1] if n + 1 < len(starts) else len(rows)
            return rows[i:end]
    raise SystemExit(f"kernel {k} not in report")


rep = sys.argv[1]


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, units = raw[0], raw[1]
ki = hdr.index("Kernel Name")
vals = next((r for r in raw[2:] if os.environ.get("NCU_KERNEL", "") in r[ki]), raw[2])
want = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]
print(f"# {rep}")
for h, u, v in zip(hdr, units, vals):
    if any(h == w or h.endswith("." + w) for w in want):
        print(f"{h:95s} {v} {u}")
print()
print("## warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active / pcsamp)")
for h, u, v in zip(hdr, units, vals):
    if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            if float(v) > 0:
                print(f"{h:95s} {v}")
        except ValueError:
            pass

src = select_kernel(list(csv.reader(io.StringIO(ncu("source")))))
# find header row
hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
cols = {n: i for i, n in enumerate(src[hi])}
rows = src[hi + 1 :]
op_exec = defaultdict(int)
op_samp = defaultdict(int)
tot_exec = tot_samp = 0
inst = []
for r in rows:
    if len(r) <= cols["Instructions Executed"]:
        continue
    sass = r[cols["Source"]].strip()
    parts = sass.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
    op = op.split(".")[0]
    ex = int(float(r[cols["Instructions Executed"]] or 0))
    sm = int(float(r[cols["Warp Stall Sampling (All Samples)"]] or 0))
    op_exec[op] += ex
    op_samp[op] += sm
    tot_exec += ex
    tot_samp += sm
    inst.append((sm, ex, sass))
print()
print("## per-opcode: warp instructions executed, share, stall samples share")
for op, ex in sorted(op_exec.items(), key=lambda kv: -kv[1])[:28]:
    print(f"{op:12s} {ex:14d} {100.0 * ex / max(tot_exec, 1):6.2f}%   samples {100.0 * op_samp[op] / max(tot_samp, 1):6.2f}%")
print()
print("## 40 hottest instructions by stall samples")
for sm, ex, sass in sorted(inst, reverse=True)[:40]:
    print(f"{100.0 * sm / max(tot_samp, 1):6.2f}%  exec {ex:12d}  {sass}")

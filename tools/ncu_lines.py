#!/usr/bin/env python
"""Attribute ncu per-instruction samples to CUDA source lines (SASS offsets -> nvdisasm -g line table).

usage: tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled kernel name substring> [top N]
"""
import csv, io, os, re, subprocess, sys, tempfile
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


rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur = None
infunc = False
for ln in dis.splitlines():
    m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kname in ln
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infunc:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = select_kernel(list(csv.reader(io.StringIO(src))))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = {n: i for i, n in enumerate(rows[hi])}
data = [r for r in rows[hi + 1 :] if len(r) > cols["Instructions Executed"]]
base = min(int(r[0], 16) for r in data)
by_line = defaultdict(lambda: [0, 0, 0.0])
tot_s = tot_e = 0
for r in data:
    off = int(r[0], 16) - base
    key = line_of.get(off)
    s_ = int(float(r[cols["Warp Stall Sampling (All Samples)"]] or 0))
    e_ = int(float(r[cols["Instructions Executed"]] or 0))
    by_line[key][0] += s_
    by_line[key][1] += e_
    by_line[key][2] += float(r[cols["Predicated-On Thread Instructions Executed"]] or 0)
    tot_s += s_
    tot_e += e_
srcs = {}
def text(key):
    if key is None:
        return "?"
    f, l = key
    for root in ("osm_renderer_b200/csrc", "include"):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            if l - 1 < len(srcs[p]):
                return srcs[p][l - 1].strip()[:100]
    return ""
print(f"# {rep}: samples {tot_s}, warp instructions {tot_e}; share of samples / share of warp instructions / avg active threads per source line")
for key, (s_, e_, t_) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*s_/max(tot_s,1):6.2f}% {100*e_/max(tot_e,1):6.2f}% thr {t_/max(e_,1):5.1f}  {key[0] if key else '?'}:{key[1] if key else 0:<5d} {text(key)}")

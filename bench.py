#!/usr/bin/env python
"""bench.py -- tiles/s of the per-tile draw path (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

Workloads (--workload):
    C2   1024 tiles 256x256, z=14, x 9888..9919, y 5104..5135 over the seeded synthetic metro `.bin` (the metric's config, default)
    C3   the same batch @2x (512x512)
    C4   zoom sweep z=10..18 over the same `.bin`, <= 1024 tiles per zoom (seeded sample), per-zoom tiles/s
    C5   z=16, 16x16 tile block with 10^5 extra building footprints + a 50k-node coastline multipolygon with 200 islands
    C2r1 round 1's sparser C2 dataset (3-8 rectangular footprints per block, one way per block side), for continuity
--style mapnik (tests/mapcss/mapnik.mapcss, the reference's test stylesheet, default) | osmosnimki (mapcss/osmosnimki-minimal.mapcss,
the stylesheet of the reference's README.md:23-25); both JOSM flavour.

One "step" = one pass of the draw path over the batch on every rank.  --scaling weak (default): every rank renders its own
full batch (request i of the interleaved global list goes to rank i mod N).  --scaling strong: ONE list of 8 x batch requests
is dealt i mod N, so the work per rank shrinks with N.  Tiles are independent; no collective touches the data path.

JSON line: `value` = tiles/s of the whole Drawer::draw_to_pixels (area passes + label pass) with the batch description resident
in HBM and the output left in HBM (wall clock between barriers; `device_only` is the CUDA-event figure); `e2e` = the same through
the reference-facing call osmr_draw_tiles_labeled with pinned HOST buffers, copies inside the timed region; siblings
`value_area_only` / `e2e_area_only` (Fill, Casing, Stroke passes only: round 1's headline), `e2e_auto` (tile list in),
`e2e_png` / `e2e_auto_png` (PNG files out), `e2e_auto_labeled` / `e2e_auto_labeled_png` (tile list in, label pass included), `latency_ms` (p50 of single small calls);
`sustained` = the resident leg repeated for >= --min-seconds with median / min per step; `roofline` = dominant kernel against
the measured HBM peak with SURVEY 8(d)'s B_tile next to the kernel-local byte count; `cpu_baseline` = the oracle (C++
restatement of the reference CPU path) on this box's host cores, T = nproc and T = 1.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dict(zoom, x0, y0, n, scale, metro=kwargs of synth.make_metro)
    "C2": dict(zoom=14, x0=9888, y0=5104, n=32, scale=1, metro={}),
    "C2r1": dict(zoom=14, x0=9888, y0=5104, n=32, scale=1, metro={"profile": "r1"}),
    "C3": dict(zoom=14, x0=9888, y0=5104, n=32, scale=2, metro={}),
    "C4": dict(zoom=None, x0=9888, y0=5104, n=32, scale=1, metro={}),  # z10..18 over the C2 dataset
    "C5": dict(zoom=16, x0=9888 * 4 + 56, y0=5104 * 4 + 56, n=16, scale=1,
               metro={"zoom": 16, "x0": 9888 * 4 + 56, "y0": 5104 * 4 + 56, "n": 16, "extra_footprints": 100000, "coastline_nodes": 50000}),
}
STYLES = {"mapnik": "mapnik_rules.json.gz", "osmosnimki": "osmosnimki_rules.json.gz"}
HEADLINE_LABELED = True  # `value` / `e2e` are the WHOLE Drawer::draw_to_pixels (area passes + label pass, drawer.rs:60-131), as the
# reference always runs it; the area passes alone (round 1's headline) are reported as value_area_only / e2e_area_only


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------
def c4_tiles(seed=0xB20005A1):
    """for z = 10..18 the tiles covering the bbox of the C2 block, at most 1024 per zoom (seeded sample)."""
    rng = np.random.default_rng(seed)
    out = []
    for z in range(10, 19):
        if z <= 14:
            sh = 14 - z
            xa, xb = 9888 >> sh, 9919 >> sh
            ya, yb = 5104 >> sh, 5135 >> sh
        else:
            sh = z - 14
            xa, xb = 9888 << sh, ((9919 + 1) << sh) - 1
            ya, yb = 5104 << sh, ((5135 + 1) << sh) - 1
        w, h = xb - xa + 1, yb - ya + 1
        if w * h <= 1024:
            cells = [(x, y) for y in range(ya, yb + 1) for x in range(xa, xb + 1)]
        else:
            pick = np.sort(rng.choice(w * h, size=1024, replace=False))
            cells = [(xa + int(p % w), ya + int(p // w)) for p in pick]
        out.append((z, cells))
    return out


def build_workload(name: str, style: str = "mapnik", labels: bool = False, max_tiles: int | None = None):
    """Returns dict(bin, reader, builder, table, ltable, tiles, area_begin, areas, [label_begin, labels], canvas, caps, scale, groups).
    groups: list of (zoom, first tile, n tiles) -- one entry except for C4."""
    from osm_renderer_b200.upstream import geodata, mapcss, pipeline, styler as st, synth
    from osm_renderer_b200.wire import LABEL_DTYPE, TILE_DTYPE, LabelStyleTable, StyleTable

    spec = WORKLOADS[name]
    scale = spec["scale"]
    t0 = time.time()
    data = synth.make_metro(**spec["metro"])
    rd = geodata.GeodataReader(data)
    rules = mapcss.load_rules_json(os.path.join(ROOT, "tests", "golden", STYLES[style]))
    S = st.Styler(rules, "josm", None)
    table = StyleTable(None)  # fill-image icons are not shipped: such areas are skipped like a failed icon load
    ltable = LabelStyleTable(None)
    fb = pipeline.FastBatchBuilder(rd, S, table)
    if name == "C4":
        per_zoom = c4_tiles()
    else:
        z, x0, y0, n = spec["zoom"], spec["x0"], spec["y0"], spec["n"]
        per_zoom = [(z, [(x, y) for y in range(y0, y0 + n) for x in range(x0, x0 + n)])]
    if max_tiles is not None:  # tests: a spread sample of every group
        per_zoom = [(z, [cells[i] for i in np.unique(np.linspace(0, len(cells) - 1, min(max_tiles, len(cells))).astype(int))]) for z, cells in per_zoom]
    tiles, groups = [], []
    for z, cells in per_zoom:
        groups.append((z, len(tiles), len(cells)))
        tiles += [(z, x, y, scale) for (x, y) in cells]
    parts = [fb.areas_array(z, x, y) for (z, x, y, s) in tiles]
    begins = np.zeros(len(tiles) + 1, dtype=np.uint32)
    begins[1:] = np.cumsum([len(p) for p in parts])
    areas = np.concatenate(parts)
    w = {
        "name": name, "style": style, "bin": data, "reader": rd, "builder": fb, "styler": S, "table": table, "ltable": ltable,
        "tiles": np.array(tiles, dtype=TILE_DTYPE), "area_begin": begins, "areas": areas,
        "canvas": S.canvas_fill_color, "caps": S.use_caps_for_dashes, "scale": scale, "groups": groups,
    }
    if labels:
        lparts = [fb.labels_array(z, x, y, ltable) for (z, x, y, s) in tiles]
        lb = np.zeros(len(tiles) + 1, dtype=np.uint32)
        lb[1:] = np.cumsum([len(p) for p in lparts])
        w["label_begin"] = lb
        w["labels"] = np.concatenate(lparts) if lparts else np.zeros(0, dtype=LABEL_DTYPE)
        w["font"] = np.load(os.path.join(ROOT, "tests", "golden", "label_inputs.npz"))["font"].tobytes()
    log(f"[bench] workload {name}/{style}: {len(tiles)} tiles, {len(areas)} styled areas, {len(data) / 1e6:.0f} MB geodata, "
        f"{len(table.rows)} styles" + (f", {len(w['labels'])} label generations" if labels else "") + f", built in {time.time() - t0:.1f}s")
    return w


def sub_batch(w, sel, with_labels=False):
    """(tiles, area_begin, areas[, label_begin, labels]) of the tiles `sel` of workload w."""
    ab = w["area_begin"]
    parts = [w["areas"][ab[i]: ab[i + 1]] for i in sel]
    begins = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    out = [np.ascontiguousarray(w["tiles"][sel]), begins, np.concatenate(parts) if parts else w["areas"][:0]]
    if with_labels:
        lb = w["label_begin"]
        lparts = [w["labels"][lb[i]: lb[i + 1]] for i in sel]
        out += [np.concatenate([[0], np.cumsum([len(p) for p in lparts])]).astype(np.uint32),
                np.concatenate(lparts) if lparts else w["labels"][:0]]
    return out


def oracle_labels(w, label_begin, labels):
    """The ABI label list (entity, label style) as the records oracle.draw_tiles_with_labels takes (its own tag lookup is
    done here, once per distinct (entity, key)); returns (labels48, texts blob)."""
    import oracle
    from osm_renderer_b200.wire import OSMR_AREA_MULTIPOLYGON, OSMR_LABEL_NODE, OSMR_LSTYLE_FONT_SIZE, OSMR_LSTYLE_TEXT, OSMR_LSTYLE_TEXT_COLOR

    rows = w["ltable"].styles_array()
    strings = bytes(w["ltable"].strings)
    rd = w["reader"]
    ent, sty = labels["entity"], labels["style"]
    out = np.zeros(len(labels), dtype=oracle.LABEL_DTYPE)
    is_node = ((ent & OSMR_LABEL_NODE) != 0) & ((ent & OSMR_AREA_MULTIPOLYGON) == 0)
    is_mp = (ent & OSMR_AREA_MULTIPOLYGON) != 0
    out["kind"] = np.where(is_node, 2, np.where(is_mp, 1, 0))
    out["entity"] = ent & ~np.uint32(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE)
    r = rows[sty]
    out["icon"] = r["icon"]
    out["has_text_style"] = (r["flags"] & OSMR_LSTYLE_TEXT) != 0
    out["has_font_size"] = (r["flags"] & OSMR_LSTYLE_FONT_SIZE) != 0
    out["text_pos"] = r["text_position"]
    out["text_color"] = r["text_color"]
    out["has_text_color"] = (r["flags"] & OSMR_LSTYLE_TEXT_COLOR) != 0
    out["font_size"] = r["font_size"]
    texts = bytearray()
    toff: dict = {}
    memo: dict = {}
    for i in np.nonzero(out["has_text_style"])[0]:
        key = (int(out["kind"][i]), int(out["entity"][i]), int(r["text_key_off"][i]), int(r["text_key_len"][i]))
        hit = memo.get(key)
        if hit is None:
            k = strings[key[2]: key[2] + key[3]].decode("utf-8")
            tags = rd.node_tags(key[1]) if key[0] == 2 else (rd.multipolygon_tags(key[1]) if key[0] == 1 else rd.way_tags(key[1]))
            v = tags.get(k)
            if v is None:
                hit = (0, 0, 0)
            else:
                t = toff.get(v)
                if t is None:
                    b = v.encode("utf-8")
                    t = (len(texts), len(b))
                    texts += b
                    toff[v] = t
                hit = (1, t[0], t[1])
            memo[key] = hit
        out["has_text"][i], out["text_off"][i], out["text_len"][i] = hit
    return out, bytes(texts)


def distinct_nodes_per_batch(w) -> int:
    """U of SURVEY.md 8(d): distinct nodes referenced by the styled areas of each tile, summed over tiles."""
    rd = w["reader"]
    way_off = rd.ways["off"].astype(np.int64)
    way_len = rd.ways["len"].astype(np.int64)
    total = 0
    ab = w["area_begin"]
    ent = w["areas"]["entity"]
    for t in range(len(w["tiles"])):
        e = np.unique(ent[ab[t]: ab[t + 1]])
        e = e[e < 0x80000000].astype(np.int64)
        lens = way_len[e]
        tot = int(lens.sum())
        if tot == 0:
            continue
        rep = np.repeat(np.arange(len(e)), lens)
        start = np.concatenate([[0], np.cumsum(lens)[:-1]])
        idx = way_off[e][rep] + (np.arange(tot) - start[rep])
        total += len(np.unique(rd.ints[idx]))
    return total


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe), sampled every 50 ms from BEFORE the warm-up on; only the
    samples whose arrival time falls inside a timed region (mark()/unmark()) count."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.windows = []
        self._open = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        self._open = time.perf_counter()

    def unmark(self):
        if self._open is not None:
            self.windows.append((self._open, time.perf_counter()))
            self._open = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)  # let the last samples of the timed region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ts, r in self.rows:
            # a sample describes the ~50 ms before it arrived
            if not any(a <= ts <= b + 0.06 for a, b in self.windows):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "power_w_max": float(max(power)) if power else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
            "seconds_sampled": float(sum(b - a for a, b in self.windows)),
        }


# ----------------------------------------------------------------------------------------------------------
def cpu_port_throughput(w, sel, threads: int, labeled: bool = False):
    """tiles/s of the oracle (C++ restatement of the reference CPU path) on `threads` host threads over the tiles `sel`."""
    import oracle

    if labeled:
        tiles, begins, areas, lb, labels = sub_batch(w, sel, True)
        memo = w.setdefault("_oracle_labels", {})
        key = (len(sel), int(sel[0]), int(sel[-1]))
        if key not in memo:
            memo[key] = oracle_labels(w, lb, labels)
        l48, texts = memo[key]
        t0 = time.perf_counter()
        imgs = oracle.draw_tiles_with_labels(w["bin"], w["table"], tiles, begins, areas, w["canvas"], w["caps"], w["font"],
                                             w["ltable"].icons, lb, l48, texts, n_threads=threads)
    else:
        tiles, begins, areas = sub_batch(w, sel)
        t0 = time.perf_counter()
        imgs = oracle.draw_tiles(w["bin"], w["table"], tiles, begins, areas, w["canvas"], w["caps"], n_threads=threads)
    dt = time.perf_counter() - t0
    return len(sel) / dt, dt, imgs


def spread(n_tiles: int, n_sample: int) -> np.ndarray:
    return np.unique(np.linspace(0, n_tiles - 1, max(1, min(n_sample, n_tiles))).astype(int))


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def pin_to_gpu_cpus(local_rank: int):
    """Run this rank on the CPUs next to its GPU (and allocate the page-locked buffers from there afterwards): the raw-RGB
    end-to-end leg is bound by the host side of the D2H copies when several ranks share one socket."""
    info = {"applied": False}
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        numa = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        info.update({"gpu_pci": bdf, "gpu_numa_node": numa, "allowed_cpus": len(allowed), "gpu_local_cpus_allowed": len(use)})
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["applied"] = True
    except Exception as exc:
        info["error"] = str(exc)[:120]
    return info


def metric_name(w) -> str:
    D = 256 * w["scale"]
    if w["name"] == "C4":
        return f"{D}x{D} tiles/sec, zoom sweep z=10..18"
    return f"{D}x{D} tiles/sec at z={WORKLOADS[w['name']]['zoom']}"


def config_of(args, w, world) -> dict:
    spec = WORKLOADS[args.workload]
    D = 256 * spec["scale"]
    n_tiles = len(w["tiles"])
    desc = {
        "C2": f"C2: 1024 tiles {D}x{D} z=14 x9888..9919 y5104..5135",
        "C2r1": f"C2r1 (round-1 dataset): 1024 tiles {D}x{D} z=14 x9888..9919 y5104..5135",
        "C3": f"C3: 1024 tiles {D}x{D} (@2x) z=14 x9888..9919 y5104..5135",
        "C4": f"C4: zoom sweep z=10..18, {n_tiles} tiles {D}x{D} (<= 1024 per zoom, seeded sample)",
        "C5": f"C5: {n_tiles} tiles {D}x{D} z=16 (16x16 block), 10^5 extra footprints + 50k-node coastline multipolygon with 200 islands",
    }[args.workload]
    return {
        "workload": desc + f", synthetic metro .bin (seed 0xB20005A1), {args.style} stylesheet (JOSM)",
        "tiles_per_step_per_gpu": n_tiles if args.scaling == "weak" else None,
        "sharding": ("one full batch per rank (request i of the interleaved list -> rank i mod N)" if args.scaling == "weak"
                     else f"one list of {8 * n_tiles} requests dealt i mod N ({8 * n_tiles // world} per rank)") + "; tiles are independent, no data-path collective",
        "l2": "per-step working set (styled areas + plan scratch + walk cache + RGB output > 400 MB) exceeds the 126 MB L2",
    }


# ----------------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference is Rust and cannot be built in this image: its CPU algorithm is timed through the oracle port
    (oracle/osmr_oracle.cpp), all host threads, tiles dealt to the threads like src/http_server.rs:50-83,105-108."""
    if rank != 0:
        return
    labeled = HEADLINE_LABELED and args.workload != "C4"
    w = build_workload(args.workload, args.style, labels=True if args.workload != "C4" else False)
    threads = host_threads()
    n_tiles = len(w["tiles"])
    sample = args.cpu_sample or min(n_tiles, max(16 * threads, 256))  # several tiles per thread: the threads stay busy
    sel = spread(n_tiles, sample)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_throughput(w, spread(n_tiles, threads), threads, labeled)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        tps, dt, _ = cpu_port_throughput(w, sel, threads, labeled)
        t_tot += dt
        n_tot += len(sel)
    value = n_tot / t_tot
    # the other variant once, as a sibling
    tps_other, dt_other, _ = cpu_port_throughput(w, sel, threads, (not labeled) and args.workload != "C4")
    tps_1, _, _ = cpu_port_throughput(w, spread(n_tiles, max(4, len(sel) // threads)), 1, labeled)
    sib = {"value": tps_other, "unit": "tiles/s", "sample": f"{len(sel)} tiles, one pass"}
    line = {
        "impl": "reference",
        "metric": metric_name(w),
        "value": value,
        "unit": "tiles/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_tot / args.steps,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args, w, 1),
        "path": "draw_to_pixels with the label pass" if labeled else "area passes (Fill, Casing, Stroke) of draw_to_pixels, as timed by `e2e` of the CUDA arm",
        "cpu_baseline": {"value": value, "unit": "tiles/s", "cores": threads, "kind": "port",
                         "sample": f"{len(sel)} tiles spread over the batch per step, {args.steps} steps", "value_1_thread": tps_1},
        "e2e": {"value": value, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        ("e2e_area_only" if labeled else "e2e_labeled"): sib,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--style", default="mapnik", choices=sorted(STYLES))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--min-seconds", type=float, default=2.0, help="length of the sustained block of the resident leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="tiles in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--lib", default=None, help="experiments only: load this build of libosmr_b200.so (e.g. other -D flags)")
    ap.add_argument("--e2e-direct", type=int, default=0, choices=[0, 1])
    ap.add_argument("--e2e-chunks", type=int, default=0, help="experiments: draw chunks of the staged e2e call (0 = library default)")
    ap.add_argument("--two-streams", type=int, default=1, choices=[0, 1])
    ap.add_argument("--resident-chunks", type=int, default=0)
    ap.add_argument("--skip-auto", action="store_true", help="skip the f3 / f4 legs")
    ap.add_argument("--skip-labeled", action="store_true", help="skip the labelled leg")
    ap.add_argument("--no-affinity", action="store_true", help="do not pin the rank to the CPUs next to its GPU")
    ap.add_argument("--debug", action="append", default=[], help="experiments: key=value for osmr_debug_set")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch

    dist = None
    if world > 1:
        # NCCL prints its version banner to STDOUT at NCCL_DEBUG >= VERSION; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    affinity = {"applied": False, "skipped": True} if args.no_affinity else pin_to_gpu_cpus(local_rank)

    from osm_renderer_b200 import _lib, sharding
    from osm_renderer_b200.drawer import GpuContext

    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)

    sampler = ClockSampler(local_rank)
    sampler.start()  # before the warm-up: nvidia-smi needs a moment to start printing

    labeled = not args.skip_labeled and args.workload != "C4"
    w = build_workload(args.workload, args.style, labels=labeled)
    spec = WORKLOADS[args.workload]
    scale = spec["scale"]
    D = 256 * scale
    n_batch = len(w["tiles"])
    if args.scaling == "weak":
        # the global request list is `world` interleaved copies of the batch; rank r renders the requests i with
        # i % world == r, i.e. exactly one full batch per GPU over the replicated dataset
        req = sharding.weak_scaling_request_list(n_batch, world)
        mine = req[sharding.shard_indices(len(req), rank, world)]
        assert (mine == np.arange(n_batch)).all()
        passes = [np.arange(n_batch)]
    else:
        # strong scaling: ONE list of 8 x batch requests (the batch repeated), request i -> rank i mod N; a rank draws its
        # requests in calls of at most one batch
        req = np.tile(np.arange(n_batch, dtype=np.int64), 8)
        mine = req[sharding.shard_indices(len(req), rank, world)]
        passes = [mine[i: i + n_batch] for i in range(0, len(mine), n_batch)]
    ctx = GpuContext(local_rank)
    ctx.set_geodata(w["bin"])
    ctx.set_table(w["table"])
    for kv in args.debug:
        k, v = kv.split("=")
        ctx.debug_set(k, int(v))
    L = ctx.L
    flags, canvas = ctx._flags(w["canvas"], w["caps"], False)
    tiles_per_step = int(sum(len(p) for p in passes))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # per-pass batches; C4 is drawn zoom group by zoom group (per-zoom tiles/s)
    def groups_of(sel):
        if args.workload != "C4":
            return [(spec["zoom"], sel)]
        out = []
        for z, first, cnt in w["groups"]:
            s = sel[(sel >= first) & (sel < first + cnt)]
            if len(s):
                out.append((z, s))
        return out

    calls = []  # (zoom, tiles, begins, areas) of every resident call of a step
    for sel in passes:
        for z, s in groups_of(sel):
            t, b, a = sub_batch(w, s)
            calls.append((z, t, b, a, s))

    # ---- value: batch description resident in HBM, output stays in HBM ----
    if args.resident_chunks:
        ctx.debug_set("resident_chunks", args.resident_chunks)
    single = len(calls) == 1
    headline_labeled = HEADLINE_LABELED and labeled
    if labeled:
        ctx.set_font(w["font"])
        ctx.set_label_table(w["ltable"])
        call_labels = [sub_batch(w, c[4], True)[3:] for c in calls]

    def upload(i, with_labels):
        z, t, b, a, s = calls[i]
        if with_labels:
            ctx.batch_upload_labeled(t, b, a, call_labels[i][0], call_labels[i][1])
        else:
            ctx.batch_upload(t, b, a)

    if single:
        upload(0, headline_labeled)

    def resident_step(acc=None, with_labels=None):
        with_labels = headline_labeled if with_labels is None else with_labels
        ev = 0.0
        for i, (z, t, b, a, s) in enumerate(calls):
            if not single:
                upload(i, with_labels)  # (several calls per step: the description is re-sent; C4 / strong scaling only)
            ms = ctx.batch_draw_labeled(w["canvas"], w["caps"]) if with_labels else ctx.batch_draw(w["canvas"], w["caps"])
            ev += ms
            if acc is not None:
                st = ctx.stats()
                acc.setdefault("zoom_ms", {}).setdefault(z, []).append(ms)
                acc.setdefault("zoom_tiles", {})[z] = acc.setdefault("zoom_tiles", {}).get(z, 0) + len(t)
                for k in ("ms_raster", "ms_plan", "ms_cover", "ms_label_device", "ms_label_cover"):
                    acc[k] = acc.get(k, 0.0) + st[k]
                for k in ("n_labels_active", "n_label_segments", "n_label_cells"):
                    acc.setdefault("lstats", {})[k] = acc.setdefault("lstats", {}).get(k, 0) + int(st[k])
                acc["launches"] = acc.get("launches", 0) + st["kernel_launches"]
                for k in ("n_tiles", "n_areas", "n_visible_ops", "n_node_refs", "geom_bytes", "mask_bytes", "walk_bytes", "walk_steps"):
                    acc.setdefault("stats", {})[k] = acc.setdefault("stats", {}).get(k, 0) + int(st[k])
        return ev

    for _ in range(max(args.warmup, 3)):
        resident_step()
    barrier()
    sampler.mark()
    t0 = time.perf_counter()
    acc = {}
    ev_ms = []
    for _ in range(args.steps):
        ev_ms.append(resident_step(acc))
    barrier()
    wall = time.perf_counter() - t0
    sampler.unmark()
    dev_s = sum(ev_ms) / 1000.0
    stats = {k: v // args.steps for k, v in acc["stats"].items()}  # per step
    raster_ms = acc["ms_raster"] / args.steps
    plan_ms = acc["ms_plan"] / args.steps
    cover_ms = acc["ms_cover"] / args.steps
    label_ms = acc["ms_label_device"] / args.steps
    label_cover_ms = acc["ms_label_cover"] / args.steps
    lstats = {k: v // args.steps for k, v in acc.get("lstats", {}).items()}
    launches = acc["launches"]

    # ---- the label kernels with nothing beside them (debug key label_serial): in the headline step they share the SMs with the
    # area kernels, so their event times there are not their own durations ----
    label_alone = None
    if headline_labeled:
        ctx.debug_set("label_serial", 1)
        try:
            for _ in range(2):
                resident_step()
            acc_s = {}
            n_serial = max(3, min(args.steps, 10))
            for _ in range(n_serial):
                resident_step(acc_s)
            label_alone = {"ms_label_pass": acc_s["ms_label_device"] / n_serial, "ms_label_cover": acc_s["ms_label_cover"] / n_serial,
                           "note": "label pass finished before the area passes start; not part of any timed leg"}
        finally:
            ctx.debug_set("label_serial", 0)

    # ---- value_area_only: the Fill / Casing / Stroke passes alone (round 1's headline), same protocol ----
    area_only = None
    if headline_labeled:
        if single:
            upload(0, False)
        for _ in range(3):
            resident_step(None, False)
        barrier()
        t0a = time.perf_counter()
        acc_a = {}
        for _ in range(args.steps):
            resident_step(acc_a, False)
        barrier()
        area_wall = time.perf_counter() - t0a
        # (stage times of the area kernels WITHOUT the label stream beside them: what their roofline is quoted on)
        area_only = {"wall": area_wall, "ms_cover": acc_a["ms_cover"] / args.steps, "ms_raster": acc_a["ms_raster"] / args.steps,
                     "ms_plan": acc_a["ms_plan"] / args.steps}
        if single:
            upload(0, True)

    # ---- sustained: the same step repeated for >= min_seconds (power / thermal behaviour, median and min per step) ----
    sustained = None
    if args.min_seconds > 0:
        per = []
        barrier()
        sampler.mark()
        ts = time.perf_counter()
        while True:
            t1 = time.perf_counter()
            resident_step()
            torch.cuda.synchronize()
            per.append(time.perf_counter() - t1)
            if time.perf_counter() - ts >= args.min_seconds and len(per) >= 5:
                break
        barrier()
        sus_wall = time.perf_counter() - ts
        sampler.unmark()
        sustained = {"seconds": sus_wall, "steps": len(per), "tiles_per_s": tiles_per_step * len(per) / sus_wall,
                     "ms_per_step_median": 1000.0 * float(np.median(per)), "ms_per_step_min": 1000.0 * float(np.min(per))}
    clocks = sampler.stop()

    # ---- e2e: pinned host buffers through osmr_draw_tiles (H2D inputs + D2H RGB inside the timed region) ----
    # (page-locked buffers are allocated after the affinity was set: first touch puts them next to this rank's GPU)
    max_call = max(len(c[1]) for c in calls)
    out_bytes_call = max_call * D * D * 3
    pin_out = L.osmr_alloc_pinned(out_bytes_call)

    def pinned_copy(a):
        a = np.ascontiguousarray(a)
        p = L.osmr_alloc_pinned(max(a.nbytes, 16))
        C.memmove(p, a.ctypes.data, a.nbytes)
        return p, a.nbytes

    e2e_calls = []
    h2d = 0
    for z, t, b, a, s in calls:
        pt, nt = pinned_copy(t)
        pb, nb = pinned_copy(b)
        pa, na = pinned_copy(a)
        e2e_calls.append((z, len(t), pt, pb, pa, nt))
        h2d += nt + nb + na
    d2h = tiles_per_step * D * D * 3
    ctx.debug_set("direct_out", args.e2e_direct)
    ctx.debug_set("two_streams", args.two_streams)
    if args.e2e_chunks:
        ctx.debug_set("host_chunks", args.e2e_chunks)

    def check(rc):
        if rc != 0:
            raise RuntimeError(L.osmr_last_error(ctx.h))

    def e2e_step():
        for z, n, pt, pb, pa, _ in e2e_calls:
            check(L.osmr_draw_tiles(ctx.h, pt, n, pb, pa, canvas.ctypes.data, flags, pin_out))

    def timed(step_fn, warm=2):
        for _ in range(max(1, min(args.warmup, warm))):
            step_fn()
        barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            step_fn()
        barrier()
        return time.perf_counter() - t

    e2e_wall = timed(e2e_step)
    checksum = int(np.frombuffer((C.c_uint8 * 4096).from_address(pin_out), dtype=np.uint8).sum())
    # the GPU tiles of the last call (pin_out) against the CPU render later on
    last_sel = calls[-1][4]
    gpu_last = np.frombuffer((C.c_uint8 * (len(last_sel) * D * D * 3)).from_address(pin_out), dtype=np.uint8).reshape(len(last_sel), D, D, 3).copy()

    # ---- latency: the reference's own call shape (n_tiles = 1) and small batches, p50 of 30 calls ----
    latency = {}
    z0, n0, pt0, pb0, pa0, _ = e2e_calls[0]
    t_first, b_first, a_first = calls[0][1], calls[0][2], calls[0][3]
    for nb_ in (1, 8, 64):
        if nb_ > n0:
            continue
        bb = np.ascontiguousarray(b_first[: nb_ + 1])
        ts_ = []
        for i in range(33):
            t1 = time.perf_counter()
            check(L.osmr_draw_tiles(ctx.h, pt0, nb_, bb.ctypes.data, pa0, canvas.ctypes.data, flags, pin_out))
            ts_.append(time.perf_counter() - t1)
        latency[f"n_tiles_{nb_}_p50"] = 1000.0 * float(np.median(ts_[3:]))

    # ---- context for e2e: what the bus of this box gives a plain pinned copy of the step's output / input ----
    pcie = None
    try:
        dev_buf = torch.empty(out_bytes_call, dtype=torch.uint8, device="cuda")
        host_buf = torch.empty(out_bytes_call, dtype=torch.uint8, pin_memory=True)
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        host_buf.copy_(dev_buf, non_blocking=True)
        torch.cuda.synchronize()
        barrier()  # all ranks probe at the same time: the figure is what a rank gets while its neighbours copy too
        ev0.record()
        host_buf.copy_(dev_buf, non_blocking=True)
        ev1.record()
        nh = min(h2d, out_bytes_call)
        dev_buf[:nh].copy_(host_buf[:nh], non_blocking=True)
        ev2.record()
        torch.cuda.synchronize()
        pcie = {"d2h_gbs": out_bytes_call / ev0.elapsed_time(ev1) / 1e6, "h2d_gbs": nh / ev1.elapsed_time(ev2) / 1e6,
                "d2h_ms_for_call_output": ev0.elapsed_time(ev1), "concurrent_ranks": world}
        del dev_buf, host_buf
    except Exception as exc:  # the probe is context, never a reason to fail the bench
        pcie = {"error": str(exc)}

    # ---- e2e_labeled: the whole Drawer::draw_to_pixels (area passes + label pass) through osmr_draw_tiles_labeled ----
    lab = None
    lab_wall = None
    if labeled:
        lab_calls = []
        lab_h2d = 0
        for (z, t, b, a, s), (_, n, pt, pb, pa, nt) in zip(calls, e2e_calls):
            lb_, ll_ = sub_batch(w, s, True)[3:]
            plb, nlb = pinned_copy(lb_)
            pll, nll = pinned_copy(ll_)
            lab_calls.append((n, pt, pb, pa, plb, pll))
            lab_h2d += nlb + nll
        pin_lab = L.osmr_alloc_pinned(out_bytes_call)
        lab_ms = {"layout": [], "device": []}

        def lab_step():
            for n, pt, pb, pa, plb, pll in lab_calls:
                check(L.osmr_draw_tiles_labeled(ctx.h, pt, n, pb, pa, plb, pll, canvas.ctypes.data, flags, pin_lab))
                st_ = ctx.stats()
                lab_ms["layout"].append(st_["ms_label_layout"])
                lab_ms["device"].append(st_["ms_label_device"])

        lab_wall = timed(lab_step)
        gpu_lab_last = np.frombuffer((C.c_uint8 * (len(last_sel) * D * D * 3)).from_address(pin_lab), dtype=np.uint8).reshape(len(last_sel), D, D, 3).copy()
        n_lab_calls = len(lab_calls)
        lab = {"unit": "tiles/s", "h2d_bytes_per_step": h2d + lab_h2d, "d2h_bytes_per_step": d2h,
               "api": "osmr_draw_tiles_labeled (area passes + label pass = the reference's whole draw_to_pixels, drawer.rs:60-131)",
               "ms_label_layout": float(np.mean(lab_ms["layout"][-args.steps * n_lab_calls:])) * n_lab_calls,
               "ms_label_device": float(np.mean(lab_ms["device"][-args.steps * n_lab_calls:])) * n_lab_calls,
               "label_generations_per_step": int(sum(len(sub_batch(w, c[4], True)[4]) for c in calls)),
               "pixels_changed_by_labels_in_last_call": int((gpu_lab_last != gpu_last).any(axis=-1).sum())}

    # ---- e2e_auto / e2e_png / e2e_auto_png (f3, f4): tile list in and / or PNG files out ----
    auto = png = auto_png = auto_lab = auto_lab_png = None
    auto_wall = png_wall = auto_png_wall = auto_lab_wall = auto_lab_png_wall = None
    if not args.skip_auto:
        from osm_renderer_b200.upstream import pipeline

        t_cls = time.perf_counter()
        for z in sorted({c[0] for c in calls}):
            wc, mc, cb, cs = pipeline.zoom_class_tables(w["builder"], z)
            ctx.set_table(w["table"])
            ctx.set_zoom_styles(z, wc, mc, cb, cs)
        t_cls = time.perf_counter() - t_cls
        pin_auto = L.osmr_alloc_pinned(out_bytes_call)

        def auto_step():
            for z, n, pt, pb, pa, _ in e2e_calls:
                check(L.osmr_draw_tiles_auto(ctx.h, pt, n, canvas.ctypes.data, flags, pin_auto))

        auto_wall = timed(auto_step)
        ast = ctx.stats()
        same = bool((np.frombuffer((C.c_uint8 * gpu_last.size).from_address(pin_auto), dtype=np.uint8) == gpu_last.reshape(-1)).all())
        auto = {"unit": "tiles/s", "h2d_bytes_per_step": int(sum(c[5] for c in e2e_calls)), "d2h_bytes_per_step": d2h,
                "api": "osmr_draw_tiles_auto (tile list only; styled-area lists built on the device)",
                "ms_auto_stage_last_call": float(ast["ms_auto"]), "styled_areas_after_culling_last_call": int(ast["n_areas"]),
                "identical_to_e2e_output": same, "class_tables_host_s": t_cls}

        cap = max_call * int(L.osmr_png_bound(scale))
        pin_png = L.osmr_alloc_pinned(cap)
        offs = np.zeros(max_call + 1, dtype=np.uint64)
        png_bytes = [0]

        def png_step():
            png_bytes[0] = 0
            for z, n, pt, pb, pa, _ in e2e_calls:
                check(L.osmr_draw_tiles_png(ctx.h, pt, n, pb, pa, canvas.ctypes.data, flags, pin_png, cap, offs.ctypes.data))
                png_bytes[0] += int(offs[n])

        png_wall = timed(png_step)
        pst = ctx.stats()
        png = {"unit": "tiles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": png_bytes[0], "mean_png_bytes": png_bytes[0] / tiles_per_step,
               "ms_png_stage_last_call": float(pst["ms_png"]), "api": "osmr_draw_tiles_png (filter + deflate + checksums on the device)"}

        def auto_png_step():
            png_bytes[0] = 0
            for z, n, pt, pb, pa, _ in e2e_calls:
                check(L.osmr_draw_tiles_auto_png(ctx.h, pt, n, canvas.ctypes.data, flags, pin_png, cap, offs.ctypes.data))
                png_bytes[0] += int(offs[n])

        auto_png_wall = timed(auto_png_step)
        auto_png = {"unit": "tiles/s", "h2d_bytes_per_step": int(sum(c[5] for c in e2e_calls)), "d2h_bytes_per_step": png_bytes[0],
                    "api": "osmr_draw_tiles_auto_png (tile list in, PNG files out = Drawer::draw_tile behind the server's lookup)"}

        # ---- e2e_auto_labeled{,_png}: the whole draw_to_pixels / draw_tile from a tile list (f3 for the label lists as well) ----
        if labeled:
            t_lcls = time.perf_counter()
            for z in sorted({c[0] for c in calls}):
                nc_, wc_, mc_, cb_, cs_ = pipeline.zoom_label_class_tables(w["builder"], z, w["ltable"])
                ctx.set_label_table(w["ltable"])
                ctx.set_zoom_label_styles(z, nc_, wc_, mc_, cb_, cs_)
            t_lcls = time.perf_counter() - t_lcls

            def auto_lab_step():
                for z, n, pt, pb, pa, _ in e2e_calls:
                    check(L.osmr_draw_tiles_auto_labeled(ctx.h, pt, n, canvas.ctypes.data, flags, pin_auto))

            auto_lab_wall = timed(auto_lab_step)
            alst = ctx.stats()
            same = bool((np.frombuffer((C.c_uint8 * gpu_lab_last.size).from_address(pin_auto), dtype=np.uint8) == gpu_lab_last.reshape(-1)).all())
            auto_lab = {"unit": "tiles/s", "h2d_bytes_per_step": int(sum(c[5] for c in e2e_calls)), "d2h_bytes_per_step": d2h,
                        "api": "osmr_draw_tiles_auto_labeled (tile list only; styled-area AND label lists built on the device, then the whole draw_to_pixels)",
                        "ms_auto_stage_last_call": float(alst["ms_auto"]), "live_label_generations_last_call": int(alst["n_labels_active"]),
                        "label_path_last_call": int(alst["label_path"]), "identical_to_e2e_labeled_output": same, "label_class_tables_host_s": t_lcls}

            def auto_lab_png_step():
                png_bytes[0] = 0
                for z, n, pt, pb, pa, _ in e2e_calls:
                    check(L.osmr_draw_tiles_auto_labeled_png(ctx.h, pt, n, canvas.ctypes.data, flags, pin_png, cap, offs.ctypes.data))
                    png_bytes[0] += int(offs[n])

            auto_lab_png_wall = timed(auto_lab_png_step)
            auto_lab_png = {"unit": "tiles/s", "h2d_bytes_per_step": int(sum(c[5] for c in e2e_calls)), "d2h_bytes_per_step": png_bytes[0],
                            "mean_png_bytes": png_bytes[0] / tiles_per_step,
                            "api": "osmr_draw_tiles_auto_labeled_png (tile list in, PNG files out, label pass included = Drawer::draw_tile complete)"}

    # ---- max over ranks (time), sum over ranks (tiles): the only collectives of the whole job ----
    job_tiles = tiles_per_step * args.steps
    red = lambda secs: sharding.reduce_job(dist, secs, job_tiles, device="cuda")
    dev_s_max, total_tiles = red(dev_s)
    wall, _ = red(wall)
    e2e_wall_max, _ = red(e2e_wall)
    per_rank_d2h_gbs = d2h * args.steps / e2e_wall / 1e9  # this rank's own rate (rank 0 prints its own)
    for leg, lw in ((lab, lab_wall), (auto, auto_wall), (png, png_wall), (auto_png, auto_png_wall), (auto_lab, auto_lab_wall), (auto_lab_png, auto_lab_png_wall)):
        if leg is not None:  # whole-job numbers for every leg: max wall over ranks, tiles of all ranks
            mw, mt = red(lw)
            leg["value"], leg["ms_per_step"] = mt / mw, 1000.0 * mw / args.steps
            leg["d2h_gbs_per_rank"] = leg["d2h_bytes_per_step"] * args.steps / lw / 1e9
    raster_mean_ms, _ = red(raster_ms)
    cover_mean_ms, _ = red(cover_ms)
    label_cover_mean_ms, _ = red(label_alone["ms_label_cover"] if label_alone else label_cover_ms)
    if area_only is not None:
        aw, at = red(area_only["wall"])
        area_only = {"value": at / aw, "unit": "tiles/s", "ms_per_step": 1000.0 * aw / args.steps,
                     "path": "Fill, Casing and Stroke passes only (osmr_batch_draw), resident: round 1's headline leg",
                     "ms_cover": area_only["ms_cover"], "ms_raster": area_only["ms_raster"], "ms_plan": area_only["ms_plan"],
                     "stage_ms_note": "this rank's CUDA-event stage times with no label stream beside the area kernels"}
    if sustained is not None:
        sw, stl = sharding.reduce_job(dist, sustained["seconds"], tiles_per_step * sustained["steps"], device="cuda")
        sustained["tiles_per_s"] = stl / sw

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = total_tiles / wall
    e2e_value = total_tiles / e2e_wall_max

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth of this pool's B200)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # B_tile = 4 D^2 + 16 U + 4 R + 64 G (+ 8 per dash number: a few hundred bytes per batch, omitted), summed
    # over the step's tiles.  U = distinct nodes per tile, R = node references of the styled areas, G = visible generations.
    U = distinct_nodes_per_batch(w) * (tiles_per_step / n_batch) if args.workload != "C4" else None
    R, G = stats["n_node_refs"], stats["n_visible_ops"]
    b_step = (4 * D * D * tiles_per_step + 16 * U + 4 * R + 64 * G) if U is not None else None
    # kernel-local bytes: what the dominant kernel itself reads and writes once per launch; the walk-cache
    # alphas are an INTERMEDIATE the design created (written by line_cover_kernel, read by raster_kernel), not algorithmic input
    walk_alpha_bytes = 8 * stats["walk_steps"]
    raster_local = tiles_per_step * D * D * 3 + 32 * G + stats["geom_bytes"] + stats["mask_bytes"] + walk_alpha_bytes
    cover_local = stats["geom_bytes"] + walk_alpha_bytes + stats["walk_bytes"] // (8 * 4)
    # (in a labelled step the area kernels share the SMs with the label stream: their own durations come from the area-only leg)
    a_cover_ms = area_only["ms_cover"] if area_only else cover_mean_ms
    a_raster_ms = area_only["ms_raster"] if area_only else raster_mean_ms
    if a_cover_ms > a_raster_ms:
        dom, dom_ms, local_bytes = "line_cover_kernel", a_cover_ms, cover_local
    else:
        dom, dom_ms, local_bytes = "raster_kernel", a_raster_ms, raster_local
    algo_bytes = b_step if b_step is not None else local_bytes - walk_alpha_bytes
    area_dom = {"kernel": dom, "kernel_ms": dom_ms, "kernel_local_bytes": int(local_bytes), "algorithmic_bytes_B_tile": None if b_step is None else int(b_step),
                "frac_B_tile": None if b_step is None else b_step / (dom_ms / 1000.0) / 1e9 / peak, "kernel_local_frac": local_bytes / (dom_ms / 1000.0) / 1e9 / peak}
    if headline_labeled and label_cover_mean_ms > dom_ms:
        # the longest kernel of a labelled step is the glyph coverage: it reads every outline segment once (32 B, the reference's
        # draw_line call stream) and writes / sweeps the coverage cells (2 x 8 B); SURVEY 8(d)'s B_tile has no label term, so the
        # label inputs are added to it: 8 B per label generation listed + 32 B per segment + 16 B per coverage cell
        dom, dom_ms = "label_cover_kernel", label_cover_mean_ms
        local_bytes = 32 * lstats.get("n_label_segments", 0) + 16 * lstats.get("n_label_cells", 0)
        algo_bytes = local_bytes
        walk_alpha_bytes = 0
    achieved = algo_bytes / (dom_ms / 1000.0) / 1e9
    traffic = None
    prof_json = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof_json):
        try:
            pj = json.load(open(prof_json))
            if pj.get("workload") == args.workload and pj.get("style", "mapnik") == args.style:
                traffic = pj.get(dom + "_dram_bytes_per_launch")
        except Exception:
            pass
    step_ms = 1000.0 * wall / args.steps
    roofline = {
        "bound": "hbm",
        "kernel": dom,
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "traffic_source": "ncu --set full capture of the same workload (profiles/roofline_traffic.json), per launch" if traffic else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": int(algo_bytes),
        "B_tile_terms": {"out_4DD": 4 * D * D * tiles_per_step, "sum_U": U, "sum_R": R, "sum_G": G,
                         "formula": "4*D^2*tiles + 16*U + 4*R + 64*G (SURVEY.md 8d), one launch covers the whole step"},
        "kernel_local_bytes": int(local_bytes),
        "walk_cache_intermediate_bytes": int(walk_alpha_bytes),
        "kernel_local_frac": local_bytes / (dom_ms / 1000.0) / 1e9 / peak,
        "kernel_ms": dom_ms,
        "kernel_share_of_step": dom_ms / (1000.0 * dev_s_max / args.steps),
        "whole_step_frac": (b_step / (step_ms / 1000.0) / 1e9 / peak) if b_step is not None else None,
        # the write-only variant north_star quotes: whole-path tiles/s x 4*D^2 output bytes against the HBM peak
        "north_star_write_frac": (value / world) * (4.0 * D * D) / (peak * 1e9),
        "area_passes_dominant_kernel": area_dom,
        "note": "the path is FP64/integer-issue and latency bound, not HBM bound (SURVEY.md F6); see profiles/",
    }

    # ---- CPU baseline beside it (bounded sample, rank 0 only) ----
    cpu = None
    max_abs_diff = None
    max_abs_diff_lab = None
    if not args.skip_cpu_baseline and world == 1:  # the CPU baseline belongs to the N=1 line only
        threads = host_threads()
        n_last = len(last_sel)
        probe_tps, _, _ = cpu_port_throughput(w, last_sel[spread(n_last, threads)], threads)
        n_sample = args.cpu_sample or int(min(n_last, max(threads, probe_tps * 12.0)))
        pick = spread(n_last, n_sample)
        tps, dt, imgs = cpu_port_throughput(w, last_sel[pick], threads)
        pick1 = spread(n_last, max(2, int(n_sample / threads / 2)))
        tps1, dt1, _ = cpu_port_throughput(w, last_sel[pick1], 1)
        cpu = {"value": tps, "unit": "tiles/s", "cores": threads, "kind": "port",
               "sample": f"{len(pick)} tiles spread over the same batch, {dt:.1f}s, C++ restatement of the reference CPU path, tiles dealt to the threads round-robin",
               "value_1_thread": tps1, "sample_1_thread": f"{len(pick1)} tiles, {dt1:.1f}s"}
        # second half of BASELINE.json's metric: max |dRGB| of the GPU tiles (e2e output) against the CPU render
        max_abs_diff = int(max(np.abs(gpu_last[i].astype(np.int16) - im.astype(np.int16)).max() for i, im in zip(pick, imgs)))
        if lab is not None:
            pickl = spread(n_last, max(threads, n_sample // 2))
            tpsl, dtl, imgsl = cpu_port_throughput(w, last_sel[pickl], threads, labeled=True)
            max_abs_diff_lab = int(max(np.abs(gpu_lab_last[i].astype(np.int16) - im.astype(np.int16)).max() for i, im in zip(pickl, imgsl)))
            if headline_labeled:  # the headline path is the whole draw_to_pixels: its CPU figure is `value`
                tps1l, dt1l, _ = cpu_port_throughput(w, last_sel[pick1], 1, labeled=True)
                cpu = {"value": tpsl, "unit": "tiles/s", "cores": threads, "kind": "port",
                       "sample": f"{len(pickl)} tiles spread over the same batch, {dtl:.1f}s, C++ restatement of the reference CPU path (draw_to_pixels with the label pass), tiles dealt to the threads round-robin",
                       "value_1_thread": tps1l, "sample_1_thread": f"{len(pick1)} tiles, {dt1l:.1f}s",
                       "value_area_only": tps, "value_area_only_1_thread": tps1}
            else:
                cpu["value_labeled"] = tpsl
                cpu["sample_labeled"] = f"{len(pickl)} tiles, {dtl:.1f}s, draw_to_pixels with the label pass"

    per_zoom = None
    if args.workload == "C4":
        per_zoom = {}
        for z, first, cnt in w["groups"]:
            ms = acc["zoom_ms"].get(z)
            if ms:  # rank 0's own calls (device time, CUDA events)
                per_zoom[str(z)] = {"tiles_per_step_on_this_rank": acc["zoom_tiles"][z] // args.steps,
                                    "tiles_per_s_per_gpu": acc["zoom_tiles"][z] / (sum(ms) / 1000.0), "ms_per_call": float(np.mean(ms))}

    e2e_line = {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1000.0 * e2e_wall_max / args.steps, "d2h_gbs_per_rank": per_rank_d2h_gbs,
                "api": "osmr_draw_tiles (pinned host buffers; " + ("tiles stored straight into the host buffer by raster_kernel" if args.e2e_direct
                                                                   else "tiles staged in HBM, chunked D2H on a copy stream") + ")",
                "checksum": checksum}
    line = {
        "metric": metric_name(w),
        "value": value,
        "unit": "tiles/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args, w, world),
        "device_only": {"value": total_tiles / dev_s_max, "ms_per_step": 1000.0 * dev_s_max / args.steps,
                        "note": "sum of the per-call CUDA-event times; `value` is wall clock between barriers"},
        "sustained": sustained,
        "stage_ms": {"plan+geometry+fill_rows+bin": plan_ms, "line_cover": cover_ms, "raster": raster_ms,
                     "label_pass (own stream, beside the area stages)": label_ms, "label_cover (inside label_pass)": label_cover_ms},
        "label_kernels_alone": label_alone,
        "value_area_only": area_only,
        "label_stats": lstats,
        # the reference's perf-stats stage names (drawer.rs:51-123) where a stage is separable on the device
        "perf_stats": {"Style areas": (auto or {}).get("ms_auto_stage_last_call"), "Fill areas + Draw areas": plan_ms + cover_ms + raster_ms,
                       "Resetting TilePixels / Blend after areas / RGB export": "fused into raster_kernel",
                       "Draw labels": None if lab is None else lab["ms_label_layout"] + lab["ms_label_device"],
                       "Blend after labels": "fused into raster_kernel's export", "RGB triples to PNG": (png or {}).get("ms_png_stage_last_call")},
        "batch_stats": stats,
        "per_zoom": per_zoom,
        "clocks": clocks,
        "affinity": affinity,
        "e2e": e2e_line,
        "e2e_labeled": lab,
        "latency_ms": latency,
        "pcie_probe": pcie,
        "e2e_auto": auto,
        "e2e_png": png,
        "e2e_auto_png": auto_png,
        "e2e_auto_labeled": auto_lab,
        "e2e_auto_labeled_png": auto_lab_png,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "max_abs_diff_rgb_vs_cpu": max_abs_diff,
        "max_abs_diff_rgb_vs_cpu_labeled": max_abs_diff_lab,
    }
    if headline_labeled and lab is not None:
        line["e2e_area_only"] = e2e_line
        line["e2e"] = lab
        del line["e2e_labeled"]
        line["path"] = "the whole Drawer::draw_to_pixels (drawer.rs:60-131): Fill, Casing, Stroke passes + label pass"
    else:
        line["path"] = "Fill, Casing and Stroke passes of Drawer::draw_to_pixels (no label lists in this workload)"
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

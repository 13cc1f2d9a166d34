#!/usr/bin/env python
"""bench.py -- tiles/s of the per-tile draw path (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

Workload (config.workload "C2"): 1024 tiles 256x256 at z=14 (x 9888..9919, y 5104..5135) over the seeded synthetic
metro geodata image (osm_renderer_b200/upstream/synth.py), styled with the reference's mapnik.mapcss test
stylesheet (JOSM flavour).  One "step" = one pass of the draw path over that batch on every rank (weak scaling:
each rank renders its own full batch; tiles are independent, there is no data-path collective).

JSON line: `value` = tiles/s with the batch description resident in HBM and the output left in HBM;
`e2e` = tiles/s through osmr_draw_tiles with pinned HOST buffers (H2D of the styled-area lists + D2H of the RGB
tiles inside the timed region); `roofline` = raster kernel (dominant) against the measured HBM peak;
`cpu_baseline` = the oracle (C++ restatement of the reference CPU path) on this box's host cores.
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
    # name: (zoom, x0, y0, n, scale)
    "C2": (14, 9888, 5104, 32, 1),
    "C3": (14, 9888, 5104, 32, 2),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------
def build_workload(name: str, cache_dir: str | None = None):
    """Returns dict(bin, table, tiles, area_begin, areas, canvas, caps)."""
    from osm_renderer_b200.upstream import geodata, mapcss, pipeline, styler as st, synth
    from osm_renderer_b200.wire import TILE_DTYPE, StyleTable

    zoom, x0, y0, n, scale = WORKLOADS[name]
    t0 = time.time()
    data = synth.make_metro()
    rd = geodata.GeodataReader(data)
    rules = mapcss.load_rules_json(os.path.join(ROOT, "tests", "golden", "mapnik_rules.json.gz"))
    S = st.Styler(rules, "josm", None)
    table = StyleTable(None)  # fill-image icons are not shipped: such areas are skipped like a failed icon load
    fb = pipeline.FastBatchBuilder(rd, S, table)
    tiles = [(zoom, x, y, scale) for y in range(y0, y0 + n) for x in range(x0, x0 + n)]
    parts = [fb.areas_array(z, x, y) for (z, x, y, s) in tiles]
    begins = np.zeros(len(tiles) + 1, dtype=np.uint32)
    begins[1:] = np.cumsum([len(p) for p in parts])
    areas = np.concatenate(parts)
    log(f"[bench] workload {name}: {len(tiles)} tiles, {len(areas)} styled areas, {len(data) / 1e6:.0f} MB geodata, "
        f"{len(table.rows)} styles, built in {time.time() - t0:.1f}s")
    return {
        "bin": data,
        "reader": rd,
        "builder": fb,
        "table": table,
        "tiles": np.array(tiles, dtype=TILE_DTYPE),
        "area_begin": begins,
        "areas": areas,
        "canvas": S.canvas_fill_color,
        "caps": S.use_caps_for_dashes,
        "scale": scale,
    }


def distinct_nodes_per_batch(w) -> int:
    """U of SURVEY.md 8(d): distinct nodes referenced by the styled areas of each tile, summed over tiles."""
    rd = w["reader"]
    way_off = rd.ways["off"].astype(np.int64)
    way_len = rd.ways["len"].astype(np.int64)
    total = 0
    ab = w["area_begin"]
    ent = w["areas"]["entity"]
    for t in range(len(w["tiles"])):
        e = np.unique(ent[ab[t] : ab[t + 1]])
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
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

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
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ----------------------------------------------------------------------------------------------------------
def cpu_port_throughput(w, n_sample: int, threads: int, repeats: int = 1):
    """tiles/s of the oracle (C++ restatement of the reference CPU path) on `threads` host threads."""
    import oracle

    ab = w["area_begin"]
    n_tiles = len(w["tiles"])
    sel = np.linspace(0, n_tiles - 1, n_sample).astype(int)  # spread over the batch
    parts = [w["areas"][ab[i] : ab[i + 1]] for i in sel]
    begins = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    areas = np.concatenate(parts)
    tiles = w["tiles"][sel]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        imgs = oracle.draw_tiles(w["bin"], w["table"], tiles, begins, areas, w["canvas"], w["caps"], n_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    w["_last_cpu_sample"] = (sel, imgs)
    return n_sample / best, best


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="tiles in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--lib", default=None, help="experiments only: load this build of libosmr_b200.so (e.g. other -D flags)")
    ap.add_argument("--e2e-direct", type=int, default=0, choices=[0, 1],
                    help="e2e leg: 0 = tiles staged in HBM, chunk-wise D2H on its own stream while later chunks are drawn (default), "
                         "1 = raster_kernel stores the tiles straight into the page-locked host buffer")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="experiments: draw chunks of the staged e2e call (0 = library default)")
    ap.add_argument("--two-streams", type=int, default=1, choices=[0, 1], help="experiments: draw chunks of the e2e call on two streams")
    ap.add_argument("--resident-chunks", type=int, default=0, help="experiments: draw chunks of the resident (value) leg (0 = library default)")
    ap.add_argument("--skip-auto", action="store_true", help="skip the osmr_draw_tiles_auto (f3) and osmr_draw_tiles_png (f4) legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus

    zoom, x0, y0, n, scale = WORKLOADS[args.workload]
    D = 256 * scale
    cfg = {
        "workload": f"{args.workload}: {n * n} tiles {D}x{D} z={zoom} x{x0}..{x0 + n - 1} y{y0}..{y0 + n - 1}, synthetic metro .bin (seed 0xB20005A1), mapnik.mapcss (JOSM), area passes",
        "tiles_per_step_per_gpu": n * n,
        "sharding": "one full batch per rank (tiles are independent; no data-path collective)",
        "l2": "per-step working set (styled areas + plan scratch + RGB output > 400 MB) exceeds the 126 MB L2",
    }

    # ------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        # The reference is Rust and cannot be built in this image: its CPU algorithm is timed through the oracle
        # port (oracle/osmr_oracle.cpp), all host threads, one tile per thread like src/http_server.rs:50-83.
        if rank != 0:
            return
        w = build_workload(args.workload)
        threads = host_threads()
        sample = args.cpu_sample or min(n * n, max(16 * threads, 256))  # several tiles per thread: the threads stay busy
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_port_throughput(w, min(sample, threads), threads)
        t_tot, n_tot = 0.0, 0
        for _ in range(args.steps):
            tps, dt = cpu_port_throughput(w, sample, threads)
            t_tot += dt
            n_tot += sample
        value = n_tot / t_tot
        line = {
            "impl": "reference",
            "metric": f"{D}x{D} tiles/sec at z=14",
            "value": value,
            "unit": "tiles/s",
            "n_gpus": n_gpus,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": 1000.0 * t_tot / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "tiles/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} tiles of the batch per step, {args.steps} steps"},
            "e2e": {"value": value, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------------------------------------------
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

    from osm_renderer_b200 import _lib, sharding
    from osm_renderer_b200.drawer import GpuContext

    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)

    w = build_workload(args.workload)
    # weak scaling: the global request list is `world` interleaved copies of the batch; rank r renders the requests
    # i with i % world == r, i.e. exactly one full batch per GPU over the replicated dataset
    req = sharding.weak_scaling_request_list(len(w["tiles"]), world)
    mine = req[sharding.shard_indices(len(req), rank, world)]
    assert (mine == np.arange(len(w["tiles"]))).all()
    ctx = GpuContext(local_rank)
    ctx.set_geodata(w["bin"])
    ctx.set_table(w["table"])
    n_tiles = len(mine)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: batch resident in HBM, output stays in HBM ----
    if args.resident_chunks:
        ctx.debug_set("resident_chunks", args.resident_chunks)
    ctx.batch_upload(w["tiles"], w["area_begin"], w["areas"])
    for _ in range(args.warmup):
        ctx.batch_draw(w["canvas"], w["caps"])
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ev_ms, raster_ms, plan_ms, cover_ms = [], [], [], []
    launches = 0
    for _ in range(args.steps):
        ev_ms.append(ctx.batch_draw(w["canvas"], w["caps"]))
        st = ctx.stats()
        raster_ms.append(st["ms_raster"])
        plan_ms.append(st["ms_plan"])
        cover_ms.append(st["ms_cover"])
        launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()  # sampled every 50 ms, only while the timed resident steps run (idle gaps would drag the median down)
    stats = ctx.stats()
    dev_s = sum(ev_ms) / 1000.0

    # ---- e2e: pinned host buffers through osmr_draw_tiles (H2D inputs + D2H RGB inside the timed region) ----
    L = ctx.L
    out_bytes = n_tiles * D * D * 3
    pin_out = L.osmr_alloc_pinned(out_bytes)
    in_arrays = [np.ascontiguousarray(w["tiles"]), np.ascontiguousarray(w["area_begin"]), np.ascontiguousarray(w["areas"])]
    pins = []
    for a in in_arrays:
        p = L.osmr_alloc_pinned(a.nbytes)
        C.memmove(p, a.ctypes.data, a.nbytes)
        pins.append(p)
    h2d = int(sum(a.nbytes for a in in_arrays))
    flags, canvas = ctx._flags(w["canvas"], w["caps"], False)
    ctx.debug_set("direct_out", args.e2e_direct)
    ctx.debug_set("two_streams", args.two_streams)
    if args.e2e_chunks:
        ctx.debug_set("host_chunks", args.e2e_chunks)

    def e2e_step():
        rc = L.osmr_draw_tiles(ctx.h, pins[0], n_tiles, pins[1], pins[2], canvas.ctypes.data, flags, pin_out)
        if rc != 0:
            raise RuntimeError(L.osmr_last_error(ctx.h))

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_wall = time.perf_counter() - t1
    checksum = int(np.frombuffer((C.c_uint8 * 4096).from_address(pin_out), dtype=np.uint8).sum())

    # ---- context for e2e: what the bus of this box gives a plain pinned copy of the step's output / input ----
    pcie = None
    try:
        dev_buf = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
        host_buf = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
        ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        host_buf.copy_(dev_buf, non_blocking=True)
        torch.cuda.synchronize()
        ev0.record()
        host_buf.copy_(dev_buf, non_blocking=True)
        ev1.record()
        dev_buf[:h2d].copy_(host_buf[:h2d], non_blocking=True)
        ev2.record()
        torch.cuda.synchronize()
        pcie = {"d2h_gbs": out_bytes / ev0.elapsed_time(ev1) / 1e6, "h2d_gbs": h2d / ev1.elapsed_time(ev2) / 1e6,
                "d2h_ms_for_step_output": ev0.elapsed_time(ev1)}
        del dev_buf, host_buf
    except Exception as exc:  # the probe is context, never a reason to fail the bench
        pcie = {"error": str(exc)}

    # ---- e2e_auto (f3): only the tile list crosses the bus; candidate lookup + painter's order on the device ----
    auto = None
    if not args.skip_auto:
        from osm_renderer_b200.upstream import pipeline

        t_cls = time.perf_counter()
        wc, mc, cb, cs = pipeline.zoom_class_tables(w["builder"], zoom)
        ctx.set_table(w["table"])
        ctx.set_zoom_styles(zoom, wc, mc, cb, cs)
        t_cls = time.perf_counter() - t_cls
        pin_auto = L.osmr_alloc_pinned(out_bytes)

        def auto_step():
            rc = L.osmr_draw_tiles_auto(ctx.h, pins[0], n_tiles, canvas.ctypes.data, flags, pin_auto)
            if rc != 0:
                raise RuntimeError(L.osmr_last_error(ctx.h))

        for _ in range(max(1, min(args.warmup, 2))):
            auto_step()
        barrier()
        t2 = time.perf_counter()
        for _ in range(args.steps):
            auto_step()
        barrier()
        auto_wall = time.perf_counter() - t2
        ast = ctx.stats()
        same = bool((np.frombuffer((C.c_uint8 * out_bytes).from_address(pin_auto), dtype=np.uint8)
                     == np.frombuffer((C.c_uint8 * out_bytes).from_address(pin_out), dtype=np.uint8)).all())
        auto = {"value": n_tiles * args.steps / auto_wall, "unit": "tiles/s", "ms_per_step": 1000.0 * auto_wall / args.steps,
                "h2d_bytes_per_step": int(in_arrays[0].nbytes), "d2h_bytes_per_step": out_bytes,
                "api": "osmr_draw_tiles_auto (tile list only; styled-area lists built on the device)",
                "ms_auto_stage": float(ast["ms_auto"]), "styled_areas_after_culling": int(ast["n_areas"]),
                "identical_to_e2e_output": same, "class_tables_host_s": t_cls}

    # ---- e2e_png (f4): PNG files instead of RGB triples; the files of the batch come back packed ----
    png = None
    if not args.skip_auto:
        cap = n_tiles * int(L.osmr_png_bound(scale))
        pin_png = L.osmr_alloc_pinned(cap)
        offs = np.zeros(n_tiles + 1, dtype=np.uint64)

        def png_step():
            rc = L.osmr_draw_tiles_png(ctx.h, pins[0], n_tiles, pins[1], pins[2], canvas.ctypes.data, flags, pin_png, cap, offs.ctypes.data)
            if rc != 0:
                raise RuntimeError(L.osmr_last_error(ctx.h))

        for _ in range(max(1, min(args.warmup, 2))):
            png_step()
        barrier()
        t3 = time.perf_counter()
        for _ in range(args.steps):
            png_step()
        barrier()
        png_wall = time.perf_counter() - t3
        pst = ctx.stats()
        png = {"value": n_tiles * args.steps / png_wall, "unit": "tiles/s", "ms_per_step": 1000.0 * png_wall / args.steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(offs[-1]), "mean_png_bytes": float(offs[-1]) / n_tiles,
               "ms_png_stage": float(pst["ms_png"]), "api": "osmr_draw_tiles_png (filter + deflate + checksums on the device)"}

    # ---- max over ranks (time), sum over ranks (tiles): the only collectives of the whole job ----
    job_tiles = n_tiles * args.steps
    dev_s, total_tiles = sharding.reduce_job(dist, dev_s, job_tiles, device="cuda")
    wall, _ = sharding.reduce_job(dist, wall, job_tiles, device="cuda")
    e2e_wall, _ = sharding.reduce_job(dist, e2e_wall, job_tiles, device="cuda")
    if auto is not None:  # whole-job numbers for the extra legs too: max wall over ranks, tiles of all ranks
        aw, at = sharding.reduce_job(dist, auto_wall, job_tiles, device="cuda")
        auto["value"], auto["ms_per_step"] = at / aw, 1000.0 * aw / args.steps
    if png is not None:
        pw, pt = sharding.reduce_job(dist, png_wall, job_tiles, device="cuda")
        png["value"], png["ms_per_step"] = pt / pw, 1000.0 * pw / args.steps
    raster_mean_ms, _ = sharding.reduce_job(dist, float(np.mean(raster_ms)), job_tiles, device="cuda")
    cover_mean_ms, _ = sharding.reduce_job(dist, float(np.mean(cover_ms)), job_tiles, device="cuda")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = total_tiles / dev_s
    e2e_value = total_tiles / e2e_wall

    # ---- roofline of the dominant kernel (raster_kernel) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth of this pool's B200)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # algorithmic bytes of one launch of the dominant kernel:
    #   raster_kernel: RGB written once + visible-op records, geometry records and row masks read once + one 8-byte alpha
    #                  per stored walk step read once
    #   line_cover_kernel: line records read once + one 8-byte alpha per in-line walk step and one length byte per walk written
    raster_bytes = (n_tiles * D * D * 3 + 32 * stats["n_visible_ops"] + stats["geom_bytes"] + stats["mask_bytes"]
                    + 8 * stats["walk_steps"])
    cover_bytes = stats["geom_bytes"] + 8 * stats["walk_steps"] + stats["walk_bytes"] // (8 * 4)
    if cover_mean_ms > raster_mean_ms:
        dom, dom_ms, algo_bytes = "line_cover_kernel", cover_mean_ms, cover_bytes
    else:
        dom, dom_ms, algo_bytes = "raster_kernel", raster_mean_ms, raster_bytes
    achieved = algo_bytes / (dom_ms / 1000.0) / 1e9
    traffic = None
    prof_json = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof_json):
        try:
            pj = json.load(open(prof_json))
            if pj.get("workload") == args.workload:
                traffic = pj.get(dom + "_dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {
        "bound": "hbm",
        "kernel": dom,
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": int(algo_bytes),
        "kernel_ms": dom_ms,
        "kernel_share_of_step": dom_ms / (1000.0 * dev_s / args.steps),
        # the write-only variant north_star quotes: whole-path tiles/s x 4*D^2 output bytes against the HBM peak
        "north_star_write_frac": (value / world) * (4.0 * D * D) / (peak * 1e9),
        "note": "the path is FP64/integer-issue and shared-memory bound, not HBM bound (SURVEY.md F6); see profiles/",
    }

    # ---- CPU baseline beside it (bounded sample, rank 0 only) ----
    cpu = None
    max_abs_diff = None
    if not args.skip_cpu_baseline and world == 1:  # the CPU baseline belongs to the N=1 line only
        threads = host_threads()
        probe_tps, _ = cpu_port_throughput(w, threads, threads)
        sample = args.cpu_sample or int(min(n_tiles, max(threads, probe_tps * 15.0)))
        tps, dt = cpu_port_throughput(w, sample, threads)
        cpu = {"value": tps, "unit": "tiles/s", "cores": threads, "kind": "port",
               "sample": f"{sample} tiles spread over the same batch, {dt:.1f}s, C++ restatement of the reference CPU path, one tile per thread"}
        # second half of BASELINE.json's metric: max |dRGB| of the GPU tiles (e2e output) against the CPU render
        sel, imgs = w["_last_cpu_sample"]
        gpu_out = np.frombuffer((C.c_uint8 * out_bytes).from_address(pin_out), dtype=np.uint8).reshape(n_tiles, D, D, 3)
        max_abs_diff = int(max(np.abs(gpu_out[i].astype(np.int16) - im.astype(np.int16)).max() for i, im in zip(sel, imgs)))

    # U (distinct nodes) for the SURVEY 8(d) whole-path byte count
    line = {
        "metric": f"{D}x{D} tiles/sec at z=14",
        "value": value,
        "unit": "tiles/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1000.0 * dev_s / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": cfg,
        "wall_ms_per_step": 1000.0 * wall / args.steps,
        "stage_ms": {"plan+geometry+fill_rows": float(np.mean(plan_ms)), "line_cover": float(np.mean(cover_ms)),
                     "raster": float(np.mean(raster_ms))},
        "batch_stats": {k: int(stats[k]) for k in ("n_tiles", "n_areas", "n_visible_ops", "n_node_refs", "geom_bytes", "mask_bytes",
                                                   "walk_bytes", "walk_steps")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_bytes,
                "ms_per_step": 1000.0 * e2e_wall / args.steps, "api": "osmr_draw_tiles (pinned host buffers; " + ("tiles stored straight into the host buffer by raster_kernel" if args.e2e_direct
                                                                      else "tiles staged in HBM, chunked D2H on a copy stream") + ")",
                "checksum": checksum},
        "pcie_probe": pcie,
        "e2e_auto": auto,
        "e2e_png": png,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "max_abs_diff_rgb_vs_cpu": max_abs_diff,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Wall-clock of osmr_draw_tiles vs osmr_draw_tiles_labeled on the fixture tile sets (run on the GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CONFIG_NAMES, FixtureInputs  # noqa: E402
from osm_renderer_b200.drawer import GpuContext  # noqa: E402

fx = FixtureInputs()
lt, font, per = fx.labels()
ctx = GpuContext(0)
ctx.set_geodata(fx.bin)
ctx.set_table(fx.table)
ctx.set_font(font)
ctx.set_label_table(lt)
for name in (sys.argv[1].split(',') if len(sys.argv) > 1 else CONFIG_NAMES):
    tiles, begins, areas = fx.batches[name]
    lb, labels = per[name]
    for _ in range(2):
        ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
        ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
    t1 = time.perf_counter()
    for _ in range(5):
        ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
    t2 = time.perf_counter()
    n = len(tiles)
    print(f"{name:6s} {n:3d} tiles  {len(labels):6d} label generations  areas only {1e3 * (t1 - t0) / 5 / n:7.3f} ms/tile   "
          f"with labels {1e3 * (t2 - t1) / 5 / n:7.3f} ms/tile", end="")
    st = ctx.stats()
    print(f"   [last draw: host layout {st['ms_label_layout'] / n:6.3f}  label kernels {st['ms_label_device'] / n:6.3f} ms/tile]")

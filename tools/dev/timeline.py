"""experiment (GPU box): timeline of one labelled host-output call (OSMR_TIMELINE=1), for label_chunks in argv"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["OSMR_TIMELINE"] = "1"
import bench
from osm_renderer_b200.drawer import GpuContext
w = bench.build_workload("C2", "mapnik", labels=True)
n = len(w["tiles"])
t, b, a, lb, ln = bench.sub_batch(w, np.arange(n), True)
ctx = GpuContext(0)
ctx.set_geodata(w["bin"]); ctx.set_table(w["table"]); ctx.set_font(w["font"]); ctx.set_label_table(w["ltable"])
for lc in [int(x) for x in sys.argv[1:]] or [1]:
    ctx.debug_set("label_chunks", lc)
    os.environ.pop("OSMR_TIMELINE", None)
    for _ in range(3):
        ctx.draw_tiles_labeled(t, b, a, lb, ln, w["canvas"], w["caps"])
    os.environ["OSMR_TIMELINE"] = "1"
    sys.stderr.write(f"==== label_chunks={lc}\n")
    t0 = time.perf_counter()
    ctx.draw_tiles_labeled(t, b, a, lb, ln, w["canvas"], w["caps"])
    sys.stderr.write(f"wall {1000 * (time.perf_counter() - t0):.2f} ms (pageable numpy buffers)\n")

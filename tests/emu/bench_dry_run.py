"""Dry run of bench.py's whole flow in the GPU-less container: torch.cuda is stubbed, the library is the emulated one and
the C2 workload is cut to 3x3 tiles.  Checks that the script assembles its JSON line (every key the driver reads) --
the numbers mean nothing.

    python tests/emu/bench_dry_run.py
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402

import build_emu  # noqa: E402


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self):
        pass

    def elapsed_time(self, other):
        return 1.0


def main():
    lib = build_emu.build()
    torch.cuda.set_device = lambda *_a, **_k: None
    torch.cuda.synchronize = lambda *_a, **_k: None
    torch.cuda.Event = _Event
    real_empty = torch.empty

    def empty(*a, **k):  # no CUDA tensors, no pinning
        k.pop("pin_memory", None)
        k["device"] = "cpu"
        return real_empty(*a, **k)

    torch.empty = empty
    import bench

    bench.WORKLOADS["C2"] = dict(zoom=14, x0=9900, y0=5118, n=3, scale=1, metro={})  # 9 tiles around the fixture tile
    sys.argv = ["bench.py", "--steps", "1", "--warmup", "1", "--lib", lib, "--cpu-sample", "4", "--min-seconds", "0.01", "--no-affinity"]
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "north_star_write_frac"):
        assert k in d["roofline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["max_abs_diff_rgb_vs_cpu"] == 0 and d["e2e_auto"]["identical_to_e2e_output"] and d["e2e_png"]["mean_png_bytes"] > 0
    assert d["max_abs_diff_rgb_vs_cpu_labeled"] == 0 and d["e2e_area_only"]["value"] > 0 and d["e2e_auto_png"]["value"] > 0
    assert "draw_tiles_labeled" in d["e2e"]["api"] and d["value_area_only"]["value"] > 0 and d["cpu_baseline"]["value_area_only"] > 0
    assert d["e2e_auto_labeled"]["identical_to_e2e_labeled_output"] and d["e2e_auto_labeled"]["label_path_last_call"] == 1
    assert d["e2e_auto_labeled"]["live_label_generations_last_call"] > 0 and d["e2e_auto_labeled_png"]["mean_png_bytes"] > 0
    print("e2e_auto_labeled:", d["e2e_auto_labeled"])
    assert d["sustained"]["steps"] >= 5 and d["latency_ms"] and d["roofline"]["B_tile_terms"]["sum_U"] > 0
    print("bench dry run ok:", {k: d[k] for k in ("metric", "n_gpus", "gpu_launches", "max_abs_diff_rgb_vs_cpu")})


if __name__ == "__main__":
    main()

"""Kernel logic on the host SIMT emulator (tests/emu/): the CUDA sources of libosmr_b200.so, compiled by g++ against
a fiber-per-thread emulator of warps / blocks, must render fixture tiles bit-for-bit like the oracle.  This checks the
kernels' control flow and integer / f64 arithmetic in the GPU-less container (divergent or deadlocking warp
collectives abort the run); it says nothing about the GPU build itself -- the `-m gpu` tests do that on the B200.
The whole `-m gpu` suite can be run the same way:  OSMR_TEST_EMU=1 python -m pytest tests -m gpu
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [["17", "0", "7", "12"], ["14", "0"], ["18_2x", "5"]])
def test_emulated_kernels_equal_oracle(args):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    build_emu.build()
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_emu.py")] + args, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing pixels vs oracle: 0" in res.stdout
    assert "[emu]" not in res.stderr, res.stderr  # divergent collective / deadlock reports


def _run(script_args, timeout=900):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    build_emu.build()
    return subprocess.run([sys.executable] + script_args, capture_output=True, text=True, timeout=timeout)


def test_emulated_device_side_lists_equal_host_lists():
    """f3 (osmr_draw_tiles_auto) on the emulator: same pixels and the same relative order as the host-built lists."""
    res = _run([os.path.join(ROOT, "tests", "emu", "run_auto.py"), "17", "4"])
    assert res.returncode == 0, res.stdout + res.stderr
    assert "order violations 0, differing pixels auto vs host lists: 0" in res.stdout
    assert "[emu]" not in res.stderr, res.stderr


def test_emulated_label_lists_from_a_tile_list_give_the_reference_golden():
    """f3 for the label pass (osmr_draw_tiles_auto_labeled) on the emulator: three z17 tiles from nothing but their coordinates
    equal the reference's golden render on every pixel; the lists are the host styler's minus the dead generations."""
    code = (
        "import sys, os, numpy as np\n"
        f"sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'tests')!r}, {os.path.join(ROOT, 'tests', 'emu')!r}]\n"
        "from run_emu import emu_context\n"
        "from conftest import FixtureInputs\n"
        "from autocheck import fixture_builder, subsequence_violations\n"
        "from osm_renderer_b200.upstream import pipeline\n"
        "from osm_renderer_b200.wire import LabelStyleTable\n"
        "fx = FixtureInputs(); ctx = emu_context()\n"
        "data, rd, S, table, fb = fixture_builder(icon_loader=fx.icon_loader())\n"
        "ltable = LabelStyleTable(None, icon_loader=fx.icon_loader())\n"
        "ac = pipeline.zoom_class_tables(fb, 17); lc = pipeline.zoom_label_class_tables(fb, 17, ltable)\n"
        "ctx.set_geodata(data); ctx.set_table(table); ctx.set_font(fx.labels()[1]); ctx.set_label_table(ltable)\n"
        "ctx.set_zoom_styles(17, *ac); ctx.set_zoom_label_styles(17, *lc)\n"
        "sel = [0, 7, 12]; tiles = fx.batches['17'][0][sel]\n"
        "got = ctx.draw_tiles_auto_labeled(tiles, S.canvas_fill_color, S.use_caps_for_dashes)\n"
        "golden = fx.golden('17')[0][sel]\n"
        "diff = (got != golden).any(axis=-1); diff[:, 0, :] = False; diff[:, :, 255] = False\n"
        "lb, labels = ctx.auto_readback_labels()\n"
        "viol = sum(subsequence_violations(fb.labels_array(z, x, y, ltable), labels[lb[t]:lb[t + 1]]) for t, (z, x, y, s) in enumerate(tiles.tolist()))\n"
        "st = ctx.stats()\n"
        "print('auto labelled: differing pixels vs golden', int(diff.sum()), 'order violations', viol, 'label path', st['label_path'], 'live', int(lb[-1]), st['n_labels_active'])\n"
        "assert diff.sum() == 0 and viol == 0 and st['label_path'] == 1 and int(lb[-1]) == st['n_labels_active'] > 0\n"
    )
    res = _run(["-c", code])
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing pixels vs golden 0 order violations 0" in res.stdout
    assert "[emu]" not in res.stderr, res.stderr


def test_emulated_sliced_plan_and_bin_lists():
    """the low-zoom form of plan_ops / bin_ops (a CTA per slice of a tile's op lists, debug key plan_slice_areas) draws the
    oracle's tiles"""
    code = (
        "import sys, os, numpy as np\n"
        f"sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'tests')!r}, {os.path.join(ROOT, 'tests', 'emu')!r}]\n"
        "from run_emu import emu_context\n"
        "from conftest import FixtureInputs\n"
        "import oracle\n"
        "fx = FixtureInputs(); ctx = emu_context(); ctx.set_geodata(fx.bin); ctx.set_table(fx.table)\n"
        "tiles, begins, areas = fx.batches['15']\n"
        "want = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, True, n_threads=4))\n"
        "ctx.debug_set('plan_slice_areas', 500)\n"
        "got = ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)\n"
        "print('sliced: differing pixels vs oracle', int((got != want).any(axis=-1).sum()), 'launches', ctx.stats()['kernel_launches'])\n"
        "assert (got == want).all() and ctx.stats()['kernel_launches'] > 8\n"
    )
    res = _run(["-c", code])
    assert res.returncode == 0, res.stdout + res.stderr
    assert "sliced: differing pixels vs oracle 0" in res.stdout
    assert "[emu]" not in res.stderr, res.stderr


def test_emulated_png_encoder_round_trips():
    """f4 (osmr_draw_tiles_png) on the emulator: the files pass a strict decoder and decode to the RGB tiles."""
    code = (
        "import sys, os, numpy as np\n"
        f"sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'tests')!r}, {os.path.join(ROOT, 'tests', 'emu')!r}]\n"
        "from run_emu import emu_context\n"
        "from conftest import FixtureInputs\n"
        "from test_gpu_png import decode_png\n"
        "fx = FixtureInputs(); ctx = emu_context(); ctx.set_geodata(fx.bin); ctx.set_table(fx.table)\n"
        "tiles, begins, areas = fx.batches['17']\n"
        "n = 2\n"
        "want = ctx.draw_tiles(tiles[:n], begins[:n + 1], areas[:begins[n]], fx.canvas_rgb, True)\n"
        "files = ctx.draw_tiles_png(tiles[:n], begins[:n + 1], areas[:begins[n]], fx.canvas_rgb, True)\n"
        "assert all((decode_png(f) == w).all() for f, w in zip(files, want))\n"
        "from test_gpu_png import _edge_case_images\n"
        "imgs = _edge_case_images()\n"
        "assert all((decode_png(f) == im).all() for f, im in zip(ctx.rgb_to_png(imgs), imgs))\n"
        "print('png ok', [len(f) for f in files])\n"
    )
    res = _run(["-c", code])
    assert res.returncode == 0, res.stdout + res.stderr
    assert "png ok" in res.stdout


def test_bench_script_assembles_its_json_line():
    """bench.py's whole flow (value, e2e, auto, png, pcie probe, roofline, cpu baseline) on the emulator with a 9-tile cut
    of the workload and torch.cuda stubbed: the driver-facing keys are all there, the outputs agree with the CPU."""
    res = _run([os.path.join(ROOT, "tests", "emu", "bench_dry_run.py")], timeout=1500)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "bench dry run ok" in res.stdout


def test_emulated_chunked_host_pipeline():
    """148 tiles in one call: the draw is split into chunks that alternate between the two scratch sets (the emulator runs them
    in order, so this checks the chunk plan, the per-chunk slices and offsets -- not the concurrency, which the GPU test does)."""
    idx = [str(i) for i in list(range(64)) + list(range(64)) + list(range(20))]
    res = _run([os.path.join(ROOT, "tests", "emu", "run_emu.py"), "18"] + idx, timeout=1500)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "differing pixels vs oracle: 0" in res.stdout

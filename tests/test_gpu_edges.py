"""-m gpu: edge cases of the C-ABI path (empty tiles, scratch overflow + redo, huge coordinates, resident batch API,
argument validation)."""
import numpy as np
import pytest

import oracle
from osm_renderer_b200.wire import AREA_DTYPE, OSMR_STYLE_COLOR, OSMR_STYLE_DASHES, OSMR_STYLE_WIDTH, STYLE_DTYPE, TILE_DTYPE, StyleTable

pytestmark = pytest.mark.gpu


def test_tiles_without_areas_are_canvas_colour(fx, gpu_ctx):
    tiles, begins, areas = fx.batches["16"]
    # tile 0 keeps its areas, tiles 1 and 2 get none
    b = np.array([0, begins[1], begins[1], begins[1]], dtype=np.uint32)
    got = gpu_ctx.draw_tiles(tiles[:3], b, areas[: begins[1]], fx.canvas_rgb, True)
    want = oracle.draw_tiles(fx.bin, fx.table, tiles[:3], b, areas[: begins[1]], fx.canvas_rgb, True)
    assert (got == np.stack(want)).all()
    assert (got[1] == np.array(fx.canvas_rgb, dtype=np.uint8)).all() and (got[2] == got[1]).all()


def test_scratch_overflow_grows_and_redoes_the_batch(fx):
    from osm_renderer_b200.drawer import GpuContext

    tiles, begins, areas = fx.batches["17"]
    want = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, True, n_threads=8))
    ctx = GpuContext(0)
    try:
        ctx.set_geodata(fx.bin)
        ctx.set_table(fx.table)
        ctx.debug_set("scratch_units", 64)  # far too small: every bump allocator overflows, the library must grow and redo
        ctx.debug_set("work_items", 16)  # ... and so do the work lists of the per-op kernels
        got = ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
    finally:
        ctx.close()
    assert (got == want).all()


def test_sliced_plan_lists_equal_the_single_cta_lists(fx):
    """debug key plan_slice_areas: every (tile, pass) list is cut into slices with a CTA each (the low-zoom form of
    plan_ops_kernel, C4) -- same tiles as the oracle's, also through the grow-and-redo path and the chunked host pipeline"""
    from osm_renderer_b200.drawer import GpuContext

    for name, per in (("14", 300), ("17", 256), ("16", 1000), ("18_2x", 200)):
        tiles, begins, areas = fx.batches[name]
        if name == "18_2x":  # @2x: four block areas per tile; six tiles are enough
            tiles, begins = tiles[30:36], begins[30:37] - begins[30]
            areas = areas[fx.batches[name][1][30] : fx.batches[name][1][36]]
        want = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, True, n_threads=8))
        ctx = GpuContext(0)
        try:
            ctx.set_geodata(fx.bin)
            ctx.set_table(fx.table)
            ctx.debug_set("plan_slice_areas", per)
            got = ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
            assert (got == want).all(), name
            assert ctx.stats()["kernel_launches"] > 8  # plan_count_kernel + plan_scan_kernel ran
            ctx.debug_set("scratch_units", 64)
            ctx.debug_set("work_items", 16)
            got = ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
            assert (got == want).all(), name + " (redo)"
            ctx.debug_set("host_chunks", 3)
            got = ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
            assert (got == want).all(), name + " (chunks)"
        finally:
            ctx.close()


def test_resident_batch_api_equals_one_shot(fx, gpu_ctx):
    tiles, begins, areas = fx.batches["16"]
    one = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
    gpu_ctx.batch_upload(tiles, begins, areas)
    out = np.empty_like(one)
    ms1 = gpu_ctx.batch_draw(fx.canvas_rgb, True, out=out)
    ms2 = gpu_ctx.batch_draw(fx.canvas_rgb, True)  # output stays in HBM
    assert (out == one).all() and ms1 > 0 and ms2 > 0
    st = gpu_ctx.stats()
    assert st["n_tiles"] == len(tiles) and st["n_areas"] == len(areas) and st["kernel_launches"] == 8  # style_calc, plan, geometry, fill_rows, bin, cover, raster, counters export


def test_huge_coordinates_take_the_exact_i64_path():
    """A way whose far end lies a quarter of the planet away: tile-relative coordinates beyond 2^24 at z18 @2x."""
    from osm_renderer_b200.drawer import GpuContext
    from osm_renderer_b200.upstream import synth
    from synthgeom import TX, TY

    b = synth._Builder()
    ts = b.tagset({"k": "v"})
    ox, oy = TX * 256, TY * 256
    for (x0, y0, x1, y1) in [(100, 40, 100 + 17_000_000, 40 + 6_000_000), (10, 200, 10 - 17_500_000, 200 - 9_000_000), (5, 5, 250, 240)]:
        ids = b.add_nodes([x0 + ox, x1 + ox], [y0 + oy, y1 + oy])
        b.way_nodes.append(ids)
        b.way_tags.append(ts)
    image = synth._serialise(b, with_index=False)  # a planet-sized bbox must not be enumerated into a tile index
    table = StyleTable(None)
    rows = np.zeros(2, dtype=STYLE_DTYPE)
    rows[0]["flags"] = OSMR_STYLE_COLOR | OSMR_STYLE_WIDTH
    rows[0]["color"] = (200, 30, 30)
    rows[0]["width"] = 5.0
    rows[0]["line_cap"] = 2
    rows[1]["flags"] = OSMR_STYLE_COLOR | OSMR_STYLE_WIDTH | OSMR_STYLE_DASHES
    rows[1]["color"] = (20, 60, 220)
    rows[1]["width"] = 2.0
    rows[1]["dashes_off"], rows[1]["dashes_len"] = 0, 2
    rows[1]["line_cap"] = 3
    table.rows = list(rows)
    table.dashes = [6.0, 3.0]
    areas = np.array([(0, 0), (1, 1), (2, 0), (0, 1)], dtype=AREA_DTYPE)
    for scale in (1, 2):
        tiles = np.array([(18, TX, TY, scale)], dtype=TILE_DTYPE)
        begins = np.array([0, len(areas)], dtype=np.uint32)
        ctx = GpuContext(0)
        try:
            ctx.set_geodata(image)
            ctx.set_table(table)
            got = ctx.draw_tiles(tiles, begins, areas, (255, 255, 255), True)
        finally:
            ctx.close()
        want = np.stack(oracle.draw_tiles(image, table, tiles, begins, areas, (255, 255, 255), True))
        assert (got == want).all(), scale
        assert (want != 255).any()


def test_argument_validation(fx, gpu_ctx):
    from osm_renderer_b200._lib import OsmrError

    tiles, begins, areas = fx.batches["14"]
    bad = tiles.copy()
    bad["scale"][1] = 2
    with pytest.raises(OsmrError):
        gpu_ctx.draw_tiles(bad, begins, areas, fx.canvas_rgb, True)
    bad = tiles.copy()
    bad["zoom"][0] = 40
    with pytest.raises(OsmrError):
        gpu_ctx.draw_tiles(bad, begins, areas, fx.canvas_rgb, True)
    badb = begins.copy()
    badb[2] = 1
    with pytest.raises(OsmrError):
        gpu_ctx.draw_tiles(tiles, badb, areas, fx.canvas_rgb, True)
    with pytest.raises(OsmrError):
        gpu_ctx.debug_set("fill_cap", 9999)


def _two_line_scene(width):
    from osm_renderer_b200.upstream import synth
    from osm_renderer_b200.wire import AREA_DTYPE, OSMR_STYLE_COLOR, OSMR_STYLE_DASHES, OSMR_STYLE_WIDTH, STYLE_DTYPE, TILE_DTYPE, StyleTable
    from synthgeom import TX, TY

    b = synth._Builder()
    ts = b.tagset({"k": "v"})
    ox, oy = TX * 256, TY * 256
    for (x0, y0, x1, y1) in [(-300, 40, 500, 200), (128, -200, 100, 400)]:
        ids = b.add_nodes([x0 + ox, x1 + ox], [y0 + oy, y1 + oy])
        b.way_nodes.append(ids)
        b.way_tags.append(ts)
    image = synth._serialise(b, with_index=False)
    table = StyleTable(None)
    rows = np.zeros(2, dtype=STYLE_DTYPE)
    rows[0]["flags"] = OSMR_STYLE_COLOR | OSMR_STYLE_WIDTH
    rows[0]["color"] = (200, 30, 30)
    rows[0]["width"] = width
    rows[0]["line_cap"] = 2
    rows[1]["flags"] = OSMR_STYLE_COLOR | OSMR_STYLE_WIDTH | OSMR_STYLE_DASHES
    rows[1]["color"] = (20, 60, 220)
    rows[1]["width"] = width / 2
    rows[1]["opacity"] = 1.0
    rows[1]["dashes_off"], rows[1]["dashes_len"] = 0, 2
    rows[1]["line_cap"] = 2
    table.rows = list(rows)
    table.dashes = [40.0, 25.0]
    areas = np.array([(0, 0), (1, 1)], dtype=AREA_DTYPE)
    tiles = np.array([(18, TX, TY, 1)], dtype=TILE_DTYPE)
    return image, table, tiles, np.array([0, 2], dtype=np.uint32), areas


@pytest.mark.parametrize("width", [240.0, 260.0, 700.0])
def test_very_wide_lines_are_drawn_like_the_reference(width):
    """Round 1 cached walk lengths in 7 bits and refused lines wider than 248 px; the reference has no such limit.  Fragments
    carry no length, so any width the reference draws is drawn (bit-exact) -- here up to lines wider than the whole tile."""
    from osm_renderer_b200.drawer import GpuContext

    image, table, tiles, begins, areas = _two_line_scene(width)
    ctx = GpuContext(0)
    try:
        ctx.set_geodata(image)
        ctx.set_table(table)
        got = ctx.draw_tiles(tiles, begins, areas, (255, 255, 255), True)
    finally:
        ctx.close()
    want = np.stack(oracle.draw_tiles(image, table, tiles, begins, areas, (255, 255, 255), True))
    assert (got == want).all()


def test_chunked_host_output_pipeline(fx, gpu_ctx):
    """148 tiles in one call: the library draws them in chunks on two compute streams with separate scratch while a third
    stream copies finished chunks back; every tile must still be the oracle's."""
    tiles, begins, areas = fx.batches["18"]
    sel = list(range(64)) + list(range(64)) + list(range(20))
    parts = [areas[begins[i] : begins[i + 1]] for i in sel]
    b = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    a = np.concatenate(parts)
    uniq = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles[:64], begins[:65], areas[: begins[64]], fx.canvas_rgb, True, n_threads=8))
    for two in (1, 0):
        gpu_ctx.debug_set("two_streams", two)
        got = gpu_ctx.draw_tiles(tiles[sel], b, a, fx.canvas_rgb, True)
        assert (got == uniq[sel]).all(), two
    gpu_ctx.debug_set("two_streams", 1)
    gpu_ctx.debug_set("host_chunks", 2)
    got = gpu_ctx.draw_tiles(tiles[sel], b, a, fx.canvas_rgb, True)
    gpu_ctx.debug_set("host_chunks", 0)
    assert (got == uniq[sel]).all()


def test_opacity_outside_the_exact_range_is_refused(fx):
    """The compositor keeps RGB only, which is exact while every source alpha a satisfies a + fl(1 - a) == 1 (a in [0, 2]);
    the reference does not clamp opacity, so a table outside that range is refused instead of drawn differently."""
    import copy

    from osm_renderer_b200._lib import OsmrError
    from osm_renderer_b200.drawer import GpuContext
    from osm_renderer_b200.wire import OSMR_STYLE_FILL_OPACITY, OSMR_STYLE_OPACITY, StyleTable

    ctx = GpuContext(0)
    for field, flag, val in (("opacity", OSMR_STYLE_OPACITY, 2.5), ("fill_opacity", OSMR_STYLE_FILL_OPACITY, -0.25), ("opacity", OSMR_STYLE_OPACITY, float("nan"))):
        t = StyleTable(None)
        t.rows = [r.copy() for r in fx.table.rows]
        t.dashes = list(fx.table.dashes)
        t.rows[0][field] = val
        t.rows[0]["flags"] |= flag
        with pytest.raises(OsmrError, match="outside"):
            ctx.set_table(t)
    t = StyleTable(None)  # the boundary values are fine
    t.rows = [r.copy() for r in fx.table.rows]
    t.dashes = list(fx.table.dashes)
    t.rows[0]["opacity"] = 2.0
    t.rows[0]["flags"] |= OSMR_STYLE_OPACITY
    ctx.set_table(t)
    ctx.close()


def test_worker_contexts_share_one_resident_dataset(fx):
    """osmr_ctx_create_shared: the reference shares reader / styler / drawer between worker threads and gives each thread only
    its own TilePixels (http_server.rs:42-48,69-72).  A worker context draws from the parent's geodata without uploading it
    again, from its own host thread, and replacing the dataset of one context leaves the other one alone."""
    import threading

    from osm_renderer_b200.drawer import GpuContext
    from osm_renderer_b200.upstream import synth

    tiles, begins, areas = fx.batches["17"]
    want = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, True, n_threads=8))
    parent = GpuContext(0)
    parent.set_geodata(fx.bin)
    parent.set_table(fx.table)
    workers = [GpuContext(0, share_dataset_of=parent) for _ in range(3)]
    results = [None] * len(workers)

    def run(i):
        w = workers[i]
        w.set_table(fx.table)  # style tables are per context
        results[i] = w.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)

    import os

    if os.environ.get("OSMR_TEST_EMU") == "1":  # the host emulator is a single fiber engine: one context at a time
        for i in range(len(workers)):
            run(i)
        mine = parent.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
    else:
        threads = [threading.Thread(target=run, args=(i,)) for i in range(len(workers))]
        for t in threads:
            t.start()
        mine = parent.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)  # the parent draws at the same time
        for t in threads:
            t.join()
    assert (mine == want).all()
    for r in results:
        assert r is not None and (r == want).all()
    # a new dataset on one worker: the others (and the parent) keep drawing the old one
    other = synth.make_metro(n=1)
    workers[0].set_geodata(other)
    assert (workers[1].draw_tiles(tiles, begins, areas, fx.canvas_rgb, True) == want).all()
    assert (parent.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True) == want).all()
    # the parent goes away first: the shared dataset lives as long as a context uses it
    parent.close()
    assert (workers[2].draw_tiles(tiles, begins, areas, fx.canvas_rgb, True) == want).all()
    for w in workers:
        w.close()

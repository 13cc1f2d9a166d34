"""SURVEY.md 8(f) row f3 for the LABEL pass on the GPU: osmr_draw_tiles_auto_labeled takes a tile list and builds the styled-area
lists AND the label lists on the device (reference reader.rs:60-100 with nodes, styler.rs:115-203 with for_labels = true,
drawer.rs:106-119), then runs the whole draw_to_pixels.  Checked against the REFERENCE'S OWN golden renders on every pixel,
against the labelled draw from host-built lists, and list by list against the host styler's order."""
import numpy as np
import pytest

from autocheck import fixture_builder, subsequence_violations
from conftest import CONFIG_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def built(fx):
    from osm_renderer_b200.wire import LabelStyleTable

    data, rd, S, table, fb = fixture_builder(icon_loader=fx.icon_loader())
    ltable = LabelStyleTable(None, icon_loader=fx.icon_loader())
    return data, rd, S, table, fb, ltable


@pytest.fixture()
def ctx():
    from osm_renderer_b200.drawer import GpuContext

    c = GpuContext(0)
    yield c
    c.close()


def prepare(ctx, fx, built, zoom):
    from osm_renderer_b200.upstream import pipeline

    data, rd, S, table, fb, ltable = built
    area_classes = pipeline.zoom_class_tables(fb, zoom)  # (styles every entity once: completes `table` / `ltable`)
    label_classes = pipeline.zoom_label_class_tables(fb, zoom, ltable)
    ctx.set_geodata(data)
    ctx.set_table(table)
    ctx.set_font(fx.labels()[1])
    ctx.set_label_table(ltable)
    ctx.set_zoom_styles(zoom, *area_classes)
    ctx.set_zoom_label_styles(zoom, *label_classes)
    return S, fb, ltable


def grid_mask(d):
    grid = np.zeros((d, d), dtype=bool)  # the red test grid of tests/test_rendering.rs:109-114
    grid[0, :] = True
    grid[:, d - 1] = True
    return grid


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_tile_list_in_reference_golden_out(fx, built, ctx, name):
    tiles = fx.batches[name][0]
    zoom = int(tiles["zoom"][0])
    S, fb, ltable = prepare(ctx, fx, built, zoom)
    got = ctx.draw_tiles_auto_labeled(tiles, S.canvas_fill_color, S.use_caps_for_dashes)
    st = ctx.stats()
    golden, _ = fx.golden(name)
    diff = (got != golden).any(axis=-1) & ~grid_mask(golden.shape[1])[None]
    assert diff.sum() == 0, f"{name}: {diff.sum()} pixels differ from the reference golden in tiles {sorted(set(np.argwhere(diff)[:, 0].tolist()))[:10]}"
    assert st["label_path"] == 1 and st["ms_auto"] > 0
    # the lists themselves: live generations only, in the host styler's order
    lb, labels = ctx.auto_readback_labels()
    n_host = 0
    for t, (z, x, y, s) in enumerate(tiles.tolist()):
        host = fb.labels_array(z, x, y, ltable)
        n_host += len(host)
        assert subsequence_violations(host, labels[lb[t] : lb[t + 1]]) == 0, f"tile {t}"
    assert 0 < int(lb[-1]) <= n_host
    assert st["n_labels_active"] == int(lb[-1])  # nothing listed that label_select_kernel then drops


def test_auto_labeled_equals_host_lists_and_png(fx, built, ctx):
    """the same pixels as osmr_draw_tiles_labeled from host-built lists; the PNG form decodes to them"""
    tiles, begins, areas = fx.batches["17"]
    S, fb, ltable = prepare(ctx, fx, built, 17)
    parts = [fb.labels_array(z, x, y, ltable) for (z, x, y, s) in tiles.tolist()]
    lb = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    aparts = [fb.areas_array(z, x, y) for (z, x, y, s) in tiles.tolist()]
    ab = np.concatenate([[0], np.cumsum([len(p) for p in aparts])]).astype(np.uint32)
    want = ctx.draw_tiles_labeled(tiles, ab, np.concatenate(aparts), lb, np.concatenate(parts), S.canvas_fill_color, S.use_caps_for_dashes)
    got = ctx.draw_tiles_auto_labeled(tiles, S.canvas_fill_color, S.use_caps_for_dashes)
    assert (got == want).all()
    files = ctx.draw_tiles_auto_labeled_png(tiles[:4], S.canvas_fill_color, S.use_caps_for_dashes)
    from test_gpu_png import decode_png

    for t, png in enumerate(files):
        assert (decode_png(png) == want[t]).all()


def test_auto_labeled_host_layout_fallback(fx, built, ctx):
    """debug key label_host: the device-built lists come back once and the host lays the labels out -- same pixels"""
    tiles = fx.batches["16"][0]
    S, fb, ltable = prepare(ctx, fx, built, 16)
    dev = ctx.draw_tiles_auto_labeled(tiles, S.canvas_fill_color, S.use_caps_for_dashes)
    assert ctx.stats()["label_path"] == 1
    ctx.debug_set("label_host", 1)
    try:
        host = ctx.draw_tiles_auto_labeled(tiles, S.canvas_fill_color, S.use_caps_for_dashes)
        assert ctx.stats()["label_path"] == 2
    finally:
        ctx.debug_set("label_host", 0)
    assert (host == dev).all()


def test_auto_labeled_errors(fx, built, ctx):
    from osm_renderer_b200._lib import OsmrError
    from osm_renderer_b200.upstream import pipeline
    from osm_renderer_b200.wire import TILE_DTYPE

    data, rd, S, table, fb, ltable = built
    tiles = np.array([(17, 79229, 40974, 1)], dtype=TILE_DTYPE)
    area_classes = pipeline.zoom_class_tables(fb, 17)
    nc, wc, mc, cb, cs = pipeline.zoom_label_class_tables(fb, 17, ltable)
    ctx.set_geodata(data)
    ctx.set_table(table)
    ctx.set_zoom_styles(17, *area_classes)
    with pytest.raises(OsmrError, match="osmr_set_font"):
        ctx.draw_tiles_auto_labeled(tiles, None)
    ctx.set_font(fx.labels()[1])
    ctx.set_label_table(ltable)
    with pytest.raises(OsmrError, match="osmr_set_zoom_label_styles"):
        ctx.draw_tiles_auto_labeled(tiles, None)
    bad = cs.copy()
    bad["order"][0] = 1 << 19
    with pytest.raises(OsmrError, match="below 2\\^19"):
        ctx.set_zoom_label_styles(17, nc, wc, mc, cb, bad)
    bad = cs.copy()
    bad["style"][:] = 100_000
    ctx.set_zoom_label_styles(17, nc, wc, mc, cb, bad)
    with pytest.raises(OsmrError, match="label style that does not exist"):
        ctx.draw_tiles_auto_labeled(tiles, None)
    ctx.set_zoom_label_styles(17, nc, wc, mc, cb, cs)
    ctx.draw_tiles_auto_labeled(tiles, None)


def test_auto_labeled_on_the_synthetic_metro(ctx):
    """32 tiles of the bench workload (C2): named streets along the line, building names centred, thousands of dead generations"""
    import bench
    from osm_renderer_b200.upstream import pipeline

    w = bench.build_workload("C2", labels=True, max_tiles=32)
    tiles, ab, areas, lb, labels = bench.sub_batch(w, np.arange(len(w["tiles"])), True)
    zoom = int(tiles["zoom"][0])
    area_classes = pipeline.zoom_class_tables(w["builder"], zoom)
    label_classes = pipeline.zoom_label_class_tables(w["builder"], zoom, w["ltable"])
    ctx.set_geodata(w["bin"])
    ctx.set_table(w["table"])
    ctx.set_font(w["font"])
    ctx.set_label_table(w["ltable"])
    ctx.set_zoom_styles(zoom, *area_classes)
    ctx.set_zoom_label_styles(zoom, *label_classes)
    want = ctx.draw_tiles_labeled(tiles, ab, areas, lb, labels, w["canvas"], w["caps"])
    n_active = ctx.stats()["n_labels_active"]
    got = ctx.draw_tiles_auto_labeled(tiles, w["canvas"], w["caps"])
    assert (got == want).all()
    dlb, dl = ctx.auto_readback_labels()
    assert int(dlb[-1]) == n_active and 0 < n_active < len(labels)
    for t in range(len(tiles)):
        assert subsequence_violations(labels[lb[t] : lb[t + 1]], dl[dlb[t] : dlb[t + 1]]) == 0

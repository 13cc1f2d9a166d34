"""SURVEY.md 8(f) row f3 on the GPU: osmr_draw_tiles_auto builds the styled-area lists on the device (candidate lookup in
the .bin tile index, dedup, painter's order) and must draw exactly what osmr_draw_tiles draws from the host-built lists
(reference reader.rs:60-180, styler.rs:115-203,246-272 as restated in osm_renderer_b200/upstream)."""
import numpy as np
import pytest

from autocheck import check_auto, fixture_builder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def built():
    return fixture_builder()


@pytest.fixture()
def ctx():
    from osm_renderer_b200.drawer import GpuContext

    c = GpuContext(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["14", "15", "16", "17", "18", "18_2x"])
def test_auto_equals_host_lists_on_every_fixture_tile(fx, built, ctx, name):
    data, rd, S, table, fb = built
    tiles = [tuple(int(v) for v in t) for t in fx.batches[name][0]]
    bad, bad_order, n_dev, n_host = check_auto(ctx, data, S, table, fb, tiles)
    assert bad == 0 and bad_order == 0
    assert 0 < n_dev <= n_host  # areas that cannot touch the tile are dropped before ordering


def test_auto_on_the_synthetic_metro(ctx):
    """32 tiles of the bench workload (C2): dense candidate lists, multipolygons, every street class."""
    import bench

    w = bench.build_workload("C2", max_tiles=32)
    from osm_renderer_b200.upstream import pipeline
    from osm_renderer_b200.wire import TILE_DTYPE

    sel = np.arange(len(w["tiles"]))
    ab = w["area_begin"]
    parts = [w["areas"][ab[i] : ab[i + 1]] for i in sel]
    begins = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    tiles = w["tiles"][sel]
    wc, mc, cb, cs = pipeline.zoom_class_tables(w["builder"], int(tiles["zoom"][0]))
    ctx.set_geodata(w["bin"])
    ctx.set_table(w["table"])
    ctx.set_zoom_styles(int(tiles["zoom"][0]), wc, mc, cb, cs)
    want = ctx.draw_tiles(tiles, begins, np.concatenate(parts), w["canvas"], w["caps"])
    got = ctx.draw_tiles_auto(tiles, w["canvas"], w["caps"])
    assert (got == want).all()
    st = ctx.stats()
    assert st["n_areas"] < len(np.concatenate(parts)) and st["ms_auto"] > 0


def test_auto_errors(fx, built, ctx):
    from osm_renderer_b200._lib import OsmrError
    from osm_renderer_b200.upstream import pipeline
    from osm_renderer_b200.wire import TILE_DTYPE

    data, rd, S, table, fb = built
    tiles = np.array([(17, 79222, 40978, 1)], dtype=TILE_DTYPE)
    with pytest.raises(OsmrError, match="osmr_set_geodata"):
        ctx.draw_tiles_auto(tiles, None)
    wc, mc, cb, cs = pipeline.zoom_class_tables(fb, 17)
    ctx.set_geodata(data)
    ctx.set_table(table)
    with pytest.raises(OsmrError, match="osmr_set_zoom_styles"):
        ctx.draw_tiles_auto(tiles, None)
    ctx.set_zoom_styles(17, wc, mc, cb, cs)
    ctx.draw_tiles_auto(tiles, None)
    mixed = np.array([(17, 79222, 40978, 1), (16, 39611, 20489, 1)], dtype=TILE_DTYPE)
    with pytest.raises(OsmrError, match="one zoom"):
        ctx.draw_tiles_auto(mixed, None)
    bad = cs.copy()
    bad["style"][0] = 10_000
    ctx.set_zoom_styles(17, wc, mc, cb, bad)
    with pytest.raises(OsmrError, match="style that does not exist"):
        ctx.draw_tiles_auto(tiles, None)
    with pytest.raises(OsmrError, match="zoom must be"):
        ctx.set_zoom_styles(19, wc, mc, cb, cs)

"""-m gpu: the remaining BASELINE.json configs as parity cases (bit-exact vs the oracle) on a reduced synthetic city.

C3  @2x tiles (512x512)                                  -- widths / dashes x2, patterns unscaled
C4  zoom sweep z10..18 over the same geodata image        -- from everything-in-a-few-pixels to a few huge polygons
C5  dense polygons: thousands of footprints + a long noisy coastline multipolygon with islands, z16
The full-size versions are bench workloads (bench.py --workload C2|C3); here the sizes are chosen so that the CPU
oracle finishes in seconds."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def city():
    from osm_renderer_b200.upstream import geodata, mapcss, pipeline, styler as st, synth
    from osm_renderer_b200.wire import StyleTable

    data = synth.make_metro(n=4, extra_footprints=6000, coastline_nodes=6000)  # 4x4 z14 tiles
    rd = geodata.GeodataReader(data)
    S = st.Styler(mapcss.load_rules_json(os.path.join(GOLDEN, "mapnik_rules.json.gz")), "josm", None)
    table = StyleTable(None)
    # the coastline multipolygon has no rule in mapnik.mapcss: give it a style through a landuse tag it does know
    fb = pipeline.FastBatchBuilder(rd, S, table)
    return data, rd, S, table, fb


def _batch(fb, tiles):
    from osm_renderer_b200.wire import TILE_DTYPE

    parts = [fb.areas_array(z, x, y) for (z, x, y, s) in tiles]
    begins = np.zeros(len(tiles) + 1, dtype=np.uint32)
    begins[1:] = np.cumsum([len(p) for p in parts])
    return np.array(tiles, dtype=TILE_DTYPE), begins, np.concatenate(parts)


def _check(city, tiles, extra_areas=None):
    from osm_renderer_b200.drawer import GpuContext

    data, rd, S, table, fb = city
    t, b, a = _batch(fb, tiles)
    if extra_areas is not None:
        a, b = extra_areas(t, b, a)
    ctx = GpuContext(0)
    try:
        ctx.set_geodata(data)
        ctx.set_table(table)
        got = ctx.draw_tiles(t, b, a, S.canvas_fill_color, S.use_caps_for_dashes)
    finally:
        ctx.close()
    want = np.stack(oracle.draw_tiles(data, table, t, b, a, S.canvas_fill_color, S.use_caps_for_dashes, n_threads=8))
    bad = (got != want).any(axis=-1)
    assert bad.sum() == 0, f"{bad.sum()} differing pixels in tiles {sorted(set(np.argwhere(bad)[:, 0].tolist()))}"
    return got


def test_c3_2x_tiles(city):
    _check(city, [(14, 9888 + i, 5104 + j, 2) for i in range(2) for j in range(2)])


@pytest.mark.parametrize("zoom", list(range(10, 19)))
def test_c4_zoom_sweep(city, zoom):
    # the tiles of `zoom` covering the centre of the 4x4 z14 block
    cx, cy = 9888 * 2 + 4, 5104 * 2 + 4  # centre in z15 tile units
    if zoom >= 15:
        f = 1 << (zoom - 15)
        tiles = [(zoom, cx * f + i, cy * f + j, 1) for i in range(2) for j in range(2)]
    else:
        f = 1 << (15 - zoom)
        tiles = [(zoom, cx // f, cy // f, 1)]
    got = _check(city, tiles)
    assert got.shape[0] == len(tiles)


def test_c5_dense_polygons_and_coastline(city):
    """z16 tiles with thousands of footprints and the coastline ring + islands filled as one multipolygon."""
    from osm_renderer_b200.wire import AREA_DTYPE, OSMR_AREA_MULTIPOLYGON, OSMR_STYLE_FILL_COLOR, OSMR_STYLE_FILL_OPACITY, STYLE_DTYPE

    data, rd, S, table, fb = city
    # a water-coloured fill style for the coastline multipolygon (last multipolygon of the image), drawn first
    row = np.zeros((), dtype=STYLE_DTYPE)
    row["flags"] = OSMR_STYLE_FILL_COLOR | OSMR_STYLE_FILL_OPACITY
    row["fill_color"] = (170, 200, 230)
    row["fill_opacity"] = 0.8
    row["fill_image"] = -1
    table.rows.append(row)
    coast_style = len(table.rows) - 1
    coast_entity = (len(rd.multipolygons) - 1) | OSMR_AREA_MULTIPOLYGON

    def add_coast(t, b, a):
        parts, begins = [], [0]
        for i in range(len(t)):
            extra = np.array([(coast_entity, coast_style)], dtype=AREA_DTYPE)
            parts += [extra, a[b[i] : b[i + 1]]]
            begins.append(begins[-1] + 1 + int(b[i + 1] - b[i]))
        return np.concatenate(parts), np.asarray(begins, dtype=np.uint32)

    tiles = [(16, 9888 * 4 + i, 5104 * 4 + j, 1) for i in (2, 7, 11) for j in (3, 8, 12)]
    got = _check(city, tiles, add_coast)
    assert ((got == np.array([170, 200, 230])).all(axis=-1)).sum() == 0  # 0.8 opacity: never the pure colour
    assert (got != np.array(S.canvas_fill_color, dtype=np.uint8)).any()

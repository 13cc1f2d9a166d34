"""Shared helpers of the f3 tests (device-side candidate lookup + painter's order, osmr_draw_tiles_auto)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def fixture_builder(icon_loader=None):
    """nano_moscow.bin + the committed mapnik rules -> (image, reader, styler, style table, FastBatchBuilder);
    `icon_loader` (name -> (w, h, rgba) or None) resolves fill-image patterns (without one they fail to load, which the
    reference treats as "no fill")"""
    from osm_renderer_b200.upstream import geodata, mapcss, pipeline, styler as st
    from osm_renderer_b200.wire import StyleTable

    with open(os.path.join(GOLDEN, "nano_moscow.bin"), "rb") as f:
        data = f.read()
    rd = geodata.GeodataReader(data)
    rules = mapcss.load_rules_json(os.path.join(GOLDEN, "mapnik_rules.json.gz"))
    S = st.Styler(rules, "josm", None)
    table = StyleTable(None, icon_loader=icon_loader)
    return data, rd, S, table, pipeline.FastBatchBuilder(rd, S, table)


def subsequence_violations(host, dev) -> int:
    """number of device entries that do not continue a common subsequence of the host list (same relative order)"""
    hk = host["entity"].astype(np.uint64) << np.uint64(32) | host["style"].astype(np.uint64)
    dk = dev["entity"].astype(np.uint64) << np.uint64(32) | dev["style"].astype(np.uint64)
    pos = {}
    for i, k in enumerate(hk):
        pos.setdefault(int(k), []).append(i)
    last, bad = -1, 0
    for k in dk:
        lst = pos.get(int(k))
        nxt = next((p for p in lst if p > last), None) if lst else None
        if nxt is None:
            bad += 1
        else:
            last = nxt
    return bad


def check_auto(ctx, data, S, table, fb, tiles):
    """Draw `tiles` [(zoom, x, y, scale)] through osmr_draw_tiles (host-built lists) and osmr_draw_tiles_auto.
    Returns (#differing pixels, #order violations, #device styled areas, #host styled areas)."""
    from osm_renderer_b200.upstream import pipeline
    from osm_renderer_b200.wire import TILE_DTYPE

    zoom = tiles[0][0]
    parts = [fb.areas_array(z, x, y) for (z, x, y, s) in tiles]
    begins = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    areas = np.concatenate(parts)
    tarr = np.array(tiles, dtype=TILE_DTYPE)
    wc, mc, cb, cs = pipeline.zoom_class_tables(fb, zoom)  # styles every entity once: completes `table`
    ctx.set_geodata(data)
    ctx.set_table(table)
    ctx.set_zoom_styles(zoom, wc, mc, cb, cs)
    want = ctx.draw_tiles(tarr, begins, areas, S.canvas_fill_color, S.use_caps_for_dashes)
    got = ctx.draw_tiles_auto(tarr, S.canvas_fill_color, S.use_caps_for_dashes)
    ab, aa = ctx.auto_readback()
    bad_order = sum(subsequence_violations(parts[t], aa[ab[t] : ab[t + 1]]) for t in range(len(tiles)))
    return int((got != want).any(axis=-1).sum()), bad_order, int(ab[-1]), len(areas)

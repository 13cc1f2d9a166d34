"""Random integer-pixel geometry + random styles for parity fuzzing (shared by CPU and GPU tests).

Nodes are placed at integer z18 pixel positions, so Point::from_node lands exactly on the intended pixel of the
z18 test tile (and on half pixels of lower zooms, exercising the round-half-away rule)."""
import numpy as np

from osm_renderer_b200.upstream import synth
from osm_renderer_b200.wire import (
    AREA_DTYPE,
    OSMR_AREA_MULTIPOLYGON,
    OSMR_STYLE_CASING_COLOR,
    OSMR_STYLE_CASING_DASHES,
    OSMR_STYLE_CASING_WIDTH,
    OSMR_STYLE_COLOR,
    OSMR_STYLE_DASHES,
    OSMR_STYLE_FILL_COLOR,
    OSMR_STYLE_FILL_IMAGE,
    OSMR_STYLE_FILL_OPACITY,
    OSMR_STYLE_OPACITY,
    OSMR_STYLE_WIDTH,
    STYLE_DTYPE,
    TILE_DTYPE,
    StyleTable,
)

TX, TY = 158460, 81950  # z18 tile the geometry is built around


def random_scene(seed: int, n_ways: int = 60, n_mps: int = 6, span: int = 420, scale: int = 1, zoom: int = 18):
    """Returns (bin image, StyleTable, tiles, area_begin, areas)."""
    rng = np.random.default_rng(seed)
    b = synth._Builder()
    ox, oy = ((TX >> (18 - zoom)) << (18 - zoom)) * 256, ((TY >> (18 - zoom)) << (18 - zoom)) * 256  # origin of the tile at `zoom`
    ts = b.tagset({"k": "v"})

    shift = 18 - zoom
    unit = float(1 << shift)  # z18 pixels per pixel of `zoom`

    def pts(k, lo=-120, hi=span):
        # integer pixels of the target zoom plus a fraction in {0, 1/8, 5/16} (x scale 1 or 2): never an exact .5 tie, where the last
        # bit of the device's tan/log (not glibc's) would decide the rounding
        fx = rng.choice([0.0, 0.125, 0.3125], size=k) if shift else 0.0
        fy = rng.choice([0.0, 0.125, 0.3125], size=k) if shift else 0.0
        return (rng.integers(lo, hi, size=k) + fx) * unit + ox, (rng.integers(lo, hi, size=k) + fy) * unit + oy

    ways = []
    for i in range(n_ways):
        kind = rng.choice(["poly", "line", "long", "degenerate", "axis"], p=[0.35, 0.35, 0.1, 0.1, 0.1])
        if kind == "poly":
            k = int(rng.integers(3, 14))
            xs, ys = pts(k)
            ids = b.add_nodes(xs, ys)
            ids = np.concatenate([ids, ids[:1]])
        elif kind == "line":
            k = int(rng.integers(2, 9))
            xs, ys = pts(k)
            ids = b.add_nodes(xs, ys)
        elif kind == "long":  # endpoints far outside the tile
            xs, ys = pts(2, -40000, 40000)
            ids = b.add_nodes(xs, ys)
        elif kind == "axis":  # horizontal / vertical / 45 degree runs (ties in every Bresenham variant)
            x0, y0 = int(rng.integers(0, 256)), int(rng.integers(0, 256))
            d = int(rng.integers(5, 200))
            dirs = [(d, 0), (0, d), (d, d), (d, -d), (-d, 0), (0, -d)]
            dx, dy = dirs[int(rng.integers(0, len(dirs)))]
            ids = b.add_nodes(
                [x0 * unit + ox, (x0 + dx) * unit + ox, (x0 + dx) * unit + ox],
                [y0 * unit + oy, (y0 + dy) * unit + oy, (y0 + dy + int(rng.integers(0, 2))) * unit + oy],
            )
        else:  # repeated points -> degenerate pairs, also first/last (caps are skipped, line.rs:33)
            xs, ys = pts(3)
            ids = b.add_nodes(np.repeat(xs, 2), np.repeat(ys, 2))
        b.way_nodes.append(np.asarray(ids, dtype=np.int64))
        b.way_tags.append(ts)
        ways.append(i)
    for _ in range(n_mps):
        pids = []
        for _ in range(int(rng.integers(1, 5))):
            k = int(rng.integers(3, 10))
            xs, ys = pts(k)
            ids = b.add_nodes(xs, ys)
            pids.append(len(b.polys))
            b.polys.append(np.concatenate([ids, ids[:1]]))
        b.mps.append((pids, ts))
    image = synth._serialise(b)

    table = StyleTable(None)
    icon = rng.integers(0, 256, size=(7, 5, 4), dtype=np.uint8)
    icon[..., 3] = rng.choice([0, 77, 255], size=(7, 5))
    table.add_raw_icon("pattern", icon)
    n_styles = 24
    rows = np.zeros(n_styles, dtype=STYLE_DTYPE)
    dashes = []
    dash_choices = [[4.0, 2.0], [0.0, 12.0, 10.0, 152.0], [1.0, 3.0], [7.5], [2.0, 2.0, 6.0, 2.0], [0.5, 0.5], [30.0, 1.0, 0.0, 4.0]]
    for i in range(n_styles):
        f = 0
        r = rows[i]
        if rng.random() < 0.7:
            f |= OSMR_STYLE_COLOR
            r["color"] = rng.integers(0, 256, 3)
        if rng.random() < 0.8:
            f |= OSMR_STYLE_WIDTH
            r["width"] = float(rng.choice([0.2, 0.5, 1.0, 1.5, 2.0, 3.0, 4.5, 7.0, 13.0, 16.0]))
        if rng.random() < 0.3:
            f |= OSMR_STYLE_OPACITY
            r["opacity"] = float(rng.choice([0.25, 0.5, 0.9, 1.0]))
        if rng.random() < 0.4:
            f |= OSMR_STYLE_DASHES
            d = dash_choices[int(rng.integers(0, len(dash_choices)))]
            r["dashes_off"], r["dashes_len"] = len(dashes), len(d)
            dashes += d
        r["line_cap"] = int(rng.integers(0, 4))
        if rng.random() < 0.4:
            f |= OSMR_STYLE_CASING_COLOR | OSMR_STYLE_CASING_WIDTH
            r["casing_color"] = rng.integers(0, 256, 3)
            r["casing_width"] = float(r["width"]) + float(rng.choice([0.5, 1.0, 2.0, 3.0]))
            r["casing_line_cap"] = int(rng.integers(0, 4))
            if rng.random() < 0.3:
                f |= OSMR_STYLE_CASING_DASHES
                d = dash_choices[int(rng.integers(0, len(dash_choices)))]
                r["casing_dashes_off"], r["casing_dashes_len"] = len(dashes), len(d)
                dashes += d
        u = rng.random()
        r["fill_image"] = -1
        if u < 0.45:
            f |= OSMR_STYLE_FILL_COLOR
            r["fill_color"] = rng.integers(0, 256, 3)
        elif u < 0.6:
            f |= OSMR_STYLE_FILL_IMAGE
            r["fill_image"] = 0 if rng.random() < 0.8 else -1  # -1: icon failed to load -> area skipped
        if rng.random() < 0.5:
            f |= OSMR_STYLE_FILL_OPACITY
            r["fill_opacity"] = float(rng.choice([0.1, 0.5, 0.9, 1.0]))
        r["flags"] = f
    table.rows = list(rows)
    table.dashes = dashes

    n_areas = n_ways * 2 + n_mps * 2
    ent = np.concatenate([rng.integers(0, n_ways, n_ways * 2), rng.integers(0, n_mps, n_mps * 2) | OSMR_AREA_MULTIPOLYGON]).astype(np.uint32)
    rng.shuffle(ent)
    areas = np.zeros(n_areas, dtype=AREA_DTYPE)
    areas["entity"] = ent
    areas["style"] = rng.integers(0, n_styles, n_areas)
    tiles = np.array([(zoom, TX >> shift, TY >> shift, scale)], dtype=TILE_DTYPE)
    return image, table, tiles, np.array([0, n_areas], dtype=np.uint32), areas

"""-m gpu: random geometry / style fuzzing of the CUDA path against the oracle (bit-exact RGB).

Covers what the fixture tiles cannot: every line cap x dash combination incl. casing dashes, opacities, widths from
0.2 to 16 px, pattern fills (and the failed-icon skip), degenerate and far-out-of-tile geometry, axis-aligned and
diagonal ties, multipolygons with several rings, @2x and zooms below 18 (half-pixel rounding)."""
import numpy as np
import pytest

import oracle
from synthgeom import random_scene

pytestmark = pytest.mark.gpu


def _run(ctx_cls, scene, canvas=(241, 238, 232), caps=True):
    image, table, tiles, begins, areas = scene
    ctx = ctx_cls(0)
    try:
        ctx.set_geodata(image)
        ctx.set_table(table)
        got = ctx.draw_tiles(tiles, begins, areas, canvas, caps)
    finally:
        ctx.close()
    want = np.stack(oracle.draw_tiles(image, table, tiles, begins, areas, canvas, caps))
    return got, want


@pytest.mark.parametrize("seed", range(24))
def test_random_scene_matches_oracle(seed):
    from osm_renderer_b200.drawer import GpuContext

    got, want = _run(GpuContext, random_scene(seed))
    bad = (got != want).any(axis=-1)
    assert bad.sum() == 0, f"seed {seed}: {bad.sum()} px differ, first at {np.argwhere(bad)[:5].tolist()}"


@pytest.mark.parametrize("seed,scale,zoom", [(100, 2, 18), (101, 2, 18), (102, 1, 17), (103, 1, 16), (104, 2, 15), (105, 4, 18)])
def test_random_scene_scales_and_zooms(seed, scale, zoom):
    from osm_renderer_b200.drawer import GpuContext

    got, want = _run(GpuContext, random_scene(seed, scale=scale, zoom=zoom))
    assert (got != want).sum() == 0


def test_no_caps_for_dashes_flag_and_no_canvas():
    from osm_renderer_b200.drawer import GpuContext

    got, want = _run(GpuContext, random_scene(7), canvas=None, caps=False)
    assert (got != want).sum() == 0


def test_dense_polygon_rows_streaming_path():
    """Many spans per row (> the 128 ranked in shared memory): the counting form of the even-odd rule."""
    from osm_renderer_b200.drawer import GpuContext

    scene = random_scene(9, n_ways=4, n_mps=3)
    # one multipolygon with 90 random rings of 3..9 points -> hundreds of edges crossing most rows
    image, table, tiles, begins, areas = random_scene(11, n_ways=2, n_mps=1)
    from osm_renderer_b200.upstream import synth
    from synthgeom import TX, TY

    rng = np.random.default_rng(3)
    b = synth._Builder()
    ts = b.tagset({"k": "v"})
    pids = []
    for _ in range(90):
        k = int(rng.integers(3, 10))
        ids = b.add_nodes(rng.integers(-50, 300, k) + TX * 256, rng.integers(-50, 300, k) + TY * 256)
        pids.append(len(b.polys))
        b.polys.append(np.concatenate([ids, ids[:1]]))
    b.mps.append((pids, ts))
    ids = b.add_nodes([TX * 256, TX * 256 + 10], [TY * 256, TY * 256 + 10])
    b.way_nodes.append(ids)
    b.way_tags.append(ts)
    image = synth._serialise(b)
    areas = np.zeros(1, dtype=areas.dtype)
    areas["entity"] = 0x80000000
    areas["style"] = next(i for i, r in enumerate(table.rows) if r["flags"] & 2)  # a fill-colour style
    got, want = _run(GpuContext, (image, table, tiles, np.array([0, 1], dtype=np.uint32), areas))
    assert (got != want).sum() == 0
    assert (want != np.array([241, 238, 232], dtype=np.uint8)).any()

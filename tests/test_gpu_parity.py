"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle and the reference's goldens.

Bar: bit-exact RGB (integer/byte output; the f64 arithmetic is reproduced operation by operation).
"""
import numpy as np
import pytest

import oracle
from conftest import CONFIG_NAMES

pytestmark = pytest.mark.gpu


def _oracle(fx, name, sel=None, **kw):
    tiles, begins, areas = fx.batches[name]
    if sel is not None:
        parts = [areas[begins[i] : begins[i + 1]] for i in sel]
        tiles = tiles[sel]
        begins = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
        areas = np.concatenate(parts)
    imgs = oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes, n_threads=8, **kw)
    return tiles, begins, areas, np.stack(imgs)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_fixture_tiles_match_oracle_exactly(fx, gpu_ctx, name):
    tiles, begins, areas, want = _oracle(fx, name)
    got = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes)
    diff = (got != want).any(axis=-1)
    assert diff.sum() == 0, f"{name}: {diff.sum()} differing pixels, max |d| = {np.abs(got.astype(int) - want.astype(int)).max()}"


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_fixture_tiles_match_reference_golden_outside_labels(fx, gpu_ctx, name):
    tiles, begins, areas = fx.batches[name]
    golden, label_mask = fx.golden(name)
    got = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes)
    diff = (got != golden).any(axis=-1) & ~label_mask
    assert diff.sum() == 0


def test_rgba_output_equals_rgb_plus_opaque_alpha(fx, gpu_ctx):
    tiles, begins, areas = fx.batches["16"]
    rgb = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes)
    rgba = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes, rgba=True)
    assert (rgba[..., :3] == rgb).all() and (rgba[..., 3] == 255).all()


@pytest.mark.parametrize("cap", [0, 1, 3])
def test_streaming_even_odd_path_equals_ranked_path(fx, gpu_ctx, cap):
    """fill_cap forces fill_rows_kernel onto its order-free counting form; the result must not change."""
    tiles, begins, areas, want = _oracle(fx, "17")
    try:
        gpu_ctx.debug_set("fill_cap", cap)
        got = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes)
    finally:
        gpu_ctx.debug_set("fill_cap", 128)
    assert (got == want).all()


@pytest.mark.parametrize("tile", [(14, 9903, 5121, 1), (18, 158460, 81950, 1), (18, 158460, 81950, 2), (10, 619, 320, 1), (0, 0, 0, 1)])
def test_projection_equals_reference_arithmetic(fx, gpu_ctx, tile):
    """a1: Point::from_node of all 19,184 fixture nodes.  The device tan/log are not glibc's, so a differing last
    bit before rounding is possible in principle; the acceptance bar is zero differing integer pixels here."""
    want = oracle.project_nodes(fx.bin, tile)
    got = gpu_ctx.project_nodes(tile)
    assert (got != want).sum() == 0


def test_single_tile_call_equals_batched_call(fx, gpu_ctx):
    tiles, begins, areas = fx.batches["17"]
    batched = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes)
    for i in (0, 7, 19):
        a = areas[begins[i] : begins[i + 1]]
        one = gpu_ctx.draw_tiles(tiles[i : i + 1], np.array([0, len(a)], dtype=np.uint32), a, fx.canvas_rgb, fx.use_caps_for_dashes)
        assert (one[0] == batched[i]).all()


def test_no_canvas_colour_and_no_caps_flag(fx, gpu_ctx):
    tiles, begins, areas = fx.batches["16"]
    imgs = oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, None, False, n_threads=8)
    got = gpu_ctx.draw_tiles(tiles, begins, areas, None, False)
    assert (got == np.stack(imgs)).all()


def test_bad_indices_are_reported_not_crashed(fx, gpu_ctx):
    from osm_renderer_b200._lib import OsmrError

    tiles, begins, areas = fx.batches["14"]
    bad = areas.copy()
    bad["style"][5] = 10_000_000
    with pytest.raises(OsmrError):
        gpu_ctx.draw_tiles(tiles, begins, bad, fx.canvas_rgb, True)
    bad = areas.copy()
    bad["entity"][7] = 0x7FFFFFF0
    with pytest.raises(OsmrError):
        gpu_ctx.draw_tiles(tiles, begins, bad, fx.canvas_rgb, True)
    # the context stays usable
    ok = gpu_ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
    assert ok.shape == (4, 256, 256, 3)

"""CPU (-m "not gpu"): the oracle against every golden vector the reference's own tests hold for this path.

* tests/rendered/*_expected.png through tests/golden/golden_*.npz: exact RGB outside the label pass + test grid
  (reference tests/test_rendering.rs:46-51,147-176)
* doc-test known answers of src/tile.rs:23-29,77-87
* the derived compositor known answer of SURVEY.md 8(c)(5)
"""
import numpy as np
import pytest

import oracle
from conftest import CONFIG_NAMES


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_oracle_equals_reference_golden_outside_labels(fx, name):
    tiles, begins, areas = fx.batches[name]
    golden, label_mask = fx.golden(name)
    imgs = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles, begins, areas, fx.canvas_rgb, fx.use_caps_for_dashes, n_threads=8))
    diff = (imgs != golden).any(axis=-1) & ~label_mask
    assert diff.sum() == 0
    # the mask (labels + red grid) must stay a small minority, otherwise the pin above means nothing
    assert label_mask.mean() < 0.04


def test_label_mask_is_small_blobs_not_geometry(fx):
    """Every masked pixel outside the grid lines lies within a few pixels of other masked pixels (glyph/icon blobs);
    a systematic area-path bug would instead mask long edges.  Checked as: per tile < 9 % masked."""
    for name in CONFIG_NAMES:
        _, label_mask = fx.golden(name)
        per_tile = label_mask.reshape(label_mask.shape[0], -1).mean(axis=1)
        assert per_tile.max() < 0.09, (name, per_tile.max())


def test_coords_to_xy_doc_vectors():
    # src/tile.rs:77-87 (floor of coords_to_xy)
    for lat, lon, zoom, want in [
        (55.747764, 37.437745, 5, (4947, 2561)),
        (55.747764, 37.437745, 18, (40533333, 20981065)),
        (40.1222, 20.6852, 0, (142, 96)),
        (-35.306536, 149.126545, 10, (239662, 158582)),
    ]:
        x, y = oracle.coords_to_xy(lat, lon, zoom)
        assert (int(x), int(y)) == want


def test_coords_to_max_zoom_tile_doc_vectors():
    # src/tile.rs:23-29
    from osm_renderer_b200.upstream.geodata import coords_to_max_zoom_tile, tile_to_max_zoom_tile_range

    assert coords_to_max_zoom_tile(55.747764, 37.437745) == (158333, 81957)
    assert coords_to_max_zoom_tile(40.1222, 20.6852) == (146134, 99125)
    assert coords_to_max_zoom_tile(-35.306536, 149.126545) == (239662, 158582)
    # src/tile.rs:41-62
    assert tile_to_max_zoom_tile_range(0, 0, 0) == (0, 262143, 0, 262143)
    assert tile_to_max_zoom_tile_range(15, 19805, 10244) == (158440, 158447, 81952, 81959)
    assert tile_to_max_zoom_tile_range(18, 239662, 158582) == (239662, 239662, 158582, 158582)


def test_building_fill_known_answer(fx):
    """#bca9a9 at fill-opacity 0.9 over canvas #f1eee8 -> (193,175,175), a top-4 colour of the z18 golden."""
    tiles, begins, areas = fx.batches["18"]
    imgs = np.stack(oracle.draw_tiles(fx.bin, fx.table, tiles[:8], begins[:9], areas[: begins[8]], fx.canvas_rgb, fx.use_caps_for_dashes, n_threads=8))
    hits = (imgs == np.array([193, 175, 175], dtype=np.uint8)).all(axis=-1).sum()
    assert hits > 10_000


def test_generation_limit_and_empty_batch(fx):
    tiles, begins, areas = fx.batches["14"]
    img0 = oracle.draw_tiles(fx.bin, fx.table, tiles[:1], begins[:2], areas[: begins[1]], fx.canvas_rgb, True, gen_limit=0)[0]
    assert (img0 == np.array(fx.canvas_rgb, dtype=np.uint8)).all()
    # a tile without any styled area is the canvas colour; without canvas colour it is black (tile_pixels.rs:231-236)
    empty = oracle.draw_tiles(fx.bin, fx.table, tiles[:1], np.array([0, 0], dtype=np.uint32), areas[:0], None, True)[0]
    assert (empty == 0).all()

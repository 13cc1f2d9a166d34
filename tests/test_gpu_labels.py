"""-m gpu: the whole Drawer::draw_to_pixels on the GPU path -- area passes AND label pass (icons, text, greedy
collisions) -- against the REFERENCE'S OWN golden renders, every pixel, no oracle in between
(tests/test_rendering.rs:46-51,147-176: exact RGB; only the red test grid of :109-114 is excluded)."""
import numpy as np
import pytest

from conftest import CONFIG_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def label_ctx(fx):
    from osm_renderer_b200.drawer import GpuContext

    ltable, font, per = fx.labels()
    ctx = GpuContext(0)
    ctx.set_geodata(fx.bin)
    ctx.set_table(fx.table)
    ctx.set_font(font)
    ctx.set_label_table(ltable)
    yield ctx, per
    ctx.close()


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_full_tiles_equal_reference_golden_on_every_pixel(fx, label_ctx, name):
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches[name]
    label_begin, labels = per[name]
    got = ctx.draw_tiles_labeled(tiles, begins, areas, label_begin, labels, fx.canvas_rgb, fx.use_caps_for_dashes)
    golden, _ = fx.golden(name)
    d = golden.shape[1]
    grid = np.zeros((d, d), dtype=bool)
    grid[0, :] = True
    grid[:, d - 1] = True
    diff = (got != golden).any(axis=-1) & ~grid[None]
    assert diff.sum() == 0, f"{name}: {diff.sum()} pixels differ from the reference golden in tiles {sorted(set(np.argwhere(diff)[:, 0].tolist()))[:10]}"


def test_label_pass_without_labels_equals_area_passes(fx, label_ctx):
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches["16"]
    plain = ctx.draw_tiles(tiles, begins, areas, fx.canvas_rgb, True)
    empty = ctx.draw_tiles_labeled(tiles, begins, areas, np.zeros(len(tiles) + 1, dtype=np.uint32), per["16"][1][:0], fx.canvas_rgb, True)
    assert (plain == empty).all()


def test_label_layout_is_independent_of_host_threading_and_batching(fx, label_ctx):
    """The layout is handed to host threads in runs of labels and tiles are batched: neither may change a pixel."""
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches["17"]
    lb, labels = per["17"]
    ref = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
    try:
        ctx.debug_set("label_threads", 1)
        serial = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
    finally:
        ctx.debug_set("label_threads", 32)
    assert (serial == ref).all()
    for t in (0, len(tiles) // 2, len(tiles) - 1):  # one tile at a time
        a0, a1 = int(begins[t]), int(begins[t + 1])
        l0, l1 = int(lb[t]), int(lb[t + 1])
        one = ctx.draw_tiles_labeled(tiles[t:t + 1], np.array([0, a1 - a0], dtype=np.uint32), areas[a0:a1],
                                     np.array([0, l1 - l0], dtype=np.uint32), labels[l0:l1], fx.canvas_rgb, True)
        assert (one[0] == ref[t]).all()


def test_label_api_rejects_bad_input(fx, label_ctx):
    ctx, per = label_ctx
    from osm_renderer_b200._lib import OsmrError

    tiles, begins, areas = fx.batches["18"]
    lb, labels = per["18"]
    bad = labels.copy()
    bad["style"][0] = 1 << 30
    with pytest.raises(OsmrError):
        ctx.draw_tiles_labeled(tiles, begins, areas, lb, bad, fx.canvas_rgb, True)
    bad = labels.copy()
    bad["entity"][0] = 0x3FFFFFFF  # way index far beyond the dataset
    with pytest.raises(OsmrError):
        ctx.draw_tiles_labeled(tiles, begins, areas, lb, bad, fx.canvas_rgb, True)
    # the context stays usable
    ok = ctx.draw_tiles_labeled(tiles[:2], begins[:3], areas[: int(begins[2])], lb[:3], labels[: int(lb[2])], fx.canvas_rgb, True)
    assert ok.shape[0] == 2


def test_truncated_and_corrupt_fonts_never_crash(fx):
    """osmr.h: no entry point aborts.  A truncated or bit-flipped TrueType file must either be refused or load with missing
    glyphs -- every offset read from the file is bounds-checked."""
    from osm_renderer_b200._lib import OsmrError
    from osm_renderer_b200.drawer import GpuContext

    ltable, font, per = fx.labels()
    ctx = GpuContext(0)
    ctx.set_geodata(fx.bin)
    ctx.set_table(fx.table)
    ctx.set_label_table(ltable)
    tiles, begins, areas = fx.batches["17"]
    lb, labels = per["17"]
    sel = slice(0, 2)
    rng = np.random.default_rng(5)
    variants = [font[:64], font[:300], font[: len(font) // 3], font[: len(font) - 1000]]
    for _ in range(6):
        b = bytearray(font)
        for pos in rng.integers(0, 4000, size=40):  # table directory, head, hhea, cmap headers live in the first kilobytes
            b[int(pos)] ^= int(rng.integers(1, 256))
        variants.append(bytes(b))
    loaded = 0
    for v in variants:
        try:
            ctx.set_font(v)
        except OsmrError:
            continue
        loaded += 1
        try:
            ctx.draw_tiles_labeled(tiles[sel], begins[:3], areas[: begins[2]], lb[:3], labels[: lb[2]], fx.canvas_rgb, True)
        except OsmrError:
            pass
    ctx.set_font(font)  # and the context is still usable afterwards
    good = ctx.draw_tiles_labeled(tiles[sel], begins[:3], areas[: begins[2]], lb[:3], labels[: lb[2]], fx.canvas_rgb, True)
    golden, _ = fx.golden("17")
    diff = (good != golden[sel]).any(axis=-1)
    diff[:, 0, :] = False
    diff[:, :, 255] = False
    assert diff.sum() == 0
    ctx.close()


@pytest.mark.parametrize("name", ["16", "18_2x"])
def test_device_layout_equals_host_layout(fx, label_ctx, name):
    """The label layout runs on the device (osmr_labels_dev.cuh); the host layout of round 1 is its fallback.  Same pixels."""
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches[name]
    lb, labels = per[name]
    dev = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
    st = ctx.stats()
    assert st["label_path"] == 1 and st["n_labels_active"] > 0 and st["n_labels_polylabel"] > 0
    try:
        ctx.debug_set("label_host", 1)
        host = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
        assert ctx.stats()["label_path"] == 2
    finally:
        ctx.debug_set("label_host", 0)
    assert (dev == host).all()


@pytest.mark.parametrize("n_chunks", [2, 5])
def test_chunked_label_pass_equals_golden(fx, label_ctx, n_chunks):
    """A host-output call runs the label pass chunk by chunk (the draw's schedule: a draw chunk waits only for the label chunk
    under it, scratch reused chunk after chunk); debug key "label_chunks" forces that on a small batch."""
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches["17"]
    lb, labels = per["17"]
    try:
        ctx.debug_set("label_chunks", n_chunks)
        got = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
        assert ctx.stats()["label_path"] == 1
    finally:
        ctx.debug_set("label_chunks", 0)
    golden, _ = fx.golden("17")
    diff = (got != golden).any(axis=-1)
    diff[:, 0, :] = False
    diff[:, :, 255] = False
    assert diff.sum() == 0


def test_curves_beyond_the_leaf_code_record_are_flattened_again(fx, label_ctx):
    """label_curve_expand_kernel writes a curve's segments from its recorded leaf codes; a curve with more leaves than the record
    holds is flattened again (flatness tests included) by one lane.  Debug key "curve_leaf_cap" shrinks the record so that
    nearly every curve takes that path."""
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches["16"]
    lb, labels = per["16"]
    try:
        ctx.debug_set("curve_leaf_cap", 6)
        got = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
        assert ctx.stats()["label_path"] == 1
    finally:
        ctx.debug_set("curve_leaf_cap", 0)
    golden, _ = fx.golden("16")
    diff = (got != golden).any(axis=-1)
    diff[:, 0, :] = False
    diff[:, :, 255] = False
    assert diff.sum() == 0


def test_label_cull_changes_no_pixel(fx, label_ctx):
    """label_cull_kernel drops the labels of the 3x3 neighbourhood that cannot reach the tile, directly or through a chain of
    collisions, before their outlines are flattened: same tiles with and without it, far fewer outline segments with it."""
    ctx, per = label_ctx
    for name in ("16", "18_2x"):
        tiles, begins, areas = fx.batches[name]
        lb, labels = per[name]
        culled = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
        st1 = ctx.stats()
        try:
            ctx.debug_set("label_cull", 0)
            full = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
            st0 = ctx.stats()
        finally:
            ctx.debug_set("label_cull", 1)
        assert (culled == full).all(), name
        assert st1["label_path"] == 1 and st0["label_path"] == 1
        assert st1["n_labels_active"] == st0["n_labels_active"] and 0 < st1["n_label_segments"] < st0["n_label_segments"] / 2


def test_label_serial_key_changes_no_pixel(fx, label_ctx):
    """debug key label_serial (bench.py's `label_kernels_alone` leg): the resident labelled draw waits for its label pass before
    the area passes start -- a scheduling change only"""
    ctx, per = label_ctx
    tiles, begins, areas = fx.batches["15"]
    lb, labels = per["15"]
    want = ctx.draw_tiles_labeled(tiles, begins, areas, lb, labels, fx.canvas_rgb, True)
    ctx.batch_upload_labeled(tiles, begins, areas, lb, labels)
    out = np.empty_like(want)
    try:
        ctx.debug_set("label_serial", 1)
        ms = ctx.batch_draw_labeled(fx.canvas_rgb, True, out=out)
    finally:
        ctx.debug_set("label_serial", 0)
    assert ms > 0 and (out == want).all() and ctx.stats()["ms_label_cover"] > 0

"""-m gpu: the HEADLINE configuration (BASELINE.json configs[1], SURVEY.md 8d C2: 1024 tiles z14 over the seeded synthetic
metro) against the CPU oracle -- 64 tiles spread over the real bench batch, bit for bit -- for every reference-facing entry
point of that batch: osmr_draw_tiles, osmr_draw_tiles_auto (f3), osmr_draw_tiles_png / _auto_png (f4, decoded),
osmr_draw_tiles_labeled (the whole draw_to_pixels); and the second stylesheet of C2 (mapcss/osmosnimki-minimal.mapcss,
reference README.md:23-25)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_SAMPLE = 64


@pytest.fixture(scope="module")
def c2():
    import bench
    import oracle

    w = bench.build_workload("C2", "mapnik", labels=True, max_tiles=N_SAMPLE)
    sel = np.arange(len(w["tiles"]))
    tiles, begins, areas, lb, labels = bench.sub_batch(w, sel, True)
    want = np.stack(oracle.draw_tiles(w["bin"], w["table"], tiles, begins, areas, w["canvas"], w["caps"], n_threads=8))
    return w, (tiles, begins, areas, lb, labels), want


@pytest.fixture()
def ctx(c2):
    from osm_renderer_b200.drawer import GpuContext

    w = c2[0]
    c = GpuContext(0)
    c.set_geodata(w["bin"])
    c.set_table(w["table"])
    yield c
    c.close()


def test_c2_batch_equals_oracle(c2, ctx):
    w, (tiles, begins, areas, lb, labels), want = c2
    assert len(tiles) == N_SAMPLE
    got = ctx.draw_tiles(tiles, begins, areas, w["canvas"], w["caps"])
    assert (got == want).all(), int((got != want).any(axis=-1).sum())
    st = ctx.stats()
    assert st["n_visible_ops"] > 1000 * len(tiles) // 4  # the metro is dense: this is not an empty-tile test


def test_c2_auto_and_png_equal_oracle(c2, ctx):
    """f3 and f4 against the ORACLE directly (not against osmr_draw_tiles)."""
    from test_gpu_png import decode_png

    from osm_renderer_b200.upstream import pipeline

    w, (tiles, begins, areas, lb, labels), want = c2
    wc, mc, cb, cs = pipeline.zoom_class_tables(w["builder"], 14)
    ctx.set_zoom_styles(14, wc, mc, cb, cs)
    got = ctx.draw_tiles_auto(tiles, w["canvas"], w["caps"])
    assert (got == want).all()
    sub = slice(0, 8)
    files = ctx.draw_tiles_png(tiles[sub], begins[:9], areas[: begins[8]], w["canvas"], w["caps"])
    for f, im in zip(files, want[sub]):
        assert (decode_png(f) == im).all()
    files = ctx.draw_tiles_auto_png(tiles[8:16], w["canvas"], w["caps"])
    assert len(files) == 8
    for f, im in zip(files, want[8:16]):
        assert (decode_png(f) == im).all()
    st = ctx.stats()
    assert st["ms_auto"] > 0 and st["ms_png"] > 0


def test_c2_labeled_equals_oracle(c2, ctx):
    """The whole draw_to_pixels (area passes + label pass: street names along the ways) on bench tiles."""
    import bench
    import oracle

    w, (tiles, begins, areas, lb, labels), want_plain = c2
    n = 16
    sel = np.arange(n)
    t, b, a, lbn, ln = bench.sub_batch(w, sel, True)
    l48, texts = bench.oracle_labels(w, lbn, ln)
    want = np.stack(oracle.draw_tiles_with_labels(w["bin"], w["table"], t, b, a, w["canvas"], w["caps"], w["font"], w["ltable"].icons, lbn, l48,
                                                  texts, n_threads=8))
    ctx.set_font(w["font"])
    ctx.set_label_table(w["ltable"])
    got = ctx.draw_tiles_labeled(t, b, a, lbn, ln, w["canvas"], w["caps"])
    assert (got == want).all(), int((got != want).any(axis=-1).sum())
    assert (want != want_plain[:n]).any(), "the sample must contain drawn labels"


def test_c2_second_stylesheet_equals_oracle(ctx):
    """mapcss/osmosnimki-minimal.mapcss (the production stylesheet of the reference's README) on bench tiles."""
    import bench
    import oracle

    w = bench.build_workload("C2", "osmosnimki", labels=False, max_tiles=16)
    tiles, begins, areas = bench.sub_batch(w, np.arange(len(w["tiles"])))
    want = np.stack(oracle.draw_tiles(w["bin"], w["table"], tiles, begins, areas, w["canvas"], w["caps"], n_threads=8))
    ctx.set_table(w["table"])
    got = ctx.draw_tiles(tiles, begins, areas, w["canvas"], w["caps"])
    assert (got == want).all()
    assert len(np.unique(want.reshape(-1, 3), axis=0)) > 3  # something was drawn


def test_c2_labeled_big_call_equals_small_calls(ctx):
    """192 labelled tiles in ONE call take the pipelined route (draw chunks, styled-area tail and label lists on their own copy
    streams, tiles copied back chunk by chunk); the same tiles in calls of 48 take the simple one.  Same pixels."""
    import bench

    w = bench.build_workload("C2", "mapnik", labels=True, max_tiles=192)
    n = len(w["tiles"])
    assert n >= 128
    ctx.set_table(w["table"])
    ctx.set_font(w["font"])
    ctx.set_label_table(w["ltable"])
    t, b, a, lb, ln = bench.sub_batch(w, np.arange(n), True)
    big = ctx.draw_tiles_labeled(t, b, a, lb, ln, w["canvas"], w["caps"])
    assert ctx.stats()["label_path"] == 1
    for first in range(0, n, 48):
        sel = np.arange(first, min(n, first + 48))
        ts, bs, as_, lbs, lns = bench.sub_batch(w, sel, True)
        small = ctx.draw_tiles_labeled(ts, bs, as_, lbs, lns, w["canvas"], w["caps"])
        assert (small == big[sel]).all(), first

"""osmr_draw_tiles_png (filter + deflate + checksums on the device) must produce valid PNG files
(signature, IHDR, one IDAT, IEND; CRC-32 of every chunk, Adler-32 of the zlib stream) that decode to exactly the RGB tiles
of osmr_draw_tiles -- what reference src/draw/png_writer.rs:4-21 guarantees for Drawer::draw_tile (drawer.rs:40-58)."""
import struct
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def decode_png(data: bytes):
    """strict decoder: checks every CRC, the chunk sequence and the filters; returns uint8 [h, w, 3]"""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        (n,) = struct.unpack(">I", data[pos : pos + 4])
        typ = data[pos + 4 : pos + 8]
        body = data[pos + 8 : pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n : pos + 12 + n])
        assert zlib.crc32(typ + body) == crc, f"bad CRC in {typ}"
        chunks.append((typ, body))
        pos += 12 + n
    assert pos == len(data)
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]  # png_writer.rs writes one IDAT
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, inter) == (8, 2, 0, 0, 0)
    raw = zlib.decompress(chunks[1][1])  # verifies the Adler-32
    stride = 3 * w
    assert len(raw) == h * (stride + 1)
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(h, stride + 1)
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        f = int(rows[y, 0])
        line = rows[y, 1:].astype(np.int32)
        cur = np.zeros(stride, dtype=np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        elif f in (1, 4):
            for x in range(stride):  # Sub / Paeth are sequential along the row
                a = cur[x - 3] if x >= 3 else 0
                if f == 1:
                    pred = a
                else:
                    b, c = prev[x], (prev[x - 3] if x >= 3 else 0)
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + pred) & 255
        else:
            raise AssertionError(f"unexpected filter {f}")
        out[y] = cur
        prev = cur
    return out.reshape(h, w, 3)


@pytest.mark.parametrize("name,sel", [("17", [0, 7, 12]), ("14", [0]), ("18_2x", [3])])
def test_png_files_decode_to_the_rgb_tiles(fx, gpu_ctx, name, sel):
    tiles, begins, areas = fx.batches[name]
    parts = [areas[begins[i] : begins[i + 1]] for i in sel]
    b = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
    a = np.concatenate(parts)
    want = gpu_ctx.draw_tiles(tiles[sel], b, a, fx.canvas_rgb, fx.use_caps_for_dashes)
    files = gpu_ctx.draw_tiles_png(tiles[sel], b, a, fx.canvas_rgb, fx.use_caps_for_dashes)
    assert len(files) == len(sel)
    raw_bytes = want[0].size
    for i, f in enumerate(files):
        got = decode_png(f)
        assert got.shape == want[i].shape and (got == want[i]).all()
        assert len(f) < raw_bytes // 2  # flat regions become runs; literals cost 8-9 bits with the fixed Huffman code
    assert gpu_ctx.stats()["ms_png"] > 0


def test_png_of_noise_fits_the_bound(gpu_ctx, fx):
    """Worst case for the encoder: a pattern fill of random texels (every byte a literal)."""
    import ctypes as C

    from osm_renderer_b200.wire import TILE_DTYPE

    n = gpu_ctx.L.osmr_png_bound(1)
    assert n > 256 * (3 * 256 + 1)  # room for an incompressible tile
    assert gpu_ctx.L.osmr_png_bound(0) == 0 and gpu_ctx.L.osmr_png_bound(9) == 0
    tiles, begins, areas = fx.batches["18"]
    files = gpu_ctx.draw_tiles_png(tiles[:2], begins[:3], areas[: begins[2]], None, True)  # no canvas colour: black background
    for f in files:
        assert len(f) <= n
        decode_png(f)


def _edge_case_images():
    """256x256 images that hit the encoder's corner cases: runs of every length around the 258 limit (the tail of a run
    shorter than 3 must become literals), runs crossing row ends, incompressible noise, a flat image, gradients that make
    each of the four filters win somewhere."""
    rng = np.random.default_rng(7)
    imgs = []
    a = np.zeros((256, 256, 3), dtype=np.uint8)  # row y: y+1 zero bytes at the start ... then a non-repeating ramp
    ramp = (np.arange(768) * 7 % 251 + 1).astype(np.uint8)
    for y in range(256):
        row = ramp.copy()
        n = 250 + y  # zero runs of 250 .. 505 bytes (after the None/Sub filter these stay runs)
        row[: min(n, 768)] = 0
        a[y] = row.reshape(256, 3)
    imgs.append(a)
    imgs.append(rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8))  # noise: every byte a literal
    imgs.append(np.full((256, 256, 3), 200, dtype=np.uint8))  # flat
    g = np.zeros((256, 256, 3), dtype=np.uint8)  # horizontal + vertical gradients, a checker in one band
    g[..., 0] = np.arange(256, dtype=np.uint8)[None, :]
    g[..., 1] = np.arange(256, dtype=np.uint8)[:, None]
    g[64:96, :, 2] = ((np.arange(256)[None, :] // 3 + np.arange(32)[:, None]) % 2 * 255).astype(np.uint8)
    imgs.append(g)
    b = rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8)  # noise rows alternating with copies of the row above (Up wins)
    b[1::2] = b[0::2]
    imgs.append(b)
    return np.stack(imgs)


def test_rgb_to_png_corner_cases(gpu_ctx):
    imgs = _edge_case_images()
    files = gpu_ctx.rgb_to_png(imgs)
    assert len(files) == len(imgs)
    bound = gpu_ctx.L.osmr_png_bound(1)
    for f, im in zip(files, imgs):
        assert len(f) <= bound
        assert (decode_png(f) == im).all()
    assert len(files[2]) < 2500  # a flat image is a handful of maximal runs per row
    assert len(files[4]) < len(files[1]) * 0.6  # the copied rows cost nothing with the Up filter

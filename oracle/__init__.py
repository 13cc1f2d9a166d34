"""ctypes binding of the CPU oracle (oracle/osmr_oracle.cpp).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from the osm_renderer_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libosmr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "osmr_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.osmr_oracle_draw_tiles.restype = C.c_int
        L.osmr_oracle_draw_tiles.argtypes = [
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
            C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
        ]
        L.osmr_oracle_project_nodes.restype = C.c_int
        L.osmr_oracle_project_nodes.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.osmr_oracle_coords_to_xy.restype = None
        L.osmr_oracle_coords_to_xy.argtypes = [C.c_double, C.c_double, C.c_uint32, C.c_void_p]
        L.osmr_oracle_fill_edge_rows.restype = None
        L.osmr_oracle_fill_edge_rows.argtypes = [C.c_int32] * 6 + [C.c_void_p]
        _lib = L
    return _lib


def draw_tiles(bin_image: bytes, table, tiles: np.ndarray, area_begin: np.ndarray, areas: np.ndarray,
               canvas_rgb, use_caps_for_dashes: bool = True, gen_limit: int = 0xFFFFFFFF, n_threads: int = 1):
    """Returns a list of uint8 arrays [D, D, 3], one per tile (reference TileRenderedPixels.triples)."""
    from osm_renderer_b200.wire import OSMR_DRAW_HAS_CANVAS_COLOR, OSMR_DRAW_USE_CAPS_FOR_DASHES

    L = lib()
    styles = table.styles_array()
    dashes = table.dashes_array()
    icons, _keep = table.icon_structs()
    tiles = np.ascontiguousarray(tiles)
    area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
    areas = np.ascontiguousarray(areas)
    sizes = [(256 * int(s)) ** 2 * 3 for s in tiles["scale"]]
    out = np.zeros(int(sum(sizes)), dtype=np.uint8)
    flags = (OSMR_DRAW_USE_CAPS_FOR_DASHES if use_caps_for_dashes else 0)
    canvas = np.zeros(3, dtype=np.uint8)
    if canvas_rgb is not None:
        flags |= OSMR_DRAW_HAS_CANVAS_COLOR
        canvas[:] = canvas_rgb
    buf = np.frombuffer(bin_image, dtype=np.uint8)
    rc = L.osmr_oracle_draw_tiles(
        buf.ctypes.data, len(bin_image), styles.ctypes.data, len(styles), dashes.ctypes.data, len(dashes),
        C.addressof(icons), len(table.icons), tiles.ctypes.data, len(tiles), area_begin.ctypes.data,
        areas.ctypes.data, canvas.ctypes.data, flags, gen_limit, n_threads, out.ctypes.data,
    )
    if rc != 0:
        raise RuntimeError(f"osmr_oracle_draw_tiles failed: {rc}")
    res = []
    off = 0
    for s, n in zip(tiles["scale"], sizes):
        d = 256 * int(s)
        res.append(out[off : off + n].reshape(d, d, 3))
        off += n
    return res


def project_nodes(bin_image: bytes, tile) -> np.ndarray:
    from osm_renderer_b200.wire import TILE_DTYPE

    L = lib()
    t = np.array([tuple(tile)], dtype=TILE_DTYPE)
    n = int(np.frombuffer(bin_image, dtype="<u4", count=1)[0])
    out = np.zeros((n, 2), dtype=np.int32)
    buf = np.frombuffer(bin_image, dtype=np.uint8)
    rc = L.osmr_oracle_project_nodes(buf.ctypes.data, len(bin_image), t.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"osmr_oracle_project_nodes failed: {rc}")
    return out


def coords_to_xy(lat: float, lon: float, zoom: int):
    out = np.zeros(2, dtype=np.float64)
    lib().osmr_oracle_coords_to_xy(lat, lon, zoom, out.ctypes.data)
    return float(out[0]), float(out[1])


def fill_edge_rows(x1, y1, x2, y2, min_y, max_y) -> np.ndarray:
    out = np.zeros((max_y - min_y + 1, 3), dtype=np.int32)
    lib().osmr_oracle_fill_edge_rows(x1, y1, x2, y2, min_y, max_y, out.ctypes.data)
    return out


# ---------------------------------------------------------------------------------------------------------------
# label pass (oracle only)
# ---------------------------------------------------------------------------------------------------------------
LABEL_DTYPE = np.dtype(
    [
        ("kind", "<u4"), ("entity", "<u4"), ("icon", "<i4"), ("has_text_style", "<u4"), ("has_font_size", "<u4"),
        ("has_text", "<u4"), ("text_off", "<u4"), ("text_len", "<u4"), ("text_pos", "<u4"),
        ("text_color", "u1", (3,)), ("has_text_color", "u1"), ("font_size", "<f8"),
    ],
    align=True,
)
assert LABEL_DTYPE.itemsize == 48


def draw_tiles_with_labels(bin_image, table, tiles, area_begin, areas, canvas_rgb, use_caps_for_dashes, font_bytes,
                           label_icons, label_begin, labels, texts: bytes, n_threads: int = 1):
    """Full draw_to_pixels (area passes + label pass).  label_icons: list of (w, h, rgba ndarray)."""
    from osm_renderer_b200.wire import OSMR_DRAW_HAS_CANVAS_COLOR, OSMR_DRAW_USE_CAPS_FOR_DASHES, IconStruct

    L = lib()
    L.osmr_oracle_draw_tiles_labels.restype = C.c_int
    L.osmr_oracle_draw_tiles_labels.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                                 C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                 C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_char_p,
                                                 C.c_int, C.c_void_p]
    styles = table.styles_array()
    dashes = table.dashes_array()
    icons, _keep = table.icon_structs()
    licons = (IconStruct * max(1, len(label_icons)))()
    for i, (w, h, px) in enumerate(label_icons):
        licons[i].width, licons[i].height, licons[i].rgba = w, h, px.ctypes.data
    tiles = np.ascontiguousarray(tiles)
    area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
    areas = np.ascontiguousarray(areas)
    label_begin = np.ascontiguousarray(label_begin, dtype=np.uint32)
    labels = np.ascontiguousarray(labels, dtype=LABEL_DTYPE)
    sizes = [(256 * int(s)) ** 2 * 3 for s in tiles["scale"]]
    out = np.zeros(int(sum(sizes)), dtype=np.uint8)
    flags = OSMR_DRAW_USE_CAPS_FOR_DASHES if use_caps_for_dashes else 0
    canvas = np.zeros(3, dtype=np.uint8)
    if canvas_rgb is not None:
        flags |= OSMR_DRAW_HAS_CANVAS_COLOR
        canvas[:] = canvas_rgb
    buf = np.frombuffer(bin_image, dtype=np.uint8)
    fbuf = np.frombuffer(font_bytes, dtype=np.uint8)
    rc = L.osmr_oracle_draw_tiles_labels(
        buf.ctypes.data, len(bin_image), styles.ctypes.data, len(styles), dashes.ctypes.data, len(dashes), C.addressof(icons),
        len(table.icons), tiles.ctypes.data, len(tiles), area_begin.ctypes.data, areas.ctypes.data, canvas.ctypes.data, flags,
        fbuf.ctypes.data, len(font_bytes), C.addressof(licons), len(label_icons), label_begin.ctypes.data, labels.ctypes.data, texts,
        n_threads, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"osmr_oracle_draw_tiles_labels failed: {rc}")
    res, off = [], 0
    for s, n in zip(tiles["scale"], sizes):
        d = 256 * int(s)
        res.append(out[off : off + n].reshape(d, d, 3))
        off += n
    return res

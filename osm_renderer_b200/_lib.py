"""ctypes loader of libosmr_b200.so (the C ABI of include/osmr.h).

Fails loudly when the CUDA library is missing or cannot be loaded: there is no CPU fallback and nothing in
this package imports the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os

from .wire import StatsStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libosmr_b200.so")

EXPORTS = [
    "osmr_abi_version",
    "osmr_ctx_create",
    "osmr_ctx_destroy",
    "osmr_last_error",
    "osmr_set_geodata",
    "osmr_set_icons",
    "osmr_set_styles",
    "osmr_draw_tiles",
    "osmr_batch_upload",
    "osmr_batch_draw",
    "osmr_batch_output",
    "osmr_get_stats",
    "osmr_project_nodes",
    "osmr_alloc_pinned",
    "osmr_free_pinned",
    "osmr_debug_set",
    "osmr_set_font",
    "osmr_set_label_icons",
    "osmr_set_label_styles",
    "osmr_draw_tiles_labeled",
    "osmr_set_zoom_styles",
    "osmr_draw_tiles_auto",
    "osmr_auto_readback",
    "osmr_png_bound",
    "osmr_draw_tiles_png",
    "osmr_rgb_to_png",
    "osmr_draw_tiles_auto_png",
    "osmr_ctx_create_shared",
    "osmr_batch_upload_labeled",
    "osmr_batch_draw_labeled",
    "osmr_set_zoom_label_styles",
    "osmr_draw_tiles_auto_labeled",
    "osmr_draw_tiles_auto_labeled_png",
    "osmr_auto_readback_labels",
]

_lib = None


class OsmrError(RuntimeError):
    pass


def load():
    """Load (never build) the shared library; raise if it is absent -- run __graft_entry__.build() first."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OsmrError(
            f"{LIB_PATH} is missing: build it with `python -m osm_renderer_b200.build` (needs nvcc). "
            "There is no CPU fallback for the draw path."
        )
    L = C.CDLL(LIB_PATH)
    vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
    L.osmr_abi_version.restype = u32
    L.osmr_abi_version.argtypes = []
    L.osmr_ctx_create.restype = C.c_int
    L.osmr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.osmr_ctx_destroy.restype = None
    L.osmr_ctx_destroy.argtypes = [vp]
    L.osmr_last_error.restype = C.c_char_p
    L.osmr_last_error.argtypes = [vp]
    L.osmr_set_geodata.restype = C.c_int
    L.osmr_set_geodata.argtypes = [vp, vp, sz]
    L.osmr_set_icons.restype = C.c_int
    L.osmr_set_icons.argtypes = [vp, vp, u32]
    L.osmr_set_styles.restype = C.c_int
    L.osmr_set_styles.argtypes = [vp, vp, u32, vp, u32]
    L.osmr_draw_tiles.restype = C.c_int
    L.osmr_draw_tiles.argtypes = [vp, vp, u32, vp, vp, vp, u32, vp]
    L.osmr_batch_upload.restype = C.c_int
    L.osmr_batch_upload.argtypes = [vp, vp, u32, vp, vp]
    L.osmr_batch_draw.restype = C.c_int
    L.osmr_batch_draw.argtypes = [vp, vp, u32, vp, C.POINTER(C.c_float)]
    L.osmr_batch_output.restype = C.c_int
    L.osmr_batch_output.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.osmr_get_stats.restype = C.c_int
    L.osmr_get_stats.argtypes = [vp, C.POINTER(StatsStruct)]
    L.osmr_project_nodes.restype = C.c_int
    L.osmr_project_nodes.argtypes = [vp, vp, vp]
    L.osmr_alloc_pinned.restype = vp
    L.osmr_alloc_pinned.argtypes = [sz]
    L.osmr_free_pinned.restype = None
    L.osmr_free_pinned.argtypes = [vp]
    L.osmr_debug_set.restype = C.c_int
    L.osmr_debug_set.argtypes = [vp, C.c_char_p, C.c_int]
    L.osmr_set_font.restype = C.c_int
    L.osmr_set_font.argtypes = [vp, vp, sz]
    L.osmr_set_label_icons.restype = C.c_int
    L.osmr_set_label_icons.argtypes = [vp, vp, u32]
    L.osmr_set_label_styles.restype = C.c_int
    L.osmr_set_label_styles.argtypes = [vp, vp, u32, C.c_char_p, sz]
    L.osmr_draw_tiles_labeled.restype = C.c_int
    L.osmr_draw_tiles_labeled.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, u32, vp]
    L.osmr_set_zoom_styles.restype = C.c_int
    L.osmr_set_zoom_styles.argtypes = [vp, u32, vp, vp, vp, vp, u32]
    L.osmr_draw_tiles_auto.restype = C.c_int
    L.osmr_draw_tiles_auto.argtypes = [vp, vp, u32, vp, u32, vp]
    L.osmr_auto_readback.restype = C.c_int
    L.osmr_auto_readback.argtypes = [vp, vp, vp, u32]
    L.osmr_png_bound.restype = sz
    L.osmr_png_bound.argtypes = [u32]
    L.osmr_draw_tiles_png.restype = C.c_int
    L.osmr_draw_tiles_png.argtypes = [vp, vp, u32, vp, vp, vp, u32, vp, sz, vp]
    L.osmr_rgb_to_png.restype = C.c_int
    L.osmr_rgb_to_png.argtypes = [vp, vp, u32, u32, vp, sz, vp]
    L.osmr_batch_upload_labeled.restype = C.c_int
    L.osmr_batch_upload_labeled.argtypes = [vp, vp, u32, vp, vp, vp, vp]
    L.osmr_batch_draw_labeled.restype = C.c_int
    L.osmr_batch_draw_labeled.argtypes = [vp, vp, u32, vp, C.POINTER(C.c_float)]
    L.osmr_ctx_create_shared.restype = C.c_int
    L.osmr_ctx_create_shared.argtypes = [vp, C.POINTER(vp)]
    L.osmr_draw_tiles_auto_png.restype = C.c_int
    L.osmr_draw_tiles_auto_png.argtypes = [vp, vp, u32, vp, u32, vp, sz, vp]
    L.osmr_set_zoom_label_styles.restype = C.c_int
    L.osmr_set_zoom_label_styles.argtypes = [vp, u32, vp, vp, vp, vp, vp, u32]
    L.osmr_draw_tiles_auto_labeled.restype = C.c_int
    L.osmr_draw_tiles_auto_labeled.argtypes = [vp, vp, u32, vp, u32, vp]
    L.osmr_draw_tiles_auto_labeled_png.restype = C.c_int
    L.osmr_draw_tiles_auto_labeled_png.argtypes = [vp, vp, u32, vp, u32, vp, sz, vp]
    L.osmr_auto_readback_labels.restype = C.c_int
    L.osmr_auto_readback_labels.argtypes = [vp, vp, vp, u32]
    _lib = L
    return L

"""Wire structures of the C ABI (include/osmr.h) as numpy dtypes + ctypes mirrors, and the host-side
interning of styler output (`Arc<Style>` -> table index, icon name -> icon index).

Shared by the product binding (osm_renderer_b200/drawer.py) and by the test-only oracle binding
(oracle/__init__.py) so both sides receive byte-identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

OSMR_STYLE_COLOR = 1 << 0
OSMR_STYLE_FILL_COLOR = 1 << 1
OSMR_STYLE_FILL_IMAGE = 1 << 2
OSMR_STYLE_CASING_COLOR = 1 << 3
OSMR_STYLE_CASING_WIDTH = 1 << 4
OSMR_STYLE_WIDTH = 1 << 5
OSMR_STYLE_OPACITY = 1 << 6
OSMR_STYLE_FILL_OPACITY = 1 << 7
OSMR_STYLE_DASHES = 1 << 8
OSMR_STYLE_CASING_DASHES = 1 << 9

OSMR_AREA_MULTIPOLYGON = 0x80000000

OSMR_DRAW_USE_CAPS_FOR_DASHES = 1 << 0
OSMR_DRAW_HAS_CANVAS_COLOR = 1 << 1
OSMR_DRAW_OUT_RGBA = 1 << 2
OSMR_DRAW_OUT_DEVICE = 1 << 3

TILE_DTYPE = np.dtype([("zoom", "<u4"), ("x", "<u4"), ("y", "<u4"), ("scale", "<u4")])

STYLE_DTYPE = np.dtype(
    [
        ("flags", "<u4"),
        ("color", "u1", (3,)),
        ("line_cap", "u1"),
        ("fill_color", "u1", (3,)),
        ("casing_line_cap", "u1"),
        ("casing_color", "u1", (3,)),
        ("reserved0", "u1"),
        ("fill_image", "<i4"),
        ("width", "<f8"),
        ("opacity", "<f8"),
        ("fill_opacity", "<f8"),
        ("casing_width", "<f8"),
        ("dashes_off", "<u4"),
        ("dashes_len", "<u4"),
        ("casing_dashes_off", "<u4"),
        ("casing_dashes_len", "<u4"),
    ],
    align=True,
)
assert STYLE_DTYPE.itemsize == 72, STYLE_DTYPE.itemsize

AREA_DTYPE = np.dtype([("entity", "<u4"), ("style", "<u4")])

# label pass (include/osmr.h: osmr_label_style, osmr_label)
OSMR_LABEL_NODE = 0x40000000
OSMR_LSTYLE_TEXT = 1 << 0
OSMR_LSTYLE_FONT_SIZE = 1 << 1
OSMR_LSTYLE_TEXT_COLOR = 1 << 2
LABEL_STYLE_DTYPE = np.dtype(
    [
        ("icon", "<i4"), ("flags", "<u4"), ("text_key_off", "<u4"), ("text_key_len", "<u4"),
        ("text_color", "u1", (3,)), ("text_position", "u1"), ("reserved0", "<u4"), ("font_size", "<f8"),
    ],
    align=True,
)
assert LABEL_STYLE_DTYPE.itemsize == 32, LABEL_STYLE_DTYPE.itemsize
LABEL_DTYPE = np.dtype([("entity", "<u4"), ("style", "<u4")])


class IconStruct(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba", C.c_void_p)]


CLASS_STYLE_DTYPE = np.dtype([("style", "<u4"), ("order", "<u4")])  # osmr_class_style
NO_CLASS = 0xFFFFFFFF


class StatsStruct(C.Structure):
    _fields_ = [
        ("n_tiles", C.c_uint64),
        ("n_areas", C.c_uint64),
        ("n_visible_ops", C.c_uint64),
        ("n_node_refs", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("geom_bytes", C.c_uint64),
        ("mask_bytes", C.c_uint64),
        ("walk_bytes", C.c_uint64),
        ("walk_steps", C.c_uint64),
        ("ms_plan", C.c_float),
        ("ms_raster", C.c_float),
        ("ms_total", C.c_float),
        ("ms_label_layout", C.c_float),
        ("ms_label_device", C.c_float),
        ("ms_cover", C.c_float),
        ("ms_auto", C.c_float),
        ("ms_png", C.c_float),
        ("label_path", C.c_uint32),
        ("n_labels_active", C.c_uint32),
        ("n_labels_polylabel", C.c_uint32),
        ("label_attempts", C.c_uint32),
        ("ms_label_cover", C.c_float),
        ("n_label_segments", C.c_uint32),
        ("n_label_cells", C.c_uint64),
    ]


def load_icon_rgba(path: str):
    """Decode a PNG the way reference src/draw/icon.rs:14-58 does (png crate, `normalize_to_color8`):
    palette -> RGB (RGBA with tRNS), 16-bit -> 8-bit, grey+tRNS -> grey+alpha; then only RGB, RGBA and
    GrayscaleAlpha are accepted -- plain Grayscale makes Icon::load fail (icon.rs:46).
    Returns (width, height, uint8[h, w, 4]) or None when the reference would fail to load the icon.
    """
    from PIL import Image

    try:
        im = Image.open(path)
        im.load()
    except Exception:
        return None
    mode = im.mode
    has_trns = "transparency" in im.info
    if mode == "P":
        im = im.convert("RGBA" if has_trns else "RGB")
    elif mode in ("L", "1", "I;16", "I"):
        if not has_trns:
            return None  # ColorType::Grayscale -> "Unknown color type"
        im = im.convert("LA")
    elif mode == "RGB" and has_trns:
        im = im.convert("RGBA")
    mode = im.mode
    arr = np.asarray(im)
    h, w = arr.shape[0], arr.shape[1]
    out = np.empty((h, w, 4), dtype=np.uint8)
    if mode == "RGB":
        out[..., :3] = arr
        out[..., 3] = 255
    elif mode == "RGBA":
        out[...] = arr
    elif mode == "LA":
        out[..., 0] = out[..., 1] = out[..., 2] = arr[..., 0]
        out[..., 3] = arr[..., 1]
    else:
        return None
    return w, h, np.ascontiguousarray(out)


class StyleTable:
    """Interns styler.Style objects (the reference's Arc<Style>) into the flat table of osmr_set_styles and
    icon names into the icon table of osmr_set_icons (reference IconCache, src/draw/icon_cache.rs:21-45:
    lazily loaded relative to the stylesheet directory, failures cached as None)."""

    def __init__(self, icon_base_path: str | None = None, icon_loader=None):
        self.icon_base_path = icon_base_path
        self.icon_loader = icon_loader  # optional: name -> (w, h, rgba) | None, instead of reading files
        self.icon_names: list = []
        self._style_ids: dict[int, int] = {}
        self._styles_keepalive: list = []
        self.rows: list = []
        self.dashes: list[float] = []
        self._icon_ids: dict[str, int] = {}
        self.icons: list = []  # (w, h, rgba ndarray)

    def icon_id(self, name: str) -> int:
        i = self._icon_ids.get(name)
        if i is None:
            ic = None
            if self.icon_loader is not None:
                ic = self.icon_loader(name)
            elif self.icon_base_path is not None:
                ic = load_icon_rgba(os.path.join(self.icon_base_path, name))
            if ic is None:
                i = -1
            else:
                i = len(self.icons)
                self.icons.append(ic)
                self.icon_names.append(name)
            self._icon_ids[name] = i
        return i

    def add_raw_icon(self, name: str, rgba: np.ndarray) -> int:
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        i = len(self.icons)
        self.icons.append((rgba.shape[1], rgba.shape[0], rgba))
        self._icon_ids[name] = i
        return i

    def style_id(self, s) -> int:
        k = id(s)
        i = self._style_ids.get(k)
        if i is not None:
            return i
        row = np.zeros((), dtype=STYLE_DTYPE)
        flags = 0
        if s.color is not None:
            flags |= OSMR_STYLE_COLOR
            row["color"] = s.color
        if s.fill_color is not None:
            flags |= OSMR_STYLE_FILL_COLOR
            row["fill_color"] = s.fill_color
        row["fill_image"] = -1
        if s.fill_image is not None:
            flags |= OSMR_STYLE_FILL_IMAGE
            # the reference only consults the icon cache when there is no fill colour (drawer.rs:176-184)
            row["fill_image"] = self.icon_id(s.fill_image) if s.fill_color is None else -1
        if s.casing_color is not None:
            flags |= OSMR_STYLE_CASING_COLOR
            row["casing_color"] = s.casing_color
        if s.casing_width is not None:
            flags |= OSMR_STYLE_CASING_WIDTH
            row["casing_width"] = s.casing_width
        if s.width is not None:
            flags |= OSMR_STYLE_WIDTH
            row["width"] = s.width
        if s.opacity is not None:
            flags |= OSMR_STYLE_OPACITY
            row["opacity"] = s.opacity
        if s.fill_opacity is not None:
            flags |= OSMR_STYLE_FILL_OPACITY
            row["fill_opacity"] = s.fill_opacity
        if s.dashes is not None:
            flags |= OSMR_STYLE_DASHES
            row["dashes_off"] = len(self.dashes)
            row["dashes_len"] = len(s.dashes)
            self.dashes.extend(s.dashes)
        if s.casing_dashes is not None:
            flags |= OSMR_STYLE_CASING_DASHES
            row["casing_dashes_off"] = len(self.dashes)
            row["casing_dashes_len"] = len(s.casing_dashes)
            self.dashes.extend(s.casing_dashes)
        row["line_cap"] = s.line_cap
        row["casing_line_cap"] = s.casing_line_cap
        row["flags"] = flags
        i = len(self.rows)
        self.rows.append(row)
        self._style_ids[k] = i
        self._styles_keepalive.append(s)
        return i

    def styles_array(self) -> np.ndarray:
        if not self.rows:
            return np.zeros(0, dtype=STYLE_DTYPE)
        return np.array(self.rows, dtype=STYLE_DTYPE)

    def dashes_array(self) -> np.ndarray:
        return np.asarray(self.dashes, dtype=np.float64)

    def icon_structs(self):
        """(ctypes array of osmr_icon, keepalive list)"""
        arr = (IconStruct * max(1, len(self.icons)))()
        for i, (w, h, px) in enumerate(self.icons):
            arr[i].width = w
            arr[i].height = h
            arr[i].rgba = px.ctypes.data
        return arr, list(self.icons)


def styled_areas_to_array(styled, table: StyleTable) -> np.ndarray:
    """[(entity tuple (kind, local_id, global_id, tags), Style)] -> AREA_DTYPE array in styler order."""
    from .upstream.styler import KIND_MULTIPOLYGON

    out = np.empty(len(styled), dtype=AREA_DTYPE)
    for i, (ent, s) in enumerate(styled):
        e = ent[1] | (OSMR_AREA_MULTIPOLYGON if ent[0] == KIND_MULTIPOLYGON else 0)
        out[i] = (e, table.style_id(s))
    return out


class LabelStyleTable:
    """Interns styler.Style objects into the label-style table of osmr_set_label_styles and icon names into the label
    icon table of osmr_set_label_icons (reference IconCache for `icon-image`, labeler.rs:46-50)."""

    def __init__(self, icon_base_path: str | None = None, icon_loader=None):
        self.icon_base_path = icon_base_path
        self.icon_loader = icon_loader
        self.icon_names: list = []
        self._ids: dict[int, int] = {}
        self._keep: list = []
        self.rows: list = []
        self.strings = bytearray()
        self._key_off: dict[str, tuple] = {}
        self._icon_ids: dict[str, int] = {}
        self.icons: list = []

    def icon_id(self, name) -> int:
        if name is None:
            return -2
        i = self._icon_ids.get(name)
        if i is None:
            if self.icon_loader is not None:
                ic = self.icon_loader(name)
            else:
                ic = load_icon_rgba(os.path.join(self.icon_base_path, name)) if self.icon_base_path else None
            i = -1
            if ic is not None:
                i = len(self.icons)
                self.icons.append(ic)
                self.icon_names.append(name)
            self._icon_ids[name] = i
        return i

    def style_id(self, s) -> int:
        k = id(s)
        i = self._ids.get(k)
        if i is not None:
            return i
        row = np.zeros((), dtype=LABEL_STYLE_DTYPE)
        row["icon"] = self.icon_id(s.icon_image)
        ts = s.text_style
        flags = 0
        if ts is not None:
            flags |= OSMR_LSTYLE_TEXT
            ko = self._key_off.get(ts.text)
            if ko is None:
                b = ts.text.encode("utf-8")
                ko = (len(self.strings), len(b))
                self.strings += b
                self._key_off[ts.text] = ko
            row["text_key_off"], row["text_key_len"] = ko
            if ts.font_size is not None:
                flags |= OSMR_LSTYLE_FONT_SIZE
                row["font_size"] = ts.font_size
            if ts.text_color is not None:
                flags |= OSMR_LSTYLE_TEXT_COLOR
                row["text_color"] = ts.text_color
            row["text_position"] = 0 if ts.text_position is None else (1 if ts.text_position == "center" else 2)
        row["flags"] = flags
        i = len(self.rows)
        self.rows.append(row)
        self._ids[k] = i
        self._keep.append(s)
        return i

    def styles_array(self) -> np.ndarray:
        return np.array(self.rows, dtype=LABEL_STYLE_DTYPE) if self.rows else np.zeros(0, dtype=LABEL_STYLE_DTYPE)

    def icon_structs(self):
        arr = (IconStruct * max(1, len(self.icons)))()
        for i, (w, h, px) in enumerate(self.icons):
            arr[i].width = w
            arr[i].height = h
            arr[i].rgba = px.ctypes.data
        return arr, list(self.icons)

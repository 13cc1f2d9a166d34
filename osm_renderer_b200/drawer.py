"""Host-side mirror of the reference's draw interface on top of the C ABI.

Same names, argument meaning and error behaviour as the reference (file:line in /root/reference):
    Tile ................ src/tile.rs:8-13
    TilePixels(scale) ... src/draw/tile_pixels.rs:57   per-worker scratch; here it owns the GPU context
    Drawer(base_path) ... src/draw/drawer.rs:33        owns the icon cache (fill-image patterns)
    Drawer.draw_to_pixels(entities, tile, pixels, scale, styler) -> TileRenderedPixels   drawer.rs:60-131
plus the batched extension Drawer.draw_tiles_to_pixels (many tiles per launch).

Everything numeric happens in libosmr_b200.so (hand-written CUDA + its host C++); this file only marshals the
styler's output.  With a font (`Drawer(base_path, font=...)`, the reference embeds NotoSans-Regular.ttf,
text_placer.rs:299) draw_to_pixels is the complete reference call: area passes + label pass; without one it returns
the tile after the area passes (osmr_draw_tiles).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .wire import (
    AREA_DTYPE,
    OSMR_DRAW_HAS_CANVAS_COLOR,
    OSMR_DRAW_OUT_RGBA,
    OSMR_DRAW_USE_CAPS_FOR_DASHES,
    TILE_DTYPE,
    StatsStruct,
    StyleTable,
    styled_areas_to_array,
)


@dataclass(frozen=True)
class Tile:
    zoom: int
    x: int
    y: int


@dataclass
class TileRenderedPixels:
    triples: np.ndarray  # uint8 [dimension, dimension, 3]  (reference: Vec<(u8,u8,u8)>)
    dimension: int


@dataclass
class OsmEntities:
    """reader.rs:21-25: what get_entities_in_tile_with_neighbors returns (local ids into the `.bin`)."""

    reader: object
    nodes: np.ndarray
    ways: np.ndarray
    multipolygons: np.ndarray


class GpuContext:
    """Thin RAII wrapper of osmr_ctx."""

    def __init__(self, device: int = 0, share_dataset_of: "GpuContext | None" = None):
        """share_dataset_of: another context whose resident geodata this one uses by reference (osmr_ctx_create_shared: the
        reference shares its GeodataReader between worker threads, http_server.rs:42-48)."""
        self.L = _lib.load()
        h = C.c_void_p()
        if share_dataset_of is not None:
            rc = self.L.osmr_ctx_create_shared(share_dataset_of.h, C.byref(h))
        else:
            rc = self.L.osmr_ctx_create(device, C.byref(h))
        if rc != 0:
            raise _lib.OsmrError(f"osmr_ctx_create(device={device}) failed with {rc}: no usable CUDA device (no CPU fallback)")
        self.h = h
        if share_dataset_of is not None:
            self.n_nodes = getattr(share_dataset_of, "n_nodes", 0)
        self._geodata_id = None
        # what this context holds on the device: (identity of the table, rows, icons).  A context is one worker's scratch
        # (TilePixels); several of them may serve one shared Drawer, and a Drawer may be replaced by another with equally
        # long tables -- so the key is the table object and its length, tracked PER CONTEXT.
        self._style_key = None
        self._icon_key = None
        self._label_key = None
        self._font_key = None
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.osmr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.L.osmr_last_error(self.h)
            raise _lib.OsmrError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    # ---- residency ------------------------------------------------------------------------------------------
    def set_geodata(self, image: bytes):
        buf = np.frombuffer(image, dtype=np.uint8)
        self._check(self.L.osmr_set_geodata(self.h, buf.ctypes.data, len(image)), "osmr_set_geodata")
        self._geodata_id = id(image)
        self.n_nodes = int(np.frombuffer(image, dtype="<u4", count=1)[0])

    def set_table(self, table: StyleTable):
        icon_key = (id(table), len(table.icons))
        if icon_key != self._icon_key:
            icons, keep = table.icon_structs()
            self._check(self.L.osmr_set_icons(self.h, C.addressof(icons), len(table.icons)), "osmr_set_icons")
            self._icon_key = icon_key
            self._keep_table = table  # keeps id(table) from being reused while it is the key
        style_key = (id(table), len(table.rows), len(table.dashes))
        if style_key != self._style_key:
            styles = table.styles_array()
            dashes = table.dashes_array()
            self._check(
                self.L.osmr_set_styles(self.h, styles.ctypes.data, len(styles), dashes.ctypes.data, len(dashes)),
                "osmr_set_styles",
            )
            self._style_key = style_key
            self._keep_table = table

    # ---- drawing ------------------------------------------------------------------------------------------------
    @staticmethod
    def _flags(canvas_rgb, use_caps_for_dashes: bool, rgba: bool):
        flags = OSMR_DRAW_USE_CAPS_FOR_DASHES if use_caps_for_dashes else 0
        canvas = np.zeros(3, dtype=np.uint8)
        if canvas_rgb is not None:
            flags |= OSMR_DRAW_HAS_CANVAS_COLOR
            canvas[:] = canvas_rgb
        if rgba:
            flags |= OSMR_DRAW_OUT_RGBA
        return flags, canvas

    def draw_tiles(self, tiles, area_begin, areas, canvas_rgb, use_caps_for_dashes=True, rgba=False, out=None):
        """osmr_draw_tiles with host buffers: returns uint8 [n_tiles, D, D, 3|4]."""
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
        areas = np.ascontiguousarray(areas, dtype=AREA_DTYPE)
        n = len(tiles)
        d = 256 * int(tiles["scale"][0]) if n else 256
        ch = 4 if rgba else 3
        if out is None:
            out = np.empty((n, d, d, ch), dtype=np.uint8)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, rgba)
        self._check(
            self.L.osmr_draw_tiles(
                self.h, tiles.ctypes.data, n, area_begin.ctypes.data, areas.ctypes.data, canvas.ctypes.data, flags, out.ctypes.data
            ),
            "osmr_draw_tiles",
        )
        return out

    def batch_upload(self, tiles, area_begin, areas):
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
        areas = np.ascontiguousarray(areas, dtype=AREA_DTYPE)
        self._check(
            self.L.osmr_batch_upload(self.h, tiles.ctypes.data, len(tiles), area_begin.ctypes.data, areas.ctypes.data),
            "osmr_batch_upload",
        )
        self._batch_shape = (len(tiles), 256 * int(tiles["scale"][0]))

    def batch_draw(self, canvas_rgb, use_caps_for_dashes=True, rgba=False, out=None) -> float:
        """Draw the resident batch; returns device milliseconds (CUDA events on the context stream)."""
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, rgba)
        ms = C.c_float(0.0)
        ptr = out.ctypes.data if out is not None else None
        self._check(self.L.osmr_batch_draw(self.h, canvas.ctypes.data, flags, ptr, C.byref(ms)), "osmr_batch_draw")
        return float(ms.value)

    def batch_upload_labeled(self, tiles, area_begin, areas, label_begin, labels):
        from .wire import LABEL_DTYPE

        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
        areas = np.ascontiguousarray(areas, dtype=AREA_DTYPE)
        label_begin = np.ascontiguousarray(label_begin, dtype=np.uint32)
        labels = np.ascontiguousarray(labels, dtype=LABEL_DTYPE)
        self._check(
            self.L.osmr_batch_upload_labeled(self.h, tiles.ctypes.data, len(tiles), area_begin.ctypes.data, areas.ctypes.data,
                                             label_begin.ctypes.data, labels.ctypes.data),
            "osmr_batch_upload_labeled",
        )
        self._batch_shape = (len(tiles), 256 * int(tiles["scale"][0]))

    def batch_draw_labeled(self, canvas_rgb, use_caps_for_dashes=True, rgba=False, out=None) -> float:
        """Draw the resident labelled batch (area passes + label pass); returns device milliseconds."""
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, rgba)
        ms = C.c_float(0.0)
        ptr = out.ctypes.data if out is not None else None
        self._check(self.L.osmr_batch_draw_labeled(self.h, canvas.ctypes.data, flags, ptr, C.byref(ms)), "osmr_batch_draw_labeled")
        return float(ms.value)

    def stats(self) -> dict:
        st = StatsStruct()
        self._check(self.L.osmr_get_stats(self.h, C.byref(st)), "osmr_get_stats")
        return {k: getattr(st, k) for k, _ in StatsStruct._fields_}

    def project_nodes(self, tile) -> np.ndarray:
        t = np.array([tuple(tile)], dtype=TILE_DTYPE)
        out = np.zeros((self.n_nodes, 2), dtype=np.int32)
        self._check(self.L.osmr_project_nodes(self.h, t.ctypes.data, out.ctypes.data), "osmr_project_nodes")
        return out

    # ---- label pass ---------------------------------------------------------------------------------------------
    def set_font(self, ttf: bytes):
        key = (id(ttf), len(ttf))
        if key == self._font_key:
            return
        buf = np.frombuffer(ttf, dtype=np.uint8)
        self._check(self.L.osmr_set_font(self.h, buf.ctypes.data, len(ttf)), "osmr_set_font")
        self._font_key = key
        self._keep_font = ttf

    def set_label_table(self, ltable):
        key = (id(ltable), len(ltable.rows), len(ltable.icons), len(ltable.strings))
        if key == self._label_key:
            return
        icons, keep = ltable.icon_structs()
        self._check(self.L.osmr_set_label_icons(self.h, C.addressof(icons), len(ltable.icons)), "osmr_set_label_icons")
        styles = ltable.styles_array()
        blob = bytes(ltable.strings)
        self._check(self.L.osmr_set_label_styles(self.h, styles.ctypes.data, len(styles), blob, len(blob)), "osmr_set_label_styles")
        self._label_key = key
        self._keep_ltable = ltable

    def draw_tiles_labeled(self, tiles, area_begin, areas, label_begin, labels, canvas_rgb, use_caps_for_dashes=True, rgba=False):
        """osmr_draw_tiles_labeled: area passes + label pass, host buffers."""
        from .wire import LABEL_DTYPE

        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
        areas = np.ascontiguousarray(areas, dtype=AREA_DTYPE)
        label_begin = np.ascontiguousarray(label_begin, dtype=np.uint32)
        labels = np.ascontiguousarray(labels, dtype=LABEL_DTYPE)
        n = len(tiles)
        d = 256 * int(tiles["scale"][0])
        out = np.empty((n, d, d, 4 if rgba else 3), dtype=np.uint8)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, rgba)
        self._check(
            self.L.osmr_draw_tiles_labeled(self.h, tiles.ctypes.data, n, area_begin.ctypes.data, areas.ctypes.data,
                                           label_begin.ctypes.data, labels.ctypes.data, canvas.ctypes.data, flags, out.ctypes.data),
            "osmr_draw_tiles_labeled",
        )
        return out

    # ---- f3: styled-area lists built on the device ---------------------------------------------------------------
    def set_zoom_styles(self, zoom: int, way_class, mp_class, class_begin, class_styles):
        """osmr_set_zoom_styles: per-zoom style classes of every way / multipolygon (the host styler's cache)."""
        from .wire import CLASS_STYLE_DTYPE

        way_class = np.ascontiguousarray(way_class, dtype=np.uint32)
        mp_class = np.ascontiguousarray(mp_class, dtype=np.uint32)
        class_begin = np.ascontiguousarray(class_begin, dtype=np.uint32)
        class_styles = np.ascontiguousarray(class_styles, dtype=CLASS_STYLE_DTYPE)
        self._check(
            self.L.osmr_set_zoom_styles(self.h, zoom, way_class.ctypes.data, mp_class.ctypes.data, class_begin.ctypes.data,
                                        class_styles.ctypes.data, len(class_begin) - 1),
            "osmr_set_zoom_styles",
        )

    def draw_tiles_auto(self, tiles, canvas_rgb, use_caps_for_dashes=True, rgba=False, out=None):
        """osmr_draw_tiles_auto: only the tile list is sent; candidates, styling order and drawing happen on the device."""
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        n = len(tiles)
        d = 256 * int(tiles["scale"][0]) if n else 256
        if out is None:
            out = np.empty((n, d, d, 4 if rgba else 3), dtype=np.uint8)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, rgba)
        self._check(self.L.osmr_draw_tiles_auto(self.h, tiles.ctypes.data, n, canvas.ctypes.data, flags, out.ctypes.data), "osmr_draw_tiles_auto")
        self._auto_tiles = n
        return out

    def auto_readback(self):
        """(area_begin, areas) of the last draw_tiles_auto call."""
        begins = np.zeros(self._auto_tiles + 1, dtype=np.uint32)
        self._check(self.L.osmr_auto_readback(self.h, begins.ctypes.data, None, 0), "osmr_auto_readback")
        areas = np.zeros(int(begins[-1]), dtype=AREA_DTYPE)
        self._check(self.L.osmr_auto_readback(self.h, begins.ctypes.data, areas.ctypes.data, len(areas)), "osmr_auto_readback")
        return begins, areas

    # ---- f4: PNG files ------------------------------------------------------------------------------------------------
    def draw_tiles_png(self, tiles, area_begin, areas, canvas_rgb, use_caps_for_dashes=True):
        """osmr_draw_tiles_png: list of `bytes`, one PNG file per tile (Drawer::draw_tile, drawer.rs:40-58)."""
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        area_begin = np.ascontiguousarray(area_begin, dtype=np.uint32)
        areas = np.ascontiguousarray(areas, dtype=AREA_DTYPE)
        n = len(tiles)
        cap = n * int(self.L.osmr_png_bound(int(tiles["scale"][0]))) if n else 0
        buf = np.empty(cap, dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, False)
        self._check(
            self.L.osmr_draw_tiles_png(self.h, tiles.ctypes.data, n, area_begin.ctypes.data, areas.ctypes.data, canvas.ctypes.data, flags,
                                       buf.ctypes.data, cap, offs.ctypes.data),
            "osmr_draw_tiles_png",
        )
        return [buf[int(offs[i]) : int(offs[i + 1])].tobytes() for i in range(n)]

    def draw_tiles_auto_png(self, tiles, canvas_rgb, use_caps_for_dashes=True):
        """osmr_draw_tiles_auto_png: tile list in, one PNG file per tile out (Drawer::draw_tile behind the server's lookup)."""
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        n = len(tiles)
        cap = n * int(self.L.osmr_png_bound(int(tiles["scale"][0]))) if n else 0
        buf = np.empty(cap, dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, False)
        self._check(
            self.L.osmr_draw_tiles_auto_png(self.h, tiles.ctypes.data, n, canvas.ctypes.data, flags, buf.ctypes.data, cap, offs.ctypes.data),
            "osmr_draw_tiles_auto_png",
        )
        self._auto_tiles = n
        return [buf[int(offs[i]) : int(offs[i + 1])].tobytes() for i in range(n)]

    # ---- f3 for the label pass: label lists built on the device ---------------------------------------------------------
    def set_zoom_label_styles(self, zoom: int, node_class, way_class, mp_class, class_begin, class_styles):
        """osmr_set_zoom_label_styles: per-zoom label style classes of every node / way / multipolygon."""
        from .wire import CLASS_STYLE_DTYPE

        node_class = np.ascontiguousarray(node_class, dtype=np.uint32)
        way_class = np.ascontiguousarray(way_class, dtype=np.uint32)
        mp_class = np.ascontiguousarray(mp_class, dtype=np.uint32)
        class_begin = np.ascontiguousarray(class_begin, dtype=np.uint32)
        class_styles = np.ascontiguousarray(class_styles, dtype=CLASS_STYLE_DTYPE)
        self._check(
            self.L.osmr_set_zoom_label_styles(self.h, zoom, node_class.ctypes.data, way_class.ctypes.data, mp_class.ctypes.data,
                                              class_begin.ctypes.data, class_styles.ctypes.data, len(class_begin) - 1),
            "osmr_set_zoom_label_styles",
        )

    def draw_tiles_auto_labeled(self, tiles, canvas_rgb, use_caps_for_dashes=True, rgba=False, out=None):
        """osmr_draw_tiles_auto_labeled: the whole draw_to_pixels from a tile list (area AND label lists built on the device)."""
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        n = len(tiles)
        d = 256 * int(tiles["scale"][0]) if n else 256
        if out is None:
            out = np.empty((n, d, d, 4 if rgba else 3), dtype=np.uint8)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, rgba)
        self._check(self.L.osmr_draw_tiles_auto_labeled(self.h, tiles.ctypes.data, n, canvas.ctypes.data, flags, out.ctypes.data),
                    "osmr_draw_tiles_auto_labeled")
        self._auto_tiles = n
        return out

    def auto_readback_labels(self):
        """(label_begin, labels) of the last draw_tiles_auto_labeled* call."""
        from .wire import LABEL_DTYPE

        begins = np.zeros(self._auto_tiles + 1, dtype=np.uint32)
        self._check(self.L.osmr_auto_readback_labels(self.h, begins.ctypes.data, None, 0), "osmr_auto_readback_labels")
        labels = np.zeros(int(begins[-1]), dtype=LABEL_DTYPE)
        self._check(self.L.osmr_auto_readback_labels(self.h, begins.ctypes.data, labels.ctypes.data, len(labels)), "osmr_auto_readback_labels")
        return begins, labels

    def draw_tiles_auto_labeled_png(self, tiles, canvas_rgb, use_caps_for_dashes=True):
        """osmr_draw_tiles_auto_labeled_png: tile list in, one PNG file per tile out, label pass included (Drawer::draw_tile)."""
        tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
        n = len(tiles)
        cap = n * int(self.L.osmr_png_bound(int(tiles["scale"][0]))) if n else 0
        buf = np.empty(cap, dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        flags, canvas = self._flags(canvas_rgb, use_caps_for_dashes, False)
        self._check(
            self.L.osmr_draw_tiles_auto_labeled_png(self.h, tiles.ctypes.data, n, canvas.ctypes.data, flags, buf.ctypes.data, cap, offs.ctypes.data),
            "osmr_draw_tiles_auto_labeled_png",
        )
        self._auto_tiles = n
        return [buf[int(offs[i]) : int(offs[i + 1])].tobytes() for i in range(n)]

    def rgb_to_png(self, images):
        """osmr_rgb_to_png: uint8 [n, D, D, 3] (D = 256 * scale) -> list of PNG files (png_writer.rs:4-21)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        n, d = images.shape[0], images.shape[1]
        assert images.shape == (n, d, d, 3) and d % 256 == 0
        scale = d // 256
        cap = n * int(self.L.osmr_png_bound(scale))
        buf = np.empty(cap, dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        self._check(self.L.osmr_rgb_to_png(self.h, images.ctypes.data, n, scale, buf.ctypes.data, cap, offs.ctypes.data), "osmr_rgb_to_png")
        return [buf[int(offs[i]) : int(offs[i + 1])].tobytes() for i in range(n)]

    def debug_set(self, key: str, value: int):
        self._check(self.L.osmr_debug_set(self.h, key.encode(), value), "osmr_debug_set")


class TilePixels:
    """Per-worker scratch (tile_pixels.rs:57).  Reallocated by the caller when the scale changes, exactly like
    the reference server does (http_server.rs:156-160); here it owns the GPU context and its device buffers."""

    def __init__(self, scale: int, device: int = 0):
        self.scale = int(scale)
        self.ctx = GpuContext(device)

    def dimension(self) -> int:
        return 256 * self.scale


class Drawer:
    def __init__(self, base_path: str | None, font: bytes | None = None, icon_loader=None):
        from .wire import LabelStyleTable

        self.table = StyleTable(base_path, icon_loader)  # icon cache (fill patterns) + interned styles
        self.ltable = LabelStyleTable(base_path, icon_loader)  # icon cache (label icons) + interned label styles
        self.font = font

    def _builders(self, reader, styler):
        from .upstream.pipeline import LabelListBuilder, TileStyler

        ts = getattr(reader, "_tile_styler", None)
        if ts is None or ts.styler is not styler or ts.table is not self.table:
            ts = TileStyler(reader, styler, self.table)
            reader._tile_styler = ts
            reader._label_builder = LabelListBuilder(ts, None)
        return ts, reader._label_builder

    def _areas_for(self, entities: OsmEntities, tile: Tile, styler):
        ts, _ = self._builders(entities.reader, styler)
        way_ents = [ts.way_entity(int(w)) for w in entities.ways]
        mp_ents = [ts.mp_entity(int(m)) for m in entities.multipolygons]
        styled = styler.style_areas(way_ents, mp_ents, tile.zoom, False)  # drawer.rs:75-78
        return styled_areas_to_array(styled, self.table)

    def _labels_for(self, entities: OsmEntities, tile: Tile, styler):
        """drawer.rs:106-119: areas styled for labels, then nodes, as osmr_label records."""
        from .upstream.styler import KIND_MULTIPOLYGON, KIND_NODE
        from .wire import LABEL_DTYPE, OSMR_AREA_MULTIPOLYGON, OSMR_LABEL_NODE

        ts, lb = self._builders(entities.reader, styler)
        way_ents = [ts.way_entity(int(w)) for w in entities.ways]
        mp_ents = [ts.mp_entity(int(m)) for m in entities.multipolygons]
        seq = list(styler.style_areas(way_ents, mp_ents, tile.zoom, True))
        seq += list(styler.style_entities([lb.node_entity(int(n)) for n in entities.nodes], tile.zoom, True))
        out = np.empty(len(seq), dtype=LABEL_DTYPE)
        for i, (ent, s) in enumerate(seq):
            flag = OSMR_LABEL_NODE if ent[0] == KIND_NODE else (OSMR_AREA_MULTIPOLYGON if ent[0] == KIND_MULTIPOLYGON else 0)
            out[i] = (ent[1] | flag, self.ltable.style_id(s))
        return out

    def draw_to_pixels(self, entities: OsmEntities, tile: Tile, pixels: TilePixels, scale: int, styler) -> TileRenderedPixels:
        out = self.draw_tiles_to_pixels([entities], [tile], pixels, scale, styler)
        return out[0]

    def draw_tiles_to_pixels(self, entities_list, tiles, pixels: TilePixels, scale: int, styler):
        """Batched form: one launch for many (entities, tile) pairs sharing a reader, scale and styler."""
        if scale != pixels.scale:
            raise ValueError("TilePixels was created for another scale (reference: http_server.rs:156-160)")
        ctx = pixels.ctx
        reader = entities_list[0].reader
        if ctx._geodata_id != id(reader.data):
            ctx.set_geodata(reader.data)
        parts = [self._areas_for(e, t, styler) for e, t in zip(entities_list, tiles)]
        ctx.set_table(self.table)
        begins = np.zeros(len(parts) + 1, dtype=np.uint32)
        begins[1:] = np.cumsum([len(p) for p in parts])
        areas = np.concatenate(parts) if parts else np.zeros(0, dtype=AREA_DTYPE)
        tarr = np.array([(t.zoom, t.x, t.y, scale) for t in tiles], dtype=TILE_DTYPE)
        if self.font is None:
            img = ctx.draw_tiles(tarr, begins, areas, styler.canvas_fill_color, styler.use_caps_for_dashes)
        else:
            lparts = [self._labels_for(e, t, styler) for e, t in zip(entities_list, tiles)]
            lbegins = np.zeros(len(lparts) + 1, dtype=np.uint32)
            lbegins[1:] = np.cumsum([len(p) for p in lparts])
            labels = np.concatenate(lparts)
            ctx.set_font(self.font)  # both are no-ops when THIS context already holds them (tracked per context:
            ctx.set_label_table(self.ltable)  # one Drawer is shared by many TilePixels, http_server.rs:42-48,69-72)
            img = ctx.draw_tiles_labeled(tarr, begins, areas, lbegins, labels, styler.canvas_fill_color, styler.use_caps_for_dashes)
        d = 256 * scale
        return [TileRenderedPixels(img[i], d) for i in range(len(tiles))]

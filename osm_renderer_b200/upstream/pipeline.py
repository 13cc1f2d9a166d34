"""Glue of the upstream restatement: tile -> candidate entities -> ordered styled areas.

Mirrors the caller side of the draw path: reference src/http_server.rs:150-171 /
tests/test_rendering.rs:88-91 (`get_entities_in_tile_with_neighbors` then `draw_to_pixels`, whose first
step is `styler.style_areas(ways, multipolygons, zoom, false)`, src/draw/drawer.rs:75-78).
"""
from __future__ import annotations

import numpy as np

from ..wire import AREA_DTYPE, StyleTable, styled_areas_to_array
from .geodata import GeodataReader
from .styler import KIND_MULTIPOLYGON, KIND_WAY_CLOSED, KIND_WAY_OPEN, Styler


class TileStyler:
    def __init__(self, reader: GeodataReader, styler: Styler, table: StyleTable):
        self.reader = reader
        self.styler = styler
        self.table = table
        self._way_ent: dict = {}
        self._mp_ent: dict = {}

    def way_entity(self, w: int):
        e = self._way_ent.get(w)
        if e is None:
            rd = self.reader
            kind = KIND_WAY_CLOSED if rd.way_is_closed(w) else KIND_WAY_OPEN
            e = (kind, w, int(rd.ways[w]["id"]), rd.way_tags(w))
            self._way_ent[w] = e
        return e

    def mp_entity(self, m: int):
        e = self._mp_ent.get(m)
        if e is None:
            rd = self.reader
            e = (KIND_MULTIPOLYGON, m, int(rd.multipolygons[m]["id"]), rd.multipolygon_tags(m))
            self._mp_ent[m] = e
        return e

    def styled_areas(self, zoom: int, x: int, y: int, for_labels: bool = False):
        _, ways, mps = self.reader.get_entities_in_tile_with_neighbors(zoom, x, y)
        way_ents = [self.way_entity(int(w)) for w in ways]
        mp_ents = [self.mp_entity(int(m)) for m in mps]
        return self.styler.style_areas(way_ents, mp_ents, zoom, for_labels)

    def areas_array(self, zoom: int, x: int, y: int) -> np.ndarray:
        return styled_areas_to_array(self.styled_areas(zoom, x, y, False), self.table)


def build_batch(ts: TileStyler, tiles):
    """tiles: iterable of (zoom, x, y, scale) -> (TILE array, area_begin u32[n+1], AREA array)."""
    from ..wire import TILE_DTYPE

    tiles = list(tiles)
    tarr = np.array(tiles, dtype=TILE_DTYPE) if tiles else np.zeros(0, dtype=TILE_DTYPE)
    begins = [0]
    parts = []
    cache: dict = {}
    for (z, x, y, s) in tiles:
        a = cache.get((z, x, y))
        if a is None:
            a = ts.areas_array(z, x, y)
            cache[(z, x, y)] = a
        parts.append(a)
        begins.append(begins[-1] + len(a))
    areas = np.concatenate(parts) if parts else np.zeros(0, dtype=AREA_DTYPE)
    return tarr, np.asarray(begins, dtype=np.uint32), areas

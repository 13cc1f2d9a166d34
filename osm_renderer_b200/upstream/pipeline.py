"""Glue of the upstream restatement: tile -> candidate entities -> ordered styled areas.

Mirrors the caller side of the draw path: reference src/http_server.rs:150-171 /
tests/test_rendering.rs:88-91 (`get_entities_in_tile_with_neighbors` then `draw_to_pixels`, whose first
step is `styler.style_areas(ways, multipolygons, zoom, false)`, src/draw/drawer.rs:75-78).
"""
from __future__ import annotations

import os

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


class FastBatchBuilder:
    """Vectorised equivalent of TileStyler.areas_array for large datasets (the synthetic metro).

    Same semantics as reader.rs:60-100 + styler.rs:115-203; per distinct (closedness, tag list) the styles are
    computed once through Styler.styles_for, the per-tile work is numpy (candidate selection, expansion, and
    the painter's-order sort of styler.rs:246-272).  tests/test_upstream.py checks it against TileStyler.
    """

    def __init__(self, reader: GeodataReader, styler: Styler, table: StyleTable):
        self.rd = reader
        self.styler = styler
        self.table = table
        rd = reader
        ints = rd.ints
        w = rd.ways
        wlen = w["len"].astype(np.int64)
        woff = w["off"].astype(np.int64)
        first = ints[np.minimum(woff, max(len(ints) - 1, 0))] if len(ints) else np.zeros(len(w), dtype=np.uint32)
        last = ints[np.clip(woff + wlen - 1, 0, max(len(ints) - 1, 0))] if len(ints) else first
        nlat, nlon = rd.nodes["lat"], rd.nodes["lon"]
        closed = (wlen > 2) & (nlat[first] == nlat[last]) & (nlon[first] == nlon[last])
        self.way_kind = np.where(closed, KIND_WAY_CLOSED, KIND_WAY_OPEN).astype(np.int64)
        self.way_gid = w["id"].astype(np.int64)
        self.way_tagkey = (w["tags_off"].astype(np.int64) << 32) | w["tags_len"].astype(np.int64)
        m = rd.multipolygons
        self.mp_gid = m["id"].astype(np.int64)
        self.mp_tagkey = (m["tags_off"].astype(np.int64) << 32) | m["tags_len"].astype(np.int64)
        self._tx = rd.tiles["x"].astype(np.int64)
        self._ty = rd.tiles["y"].astype(np.int64)
        self._cache: dict = {}

    def _candidates(self, zoom, x, y):
        rd = self.rd
        mul = 1 << (18 - zoom)
        xa, xb = (x - 1) * mul, (x + 2) * mul - 1
        ya, yb = (y - 1) * mul, (y + 2) * mul - 1
        lo = np.searchsorted(self._tx, max(xa, 0), side="left")
        hi = np.searchsorted(self._tx, xb, side="right")
        sel = np.nonzero((self._ty[lo:hi] >= ya) & (self._ty[lo:hi] <= yb))[0] + lo
        recs = rd.tiles[sel]

        def gather(off_name, len_name):
            offs = recs[off_name].astype(np.int64)
            lens = recs[len_name].astype(np.int64)
            tot = int(lens.sum())
            if tot == 0:
                return np.zeros(0, dtype=np.int64)
            rep = np.repeat(np.arange(len(offs)), lens)
            start = np.concatenate([[0], np.cumsum(lens)[:-1]])
            idx = offs[rep] + (np.arange(tot) - start[rep])
            return np.unique(rd.ints[idx]).astype(np.int64)

        ways = gather("w_off", "w_len")
        mps = gather("m_off", "m_len")
        if len(mps):
            mps = mps[rd.multipolygons["len"][mps] > 0]
        return ways, mps

    def _styles_csr(self, zoom, kind, tagkeys):
        """for an array of tag keys: CSR (ptr, style ids, layer, fg, z) through the memo."""
        uniq, inv = np.unique(tagkeys, return_inverse=True)
        lists = []
        for k in uniq:
            ck = (zoom, kind, int(k))
            ent = self._cache.get(ck)
            if ent is None:
                tags = self.rd.tags_of(int(k) >> 32, int(k) & 0xFFFFFFFF)
                styles = self.styler.styles_for(tags, zoom, kind)
                ent = (
                    np.array([self.table.style_id(s) for s in styles], dtype=np.int64),
                    np.array([s.layer or 0 for s in styles], dtype=np.int64),
                    np.array([1 if s.is_foreground_fill else 0 for s in styles], dtype=np.int64),
                    np.array([s.z_index for s in styles], dtype=np.float64),
                )
                self._cache[ck] = ent
            lists.append(ent)
        return inv, lists

    def _expand(self, zoom, ids, kinds, gids, tagkeys, is_way):
        if len(ids) == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z, z, np.zeros(0), z, z
        out = []
        for kind in np.unique(kinds):
            m = kinds == kind
            inv, lists = self._styles_csr(zoom, int(kind), tagkeys[m])
            cnt = np.array([len(l[0]) for l in lists], dtype=np.int64)[inv]
            tot = int(cnt.sum())
            if tot == 0:
                continue
            rep = np.repeat(np.arange(len(inv)), cnt)
            start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            within = np.arange(tot) - start[rep]
            # ragged gather through a padded table
            width = max(len(l[0]) for l in lists)
            pad = lambda j, dt: np.array([np.pad(l[j], (0, width - len(l[j]))) for l in lists], dtype=dt)
            sid, lay, fg, zi = pad(0, np.int64), pad(1, np.int64), pad(2, np.int64), pad(3, np.float64)
            u = inv[rep]
            ent_ids = ids[m][rep]
            out.append((ent_ids, sid[u, within], lay[u, within], fg[u, within], zi[u, within], gids[m][rep], within))
        if not out:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z, z, np.zeros(0), z, z
        return tuple(np.concatenate([o[i] for o in out]) for i in range(7))

    def areas_array(self, zoom, x, y) -> np.ndarray:
        from ..wire import OSMR_AREA_MULTIPOLYGON

        ways, mps = self._candidates(zoom, x, y)
        we = self._expand(zoom, ways, self.way_kind[ways], self.way_gid[ways], self.way_tagkey[ways], True)
        me = self._expand(zoom, mps, np.full(len(mps), KIND_MULTIPOLYGON), self.mp_gid[mps], self.mp_tagkey[mps], False)
        ent = np.concatenate([me[0] | OSMR_AREA_MULTIPOLYGON, we[0]])
        sid = np.concatenate([me[1], we[1]])
        lay = np.concatenate([me[2], we[2]])
        fg = np.concatenate([me[3], we[3]])
        zi = np.concatenate([me[4], we[4]])
        gid = np.concatenate([me[5], we[5]])
        is_way = np.concatenate([np.zeros(len(me[0]), dtype=np.int64), np.ones(len(we[0]), dtype=np.int64)])
        loc = np.concatenate([me[0], we[0]])  # local id: entities are visited in local-id order (stable sort)
        within = np.concatenate([me[6], we[6]])
        order = np.lexsort((within, loc, is_way, gid, zi, fg, lay))
        out = np.empty(len(order), dtype=AREA_DTYPE)
        out["entity"] = ent[order]
        out["style"] = sid[order]
        return out


def _label_methods():
    """labels_array for FastBatchBuilder (kept next to it; bound below)."""
    from ..wire import LABEL_DTYPE, OSMR_AREA_MULTIPOLYGON, OSMR_LABEL_NODE
    from .styler import KIND_NODE

    def _label_lists(self, zoom, kind, tagkeys, ltable):
        """per distinct tag list: (label style ids, layer, z_index) through a memo -- the reference's StyleCache entry"""
        uniq, inv = np.unique(tagkeys, return_inverse=True)
        lists = []
        for k in uniq:
            ck = ("L", zoom, kind, int(k), id(ltable))
            ent = self._cache.get(ck)
            if ent is None:
                tags = self.rd.tags_of(int(k) >> 32, int(k) & 0xFFFFFFFF)
                styles = self.styler.styles_for(tags, zoom, kind)
                ent = (
                    np.array([ltable.style_id(s) for s in styles], dtype=np.int64),
                    np.array([s.layer or 0 for s in styles], dtype=np.int64),
                    np.array([s.z_index for s in styles], dtype=np.float64),
                )
                self._cache[ck] = ent
            lists.append(ent)
        return inv, lists

    def _expand_labels(self, zoom, ids, kinds, gids, tagkeys, ltable):
        out = []
        for kind in np.unique(kinds):
            m = kinds == kind
            inv, lists = _label_lists(self, zoom, int(kind), tagkeys[m], ltable)
            cnt = np.array([len(l[0]) for l in lists], dtype=np.int64)[inv]
            tot = int(cnt.sum())
            if tot == 0:
                continue
            rep = np.repeat(np.arange(len(inv)), cnt)
            start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
            within = np.arange(tot) - start[rep]
            width = max(len(l[0]) for l in lists)
            pad = lambda j, dt: np.array([np.pad(l[j], (0, width - len(l[j]))) for l in lists], dtype=dt)
            sid, lay, zi = pad(0, np.int64), pad(1, np.int64), pad(2, np.float64)
            u = inv[rep]
            out.append((ids[m][rep], sid[u, within], lay[u, within], zi[u, within], gids[m][rep], within))
        if not out:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z, np.zeros(0), z, z
        return tuple(np.concatenate([o[i] for o in out]) for i in range(6))

    def labels_array(self, zoom, x, y, ltable) -> np.ndarray:
        """The label generations of a tile as osmr_label records, in the reference's order (drawer.rs:106-119): the areas
        styled with for_labels = true (styler.rs:168-203: sorted by (layer, z_index, global id) -- is_foreground_fill is
        ignored, styler.rs:263 --, multipolygons first on ties), then the styled nodes (styler.rs:115-166)."""
        rd = self.rd
        ways, mps = self._candidates(zoom, x, y)
        we = _expand_labels(self, zoom, ways, self.way_kind[ways], self.way_gid[ways], self.way_tagkey[ways], ltable)
        me = _expand_labels(self, zoom, mps, np.full(len(mps), KIND_MULTIPOLYGON), self.mp_gid[mps], self.mp_tagkey[mps], ltable)
        ent = np.concatenate([me[0] | OSMR_AREA_MULTIPOLYGON, we[0]])
        sid = np.concatenate([me[1], we[1]])
        lay = np.concatenate([me[2], we[2]])
        zi = np.concatenate([me[3], we[3]])
        gid = np.concatenate([me[4], we[4]])
        is_way = np.concatenate([np.zeros(len(me[0]), dtype=np.int64), np.ones(len(we[0]), dtype=np.int64)])
        loc = np.concatenate([me[0], we[0]])
        within = np.concatenate([me[5], we[5]])
        order = np.lexsort((within, loc, is_way, gid, zi, lay))
        parts_e, parts_s = [ent[order]], [sid[order]]
        # nodes of the 3x3 neighbourhood that carry tags (an untagged node matches no `node` selector with a condition;
        # unconditional node rules would style it, so keep every node when the stylesheet has such a rule)
        nodes = self._candidate_nodes(zoom, x, y)
        if len(nodes):
            ntl = rd.nodes["tags_len"][nodes].astype(np.int64)
            ntk = np.where(ntl > 0, (rd.nodes["tags_off"][nodes].astype(np.int64) << 32) | ntl, 0)  # untagged nodes: one class
            ne = _expand_labels(self, zoom, nodes, np.full(len(nodes), KIND_NODE), rd.nodes["id"][nodes].astype(np.int64), ntk, ltable)
            order_n = np.lexsort((ne[5], ne[0], ne[4], ne[3], ne[2]))
            parts_e.append(ne[0][order_n] | OSMR_LABEL_NODE)
            parts_s.append(ne[1][order_n])
        out = np.empty(sum(len(p) for p in parts_e), dtype=LABEL_DTYPE)
        out["entity"] = np.concatenate(parts_e)
        out["style"] = np.concatenate(parts_s)
        return out

    def _candidate_nodes(self, zoom, x, y):
        rd = self.rd
        mul = 1 << (18 - zoom)
        xa, xb = (x - 1) * mul, (x + 2) * mul - 1
        ya, yb = (y - 1) * mul, (y + 2) * mul - 1
        lo = np.searchsorted(self._tx, max(xa, 0), side="left")
        hi = np.searchsorted(self._tx, xb, side="right")
        sel = np.nonzero((self._ty[lo:hi] >= ya) & (self._ty[lo:hi] <= yb))[0] + lo
        recs = rd.tiles[sel]
        offs = recs["n_off"].astype(np.int64)
        lens = recs["n_len"].astype(np.int64)
        tot = int(lens.sum())
        if tot == 0:
            return np.zeros(0, dtype=np.int64)
        rep = np.repeat(np.arange(len(offs)), lens)
        start = np.concatenate([[0], np.cumsum(lens)[:-1]])
        idx = offs[rep] + (np.arange(tot) - start[rep])
        return np.unique(rd.ints[idx]).astype(np.int64)

    _label_methods.label_lists = _label_lists
    return labels_array, _candidate_nodes


FastBatchBuilder.labels_array, FastBatchBuilder._candidate_nodes = _label_methods()


def zoom_class_tables(fb: "FastBatchBuilder", zoom: int):
    """The per-zoom style classes osmr_set_zoom_styles takes: what the reference's StyleCache
    (style_cache.rs:68-87) would hold after every way and multipolygon has been styled once at `zoom`.

    Returns (way_class u32[n_ways], mp_class u32[n_mps], class_begin u32[n_classes+1], class_styles CLASS_STYLE_DTYPE).
    A class is a distinct (closedness, tag list); `order` is the dense rank of (layer, is_foreground_fill, z_index).
    """
    from ..wire import CLASS_STYLE_DTYPE, NO_CLASS

    lists = []  # per class: (style ids, layer, fg, z)
    def classes_of(kinds, tagkeys):
        out = np.full(len(kinds), NO_CLASS, dtype=np.uint32)
        for kind in np.unique(kinds):
            m = np.nonzero(kinds == kind)[0]
            inv, ls = fb._styles_csr(zoom, int(kind), tagkeys[m])
            base = len(lists)
            lists.extend(ls)
            out[m] = (base + inv).astype(np.uint32)
        return out

    way_class = classes_of(fb.way_kind, fb.way_tagkey)
    mp_class = classes_of(np.full(len(fb.mp_gid), KIND_MULTIPOLYGON, dtype=np.int64), fb.mp_tagkey)
    counts = np.array([len(l[0]) for l in lists], dtype=np.int64)
    class_begin = np.zeros(len(lists) + 1, dtype=np.uint32)
    class_begin[1:] = np.cumsum(counts)
    n = int(class_begin[-1])
    cs = np.zeros(n, dtype=CLASS_STYLE_DTYPE)
    if n:
        sid = np.concatenate([l[0] for l in lists])
        lay = np.concatenate([l[1] for l in lists])
        fg = np.concatenate([l[2] for l in lists])
        zi = np.concatenate([l[3] for l in lists])
        trip = np.stack([lay.astype(np.float64), fg.astype(np.float64), zi], axis=1)
        _, order = np.unique(trip, axis=0, return_inverse=True)  # rows sort lexicographically: layer, fg, z_index
        cs["style"] = sid
        cs["order"] = np.asarray(order).reshape(-1)
    return way_class, mp_class, class_begin, cs


def zoom_label_class_tables(fb: "FastBatchBuilder", zoom: int, ltable):
    """The per-zoom LABEL style classes osmr_set_zoom_label_styles takes: the StyleCache contents (style_cache.rs:68-87) for
    every node, way and multipolygon at `zoom`, with the styles interned in `ltable` (a wire.LabelStyleTable).

    Returns (node_class, way_class, mp_class, class_begin, class_styles); `order` is the dense rank of (layer, z_index) --
    compare_styled_entities with for_labels = true (styler.rs:246-272)."""
    from ..wire import CLASS_STYLE_DTYPE, NO_CLASS
    from .styler import KIND_NODE

    lists = []  # per class: (label style ids, layer, z)
    label_lists = _label_methods.label_lists

    def classes_of(kinds, tagkeys):
        out = np.full(len(kinds), NO_CLASS, dtype=np.uint32)
        for kind in np.unique(kinds):
            m = np.nonzero(kinds == kind)[0]
            inv, ls = label_lists(fb, zoom, int(kind), tagkeys[m], ltable)
            base = len(lists)
            lists.extend(ls)
            out[m] = (base + inv).astype(np.uint32)
        return out

    rd = fb.rd
    ntl = rd.nodes["tags_len"].astype(np.int64)
    ntk = np.where(ntl > 0, (rd.nodes["tags_off"].astype(np.int64) << 32) | ntl, 0)  # untagged nodes: one class
    node_class = classes_of(np.full(len(ntk), KIND_NODE, dtype=np.int64), ntk)
    way_class = classes_of(fb.way_kind, fb.way_tagkey)
    mp_class = classes_of(np.full(len(fb.mp_gid), KIND_MULTIPOLYGON, dtype=np.int64), fb.mp_tagkey)
    counts = np.array([len(l[0]) for l in lists], dtype=np.int64)
    class_begin = np.zeros(len(lists) + 1, dtype=np.uint32)
    class_begin[1:] = np.cumsum(counts)
    n = int(class_begin[-1])
    cs = np.zeros(n, dtype=CLASS_STYLE_DTYPE)
    if n:
        sid = np.concatenate([l[0] for l in lists])
        lay = np.concatenate([l[1] for l in lists])
        zi = np.concatenate([l[2] for l in lists])
        pair = np.stack([lay.astype(np.float64), zi], axis=1)
        _, order = np.unique(pair, axis=0, return_inverse=True)
        cs["style"] = sid
        cs["order"] = np.asarray(order).reshape(-1)
    return node_class, way_class, mp_class, class_begin, cs


class LabelListBuilder:
    """The label generations of a tile in the order the reference draws them (drawer.rs:106-119, 221-262):
    styled areas with for_labels=true (ways: text on the line, multipolygons: centred), then styled nodes.
    Oracle-side only: the CUDA library has no label pass yet."""

    def __init__(self, ts: TileStyler, icon_base_path: str | None):
        from ..wire import load_icon_rgba

        self.ts = ts
        self._load = load_icon_rgba
        self.icon_base_path = icon_base_path
        self.icon_ids: dict = {}
        self.icons: list = []
        self.texts = bytearray()
        self._text_off: dict = {}
        self._node_ent: dict = {}

    def _icon(self, name):
        if name is None:
            return -2
        i = self.icon_ids.get(name)
        if i is None:
            ic = self._load(os.path.join(self.icon_base_path, name)) if self.icon_base_path else None
            i = -1
            if ic is not None:
                i = len(self.icons)
                self.icons.append(ic)
            self.icon_ids[name] = i
        return i

    def _text(self, s: str):
        t = self._text_off.get(s)
        if t is None:
            b = s.encode("utf-8")
            t = (len(self.texts), len(b))
            self.texts += b
            self._text_off[s] = t
        return t

    def node_entity(self, n: int):
        from .styler import KIND_NODE

        e = self._node_ent.get(n)
        if e is None:
            rd = self.ts.reader
            e = (KIND_NODE, n, int(rd.nodes[n]["id"]), rd.node_tags(n))
            self._node_ent[n] = e
        return e

    def labels_abi(self, zoom, x, y, ltable):
        """The same label generations as `labels`, as osmr_label records (entity, label-style index) for the C ABI."""
        from ..wire import LABEL_DTYPE, OSMR_AREA_MULTIPOLYGON, OSMR_LABEL_NODE
        from .styler import KIND_MULTIPOLYGON, KIND_NODE

        ts = self.ts
        nodes, _, _ = ts.reader.get_entities_in_tile_with_neighbors(zoom, x, y)
        styled = ts.styled_areas(zoom, x, y, for_labels=True)
        styled_nodes = ts.styler.style_entities([self.node_entity(int(n)) for n in nodes], zoom, True)
        seq = list(styled) + list(styled_nodes)
        out = np.empty(len(seq), dtype=LABEL_DTYPE)
        for i, (ent, s) in enumerate(seq):
            flag = OSMR_LABEL_NODE if ent[0] == KIND_NODE else (OSMR_AREA_MULTIPOLYGON if ent[0] == KIND_MULTIPOLYGON else 0)
            out[i] = (ent[1] | flag, ltable.style_id(s))
        return out

    def labels(self, zoom, x, y):
        from .styler import KIND_MULTIPOLYGON, KIND_NODE

        ts = self.ts
        nodes, _, _ = ts.reader.get_entities_in_tile_with_neighbors(zoom, x, y)
        styled = ts.styled_areas(zoom, x, y, for_labels=True)
        styled_nodes = ts.styler.style_entities([self.node_entity(int(n)) for n in nodes], zoom, True)
        rows = []
        for ent, s in list(styled) + list(styled_nodes):
            kind = 2 if ent[0] == KIND_NODE else (1 if ent[0] == KIND_MULTIPOLYGON else 0)
            tstyle = s.text_style
            has_text = False
            toff = tlen = 0
            if tstyle is not None and tstyle.text in ent[3]:
                has_text = True
                toff, tlen = self._text(ent[3][tstyle.text])
            rows.append((
                kind, ent[1], self._icon(s.icon_image), 1 if tstyle is not None else 0,
                1 if (tstyle is not None and tstyle.font_size is not None) else 0, 1 if has_text else 0, toff, tlen,
                0 if (tstyle is None or tstyle.text_position is None) else (1 if tstyle.text_position == "center" else 2),
                tuple(tstyle.text_color) if (tstyle is not None and tstyle.text_color is not None) else (0, 0, 0),
                1 if (tstyle is not None and tstyle.text_color is not None) else 0,
                float(tstyle.font_size) if (tstyle is not None and tstyle.font_size is not None) else 0.0,
            ))
        return rows

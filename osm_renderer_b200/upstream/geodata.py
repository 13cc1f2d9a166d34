"""Upstream of the draw path: OSM XML -> geodata `.bin` -> per-tile entity candidates.

This is NOT part of the accelerated hot path (SURVEY.md section 8 marks it out of scope).  It exists
only because the reference's fixtures (`tests/osm/nano_moscow.osm`) must be turned into the exact
inputs the reference's `Drawer::draw_to_pixels` would see, and no Rust toolchain exists here.

Behaviour follows (file:line relative to /root/reference):
  * XML import ............ src/geodata/importer.rs:186-353, 517-536
  * multipolygon rings .... src/geodata/find_polygons.rs:32-196
  * `.bin` writer ......... src/geodata/saver.rs:21-226   (wire format: SURVEY.md appendix A.1)
  * `.bin` reader ......... src/geodata/reader.rs:60-180, 264-335
  * z18 tile of a node .... src/tile.rs:30-38, 88-101

The `.bin` image produced here is the *input wire format* of the C ABI (`osmr_set_geodata`).
"""
from __future__ import annotations

import math
import struct
import sys
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

MAX_ZOOM = 18
TILE_SIZE = 256

NODE_SIZE = 32
WAY_SIZE = 24
POLYGON_SIZE = 8
MULTIPOLYGON_SIZE = 24
TILE_REC_SIZE = 32


# ----------------------------------------------------------------------------------------------
# Projection used by the importer for the z18 tile index (src/tile.rs:88-101, 30-38).
# Python floats are IEEE doubles and math.tan/math.log are the platform libm, exactly what Rust
# calls on Linux, so this is the same arithmetic as the reference importer.
# ----------------------------------------------------------------------------------------------
_RADS_PER_DEG = math.pi / 180.0


def coords_to_xy(lat: float, lon: float, zoom: int) -> tuple[float, float]:
    lat_rad = lat * _RADS_PER_DEG
    lon_rad = lon * _RADS_PER_DEG
    x = lon_rad + math.pi
    y = math.pi - math.log(math.tan((math.pi / 4.0) + (lat_rad / 2.0)))
    dim = float(TILE_SIZE * (1 << zoom))
    return (x / (2.0 * math.pi)) * dim, (y / (2.0 * math.pi)) * dim


def _as_u32(v: float) -> int:
    # Rust `f64 as u32`: truncate toward zero, saturate, NaN -> 0.
    if v != v:
        return 0
    if v <= 0.0:
        return 0
    if v >= 4294967295.0:
        return 4294967295
    return int(v)


def coords_to_max_zoom_tile(lat: float, lon: float) -> tuple[int, int]:
    x, y = coords_to_xy(lat, lon, MAX_ZOOM)
    return _as_u32(x) // TILE_SIZE, _as_u32(y) // TILE_SIZE


def tile_to_max_zoom_tile_range(zoom: int, x: int, y: int) -> tuple[int, int, int, int]:
    """src/tile.rs:63-73 -> (min_x, max_x, min_y, max_y)."""
    mul = 1 << (MAX_ZOOM - zoom)
    min_x, min_y = x * mul, y * mul
    return min_x, min_x + mul - 1, min_y, min_y + mul - 1


# ----------------------------------------------------------------------------------------------
# Raw entities (importer.rs:484-546)
# ----------------------------------------------------------------------------------------------
@dataclass
class RawNode:
    global_id: int
    lat: float
    lon: float
    tags: dict = field(default_factory=dict)


@dataclass
class RawWay:
    global_id: int
    node_ids: list = field(default_factory=list)
    tags: dict = field(default_factory=dict)


@dataclass
class RawMultipolygon:
    global_id: int
    polygon_ids: list = field(default_factory=list)
    tags: dict = field(default_factory=dict)


class EntityStorages:
    def __init__(self):
        self.nodes: list[RawNode] = []
        self.node_g2l: dict[int, int] = {}
        self.ways: list[RawWay] = []
        self.way_g2l: dict[int, int] = {}
        self.polygons: list[list[int]] = []
        self.multipolygons: list[RawMultipolygon] = []

    def add_node(self, n: RawNode):
        self.node_g2l[n.global_id] = len(self.nodes)
        self.nodes.append(n)

    def add_way(self, w: RawWay):
        self.way_g2l[w.global_id] = len(self.ways)
        self.ways.append(w)


def _postprocess_node_refs(refs: list[int]) -> list[int]:
    """importer.rs:334-353 -- drop a node whose (cur, prev) pair (either direction) was seen before."""
    if not refs:
        return refs
    seen = set()
    out = [refs[0]]
    for idx in range(1, len(refs)):
        cur, prev = refs[idx], refs[idx - 1]
        if (cur, prev) not in seen and (prev, cur) not in seen:
            seen.add((cur, prev))
            out.append(cur)
    return out


def _f64_bits(v: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", v))[0]


def find_polygons_in_multipolygon(relation_id: int, segments: list) -> list[list[int]] | None:
    """find_polygons.rs:32-196.  `segments` = [(id1, pos1, id2, pos2, is_inner)], pos = (lat bits, lon bits)."""
    connections: dict = {}
    for idx, (id1, pos1, id2, pos2, is_inner) in enumerate(segments):
        connections.setdefault(pos1, []).append((pos2, idx, is_inner))
        connections.setdefault(pos2, []).append((pos1, idx, is_inner))

    available = [True] * len(segments)
    rings = []
    unmatched = len(segments)
    for start_idx in range(len(segments)):
        if not available[start_idx]:
            continue
        available[start_idx] = False
        id1, pos1, id2, pos2, inner = segments[start_idx]
        used_segments = [start_idx]
        used_vertices = {pos1, pos2}
        first_pos = pos1
        cur = pos2
        ok = False
        while True:
            nxt = None
            for (other, seg_idx, seg_inner) in connections.get(cur, ()):
                can_use = seg_inner == inner and available[seg_idx]
                is_dup = (other in used_vertices) and other != first_pos
                if can_use and not is_dup:
                    nxt = (other, seg_idx)
                    break
            if nxt is None:
                ok = False
                break
            other, seg_idx = nxt
            available[seg_idx] = False
            used_segments.append(seg_idx)
            used_vertices.add(other)
            if other == first_pos:
                ok = len(used_segments) >= 3
                break
            cur = other
        if not ok:
            print(
                f"Relation #{relation_id} is not a valid multipolygon (built {len(rings)} complete rings, "
                f"but {unmatched} segments are unmatched)",
                file=sys.stderr,
            )
            return None
        unmatched -= len(used_segments)
        rings.append(used_segments)

    polygons = []
    for ring in rings:
        poly: list[int] = []
        for i, seg_i in enumerate(ring):
            id1, _, id2, _, _ = segments[seg_i]
            if i == 0:
                poly.append(id1)
            last = poly[-1]
            poly.append(id2 if last == id1 else id1)
        polygons.append(poly)
    return polygons


def parse_osm_xml(path: str) -> EntityStorages:
    """importer.rs:186-299."""
    st = EntityStorages()
    for _, el in ET.iterparse(path, events=("end",)):
        tag = el.tag
        if tag == "node":
            n = RawNode(int(el.attrib["id"]), float(el.attrib["lat"]), float(el.attrib["lon"]))
            for sub in el:
                if sub.tag == "tag":
                    n.tags[sub.attrib["k"]] = sub.attrib["v"]
            st.add_node(n)
            el.clear()
        elif tag == "way":
            w = RawWay(int(el.attrib["id"]))
            for sub in el:
                if sub.tag == "tag":
                    w.tags[sub.attrib["k"]] = sub.attrib["v"]
                elif sub.tag == "nd":
                    r = st.node_g2l.get(int(sub.attrib["ref"]))
                    if r is not None:
                        w.node_ids.append(r)
            w.node_ids = _postprocess_node_refs(w.node_ids)
            st.add_way(w)
            el.clear()
        elif tag == "relation":
            gid = int(el.attrib["id"])
            tags: dict = {}
            way_refs = []
            for sub in el:
                if sub.tag == "tag":
                    tags[sub.attrib["k"]] = sub.attrib["v"]
                elif sub.tag == "member" and sub.attrib["type"] == "way":
                    r = st.way_g2l.get(int(sub.attrib["ref"]))
                    if r is not None:
                        way_refs.append((r, sub.attrib.get("role", "") == "inner"))
            if tags.get("type") == "multipolygon":
                segments = []
                for way_id, is_inner in way_refs:
                    way = st.ways[way_id]
                    for idx in range(1, len(way.node_ids)):
                        a, b = way.node_ids[idx - 1], way.node_ids[idx]
                        na, nb = st.nodes[a], st.nodes[b]
                        segments.append(
                            (a, (_f64_bits(na.lat), _f64_bits(na.lon)), b, (_f64_bits(nb.lat), _f64_bits(nb.lon)), is_inner)
                        )
                polys = find_polygons_in_multipolygon(gid, segments)
                if polys is not None:
                    mp = RawMultipolygon(gid, [], tags)
                    for p in polys:
                        mp.polygon_ids.append(len(st.polygons))
                        st.polygons.append(p)
                    st.multipolygons.append(mp)
            el.clear()
    return st


# ----------------------------------------------------------------------------------------------
# `.bin` writer (saver.rs:21-226)
# ----------------------------------------------------------------------------------------------
class _Buffered:
    def __init__(self):
        self.ints: list[int] = []
        self.str_off: dict[str, int] = {}
        self.strings = bytearray()

    def add_string(self, s: str) -> tuple[int, int]:
        b = s.encode("utf-8")
        off = self.str_off.get(s)
        if off is None:
            off = len(self.strings)
            self.str_off[s] = off
            self.strings += b
        return off, len(b)

    def refs(self, values) -> bytes:
        off = len(self.ints)
        self.ints.extend(values)
        return struct.pack("<II", off, len(self.ints) - off)

    def tags(self, tags: dict) -> bytes:
        kv = []
        for k in sorted(tags.keys()):  # BTreeMap order == byte order of UTF-8 == code-point order
            ko, kl = self.add_string(k)
            vo, vl = self.add_string(tags[k])
            kv.extend((ko, kl, vo, vl))
        return self.refs(kv)


def _tile_references(st: EntityStorages):
    """saver.rs:167-226: entity -> every z18 tile of the bbox of its nodes' z18 tiles."""
    node_tiles = [coords_to_max_zoom_tile(n.lat, n.lon) for n in st.nodes]
    refs: dict[tuple[int, int], list[set]] = {}

    def slot(xy):
        r = refs.get(xy)
        if r is None:
            r = [set(), set(), set()]
            refs[xy] = r
        return r

    for i, xy in enumerate(node_tiles):
        slot(xy)[0].add(i)

    def insert(node_ids, which, entity_id):
        if not node_ids:
            return
        xs = [node_tiles[n][0] for n in node_ids]
        ys = [node_tiles[n][1] for n in node_ids]
        for x in range(min(xs), max(xs) + 1):
            for y in range(min(ys), max(ys) + 1):
                slot((x, y))[which].add(entity_id)

    for i, w in enumerate(st.ways):
        insert(w.node_ids, 1, i)
    for i, mp in enumerate(st.multipolygons):
        ids = [n for pid in mp.polygon_ids for n in st.polygons[pid]]
        insert(ids, 2, i)
    return refs


def save_to_internal_format(st: EntityStorages) -> bytes:
    buf = _Buffered()
    out = bytearray()
    out += struct.pack("<I", len(st.nodes))
    for n in st.nodes:
        out += struct.pack("<Qdd", n.global_id, n.lat, n.lon)
        out += buf.tags(n.tags)
    out += struct.pack("<I", len(st.ways))
    for w in st.ways:
        out += struct.pack("<Q", w.global_id)
        out += buf.refs(w.node_ids)
        out += buf.tags(w.tags)
    out += struct.pack("<I", len(st.polygons))
    for p in st.polygons:
        out += buf.refs(p)
    out += struct.pack("<I", len(st.multipolygons))
    for mp in st.multipolygons:
        out += struct.pack("<Q", mp.global_id)
        out += buf.refs(mp.polygon_ids)
        out += buf.tags(mp.tags)
    refs = _tile_references(st)
    out += struct.pack("<I", len(refs))
    for (x, y) in sorted(refs.keys()):
        r = refs[(x, y)]
        out += struct.pack("<II", x, y)
        out += buf.refs(sorted(r[0]))
        out += buf.refs(sorted(r[1]))
        out += buf.refs(sorted(r[2]))
    out += struct.pack("<I", len(buf.ints))
    out += np.asarray(buf.ints, dtype="<u4").tobytes()
    out += bytes(buf.strings)
    return bytes(out)


def import_osm(path: str) -> bytes:
    """importer.rs:19-43 (`import`) returning the `.bin` image instead of writing a file."""
    return save_to_internal_format(parse_osm_xml(path))


# ----------------------------------------------------------------------------------------------
# `.bin` reader (reader.rs)
# ----------------------------------------------------------------------------------------------
_NODE_DT = np.dtype([("id", "<u8"), ("lat", "<f8"), ("lon", "<f8"), ("tags_off", "<u4"), ("tags_len", "<u4")])
_WAY_DT = np.dtype([("id", "<u8"), ("off", "<u4"), ("len", "<u4"), ("tags_off", "<u4"), ("tags_len", "<u4")])
_POLY_DT = np.dtype([("off", "<u4"), ("len", "<u4")])
_TILE_DT = np.dtype(
    [("x", "<u4"), ("y", "<u4"), ("n_off", "<u4"), ("n_len", "<u4"), ("w_off", "<u4"), ("w_len", "<u4"), ("m_off", "<u4"), ("m_len", "<u4")]
)


class GeodataReader:
    """Zero-copy numpy views over a `.bin` image (reader.rs:264-335)."""

    def __init__(self, data: bytes):
        self.data = data
        pos = 0

        def table(dt):
            nonlocal pos
            (cnt,) = struct.unpack_from("<I", data, pos)
            pos += 4
            arr = np.frombuffer(data, dtype=dt, count=cnt, offset=pos)
            pos += cnt * dt.itemsize
            return arr

        self.nodes = table(_NODE_DT)
        self.ways = table(_WAY_DT)
        self.polygons = table(_POLY_DT)
        self.multipolygons = table(_WAY_DT)
        self.tiles = table(_TILE_DT)
        (n_ints,) = struct.unpack_from("<I", data, pos)
        pos += 4
        self.ints = np.frombuffer(data, dtype="<u4", count=n_ints, offset=pos)
        pos += 4 * n_ints
        self.strings = data[pos:]
        self._tile_key = (self.tiles["x"].astype(np.uint64) << np.uint64(32)) | self.tiles["y"].astype(np.uint64)
        self._tags_cache: dict = {}

    @classmethod
    def load(cls, path: str) -> "GeodataReader":
        with open(path, "rb") as f:
            return cls(f.read())

    # -- tags ----------------------------------------------------------------------------------
    def tags_of(self, tags_off: int, tags_len: int) -> dict:
        key = (tags_off, tags_len)
        t = self._tags_cache.get(key)
        if t is None:
            kv = self.ints[tags_off : tags_off + tags_len]
            s = self.strings
            t = {}
            for i in range(0, len(kv), 4):
                ko, kl, vo, vl = (int(v) for v in kv[i : i + 4])
                t[s[ko : ko + kl].decode("utf-8")] = s[vo : vo + vl].decode("utf-8")
            self._tags_cache[key] = t
        return t

    def way_tags(self, idx: int) -> dict:
        w = self.ways[idx]
        return self.tags_of(int(w["tags_off"]), int(w["tags_len"]))

    def multipolygon_tags(self, idx: int) -> dict:
        w = self.multipolygons[idx]
        return self.tags_of(int(w["tags_off"]), int(w["tags_len"]))

    def node_tags(self, idx: int) -> dict:
        n = self.nodes[idx]
        return self.tags_of(int(n["tags_off"]), int(n["tags_len"]))

    # -- geometry ------------------------------------------------------------------------------
    def way_node_ids(self, idx: int) -> np.ndarray:
        w = self.ways[idx]
        return self.ints[int(w["off"]) : int(w["off"]) + int(w["len"])]

    def way_is_closed(self, idx: int) -> bool:
        """reader.rs:474-483."""
        ids = self.way_node_ids(idx)
        if len(ids) <= 2:
            return False
        a, b = self.nodes[int(ids[0])], self.nodes[int(ids[-1])]
        return a["lat"] == b["lat"] and a["lon"] == b["lon"]

    def multipolygon_polygon_ids(self, idx: int) -> np.ndarray:
        m = self.multipolygons[idx]
        return self.ints[int(m["off"]) : int(m["off"]) + int(m["len"])]

    def polygon_node_ids(self, idx: int) -> np.ndarray:
        p = self.polygons[idx]
        return self.ints[int(p["off"]) : int(p["off"]) + int(p["len"])]

    # -- tile query ------------------------------------------------------------------------------
    def get_entities_in_tile_with_neighbors(self, zoom: int, x: int, y: int):
        """reader.rs:60-100: union over the 3x3 neighbourhood, sorted + deduped local ids.

        Returns (node_ids, way_ids, multipolygon_ids) as sorted uint32 arrays; multipolygons with zero
        polygons are dropped (reader.rs:86-93).  The per-column binary search of the reference
        (reader.rs:102-180) enumerates exactly the index records with x18,y18 inside each neighbour's
        z18 range, which is what the vectorised range test below selects.
        """
        tx, ty = self.tiles["x"], self.tiles["y"]
        sel = np.zeros(len(self.tiles), dtype=bool)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                nx, ny = (x + dx) & 0xFFFFFFFF, (y + dy) & 0xFFFFFFFF
                mul = 1 << (MAX_ZOOM - zoom)
                # u32 arithmetic of tile.rs:64 wraps; a wrapped neighbour selects nothing in practice.
                min_x, min_y = (nx * mul) & 0xFFFFFFFF, (ny * mul) & 0xFFFFFFFF
                max_x, max_y = min_x + mul - 1, min_y + mul - 1
                sel |= (tx >= min_x) & (tx <= max_x) & (ty >= min_y) & (ty <= max_y)
        recs = self.tiles[sel]

        def gather(off_name, len_name):
            parts = [self.ints[int(o) : int(o) + int(l)] for o, l in zip(recs[off_name], recs[len_name]) if l]
            if not parts:
                return np.zeros(0, dtype=np.uint32)
            return np.unique(np.concatenate(parts)).astype(np.uint32)

        nodes = gather("n_off", "n_len")
        ways = gather("w_off", "w_len")
        mps = gather("m_off", "m_len")
        if len(mps):
            mps = mps[self.multipolygons["len"][mps] > 0]
        return nodes, ways, mps

"""Seeded synthetic "metro area" geodata image for the throughput configs of BASELINE.json (C2-C5).

Not part of the accelerated path: this only manufactures a `.bin` image in the reference's on-disk format
(src/geodata/saver.rs:21-165) so the benchmark has a city-sized dataset without any
network access.  Everything is drawn from numpy's PCG64 seeded with 0xB20005A1, so every run (here and on the
GPU box) produces the same bytes.

Content of the default metro (z14 block x in [9888,9919], y in [5104,5135] around the reference's fixture tile):
  * jittered street grid every 120-250 m, a way runs along 1-4 blocks of its grid line with 2-10 nodes,
    highway=residential/tertiary/secondary/primary (p = .70/.15/.10/.05), 30 % of the streets carry one of 64 names
    (5-18 characters, like real street names)
  * 4-20 rotated building footprints per block (building=yes, closed ways of 5-13 nodes: rectangles, L-, U- and
    cross-shaped outlines, sides 8-40 m)
    [round 1 used `profile="r1"`: one way per block side, 3-8 rectangular footprints, "Synthetic Street N" names]
  * 6 % of the blocks carry a park / residential-landuse / water polygon of 16-48 nodes
  * 1 % of the blocks carry a multipolygon (natural=water outer ring, 1-3 inner rings)
  * one meandering river (waterway=river, 400 nodes) and one railway (railway=rail, 300 nodes, dashed in both
    stylesheets)
"""
from __future__ import annotations

import math
import struct

import numpy as np

SEED = 0xB20005A1
MAX_ZOOM = 18


def _px18_to_latlon(px, py):
    """inverse of src/tile.rs:88-101 at zoom 18 (only used to *place* synthetic nodes)."""
    dim = 256.0 * (1 << MAX_ZOOM)
    lon = (px / dim) * 360.0 - 180.0
    n = math.pi - 2.0 * math.pi * (py / dim)
    lat = np.degrees(np.arctan(np.sinh(n)))
    return lat, lon


def _latlon_to_tile18(lat, lon):
    """src/tile.rs:30-38 vectorised (numpy libm; the index only selects candidates, exactness is irrelevant)."""
    lat_rad = lat * (math.pi / 180.0)
    lon_rad = lon * (math.pi / 180.0)
    x = lon_rad + math.pi
    y = math.pi - np.log(np.tan((math.pi / 4.0) + (lat_rad / 2.0)))
    dim = 256.0 * (1 << MAX_ZOOM)
    fx = (x / (2.0 * math.pi)) * dim
    fy = (y / (2.0 * math.pi)) * dim
    return (fx.astype(np.int64) // 256).astype(np.uint32), (fy.astype(np.int64) // 256).astype(np.uint32)


class _Builder:
    def __init__(self):
        self.px = []  # node coordinates in z18 pixels
        self.py = []
        self.n_nodes = 0
        self.way_nodes = []  # list of int arrays (local node ids)
        self.way_tags = []  # tagset id per way
        self.polys = []  # rings of multipolygons
        self.mps = []  # (list of polygon ids, tagset id)
        self.tagsets: dict = {}
        self.tagset_list = []

    def tagset(self, tags: dict) -> int:
        key = tuple(sorted(tags.items()))
        i = self.tagsets.get(key)
        if i is None:
            i = len(self.tagset_list)
            self.tagsets[key] = i
            self.tagset_list.append(dict(key))
        return i

    def add_nodes(self, xs, ys) -> np.ndarray:
        xs = np.asarray(xs, dtype=np.float64).ravel()
        ys = np.asarray(ys, dtype=np.float64).ravel()
        ids = np.arange(self.n_nodes, self.n_nodes + len(xs), dtype=np.int64)
        self.px.append(xs)
        self.py.append(ys)
        self.n_nodes += len(xs)
        return ids

    def add_way(self, node_ids, tags: dict):
        self.way_nodes.append(np.asarray(node_ids, dtype=np.int64))
        self.way_tags.append(self.tagset(tags))

    def add_ways_fixed(self, ids2d: np.ndarray, tagset_ids: np.ndarray):
        """many ways with the same node count (rows of ids2d)"""
        for row, ts in zip(ids2d, tagset_ids):
            self.way_nodes.append(row)
            self.way_tags.append(int(ts))


_STREET_WORDS = ["Oak", "Elm", "Main", "Mill", "Park", "Lake", "Hill", "High", "Church", "Station", "Market", "Bridge", "Garden",
                 "Linden", "Maple", "Cedar", "Victoria", "Harbour", "Orchard", "Meadow", "Castle", "Abbey"]
_STREET_KINDS = ["St", "Ave", "Rd", "Lane", "Street", "Way", "Close", "Drive"]


def street_names() -> list:
    """64 deterministic names of 5-18 characters (no RNG: the same list everywhere)."""
    out = []
    for i in range(64):
        out.append(f"{_STREET_WORDS[(i * 7) % len(_STREET_WORDS)]} {_STREET_KINDS[(i * 3 + i // 8) % len(_STREET_KINDS)]}")
    return out


# building outlines on the unit square [-1, 1]^2, counter-clockwise, closing node added by the generator (5-13 nodes)
_OUTLINES = [
    np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64),  # rectangle: 5 nodes
    np.array([[-1, -1], [1, -1], [1, 0], [0, 0], [0, 1], [-1, 1]], dtype=np.float64),  # L: 7 nodes
    np.array([[-1, -1], [1, -1], [1, 1], [0.4, 1], [0.4, -0.2], [-0.4, -0.2], [-0.4, 1], [-1, 1]], dtype=np.float64),  # U: 9 nodes
    np.array([[-0.4, -1], [0.4, -1], [0.4, -0.4], [1, -0.4], [1, 0.4], [0.4, 0.4], [0.4, 1], [-0.4, 1], [-0.4, 0.4], [-1, 0.4],
              [-1, -0.4], [-0.4, -0.4]], dtype=np.float64),  # cross: 13 nodes
]


def make_metro(seed: int = SEED, zoom: int = 14, x0: int = 9888, y0: int = 5104, n: int = 32, buildings=None,
               extra_footprints: int = 0, coastline_nodes: int = 0, profile: str = "r2") -> bytes:
    """Returns the `.bin` image.  extra_footprints / coastline_nodes add the C5 stress content.
    profile "r2" is SURVEY.md 8(d)'s C2 as written; "r1" reproduces the round-1 dataset (sparser, see the module docstring)."""
    if buildings is None:
        buildings = (4, 20) if profile == "r2" else (3, 8)
    rng = np.random.default_rng(seed)
    mul = 256 << (MAX_ZOOM - zoom)  # z18 pixels per tile of `zoom`
    X0, Y0 = x0 * mul, y0 * mul
    W = n * mul
    b = _Builder()
    m_per_px = 0.336  # z18 pixel in metres at the latitude of the fixture

    # ---- street grid ------------------------------------------------------------------------------------
    def grid_lines():
        pos = [0.0]
        while pos[-1] < W:
            pos.append(pos[-1] + rng.uniform(120.0, 250.0) / m_per_px)
        return np.array(pos[:-1]) if pos[-1] > W else np.array(pos)

    gx = grid_lines()
    gy = grid_lines()
    nx, ny = len(gx), len(gy)
    names = street_names() if profile == "r2" else [f"Synthetic Street {i}" for i in range(64)]
    classes = ["residential", "tertiary", "secondary", "primary"]

    def street_tags(cls_i, name_i):
        t = {"highway": classes[cls_i]}
        if name_i >= 0:
            t["name"] = names[name_i]
        return t

    # horizontal streets: one class/name per grid line, one way per block side
    for horizontal in (True, False):
        lines = gy if horizontal else gx
        cross = gx if horizontal else gy
        for li, c in enumerate(lines):
            cls_i = int(rng.choice(4, p=[0.70, 0.15, 0.10, 0.05]))
            name_i = int(rng.integers(0, 64)) if rng.random() < 0.30 else -1
            ts = b.tagset(street_tags(cls_i, name_i))
            if profile == "r2":  # a way runs along 1-4 blocks, every block adds 1-3 nodes: 2-10 nodes per way
                j = 0
                while j < len(cross) - 1:
                    span = int(min(rng.integers(1, 5), len(cross) - 1 - j))
                    per_block = int(rng.integers(0, 3 if span <= 3 else 2))  # interior nodes per block
                    parts = [np.linspace(cross[j + q], cross[j + q + 1], per_block + 2)[:-1] for q in range(span)]
                    t = np.concatenate(parts + [np.array([cross[j + span]])])
                    off = rng.normal(0.0, 4.0, size=len(t))
                    off[0] = off[-1] = 0.0
                    if horizontal:
                        ids = b.add_nodes(X0 + t, Y0 + c + off)
                    else:
                        ids = b.add_nodes(X0 + c + off, Y0 + t)
                    b.way_nodes.append(ids)
                    b.way_tags.append(ts)
                    j += span
                continue
            k = rng.integers(2, 7, size=len(cross) - 1)  # nodes per way
            for j in range(len(cross) - 1):
                t = np.linspace(cross[j], cross[j + 1], int(k[j]))
                off = rng.normal(0.0, 4.0, size=len(t))
                off[0] = off[-1] = 0.0
                if horizontal:
                    ids = b.add_nodes(X0 + t, Y0 + c + off)
                else:
                    ids = b.add_nodes(X0 + c + off, Y0 + t)
                b.way_nodes.append(ids)
                b.way_tags.append(ts)

    # ---- per-block content -------------------------------------------------------------------------------
    ts_building = b.tagset({"building": "yes"})
    ts_park = b.tagset({"leisure": "park"})
    ts_resid = b.tagset({"landuse": "residential"})
    ts_water = b.tagset({"natural": "water"})
    bx0, bx1 = gx[:-1], gx[1:]
    by0, by1 = gy[:-1], gy[1:]
    BX0, BY0 = np.meshgrid(bx0, by0, indexing="ij")
    BX1, BY1 = np.meshgrid(bx1, by1, indexing="ij")
    BX0, BY0, BX1, BY1 = BX0.ravel(), BY0.ravel(), BX1.ravel(), BY1.ravel()
    n_blocks = len(BX0)
    nb = rng.integers(buildings[0], buildings[1] + 1, size=n_blocks)
    blk = np.repeat(np.arange(n_blocks), nb)
    nbt = len(blk)
    margin = 30.0
    cx = BX0[blk] + margin + rng.random(nbt) * np.maximum(BX1[blk] - BX0[blk] - 2 * margin, 1.0)
    cy = BY0[blk] + margin + rng.random(nbt) * np.maximum(BY1[blk] - BY0[blk] - 2 * margin, 1.0)
    hw = rng.uniform(8.0, 40.0, nbt) / m_per_px / 2.0
    hh = rng.uniform(8.0, 40.0, nbt) / m_per_px / 2.0
    ang = rng.uniform(-15.0, 15.0, nbt) * math.pi / 180.0
    ca, sa = np.cos(ang), np.sin(ang)
    corners = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64)
    shape = rng.choice(4, size=nbt, p=[0.60, 0.22, 0.12, 0.06]) if profile == "r2" else np.zeros(nbt, dtype=np.int64)
    for si, outline in enumerate(_OUTLINES):
        m = np.nonzero(shape == si)[0]
        if len(m) == 0:
            continue
        ux, uy = outline[None, :, 0] * hw[m, None], outline[None, :, 1] * hh[m, None]
        rx = cx[m, None] + ux * ca[m, None] - uy * sa[m, None]
        ry = cy[m, None] + ux * sa[m, None] + uy * ca[m, None]
        ids = b.add_nodes(X0 + rx, Y0 + ry).reshape(len(m), len(outline))
        b.add_ways_fixed(np.concatenate([ids, ids[:, :1]], axis=1), np.full(len(m), ts_building))

    def ring(cxv, cyv, rad, k):
        th = np.sort(rng.random(k)) * 2.0 * math.pi
        r = rad * rng.uniform(0.6, 1.0, k)
        return cxv + r * np.cos(th), cyv + r * np.sin(th)

    special = rng.random(n_blocks)
    for bi in np.nonzero(special < 0.06)[0]:
        k = int(rng.integers(16, 49))
        rad = 0.45 * min(BX1[bi] - BX0[bi], BY1[bi] - BY0[bi])
        xs, ys = ring(0.5 * (BX0[bi] + BX1[bi]), 0.5 * (BY0[bi] + BY1[bi]), rad, k)
        ids = b.add_nodes(X0 + xs, Y0 + ys)
        b.way_nodes.append(np.concatenate([ids, ids[:1]]))
        b.way_tags.append(int(rng.choice([ts_park, ts_resid, ts_water])))
    for bi in np.nonzero((special >= 0.06) & (special < 0.07))[0]:
        rad = 0.45 * min(BX1[bi] - BX0[bi], BY1[bi] - BY0[bi])
        ccx, ccy = 0.5 * (BX0[bi] + BX1[bi]), 0.5 * (BY0[bi] + BY1[bi])
        xs, ys = ring(ccx, ccy, rad, 24)
        ids = b.add_nodes(X0 + xs, Y0 + ys)
        pids = [len(b.polys)]
        b.polys.append(np.concatenate([ids, ids[:1]]))
        for _ in range(int(rng.integers(1, 4))):
            ox, oy = rng.uniform(-0.3, 0.3, 2) * rad
            xs, ys = ring(ccx + ox, ccy + oy, 0.15 * rad, 8)
            ids = b.add_nodes(X0 + xs, Y0 + ys)
            pids.append(len(b.polys))
            b.polys.append(np.concatenate([ids, ids[:1]]))
        b.mps.append((pids, ts_water))

    # ---- river and railway ----------------------------------------------------------------------------------
    t = np.linspace(0.0, 1.0, 400)
    ids = b.add_nodes(X0 + t * W, Y0 + W * (0.55 + 0.12 * np.sin(t * 9.0) + 0.03 * np.sin(t * 41.0)))
    b.add_way(ids, {"waterway": "river", "name": "Synthetic River"})
    t = np.linspace(0.0, 1.0, 300)
    ids = b.add_nodes(X0 + W * (0.30 + 0.25 * t + 0.02 * np.sin(t * 17.0)), Y0 + t * W)
    b.add_way(ids, {"railway": "rail"})

    # ---- C5 stress: dense footprints + a long coastline multipolygon -------------------------------------------
    if extra_footprints:
        k = extra_footprints
        cx = rng.random(k) * W
        cy = rng.random(k) * W
        hw = rng.uniform(8.0, 40.0, k) / m_per_px / 2.0
        hh = rng.uniform(8.0, 40.0, k) / m_per_px / 2.0
        rx = cx[:, None] + corners[None, :, 0] * hw[:, None]
        ry = cy[:, None] + corners[None, :, 1] * hh[:, None]
        ids = b.add_nodes(X0 + rx, Y0 + ry).reshape(k, 4)
        b.add_ways_fixed(np.concatenate([ids, ids[:, :1]], axis=1), np.full(k, ts_building))
    if coastline_nodes:
        k = coastline_nodes
        half = k // 2
        t = np.linspace(0.0, 1.0, half)
        noise = np.cumsum(rng.normal(0.0, 6.0, half))
        noise -= np.linspace(0.0, noise[-1], half)
        xs = np.concatenate([t * W, (1.0 - t) * W])
        ys = np.concatenate([t * W + noise, np.clip((1.0 - t) * W + 0.35 * W + noise[::-1], 0, None)])
        ids = b.add_nodes(X0 + xs, Y0 + ys)
        pids = [len(b.polys)]
        b.polys.append(np.concatenate([ids, ids[:1]]))
        for _ in range(200):
            kk = int(rng.integers(20, 201))
            u = rng.random()
            ccx, ccy = u * W, u * W + 0.17 * W + rng.uniform(-0.1, 0.1) * W
            xs, ys = ring(ccx, ccy, rng.uniform(80.0, 400.0), kk)
            ids = b.add_nodes(X0 + xs, Y0 + ys)
            pids.append(len(b.polys))
            b.polys.append(np.concatenate([ids, ids[:1]]))
        b.mps.append((pids, b.tagset({"natural": "coastline"})))

    return _serialise(b)


def _serialise(b: _Builder, with_index: bool = True) -> bytes:
    """with_index=False leaves the z18 tile index empty (the draw path never reads it; tests with planet-sized
    ways would otherwise enumerate billions of index cells)."""
    px = np.concatenate(b.px)
    py = np.concatenate(b.py)
    lat, lon = _px18_to_latlon(px, py)
    n_nodes = len(px)
    n_ways = len(b.way_nodes)

    # strings + tag kv lists (shared per distinct tag set; the format only stores (off, len) pairs)
    strings = bytearray()
    str_off: dict = {}

    def add_string(s: str):
        o = str_off.get(s)
        bs = s.encode("utf-8")
        if o is None:
            o = len(strings)
            str_off[s] = o
            strings.extend(bs)
        return o, len(bs)

    ints_parts = []
    n_ints = 0

    def push(arr) -> tuple[int, int]:
        nonlocal n_ints
        arr = np.asarray(arr, dtype=np.uint32)
        off = n_ints
        ints_parts.append(arr)
        n_ints += len(arr)
        return off, len(arr)

    ts_ref = []
    for tags in b.tagset_list:
        kv = []
        for k in sorted(tags):
            ko, kl = add_string(k)
            vo, vl = add_string(tags[k])
            kv += [ko, kl, vo, vl]
        ts_ref.append(push(kv))

    nodes = np.zeros(n_nodes, dtype=[("id", "<u8"), ("lat", "<f8"), ("lon", "<f8"), ("to", "<u4"), ("tl", "<u4")])
    nodes["id"] = np.arange(1, n_nodes + 1)
    nodes["lat"] = lat
    nodes["lon"] = lon

    way_len = np.array([len(w) for w in b.way_nodes], dtype=np.int64)
    way_off0 = n_ints
    all_way_nodes = np.concatenate(b.way_nodes) if n_ways else np.zeros(0, dtype=np.int64)
    push(all_way_nodes)
    way_off = way_off0 + np.concatenate([[0], np.cumsum(way_len)[:-1]]) if n_ways else np.zeros(0, dtype=np.int64)
    ways = np.zeros(n_ways, dtype=[("id", "<u8"), ("o", "<u4"), ("l", "<u4"), ("to", "<u4"), ("tl", "<u4")])
    ways["id"] = np.arange(1, n_ways + 1)
    ways["o"] = way_off
    ways["l"] = way_len
    wt = np.asarray(b.way_tags, dtype=np.int64)
    tsr = np.asarray(ts_ref, dtype=np.int64).reshape(-1, 2)
    ways["to"] = tsr[wt, 0]
    ways["tl"] = tsr[wt, 1]

    polys = np.zeros(len(b.polys), dtype=[("o", "<u4"), ("l", "<u4")])
    for i, p in enumerate(b.polys):
        polys[i] = push(p)
    mps = np.zeros(len(b.mps), dtype=[("id", "<u8"), ("o", "<u4"), ("l", "<u4"), ("to", "<u4"), ("tl", "<u4")])
    for i, (pids, ts) in enumerate(b.mps):
        o, l = push(pids)
        mps[i] = (n_ways + 1 + i, o, l, ts_ref[ts][0], ts_ref[ts][1])

    # ---- z18 tile index (saver.rs:167-226) ----
    tx, ty = _latlon_to_tile18(lat, lon)
    if not with_index:
        tx, ty = tx[:0], ty[:0]
    ent_x, ent_y, ent_id, ent_kind = [tx], [ty], [np.arange(len(tx), dtype=np.int64)], [np.zeros(len(tx), dtype=np.int8)]

    def add_bbox_entries(x0, x1, y0, y1, ids, kind):
        wx = (x1 - x0 + 1).astype(np.int64)
        wy = (y1 - y0 + 1).astype(np.int64)
        cnt = wx * wy
        if int(cnt.sum()) > 200_000_000:
            raise MemoryError("tile index would need %d cells; build this image with with_index=False" % int(cnt.sum()))
        rep = np.repeat(np.arange(len(ids)), cnt)
        start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        local = np.arange(cnt.sum()) - start[rep]
        ent_x.append((x0[rep] + local // wy[rep]).astype(np.uint32))
        ent_y.append((y0[rep] + local % wy[rep]).astype(np.uint32))
        ent_id.append(ids[rep])
        ent_kind.append(np.full(len(rep), kind, dtype=np.int8))

    if n_ways and with_index:
        seg = np.repeat(np.arange(n_ways), way_len)
        wx0 = np.full(n_ways, 2**32 - 1, dtype=np.int64)
        wx1 = np.zeros(n_ways, dtype=np.int64)
        wy0 = wx0.copy()
        wy1 = wx1.copy()
        nx_ = tx[all_way_nodes].astype(np.int64)
        ny_ = ty[all_way_nodes].astype(np.int64)
        np.minimum.at(wx0, seg, nx_)
        np.maximum.at(wx1, seg, nx_)
        np.minimum.at(wy0, seg, ny_)
        np.maximum.at(wy1, seg, ny_)
        add_bbox_entries(wx0, wx1, wy0, wy1, np.arange(n_ways, dtype=np.int64), 1)
    for i, (pids, _) in enumerate(b.mps if with_index else []):
        nn = np.concatenate([b.polys[p] for p in pids])
        add_bbox_entries(
            np.array([tx[nn].min()], dtype=np.int64), np.array([tx[nn].max()], dtype=np.int64),
            np.array([ty[nn].min()], dtype=np.int64), np.array([ty[nn].max()], dtype=np.int64),
            np.array([i], dtype=np.int64), 2,
        )
    ex = np.concatenate(ent_x).astype(np.int64)
    ey = np.concatenate(ent_y).astype(np.int64)
    eid = np.concatenate(ent_id)
    ek = np.concatenate(ent_kind).astype(np.int64)
    order = np.lexsort((eid, ek, ey, ex))
    ex, ey, eid, ek = ex[order], ey[order], eid[order], ek[order]
    key = (ex << 32) | ey
    tile_start = np.concatenate([[0], np.nonzero(np.diff(key))[0] + 1]) if len(key) else np.zeros(0, dtype=np.int64)
    n_tiles = len(tile_start)
    tile_end = np.concatenate([tile_start[1:], [len(key)]]).astype(np.int64) if n_tiles else tile_start
    idx_off = n_ints
    push(eid)
    tiles = np.zeros(n_tiles, dtype=[("x", "<u4"), ("y", "<u4"), ("no", "<u4"), ("nl", "<u4"), ("wo", "<u4"), ("wl", "<u4"), ("mo", "<u4"), ("ml", "<u4")])
    tiles["x"] = ex[tile_start]
    tiles["y"] = ey[tile_start]
    # inside one tile the entries are sorted by kind then id: three contiguous runs
    tile_of = np.repeat(np.arange(n_tiles), tile_end - tile_start)
    for kind, (o_name, l_name) in enumerate((("no", "nl"), ("wo", "wl"), ("mo", "ml"))):
        cnt = np.bincount(tile_of[ek == kind], minlength=n_tiles)
        tiles[l_name] = cnt
    tiles["no"] = idx_off + tile_start
    tiles["wo"] = tiles["no"] + tiles["nl"]
    tiles["mo"] = tiles["wo"] + tiles["wl"]

    ints = np.concatenate(ints_parts).astype("<u4") if ints_parts else np.zeros(0, dtype="<u4")
    out = bytearray()
    out += struct.pack("<I", n_nodes) + nodes.tobytes()
    out += struct.pack("<I", n_ways) + ways.tobytes()
    out += struct.pack("<I", len(polys)) + polys.tobytes()
    out += struct.pack("<I", len(mps)) + mps.tobytes()
    out += struct.pack("<I", n_tiles) + tiles.tobytes()
    out += struct.pack("<I", len(ints)) + ints.tobytes()
    out += bytes(strings)
    return bytes(out)

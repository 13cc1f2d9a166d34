// osmr_labels_dev.cuh -- label LAYOUT on the device (the tile-dependent half of labeler.rs / labelable.rs / font/text_placer.rs).
//
// Round 1 laid every label out on the host, per tile and per call (osmr_labels_host.hpp).  What the reference computes for a
// label splits into a part that depends only on (entity, style, zoom, scale) and a part that depends on the tile:
//   * tile-INdependent: the text of the entity (tag lookup), its glyph ids, advances and kerning (font tables), the glyph
//     outlines, and -- for text along a way -- everything that is a function of the way's INTEGER pixel differences: segment
//     lengths, the fit test, the segment a glyph lands on, the ratio along it and the glyph's rotation (atan2 / sin / cos of
//     integer differences).  The integer pixels of a node in two tiles of one zoom differ by an exact multiple of 256 * scale
//     (tile.rs:88-106: factor * 2^(8+z) is exact, the subtraction of the tile origin is exact because both operands are
//     multiples of the minuend's ulp, `* scale` is exact for scale in {1, 2, 4, 8}), so the differences are tile-invariant.
//   * tile-DEPENDENT (in the last bits): every f64 that has the tile-relative position added to it -- the glyph origin
//     wx = x_i + dx * ratio, every transformed outline point wx + (...), the flatness test of the quadratic curves, the
//     polylabel anchor of an area (labelable.rs:125-204 runs on tile-relative f64 coordinates).  These are IEEE basic
//     operations (+, -, *, /, sqrt, comparisons), which the device reproduces bit for bit (-fmad=false).
// So the host keeps only what needs strings, font tables or the platform libm -- as RESIDENT tables built once per dataset /
// font / style table (text runs, glyph outlines) and once per zoom (sin / cos of every named way's segment directions, from
// glibc like the reference's) -- and the per-call, per-tile work below runs on the device:
//   label_select_kernel   which label generations of a tile can draw or collide at all (icon, or text that exists)
//   label_layout_kernel   anchor (node position or polylabel), icon rectangle, glyph placement along the way / in wrapped rows
//   label_vfill / vline / curve / scan kernels   glyph outlines -> the reference's Rasterizer::draw_line call stream (quadratic
//                         curves flattened by rasterizer.rs:86-107's recursive midpoint rule): counted, scanned, written
//   label_finish_kernel   per label: segment range, pixel bbox, coverage storage, the work list of label_cover_kernel
// The flatness rule compares platform-libm hypot values; the device decides it with sqrt whenever the two sides differ by more
// than 1e-12 relative (the host does the same, osmr_labels_host.hpp flat_enough) and raises LCNT_FALLBACK on a near tie: the
// call is then laid out by the host path, which asks glibc.  Scales that are not a power of two also take the host path.
#pragma once
#include "osmr_kernels.cuh"

namespace osmr {

struct DevLabelStyle {  // osmr_label_style with the text key interned
    int icon;           // -2 none, -1 failed to load, >= 0 label icon index
    unsigned flags;     // OSMR_LSTYLE_*
    int key;            // index of the distinct text key, -1 none
    unsigned rgb;       // 0x00BBGGRR (0 when TextStyle.text_color is None: black, text_placer.rs:52-55)
    unsigned text_position;
    unsigned pad;
    double font_size;
};
struct DevGlyphRec {  // one character of a text run
    int slot;         // index into glyph_vbegin (-1: the glyph has no outline)
    int advance;      // hmtx advance width, font units
    int kern;         // kern(previous glyph, this glyph), font units; 0 for the first
    unsigned ws;      // char::is_whitespace
};
struct DevVertex {  // stb_truetype vertex
    short x, y, cx, cy;
    int type;  // 1 move, 2 line, 3 curve
};

enum {
    LCNT_PLACES = 0,    // GlyphPlace entries handed out
    LCNT_SEGS = 1,      // segments handed out
    LCNT_ROWRECS = 2,   // (label, row) work items handed out (multiple of 32)
    LCNT_CELLS_LO = 4,  // 64-bit (4,5): coverage cells handed out
    LCNT_OVERFLOW = 6,  // bit0 places, bit1 segments, bit2 rows, bit3 cells, bit4 polylabel rings, bit5 polylabel heap, bit6 vertex instances, bit7 curves
    LCNT_FALLBACK = 7,  // bit0: flatness near-tie (needs libm hypot); bit1: input the device path does not take
    LCNT_BAD = 8,       // label references an entity / style / icon that does not exist
    LCNT_RING_PTS = 9,  // polylabel ring points handed out
    LCNT_ACTIVE = 10,   // statistics: active labels
    LCNT_POLY = 11,     // labels that ran polylabel (= heaps handed out)
    LCNT_COVER = 12,    // labels whose text coverage label_cover_kernel computes
    LCNT_VERTS = 13,    // outline vertex instances handed out (one per vertex of every placed glyph)
    LCNT_CURVES = 14,   // of those, curves (the expensive ones: they get a compact work list of their own)
    LCNT_COVER_ERR = 3,  // label_cover_kernel: a pair fell outside its proven window (a logic error, never silent)
    LCNT_CURVE_CURSOR_C = 16,  // work distribution of the curve kernels (counting / writing sweep)
    LCNT_CURVE_CURSOR_W = 17,
    LCNT_SCAN_OVF = 15,  // (overflow word of the block-sum scan: segment counts beyond 2^32 also trip the capacity check)
    LCNT_CULLED = 18,    // statistics: active labels label_cull_kernel took out
    LCNT_PRECULLED = 19, // of those, taken out before the layout (label_precull_kernel)
    LCNT_COUNT = 20
};

struct ActLabel {  // a label generation that can draw or collide
    unsigned entity, style;
    unsigned text;      // text run id, 0xffffffff: no text
    unsigned n_glyphs;
};
struct GlyphPlace {  // 48 bytes: where one glyph goes
    double a, b, c, d, e;  // text on a line: wx, wy, sin(-angle), cos(-angle), gcx; centred: x origin, baseline
    int slot;              // glyph outline, -1 none
    unsigned label;        // slot of the label in the per-tile regions
};
struct LabelPlace {  // layout result of an active label
    int icon, ix, iy;
    unsigned mode;       // 0: no text geometry, 1: along the way, 2: centred rows
    unsigned place_off;  // first GlyphPlace
    unsigned n_places;
    unsigned rgb;
    unsigned dead;  // label_cull_kernel: the label cannot reach the tile, directly or through a chain of collisions
    double scale;  // font units -> pixels
    double gcy;    // (descent + ascent) / 2 (text on a line)
    unsigned vinst_off;  // first outline vertex instance of the label (its glyphs' vertices, glyph after glyph)
    unsigned n_vinst;
};
// (GlyphOut -- a glyph's segment range and bounds -- is declared in osmr_kernels.cuh next to its reader, label_cover_kernel)

struct alignas(16) CurveRoot {  // draw_quad(to, control, from) of a placed glyph's curve vertex (text_placer.rs:211-231)
    double x0, y0, x1, y1, x2, y2;
};
struct LabelDev {
    // resident tables
    const DevLabelStyle* styles;
    unsigned n_styles;
    const unsigned* text_id;  // [key][ent_total]; entity slot = node | way_base + way | mp_base + mp
    unsigned n_keys, ent_total, way_base, mp_base;
    const unsigned* text_begin;
    unsigned n_texts;
    const DevGlyphRec* glyphs;
    const unsigned* glyph_vbegin;
    const DevVertex* verts;
    int ascent, descent, line_gap;
    const unsigned* way_angle_off[19];  // per zoom: first (sin, cos) pair of the way's oriented segments, 0xffffffff none
    const double2* sincos[19];
    const DevIcon* icons;
    unsigned n_icons;
    // batch
    const unsigned* label_begin;  // per tile (absolute)
    const osmr_label* labels;
    unsigned n_tiles;
    // scratch
    ActLabel* act;        // per-tile regions: act[label_begin[t] .. + act_cnt[t])
    unsigned* act_cnt;    // per tile
    LabelPlace* place;    // same indexing as act
    GlyphPlace* gplace;
    unsigned* place_vinst;  // per GlyphPlace: its first vertex instance
    unsigned gplace_cap;
    unsigned* vinst_place;  // per vertex instance: its GlyphPlace
    unsigned* vcnt;         // per vertex instance: segments it draws; after the scan: offset of its first segment ([n]: total)
    double4* vbox;          // per vertex instance: bounds of its segments (min x, max x, min y, max y)
    unsigned* curve_list;   // vertex instances that are curves
    CurveRoot* curve_root;  // per curve: its control points in pixel space (label_vfill_kernel)
    unsigned long long* curve_codes;  // per curve: kCurveLeafCap 16-bit leaf codes ((1 << depth) | path) of the segments it draws, in order
    unsigned char* curve_deep;        // per curve: 1 when the codes do not describe it (too many leaves, too deep): flattened again
    unsigned curves_cap;
    unsigned leaf_cap;                // <= kCurveLeafCap (smaller only in tests: forces the flatten-again path)
    unsigned verts_cap;
    unsigned* scan_blocks;  // block sums of the segment-offset scan
    unsigned n_scan_blocks;
    DevSeg* segs;
    unsigned segs_cap;
    DevLabel* out_labels;  // same indexing as act
    unsigned* cover_list;  // slots of the labels with text coverage (as long as the label list)
    unsigned rowrecs_cap;  // rows of kmin / kmax
    unsigned long long cells_cap;
    double2* ring_pts;  // polylabel scratch
    unsigned ring_cap;
    unsigned char* heap;  // polylabel heaps: kPolyHeapCap cells per label that needs one
    unsigned heap_slots;
    unsigned* counters;  // LCNT_*
    unsigned cull;       // 0: every active label gets its outlines and coverage (debug key "label_cull")
    unsigned char* predead;  // same indexing as act: label_precull_kernel's verdict (1: no layout either)
    int font_reach;          // font units: no outline point of a placed glyph lies farther from the glyph's anchor on the way /
                             // its pen position than font_reach * scale (outline extents, advances, ascent, descent)
};

// ------------------------------------------------------------------------------------------------------
// label_select_kernel: one CTA per tile, ordered compaction of the label generations that matter.
// A generation without an icon and without text geometry is an empty successful generation: no pixel, no collision
// (SURVEY.md A.6), so dropping it cannot change the image -- the host layout drops the same ones.
// ------------------------------------------------------------------------------------------------------
constexpr int kLabelSelThreads = 256;

__device__ __forceinline__ bool label_entity_slot(const Scene& s, const LabelDev& ld, unsigned entity, unsigned& slot, unsigned& kind) {
    const bool is_mp = (entity & OSMR_AREA_MULTIPOLYGON) != 0;
    const bool is_node = !is_mp && (entity & OSMR_LABEL_NODE) != 0;
    const unsigned idx = entity & ~(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE);
    kind = is_node ? 0u : (is_mp ? 2u : 1u);
    if (is_node) {
        if (idx >= s.n_nodes) return false;
        slot = idx;
    } else if (is_mp) {
        if (idx >= s.n_mps) return false;
        slot = ld.mp_base + idx;
    } else {
        if (idx >= s.n_ways) return false;
        slot = ld.way_base + idx;
    }
    return true;
}

__global__ void __launch_bounds__(kLabelSelThreads) label_select_kernel(Scene s, LabelDev ld) {
    __shared__ unsigned warp_cnt[kLabelSelThreads / 32];
    __shared__ unsigned running;
    const unsigned t = blockIdx.x;
    const unsigned first = ld.label_begin[t], last = ld.label_begin[t + 1];
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (unsigned start = first; start < last; start += kLabelSelThreads) {
        const unsigned li = start + threadIdx.x;
        bool active = false;
        ActLabel a;
        a.entity = a.style = 0;
        a.text = 0xffffffffu;
        a.n_glyphs = 0;
        if (li < last) {
            const osmr_label L = ld.labels[li];
            unsigned slot = 0, kind = 0;
            if (L.style >= ld.n_styles || !label_entity_slot(s, ld, L.entity, slot, kind)) {
                atomicOr(&ld.counters[LCNT_BAD], 1u);
            } else {
                const DevLabelStyle st = ld.styles[L.style];
                if (st.icon >= 0 && (unsigned)st.icon >= ld.n_icons) atomicOr(&ld.counters[LCNT_BAD], 1u);
                a.entity = L.entity;
                a.style = L.style;
                if ((st.flags & OSMR_LSTYLE_TEXT) && (st.flags & OSMR_LSTYLE_FONT_SIZE) && st.key >= 0) {
                    const unsigned tid = ld.text_id[(size_t)st.key * ld.ent_total + slot];
                    if (tid != 0xffffffffu) {
                        a.text = tid;
                        a.n_glyphs = ld.text_begin[tid + 1] - ld.text_begin[tid];
                    }
                }
                active = (st.icon >= 0 && (unsigned)st.icon < ld.n_icons) || a.text != 0xffffffffu;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, active);
        const unsigned w = threadIdx.x >> 5;
        if (lane_id() == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        unsigned pos = running;
        for (unsigned k = 0; k < w; ++k) pos += warp_cnt[k];
        pos += __popc(bal & ((1u << lane_id()) - 1u));
        if (active) ld.act[first + pos] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned add = 0;
            for (unsigned k = 0; k < kLabelSelThreads / 32; ++k) add += warp_cnt[k];
            running += add;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ld.act_cnt[t] = running;
        atomicAdd(&ld.counters[LCNT_ACTIVE], running);
    }
}

// ------------------------------------------------------------------------------------------------------
// polylabel (labelable.rs:125-232) on tile-relative f64 coordinates, one thread per label.
// std::collections::BinaryHeap's sift order is reproduced (push: sift up; pop: move the last element to the root, sift it to
// the bottom along the larger children, sift it back up), because equal max_fitness values are common on regular shapes.
// ------------------------------------------------------------------------------------------------------
constexpr unsigned kPolyHeapCap = 1024;  // cells per heap (40 KB); a larger heap sends the call to the host path
constexpr unsigned kMaxPolyRings = 64;

struct PolyCell {
    double cx, cy, half, fit, max_fit;
};

__device__ __forceinline__ double poly_seg_dist_sq(double px, double py, double2 a, double2 b) {  // labelable.rs:313-349
    double x = a.x, y = a.y, dx = b.x - x, dy = b.y - y;
    if (dx != 0.0 || dy != 0.0) {
        const double t = ((px - x) * dx + (py - y) * dy) / (dx * dx + dy * dy);
        if (t > 1.0) {
            x = b.x;
            y = b.y;
        } else if (t > 0.0) {
            x += dx * t;
            y += dy * t;
        }
    }
    dx = px - x;
    dy = py - y;
    return dx * dx + dy * dy;
}

struct PolyHeap {
    PolyCell* d;
    unsigned n;
    bool overflow;
    __device__ static bool le(const PolyCell& a, const PolyCell& b) { return !(a.max_fit > b.max_fit); }
    __device__ void up(unsigned pos) {
        const PolyCell e = d[pos];
        while (pos > 0) {
            const unsigned parent = (pos - 1) / 2;
            if (le(e, d[parent])) break;
            d[pos] = d[parent];
            pos = parent;
        }
        d[pos] = e;
    }
    __device__ void push(const PolyCell& v) {
        if (n >= kPolyHeapCap) {
            overflow = true;
            return;
        }
        d[n] = v;
        up(n);
        ++n;
    }
    __device__ bool pop(PolyCell& out) {
        if (n == 0) return false;
        const PolyCell last = d[n - 1];
        --n;
        if (n == 0) {
            out = last;
            return true;
        }
        out = d[0];
        const unsigned end = n;
        unsigned pos = 0, child = 1;
        while (end >= 2 && child <= end - 2) {
            if (le(d[child], d[child + 1])) ++child;
            d[pos] = d[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            d[pos] = d[child];
            pos = child;
        }
        d[pos] = last;
        up(pos);
        return true;
    }
};

// get_label_position (labelable.rs:191-204) for an area.  `pts` is this label's private copy of the rings (it is permuted).
// Returns false when the entity has no points.  *overflow: the heap did not fit.
__device__ bool poly_label_anchor(double2* pts, unsigned* off, unsigned n_rings, double scale, PolyCell* heap_mem, double& ax, double& ay,
                                  bool& overflow) {
    overflow = false;
    if (n_rings == 0 || off[1] - off[0] == 0) return false;
    // filter_polygons (labelable.rs:206-232): the ring of the largest area first, then the rings that lie inside it.  Rings are
    // permuted as whole (offset, length) records: ring order among the kept ones follows the reference's swaps.
    unsigned ro[kMaxPolyRings], rl[kMaxPolyRings];
    for (unsigned k = 0; k < n_rings; ++k) {
        ro[k] = off[k];
        rl[k] = off[k + 1] - off[k];
    }
    auto ring_area = [&](unsigned k) {
        double sacc = 0.0;
        for (unsigned i = 1; i < rl[k]; ++i) {
            const double2 a = pts[ro[k] + i], b = pts[ro[k] + i - 1];
            sacc += a.x * b.y - b.x * a.y;
        }
        return fabs(sacc);
    };
    unsigned big = 0;
    double big_area = ring_area(0);
    for (unsigned k = 1; k < n_rings; ++k) {
        const double a = ring_area(k);
        if (a > big_area) {
            big = k;
            big_area = a;
        }
    }
    {
        unsigned t0 = ro[0], t1 = rl[0];
        ro[0] = ro[big];
        rl[0] = rl[big];
        ro[big] = t0;
        rl[big] = t1;
    }
    // the distance functions below take PolyRings with prefix offsets: build it over an index table instead of moving points
    struct View {
        const double2* pts;
        const unsigned *ro, *rl;
    } v{pts, ro, rl};
    auto signed_dist = [&](double px, double py, unsigned n_use) {
        bool inside = false;
        double best = __longlong_as_double(0x7ff0000000000000LL);
        for (unsigned k = 0; k < n_use; ++k) {
            const unsigned o = v.ro[k], n = v.rl[k];
            for (unsigned i = 1; i < n; ++i) {
                const double2 a = v.pts[o + i], b = v.pts[o + i - 1];
                if ((a.y > py) != (b.y > py) && px < (b.x - a.x) * (py - a.y) / (b.y - a.y) + a.x) inside = !inside;
                best = fmin(best, poly_seg_dist_sq(px, py, a, b));
            }
        }
        return (inside ? 1.0 : -1.0) * sqrt(best);
    };
    unsigned keep = 1;
    for (unsigned k = 1; k < n_rings; ++k) {
        bool all_inside = true;
        for (unsigned i = 0; i < rl[k]; ++i) {
            const double2 p = pts[ro[k] + i];
            if (!(signed_dist(p.x, p.y, 1) >= 0.0)) {
                all_inside = false;
                break;
            }
        }
        if (all_inside) {
            unsigned t0 = ro[k], t1 = rl[k];
            ro[k] = ro[keep];
            rl[k] = rl[keep];
            ro[keep] = t0;
            rl[keep] = t1;
            ++keep;
        }
    }
    const unsigned n_use = keep;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double min_x = inf, max_x = -inf, min_y = inf, max_y = -inf;
    for (unsigned i = 0; i < rl[0]; ++i) {
        const double2 p = pts[ro[0] + i];
        min_x = fmin(min_x, p.x);
        max_x = fmax(max_x, p.x);
        min_y = fmin(min_y, p.y);
        max_y = fmax(max_y, p.y);
    }
    const double w = max_x - min_x, h = max_y - min_y;
    const double precision = fmax(w, h) / 100.0 * scale;
    const double cell = fmin(w, h), max_size = fmax(w, h);
    if (cell == 0.0) {
        ax = min_x;
        ay = min_y;
        return true;
    }
    double cen_x, cen_y;
    {
        double area = 0.0, cx = 0.0, cy = 0.0;
        for (unsigned i = 1; i < rl[0]; ++i) {
            const double2 a = pts[ro[0] + i], b = pts[ro[0] + i - 1];
            const double c = a.x * b.y - b.x * a.y;
            cx += (a.x + b.x) * c;
            cy += (a.y + b.y) * c;
            area += c * 3.0;
        }
        if (area == 0.0) {
            cen_x = pts[ro[0]].x;
            cen_y = pts[ro[0]].y;
        } else {
            cen_x = cx / area;
            cen_y = cy / area;
        }
    }
    auto fitness = [&](double cx, double cy, double d) {
        if (d <= 0.0) return d;
        const double dx = cx - cen_x, dy = cy - cen_y;
        return d * (1.0 - sqrt(dx * dx + dy * dy) / max_size);
    };
    auto cell_at = [&](double cx, double cy, double half) {
        const double d = signed_dist(cx, cy, n_use);
        PolyCell c;
        c.cx = cx;
        c.cy = cy;
        c.half = half;
        c.fit = fitness(cx, cy, d);
        c.max_fit = fitness(cx, cy, d + half * 1.41421356237309504880168872420969808);
        return c;
    };
    PolyHeap heap{heap_mem, 0u, false};
    double half = cell / 2.0;
    for (double x = min_x; x < max_x; x += cell)
        for (double y = min_y; y < max_y; y += cell) {
            heap.push(cell_at(x + half, y + half, half));
            if (heap.overflow) {
                overflow = true;
                return true;
            }
        }
    PolyCell best = cell_at(cen_x, cen_y, 0.0), cur;
    while (heap.pop(cur)) {
        if (cur.fit > best.fit) best = cur;
        if (cur.max_fit - best.fit <= precision) continue;
        half = cur.half / 2.0;
        for (int ix = 0; ix < 2; ++ix)
            for (int iy = 0; iy < 2; ++iy) {
                const double dx = ix ? 1.0 : -1.0, dy = iy ? 1.0 : -1.0;
                heap.push(cell_at(cur.cx + dx * half, cur.cy + dy * half, half));
            }
        if (heap.overflow) {
            overflow = true;
            return true;
        }
    }
    ax = best.cx;
    ay = best.cy;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// label_layout_kernel: one thread per active label (persistent grid over (tile, label) pairs)
// ------------------------------------------------------------------------------------------------------
constexpr int kLayoutThreads = 64;

__device__ __forceinline__ double2 tile_rel(const double2 m, const TileXform& t) {  // coords_to_xy_tile_relative * scale
    double2 r;
    r.x = (m.x * t.dim - t.tx256) * t.scale;
    r.y = (m.y * t.dim - t.ty256) * t.scale;
    return r;
}

__global__ void __launch_bounds__(kLayoutThreads) label_layout_kernel(Scene s, LabelDev ld) {
    const unsigned t = blockIdx.x;
    const unsigned first = ld.label_begin[t];
    const unsigned n_act = ld.act_cnt[t];
    const osmr_tile tile = s.tiles[t];
    const TileXform xf = make_xform(tile);
    const double gscale = (double)tile.scale;
    for (unsigned ai = threadIdx.x; ai < n_act; ai += kLayoutThreads) {
        const ActLabel a = ld.act[first + ai];
        const DevLabelStyle st = ld.styles[a.style];
        const bool is_mp = (a.entity & OSMR_AREA_MULTIPOLYGON) != 0;
        const bool is_node = !is_mp && (a.entity & OSMR_LABEL_NODE) != 0;
        const unsigned idx = a.entity & ~(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE);
        LabelPlace lp;
        lp.icon = -1;
        lp.ix = lp.iy = 0;
        lp.mode = 0;
        lp.place_off = 0;
        lp.n_places = 0;
        lp.rgb = st.rgb;
        lp.dead = 0;
        lp.scale = 0.0;
        lp.gcy = 0.0;
        lp.vinst_off = 0;
        lp.n_vinst = 0;
        if (ld.cull && ld.predead[first + ai]) {  // label_precull_kernel: cannot matter whatever its exact layout
            lp.dead = 1u;
            ld.place[first + ai] = lp;
            continue;
        }
        // label anchor (labelable.rs), lazily: polylabel is expensive
        bool anchor_done = false, anchor_ok = false;
        double anx = 0.0, any = 0.0;
        auto get_anchor = [&]() {
            if (anchor_done) return anchor_ok;
            anchor_done = true;
            if (is_node) {
                const int2 p = project_point(s.merc[idx], xf);
                anx = (double)p.x;
                any = (double)p.y;
                anchor_ok = true;
                return true;
            }
            // rings of the area as tile-relative f64 pixels, in this label's private scratch
            RingIter it(s, is_mp ? (idx | OSMR_AREA_MULTIPOLYGON) : idx);
            if (it.n_rings > kMaxPolyRings) {
                atomicOr(&ld.counters[LCNT_FALLBACK], 2u);
                return false;
            }
            unsigned total = 0;
            for (unsigned k = 0; k < it.n_rings; ++k) total += it.ring(k).y;
            const unsigned base = atomicAdd(&ld.counters[LCNT_RING_PTS], total);
            if (base + total > ld.ring_cap || base + total < base) {
                atomicOr(&ld.counters[LCNT_OVERFLOW], 16u);
                return false;
            }
            const unsigned hslot = atomicAdd(&ld.counters[LCNT_POLY], 1u);  // a heap of its own for this label
            if (hslot >= ld.heap_slots) {
                atomicOr(&ld.counters[LCNT_OVERFLOW], 32u);
                return false;
            }
            double2* pts = ld.ring_pts + base;
            unsigned off[kMaxPolyRings + 1];
            unsigned w = 0;
            for (unsigned k = 0; k < it.n_rings; ++k) {
                const uint2 r = it.ring(k);
                off[k] = w;
                for (unsigned q = 0; q < r.y; ++q) pts[w++] = tile_rel(s.merc[s.ints[r.x + q]], xf);
            }
            off[it.n_rings] = w;
            bool ovf = false;
            PolyCell* heap = reinterpret_cast<PolyCell*>(ld.heap) + (size_t)hslot * kPolyHeapCap;
            anchor_ok = poly_label_anchor(pts, off, it.n_rings, gscale, heap, anx, any, ovf);
            if (ovf) {
                atomicOr(&ld.counters[LCNT_FALLBACK], 2u);  // an unusually large heap: let the host lay this call out
                anchor_ok = false;
            }
            return anchor_ok;
        };
        unsigned y_offset = 0;
        // label_with_icon (labeler.rs:39-68)
        if (st.icon >= 0 && (unsigned)st.icon < ld.n_icons) {
            if (get_anchor()) {
                const DevIcon ic = ld.icons[st.icon];
                lp.icon = st.icon;
                lp.ix = f64_as_i32(anx - ((double)ic.w / 2.0));
                lp.iy = f64_as_i32(any - ((double)ic.h / 2.0));
                y_offset = ic.h / 2u;
            }
        }
        // label_with_text -> TextPlacer::place (text_placer.rs:24-168)
        if (a.text != 0xffffffffu) {
            const double font_size = st.font_size * gscale;
            const float fs = (float)font_size;
            const float fscale = fs / (float)(ld.ascent - ld.descent);  // scale_for_pixel_height: an f32 division
            const double scale = (double)fscale;
            const DevGlyphRec* gl = ld.glyphs + ld.text_begin[a.text];
            const unsigned ng = a.n_glyphs;
            auto width_of = [&](unsigned k) {
                double w = (double)gl[k].advance * scale;
                if (k > 0) w += (double)gl[k].kern * scale;
                return w;
            };
            double total_width = 0.0;
            for (unsigned k = 0; k < ng; ++k) total_width += width_of(k);
            const double ascent = (double)ld.ascent * scale, descent = (double)ld.descent * scale, line_gap = (double)ld.line_gap * scale;
            lp.scale = scale;
            const unsigned pos = st.text_position ? st.text_position : ((is_node || is_mp) ? (unsigned)OSMR_TEXT_POS_CENTER : (unsigned)OSMR_TEXT_POS_LINE);
            // the label's GlyphPlace block (one entry per character)
            unsigned poff = 0;
            bool have_block = false;
            auto take_block = [&]() {
                poff = atomicAdd(&ld.counters[LCNT_PLACES], ng);
                if (poff + ng > ld.gplace_cap || poff + ng < poff) {
                    atomicOr(&ld.counters[LCNT_OVERFLOW], 1u);
                    return false;
                }
                have_block = true;
                return true;
            };
            if (pos == OSMR_TEXT_POS_LINE) {
                if (!is_node && !is_mp) {  // only ways have waypoints (labelable.rs:33-39)
                    const uint2 wr = s.ways[idx];
                    const unsigned len = wr.y;
                    const unsigned aoff = (tile.zoom <= 18u && ld.way_angle_off[tile.zoom]) ? ld.way_angle_off[tile.zoom][idx] : 0xffffffffu;
                    if (len >= 2) {
                        if (aoff == 0xffffffffu) {
                            atomicOr(&ld.counters[LCNT_FALLBACK], 2u);  // no direction table for this way: host path
                        } else {
                            const double2* sc = ld.sincos[tile.zoom] + aoff;
                            const int2 pf = project_point(s.merc[s.ints[wr.x]], xf), pb = project_point(s.merc[s.ints[wr.x + len - 1]], xf);
                            const bool rev = pf.x > pb.x;
                            auto pt = [&](unsigned i) { return project_point(s.merc[s.ints[wr.x + (rev ? len - 1 - i : i)]], xf); };
                            double way_len = 0.0;
                            {
                                int2 prev = pt(0);
                                for (unsigned i = 1; i < len; ++i) {
                                    const int2 cur = pt(i);
                                    way_len += point_dist(prev.x, prev.y, cur.x, cur.y);
                                    prev = cur;
                                }
                            }
                            // The direction table holds sin / cos of the INTEGER pixel differences as seen from tile (0, 0).  They
                            // are the same in every tile except when a coordinate sits on an exact half pixel left of / above the
                            // tile origin (round-half-away-from-zero mirrors there): such a way is laid out by the host.
                            for (unsigned i = 0; i < len; ++i) {
                                const double2 r = tile_rel(s.merc[s.ints[wr.x + i]], xf);
                                if ((r.x < 0.0 && r.x - floor(r.x) == 0.5) || (r.y < 0.0 && r.y - floor(r.y) == 0.5))
                                    atomicOr(&ld.counters[LCNT_FALLBACK], 2u);
                            }
                            if (!(total_width > way_len) && take_block()) {
                                double cur = (way_len - total_width) / 2.0;
                                lp.gcy = (descent + ascent) / 2.0;
                                lp.mode = 1;
                                lp.place_off = poff;
                                lp.n_places = ng;
                                for (unsigned k = 0; k < ng; ++k) {
                                    const double wk = width_of(k);
                                    const double gcx = wk / 2.0;
                                    double wx = 0.0, wy = 0.0;
                                    unsigned seg = 0;
                                    // compute_way_position (text_placer.rs:265-296)
                                    {
                                        unsigned i = 0;
                                        double left = cur + gcx;
                                        bool found = false;
                                        int2 p0 = pt(0);
                                        while (left > 0.0 && i + 1 < len) {
                                            const int2 p1 = pt(i + 1);
                                            const double sd = point_dist(p0.x, p0.y, p1.x, p1.y);
                                            if (sd >= left) {
                                                const double ratio = left / sd;
                                                wx = (double)p0.x + (double)wsub(p1.x, p0.x) * ratio;
                                                wy = (double)p0.y + (double)wsub(p1.y, p0.y) * ratio;
                                                seg = i;
                                                found = true;
                                                break;
                                            }
                                            left -= sd;
                                            ++i;
                                            p0 = p1;
                                        }
                                        if (!found) {
                                            const int2 pl = pt(len - 1);
                                            wx = (double)pl.x;
                                            wy = (double)pl.y;
                                            seg = len - 2;
                                        }
                                    }
                                    GlyphPlace gp;
                                    gp.a = wx;
                                    gp.b = wy;
                                    gp.c = sc[seg].x;
                                    gp.d = sc[seg].y;
                                    gp.e = gcx;
                                    gp.slot = gl[k].slot;
                                    gp.label = first + ai;
                                    ld.gplace[poff + k] = gp;
                                    cur += wk;
                                }
                            }
                        }
                    }
                }
            } else if (get_anchor() && take_block()) {
                // centred rows with wrapping (text_placer.rs:101-168)
                const double max_text_width = 256.0 / 8.0;  // TILE_SIZE / 8, not scaled (text_placer.rs:298)
                unsigned n_rows = 0;
                {
                    double rw = 0.0;
                    for (unsigned k = 0; k < ng; ++k) {
                        const double wk = width_of(k);
                        rw += wk;
                        const bool last = k + 1 == ng;
                        const bool brk = gl[k].ws && (rw + wk > max_text_width);
                        if (brk || last) {
                            ++n_rows;
                            rw = 0.0;
                        }
                    }
                }
                const double row_h = ascent - descent + line_gap;
                double cur_y = any;
                if (y_offset > 0)
                    cur_y += (double)y_offset;
                else
                    cur_y -= (row_h * (double)n_rows) / 2.0;
                lp.mode = 2;
                lp.place_off = poff;
                lp.n_places = ng;
                unsigned row_first = 0;
                double rw = 0.0;
                for (unsigned k = 0; k < ng; ++k) {
                    const double wk = width_of(k);
                    rw += wk;
                    const bool last = k + 1 == ng;
                    const bool brk = gl[k].ws && (rw + wk > max_text_width);
                    if (brk || last) {
                        double cur_x = anx - rw / 2.0;
                        const double baseline = cur_y + ascent;
                        for (unsigned q = row_first; q <= k; ++q) {
                            GlyphPlace gp;
                            gp.a = cur_x;
                            gp.b = baseline;
                            gp.c = gp.d = gp.e = 0.0;
                            gp.slot = gl[q].slot;
                            gp.label = first + ai;
                            ld.gplace[poff + q] = gp;
                            cur_x += width_of(q);
                        }
                        cur_y += row_h;
                        row_first = k + 1;
                        rw = 0.0;
                    }
                }
            }
            (void)have_block;
            if (lp.mode != 0 && lp.n_places) {  // the label's outline vertex instances, glyph after glyph
                unsigned V = 0;
                for (unsigned k = 0; k < ng; ++k)
                    if (gl[k].slot >= 0) V += ld.glyph_vbegin[gl[k].slot + 1] - ld.glyph_vbegin[gl[k].slot];
                const unsigned vb = atomicAdd(&ld.counters[LCNT_VERTS], V);
                if (vb + V > ld.verts_cap || vb + V < vb) {
                    atomicOr(&ld.counters[LCNT_OVERFLOW], 64u);
                    lp.mode = 0;
                    lp.n_places = 0;
                } else {
                    unsigned o = vb;
                    for (unsigned k = 0; k < ng; ++k) {
                        ld.place_vinst[lp.place_off + k] = o;
                        if (gl[k].slot >= 0) o += ld.glyph_vbegin[gl[k].slot + 1] - ld.glyph_vbegin[gl[k].slot];
                    }
                    lp.vinst_off = vb;
                    lp.n_vinst = V;
                }
            }
        }
        ld.place[first + ai] = lp;
    }
}

// ------------------------------------------------------------------------------------------------------
// label_emit_kernel: one thread per GlyphPlace.  Glyph::rasterize (text_placer.rs:211-231): every outline vertex is scaled,
// transformed and handed to Rasterizer::draw_line / draw_quad (rasterizer.rs:27-107).  WRITE = false counts the segments and
// their bounds, WRITE = true stores them at the offsets label_finish_kernel assigned.
// ------------------------------------------------------------------------------------------------------
struct EmitSink {
    DevSeg* out;  // nullptr: count only
    unsigned n;
    double min_x, max_x, min_y, max_y;
    bool near_tie;
    __device__ void line(double x0, double y0, double x1, double y1) {
        if (y1 - y0 == 0.0) return;  // draw_line returns before touching anything (rasterizer.rs:30-32)
        if (out) {
            DevSeg sg;
            sg.x0 = x0;
            sg.y0 = y0;
            sg.x1 = x1;
            sg.y1 = y1;
            out[n] = sg;
        }
        ++n;
    }
    // Bounds of everything a vertex draws, without looking at the segments: a subdivided curve stays inside the convex hull of
    // its control points (every new point is a midpoint of two old ones, and a rounded midpoint of two doubles lies between
    // them), so the control points bound it.  The bounds only size the coverage window of the label; one pixel of slack is
    // added where they are turned into a pixel box.
    __device__ void bound(double x, double y) {
        min_x = fmin(min_x, x);  // NaN-ignoring min / max (f64::min / max)
        max_x = fmax(max_x, x);
        min_y = fmin(min_y, y);
        max_y = fmax(max_y, y);
    }
    // draw_quad's flatness test (rasterizer.rs:86-107): hypot(p0-p1) + hypot(p1-p2) <= 1.0001 * hypot(p0-p2)
    __device__ bool flat_enough(double x0, double y0, double x1, double y1, double x2, double y2) {
        const double ax = x0 - x1, ay = y0 - y1, bx = x1 - x2, by = y1 - y2;  // (only their squares are used)
        const double cx = x0 - x2, cy = y0 - y2;
        {
            // |a| + |b| <= 1.0001 |c|  <=>  2 sqrt(A B) <= T with T = 1.0001^2 C - A - B  <=>  T >= 0 and 4 A B <= T^2 (A, B, C the
            // squared lengths): no square root at all.  Both sides carry a few ulp of error; the decision is taken only with a
            // 1e-10 relative margin -- far outside anything the rounding of the reference's own hypot sum could flip -- and
            // everything closer goes to the three-root form below.
            const double A = ax * ax + ay * ay, B = bx * bx + by * by, C = cx * cx + cy * cy;
            const double Y = 1.00020001 * C, T = Y - A - B, P = 4.0 * A * B, Q = T * T;
            if (Y < 1e140 && Y > 1e-140 && A > 1e-140 && B > 1e-140) {
                if (T < -1e-10 * Y) return false;
                if (T > 1e-10 * Y) {
                    if (P < Q * (1.0 - 1e-10)) return true;
                    if (P > Q * (1.0 + 1e-10)) return false;
                }
            }
        }
        const double lhs = sqrt(ax * ax + ay * ay) + sqrt(bx * bx + by * by);
        const double rhs = 1.0001 * sqrt(cx * cx + cy * cy);
        const double big = 1e150, tiny = 1e-150;
        const bool safe = lhs < big && rhs < big && lhs > tiny && rhs > tiny;  // also false for NaN
        if (safe && lhs < rhs * (1.0 - 1e-12)) return true;
        if (safe && lhs > rhs * (1.0 + 1e-12)) return false;
        near_tie = true;  // only libm's hypot can decide: the call goes to the host path
        return true;
    }
    // The reference recurses (first half, then second half).  Here the subdivision tree is walked in the same order WITHOUT a
    // stack: the current node is (depth, path bits); descending to the first half is one midpoint step from the current points,
    // and after a leaf the next node -- the second half of the nearest ancestor whose first half was just finished -- gets its
    // control points by repeating the midpoint steps from the root along its path.  Same operations on the same operands as the
    // recursion, so the points are bit-identical; the explicit stack this replaces lived in local memory and its loads were
    // half of the kernel's stall samples.
    // shape of the subdivision tree, one bit per visited node in visiting order (1: subdivided).  The counting sweep records it
    // (shape_out), the writing sweep replays it (shape_in) and never evaluates a flatness test again.
    unsigned long long* shape_out;
    const unsigned long long* shape_in;
    __device__ void quad(double x0, double y0, double x1, double y1, double x2, double y2) {
        constexpr int kMaxDepth = 30;
        constexpr unsigned kShapeBits = 256;
        unsigned path = 0u;
        int depth = 0;
        double a0 = x0, b0 = y0, a1 = x1, b1 = y1, a2 = x2, b2 = y2;
        // (p1, p2) of the current node's parent, valid while the current node is a first half reached by descending: its
        // second half then follows without walking down from the root again
        double pa1 = 0.0, pb1 = 0.0, pa2 = 0.0, pb2 = 0.0;
        unsigned node = 0;  // visited nodes so far
        unsigned long long word = 0ull;
        const bool replay = shape_in != nullptr && (shape_in[3] >> 63) == 0ull;  // bit 255: the tree has more than 255 nodes
        if (shape_in) word = shape_in[0];
        for (;;) {
            bool flat;
            if (replay) {
                flat = ((word >> (node & 63u)) & 1ull) == 0ull;
            } else {
                flat = flat_enough(a0, b0, a1, b1, a2, b2);
                if (!flat && depth >= kMaxDepth) {  // absurd depth: not something the device decides
                    near_tie = true;
                    flat = true;
                }
                if (shape_out && node < kShapeBits - 1u) {
                    if (!flat) word |= 1ull << (node & 63u);
                    if ((node & 63u) == 63u) {
                        shape_out[node >> 6] = word;
                        word = 0ull;
                    }
                }
            }
            ++node;
            if (replay && (node & 63u) == 0u && node < kShapeBits) word = shape_in[node >> 6];
            if (!flat) {  // first half: (p0, (p0 + p1) / 2, m)
                const double ax = (a0 + a1) / 2.0, ay = (b0 + b1) / 2.0, bx = (a1 + a2) / 2.0, by = (b1 + b2) / 2.0;
                pa1 = a1;
                pb1 = b1;
                pa2 = a2;
                pb2 = b2;
                a2 = (ax + bx) / 2.0;
                b2 = (ay + by) / 2.0;
                a1 = ax;
                b1 = ay;
                path <<= 1;
                ++depth;
                continue;
            }
            line(a0, b0, a2, b2);
            if (depth > 0 && !(path & 1u)) {  // a first half (always reached by descending): its second half is (m, (p1 + p2) / 2, p2)
                a0 = a2;
                b0 = b2;
                a1 = (pa1 + pa2) / 2.0;
                b1 = (pb1 + pb2) / 2.0;
                a2 = pa2;
                b2 = pb2;
                path |= 1u;
                continue;
            }
            while (depth > 0 && (path & 1u)) {  // both halves of this ancestor are done
                path >>= 1;
                --depth;
            }
            if (depth == 0) break;
            path |= 1u;  // the second half of that ancestor: walk down from the root along its path
            a0 = x0;
            b0 = y0;
            a1 = x1;
            b1 = y1;
            a2 = x2;
            b2 = y2;
            for (int l = depth - 1; l >= 0; --l) {
                const double ax = (a0 + a1) / 2.0, ay = (b0 + b1) / 2.0, bx = (a1 + a2) / 2.0, by = (b1 + b2) / 2.0;
                const double mx = (ax + bx) / 2.0, my = (ay + by) / 2.0;
                if ((path >> l) & 1u) {
                    a0 = mx;
                    b0 = my;
                    a1 = bx;
                    b1 = by;
                } else {
                    a2 = mx;
                    b2 = my;
                    a1 = ax;
                    b1 = ay;
                }
            }
        }
        if (shape_out) {
            if (node >= kShapeBits - 1u) {
                shape_out[3] = 1ull << 63;  // too many nodes for the record: the writing sweep decides again
            } else {
                shape_out[node >> 6] = word;                               // the partial word (bits above `node` are zero)
                if ((node >> 6) < 3u) shape_out[3] = 0ull;                 // (word 3 carries the overflow mark: always written)
            }
        }
    }
};

// One outline vertex -> its draw_line calls (count, bounds, optionally the segments themselves at out[0..)).
__device__ __forceinline__ unsigned emit_vertex(const LabelDev& ld, const GlyphPlace& gp, const LabelPlace& lp, unsigned vi, unsigned v0, DevSeg* out,
                                                double& min_x, double& max_x, double& min_y, double& max_y, bool& near_tie,
                                                unsigned long long* shape_out = nullptr, const unsigned long long* shape_in = nullptr) {
    const DevVertex v = ld.verts[vi];
    if (v.type != 2 && v.type != 3) return 0u;  // a move draws nothing
    const double scale = lp.scale;
    const bool on_line = lp.mode == 1;
    const double wx = gp.a, wy = gp.b, sn = gp.c, cs = gp.d, gcx = gp.e, gcy = lp.gcy;
    auto tr = [&](double px, double py, double& ox, double& oy) {
        if (on_line) {  // text_placer.rs:76-93
            const double tx = px - gcx, ty = py - gcy;
            ox = wx + (tx * cs - ty * sn);
            oy = wy - (ty * cs + tx * sn);
        } else {  // text_placer.rs:150-160
            ox = wx + px;
            oy = wy - py;
        }
    };
    // `from` is the previous vertex's point ((0, 0) before the first one, text_placer.rs:213)
    double fx = 0.0, fy = 0.0;
    if (vi > v0) {
        const DevVertex pv = ld.verts[vi - 1];
        fx = (double)pv.x * scale;
        fy = (double)pv.y * scale;
    }
    const double tx = (double)v.x * scale, ty = (double)v.y * scale;
    EmitSink sink;
    sink.out = out;
    sink.n = 0;
    sink.min_x = min_x;
    sink.max_x = max_x;
    sink.min_y = min_y;
    sink.max_y = max_y;
    sink.near_tie = false;
    sink.shape_out = shape_out;
    sink.shape_in = shape_in;
    if (v.type == 2) {
        double p1x, p1y, p0x, p0y;
        tr(fx, fy, p1x, p1y);
        tr(tx, ty, p0x, p0y);
        sink.line(p0x, p0y, p1x, p1y);
        if (sink.n) {
            sink.bound(p0x, p0y);
            sink.bound(p1x, p1y);
        }
    } else {
        double p2x, p2y, p1x, p1y, p0x, p0y;
        tr(fx, fy, p2x, p2y);
        tr((double)v.cx * scale, (double)v.cy * scale, p1x, p1y);
        tr(tx, ty, p0x, p0y);
        sink.quad(p0x, p0y, p1x, p1y, p2x, p2y);
        if (sink.n) {
            sink.bound(p0x, p0y);
            sink.bound(p1x, p1y);
            sink.bound(p2x, p2y);
        }
    }
    min_x = sink.min_x;
    max_x = sink.max_x;
    min_y = sink.min_y;
    max_y = sink.max_y;
    near_tie = near_tie || sink.near_tie;
    return sink.n;
}

// ------------------------------------------------------------------------------------------------------
// label_cull_kernel: one CTA per tile, between the layout and the outline pass.
// The reference lays out and rasterises every label of the 3x3 neighbourhood, but a tile shows only its own 256 x 256 pixels: a
// label matters if its pixels can reach the tile, or if it can block -- directly or through a chain -- a label that does.  A
// label g fails exactly when one of its pixels is held by an EARLIER successful label (tile_pixels.rs:131-148), so with
// conservative pixel boxes B:   needed(g) = B(g) meets the tile  or  B(g) meets B(g') of a needed LATER label g'.
// One backward sweep decides it (needed(g) depends on later labels only).  Everything else is marked dead: no outlines, no
// coverage, no pixels claimed -- no needed label can tell the difference.  B = the icon rectangle + the hull of every outline
// point of every glyph (the bounds label_finish_kernel derives its proven coverage window from, two pixels wider).
// The neighbourhood is nine times the tile, so most labels of a tile's list are dead.
// ------------------------------------------------------------------------------------------------------
constexpr unsigned kCullCap = 1024;  // labels of a tile decided in shared memory; a longer list is not culled
constexpr int kCullThreads = 128;

// the backward sweep of the cull kernels: one warp, a label at a time, the lanes over the later labels
__device__ __forceinline__ void cull_sweep(const int4* box, unsigned char* needed, unsigned n_act, int D) {
    const unsigned lane = threadIdx.x & 31u;
    for (int g = (int)n_act - 1; g >= 0; --g) {
        const int4 b = box[g];
        bool need = false;
        if (b.x <= b.z && b.y <= b.w) {
            need = b.x <= D - 1 && b.z >= 0 && b.y <= D - 1 && b.w >= 0;
            for (unsigned h = (unsigned)g + 1u + lane; !need && h - lane < n_act; h += 32) {
                bool hit = false;
                if (h < n_act && needed[h]) {
                    const int4 o = box[h];
                    hit = o.x <= o.z && b.x <= o.z && b.z >= o.x && b.y <= o.w && b.w >= o.y;
                }
                need = __any_sync(0xffffffffu, hit);
            }
        }
        if (lane == 0) needed[g] = need ? 1 : 0;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------
// label_precull_kernel: the same decision BEFORE the layout, from boxes that need no layout: where the label's anchor can be
// (the node's pixel; for text along a way the pixel box of the way -- the glyphs sit on it) widened by what the text can add (along a way: font_reach * scale around a point of the way;
// centred rows: the summed |advance| to either side, one row height per possible row, font_reach * scale around every pen
// position) and by the icon.  These boxes contain the exact ones, so every label the exact sweep would keep is kept here;
// the rest skips label_layout_kernel (glyph placement along the way, polylabel) as well.  label_cull_kernel then decides
// among the survivors with the exact boxes.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCullThreads) label_precull_kernel(Scene s, LabelDev ld) {
    __shared__ int4 box[kCullCap];
    __shared__ unsigned char needed[kCullCap];
    const unsigned t = blockIdx.x;
    const unsigned first = ld.label_begin[t];
    const unsigned n_act = ld.act_cnt[t];
    const int D = s.D;
    if (!ld.cull) return;
    if (n_act > kCullCap) {
        for (unsigned ai = threadIdx.x; ai < n_act; ai += blockDim.x) ld.predead[first + ai] = 0;
        return;
    }
    if (n_act == 0) return;
    const osmr_tile tile = s.tiles[t];
    const TileXform xf = make_xform(tile);
    const double gscale = (double)tile.scale;
    for (unsigned ai = threadIdx.x; ai < n_act; ai += blockDim.x) {
        const ActLabel a = ld.act[first + ai];
        const DevLabelStyle st = ld.styles[a.style];
        const bool is_mp = (a.entity & OSMR_AREA_MULTIPOLYGON) != 0;
        const bool is_node = !is_mp && (a.entity & OSMR_LABEL_NODE) != 0;
        const unsigned idx = a.entity & ~(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE);
        // where the anchor / the way's points can be
        long long ax0, ay0, ax1, ay1;
        bool everywhere = false, has_anchor = true;
        if (is_node) {
            const int2 p = project_point(s.merc[idx], xf);
            ax0 = ax1 = p.x;
            ay0 = ay1 = p.y;
        } else {
            int x0, y0, x1, y1;
            entity_pixel_bbox((is_mp ? s.mp_box : s.way_box)[idx], xf, x0, y0, x1, y1);
            has_anchor = x0 <= x1 && y0 <= y1;
            ax0 = (long long)x0 - 1;  // (polylabel works on the unrounded coordinates: half a pixel off the integer box)
            ay0 = (long long)y0 - 1;
            ax1 = (long long)x1 + 1;
            ay1 = (long long)y1 + 1;
        }
        long long bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
        auto widen = [&](double l, double r, double u, double d) {  // anchor box widened by l, r, u, d pixels
            if (!(l >= 0.0 && l < 1.0e8 && r >= 0.0 && r < 1.0e8 && u >= 0.0 && u < 1.0e8 && d >= 0.0 && d < 1.0e8)) {
                everywhere = true;
                return;
            }
            bx0 = min(bx0, ax0 - (long long)ceil(l) - 4);
            bx1 = max(bx1, ax1 + (long long)ceil(r) + 4);
            by0 = min(by0, ay0 - (long long)ceil(u) - 4);
            by1 = max(by1, ay1 + (long long)ceil(d) + 4);
        };
        {
            // An anchor that comes from polylabel cannot be bounded without running it: for a degenerate ring (no interior, or a
            // signed area that cancels) the answer is the ring's centroid (labelable.rs:163), which need not lie in any box of the
            // entity -- or is not a number at all.  Such a label is decided by label_cull_kernel, from its real position.
            const bool has_icon = st.icon >= 0 && (unsigned)st.icon < ld.n_icons;
            const unsigned tpos = st.text_position ? st.text_position : ((is_node || is_mp) ? (unsigned)OSMR_TEXT_POS_CENTER : (unsigned)OSMR_TEXT_POS_LINE);
            if (!is_node && (has_icon || (a.text != 0xffffffffu && tpos != OSMR_TEXT_POS_LINE))) everywhere = true;
        }
        unsigned icon_h = 0;
        if (has_anchor && st.icon >= 0 && (unsigned)st.icon < ld.n_icons) {
            const DevIcon ic = ld.icons[st.icon];
            icon_h = ic.h;
            widen((double)ic.w, (double)ic.w, (double)ic.h, (double)ic.h);
        }
        if (has_anchor && a.text != 0xffffffffu) {
            const double font_size = st.font_size * gscale;
            const float fs = (float)font_size;
            const float fscale = fs / (float)(ld.ascent - ld.descent);
            const double scale = fabs((double)fscale);
            const double reach = (double)ld.font_reach * scale;
            const unsigned pos = st.text_position ? st.text_position : ((is_node || is_mp) ? (unsigned)OSMR_TEXT_POS_CENTER : (unsigned)OSMR_TEXT_POS_LINE);
            if (pos == OSMR_TEXT_POS_LINE) {
                if (!is_node && !is_mp) widen(reach, reach, reach, reach);
            } else {
                const DevGlyphRec* gl = ld.glyphs + ld.text_begin[a.text];
                double tw = 0.0;
                unsigned rows = 1;
                for (unsigned k = 0; k < a.n_glyphs; ++k) {
                    tw += (fabs((double)gl[k].advance) + fabs((double)gl[k].kern)) * scale;
                    rows += gl[k].ws ? 1u : 0u;
                }
                const double row_h = (fabs((double)ld.ascent) + fabs((double)ld.descent) + fabs((double)ld.line_gap)) * scale;
                const double v = row_h * (double)(rows + 1u) + (double)icon_h + reach;
                widen(tw + reach, tw + reach, v, v);
            }
        }
        if (everywhere || !(ax0 > -1000000000ll && ax1 < 1000000000ll && ay0 > -1000000000ll && ay1 < 1000000000ll)) {
            if (bx0 <= bx1 || everywhere) {
                bx0 = by0 = -0x7fffffff;
                bx1 = by1 = 0x7fffffff;
            }
        }
        box[ai] = make_int4((int)max(bx0, -0x7fffffffll), (int)max(by0, -0x7fffffffll), (int)min(bx1, 0x7fffffffll), (int)min(by1, 0x7fffffffll));
    }
    __syncthreads();
    if (threadIdx.x < 32) cull_sweep(box, needed, n_act, D);
    __syncthreads();
    unsigned pre = 0;
    for (unsigned ai = threadIdx.x; ai < n_act; ai += blockDim.x) {
        ld.predead[first + ai] = needed[ai] ? 0 : 1;
        pre += needed[ai] ? 0u : 1u;
    }
    if (pre) atomicAdd(&ld.counters[LCNT_PRECULLED], pre);
}

__global__ void __launch_bounds__(kCullThreads) label_cull_kernel(Scene s, LabelDev ld) {
    __shared__ int4 box[kCullCap];  // x0, y0, x1, y1 (x0 > x1: no pixels)
    __shared__ unsigned char needed[kCullCap];
    const unsigned t = blockIdx.x;
    const unsigned first = ld.label_begin[t];
    const unsigned n_act = ld.act_cnt[t];
    const int D = s.D;
    if (!ld.cull || n_act > kCullCap || n_act == 0) return;
    if (ld.counters[LCNT_OVERFLOW] & 65u) return;  // (places or vertex instances overflowed: the attempt is redone)
    const double lim = 1.0e9;
    for (unsigned ai = threadIdx.x; ai < n_act; ai += blockDim.x) {
        const LabelPlace lp = ld.place[first + ai];
        long long x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
        bool everywhere = false;  // coordinates beyond any canvas or not finite: never culled, meets everything
        if (lp.dead) {  // label_precull_kernel took it out: no box, it stays dead
            box[ai] = make_int4(1, 1, 0, 0);
            continue;
        }
        if (lp.icon >= 0 && (unsigned)lp.icon < ld.n_icons) {
            const DevIcon ic = ld.icons[lp.icon];
            x0 = lp.ix;
            y0 = lp.iy;
            x1 = (long long)lp.ix + (long long)ic.w - 1;
            y1 = (long long)lp.iy + (long long)ic.h - 1;
            if (lp.ix > 1000000000 || lp.ix < -1000000000 || lp.iy > 1000000000 || lp.iy < -1000000000) everywhere = true;
        }
        if (lp.mode != 0) {
            const double inf = __longlong_as_double(0x7ff0000000000000LL);
            double mnx = inf, mxx = -inf, mny = inf, mxy = -inf;
            const bool on_line = lp.mode == 1;
            const double scale = lp.scale, gcy = lp.gcy;
            for (unsigned gi = lp.place_off; gi < lp.place_off + lp.n_places; ++gi) {
                const GlyphPlace gp = ld.gplace[gi];
                if (gp.slot < 0) continue;
                const double wx = gp.a, wy = gp.b, sn = gp.c, cs = gp.d, gcx = gp.e;
                auto bound = [&](double px, double py) {  // emit_vertex's transform
                    double ox, oy;
                    if (on_line) {
                        const double tx = px - gcx, ty = py - gcy;
                        ox = wx + (tx * cs - ty * sn);
                        oy = wy - (ty * cs + tx * sn);
                    } else {
                        ox = wx + px;
                        oy = wy - py;
                    }
                    if (!(ox > -lim && ox < lim && oy > -lim && oy < lim)) everywhere = true;  // (NaN included)
                    mnx = fmin(mnx, ox);
                    mxx = fmax(mxx, ox);
                    mny = fmin(mny, oy);
                    mxy = fmax(mxy, oy);
                };
                const unsigned v0 = ld.glyph_vbegin[gp.slot], v1 = ld.glyph_vbegin[gp.slot + 1];
                if (v1 > v0) bound(0.0, 0.0);  // `from` of the first vertex
                for (unsigned vi = v0; vi < v1; ++vi) {
                    const DevVertex v = ld.verts[vi];
                    bound((double)v.x * scale, (double)v.y * scale);
                    if (v.type == 3) bound((double)v.cx * scale, (double)v.cy * scale);
                }
            }
            if (mnx <= mxx && !everywhere) {  // label_finish_kernel's window (floor - 1 .. floor + 2), two pixels wider
                x0 = min(x0, (long long)floor(mnx) - 3);
                y0 = min(y0, (long long)floor(mny) - 3);
                x1 = max(x1, (long long)floor(mxx) + 4);
                y1 = max(y1, (long long)floor(mxy) + 4);
            }
        }
        if (everywhere) {
            x0 = y0 = -0x7fffffff;
            x1 = y1 = 0x7fffffff;
        }
        box[ai] = make_int4((int)x0, (int)y0, (int)x1, (int)y1);
    }
    __syncthreads();
    if (threadIdx.x < 32) cull_sweep(box, needed, n_act, D);
    __syncthreads();
    unsigned culled = 0;
    for (unsigned ai = threadIdx.x; ai < n_act; ai += blockDim.x) {
        if (!needed[ai]) {
            ld.place[first + ai].dead = 1u;
            ++culled;
        }
    }
    if (culled) atomicAdd(&ld.counters[LCNT_CULLED], culled);
}

// The outline pass.  A glyph is ~30 vertices: straight lines (one draw_line each) and quadratic curves (~64 draw_line calls each
// under the 1.0001 flatness rule, ~20k instructions).  Lanes that hold lines would idle next to lanes that hold curves, so the
// curves get a compact work list of their own and every lane of the curve kernels flattens a curve:
//   label_vfill_kernel    thread per GlyphPlace: back pointers of its vertex instances, curves appended to the curve list
//   label_vline_kernel    thread per vertex instance that is not a curve  } COUNT: segments + bounds per instance,
//   label_curve_kernel    thread per curve                                } WRITE: the segments at their scanned offsets
//   label_scan_*          exclusive scan of the per-instance counts -> segment offsets: a label's segments are contiguous and in
//                         the reference's order (glyph, vertex, subdivision order)
__global__ void __launch_bounds__(128) label_vfill_kernel(LabelDev ld) {
    const unsigned n_places = min(ld.counters[LCNT_PLACES], ld.gplace_cap);
    if (ld.counters[LCNT_OVERFLOW] & 65u) return;
    for (unsigned gi = blockIdx.x * blockDim.x + threadIdx.x; gi < n_places; gi += gridDim.x * blockDim.x) {
        const GlyphPlace gp = ld.gplace[gi];
        if (gp.slot < 0) continue;
        const LabelPlace lp = ld.place[gp.label];
        if (lp.mode == 0) continue;  // the label lost its vertex block to an overflow
        const unsigned v0 = ld.glyph_vbegin[gp.slot], v1 = ld.glyph_vbegin[gp.slot + 1];
        const unsigned vo = ld.place_vinst[gi];
        if (lp.dead) {  // label_cull_kernel: its vertex instances draw nothing and belong to no glyph
            for (unsigned vi = v0; vi < v1; ++vi) {
                ld.vinst_place[vo + (vi - v0)] = 0xffffffffu;
                ld.vcnt[vo + (vi - v0)] = 0u;
            }
            continue;
        }
        const double scale = lp.scale;
        const bool on_line = lp.mode == 1;
        const double wx = gp.a, wy = gp.b, sn = gp.c, cs = gp.d, gcx = gp.e, gcy = lp.gcy;
        auto tr = [&](double px, double py, double& ox, double& oy) {
            if (on_line) {  // text_placer.rs:76-93
                const double tx = px - gcx, ty = py - gcy;
                ox = wx + (tx * cs - ty * sn);
                oy = wy - (ty * cs + tx * sn);
            } else {  // text_placer.rs:150-160
                ox = wx + px;
                oy = wy - py;
            }
        };
        double fx = 0.0, fy = 0.0;  // `from`: the previous vertex's point ((0, 0) before the first one)
        for (unsigned vi = v0; vi < v1; ++vi) {
            const unsigned inst = vo + (vi - v0);
            ld.vinst_place[inst] = gi;
            const DevVertex v = ld.verts[vi];
            const double tx = (double)v.x * scale, ty = (double)v.y * scale;
            if (v.type == 3) {
                // the curve's work item: Glyph::rasterize (text_placer.rs:211-231) calls draw_quad(to, control, from); the
                // curve kernels start from these points without walking the placement tables again
                const unsigned k = atomicAdd(&ld.counters[LCNT_CURVES], 1u);
                if (k >= ld.curves_cap) {  // (the counter keeps counting: the redo knows what to allocate)
                    atomicOr(&ld.counters[LCNT_OVERFLOW], 128u);
                    fx = tx;
                    fy = ty;
                    continue;
                }
                CurveRoot r;
                tr(tx, ty, r.x0, r.y0);
                tr((double)v.cx * scale, (double)v.cy * scale, r.x1, r.y1);
                tr(fx, fy, r.x2, r.y2);
                ld.curve_list[k] = inst;
                ld.curve_root[k] = r;
                // bounds from the control points (a subdivided curve stays inside their hull; EmitSink::bound).  The counting
                // sweep empties them again for a curve that draws nothing.
                ld.vbox[inst] = make_double4(fmin(fmin(r.x0, r.x1), r.x2), fmax(fmax(r.x0, r.x1), r.x2), fmin(fmin(r.y0, r.y1), r.y2),
                                             fmax(fmax(r.y0, r.y1), r.y2));
            }
            fx = tx;
            fy = ty;
        }
    }
}

template <bool WRITE>
__device__ __forceinline__ void label_vertex_instance(const LabelDev& ld, unsigned inst) {
    const unsigned gi = ld.vinst_place[inst];
    const GlyphPlace gp = ld.gplace[gi];
    const LabelPlace lp = ld.place[gp.label];
    const unsigned v0 = ld.glyph_vbegin[gp.slot];
    const unsigned vi = v0 + (inst - ld.place_vinst[gi]);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double mnx = inf, mxx = -inf, mny = inf, mxy = -inf;
    bool tie = false;
    if (WRITE) {
        const unsigned off = ld.vcnt[inst], n = ld.vcnt[inst + 1] - off;
        if (n) emit_vertex(ld, gp, lp, vi, v0, ld.segs + off, mnx, mxx, mny, mxy, tie);
    } else {
        const unsigned n = emit_vertex(ld, gp, lp, vi, v0, nullptr, mnx, mxx, mny, mxy, tie);
        ld.vcnt[inst] = n;
        ld.vbox[inst] = make_double4(mnx, mxx, mny, mxy);
        if (tie) atomicOr(&ld.counters[LCNT_FALLBACK], 1u);
    }
}

template <bool WRITE>
__device__ __forceinline__ void label_vline_body(const LabelDev& ld) {
    const unsigned n_verts = min(ld.counters[LCNT_VERTS], ld.verts_cap);
    if (ld.counters[LCNT_OVERFLOW] & 193u) return;
    if (WRITE && (ld.counters[LCNT_OVERFLOW] || ld.counters[LCNT_FALLBACK])) return;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_verts; i += gridDim.x * blockDim.x) {
        const unsigned gi = ld.vinst_place[i];
        if (gi == 0xffffffffu) continue;  // a culled label's instance
        const unsigned vi = ld.glyph_vbegin[ld.gplace[gi].slot] + (i - ld.place_vinst[gi]);
        if (ld.verts[vi].type == 3) continue;  // curves: label_curve_kernel
        label_vertex_instance<WRITE>(ld, i);
    }
}
// The curve kernels.  draw_quad (rasterizer.rs:86-107) is a recursion whose shape differs from curve to curve (8 .. 128
// draw_line calls), decided by a flatness test per node.
//
// label_curve_count_kernel runs the recursion as a state machine, the warp in lock step: every iteration every lane visits one
// node of ITS curve's subdivision tree -- the flatness decision, then either the midpoint step into the first half or the leaf
// and the move to the next second half -- and a lane whose curve is finished takes the next one.  The recursion's stack is
// (p1, p2) per level in shared memory, so the second half of ANY ancestor follows its first half in one step: it starts where
// the last leaf ended (the midpoint is handed down unchanged) and its control point is (p1 + p2) / 2.  The warp takes 32 curves
// of the list with ONE atomic and keeps them staged, one per lane (control points precomputed by label_vfill_kernel); a lane
// gets its next curve by shuffle, and the loads of a freshly taken chunk are in flight while the warp works on.  Besides the
// count, every leaf that draws leaves its CODE behind: (1 << depth) | path, 16 bits.
//
// label_curve_expand_kernel then writes the segments without any tree walk: a lane per leaf goes from the curve's control points
// down the leaf's path (depth midpoint steps in registers) and stores its segment at the scanned offset -- no stack, no
// divergence, coalesced stores.  Every node's points come from its parent's by the same midpoint operations whatever the
// traversal order, so the leaves are bit-identical to the recursion's.  A curve with more leaves than the code record holds (or
// deeper than 15 levels) is flattened again by one lane, flatness tests included.
constexpr int kCurveStackLevels = 7;  // levels of the subdivision stack kept in shared memory (a 90 degree curve needs 6; 7 CTAs per SM fit)
constexpr int kCurveMaxDepth = 30;
constexpr unsigned kCurveLeafCap = 128;  // leaf codes per curve
struct CurveState {
    double a0, b0, a1, b1, a2, b2;  // control points of the current node
    unsigned path, n_segs, inst, curve;
    int depth;
    unsigned long long pack;  // leaf codes not yet stored (four per word)
    bool deep, tie;
};

// levels beyond the shared-memory stack (out of line: real glyph curves never get here, and the common path pays nothing for them)
__device__ __noinline__ void curve_deep_push(double* q, double a1, double b1, double a2, double b2) {
    q[0] = a1;
    q[1] = b1;
    q[2] = a2;
    q[3] = b2;
}
__device__ __noinline__ void curve_deep_pop(const double* q, double& p1x, double& p1y, double& p2x, double& p2y) {
    p1x = q[0];
    p1y = q[1];
    p2x = q[2];
    p2y = q[3];
}

__global__ void __launch_bounds__(128) label_curve_count_kernel(LabelDev ld) {
    __shared__ double2 stk[2 * kCurveStackLevels * 128];  // [2][level][thread]: (p1, p2) of the ancestors of the current node
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned n_curves = min(ld.counters[LCNT_CURVES], ld.curves_cap);
    if (ld.counters[LCNT_OVERFLOW] & 193u) return;
    unsigned* cursor = &ld.counters[LCNT_CURVE_CURSOR_C];
    const unsigned lane = lane_id();
    double2* my_stk = stk + threadIdx.x;
    double deep[(kCurveMaxDepth + 1 - kCurveStackLevels) * 4];  // absurdly deep trees only (local memory, never touched otherwise)
    CurveState c;
    bool active = false;
    const EmitSink tester{};  // (only its flatness test is used)
    CurveRoot st_r = {};
    unsigned st_inst = 0;
    bool st_valid = false;
    unsigned pool_pos = 32;  // lanes pool_pos .. 31 hold staged curves that nobody has taken yet
    unsigned pool_base = 0;  // list index of lane 0's staged curve
    bool list_done = false;
    for (;;) {
        const unsigned need = __ballot_sync(kFull, !active);
        if (need) {
            if (pool_pos < 32u) {
                const unsigned rank = (unsigned)__popc(need & ((1u << lane) - 1u));
                const unsigned take = min((unsigned)__popc(need), 32u - pool_pos);
                const bool taker = !active && rank < take;
                const int src = taker ? (int)(pool_pos + rank) : (int)lane;
                const double x0 = __shfl_sync(kFull, st_r.x0, src), y0 = __shfl_sync(kFull, st_r.y0, src);
                const double x1 = __shfl_sync(kFull, st_r.x1, src), y1 = __shfl_sync(kFull, st_r.y1, src);
                const double x2 = __shfl_sync(kFull, st_r.x2, src), y2 = __shfl_sync(kFull, st_r.y2, src);
                const unsigned inst = __shfl_sync(kFull, st_inst, src);
                const bool valid = __shfl_sync(kFull, (int)st_valid, src) != 0;
                if (taker && valid) {
                    c.curve = pool_base + (unsigned)src;
                    c.inst = inst;
                    c.a0 = x0; c.b0 = y0; c.a1 = x1; c.b1 = y1; c.a2 = x2; c.b2 = y2;
                    c.path = 0u;
                    c.n_segs = 0u;
                    c.depth = 0;
                    c.pack = 0ull;
                    c.deep = false;
                    c.tie = false;
                    active = true;
                }
                pool_pos += take;
            }
            if (pool_pos == 32u && !list_done) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(cursor, 32u);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n_curves) {
                    list_done = true;
                } else {
                    pool_base = base;
                    pool_pos = 0u;
                    const unsigned k = base + lane;
                    st_valid = k < n_curves;
                    if (st_valid) {
                        st_r = ld.curve_root[k];
                        st_inst = ld.curve_list[k];
                    }
                }
            }
        }
        if (pool_pos == 32u && list_done && __all_sync(kFull, !active)) break;
        if (active) {
            // ---- the current node: flat? ----
            EmitSink t2 = tester;
            t2.near_tie = false;
            bool flat = t2.flat_enough(c.a0, c.b0, c.a1, c.b1, c.a2, c.b2);
            if (t2.near_tie) c.tie = true;
            if (!flat && c.depth >= kCurveMaxDepth) {  // absurd depth: not something the device decides
                c.tie = true;
                flat = true;
            }
            if (!flat) {
                // descend into the first half (p0, (p0 + p1) / 2, m); (p1, p2) stay behind for the second half
                if (c.depth < kCurveStackLevels) {
                    double2* q = my_stk + c.depth * 128;
                    q[0] = make_double2(c.a1, c.b1);
                    q[kCurveStackLevels * 128] = make_double2(c.a2, c.b2);
                } else {
                    curve_deep_push(deep + (c.depth - kCurveStackLevels) * 4, c.a1, c.b1, c.a2, c.b2);
                }
                const double ax = (c.a0 + c.a1) / 2.0, ay = (c.b0 + c.b1) / 2.0, bx = (c.a1 + c.a2) / 2.0, by = (c.b1 + c.b2) / 2.0;
                c.a2 = (ax + bx) / 2.0;
                c.b2 = (ay + by) / 2.0;
                c.a1 = ax;
                c.b1 = ay;
                c.path <<= 1;
                ++c.depth;
            } else {
                // draw_line(p0, p2) of this leaf (rasterizer.rs:30-32: nothing happens when y does not change)
                if (c.b2 - c.b0 != 0.0) {
                    if (c.n_segs < ld.leaf_cap && c.depth <= 15) {
                        c.pack |= (unsigned long long)((1u << c.depth) | c.path) << (16u * (c.n_segs & 3u));
                        if ((c.n_segs & 3u) == 3u) {
                            ld.curve_codes[(size_t)c.curve * (kCurveLeafCap / 4u) + (c.n_segs >> 2)] = c.pack;
                            c.pack = 0ull;
                        }
                    } else {
                        c.deep = true;
                    }
                    ++c.n_segs;
                }
                // second halves that are finished take their ancestors with them
                const int up = min(__ffs((int)~c.path) - 1, c.depth);
                c.path >>= up;
                c.depth -= up;
                if (c.depth == 0) {
                    // ---- the curve is finished ----
                    ld.vcnt[c.inst] = c.n_segs;
                    if (!c.n_segs) {  // (label_vfill_kernel wrote the hull of the control points)
                        const double inf = __longlong_as_double(0x7ff0000000000000LL);
                        ld.vbox[c.inst] = make_double4(inf, -inf, inf, -inf);
                    }
                    if ((c.n_segs & 3u) && c.n_segs < kCurveLeafCap) ld.curve_codes[(size_t)c.curve * (kCurveLeafCap / 4u) + (c.n_segs >> 2)] = c.pack;
                    ld.curve_deep[c.curve] = c.deep ? 1 : 0;
                    if (c.tie) atomicOr(&ld.counters[LCNT_FALLBACK], 1u);
                    active = false;
                } else {
                    // The node is a first half whose subtree is complete: the leaf just drawn ends in the parent's midpoint m
                    // (second halves hand p2 down unchanged), and the parent's second half is (m, (p1 + p2) / 2, p2).
                    double p1x, p1y, p2x, p2y;
                    const int lv = c.depth - 1;
                    if (lv < kCurveStackLevels) {
                        const double2* q = my_stk + lv * 128;
                        const double2 p1 = q[0], p2 = q[kCurveStackLevels * 128];
                        p1x = p1.x;
                        p1y = p1.y;
                        p2x = p2.x;
                        p2y = p2.y;
                    } else {
                        curve_deep_pop(deep + (lv - kCurveStackLevels) * 4, p1x, p1y, p2x, p2y);
                    }
                    c.a0 = c.a2;
                    c.b0 = c.b2;
                    c.a1 = (p1x + p2x) / 2.0;
                    c.b1 = (p1y + p2y) / 2.0;
                    c.a2 = p2x;
                    c.b2 = p2y;
                    c.path |= 1u;
                }
            }
        }
    }
}

// a curve the leaf codes do not describe: the recursion itself, flatness tests included (EmitSink::quad)
__device__ __noinline__ void curve_flatten_again(const CurveRoot& r, DevSeg* out) {
    EmitSink sink;
    sink.out = out;
    sink.n = 0;
    sink.near_tie = false;
    sink.shape_out = nullptr;
    sink.shape_in = nullptr;
    sink.quad(r.x0, r.y0, r.x1, r.y1, r.x2, r.y2);
}

// a warp per group of 8 curves of the list, a lane per leaf of the group (a curve alone would leave a third of the lanes idle)
constexpr unsigned kExpandGroup = 8;
__global__ void __launch_bounds__(256, 4) label_curve_expand_kernel(LabelDev ld) {
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned n_curves = min(ld.counters[LCNT_CURVES], ld.curves_cap);
    if (ld.counters[LCNT_OVERFLOW] || ld.counters[LCNT_FALLBACK]) return;
    const unsigned lane = lane_id();
    const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
    const unsigned n_groups = (n_curves + kExpandGroup - 1) / kExpandGroup;
    for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += n_warps) {
        const unsigned k0 = g * kExpandGroup;
        // lanes 0 .. 7: the curves of the group
        unsigned off = 0, n = 0;
        bool deep = false;
        if (lane < kExpandGroup && k0 + lane < n_curves) {
            const unsigned inst = ld.curve_list[k0 + lane];
            off = ld.vcnt[inst];
            n = ld.vcnt[inst + 1] - off;
            deep = n != 0u && ld.curve_deep[k0 + lane] != 0;
            if (deep) {  // not described by its codes: flattened again, by this lane alone
                curve_flatten_again(ld.curve_root[k0 + lane], ld.segs + off);
                n = 0;
            }
        }
        unsigned incl = n;
        for (int o = 1; o < (int)kExpandGroup; o <<= 1) {
            const unsigned y = __shfl_up_sync(kFull, incl, o);
            if ((int)lane >= o) incl += y;
        }
        const unsigned pre = incl - n;  // leaves of the group before this curve's
        const unsigned total = __shfl_sync(kFull, incl, kExpandGroup - 1);
        // leaf t of the group -> (curve j, leaf i of that curve) and its code; the NEXT round's code is loaded before this round's
        // midpoint steps (the 16-bit load and the count-leading-zeros behind it were 43 % of the kernel's stall samples)
        auto locate = [&](unsigned t, unsigned& k, unsigned& slot, unsigned& code) {
            unsigned j = 0;  // the last curve whose first leaf is not behind t
            for (unsigned step = kExpandGroup / 2; step; step >>= 1) {
                const unsigned pj = __shfl_sync(kFull, pre, j + step);
                if (pj <= t) j += step;
            }
            const unsigned pj = __shfl_sync(kFull, pre, j), oj = __shfl_sync(kFull, off, j);
            k = k0 + j;
            slot = oj + (t - pj);
            code = 0u;
            if (t < total) code = reinterpret_cast<const unsigned short*>(ld.curve_codes + (size_t)k * (kCurveLeafCap / 4u))[t - pj];
        };
        unsigned k_cur = 0, slot_cur = 0, code_cur = 0;
        if (total) locate(lane, k_cur, slot_cur, code_cur);
        for (unsigned tb = 0; tb < total; tb += 32) {
            unsigned k_nxt = 0, slot_nxt = 0, code_nxt = 0;
            if (tb + 32 < total) locate(tb + 32 + lane, k_nxt, slot_nxt, code_nxt);
            if (tb + lane < total) {
                const unsigned code = code_cur;
                const CurveRoot r = ld.curve_root[k_cur];
                double a0 = r.x0, b0 = r.y0, a1 = r.x1, b1 = r.y1, a2 = r.x2, b2 = r.y2;
                for (int lv = 30 - __clz(code); lv >= 0; --lv) {
                    const double ax = (a0 + a1) / 2.0, ay = (b0 + b1) / 2.0, bx = (a1 + a2) / 2.0, by = (b1 + b2) / 2.0;
                    const double mx = (ax + bx) / 2.0, my = (ay + by) / 2.0;
                    if ((code >> lv) & 1u) {  // the second half (m, b, p2)
                        a0 = mx;
                        b0 = my;
                        a1 = bx;
                        b1 = by;
                    } else {  // the first half (p0, a, m)
                        a2 = mx;
                        b2 = my;
                        a1 = ax;
                        b1 = ay;
                    }
                }
                DevSeg sg;
                sg.x0 = a0;
                sg.y0 = b0;
                sg.x1 = a2;
                sg.y1 = b2;
                ld.segs[slot_cur] = sg;
            }
            k_cur = k_nxt;
            slot_cur = slot_nxt;
            code_cur = code_nxt;
        }
    }
}
__global__ void __launch_bounds__(128) label_vline_count_kernel(LabelDev ld) { label_vline_body<false>(ld); }
__global__ void __launch_bounds__(128) label_vline_write_kernel(LabelDev ld) { label_vline_body<true>(ld); }

// exclusive scan of vcnt[0 .. n) in place (n = the vertex instance counter), vcnt[n] = total = the number of segments.
// Three launches: sums of blocks of 1024 elements, auto_scan_kernel over the block sums, local scans + block offsets.
constexpr unsigned kScanBlock = 1024;
__global__ void __launch_bounds__(256) label_scan_sums_kernel(LabelDev ld) {
    __shared__ unsigned wsum[8];
    const unsigned n = min(ld.counters[LCNT_VERTS], ld.verts_cap);
    const unsigned base = blockIdx.x * kScanBlock;
    unsigned v = 0;
    for (unsigned k = 0; k < kScanBlock / 256; ++k) {
        const unsigned i = base + k * 256 + threadIdx.x;
        if (i < n) v += ld.vcnt[i];
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int k = 0; k < 8; ++k) t += wsum[k];
        ld.scan_blocks[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) label_scan_apply_kernel(LabelDev ld) {
    __shared__ unsigned wsum[8];
    const unsigned n = min(ld.counters[LCNT_VERTS], ld.verts_cap);
    const unsigned base = blockIdx.x * kScanBlock;
    if (base > n) return;
    // thread t owns 4 consecutive elements
    const unsigned i0 = base + threadIdx.x * 4;
    unsigned x[4];
    unsigned mine = 0;
    for (int k = 0; k < 4; ++k) {
        x[k] = i0 + k < n ? ld.vcnt[i0 + k] : 0u;
        mine += x[k];
    }
    unsigned incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += y;
    }
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned before = ld.scan_blocks[blockIdx.x];  // already exclusive (auto_scan_kernel)
    for (unsigned k = 0; k < (threadIdx.x >> 5); ++k) before += wsum[k];
    unsigned run = before + incl - mine;
    for (int k = 0; k < 4; ++k) {
        if (i0 + k <= n) ld.vcnt[i0 + k] = run;  // (element n receives the total)
        run += x[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned total = ld.scan_blocks[ld.n_scan_blocks];
        ld.counters[LCNT_SEGS] = total;
        if (total > ld.segs_cap) atomicOr(&ld.counters[LCNT_OVERFLOW], 2u);
    }
}

// ------------------------------------------------------------------------------------------------------
// label_finish_kernel: one thread per active label -- the batch assembly the host did in osmr_draw_tiles_labeled: segment range,
// pixel bbox of the text, rows inside the label canvas, coverage cells, the work list of label_cover_kernel.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) label_finish_kernel(Scene s, LabelDev ld) {
    const unsigned t = blockIdx.x;
    const unsigned first = ld.label_begin[t];
    const unsigned n_act = ld.act_cnt[t];
    const int D = s.D;
    if (ld.counters[LCNT_OVERFLOW] & 195u) return;
    for (unsigned ai = threadIdx.x; ai < n_act; ai += blockDim.x) {
        const LabelPlace lp = ld.place[first + ai];
        DevLabel L;
        L.icon = lp.icon;
        L.ix = lp.ix;
        L.iy = lp.iy;
        L.seg_begin = 0;
        L.seg_count = 0;
        L.bx0 = 1;
        L.bx1 = 0;
        L.by0 = 1;
        L.by1 = 0;
        L.rgb = lp.rgb;
        L.ry0 = 0;
        L.rows = 0;
        L.width = 0;
        L.row_first = 0;
        L.n_ranges = 0;
        L.range_off = 0;
        L.cell_off = 0;
        if (lp.dead) L.icon = -2;  // (no icon, no segments: label_commit_kernel passes over it)
        if (lp.mode != 0 && lp.n_vinst && !lp.dead) {
            const unsigned sb = ld.vcnt[lp.vinst_off], n_segs = ld.vcnt[lp.vinst_off + lp.n_vinst] - sb;
            if (n_segs) {
                double min_x = __longlong_as_double(0x7ff0000000000000LL), min_y = min_x;
                double max_x = __longlong_as_double((long long)0xfff0000000000000ULL), max_y = max_x;
                for (unsigned k = 0; k < lp.n_vinst; ++k) {
                    const double4 bb = ld.vbox[lp.vinst_off + k];
                    min_x = fmin(min_x, bb.x);
                    max_x = fmax(max_x, bb.y);
                    min_y = fmin(min_y, bb.z);
                    max_y = fmax(max_y, bb.w);
                }
                L.seg_begin = sb;
                L.seg_count = n_segs;
                // (conservative bounds from the control points, one pixel of slack; the `s` column is one past the last `a` column)
                L.bx0 = f64_as_i32(floor(min_x)) - 1;
                L.bx1 = f64_as_i32(floor(max_x)) + 2;
                L.by0 = f64_as_i32(floor(min_y)) - 1;
                L.by1 = f64_as_i32(floor(max_y)) + 1;
                // rows outside the label canvas cannot collide or draw; columns stay complete (the sweep is a prefix sum)
                L.ry0 = max(L.by0, -D);
                const long long rows = (long long)min(L.by1, 2 * D - 1) - L.ry0 + 1;
                const long long cols = (long long)L.bx1 - L.bx0 + 1;
                const bool touches = rows > 0 && cols > 0 && L.bx1 >= -D && L.bx0 <= 2 * D - 1;
                if (touches) {
                    if (cols > (1 << 20)) {
                        atomicOr(&ld.counters[LCNT_FALLBACK], 2u);
                    } else {
                        const unsigned rb = atomicAdd(&ld.counters[LCNT_ROWRECS], (unsigned)rows);
                        const unsigned long long need = (unsigned long long)rows * (unsigned long long)cols;
                        const unsigned long long cb = atomicAdd(reinterpret_cast<unsigned long long*>(&ld.counters[LCNT_CELLS_LO]), need);
                        const unsigned ci = atomicAdd(&ld.counters[LCNT_COVER], 1u);
                        if (rb + (unsigned)rows > ld.rowrecs_cap || rb + (unsigned)rows < rb) {
                            atomicOr(&ld.counters[LCNT_OVERFLOW], 4u);
                        } else if (cb + need > ld.cells_cap) {
                            atomicOr(&ld.counters[LCNT_OVERFLOW], 8u);
                        } else {
                            L.rows = (int)rows;
                            L.width = (int)cols;
                            L.row_first = rb;
                            L.cell_off = cb;
                            ld.cover_list[ci] = first + ai;  // (at most one entry per label slot: the list is as long as the label list)
                        }
                    }
                }
            }
        }
        ld.out_labels[first + ai] = L;
    }
}

}  // namespace osmr

// osmr_kernels.cuh -- the CUDA kernels of the tile rasteriser (sm_100a).  See DESIGN.md for the data flow.
//
//   project_nodes_kernel   once per dataset   lat/lon -> Mercator factor pair (transcendental half of a1)
//   area_bbox_kernel       per batch          integer pixel bbox + point count of every (tile, styled area)
//   plan_ops_kernel        per batch          a8: (pass, area) generations -> ordered visible-op list per tile
//   build_geometry_kernel  per batch          a1 (exact half) + a2/a4 prologue: per visible op the edge / segment
//                                             records that can touch the tile, dash phase prefix sums
//   fill_rows_kernel       per batch          a2/a3: even-odd row masks (1 bit / pixel) of every visible fill op
//   raster_kernel          per batch          a3 blend, a4/a5 line coverage, a6/a7 compositor + RGB(A) export
#pragma once
#include "osmr_device.cuh"

namespace osmr {

// ------------------------------------------------------------------------------------------------------
// device-side data model
// ------------------------------------------------------------------------------------------------------
struct DevIcon {
    unsigned w, h, off, pad;  // off: first texel in icon_px
};

struct AreaInfo {  // a styled area in a tile
    int x0, y0, x1, y1;  // integer pixel bbox of all its points (x0 > x1: no points)
    unsigned npts;
};

// One visible generation.  g = pass * n_areas_of_tile + index (drawer.rs:94-100: Fill, Casing, Stroke).
struct VisOp {  // 48 bytes: what the geometry / fill / cover kernels need of an op (no look-up chain through the batch arrays)
    unsigned g;
    short x0, y0, x1, y1;  // reach bbox clamped to [-1, D]
    unsigned geom_off;     // first 16-byte unit of its records in the geometry scratch
    unsigned geom_cnt;     // records written by build_geometry_kernel
    unsigned mask_off;     // fills: first word of its row masks; lines: first cell of its (block -> bin entry) pair table
    unsigned kind;         // OP_*
    unsigned area;         // index of the styled area inside its tile (g = pass * n + area)
    unsigned pass;         // 0 Fill, 1 Casing, 2 Stroke
    unsigned tile;         // tile index inside the range being drawn
    unsigned entity;       // osmr_styled_area of the op
    unsigned style;
};
enum { OP_FILL_COLOR = 0, OP_FILL_IMAGE = 1, OP_LINE = 2 };

// The 32 bytes raster_kernel reads per op that hits a block (one aligned record instead of the chain
// vis -> areas -> styles and the style mapping of drawer.rs:156-219).
struct alignas(16) RasterOp {
    unsigned a;       // lines: geom_off; fills: mask_off
    unsigned b;       // lines: geom_cnt (written by build_geometry_kernel); image fills: icon index
    short y0, y1;     // VisOp.y0 / y1
    unsigned char kind;    // OP_*
    unsigned char rgb[3];  // colour of the pass
    unsigned short reach;  // lines: line_reach(half width), the culling margin around a segment
    unsigned short pad;
    double opacity;   // colour fills: fill-opacity (lines carry theirs in the fragment alphas)
};
static_assert(sizeof(RasterOp) == 32, "RasterOp layout");

struct SegRec {  // 64 bytes: one line segment (or outer cap line) that can touch the tile
    int x1, y1, x2, y2;
    double traveled;           // OpacityCalculator.traveled_distance when this segment is drawn (line.rs:31)
    double denom;              // center_dist_denom (line.rs:106)
    unsigned long long magic;  // floor(2^64 / (2*mx_d)) + 1: exact quotients for numerators < 2^32 (flags bit1)
    unsigned flags;            // bit0: outer cap calculator (line.rs:22,33-57); bit1: 32-bit fast path valid; bit2: coords < 2^24
    int k0;                    // first main step whose perpendiculars can reach the tile; line_cover_kernel walks k0 .. k0+n_k-1
    unsigned n_k;
    unsigned pad0;
    unsigned long long pad1;
};
static_assert(sizeof(SegRec) == 64, "SegRec layout");
// Opacity calculators depend only on (style, pass, scale, use_caps_for_dashes): style_calc_kernel builds them once per
// draw into a table; entry (style, pass) = one full OpacityCalc (dashes of the op) + header and first segment of the
// outer-cap calculator.
constexpr unsigned kCapCalcUnits = 8;                                      // header + one dash segment
constexpr unsigned kMainCalcUnits = (unsigned)(sizeof(OpacityCalc) / 16);  // 4 + 4 * kMaxDashSegs
constexpr unsigned kCalcEntryUnits = kMainCalcUnits + kCapCalcUnits;

enum {
    CNT_GEOM_USED = 0,   // 16-byte units
    CNT_MASK_USED = 1,   // words
    CNT_N_WORK = 2,      // visible ops (all kinds)
    CNT_N_FILL_WORK = 3,
    CNT_OVERFLOW = 4,    // bit0 geometry scratch, bit1 mask scratch, bit2 fragment storage, bit3 sort scratch (f3), bit4 fill / bit5 line
                         // work list, bit6 bin entries, bit7 pair tables, bit8 a (block, op) fragment list outgrew its proven capacity
    CNT_BAD_INPUT = 5,   // entity / style index out of range
    CNT_WORK_CURSOR = 6,
    CNT_FILL_CURSOR = 7,
    CNT_NODE_REFS_LO = 8,
    CNT_NODE_REFS_HI = 9,
    CNT_VISIBLE = 10,
    CNT_BIG_COORDS = 11,  // some visible segment has a coordinate >= 2^24 (raster uses exact i64 cross products)
    CNT_WALK_ALPHA = 12,  // 64-bit (12,13): fragment slots handed out by bin_ops_kernel (sum of the pair capacities)
    CNT_WALK_LEN = 14,    // 64-bit (14,15): (f3 reuses 12,13 for its sort scratch; unused by the draw path)
    CNT_N_LINE_WORK = 16,
    CNT_LINE_CURSOR = 17,
    CNT_WALK_STEPS = 20,  // 64-bit (20,21): fragments stored by line_cover_kernel (statistics)
    CNT_WALK_TRUNC = 18,  // bit0: a perpendicular walk outlived its proven bound; bit1: a fragment fell into a block the binning had
                          // ruled out (never observed; the draw fails loudly)
    CNT_BIN_ENTRIES = 19, // (block, op) entries handed out by bin_ops_kernel
    CNT_PAIR_USED = 22,   // pair table cells handed out by plan_ops_kernel
    CNT_COUNT = 24
};

// what the label pass leaves for a pixel of the centre tile: the pending pixel of the last successful label that
// touched it (tile_pixels.rs:131-148,205-223)
struct LabelPix {
    double alpha;   // text: coverage `total`; icon: unused (texel alpha is in the icon table)
    unsigned src;   // 0: none; 0x40000000 | 0xBBGGRR: text colour; 0x80000000 | texel index: icon texel
    unsigned pad;
};

// Bounding box of an entity (way: its nodes; multipolygon: the nodes of all its rings) in Mercator factor units, computed once
// per dataset.  project_point is monotone in each coordinate (multiplication by a positive power of two, subtraction of a
// constant, multiplication by the positive scale, round, saturating cast), so the pixel bbox of the entity in ANY tile is
// the projection of these two corners: no per-tile pass over the nodes.  NaN coordinates (which project to pixel 0 in every
// tile) are carried as flags.
struct alignas(16) EntBox {
    double x0, y0, x1, y1;
    unsigned npts;
    unsigned nan_flags;  // bit0: some x is NaN, bit1: some y is NaN
    unsigned pad[2];
};

struct Scene {
    // dataset
    const double2* merc;
    const EntBox* way_box;
    const EntBox* mp_box;
    const uint2* ways;
    const uint2* polys;
    const uint2* mps;
    const unsigned* ints;
    unsigned n_nodes, n_ways, n_polys, n_mps, n_ints;
    // styles
    const osmr_style* styles;
    const double* dashes;
    unsigned n_styles, n_dashes;
    const DevIcon* icons;
    const double4* icon_px;
    unsigned n_icons;
    // batch
    const osmr_tile* tiles;
    const unsigned* area_begin;
    const osmr_styled_area* areas;
    unsigned n_tiles, n_areas;  // of the tile range being drawn (tiles/area_begin point at its first tile)
    unsigned area_base;         // absolute index of the range's first styled area (area_begin values are absolute)
    int D;      // 256 * scale
    int scale;
    unsigned flags;
    unsigned char canvas[3];
    // scratch
    VisOp* vis;            // tile t owns vis[3*area_begin[t] ...)
    RasterOp* rop;         // same indexing as vis
    short4* vis_bbox;      // reach bbox of vis[i] (8 bytes; what the raster warps scan)
    unsigned* vis_count;   // per tile
    unsigned plan_slices;       // CTAs per (tile, pass) of plan_ops_kernel (1: the whole sub-list in one CTA)
    unsigned* plan_slice_base;  // [3 * tile + pass][slice]: visible ops of the slice (plan_count_kernel), then their exclusive scan
    unsigned* work;        // global indices into vis (all visible ops)
    uint2* fill_work;      // (global index into vis, chunk of 32 mask rows): work items of fill_rows_kernel
    uint2* line_work;      // (global index into vis, batch of 32 segment records): work items of line_cover_kernel
    unsigned fill_work_cap, line_work_cap;
    // Fragments (line_cover_kernel -> raster_kernel): every in-line, in-tile step of every perpendicular walk as (pixel inside its
    // 16x16 block, alpha), appended to the list of its (block, op) pair.  bin_ops_kernel sizes the lists and orders the pairs.
    double* frag_alpha;
    unsigned char* frag_pix;
    unsigned long long frag_cap;
    struct BinEntry* entries;  // per (block, op) pair, grouped by block in generation order
    unsigned* frag_cnt;        // fragments stored per entry
    unsigned entries_cap;
    uint2* blk_range;          // per (tile, block): first entry, number of entries
    unsigned* pair;            // per line op: entry index of every block of its bbox rectangle (0xffffffff: no segment reaches it)
    unsigned bin_slices;               // > 1: the ops of a block area are cut into slices with a CTA each (bin_count / bin_scan / bin_write)
    unsigned* bin_cnt_ent;             // [block area][slice][thread]: entries of the slice (bin_count_kernel), then its first entry
    unsigned long long* bin_cnt_cap;   // same indexing: fragment capacity, then the first fragment slot
    unsigned pair_cap;
    const uint4* calc_table;  // kCalcEntryUnits per (style, pass-1), built by style_calc_kernel
    uint4* geom;           // geometry scratch, 16-byte units
    unsigned geom_cap;
    unsigned* mask;        // fill row masks
    unsigned mask_cap;
    unsigned* counters;
    int fill_cap;          // <= kFillCap; lowered by tests to force the streaming path
    const struct LabelPix* label_plane;  // per tile D*D entries written by label_kernel, or nullptr (no label pass)
    const unsigned* label_mask;          // per tile D*D bits: the pixels of label_plane that hold a pending label pixel
    const double4* label_icon_px;        // premultiplied texels of the label icons
    unsigned char* out;
};

#ifndef OSMR_BW
#define OSMR_BW 16
#endif
#ifndef OSMR_BH
#define OSMR_BH 16
#endif
constexpr int kBW = OSMR_BW;  // block owned by one raster warp: kBW x kBH pixels
constexpr int kBH = OSMR_BH;
constexpr int kBP = kBW * kBH;
static_assert(kBW == 16 && kBH == 16, "block shape: a pixel index is one byte, a block row of a fill mask is 16 bits");

// blocks of the tile covered by an op's reach bbox (VisOp.x0..y1, clamped to [-1, D]): first block column / row, columns, rows
__device__ __forceinline__ void op_block_rect(int x0, int y0, int x1, int y1, int D, int& bx0, int& by0, int& nbx, int& nby) {
    bx0 = max(x0, 0) / kBW;
    by0 = max(y0, 0) / kBH;
    nbx = min(x1, D - 1) / kBW - bx0 + 1;
    nby = min(y1, D - 1) / kBH - by0 + 1;
}

// Fill masks are stored BLOCK-MAJOR: per fill op one 32-byte record for every 16x16 block of its bbox rectangle, 16 bits per
// block row (word j of a record = rows 2j and 2j+1).  raster_kernel fetches the whole record of its block with one sector read;
// round 1 stored full tile rows per op and every raster warp loaded 8 row words one after the other (the top stall after the
// fragment rewrite: 22 % of its samples).
__device__ __forceinline__ void store_mask_word(unsigned* mask_base, int rbx0, int rby0, int nbx, int y, int k, unsigned word) {
    // word k of tile row y covers block columns 2k and 2k+1
    unsigned short* m16 = reinterpret_cast<unsigned short*>(mask_base);
    const int cy = y / kBH - rby0, ry = y % kBH;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cx = 2 * k + h - rbx0;
        if (cx >= 0 && cx < nbx) m16[(size_t)(cy * nbx + cx) * kBH + ry] = (unsigned short)((word >> (16 * h)) & 0xffffu);
    }
}

struct alignas(16) BinEntry {
    unsigned op;        // global index into Scene.rop / vis
    unsigned frag_off;  // lines: first fragment slot
    unsigned frag_cap;  // lines: slots (a proven upper bound of what line_cover_kernel can store)
    unsigned pad;
};

constexpr int kFillCap = 128;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Bump allocation with ONE atomic per warp: every lane asks for n units (0: none) and gets its own offset.  Must be called by
// all 32 lanes.  (650 k single-address atomics per counter and batch were a third of plan_ops / build_geometry.)
__device__ __forceinline__ unsigned warp_alloc(unsigned* counter, unsigned n) {
    unsigned incl = n;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += v;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned base = 0;
    if (lane_id() == 0 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + incl - n;
}

// Dynamic work distribution of the persistent kernels: a warp takes kWorkBatch consecutive items per atomic.
constexpr unsigned kWorkBatch = 4;

// ------------------------------------------------------------------------------------------------------
// a1 (transcendental half): reference src/tile.rs:88-101 up to `factor`.
// ------------------------------------------------------------------------------------------------------
__global__ void project_nodes_kernel(const unsigned char* __restrict__ nodes, unsigned n, double2* __restrict__ merc) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* rec = reinterpret_cast<const double*>(nodes + (size_t)i * 32);
    double lat = rec[1], lon = rec[2];
    const double rads_per_deg = kPi / 180.0;  // f64::to_radians
    double lat_rad = lat * rads_per_deg;
    double lon_rad = lon * rads_per_deg;
    double x = lon_rad + kPi;
    double y = kPi - log(tan((kPi / 4.0) + (lat_rad / 2.0)));
    double2 m;
    m.x = x / (2.0 * kPi);
    m.y = y / (2.0 * kPi);
    merc[i] = m;
}

// Point::from_node for every node of the dataset (parity probe of a1).
__global__ void project_all_kernel(const double2* __restrict__ merc, unsigned n, osmr_tile tile, int2* __restrict__ out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TileXform xf = make_xform(tile);
    out[i] = project_point(merc[i], xf);
}

// ------------------------------------------------------------------------------------------------------
// helpers over the entity tables
// ------------------------------------------------------------------------------------------------------
struct RingIter {  // iterates the rings of a way (1 ring) or multipolygon (polygon_count rings)
    const Scene& s;
    bool is_mp;
    unsigned n_rings, poly_off;
    uint2 way;
    __device__ RingIter(const Scene& sc, unsigned entity) : s(sc) {
        is_mp = (entity & OSMR_AREA_MULTIPOLYGON) != 0;
        unsigned idx = entity & ~OSMR_AREA_MULTIPOLYGON;
        if (is_mp) {
            uint2 m = s.mps[idx];
            poly_off = m.x;
            n_rings = m.y;
        } else {
            way = s.ways[idx];
            n_rings = 1;
            poly_off = 0;
        }
    }
    __device__ uint2 ring(unsigned k) const { return is_mp ? s.polys[s.ints[poly_off + k]] : way; }
};

__device__ __forceinline__ bool entity_valid(const Scene& s, unsigned entity) {
    unsigned idx = entity & ~OSMR_AREA_MULTIPOLYGON;
    return (entity & OSMR_AREA_MULTIPOLYGON) ? idx < s.n_mps : idx < s.n_ways;
}

// ------------------------------------------------------------------------------------------------------
// entity_box_kernel: once per dataset, one thread per way / multipolygon
// ------------------------------------------------------------------------------------------------------
__global__ void entity_box_kernel(Scene s, EntBox* way_box, EntBox* mp_box) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n_ways + s.n_mps) return;
    const bool is_mp = i >= s.n_ways;
    const unsigned e = is_mp ? i - s.n_ways : i;
    EntBox b;
    b.x0 = b.y0 = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    b.x1 = b.y1 = __longlong_as_double((long long)0xfff0000000000000ULL);  // -inf
    b.npts = 0;
    b.nan_flags = 0;
    b.pad[0] = b.pad[1] = 0;
    RingIter it(s, is_mp ? (e | OSMR_AREA_MULTIPOLYGON) : e);
    for (unsigned k = 0; k < it.n_rings; ++k) {
        const uint2 r = it.ring(k);
        for (unsigned q = 0; q < r.y; ++q) {
            const double2 m = s.merc[s.ints[r.x + q]];
            if (m.x != m.x) b.nan_flags |= 1u;
            if (m.y != m.y) b.nan_flags |= 2u;
            b.x0 = fmin(b.x0, m.x);  // fmin / fmax ignore NaN
            b.y0 = fmin(b.y0, m.y);
            b.x1 = fmax(b.x1, m.x);
            b.y1 = fmax(b.y1, m.y);
        }
        b.npts += r.y;
    }
    (is_mp ? mp_box : way_box)[e] = b;
}

// pixel bbox of an entity in a tile from its EntBox (see there); x0 > x1 when it has no points
__device__ __forceinline__ void entity_pixel_bbox(const EntBox& b, const TileXform& xf, int& x0, int& y0, int& x1, int& y1) {
    x0 = y0 = 0x7fffffff;
    x1 = y1 = (int)0x80000000;
    if (b.x0 <= b.x1) {
        const int2 lo = project_point(make_double2(b.x0, b.y0), xf), hi = project_point(make_double2(b.x1, b.y1), xf);
        x0 = lo.x;
        x1 = hi.x;
    }
    if (b.y0 <= b.y1) {
        const int2 lo = project_point(make_double2(b.x0, b.y0), xf), hi = project_point(make_double2(b.x1, b.y1), xf);
        y0 = lo.y;
        y1 = hi.y;
    }
    if (b.nan_flags & 1u) {  // `as i32` of NaN is 0
        x0 = min(x0, 0);
        x1 = max(x1, 0);
    }
    if (b.nan_flags & 2u) {
        y0 = min(y0, 0);
        y1 = max(y1, 0);
    }
}

// ------------------------------------------------------------------------------------------------------
// per-op style mapping (drawer.rs:156-219)
// ------------------------------------------------------------------------------------------------------
struct LineParams {
    double width, opacity;
    const double* dashes;
    int n_dashes;
    bool has_dashes;
    unsigned cap;
    unsigned char rgb[3];
};

__device__ __forceinline__ bool line_params(const Scene& s, const osmr_style& st, int pass, LineParams& lp) {
    double scale = (double)s.scale;
    if (pass == 1) {  // Casing (drawer.rs:187-201)
        if (!(st.flags & OSMR_STYLE_CASING_COLOR) || !(st.flags & OSMR_STYLE_CASING_WIDTH)) return false;
        lp.width = st.casing_width * scale;
        lp.opacity = 1.0;
        lp.has_dashes = (st.flags & OSMR_STYLE_CASING_DASHES) != 0;
        lp.dashes = s.dashes + st.casing_dashes_off;
        lp.n_dashes = (int)st.casing_dashes_len;
        lp.cap = st.casing_line_cap;
        lp.rgb[0] = st.casing_color[0];
        lp.rgb[1] = st.casing_color[1];
        lp.rgb[2] = st.casing_color[2];
    } else {  // Stroke (drawer.rs:203-216)
        if (!(st.flags & OSMR_STYLE_COLOR)) return false;
        lp.width = scale * ((st.flags & OSMR_STYLE_WIDTH) ? st.width : 1.0);
        lp.opacity = (st.flags & OSMR_STYLE_OPACITY) ? st.opacity : 1.0;
        lp.has_dashes = (st.flags & OSMR_STYLE_DASHES) != 0;
        lp.dashes = s.dashes + st.dashes_off;
        lp.n_dashes = (int)st.dashes_len;
        lp.cap = st.line_cap;
        lp.rgb[0] = st.color[0];
        lp.rgb[1] = st.color[1];
        lp.rgb[2] = st.color[2];
    }
    return true;
}

// pixels a segment's perpendicular walks can reach beyond the segment's own bbox (see DESIGN.md "reach")
__device__ __forceinline__ int line_reach(double half_width) {
    double hw = (half_width > 0.0) ? half_width : 0.0;  // NaN -> 0
    if (hw > 1.0e6) hw = 1.0e6;
    // proven bound: a drawn pixel is < hw + 1.21 minor steps and <= that + 1 major corrections away from its main
    // pixel, an extra perpendicular starts one pixel off: ceil(hw) + 2 (tests/test_closed_forms.py); + 1 spare
    return (int)ceil(hw) + 3;
}

__device__ __forceinline__ short clamp_s(int v, int lo, int hi) { return (short)(v < lo ? lo : (v > hi ? hi : v)); }

// ------------------------------------------------------------------------------------------------------
// style_calc_kernel: one thread per (style, line pass): OpacityCalculator::new for the op's dashes and for the outer
// caps (line.rs:21-22, opacity_calculator.rs:16-30,98-143)
// ------------------------------------------------------------------------------------------------------
__global__ void style_calc_kernel(Scene s, uint4* table) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2u * s.n_styles) return;
    const osmr_style& st = s.styles[i >> 1];
    LineParams lp;
    if (!line_params(s, st, 1 + (int)(i & 1u), lp)) return;
    const double hw = lp.width / 2.0;
    OpacityCalc* cmain = reinterpret_cast<OpacityCalc*>(table + (size_t)i * kCalcEntryUnits);
    OpacityCalc* ccap = reinterpret_cast<OpacityCalc*>(table + (size_t)i * kCalcEntryUnits + kMainCalcUnits);
    unsigned cap_for_dashes = (s.flags & OSMR_DRAW_USE_CAPS_FOR_DASHES) ? lp.cap : (unsigned)OSMR_CAP_NONE;
    build_calc(*cmain, hw, lp.dashes, lp.n_dashes, (double)s.scale, lp.has_dashes, cap_for_dashes, kMaxDashSegs);
    const double zero = 0.0;
    build_calc(*ccap, hw, &zero, 1, 1.0, true, lp.cap, 1);
}

// ------------------------------------------------------------------------------------------------------
// plan_ops_kernel: one CTA per (tile, pass); ordered compaction of the visible generations of that pass (a8).
// The passes of a tile fill three separate sub-lists (pass p of a tile with n styled areas owns vis[3*base + p*n ..]), which
// raster_kernel walks one after the other: three times the CTAs of a per-tile kernel (a chunk of 256 tiles no longer
// leaves half of the SMs without a plan CTA) and no dependency between the passes.
// ------------------------------------------------------------------------------------------------------
// Low zooms over a large image put millions of styled areas into a handful of tiles (C4: z10-z12).  The host then cuts every
// (tile, pass) sub-list into `plan_slices` slices of whole 256-generation steps with a CTA each: plan_count_kernel counts the
// visible generations per slice, plan_scan_kernel turns the counts into the slices' first list positions, and the slices
// compact into the same ordered list a single CTA would have written.  (In an attempt that overflows a scratch allocator an op
// is dropped AFTER it was counted and leaves a hole -- such an attempt is discarded by the host and its consumers stop at the
// overflow flag.)
// ------------------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 256;

template <bool kCountOnly>
__device__ __forceinline__ void plan_ops_body(const Scene& s) {
    __shared__ unsigned warp_cnt[kPlanThreads / 32];
    __shared__ unsigned running;
    const unsigned S = s.plan_slices, list_id = blockIdx.x / S, slice = blockIdx.x % S;
    const unsigned t = list_id / 3u, the_pass = list_id % 3u;
    unsigned base = s.area_begin[t];
    unsigned n = s.area_begin[t + 1] - base;
    const unsigned g_begin = the_pass * n, total = g_begin + n;
    // the slice's generations: whole steps of kPlanThreads, so that the steps of all slices are the steps of the single-CTA loop
    const unsigned steps = (n + kPlanThreads - 1u) / kPlanThreads, steps_per_slice = (steps + S - 1u) / S;
    const unsigned long long s_lo = (unsigned long long)g_begin + (unsigned long long)slice * steps_per_slice * kPlanThreads;
    const unsigned long long s_hi = min((unsigned long long)total, s_lo + (unsigned long long)steps_per_slice * kPlanThreads);
    const unsigned long long list0 = 3ull * base + (unsigned long long)the_pass * n;  // first slot of this pass's sub-list
    VisOp* vis = s.vis + list0;
    const int D = s.D;
    const TileXform xf = make_xform(s.tiles[t]);
    unsigned long long refs_total = 0;
    const unsigned first_pos = (S > 1u && !kCountOnly) ? s.plan_slice_base[blockIdx.x] : 0u;
    if (threadIdx.x == 0) running = first_pos;
    __syncthreads();
    for (unsigned long long start64 = s_lo; start64 < s_hi; start64 += kPlanThreads) {
        const unsigned start = (unsigned)start64;
        unsigned g = start + threadIdx.x;
        bool visible = false;
        unsigned long long refs = 0;
        VisOp op;
        RasterOp rop;
        rop.a = rop.b = 0;
        rop.y0 = rop.y1 = 0;
        rop.kind = 0;
        rop.reach = 0;
        rop.rgb[0] = rop.rgb[1] = rop.rgb[2] = 0;
        rop.pad = 0;
        rop.opacity = 1.0;
        unsigned geom_units = 0, mask_words = 0, pair_cells = 0;
        if (g < total) {
            unsigned pass = g / n;
            unsigned i = g - pass * n;
            osmr_styled_area ar = s.areas[base + i];
            bool is_mp = (ar.entity & OSMR_AREA_MULTIPOLYGON) != 0;
            // pixel bbox + point count of the area in this tile: two projected corners of the entity's resident box
            AreaInfo info;
            info.x0 = info.y0 = 0x7fffffff;
            info.x1 = info.y1 = (int)0x80000000;
            info.npts = 0;
            if (!entity_valid(s, ar.entity) || ar.style >= s.n_styles) {
                atomicOr(&s.counters[CNT_BAD_INPUT], 1u);
            } else {
                const EntBox eb = (is_mp ? s.mp_box : s.way_box)[ar.entity & ~OSMR_AREA_MULTIPOLYGON];
                entity_pixel_bbox(eb, xf, info.x0, info.y0, info.x1, info.y1);
                info.npts = eb.npts;
                if (the_pass == 0) refs = eb.npts;  // statistics: R = node references of the batch
            }
            if (info.npts >= 2 && ar.style < s.n_styles) {
                const osmr_style& st = s.styles[ar.style];
                int x0 = info.x0, y0 = info.y0, x1 = info.x1, y1 = info.y1;
                bool active = false;
                if (pass == 0) {  // drawer.rs:172-185
                    if (st.flags & OSMR_STYLE_FILL_COLOR) {
                        active = true;
                        op.kind = OP_FILL_COLOR;
                        rop.opacity = (st.flags & OSMR_STYLE_FILL_OPACITY) ? st.fill_opacity : 1.0;
                        rop.rgb[0] = st.fill_color[0];
                        rop.rgb[1] = st.fill_color[1];
                        rop.rgb[2] = st.fill_color[2];
                    } else if ((st.flags & OSMR_STYLE_FILL_IMAGE) && st.fill_image >= 0 && (unsigned)st.fill_image < s.n_icons) {
                        active = true;
                        op.kind = OP_FILL_IMAGE;
                        rop.b = (unsigned)st.fill_image;
                    }
                    if (active) {
                        geom_units = info.npts;  // <= npts-1 edge records of 16 bytes
                        int ya = max(y0, 0), yb = min(y1, D - 1);
                        if (yb >= ya && x0 <= D - 1 && x1 >= 0) {
                            int rbx0, rby0, nbx, nby;
                            op_block_rect(x0, y0, x1, y1, D, rbx0, rby0, nbx, nby);
                            mask_words = (unsigned)(nbx * nby) * (unsigned)(kBH / 2);  // 16 bits per block row
                            rop.reach = (unsigned short)(rbx0 | (nbx << 8));           // fills: first block column, block columns
                        }
                    }
                } else if (!is_mp) {  // draw_areas(.., use_multipolygons=false) (drawer.rs:98-99,144-150)
                    LineParams lp;
                    if (line_params(s, st, (int)pass, lp)) {
                        active = true;
                        op.kind = OP_LINE;
                        double hw = lp.width / 2.0;
                        int reach = line_reach(hw);
                        if (reach > 32767) atomicOr(&s.counters[CNT_BAD_INPUT], 2u);  // half width beyond 2^15 pixels
                        rop.reach = (unsigned short)reach;
                        rop.rgb[0] = lp.rgb[0];
                        rop.rgb[1] = lp.rgb[1];
                        rop.rgb[2] = lp.rgb[2];
                        if (is_non_trivial_cap(lp.cap)) reach = 2 * reach;  // outer cap lines start hw away
                        long long lx0 = (long long)x0 - reach, ly0 = (long long)y0 - reach;
                        long long lx1 = (long long)x1 + reach, ly1 = (long long)y1 + reach;
                        x0 = (int)max(lx0, -2147483647LL);
                        y0 = (int)max(ly0, -2147483647LL);
                        x1 = (int)min(lx1, 2147483647LL);
                        y1 = (int)min(ly1, 2147483647LL);
                        geom_units = 4u * (info.npts + 1u);  // <= npts-1 segments + 2 caps of 64 bytes
                    }
                }
                if (active && x0 <= D - 1 && x1 >= 0 && y0 <= D - 1 && y1 >= 0) {
                    visible = true;
                    op.g = g;
                    op.x0 = clamp_s(x0, -1, D);
                    op.y0 = clamp_s(y0, -1, D);
                    op.x1 = clamp_s(x1, -1, D);
                    op.y1 = clamp_s(y1, -1, D);
                    op.geom_cnt = 0;
                    op.geom_off = 0;
                    op.mask_off = 0;
                    op.area = i;
                    op.pass = pass;
                    op.tile = t;
                    op.entity = ar.entity;
                    op.style = ar.style;
                    if (op.kind == OP_LINE)  // one pair-table cell per block of the reach bbox inside the tile
                        pair_cells = (unsigned)((min(x1, D - 1) / kBW - max(x0, 0) / kBW + 1) * (min(y1, D - 1) / kBH - max(y0, 0) / kBH + 1));
                }
            }
        }
        if (kCountOnly) {  // (block-uniform: every thread of the CTA takes this branch)
            const unsigned bal = __ballot_sync(0xffffffffu, visible);
            if (lane_id() == 0 && bal) atomicAdd(&running, (unsigned)__popc(bal));
            continue;
        }
        // scratch allocation (order irrelevant); failure marks the op invisible and raises the overflow flag.  Every allocator sees
        // the requests of ALL visible ops even when an earlier one has already failed: one attempt then reports the full need
        // of each, and the host grows them all before the redo.
        {
            const bool want = visible;
            const unsigned off = warp_alloc(&s.counters[CNT_GEOM_USED], want ? geom_units : 0u);
            const unsigned moff = warp_alloc(&s.counters[CNT_MASK_USED], want ? mask_words : 0u);
            const unsigned poff = warp_alloc(&s.counters[CNT_PAIR_USED], want ? pair_cells : 0u);
            if (want) {
                if (off + geom_units > s.geom_cap || off + geom_units < off) {
                    atomicOr(&s.counters[CNT_OVERFLOW], 1u);
                    visible = false;
                } else {
                    op.geom_off = off;
                }
                if (mask_words) {
                    if (moff + mask_words > s.mask_cap || moff + mask_words < moff) {
                        atomicOr(&s.counters[CNT_OVERFLOW], 2u);
                        visible = false;
                    } else {
                        op.mask_off = moff;
                    }
                }
                if (pair_cells) {
                    if (poff + pair_cells > s.pair_cap || poff + pair_cells < poff) {
                        atomicOr(&s.counters[CNT_OVERFLOW], 128u);
                        visible = false;
                    } else {
                        op.mask_off = poff;
                    }
                }
            }
        }
        refs_total += refs;
        // ordered compaction
        unsigned bal = __ballot_sync(0xffffffffu, visible);
        unsigned w = threadIdx.x >> 5;
        if (lane_id() == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        unsigned pos = running;
        for (unsigned k = 0; k < w; ++k) pos += warp_cnt[k];
        pos += __popc(bal & ((1u << lane_id()) - 1u));
        if (visible) {
            vis[pos] = op;
            rop.a = (op.kind == OP_LINE) ? op.geom_off : op.mask_off;
            rop.y0 = op.y0;
            rop.y1 = op.y1;
            rop.kind = (unsigned char)op.kind;
            s.rop[list0 + pos] = rop;
            s.vis_bbox[list0 + pos] = make_short4(op.x0, op.y0, op.x1, op.y1);
        }
        {  // work lists of the per-op kernels (their order is irrelevant)
            const unsigned gi = (unsigned)(list0 + pos);
            const bool is_line = visible && op.kind == OP_LINE, is_fill = visible && op.kind != OP_LINE;
            // big ops are split so that no single warp becomes the tail of its kernel: a fill op into chunks of 32 mask rows,
            // a line op into batches of 32 segment records (at most npts - 1 segments + 2 cap lines)
            const unsigned n_fill = is_fill ? (unsigned)((min((int)op.y1, D - 1) - max((int)op.y0, 0)) / 32 + 1) : 0u;
            const unsigned n_line = is_line ? (geom_units / 4u + 31u) / 32u : 0u;
            const unsigned wa = warp_alloc(&s.counters[CNT_N_WORK], visible ? 1u : 0u);
            const unsigned wf = warp_alloc(&s.counters[CNT_N_FILL_WORK], n_fill);
            const unsigned wl = warp_alloc(&s.counters[CNT_N_LINE_WORK], n_line);
            if (visible) s.work[wa] = gi;
            if (n_fill) {
                if (wf + n_fill > s.fill_work_cap || wf + n_fill < wf)
                    atomicOr(&s.counters[CNT_OVERFLOW], 16u);
                else
                    for (unsigned j = 0; j < n_fill; ++j) s.fill_work[wf + j] = make_uint2(gi, j);
            }
            if (n_line) {
                if (wl + n_line > s.line_work_cap || wl + n_line < wl)
                    atomicOr(&s.counters[CNT_OVERFLOW], 32u);
                else
                    for (unsigned j = 0; j < n_line; ++j) s.line_work[wl + j] = make_uint2(gi, j);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned add = 0;
            for (unsigned k = 0; k < kPlanThreads / 32; ++k) add += warp_cnt[k];
            running += add;
        }
        __syncthreads();
    }
    if (kCountOnly) {
        __syncthreads();
        if (threadIdx.x == 0) s.plan_slice_base[blockIdx.x] = running;
        return;
    }
    if (the_pass == 0) {
        for (int o = 16; o > 0; o >>= 1) refs_total += __shfl_down_sync(0xffffffffu, refs_total, o);
        if (lane_id() == 0 && refs_total) {
            unsigned lo = (unsigned)refs_total;
            unsigned old = atomicAdd(&s.counters[CNT_NODE_REFS_LO], lo);
            if (old + lo < old) atomicAdd(&s.counters[CNT_NODE_REFS_HI], 1u);
            atomicAdd(&s.counters[CNT_NODE_REFS_HI], (unsigned)(refs_total >> 32));
        }
    }
    if (threadIdx.x == 0) {
        if (slice == S - 1u) s.vis_count[list_id] = running;  // [3 * tile + pass]: the last slice ends where the list ends
        atomicAdd(&s.counters[CNT_VISIBLE], running - first_pos);
    }
}

__global__ void __launch_bounds__(kPlanThreads) plan_ops_kernel(Scene s) { plan_ops_body<false>(s); }
__global__ void __launch_bounds__(kPlanThreads) plan_count_kernel(Scene s) { plan_ops_body<true>(s); }

// one thread per (tile, pass): counts of its slices -> first list positions
__global__ void plan_scan_kernel(Scene s) {
    const unsigned l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= 3u * s.n_tiles) return;
    unsigned* c = s.plan_slice_base + (size_t)l * s.plan_slices;
    unsigned acc = 0;
    for (unsigned k = 0; k < s.plan_slices; ++k) {
        const unsigned v = c[k];
        c[k] = acc;
        acc += v;
    }
}

// ------------------------------------------------------------------------------------------------------
// build_geometry_kernel: one warp per visible op (dynamic work fetch)
// ------------------------------------------------------------------------------------------------------
constexpr int kGeomThreads = 128;
constexpr unsigned kLaneFillMax = 48;  // nodes of a way whose fill edges are produced by a single lane
#ifndef OSMR_GEOM_LANE_FILLS
#define OSMR_GEOM_LANE_FILLS 1
#endif

#ifndef OSMR_GEOM_MIN_BLOCKS
#define OSMR_GEOM_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(kGeomThreads, OSMR_GEOM_MIN_BLOCKS) build_geometry_kernel(Scene s) {
    const unsigned lane = lane_id();
    const int D = s.D;
    const unsigned n_work = s.counters[CNT_N_WORK];
    for (;;) {
        // 32 ops per fetch, one per lane.  A fill of a short way (a building: the majority of all ops) is done by its lane
        // alone -- 32 such ops side by side instead of one op on a quarter of the lanes; everything else (lines,
        // multipolygons, long rings) is then taken in turn by the whole warp.
        unsigned wi0 = 0;
        if (lane == 0) wi0 = atomicAdd(&s.counters[CNT_WORK_CURSOR], 32u);
        wi0 = __shfl_sync(0xffffffffu, wi0, 0);
        if (wi0 >= n_work) break;
        const bool have = wi0 + lane < n_work;
        const unsigned my_gi = have ? s.work[wi0 + lane] : 0u;
        bool lane_done = !have;
#if OSMR_GEOM_LANE_FILLS
        if (have) {
            VisOp& o = s.vis[my_gi];
            if (o.kind != OP_LINE && !(o.entity & OSMR_AREA_MULTIPOLYGON)) {
                const uint2 r = s.ways[o.entity];
                if (r.y <= kLaneFillMax) {
                    const TileXform xf1 = make_xform(s.tiles[o.tile]);
                    int4* out1 = reinterpret_cast<int4*>(s.geom + o.geom_off);
                    unsigned n1 = 0;
                    int2 prev = make_int2(0, 0);
                    for (unsigned q = 0; q < r.y; ++q) {
                        const int2 cur = project_point(s.merc[s.ints[r.x + q]], xf1);
                        if (q && prev.y != cur.y && max(prev.y, cur.y) >= 0 && min(prev.y, cur.y) <= D - 1)
                            out1[n1++] = make_int4(prev.x, prev.y, cur.x, cur.y);
                        prev = cur;
                    }
                    o.geom_cnt = n1;
                    lane_done = true;
                    OSMR_COUNT("geom.lane_fills", 1);
                }
            }
        }
#endif
        unsigned rest = __ballot_sync(0xffffffffu, !lane_done);
        while (rest) {
        const int src_lane = __ffs(rest) - 1;
        rest &= rest - 1;
        unsigned gi = __shfl_sync(0xffffffffu, my_gi, src_lane);
        VisOp& op = s.vis[gi];
        const unsigned pass = op.pass;
        osmr_styled_area ar;
        ar.entity = op.entity;
        ar.style = op.style;
        TileXform xf = make_xform(s.tiles[op.tile]);
        RingIter it(s, ar.entity);
        unsigned count = 0;
        if (op.kind != OP_LINE) {
            // a2 prologue: edges that visit at least one tile row and are not horizontal (a horizontal edge is
            // poisoned on its only row, fill.rs:66-72, so it never takes part in the pairing)
            int4* out = reinterpret_cast<int4*>(s.geom + op.geom_off);
            for (unsigned k = 0; k < it.n_rings; ++k) {
                uint2 r = it.ring(k);
                for (unsigned b = 0; b + 1 < r.y; b += 32) {
                    unsigned e = b + lane;
                    bool keep = false;
                    int4 rec = make_int4(0, 0, 0, 0);
                    if (e + 1 < r.y) {
                        int2 p1 = project_point(s.merc[s.ints[r.x + e]], xf);
                        int2 p2 = project_point(s.merc[s.ints[r.x + e + 1]], xf);
                        rec = make_int4(p1.x, p1.y, p2.x, p2.y);
                        keep = (p1.y != p2.y) && max(p1.y, p2.y) >= 0 && min(p1.y, p2.y) <= D - 1;
                    }
                    unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) out[count + __popc(bal & ((1u << lane) - 1u))] = rec;
                    count += __popc(bal);
                }
            }
        } else {
            // a4 prologue (line.rs:9-61): segments in order, running dash phase, outer caps
            const osmr_style& st = s.styles[ar.style];
            LineParams lp;
            line_params(s, st, (int)pass, lp);
            double hw = lp.width / 2.0;
            int reach = line_reach(hw);
            bool caps = is_non_trivial_cap(lp.cap);
            bool dashed = lp.has_dashes && lp.n_dashes > 0;
            SegRec* out = reinterpret_cast<SegRec*>(s.geom + op.geom_off);
            uint2 r = it.ring(0);
            double acc = 0.0;  // traveled_distance
            unsigned n_pairs = r.y ? r.y - 1 : 0;
            auto make_rec = [&](int ax, int ay, int bx, int by, double trav, unsigned is_cap) {
                SegRec rec;
                rec.x1 = ax;
                rec.y1 = ay;
                rec.x2 = bx;
                rec.y2 = by;
                rec.traveled = trav;
                int dx = abs(wsub(bx, ax)), dy = abs(wsub(by, ay));
                int mxd = max(dx, dy);
                const double dxf = (double)dx, dyf = (double)dy;
                rec.denom = sqrt(dyf * dyf + dxf * dxf);
                const int big = 1 << 24;
                bool small = ax > -big && ax < big && ay > -big && ay < big && bx > -big && bx < big && by > -big && by < big;
                bool fast = mxd < 32768 && mxd > 0;  // numerators 2*mn_d*n stay below 2^31
                rec.magic = fast ? (0xffffffffffffffffull / (unsigned long long)(2 * mxd)) + 1ull : 0ull;
                rec.flags = (is_cap ? 1u : 0u) | (fast ? 2u : 0u) | (small ? 4u : 0u);
                if (!small) atomicOr(&s.counters[CNT_BIG_COORDS], 1u);
                // main steps k (line.rs:133-158) whose major coordinate lies within `reach` of the tile: the only ones
                // whose perpendiculars can put a pixel into it
                const bool swp = dx > dy;
                const int mx0 = swp ? ax : ay;
                const int mx_inc = swp ? (ax <= bx ? 1 : -1) : (ay <= by ? 1 : -1);
                const long long lo = -(long long)reach, hi = (long long)D - 1 + reach;
                long long ka = mx_inc > 0 ? lo - mx0 : (long long)mx0 - hi;
                long long kb = mx_inc > 0 ? hi - mx0 : (long long)mx0 - lo;
                if (ka < 0) ka = 0;
                if (kb > mxd) kb = mxd;
                rec.k0 = (int)ka;
                rec.n_k = kb >= ka ? (unsigned)(kb - ka + 1) : 0u;
                rec.pad0 = 0;
                rec.pad1 = 0;
                return rec;
            };
            for (unsigned b = 0; b < n_pairs; b += 32) {
                unsigned e = b + lane;
                bool valid = e < n_pairs;
                int2 p1 = make_int2(0, 0), p2 = make_int2(0, 0);
                double d = 0.0;
                if (valid) {
                    p1 = project_point(s.merc[s.ints[r.x + e]], xf);
                    p2 = project_point(s.merc[s.ints[r.x + e + 1]], xf);
                    if (dashed) d = point_dist(p1.x, p1.y, p2.x, p2.y);
                }
                double trav = 0.0;
                if (dashed) {  // sequential f64 summation, same order as add_traveled_distance (line.rs:31)
                    const int nj = (int)min(32u, n_pairs - b);
                    for (int j = 0; j < nj; ++j) {
                        double dj = __shfl_sync(0xffffffffu, d, j);
                        if ((int)lane == j) trav = acc;
                        acc = acc + dj;
                    }
                }
                bool nondeg = valid && (p1.x != p2.x || p1.y != p2.y);
                auto touches = [&](int ax, int ay, int bx, int by) {
                    long long mnx = (long long)min(ax, bx) - reach, mxx = (long long)max(ax, bx) + reach;
                    long long mny = (long long)min(ay, by) - reach, mxy = (long long)max(ay, by) + reach;
                    return mnx <= D - 1 && mxx >= 0 && mny <= D - 1 && mxy >= 0;
                };
                bool keep = nondeg && touches(p1.x, p1.y, p2.x, p2.y);
                // outer caps: only for a non-degenerate first / last pair (line.rs:33-57)
                bool first_cap = caps && nondeg && e == 0;
                bool last_cap = caps && nondeg && e + 1 == n_pairs;
                int2 c1 = make_int2(0, 0), c2 = make_int2(0, 0);
                bool k1 = false, k2 = false;
                if (first_cap) {
                    c1 = push_away_from(p1.x, p1.y, p2.x, p2.y, hw);
                    k1 = (c1.x != p1.x || c1.y != p1.y) && touches(p1.x, p1.y, c1.x, c1.y);
                }
                if (last_cap) {
                    c2 = push_away_from(p2.x, p2.y, p1.x, p1.y, hw);
                    k2 = (c2.x != p2.x || c2.y != p2.y) && touches(p2.x, p2.y, c2.x, c2.y);
                }
                SegRec r0, r1, r2;
                if (keep) r0 = make_rec(p1.x, p1.y, p2.x, p2.y, trav, 0u);
                if (k1) r1 = make_rec(p1.x, p1.y, c1.x, c1.y, 0.0, 1u);
                if (k2) r2 = make_rec(p2.x, p2.y, c2.x, c2.y, 0.0, 1u);
                unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) out[count + __popc(bal & ((1u << lane) - 1u))] = r0;
                count += __popc(bal);
                unsigned b1 = __ballot_sync(0xffffffffu, k1);
                unsigned b2 = __ballot_sync(0xffffffffu, k2);
                if (k1) out[count] = r1;
                count += b1 ? 1u : 0u;
                if (k2) out[count] = r2;
                count += b2 ? 1u : 0u;
            }
        }
        if (lane == 0) {
            op.geom_cnt = count;
            if (op.kind == OP_LINE) s.rop[gi].b = count;
        }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// fill_rows_kernel (a2 + a3): even-odd coverage of every row of every visible fill op, 1 bit per pixel.
//
// Per row (fill.rs:23-46): drop poisoned spans, stable sort by x_min, pair (0,1),(2,3).., fill
// [max(e1.x_min,0), min(e2.x_max,D-1)].  One warp owns one row; lanes stream the op's edge records.
// Up to fill_cap spans are ranked in shared memory.  Rows with more spans use an equivalent order-free
// form that needs only counts (DESIGN.md "even-odd without sorting"):
//   covered(x) = [c(x) odd and c(x) < m]  or  [x in span(e) for some e of odd rank]
// with c(x) = #{spans with x_min <= x}, rank(e) = #{e' : (x_min', idx') < (x_min, idx)}.
// ------------------------------------------------------------------------------------------------------
constexpr int kFillThreads = 128;
constexpr int kFillWarps = kFillThreads / 32;
constexpr int kMaxD = 2048;  // scale <= 8
// Row-parallel path of fill_rows_kernel: a lane owns a ROW of the work item's 32 and walks the op's edges itself (the edge
// record is a broadcast load), keeping up to kRowLocal spans.  An ordinary polygon (a building: 5-13 edges, 2-4 spans per row)
// keeps all 32 lanes busy this way; with a lane per edge and a loop over the rows only ne of 32 lanes worked.  Rows with more
// spans, tiles wider than 32 * kRowWords pixels and the fill_cap test hook take the row-serial path below.
constexpr int kRowLocal = 8;
constexpr int kRowWords = 16;
#ifndef OSMR_FILL_ROW_PARALLEL
#define OSMR_FILL_ROW_PARALLEL 1
#endif

__device__ __forceinline__ unsigned bits_in_word(int from, int to, int w) {
    // bits of [from, to] that fall into word w (pixels 32w .. 32w+31); from <= to
    int lo = max(from, 32 * w), hi = min(to, 32 * w + 31);
    if (lo > hi) return 0u;
    unsigned n = (unsigned)(hi - lo + 1);
    unsigned m = (n >= 32u) ? 0xffffffffu : ((1u << n) - 1u);
    return m << (lo - 32 * w);
}

__global__ void __launch_bounds__(kFillThreads) fill_rows_kernel(Scene s) {
    __shared__ int2 act[kFillWarps][kFillCap];
    __shared__ int2 sorted[kFillWarps][kFillCap];
    __shared__ unsigned hist[kFillWarps][kMaxD / 8 + 1];  // slow path: x_min histogram in 8 passes of D/8 columns
    __shared__ int2 rspan[kFillWarps][kRowLocal][32];      // row-parallel path: the spans of row `lane`, in edge order
    __shared__ unsigned rmask[kFillWarps][kRowWords][32];  // ... and its mask words
    const unsigned lane = lane_id();
    const unsigned w = threadIdx.x >> 5;
    const int D = s.D;
    const int wpr = D / 32;
    const int cap = s.fill_cap;
    // A scratch or work-list overflow means the host grows the buffer and redoes the draw: the work lists may have holes and
    // the scratch of some ops does not exist, so the rest of this attempt is skipped instead of reading garbage.
    if (s.counters[CNT_OVERFLOW]) return;
    const unsigned n_work = s.counters[CNT_N_FILL_WORK];
    // every warp is an independent worker: fetch fill ops, produce all their rows, repeat (no CTA barrier)
    unsigned wi = 0, wi_end = 0;
    for (;; ++wi) {
        if (wi >= wi_end) {  // one op per fetch: the rows of a fill op are too uneven a load for batches (0.78 -> 0.85 ms with 4)
            if (lane == 0) wi = atomicAdd(&s.counters[CNT_FILL_CURSOR], 1u);
            wi = __shfl_sync(0xffffffffu, wi, 0);
            wi_end = wi + 1u;
        }
        if (wi >= n_work) break;
        const uint2 item = s.fill_work[wi];  // (op, chunk of 32 rows)
        const VisOp op = s.vis[item.x];
        const int4* edges = reinterpret_cast<const int4*>(s.geom + op.geom_off);
        const int ne = (int)op.geom_cnt;
        const int ya = max((int)op.y0, 0), yb = min((int)op.y1, D - 1);
        const int y_first = ya + 32 * (int)item.y, y_last = min(yb, y_first + 31);
        int rbx0, rby0, nbx, nby;
        op_block_rect(op.x0, op.y0, op.x1, op.y1, D, rbx0, rby0, nbx, nby);
        if (OSMR_FILL_ROW_PARALLEL && wpr <= kRowWords && cap == kFillCap) {
            const int y = y_first + (int)lane;
            const bool live = y <= y_last;
            int m = 0;
            for (int e = 0; e < ne; ++e) {
                const int4 ed = edges[e];
                int xmin, xmax;
                bool poisoned;
                if (live && fill_edge_row_span(ed.x, ed.y, ed.z, ed.w, y, xmin, xmax, poisoned) && !poisoned) {
                    if (m < kRowLocal) rspan[w][m][lane] = make_int2(xmin, xmax);
                    ++m;
                }
            }
            if (!__any_sync(0xffffffffu, m > kRowLocal)) {
                // stable insertion sort by x_min (fill.rs:25), then pair (0,1), (2,3), ... (fill.rs:27-45)
                for (int i = 1; i < m; ++i) {
                    const int2 v = rspan[w][i][lane];
                    int j = i - 1;
                    while (j >= 0 && rspan[w][j][lane].x > v.x) {
                        rspan[w][j + 1][lane] = rspan[w][j][lane];
                        --j;
                    }
                    rspan[w][j + 1][lane] = v;
                }
                for (int k = 0; k < wpr; ++k) rmask[w][k][lane] = 0u;
                for (int q = 0; 2 * q + 1 < m; ++q) {
                    const int from = max(rspan[w][2 * q][lane].x, 0);
                    const int to = min(rspan[w][2 * q + 1][lane].y, D - 1);
                    if (from <= to)
                        for (int wd = from >> 5; wd <= (to >> 5); ++wd) rmask[w][wd][lane] |= bits_in_word(from, to, wd);
                }
                if (live) {
                    for (int k = rbx0 / 2; k <= (rbx0 + nbx - 1) / 2; ++k) store_mask_word(s.mask + op.mask_off, rbx0, rby0, nbx, y, k, rmask[w][k][lane]);
                }
                continue;
            }
        }
        for (int y = y_first; y <= y_last; ++y) {
            // ---- gather the non-poisoned spans of this row, in edge order ----
            int m = 0;
            for (int b = 0; b < ne; b += 32) {
                int e = b + (int)lane;
                bool active = false;
                int xmin = 0, xmax = 0;
                if (e < ne) {
                    int4 ed = edges[e];
                    bool poisoned;
                    if (fill_edge_row_span(ed.x, ed.y, ed.z, ed.w, y, xmin, xmax, poisoned)) active = !poisoned;
                }
                unsigned bal = __ballot_sync(0xffffffffu, active);
                if (active) {
                    int pos = m + __popc(bal & ((1u << lane) - 1u));
                    if (pos < cap) act[w][pos] = make_int2(xmin, xmax);
                }
                m += __popc(bal);
            }
            __syncwarp();
            unsigned word = 0;  // lane l < wpr owns mask word l (wpr <= 64 -> two words per lane when D > 1024)
            unsigned word2 = 0;
            if (m <= cap) {
                // stable rank by x_min (ties keep edge order), fill.rs:25
                for (int i = (int)lane; i < m; i += 32) {
                    int2 me = act[w][i];
                    int r = 0;
                    for (int j = 0; j < m; ++j) {
                        int xj = act[w][j].x;
                        r += (xj < me.x || (xj == me.x && j < i)) ? 1 : 0;
                    }
                    sorted[w][r] = me;
                }
                __syncwarp();
                for (int q = 0; 2 * q + 1 < m; ++q) {
                    int from = max(sorted[w][2 * q].x, 0);
                    int to = min(sorted[w][2 * q + 1].y, D - 1);
                    if (from <= to) {
                        word |= bits_in_word(from, to, (int)lane);
                        word2 |= bits_in_word(from, to, (int)lane + 32);
                    }
                }
            } else {
                // ---- streaming path: counts only, no storage proportional to m ----
                // clause 1: c(x) parity, c(x) < m.  hist[] holds #{x_min == column} for one window of columns.
                const int win = kMaxD / 8;
                unsigned carry = 0;  // #{x_min < window start}
                for (int c0 = 0; c0 < D; c0 += win) {
                    for (int i = (int)lane; i <= win; i += 32) hist[w][i] = 0;
                    __syncwarp();
                    for (int b = 0; b < ne; b += 32) {
                        int e = b + (int)lane;
                        if (e < ne) {
                            int4 ed = edges[e];
                            int xmin, xmax;
                            bool poisoned;
                            if (fill_edge_row_span(ed.x, ed.y, ed.z, ed.w, y, xmin, xmax, poisoned) && !poisoned) {
                                int col = max(xmin, 0);  // x_min <= 0 counts for every pixel
                                if (col >= c0 && col < c0 + win) atomicAdd(&hist[w][col - c0], 1u);
                            }
                        }
                    }
                    __syncwarp();
                    // lane handles the columns of the words it owns inside this window
                    for (int k = 0; k < 2; ++k) {
                        int wd = (int)lane + 32 * k;
                        int x_lo = 32 * wd;
                        if (x_lo >= c0 && x_lo < min(c0 + win, D)) {
                            unsigned c = carry;
                            for (int i = 0; i < x_lo - c0; ++i) c += hist[w][i];
                            unsigned bits = 0;
                            for (int bpos = 0; bpos < 32; ++bpos) {
                                c += hist[w][x_lo - c0 + bpos];
                                if ((c & 1u) && c < (unsigned)m) bits |= 1u << bpos;
                            }
                            if (k == 0)
                                word |= bits;
                            else
                                word2 |= bits;
                        }
                    }
                    unsigned tot = 0;
                    for (int i = (int)lane; i < win; i += 32) tot += hist[w][i];
                    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    carry += tot;
                    __syncwarp();
                }
                // clause 2: own spans of odd-ranked edges that reach the tile columns
                for (int b = 0; b < ne; b += 32) {
                    int e = b + (int)lane;
                    bool cand = false;
                    int xmin = 0, xmax = 0;
                    if (e < ne) {
                        int4 ed = edges[e];
                        bool poisoned;
                        if (fill_edge_row_span(ed.x, ed.y, ed.z, ed.w, y, xmin, xmax, poisoned))
                            cand = !poisoned && xmin <= D - 1 && xmax >= 0;
                    }
                    unsigned cb = __ballot_sync(0xffffffffu, cand);
                    while (cb) {
                        int src = __ffs(cb) - 1;
                        cb &= cb - 1;
                        int txmin = __shfl_sync(0xffffffffu, xmin, src);
                        int txmax = __shfl_sync(0xffffffffu, xmax, src);
                        int te = b + src;
                        unsigned rank = 0;
                        for (int b2 = 0; b2 < ne; b2 += 32) {
                            int e2 = b2 + (int)lane;
                            bool less = false;
                            if (e2 < ne) {
                                int4 ed = edges[e2];
                                int xm, xM;
                                bool poisoned;
                                if (fill_edge_row_span(ed.x, ed.y, ed.z, ed.w, y, xm, xM, poisoned) && !poisoned)
                                    less = (xm < txmin) || (xm == txmin && e2 < te);
                            }
                            rank += __popc(__ballot_sync(0xffffffffu, less));
                        }
                        if (rank & 1u) {
                            int from = max(txmin, 0), to = min(txmax, D - 1);
                            word |= bits_in_word(from, to, (int)lane);
                            word2 |= bits_in_word(from, to, (int)lane + 32);
                        }
                    }
                }
            }
            if ((int)lane < wpr) store_mask_word(s.mask + op.mask_off, rbx0, rby0, nbx, y, (int)lane, word);
            if ((int)lane + 32 < wpr) store_mask_word(s.mask + op.mask_off, rbx0, rby0, nbx, y, (int)lane + 32, word2);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Perpendicular walks (a4/a5).  The reference draws a thick line by walking, from every pixel k of the segment's
// Bresenham main line, one perpendicular Bresenham line to each side until the opacity calculator says "not in line"
// (line.rs:65-158).  Here every walk is evaluated ONCE per tile by line_cover_kernel, which appends each covered pixel
// as a fragment (pixel, alpha) to the list of the 16x16 block it falls into; raster_kernel's blocks only stream their
// lists.  (History: up to round 1's v6 every block evaluated the walks itself -- two thirds of all opacity evaluations
// fell outside the evaluating block; v7-v11 cached the alphas per walk and the blocks replayed the integer stepping.)
// ------------------------------------------------------------------------------------------------------
struct WalkItem {  // line.rs:65-118,133-158: the state of the main line at step k
    bool swap;
    int mn_inc, mx_inc, mn_d, mx_d;
    int mn, mx;    // main-line pixel at step k (minor / major coordinate)
    int p_error;   // i32 (wrapping) like the reference's `p_error`
    bool extra;    // step k also starts the extra perpendicular of a double correction (line.rs:150-155)
};

__device__ __forceinline__ void walk_item_setup(int x1, int y1, int x2, int y2, unsigned flags, unsigned long long magic, int k,
                                                WalkItem& w) {
    const int dx = abs(wsub(x2, x1)), dy = abs(wsub(y2, y1));
    w.swap = dx > dy;
    const int mn0 = w.swap ? y1 : x1, mx0 = w.swap ? x1 : y1;
    w.mn_d = w.swap ? dy : dx;
    w.mx_d = w.swap ? dx : dy;
    const int inc_x = (x1 <= x2) ? 1 : -1, inc_y = (y1 <= y2) ? 1 : -1;
    w.mn_inc = w.swap ? inc_y : inc_x;
    w.mx_inc = w.swap ? inc_x : inc_y;
    // Bresenham state at main step k (closed form; exact quotients via the per-segment magic)
    long long c, pc;
    if (flags & 2u) {
        int num = 2 * w.mn_d * k - w.mx_d;
        c = num <= 0 ? 0 : (long long)__umul64hi((unsigned long long)(unsigned)(num + 2 * w.mx_d - 1), magic);
        int num2 = 2 * w.mn_d * (int)c - w.mx_d;
        pc = num2 <= 0 ? 0 : (long long)__umul64hi((unsigned long long)(unsigned)(num2 + 2 * w.mx_d - 1), magic);
    } else {
        c = ncorr(0, k, w.mn_d, w.mx_d);
        pc = ncorr(0, c, w.mn_d, w.mx_d);
    }
    w.mx = mx0 + w.mx_inc * k;
    w.mn = mn0 + w.mn_inc * (int)c;
    w.p_error = (int)(2ll * w.mn_d * c - 2ll * w.mx_d * pc);
    const int e_main = (int)(2ll * w.mn_d * k - 2ll * w.mx_d * c);  // main error before step k
    w.extra = k < w.mx_d && wadd(e_main, 2 * w.mn_d) > w.mx_d && wadd(w.p_error, 2 * w.mn_d) > w.mx_d;
}

// per-segment f64 constants of line.rs:98-118
struct SegConst {
    int x1, y1;
    long long numer_const, sdx, sdy;
    double denom;
    double traveled;
    bool small;  // all coordinates < 2^24: `raw as f64` can be carried exactly in f64
};

// (An interior classification under round dash caps -- approximate dash phase with an error bound, proving that a pixel well
// inside the line gets opacity_mul without the two square roots, the division and the exact `%` -- was derived, verified
// bit-exact on the whole suite and removed again: 41 % of the dashed evaluations qualify, but the lanes of a warp rarely
// qualify together, so the warp paid the test and the reference path: line_cover_kernel 1.60 -> 2.52 ms.  DESIGN.md 4.)

// draw_one_perpendicular (line.rs:89-131): evaluates the walk from its start until the first pixel that is not in the
// line (or until it has left the tile for good on its monotone axis) and appends every in-line, in-tile step with a positive
// alpha as a fragment (pixel, alpha) to the list of its 16x16 block.  Returns the number of fragments stored.
struct FragSink {  // where the fragments of one line op go: its (block -> bin entry) pair table, the last entry looked up
    const unsigned* pair;   // the op's pair table (Scene.pair + VisOp.mask_off)
    const BinEntry* entries;
    unsigned* frag_cnt;
    double* frag_alpha;
    unsigned char* frag_pix;
    unsigned* counters;
    int bx0, by0, nbx, nby; // block rectangle of the op's reach bbox
    int last_block;         // block (by * 4096 + bx) of last_entry, -1: none yet
    unsigned last_entry, last_off, last_cap;
};

__device__ __forceinline__ void frag_put(FragSink& fs, int px, int py, double alpha) {
    const int bx = px / kBW, by = py / kBH;  // px, py are inside the tile
    const int key = by * 4096 + bx;
    if (key != fs.last_block) {
        const int cx = bx - fs.bx0, cy = by - fs.by0;
        unsigned e = 0xffffffffu;
        if (cx >= 0 && cx < fs.nbx && cy >= 0 && cy < fs.nby) e = fs.pair[cy * fs.nbx + cx];
        fs.last_block = key;
        fs.last_entry = e;
        if (e != 0xffffffffu) {
            const BinEntry be = fs.entries[e];
            fs.last_off = be.frag_off;
            fs.last_cap = be.frag_cap;
        }
    }
    if (fs.last_entry == 0xffffffffu) {  // the binning proved that no segment of this op reaches this block: a logic error, never silent
        atomicOr(&fs.counters[CNT_WALK_TRUNC], 2u);
        return;
    }
    const unsigned pos = atomicAdd(&fs.frag_cnt[fs.last_entry], 1u);
    if (pos >= fs.last_cap) {  // the capacity is a proven bound; if it ever fails the draw is refused, not clipped
        atomicOr(&fs.counters[CNT_OVERFLOW], 256u);
        return;
    }
    fs.frag_alpha[(size_t)fs.last_off + pos] = alpha;
    fs.frag_pix[(size_t)fs.last_off + pos] = (unsigned char)((py % kBH) * kBW + (px % kBW));
}

__device__ __forceinline__ unsigned cover_walk(FragSink& fs, unsigned S, const WalkItem& w, const SegConst& sc, const OpacityCalc& calc,
                                               int mn, int p_error, int mul, double opacity0, int D, unsigned* trunc_flag) {
    int p_mn = w.mx;
    int p_mx = mn;
    int err = mul * p_error;  // i32 like the reference (line.rs:80-91)
    // p_mx moves by mul*mn_inc every step: once it leaves the tile on that side it never comes back
    const int step = mul * w.mn_inc;
    const int corr = -mul * w.mx_inc;
    // raw = numer_const + sdy*px - sdx*py (i64, line.rs:102-103); its per-step increments are integers
    const long long d_step = w.swap ? -sc.sdx * step : sc.sdy * step;
    const long long d_corr = w.swap ? sc.sdy * corr : -sc.sdx * corr;
    long long raw;
    {
        int px = w.swap ? p_mn : p_mx, py = w.swap ? p_mx : p_mn;
        raw = sc.numer_const + (sc.sdy * (long long)px - sc.sdx * (long long)py);
    }
    // `raw as f64` (line.rs:104): when every coordinate is below 2^24 the value is an integer below 2^53 and is
    // carried in f64 exactly, which keeps 64-bit integer maths and the i64->f64 conversion out of the loop
    double fraw = (double)raw;
    const double f_step = (double)d_step, f_corr = (double)d_corr;
    const bool dashed = calc.n_segs != 0;
    // conservative thresholds that classify a pixel without the division (only when the feather is a per-op
    // constant, i.e. no round dash cap can change the half width):  |raw| < from*denom*(1-1e-12)  =>  cd = mul*1
    const bool quick = !(dashed && calc.round_caps);
    const double t_in = calc.feather_from * sc.denom * (1.0 - 9.0e-13);
    const double t_out = calc.feather_to * sc.denom * (1.0 + 9.0e-13);
    unsigned t = 0, n_put = 0;
    for (;;) {
        if (step > 0 ? (p_mx > D - 1) : (p_mx < 0)) break;
        const double araw = fabs(sc.small ? fraw : (double)raw);
        double opacity;
        bool in_line;
        if (quick && !dashed && araw < t_in) {
            opacity = fmin(1.0, calc.opacity_mul * 1.0);  // sd = 1.0, cd = mul * 1.0
            in_line = calc.opacity_mul * 1.0 > 0.0;
        } else if (araw > t_out) {
            // cd = mul * 0.0.  Also true under round dash caps: they can only shrink the half width (sqrt(hw^2 - cap^2) <= hw,
            // NaN -> feather_to 1.0 <= the op's), so a pixel beyond the op's feather_to is beyond every per-pixel one: the
            // evaluation that ends a walk -- a quarter of all evaluations -- needs neither the dash phase nor a square root
            in_line = false;
            opacity = 0.0;
        } else {
            double center_dist = div_pos_peeled(araw, sc.denom);
            double short_start = 0.0;
            if (dashed) {
                int px = w.swap ? p_mn : p_mx, py = w.swap ? p_mx : p_mn;
                double long_start = point_dist(px, py, sc.x1, sc.y1);
                short_start = sqrt_peeled(fmax(long_start * long_start - center_dist * center_dist, 0.0));
            }
            calc_opacity(calc, sc.traveled, center_dist, short_start, opacity, in_line);
        }
        OSMR_COUNT("cover.steps_evaluated", 1);
        if (!in_line) break;
        if (t >= S) {  // cannot happen (line_reach is a proven bound on the in-line steps); never continue silently
            atomicOr(trunc_flag, 1u);
            break;
        }
        {
            const double a = opacity0 * opacity;  // RgbaColor::from_color(color, initial_opacity * opacity).a
            const int px = w.swap ? p_mn : p_mx, py = w.swap ? p_mx : p_mn;
            if (a > 0.0 && (unsigned)px < (unsigned)D && (unsigned)py < (unsigned)D) {  // set_pixel ignores everything outside the tile
                frag_put(fs, px, py, a);
                ++n_put;
            }
        }
        ++t;
        // update_error (line.rs:82-91), wrapping i32 arithmetic
        if (wadd(err, 2 * w.mn_d) > w.mx_d) {
            err = wsub(err, 2 * w.mx_d);
            p_mn += corr;
            if (sc.small) fraw += f_corr; else raw += d_corr;
        }
        err = wadd(err, 2 * w.mn_d);
        p_mx += step;
        if (sc.small) fraw += f_step; else raw += d_step;
    }
    return n_put;
}

// ------------------------------------------------------------------------------------------------------
// line_cover_kernel: one warp per visible line op (dynamic fetch).  32 segment records at a time; their
// (main step, direction) pairs are spread over the lanes, every lane evaluates its walk(s) to the end.
// ------------------------------------------------------------------------------------------------------
#ifndef OSMR_COVER_BATCH
#define OSMR_COVER_BATCH 2  // 4: 1.594 ms, 2: 1.569, 1: 1.571 (C2 batch, same box)
#endif
constexpr unsigned kCoverBatch = OSMR_COVER_BATCH;  // work items a warp takes per cursor atomic
constexpr int kCoverThreads = 128;
constexpr int kCoverWarps = kCoverThreads / 32;

struct CoverSmem {
    OpacityCalc calc[2];  // [0] dashes of the op, [1] outer caps
    SegRec recs[32];
    unsigned pre[32];
};

#ifndef OSMR_COVER_MIN_BLOCKS
#define OSMR_COVER_MIN_BLOCKS 6  // 4: 2.03 ms, 6: 1.77, 8: 1.83 (C2 batch)
#endif
__global__ void __launch_bounds__(kCoverThreads, OSMR_COVER_MIN_BLOCKS) line_cover_kernel(Scene s) {
    __shared__ CoverSmem smem[kCoverWarps];
    CoverSmem& sm = smem[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    const int D = s.D;
    if (s.counters[CNT_OVERFLOW]) return;  // the draw is redone with larger buffers (see fill_rows_kernel)
    const unsigned n_work = s.counters[CNT_N_LINE_WORK];
    unsigned wi = 0, wi_end = 0;
    for (;; ++wi) {
        if (wi >= wi_end) {
            if (lane == 0) wi = atomicAdd(&s.counters[CNT_LINE_CURSOR], kCoverBatch);
            wi = __shfl_sync(0xffffffffu, wi, 0);
            wi_end = wi + kCoverBatch;
        }
        if (wi >= n_work) break;
        const uint2 item = s.line_work[wi];  // (op, batch of 32 segment records)
        const unsigned gi = item.x;
        const VisOp op = s.vis[gi];
        if (32u * item.y >= op.geom_cnt) continue;  // the batches were counted from an upper bound of the record count
        const unsigned pass = op.pass;
        osmr_styled_area ar;
        ar.entity = op.entity;
        ar.style = op.style;
        const osmr_style& st = s.styles[ar.style];
        LineParams lp;
        line_params(s, st, (int)pass, lp);
        const double hw = lp.width / 2.0;
        const int reach = line_reach(hw);
        const unsigned S = (unsigned)reach;
        const SegRec* segs = reinterpret_cast<const SegRec*>(s.geom + op.geom_off);
        const unsigned n_seg = op.geom_cnt;
        unsigned steps_stored = 0;
        FragSink fs;
        fs.pair = s.pair + op.mask_off;
        fs.entries = s.entries;
        fs.frag_cnt = s.frag_cnt;
        fs.frag_alpha = s.frag_alpha;
        fs.frag_pix = s.frag_pix;
        fs.counters = s.counters;
        op_block_rect(op.x0, op.y0, op.x1, op.y1, D, fs.bx0, fs.by0, fs.nbx, fs.nby);
        fs.last_block = -1;
        fs.last_entry = 0xffffffffu;
        fs.last_off = fs.last_cap = 0;
        __syncwarp();  // the previous op's walks are done with sm.calc
        {  // the op's opacity calculators (built by style_calc_kernel), 16 bytes per lane and step
            const uint4* src = s.calc_table + (size_t)(2u * ar.style + (pass - 1u)) * kCalcEntryUnits;
            uint4* dst0 = reinterpret_cast<uint4*>(&sm.calc[0]);
            uint4* dst1 = reinterpret_cast<uint4*>(&sm.calc[1]);
            const unsigned n_main = 4u + 4u * (unsigned)reinterpret_cast<const OpacityCalc*>(src)->n_segs;
            for (unsigned u = lane; u < n_main; u += 32) dst0[u] = src[u];
            if (lane < kCapCalcUnits) dst1[lane] = src[kMainCalcUnits + lane];
        }
        {
            const unsigned sb = 32u * item.y;
            const unsigned si = sb + lane;
            unsigned items = 0;
            __syncwarp();  // the previous batch's items are done with sm.recs / sm.pre
            if (si < n_seg) {
                const uint4* src = reinterpret_cast<const uint4*>(&segs[si]);
                uint4* dst = reinterpret_cast<uint4*>(&sm.recs[lane]);
                for (int u = 0; u < 4; ++u) dst[u] = src[u];
                items = 2u * sm.recs[lane].n_k;
            }
            unsigned incl = items;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            sm.pre[lane] = incl - items;
            __syncwarp();
            for (unsigned item = lane; item < total; item += 32) {
                int lo = 0, hi = 32;  // largest slot with pre <= item
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (sm.pre[mid] <= item)
                        lo = mid;
                    else
                        hi = mid;
                }
                const SegRec& h = sm.recs[lo];
                const unsigned local = item - sm.pre[lo];
                const int k = h.k0 + (int)(local >> 1);
                const int mul = (local & 1u) ? -1 : 1;
                const OpacityCalc& calc = sm.calc[h.flags & 1u];
                WalkItem w;
                walk_item_setup(h.x1, h.y1, h.x2, h.y2, h.flags, h.magic, k, w);
                SegConst sc;
                sc.x1 = h.x1;
                sc.y1 = h.y1;
                sc.numer_const = (long long)h.x2 * (long long)h.y1 - (long long)h.y2 * (long long)h.x1;
                sc.sdx = (long long)h.x2 - (long long)h.x1;
                sc.sdy = (long long)h.y2 - (long long)h.y1;
                sc.denom = h.denom;
                sc.traveled = h.traveled;
                sc.small = (h.flags & 4u) != 0;
                // the walk of step k, then the extra one of a double correction (line.rs:150-155); a walk whose start is
                // more than `reach` outside the tile on its own axis cannot put a pixel into it
                int mn = w.mn, p_error = w.p_error;
                unsigned len0 = 0, len1 = 0;
                if (mn >= -reach && mn <= D - 1 + reach)
                    len0 = cover_walk(fs, S, w, sc, calc, mn, p_error, mul, lp.opacity, D, &s.counters[CNT_WALK_TRUNC]);
                if (w.extra) {
                    p_error = wadd(wsub(p_error, 2 * w.mx_d), 2 * w.mn_d);
                    mn += w.mn_inc;
                    if (mn >= -reach && mn <= D - 1 + reach)
                        len1 = cover_walk(fs, S, w, sc, calc, mn, p_error, mul, lp.opacity, D, &s.counters[CNT_WALK_TRUNC]);
                }
                steps_stored += len0 + len1;
                OSMR_COUNT("cover.walks", (len0 != 0) + (len1 != 0));
            }
        }
        for (int o = 16; o > 0; o >>= 1) steps_stored += __shfl_xor_sync(0xffffffffu, steps_stored, o);
        if (lane == 0 && steps_stored)
            atomicAdd(reinterpret_cast<unsigned long long*>(&s.counters[CNT_WALK_STEPS]), (unsigned long long)steps_stored);
    }
}

// ------------------------------------------------------------------------------------------------------
// bin_ops_kernel: per 16x16 block of every tile the ORDERED list of the generations that can touch it, and per
// (block, line op) pair the storage of its fragments.
//
// Until round 1's v11 every raster warp scanned the bboxes of all visible ops of its tile, then the segments of every hit
// line op, then replayed the integer stepping of every cached walk that could reach its block (8 of 32 lanes busy).  Now
// the scan happens once, here, a thread per block; line_cover_kernel appends every covered pixel to the list of its
// (block, op) pair; and raster_kernel only streams its block's lists.
//
// A thread owns one block and runs twice over the tile's three op lists (Fill, Casing, Stroke, in order): the first
// sweep counts its entries and their fragment capacity, a CTA scan + one atomic per CTA turn the counts into offsets,
// the second sweep writes the entries, zeroes their fragment counters and fills the line ops' pair tables.
// Fragment capacity of a (block, segment): main steps within reach of the block x 4 walks (two sides, each possibly with the
// extra perpendicular of a double correction) x the longest stretch of a walk inside the block -- a walk moves one pixel per
// step along one axis, so at most 16 of its steps lie in a 16-pixel block, and it has at most `reach` in-line steps.
// ------------------------------------------------------------------------------------------------------
constexpr int kBinThreads = 256;

// main steps k (line.rs:133-158) of a segment whose major coordinate lies within `reach` of the block at (bx0, by0), cut to the
// steps [k0, k0 + n_k) that line_cover_kernel walks; returns the number of steps (0: the segment cannot reach the block)
__device__ __forceinline__ unsigned seg_block_steps(const int4 sr, int reach, int bx0, int by0, int k0, unsigned n_k) {
    const int mnx = min(sr.x, sr.z), mxx = max(sr.x, sr.z), mny = min(sr.y, sr.w), mxy = max(sr.y, sr.w);
    if (!((long long)mnx - reach <= bx0 + kBW - 1 && (long long)mxx + reach >= bx0 && (long long)mny - reach <= by0 + kBH - 1 &&
          (long long)mxy + reach >= by0))
        return 0u;
    const int dx = abs(wsub(sr.z, sr.x)), dy = abs(wsub(sr.w, sr.y));
    const bool swap = dx > dy;
    const int mx0 = swap ? sr.x : sr.y;
    const int mxd = swap ? dx : dy;
    const int mx_inc = swap ? (sr.x <= sr.z ? 1 : -1) : (sr.y <= sr.w ? 1 : -1);
    const long long lo = (long long)(swap ? bx0 : by0) - reach;
    const long long hi = (long long)(swap ? bx0 + kBW : by0 + kBH) - 1 + reach;
    long long ka, kb;
    if (mx_inc > 0) {
        ka = lo - mx0;
        kb = hi - mx0;
    } else {
        ka = (long long)mx0 - hi;
        kb = (long long)mx0 - lo;
    }
    if (ka < 0) ka = 0;
    if (kb > mxd) kb = mxd;
    const long long c0 = k0, c1 = (long long)k0 + (long long)n_k - 1;
    if (ka < c0) ka = c0;
    if (kb > c1) kb = c1;
    return kb >= ka ? (unsigned)(kb - ka + 1) : 0u;
}

// kMode 0: a CTA per block area runs both sweeps (bin_ops_kernel).  Low zooms put millions of visible ops into a handful of
// tiles (C4: z10-z12); the host then cuts the chunk sequence of a block area (the 256-op chunks of the Fill, Casing and Stroke
// lists, in that order) into `bin_slices` slices with a CTA each: kMode 1 = the counting sweep of a slice (bin_count_kernel),
// bin_scan_kernel = per block the slices' counts -> first entry / first fragment slot plus the block area's one bump allocation,
// kMode 2 = the writing sweep (bin_write_kernel).  A block's entries stay contiguous and in generation order: slice after slice.
template <int kMode>
__device__ __forceinline__ void bin_ops_body(const Scene& s) {
    __shared__ short4 s_bb[kBinThreads];
    __shared__ uint4 s_rop[kBinThreads];  // first half of the RasterOp: a, b, (y0, y1), (kind, rgb)
    __shared__ unsigned s_reach[kBinThreads];
    __shared__ unsigned s_pair[kBinThreads];
    __shared__ unsigned w_ent[kBinThreads / 32];
    __shared__ unsigned long long w_cap[kBinThreads / 32];
    __shared__ unsigned base_ent;
    __shared__ unsigned long long base_cap;
    __shared__ int s_go;
    const int D = s.D;
    const int bpr = D / kBW, bpc = D / kBH, nblk = bpr * bpc;
    // A CTA covers an area of 16 x 16 blocks (D is a multiple of 256, so the areas tile the tile); a warp covers a compact region
    // of 8 x 4 blocks inside it, a lane one block: an op is tested against the warp's region first (one op per lane), and only
    // the survivors -- the few ops near the region -- are tested block by block.
    static_assert(kBinThreads == 256, "area layout");
    const unsigned areas_x = (unsigned)bpr / 16u;
    const unsigned groups = areas_x * ((unsigned)bpc / 16u);
    const unsigned S = kMode == 0 ? 1u : s.bin_slices;
    const unsigned cta = blockIdx.x / S, slice = blockIdx.x % S;
    const unsigned tile = cta / groups, grp = cta % groups;
    const unsigned warp_in = threadIdx.x >> 5, lane_in = threadIdx.x & 31u;
    const int rbx = (int)((grp % areas_x) * 16u + (warp_in & 1u) * 8u), rby = (int)((grp / areas_x) * 16u + (warp_in >> 1) * 4u);
    const int bxi = rbx + (int)(lane_in & 7u), byi = rby + (int)(lane_in >> 3);
    const unsigned b = (unsigned)(byi * bpr + bxi);  // my block, row-major
    const bool live = true;
    const int bx0 = bxi * kBW, by0 = byi * kBH;
    const int wx0 = rbx * kBW, wy0 = rby * kBH, wx1 = wx0 + 8 * kBW - 1, wy1 = wy0 + 4 * kBH - 1;  // the warp's region in pixels
    if (threadIdx.x == 0) s_go = s.counters[CNT_OVERFLOW] == 0u;  // (one read per CTA: other CTAs may raise the flag meanwhile)
    __syncthreads();
    if (!s_go) return;  // the draw is redone with larger buffers
    const unsigned base = s.area_begin[tile];
    const unsigned n_areas_tile = s.area_begin[tile + 1] - base;
    unsigned out_ent = 0;
    unsigned long long out_cap = 0;
    // the slice's chunks of the concatenated chunk sequence (kMode 0: all of them)
    unsigned c_lo = 0, c_hi = 0xffffffffu;
    if (kMode != 0) {
        unsigned total_chunks = 0;
        for (unsigned p = 0; p < 3u; ++p) total_chunks += (s.vis_count[3u * tile + p] + kBinThreads - 1u) / kBinThreads;
        const unsigned per = (total_chunks + S - 1u) / S;
        c_lo = min(total_chunks, slice * per);
        c_hi = min(total_chunks, c_lo + per);
    }
    if (kMode == 2) {
        out_ent = s.bin_cnt_ent[(size_t)blockIdx.x * kBinThreads + threadIdx.x];
        out_cap = s.bin_cnt_cap[(size_t)blockIdx.x * kBinThreads + threadIdx.x];
    }
    for (int sweep = (kMode == 2 ? 1 : 0); sweep < (kMode == 1 ? 1 : 2); ++sweep) {
        unsigned n_ent = 0;
        unsigned long long cap_sum = 0;
        unsigned chunk_base = 0;
        for (unsigned the_pass = 0; the_pass < 3u; ++the_pass) {
            const unsigned long long list0 = 3ull * base + (unsigned long long)the_pass * n_areas_tile;
            const unsigned n_vis = s.vis_count[3u * tile + the_pass];
            const unsigned n_chunks = (n_vis + kBinThreads - 1u) / kBinThreads;
            const unsigned ci_lo = c_lo > chunk_base ? min(n_chunks, c_lo - chunk_base) : 0u;
            const unsigned ci_hi = c_hi > chunk_base ? min(n_chunks, c_hi - chunk_base) : 0u;
            chunk_base += n_chunks;
            for (unsigned ci = ci_lo; ci < ci_hi; ++ci) {
                const unsigned chunk = ci * kBinThreads;
                __syncthreads();
                const unsigned vi = chunk + threadIdx.x;
                if (vi < n_vis) {
                    s_bb[threadIdx.x] = s.vis_bbox[list0 + vi];
                    const uint4 r0 = *reinterpret_cast<const uint4*>(&s.rop[list0 + vi]);
                    s_rop[threadIdx.x] = r0;
                    const RasterOp& full = s.rop[list0 + vi];
                    s_reach[threadIdx.x] = full.reach;
                    s_pair[threadIdx.x] = s.vis[list0 + vi].mask_off;
                }
                __syncthreads();
                const unsigned n_here = min((unsigned)kBinThreads, n_vis - chunk);
                for (unsigned sub = 0; sub < n_here; sub += 32) {
                unsigned surv;
                {
                    const unsigned t = sub + lane_in;
                    const short4 ot = s_bb[min(t, n_here - 1u)];
                    surv = __ballot_sync(0xffffffffu, t < n_here && ot.x <= wx1 && ot.z >= wx0 && ot.y <= wy1 && ot.w >= wy0);
                }
                while (surv) {
                    const unsigned q = sub + (unsigned)(__ffs(surv) - 1);
                    surv &= surv - 1;
                    const short4 o = s_bb[q];
                    const bool box_hit = live && o.x <= bx0 + kBW - 1 && o.z >= bx0 && o.y <= by0 + kBH - 1 && o.w >= by0;
                    if (!box_hit) continue;
                    const uint4 r0 = s_rop[q];
                    const unsigned kind = r0.w & 0xffu;
                    unsigned cap = 0;
                    bool hit = true;
                    if (kind == OP_LINE) {
                        const int reach = (int)s_reach[q];
                        const SegRec* segs = reinterpret_cast<const SegRec*>(s.geom + r0.x);
                        const unsigned n_seg = r0.y;
                        const unsigned per_step = 4u * (unsigned)min(reach, max(kBW, kBH));
                        unsigned long long c = 0;
                        for (unsigned si = 0; si < n_seg; ++si) {
                            const int4 sr = *reinterpret_cast<const int4*>(&segs[si]);
                            const uint4 tail = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(&segs[si]) + 40);  // flags, k0, n_k
                            c += (unsigned long long)seg_block_steps(sr, reach, bx0, by0, (int)tail.y, tail.z) * per_step;
                        }
                        hit = c != 0;
                        cap = (unsigned)min(c, 0xffffffffull);
                        // pair table cell of this block (written whether or not a segment reaches it)
                        if (sweep == 1) {
                            int rbx0, rby0, nbx, nby;
                            op_block_rect(o.x, o.y, o.z, o.w, D, rbx0, rby0, nbx, nby);
                            const int cx = bx0 / kBW - rbx0, cy = by0 / kBH - rby0;
                            s.pair[s_pair[q] + (unsigned)(cy * nbx + cx)] = hit ? (out_ent + n_ent) : 0xffffffffu;
                        }
                    }
                    if (!hit) continue;
                    if (sweep == 1) {
                        const unsigned e = out_ent + n_ent;
                        BinEntry be;
                        be.op = (unsigned)(list0 + chunk + q);
                        be.frag_off = (unsigned)(out_cap + cap_sum);
                        be.frag_cap = cap;
                        be.pad = 0;
                        s.entries[e] = be;
                        s.frag_cnt[e] = 0u;
                    }
                    ++n_ent;
                    cap_sum += cap;
                }
                }
            }
        }
        if (kMode == 1) {  // the slice's counts; bin_scan_kernel turns them into offsets
            s.bin_cnt_ent[(size_t)blockIdx.x * kBinThreads + threadIdx.x] = n_ent;
            s.bin_cnt_cap[(size_t)blockIdx.x * kBinThreads + threadIdx.x] = cap_sum;
            return;
        }
        if (sweep == 0) {
            // exclusive scan over the CTA's threads, one global bump allocation per CTA
            unsigned incl_e = n_ent;
            unsigned long long incl_c = cap_sum;
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned ye = __shfl_up_sync(0xffffffffu, incl_e, off);
                const unsigned long long yc = __shfl_up_sync(0xffffffffu, incl_c, off);
                if ((int)lane_id() >= off) {
                    incl_e += ye;
                    incl_c += yc;
                }
            }
            const unsigned w = threadIdx.x >> 5;
            if (lane_id() == 31) {
                w_ent[w] = incl_e;
                w_cap[w] = incl_c;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned te = 0;
                unsigned long long tc = 0;
                for (unsigned k = 0; k < kBinThreads / 32; ++k) {
                    te += w_ent[k];
                    tc += w_cap[k];
                }
                base_ent = te ? atomicAdd(&s.counters[CNT_BIN_ENTRIES], te) : 0u;
                base_cap = tc ? atomicAdd(reinterpret_cast<unsigned long long*>(&s.counters[CNT_WALK_ALPHA]), tc) : 0ull;
                const bool ent_fits = (unsigned long long)base_ent + te <= s.entries_cap;
                const bool cap_fits = base_cap + tc <= s.frag_cap && base_cap + tc <= 0xfffffff0ull;
                if (!ent_fits) atomicOr(&s.counters[CNT_OVERFLOW], 64u);
                if (!cap_fits) atomicOr(&s.counters[CNT_OVERFLOW], 4u);
                s_go = ent_fits && cap_fits;  // a CTA whose own allocation is out of range skips its second sweep
            }
            __syncthreads();
            unsigned before_e = base_ent;
            unsigned long long before_c = base_cap;
            for (unsigned k = 0; k < w; ++k) {
                before_e += w_ent[k];
                before_c += w_cap[k];
            }
            out_ent = before_e + incl_e - n_ent;
            out_cap = before_c + incl_c - cap_sum;
            if (!s_go) return;
            if (live) s.blk_range[(size_t)tile * nblk + b] = make_uint2(out_ent, n_ent);
        }
    }
}

__global__ void __launch_bounds__(kBinThreads) bin_ops_kernel(Scene s) { bin_ops_body<0>(s); }
__global__ void __launch_bounds__(kBinThreads) bin_count_kernel(Scene s) { bin_ops_body<1>(s); }
__global__ void __launch_bounds__(kBinThreads) bin_write_kernel(Scene s) { bin_ops_body<2>(s); }

// one CTA per block area, a thread per block (bin_ops_body's thread layout): the slices' counts of the block -> the slices' first
// entry and first fragment slot; the CTA scan and the one bump allocation per CTA are bin_ops_body's own
__global__ void __launch_bounds__(kBinThreads) bin_scan_kernel(Scene s) {
    __shared__ unsigned w_ent[kBinThreads / 32];
    __shared__ unsigned long long w_cap[kBinThreads / 32];
    __shared__ unsigned base_ent;
    __shared__ unsigned long long base_cap;
    __shared__ int s_go;
    const int D = s.D;
    const int bpr = D / kBW, bpc = D / kBH, nblk = bpr * bpc;
    const unsigned areas_x = (unsigned)bpr / 16u;
    const unsigned groups = areas_x * ((unsigned)bpc / 16u);
    const unsigned S = s.bin_slices, cta = blockIdx.x;
    const unsigned tile = cta / groups, grp = cta % groups;
    const unsigned warp_in = threadIdx.x >> 5, lane_in = threadIdx.x & 31u;
    const int rbx = (int)((grp % areas_x) * 16u + (warp_in & 1u) * 8u), rby = (int)((grp / areas_x) * 16u + (warp_in >> 1) * 4u);
    const int bxi = rbx + (int)(lane_in & 7u), byi = rby + (int)(lane_in >> 3);
    const unsigned b = (unsigned)(byi * bpr + bxi);
    if (threadIdx.x == 0) s_go = s.counters[CNT_OVERFLOW] == 0u;
    __syncthreads();
    if (!s_go) return;
    unsigned n_ent = 0;
    unsigned long long cap_sum = 0;
    for (unsigned k = 0; k < S; ++k) {
        const size_t idx = ((size_t)cta * S + k) * kBinThreads + threadIdx.x;
        n_ent += s.bin_cnt_ent[idx];
        cap_sum += s.bin_cnt_cap[idx];
    }
    unsigned incl_e = n_ent;
    unsigned long long incl_c = cap_sum;
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned ye = __shfl_up_sync(0xffffffffu, incl_e, off);
        const unsigned long long yc = __shfl_up_sync(0xffffffffu, incl_c, off);
        if ((int)lane_id() >= off) {
            incl_e += ye;
            incl_c += yc;
        }
    }
    const unsigned w = threadIdx.x >> 5;
    if (lane_id() == 31) {
        w_ent[w] = incl_e;
        w_cap[w] = incl_c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long te = 0, tc = 0;
        for (unsigned k = 0; k < kBinThreads / 32; ++k) {
            te += w_ent[k];
            tc += w_cap[k];
        }
        const bool te_ok = te <= 0xfffffff0ull;
        base_ent = (te && te_ok) ? atomicAdd(&s.counters[CNT_BIN_ENTRIES], (unsigned)te) : 0u;
        base_cap = tc ? atomicAdd(reinterpret_cast<unsigned long long*>(&s.counters[CNT_WALK_ALPHA]), tc) : 0ull;
        const bool ent_fits = te_ok && (unsigned long long)base_ent + te <= s.entries_cap;
        const bool cap_fits = base_cap + tc <= s.frag_cap && base_cap + tc <= 0xfffffff0ull;
        if (!ent_fits) atomicOr(&s.counters[CNT_OVERFLOW], 64u);
        if (!cap_fits) atomicOr(&s.counters[CNT_OVERFLOW], 4u);
        s_go = ent_fits && cap_fits;
    }
    __syncthreads();
    if (!s_go) return;  // (bin_write_kernel and everything behind it stop at the overflow flag)
    unsigned e = base_ent;
    unsigned long long c = base_cap;
    for (unsigned k = 0; k < w; ++k) {
        e += w_ent[k];
        c += w_cap[k];
    }
    e += incl_e - n_ent;
    c += incl_c - cap_sum;
    s.blk_range[(size_t)tile * nblk + b] = make_uint2(e, n_ent);
    for (unsigned k = 0; k < S; ++k) {
        const size_t idx = ((size_t)cta * S + k) * kBinThreads + threadIdx.x;
        const unsigned ve = s.bin_cnt_ent[idx];
        const unsigned long long vc = s.bin_cnt_cap[idx];
        s.bin_cnt_ent[idx] = e;
        s.bin_cnt_cap[idx] = c;
        e += ve;
        c += vc;
    }
}


// ------------------------------------------------------------------------------------------------------
// raster_kernel (a3 blend, a4 / a5 combine, a6, a7): ONE WARP (= one 32-thread CTA) per 16x16-pixel block of a tile.
//
// Compositor invariants used (tile_pixels.rs:107-129,205-223): inside one generation the surviving source of a
// pixel is the contribution with the largest alpha; generations blend in order with premultiplied over; canvas alpha stays
// exactly 1.0, so only RGB is kept and export is trunc(255*c).
//
// The warp owns its block for the whole ordered entry list bin_ops_kernel made for it: f64 canvas + f64 alpha plane in shared
// memory.  A fill entry blends straight from the op's row masks; a line entry max-combines the fragments line_cover_kernel
// appended to this (block, op) pair into the alpha plane and blends the plane.  No CTA barrier anywhere (v1 shared a 64x32
// region between 8 warps and was barrier-bound, profiles/r01_raster_v1_ncu_summary.txt); the hardware block scheduler balances
// the blocks.  The next entry's fragment list is fetched into shared memory by the bulk-copy engine (cp.async.bulk + mbarrier)
// while the current entry is combined and blended.
// ------------------------------------------------------------------------------------------------------
constexpr int kRasterThreads = 32;
#ifndef OSMR_RASTER_MIN_BLOCKS
#define OSMR_RASTER_MIN_BLOCKS 24  // resident one-warp CTAs per SM the register allocation must allow
#endif
#ifndef OSMR_RASTER_TMA
#define OSMR_RASTER_TMA 1  // 1: fragment lists are staged in shared memory by cp.async.bulk, double buffered
#endif
constexpr unsigned kStageFrags = 192;  // fragments per staging buffer (alphas 1536 B + pixel bytes 192 B)

struct alignas(16) FragStage {
    double alpha[kStageFrags];
    unsigned char pix[kStageFrags];
};

struct RasterSmem {
    double canvas[3][kBP];
    unsigned long long plane[kBP];
#if OSMR_RASTER_TMA
    FragStage stage[2];
    unsigned long long bar[2];  // mbarriers of the two staging buffers
#endif
};

#if OSMR_RASTER_TMA && !defined(OSMR_EMULATED)
// 1-D bulk copies global -> shared with mbarrier completion (sm_90+ / sm_100a): the copy engine moves the bytes, the warp
// only waits on the barrier's phase.  Sizes and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

__global__ void __launch_bounds__(kRasterThreads, OSMR_RASTER_MIN_BLOCKS) raster_kernel(Scene s) {
    __shared__ RasterSmem sm;
    const int D = s.D;
    const int bpr = D / kBW, bpc = D / kBH;  // blocks per tile row / column
    const unsigned tile = blockIdx.x / (unsigned)(bpr * bpc);
    const unsigned blk = blockIdx.x % (unsigned)(bpr * bpc);
    // consecutive CTAs cover a 2x2 group of blocks before moving on, so neighbours share masks and op records in L1/L2
    const unsigned grp = blk / 4u, in_grp = blk % 4u;
    const unsigned grp_per_row = (unsigned)bpr / 2u;
    const int bxi = (int)((grp % grp_per_row) * 2u + (in_grp % 2u)), byi = (int)((grp / grp_per_row) * 2u + (in_grp / 2u));
    const int bx0 = bxi * kBW, by0 = byi * kBH;
    const unsigned lane = threadIdx.x;
    if (s.counters[CNT_OVERFLOW]) return;  // the draw is redone with larger buffers (see fill_rows_kernel)

    // TilePixels::reset (tile_pixels.rs:89-105): canvas colour premultiplied with opacity 1.0, or (0,0,0,1)
    {
        double c[3] = {0.0, 0.0, 0.0};
        if (s.flags & OSMR_DRAW_HAS_CANVAS_COLOR)
            for (int k = 0; k < 3; ++k) c[k] = 1.0 * unit_of_u8(s.canvas[k]);
        for (int i = (int)lane; i < kBP; i += 32) {
            sm.canvas[0][i] = c[0];
            sm.canvas[1][i] = c[1];
            sm.canvas[2][i] = c[2];
            sm.plane[i] = 0ull;
        }
    }
#if OSMR_RASTER_TMA && !defined(OSMR_EMULATED)
    if (lane == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        fence_barrier_init();
    }
    unsigned phases = 0u;  // bit b: parity the next wait on staging buffer b expects
#endif
    __syncwarp();

    const uint2 range = s.blk_range[(size_t)tile * (bpr * bpc) + (unsigned)(byi * bpr + bxi)];
    const BinEntry* ents = s.entries + range.x;
    const unsigned n_ent = range.y;
    for (unsigned chunk = 0; chunk < n_ent; chunk += 32) {
        // ---- 32 entries of my block: every lane fetches the records of ITS entry now (independent loads) ----
        const unsigned ei = chunk + lane;
        BinEntry my_ent;
        my_ent.op = my_ent.frag_off = my_ent.frag_cap = my_ent.pad = 0;
        uint4 my_rop0 = make_uint4(0, 0, 0, 0), my_rop1 = make_uint4(0, 0, 0, 0);
        unsigned my_cnt = 0;
        uint4 my_m0 = make_uint4(0, 0, 0, 0), my_m1 = make_uint4(0, 0, 0, 0);  // fills: the mask record of my block (16 rows x 16 bits)
        if (ei < n_ent) {
            my_ent = ents[ei];
            const uint4* src = reinterpret_cast<const uint4*>(&s.rop[my_ent.op]);
            my_rop0 = src[0];
            my_rop1 = src[1];
            if ((my_rop0.w & 0xffu) == OP_LINE) {
                my_cnt = min(s.frag_cnt[range.x + ei], my_ent.frag_cap);
            } else {
                const int oy0 = (int)(short)(my_rop0.z & 0xffffu);
                const unsigned rect = my_rop1.x & 0xffffu;  // RasterOp.reach of a fill: first block column | block columns << 8
                const int rbx0 = (int)(rect & 0xffu), nbx = (int)(rect >> 8), rby0 = max(oy0, 0) / kBH;
                const uint4* rec = reinterpret_cast<const uint4*>(s.mask + my_rop0.x) + (size_t)((byi - rby0) * nbx + (bxi - rbx0)) * 2;
                my_m0 = rec[0];
                my_m1 = rec[1];
            }
        }
        const unsigned n_here = min(32u, n_ent - chunk);
#if OSMR_RASTER_TMA && !defined(OSMR_EMULATED)
        // Software pipeline over the fragment lists of this chunk: they are cut into pieces of <= kPiece fragments; while a
        // piece is combined, the next piece (of this entry or of the next line entry) is already on its way into the other
        // staging buffer -- and stays in flight across the fill entries in between.
        constexpr unsigned kPiece = kStageFrags - 16u;  // room for the alignment skew of both arrays
        const unsigned line_mask = __ballot_sync(0xffffffffu, my_cnt > 0u);
        auto piece_issue = [&](unsigned q, unsigned done, int buf) {
            const unsigned nq = __shfl_sync(0xffffffffu, my_cnt, q);
            const unsigned off = __shfl_sync(0xffffffffu, my_ent.frag_off, q) + done;
            const unsigned n = min(kPiece, nq - done);
            if (lane == 0) {
                // both arrays are copied from the 16-byte aligned address at or below their first element
                const unsigned a_skew = off & 1u, p_skew = off & 15u;
                const unsigned a_bytes = ((n + a_skew) * 8u + 15u) & ~15u, p_bytes = (n + p_skew + 15u) & ~15u;
                mbar_expect_tx(&sm.bar[buf], a_bytes + p_bytes);
                bulk_g2s(sm.stage[buf].alpha, s.frag_alpha + (off - a_skew), a_bytes, &sm.bar[buf]);
                bulk_g2s(sm.stage[buf].pix, s.frag_pix + (off - p_skew), p_bytes, &sm.bar[buf]);
            }
        };
        int cur_buf = 0;
        if (line_mask) piece_issue((unsigned)(__ffs(line_mask) - 1), 0u, cur_buf);
#endif
        for (unsigned q = 0; q < n_here; ++q) {
            RasterOp op;
            {
                uint4 r0, r1;
                r0.x = __shfl_sync(0xffffffffu, my_rop0.x, q);
                r0.y = __shfl_sync(0xffffffffu, my_rop0.y, q);
                r0.z = __shfl_sync(0xffffffffu, my_rop0.z, q);
                r0.w = __shfl_sync(0xffffffffu, my_rop0.w, q);
                r1.x = __shfl_sync(0xffffffffu, my_rop1.x, q);
                r1.y = __shfl_sync(0xffffffffu, my_rop1.y, q);
                r1.z = __shfl_sync(0xffffffffu, my_rop1.z, q);
                r1.w = __shfl_sync(0xffffffffu, my_rop1.w, q);
                uint4* dst = reinterpret_cast<uint4*>(&op);
                dst[0] = r0;
                dst[1] = r1;
            }
            if (op.kind != OP_LINE) {
                // ---------------- fill: blend straight from the row masks ----------------
                const int ya = max((int)op.y0, 0);
                const int wpr = D / 32;
                double src[4];
                const DevIcon* icon = nullptr;
                if (op.kind == OP_FILL_COLOR) {
                    const double opacity = op.opacity;  // fill-opacity or 1.0 (drawer.rs:172-178)
                    for (int k = 0; k < 3; ++k) src[k] = opacity * unit_of_u8(op.rgb[k]);
                    src[3] = opacity;
                } else {
                    icon = &s.icons[op.b];
                }
                const int col = (int)(lane % (unsigned)kBW);
                // the block's mask record, handed round from the lane that fetched it: word j = rows 2j, 2j+1, and lane l is
                // the pixel (row 2j + l / 16, column l % 16): bit l of word j
                unsigned mw[8];
                mw[0] = __shfl_sync(0xffffffffu, my_m0.x, q);
                mw[1] = __shfl_sync(0xffffffffu, my_m0.y, q);
                mw[2] = __shfl_sync(0xffffffffu, my_m0.z, q);
                mw[3] = __shfl_sync(0xffffffffu, my_m0.w, q);
                mw[4] = __shfl_sync(0xffffffffu, my_m1.x, q);
                mw[5] = __shfl_sync(0xffffffffu, my_m1.y, q);
                mw[6] = __shfl_sync(0xffffffffu, my_m1.z, q);
                mw[7] = __shfl_sync(0xffffffffu, my_m1.w, q);
#pragma unroll
                for (int j = 0; j < kBP / 32; ++j) {
                    const int r = (j * 32 + (int)lane) / kBW;
                    const int y = by0 + r;
                    if (y < (int)op.y0 || y > (int)op.y1) continue;  // (rows of the record outside the op's rows were never written)
                    if (!((mw[j] >> lane) & 1u)) continue;
                    const int idx = r * kBW + col;
                    double c0, c1, c2, a;
                    if (icon) {  // Filler::Image (fill.rs:36-40): texel (x mod w, y mod h), already premultiplied
                        unsigned ix = (unsigned)(bx0 + col) % icon->w, iy = (unsigned)y % icon->h;
                        double4 t = s.icon_px[icon->off + iy * icon->w + ix];
                        c0 = t.x;
                        c1 = t.y;
                        c2 = t.z;
                        a = t.w;
                    } else {
                        c0 = src[0];
                        c1 = src[1];
                        c2 = src[2];
                        a = src[3];
                    }
                    double inv = 1.0 - a;  // blend_pixel (tile_pixels.rs:211-216)
                    sm.canvas[0][idx] = c0 + inv * sm.canvas[0][idx];
                    sm.canvas[1][idx] = c1 + inv * sm.canvas[1][idx];
                    sm.canvas[2][idx] = c2 + inv * sm.canvas[2][idx];
                }
                __syncwarp();
                continue;
            }
            // ---------------- line: max-combine this pair's fragments into the alpha plane, then blend ----------------
            const unsigned n_frag = __shfl_sync(0xffffffffu, my_cnt, q);
            const unsigned f_off = __shfl_sync(0xffffffffu, my_ent.frag_off, q);
            if (!n_frag) continue;
#if OSMR_RASTER_TMA && !defined(OSMR_EMULATED)
            for (unsigned done = 0; done < n_frag; done += kPiece) {
                const unsigned n = min(kPiece, n_frag - done);
                const unsigned off = f_off + done;
                // the piece (q, done) is in flight in cur_buf; start the one after it
                if (done + kPiece < n_frag) {
                    piece_issue(q, done + kPiece, cur_buf ^ 1);
                } else {
                    const unsigned rest = q >= 31u ? 0u : (line_mask & ~((2u << q) - 1u));
                    if (rest) piece_issue((unsigned)(__ffs(rest) - 1), 0u, cur_buf ^ 1);
                }
                mbar_wait(&sm.bar[cur_buf], (phases >> cur_buf) & 1u);
                phases ^= 1u << cur_buf;
                const unsigned a_skew = off & 1u, p_skew = off & 15u;
                for (unsigned i = lane; i < n; i += 32) {
                    const double a = sm.stage[cur_buf].alpha[i + a_skew];
                    const unsigned pix = sm.stage[cur_buf].pix[i + p_skew];
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(a);
                    unsigned long long* cell = &sm.plane[pix];
                    if (bits > *cell) atomicMax(cell, bits);
                }
                __syncwarp();
                cur_buf ^= 1;
            }
#else
            for (unsigned i = lane; i < n_frag; i += 32) {
                const double a = s.frag_alpha[(size_t)f_off + i];
                const unsigned pix = s.frag_pix[(size_t)f_off + i];
                const unsigned long long bits = (unsigned long long)__double_as_longlong(a);
                unsigned long long* cell = &sm.plane[pix];
                if (bits > *cell) atomicMax(cell, bits);
            }
#endif
            __syncwarp();
            {
                // blend: pending pixel = from_color(color, alpha_max) (tile_pixels.rs:13-22), then over
                double cn[3];
                for (int k = 0; k < 3; ++k) cn[k] = unit_of_u8(op.rgb[k]);
#pragma unroll 2
                for (int j = 0; j < kBP / 32; ++j) {
                    const int idx = j * 32 + (int)lane;
                    unsigned long long bits = sm.plane[idx];
                    if (bits) {
                        double a = __longlong_as_double((long long)bits);
                        double inv = 1.0 - a;
                        sm.canvas[0][idx] = a * cn[0] + inv * sm.canvas[0][idx];
                        sm.canvas[1][idx] = a * cn[1] + inv * sm.canvas[1][idx];
                        sm.canvas[2][idx] = a * cn[2] + inv * sm.canvas[2][idx];
                        sm.plane[idx] = 0ull;
                    }
                }
            }
            __syncwarp();
        }
    }

    // ---- export (tile_pixels.rs:164-181): alpha == 1.0, so postdivide is the identity ----
    __syncwarp();
    // label pass result (drawer.rs:124 blend_unfinished_pixels(true)): one more premultiplied over per labelled pixel
    if (s.label_plane) {
        // (a tile has a few hundred labelled pixels: the warp reads the block's 256 mask bits, not 4 KB of LabelPix records)
        const LabelPix* plane = s.label_plane + (size_t)tile * D * D;
        const unsigned* pmask = s.label_mask + (size_t)tile * (size_t)(D * D / 32);
        for (int i = (int)lane; i < kBP; i += 32) {
            const size_t pix = (size_t)(by0 + i / kBW) * D + bx0 + i % kBW;
            if (!((pmask[pix >> 5] >> (pix & 31)) & 1u)) continue;
            const LabelPix lp = plane[pix];
            if (!lp.src) continue;
            double c0, c1, c2, a;
            if (lp.src & 0x80000000u) {
                const double4 tx = s.label_icon_px[lp.src & 0x3fffffffu];
                c0 = tx.x;
                c1 = tx.y;
                c2 = tx.z;
                a = tx.w;
            } else {  // RgbaColor::from_color(text_color, total)
                a = lp.alpha;
                c0 = a * unit_of_u8((unsigned char)(lp.src & 0xffu));
                c1 = a * unit_of_u8((unsigned char)((lp.src >> 8) & 0xffu));
                c2 = a * unit_of_u8((unsigned char)((lp.src >> 16) & 0xffu));
            }
            const double inv = 1.0 - a;
            sm.canvas[0][i] = c0 + inv * sm.canvas[0][i];
            sm.canvas[1][i] = c1 + inv * sm.canvas[1][i];
            sm.canvas[2][i] = c2 + inv * sm.canvas[2][i];
        }
        __syncwarp();
    }
    auto texel = [&](int ch, int idx) -> unsigned { return f64_as_u8(255.0 * (sm.canvas[ch][idx] / 1.0)); };
    if (s.flags & OSMR_DRAW_OUT_RGBA) {
        uchar4* out = reinterpret_cast<uchar4*>(s.out) + (size_t)tile * D * D;
        for (int i = (int)lane; i < kBP; i += 32) {
            int r = i / kBW, x = i % kBW;
            uchar4 v;
            v.x = (unsigned char)texel(0, i);
            v.y = (unsigned char)texel(1, i);
            v.z = (unsigned char)texel(2, i);
            v.w = 255;
            out[(size_t)(by0 + r) * D + bx0 + x] = v;
        }
    } else {
        // kBH rows of 3*kBW bytes = 3*kBW/16 aligned 128-bit words each (kBW = 16: 48 bytes = 3 words per row)
        constexpr int qpr = 3 * kBW / 16;
        unsigned char* tile_out = s.out + (size_t)tile * D * D * 3;
        for (int i = (int)lane; i < kBH * qpr; i += 32) {
            const int r = i / qpr, j = i % qpr;
            unsigned w4[4];
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                unsigned word = 0;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const int byte = 16 * j + 4 * wd + bb;
                    word |= texel(byte % 3, r * kBW + byte / 3) << (8 * bb);
                }
                w4[wd] = word;
            }
            uint4* dst = reinterpret_cast<uint4*>(tile_out + ((size_t)(by0 + r) * D + bx0) * 3);
            dst[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
    }
}


// ------------------------------------------------------------------------------------------------------
// label pass, device half (a9 + f1 + f2): label_cover_kernel + label_commit_kernel.
//   icon blit (labeler.rs:91-106) / glyph coverage (rasterizer.rs:27-84,109-148) -> collision test against the pixels of
//   earlier successful labels over the 3x3 label canvas (tile_pixels.rs:131-148) -> commit.
// Glyph coverage is the reference's exact-area accumulation: one thread per pixel row adds the contributions of the
// label's segments IN SEGMENT ORDER into dense per-row `a` / `s` arrays (the reference's BTreeMaps), tracking the
// smallest / largest touched key, then sweeps the row left to right.  Only IEEE basic operations are used; the
// segments themselves come from the host (osmr_labels_host.hpp), where the reference's libm calls live.
// ------------------------------------------------------------------------------------------------------
struct DevLabel {  // == osmr_host::LabelRec
    int icon, ix, iy;
    unsigned seg_begin, seg_count;  // host layout: the label's segments are segs[seg_begin .. + seg_count)
    int bx0, by0, bx1, by1;
    unsigned rgb;
    int ry0, rows, width;
    unsigned row_first;  // first entry of this label in kmin / kmax
    unsigned n_ranges;   // (reserved)
    unsigned range_off;
    unsigned long long cell_off;
};
struct alignas(16) DevSeg {  // (16-byte aligned: moved with two 128-bit accesses)
    double x0, y0, x1, y1;
};
struct LabelScene {
    const DevLabel* labels;
    const unsigned* label_begin;  // per tile
    const DevSeg* segs;
    const unsigned* cover_list;  // slots of the labels that have text coverage to compute
    unsigned n_cover;
    const DevIcon* icons;
    unsigned* occ;   // per tile (3D)^2 bits
    double* acc_a;   // coverage: per label rows x width cells (`a` map, then the swept totals)
    double* acc_s;   // the `s` map (only for labels whose cells do not fit in shared memory)
    int* kmin;       // per (label, row): smallest / largest touched key
    int* kmax;
    LabelPix* plane;  // per tile D*D
    unsigned* pmask;  // per tile D*D bits: which entries of `plane` are pending pixels of this draw (the rest is stale)
    int D;
    // device-side layout (osmr_labels_dev.cuh): the counts live in device memory and the labels of a tile are label_cnt[tile]
    // records starting at label_begin[tile]; nullptr: host-side layout (n_cover, label_begin is a prefix sum)
    const unsigned* n_cover_dev;
    const unsigned* label_cnt;
    const unsigned* skip_flags;  // device layout: [0] overflow, [1] fallback -- nonzero: this attempt is abandoned
    unsigned* cover_cursor;      // work distribution of label_cover_kernel (zeroed before the launch)
    unsigned* err_flag;          // raised when a pair falls outside its proven window (a logic error: the call fails loudly)
};

// ------------------------------------------------------------------------------------------------------
// label_cover_kernel (f2): exact-area glyph coverage of a label, Rasterizer::draw_line for every segment IN SEGMENT ORDER
// (rasterizer.rs:27-84: the BTreeMaps `a` and `s` are dense per-row arrays here), then the left-to-right sweep of
// save_to_figure (rasterizer.rs:109-148).
//
// The reference's flatness rule (1.0001) turns every curve into ~64 sub-pixel segments, so a label is thousands of segments that
// each add into one or two cells.  f64 addition does not commute with reordering: a cell must receive its contributions in
// segment order -- but only the ADDITIONS into one cell are ordered; the area arithmetic in front of them is not, and different
// cells do not interact.  One warp owns a label and takes its segments in batches:
//   A  lanes = (segment, row) crossings: the area arithmetic, fully parallel; every `+=` of draw_line becomes a pair
//      (cell, value) in shared memory, in the reference's order
//   B  stable counting sort of the pairs by cell (the batch's cells are a small window of the label: a glyph or two)
//   C  lanes = cells: each lane adds the pairs of its cell, in order, to the cell's running value
// Measured history of this kernel on the C2 batch: a warp per 32 rows scanning all segments 21 ms; ordered commit lane by lane
// 12.7 ms; a lane per row 13.1 ms (3.7 of 32 lanes busy: a street name is a dozen rows); a lane per (row, column bin) 17.5 ms
// (the crossings of a glyph pile up in a few buckets: 5 lanes busy).
// ------------------------------------------------------------------------------------------------------
#ifndef OSMR_COV_PAIRS
#define OSMR_COV_PAIRS 512
#endif
#ifndef OSMR_COV_CTAS
#define OSMR_COV_CTAS 24
#endif
constexpr int kCovPairs = OSMR_COV_PAIRS;     // pairs per batch
constexpr int kCovKeys = OSMR_COV_PAIRS / 2;  // cells (x 2 arrays) of the batch's window
constexpr unsigned kCovCtasPerSm = OSMR_COV_CTAS;
constexpr int kCovDirectUnits = 96;
constexpr int kCovSegs = 256;  // segments per batch (their row / slot words stay in shared memory between the two sweeps)
constexpr int kCovRows = 128;    // rows whose key range is tracked in shared memory (a taller label is processed band by band)

// draw_line of one segment restricted to pixel row y (rasterizer.rs:52-83); calls add_a(x, value) for every touched cell of `a`
// and add_s(x, value) once
template <typename FA, typename FS>
__device__ __forceinline__ void cover_segment_row(const DevSeg& sg, double slope, double rslope, int y, FA add_a, FS add_s) {
    const double y_min = fmin(sg.y0, sg.y1), y_max = fmax(sg.y0, sg.y1);
    const double sign = (sg.y0 <= sg.y1) ? 1.0 : -1.0;
    const double y_bottom = fmax((double)y, y_min);
    const double y_top = fmin((double)(y + 1), y_max);
    const double y_delta = y_top - y_bottom;
    const double x_at_bottom = sg.x0 + (y_bottom - sg.y0) * slope;
    const double x_at_top = sg.x0 + (y_top - sg.y0) * slope;
    const bool flip = !(x_at_bottom <= x_at_top);
    const double x_smallest = flip ? x_at_top : x_at_bottom;
    const double x_largest = flip ? x_at_bottom : x_at_top;
    const int x_to = f64_as_i32(floor(x_largest));
    for (int x = f64_as_i32(floor(x_smallest)); x <= x_to; ++x) {
        const double x_left = fmax((double)x, x_smallest);
        const double x_next = (double)(x + 1);
        const double x_right = fmin(x_next, x_largest);
        double pixel_area = (x_next - x_right) * y_delta;
        const double tw = x_right - x_left;
        if (tw > 0.0) {
            const double y_at_left = sg.y0 + (x_left - sg.x0) * rslope;
            const double y_at_right = sg.y0 + (x_right - sg.x0) * rslope;
            const double th = flip ? (y_top - y_at_left) + (y_top - y_at_right) : (y_at_left - y_bottom) + (y_at_right - y_bottom);
            pixel_area += tw * th / 2.0;
        }
        add_a(x, sign * pixel_area);
        if (x == 0x7fffffff) break;
    }
    add_s(x_to + 1, sign * y_delta);
}

__global__ void __launch_bounds__(32) label_cover_kernel(LabelScene ls) {
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ double s_val[kCovPairs];
    __shared__ unsigned short s_key[kCovPairs];    // window-local key: (row * win_w + column) * 2 + (0: `a`, 1: `s`); 0xffff: unused slot
    __shared__ unsigned short s_order[kCovPairs];  // pair positions sorted by key, stable
    __shared__ unsigned short s_cnt[kCovKeys];
    __shared__ unsigned short s_off[kCovKeys + 1];
    __shared__ unsigned s_seg[32];  // per lane of a chunk: first row inside the band (8 bits) | rows (8) | pair slots per row (16)
    __shared__ unsigned s_pre[33];  // prefix of crossings (rows) over the chunk's segments
    __shared__ unsigned s_slot[32]; // first pair slot of every segment of the chunk
    __shared__ unsigned s_segw[kCovSegs];  // per segment of the batch: first row | rows << 8 | pair slots per row << 16 (0: no crossing in the band)
    __shared__ unsigned char s_umap[kCovDirectUnits];  // crossing of the chunk -> its segment (chunks of few crossings)
    __shared__ unsigned short s_live[kCovKeys];  // the keys of the batch that received pairs, ascending
    __shared__ unsigned s_retry;    // a crossing needed more pair slots than the tight estimate: the batch is redone with the safe one
    __shared__ int s_kmin[kCovRows], s_kmax[kCovRows];
    if (ls.skip_flags && (ls.skip_flags[0] | ls.skip_flags[1])) return;
    const unsigned n_cover = ls.n_cover_dev ? *ls.n_cover_dev : ls.n_cover;
    const unsigned lane = threadIdx.x;
    for (;;) {
        unsigned ci = 0;
        if (lane == 0) ci = atomicAdd(ls.cover_cursor, 1u);
        ci = __shfl_sync(kFull, ci, 0);
        if (ci >= n_cover) break;
        const unsigned slot = ls.cover_list[ci];
        const DevLabel L = ls.labels[slot];
        const int W = L.width, R = L.rows;
        if (R <= 0 || W <= 0) continue;
        double* A = ls.acc_a + L.cell_off;
        double* S = ls.acc_s + L.cell_off;
        const DevSeg* segs = ls.segs + L.seg_begin;
        const unsigned nseg = L.seg_count;
        OSMR_COUNT("lcover.labels", lane == 0);
        OSMR_COUNT("lcover.segments", lane == 0 ? nseg : 0);
        for (int band = 0; band < R; band += kCovRows) {
            const int band_rows = min(kCovRows, R - band);
            const int row_lo = L.ry0 + band;  // pixel row of the band's first row
            // nobody cleared the coverage cells: the band's rows are contiguous
            {
                const size_t c0 = (size_t)band * (size_t)W, c1 = c0 + (size_t)band_rows * (size_t)W;
                for (size_t c = c0 + lane; c < c1; c += 32) {
                    A[c] = 0.0;
                    S[c] = 0.0;
                }
            }
            for (int r = (int)lane; r < band_rows; r += 32) {
                s_kmin[r] = 0x7fffffff;
                s_kmax[r] = (int)0x80000000;
            }
            __syncwarp();
            unsigned sc = 0;  // next segment
            // Pair slots per crossing: the `a` cells it can touch + one `s`.  The tight estimate (pixel extent of the segment + 2)
            // is exact unless a rounding error pushes an interpolated x across a pixel border; the safe one adds a cell on each
            // side.  A batch is formed with the tight estimate (2-3 slots per crossing instead of 5: twice the segments per batch)
            // and redone with the safe one in the rare case that a crossing ran out of slots.
            bool tight = true;
            while (sc < nseg) {
                if (lane == 0) s_retry = 0u;
                // ================= a batch: chunks of 32 segments until the pair slots or the cell window are full =================
                unsigned n_pairs = 0;  // pair slots handed out
                int win_r0 = 0x7fffffff, win_r1 = -1, win_c0 = 0x7fffffff, win_c1 = -1;  // the batch's window (rows of the band, columns)
                int win_w = 0;  // set when the window is frozen (the first chunk that does not fit ends the batch)
                // ---- first sweep: which segments go into the batch, their window ----
                unsigned n_in = 0;
                {
                    unsigned slots = 0;
                    int r0w = 0x7fffffff, r1w = -1, c0w = 0x7fffffff, c1w = -1;
                    while (sc + n_in < nseg) {
                        const unsigned j = sc + n_in + lane;
                        int r0 = 1, r1 = 0, c0 = 0, c1 = -1;
#ifndef OSMR_EMULATED
                        if (j + 64 < nseg) asm volatile("prefetch.global.L2 [%0];" ::"l"(segs + j + 64));  // the chunk after the next one
#endif
                        if (j < nseg) {
                            const DevSeg sg = segs[j];
                            r0 = max(f64_as_i32(floor(fmin(sg.y0, sg.y1))), row_lo) - row_lo;
                            r1 = min(f64_as_i32(floor(fmax(sg.y0, sg.y1))), row_lo + band_rows - 1) - row_lo;
                            // columns the segment can add into: its own x extent, one key of slack on both sides (`s` goes to
                            // x_to + 1; the interpolated x of a row may leave the segment's extent by a rounding error).  Keys
                            // outside the label's columns are dropped when the pair is made.
                            const long long xa = (long long)f64_as_i32(floor(fmin(sg.x0, sg.x1))) - 1 - L.bx0;
                            const long long xb = (long long)f64_as_i32(floor(fmax(sg.x0, sg.x1))) + 2 - L.bx0;
                            c0 = (int)max(0ll, min(xa, (long long)W - 1));
                            c1 = (int)max(0ll, min(xb, (long long)W - 1));
                        }
                        const bool has = j < nseg && r1 >= r0;
                        const unsigned need = has ? (unsigned)(r1 - r0 + 1) * (unsigned)(tight ? max(c1 - c0 - 1, 2) : c1 - c0 + 2) : 0u;  // per row: the `a` cells + one `s`
                        // the chunk is taken as a whole or not at all (except when it is the batch's first: then lane by lane)
                        const unsigned per_row_w = has ? (unsigned)(tight ? max(c1 - c0 - 1, 2) : c1 - c0 + 2) : 0u;
                        const unsigned segw = has ? ((unsigned)r0 | ((unsigned)(r1 - r0 + 1) << 8) | (per_row_w << 16)) : 0u;
                        const unsigned tot = __reduce_add_sync(kFull, need);
                        const int a0 = __reduce_min_sync(kFull, has ? r0 : 0x7fffffff), a1 = __reduce_max_sync(kFull, has ? r1 : -1);
                        const int b0 = __reduce_min_sync(kFull, has ? c0 : 0x7fffffff), b1 = __reduce_max_sync(kFull, has ? c1 : -1);
                        const int nr0 = min(r0w, a0), nr1 = max(r1w, a1), nc0 = min(c0w, b0), nc1 = max(c1w, b1);
                        const long long keys = (nr1 >= nr0) ? 2ll * (nr1 - nr0 + 1) * (nc1 - nc0 + 1) : 0ll;
                        const unsigned n_lanes = min(32u, nseg - (sc + n_in));
                        if (slots + tot <= (unsigned)kCovPairs && keys <= (long long)kCovKeys && n_in + n_lanes <= (unsigned)kCovSegs) {
                            if (lane < n_lanes) s_segw[n_in + lane] = segw;
                            slots += tot;
                            r0w = nr0;
                            r1w = nr1;
                            c0w = nc0;
                            c1w = nc1;
                            n_in += n_lanes;
                            continue;
                        }
                        if (n_in) break;  // the batch is what fitted so far
                        // the very first chunk does not fit: take its longest prefix of lanes that does (at least one lane: a
                        // single segment that is too large for a batch is processed by the slow path below)
                        unsigned take = 0;
                        for (unsigned t = 1; t <= n_lanes; ++t) {
                            unsigned ts = 0;
                            int p0 = 0x7fffffff, p1 = -1, q0 = 0x7fffffff, q1 = -1;
                            const bool in = lane < t;
                            unsigned nd = in ? need : 0u;
                            int x0 = in && has ? r0 : 0x7fffffff, x1 = in && has ? r1 : -1, y0 = in && has ? c0 : 0x7fffffff, y1 = in && has ? c1 : -1;
                            for (int o = 16; o > 0; o >>= 1) {
                                nd += __shfl_xor_sync(kFull, nd, o);
                                x0 = min(x0, __shfl_xor_sync(kFull, x0, o));
                                x1 = max(x1, __shfl_xor_sync(kFull, x1, o));
                                y0 = min(y0, __shfl_xor_sync(kFull, y0, o));
                                y1 = max(y1, __shfl_xor_sync(kFull, y1, o));
                            }
                            ts = nd;
                            p0 = x0;
                            p1 = x1;
                            q0 = y0;
                            q1 = y1;
                            const long long kk = (p1 >= p0) ? 2ll * (p1 - p0 + 1) * (q1 - q0 + 1) : 0ll;
                            if (ts <= (unsigned)kCovPairs && kk <= (long long)kCovKeys) {
                                take = t;
                                r0w = p0;
                                r1w = p1;
                                c0w = q0;
                                c1w = q1;
                            } else {
                                break;
                            }
                        }
                        n_in = take;
                        if (lane < take) s_segw[lane] = segw;
                        break;
                    }
                    win_r0 = r0w;
                    win_r1 = r1w;
                    win_c0 = c0w;
                    win_c1 = c1w;
                }
                if (n_in == 0u) {
                    // ---- slow path: ONE segment larger than a batch (a very long edge): lane 0 adds it straight into the arrays ----
                    if (lane == 0) {
                        const DevSeg sg = segs[sc];
                        const double slope = (sg.x1 - sg.x0) / (sg.y1 - sg.y0), rslope = 1.0 / slope;
                        const int r0 = max(f64_as_i32(floor(fmin(sg.y0, sg.y1))), row_lo), r1 = min(f64_as_i32(floor(fmax(sg.y0, sg.y1))), row_lo + band_rows - 1);
                        for (int y = r0; y <= r1; ++y) {
                            const int r = y - row_lo;
                            double* a = A + (size_t)(band + r) * W;
                            double* sacc = S + (size_t)(band + r) * W;
                            cover_segment_row(
                                sg, slope, rslope, y,
                                [&](int x, double v) {
                                    const int cx = x - L.bx0;
                                    if (cx >= 0 && cx < W) a[cx] += v;
                                    s_kmin[r] = min(s_kmin[r], x);
                                    s_kmax[r] = max(s_kmax[r], x);
                                },
                                [&](int x, double v) {
                                    const int cx = x - L.bx0;
                                    if (cx >= 0 && cx < W) sacc[cx] += v;
                                    s_kmin[r] = min(s_kmin[r], x);
                                    s_kmax[r] = max(s_kmax[r], x);
                                });
                        }
                    }
                    __syncwarp();
                    sc += 1;
                    continue;
                }
                OSMR_COUNT("lcover.batches", lane == 0);
                win_w = win_c1 >= win_c0 ? win_c1 - win_c0 + 1 : 1;
                const int n_keys = (win_r1 >= win_r0) ? 2 * (win_r1 - win_r0 + 1) * win_w : 0;
                for (int k = (int)lane; k < n_keys; k += 32) s_cnt[k] = 0;
                __syncwarp();
                // ---- A: the area arithmetic, a lane per (segment, row) crossing; pairs land in order ----
                for (unsigned base = 0; base < n_in; base += 32) {
                    // (rows and pair slots of the chunk's segments: the words of the first sweep)
                    const unsigned segw = base + lane < n_in ? s_segw[base + lane] : 0u;
                    const unsigned nr = (segw >> 8) & 0xffu, per_row = segw >> 16;
                    // two prefix sums in one: crossings (rows) in the low half, pair slots in the high half (a chunk has at most
                    // 32 * 128 crossings and kCovPairs slots).  Unit u of segment q starts at base_slot + (slots of the segments
                    // before q) + (u's row index inside q) * per_row(q).
                    unsigned both = nr | ((nr * per_row) << 16);
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned y = __shfl_up_sync(kFull, both, o);
                        if ((int)lane >= o) both += y;
                    }
                    const unsigned incl = both & 0xffffu, slot_incl = both >> 16;
                    __syncwarp();
                    s_seg[lane] = segw;
                    s_pre[lane] = incl - nr;
                    if (lane == 31) s_pre[32] = incl;
                    s_slot[lane] = n_pairs + slot_incl - nr * per_row;
                    const unsigned chunk_slots = __shfl_sync(kFull, slot_incl, 31);
                    // crossing -> segment: the usual chunk (glyph segments cross one or two rows) gets a direct map, a chunk with
                    // long segments a binary search in the prefix
                    const unsigned n_units = __shfl_sync(kFull, incl, 31);
                    const bool direct = n_units <= (unsigned)kCovDirectUnits;
                    if (direct)
                        for (unsigned i = 0; i < nr; ++i) s_umap[incl - nr + i] = (unsigned char)lane;
                    __syncwarp();
                    for (unsigned ub = 0; ub < n_units; ub += 32) {
                        const unsigned un = ub + lane;
                        if (un < n_units) {
                            int lo = 0;  // largest q with s_pre[q] <= un
                            if (direct) {
                                lo = (int)s_umap[un];
                            } else {
                                int hi = 32;
                                while (hi - lo > 1) {
                                    const int mid = (lo + hi) >> 1;
                                    if (s_pre[mid] <= un)
                                        lo = mid;
                                    else
                                        hi = mid;
                                }
                            }
                            const unsigned sgw = s_seg[lo];
                            const int q_r0 = (int)(sgw & 0xffu);
                            const unsigned q_per = sgw >> 16;
                            const unsigned row_i = un - s_pre[lo];
                            const int r = q_r0 + (int)row_i;
                            const unsigned slot0 = s_slot[lo] + row_i * q_per;
                            const DevSeg sg = segs[sc + base + (unsigned)lo];
                            const double slope = (sg.x1 - sg.x0) / (sg.y1 - sg.y0);  // rasterizer.rs:34-35
                            const double rslope = 1.0 / slope;
                            unsigned w = 0;  // slots written
                            int kmn = 0x7fffffff, kmx = (int)0x80000000;
                            auto put = [&](int x, double v, unsigned arr) {
                                const int cx = x - L.bx0;
                                kmn = min(kmn, x);
                                kmx = max(kmx, x);
                                if (cx < 0 || cx >= W) return;  // a key outside the label's columns: tracked, never stored
                                if (w >= q_per && tight) {  // the tight slot estimate was one short: redo the batch with the safe one
                                    s_retry = 1u;
                                    return;
                                }
                                if (cx < win_c0 || cx > win_c1 || w >= q_per) {  // outside the proven window: never silent
                                    atomicOr(ls.err_flag, 1u);
                                    return;
                                }
                                const unsigned key = (unsigned)(((r - win_r0) * win_w + (cx - win_c0)) * 2) + arr;
                                s_val[slot0 + w] = v;
                                s_key[slot0 + w] = (unsigned short)key;
                                atomicAdd(reinterpret_cast<unsigned*>(&s_cnt[key & ~1u]), (key & 1u) ? 0x10000u : 1u);  // two 16-bit counters per word
                                ++w;
                            };
                            cover_segment_row(
                                sg, slope, rslope, row_lo + r, [&](int x, double v) { put(x, v, 0u); }, [&](int x, double v) { put(x, v, 1u); });
                            for (; w < q_per; ++w) s_key[slot0 + w] = 0xffffu;
                            if (kmn != 0x7fffffff) atomicMin(&s_kmin[r], kmn);
                            if (kmx != (int)0x80000000) atomicMax(&s_kmax[r], kmx);
                        }
                    }
                    n_pairs += chunk_slots;
                    __syncwarp();
                }
                OSMR_COUNT("lcover.pairs", lane == 0 ? n_pairs : 0);
                __syncwarp();
                const unsigned retry = s_retry;
                __syncwarp();  // (lane 0 clears the flag at the top of the loop: everybody has read it by then)
                if (tight && retry) {  // (kmin / kmax updates are idempotent; everything else of the batch is rebuilt)
                    OSMR_COUNT("lcover.retries", lane == 0);
                    tight = false;
                    continue;
                }
                // ---- B: stable counting sort of the pairs by key ----
                unsigned n_live = 0;
                {
                    unsigned carry = 0;
                    for (int kb = 0; kb < n_keys; kb += 32) {
                        const int k = kb + (int)lane;
                        const unsigned v = k < n_keys ? s_cnt[k] : 0u;
                        unsigned incl = v;
                        for (int o = 1; o < 32; o <<= 1) {
                            const unsigned y = __shfl_up_sync(kFull, incl, o);
                            if ((int)lane >= o) incl += y;
                        }
                        if (k < n_keys) {
                            s_off[k] = (unsigned short)(carry + incl - v);
                            s_cnt[k] = (unsigned short)(carry + incl - v);  // becomes the running cursor
                        }
                        carry += __shfl_sync(kFull, incl, 31);
                        const unsigned lv = __ballot_sync(kFull, v != 0u);
                        if (v) s_live[n_live + (unsigned)__popc(lv & ((1u << lane) - 1u))] = (unsigned short)k;
                        n_live += (unsigned)__popc(lv);
                    }
                    if (lane == 0) s_off[n_keys] = (unsigned short)carry;
                }
                __syncwarp();
                for (unsigned pb = 0; pb < n_pairs; pb += 32) {
                    const unsigned p = pb + lane;
                    const unsigned key = p < n_pairs ? (unsigned)s_key[p] : 0xffffu;
                    const bool live = key != 0xffffu;
                    // lanes with the same key, in lane (= reference) order; dead lanes get keys of their own
                    const unsigned peers = __match_any_sync(kFull, live ? key : (0x10000u + lane));
                    if (live) s_order[(unsigned)s_cnt[key] + (unsigned)__popc(peers & ((1u << lane) - 1u))] = (unsigned short)p;
                    __syncwarp();
                    if (live && (peers & ((1u << lane) - 1u)) == 0u) s_cnt[key] = (unsigned short)(s_cnt[key] + __popc(peers));
                    __syncwarp();
                }
                // ---- C: a lane per cell that received pairs, its pairs in order ----
                for (unsigned kb = 0; kb < n_live; kb += 32) {
                    const unsigned i = kb + lane;
                    if (i < n_live) {
                        const int k = (int)s_live[i];
                        const unsigned u0 = s_off[k], u1 = s_off[k + 1];
                        const int cell = k >> 1;
                        const int r = win_r0 + cell / win_w, cx = win_c0 + cell % win_w;
                        double* dst = ((k & 1) ? S : A) + (size_t)(band + r) * W + cx;
                        double acc = *dst;
                        for (unsigned u = u0; u < u1; ++u) acc += s_val[s_order[u]];
                        *dst = acc;
                    }
                }
                __syncwarp();
                sc += n_in;
                tight = true;
            }
            // ---- sweep (save_to_figure): the lane of a row, left to right over the touched keys ----
#pragma unroll 1
            for (int r = (int)lane; r < band_rows; r += 32) {
                const int lo = s_kmin[r], hi = s_kmax[r];
                ls.kmin[L.row_first + band + r] = lo;
                ls.kmax[L.row_first + band + r] = hi;
                if (lo <= hi) {
                    double run = 0.0;
                    double* a = A + (size_t)(band + r) * W;
                    const double* sacc = S + (size_t)(band + r) * W;
                    for (int x = lo; x <= hi; ++x) {
                        const int c = x - L.bx0;
                        const bool inside = c >= 0 && c < W;  // the bbox covers every key; defensive
                        run += inside ? sacc[c] : 0.0;
                        const double total = fmin((inside ? a[c] : 0.0) + run, 1.0);
                        if (inside) a[c] = total;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// label_commit_kernel (a9 + f1): one CTA per tile walks the tile's labels IN ORDER (the collision rule is greedy):
// icon blit (labeler.rs:91-106) and the text coverage computed by label_cover_kernel are tested against the pixels of
// earlier successful labels over the 3x3 label canvas (tile_pixels.rs:131-148); survivors claim their pixels and
// leave their pending pixel for the export of raster_kernel.
// ------------------------------------------------------------------------------------------------------
constexpr int kLabelThreads = 256;

__global__ void __launch_bounds__(kLabelThreads) label_commit_kernel(LabelScene ls) {
    // (two flags: the icon's is written before the barrier that the text phase reads it behind, the text's is written while other
    // warps may still be reading the icon's -- one flag would race)
    __shared__ volatile int fail, fail_text;
    const unsigned tile = blockIdx.x;
    const int D = ls.D, E = 3 * D;
    const unsigned occ_words = (unsigned)((size_t)E * E / 32);
    unsigned* occ = ls.occ + (size_t)tile * occ_words;
    LabelPix* plane = ls.plane + (size_t)tile * D * D;
    for (unsigned i = threadIdx.x; i < occ_words; i += kLabelThreads) occ[i] = 0u;
    unsigned* pmask = ls.pmask + (size_t)tile * (size_t)(D * D / 32);
    for (int i = threadIdx.x; i < D * D / 32; i += kLabelThreads) pmask[i] = 0u;  // (the plane itself keeps stale records)
    __syncthreads();
    auto in_canvas = [&](int x, int y) { return x >= -D && x <= 2 * D - 1 && y >= -D && y <= 2 * D - 1; };  // labels_bb
    auto occ_index = [&](int x, int y) { return (size_t)(y + D) * E + (size_t)(x + D); };

    const bool abandoned = ls.skip_flags && (ls.skip_flags[0] | ls.skip_flags[1]);  // (the planes above are still cleared)
    const unsigned li_begin = ls.label_begin[tile];
    const unsigned li_end = abandoned ? li_begin : (ls.label_cnt ? li_begin + ls.label_cnt[tile] : ls.label_begin[tile + 1]);
    for (unsigned li = li_begin; li < li_end; ++li) {
        const DevLabel L = ls.labels[li];
        if (threadIdx.x == 0) {
            fail = 0;
            fail_text = 0;
        }
        __syncthreads();
        int iw = 0, ih = 0;
        if (L.icon >= 0) {  // labeler.rs:94-103: every texel (transparent ones too) claims its pixel
            iw = (int)ls.icons[L.icon].w;
            ih = (int)ls.icons[L.icon].h;
            for (int p = threadIdx.x; p < iw * ih; p += kLabelThreads) {
                int x = wadd(L.ix, p / ih), y = wadd(L.iy, p % ih);
                if (in_canvas(x, y)) {
                    size_t b = occ_index(x, y);
                    if (occ[b >> 5] & (1u << (b & 31))) fail = 1;
                }
            }
        }
        __syncthreads();
        const int W = L.width, R = L.rows;
        const bool has_text = L.seg_count != 0 && R > 0 && W > 0 && !fail;  // a failed icon fails the label before its text
        const double* tot = ls.acc_a + L.cell_off;
        const int* kmin = ls.kmin + L.row_first;
        const int* kmax = ls.kmax + L.row_first;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (has_text) {  // a warp per stored row, lanes over the stripe's key range (nothing is set outside it)
            // (a warp stops at its own first collision; the shared flag is only written here and read behind the barrier)
            bool mine = false;
            for (int r = warp; r < R && !mine; r += kLabelThreads / 32) {
                const int y = L.ry0 + r;
                const int lo = max(kmin[r], max(L.bx0, -D)), hi = min(kmax[r], min(L.bx0 + W - 1, 2 * D - 1));
                if (lo > hi) continue;  // untouched stripe (kmin = INT_MAX)
                const double* row = tot + (size_t)r * W - L.bx0;
                bool hit = false;
                for (int x = lo + lane; x <= hi; x += 32) {
                    if (row[x] > 0.0) {
                        size_t b = occ_index(x, y);
                        if (occ[b >> 5] & (1u << (b & 31))) hit = true;
                    }
                }
                mine = __any_sync(0xffffffffu, hit);
            }
            if (mine && lane == 0) fail_text = 1;
        }
        __syncthreads();
        if (!fail && !fail_text) {  // bump_label_generation(true): the label's pixels become final
            if (L.icon >= 0) {
                const unsigned tex0 = ls.icons[L.icon].off;
                for (int p = threadIdx.x; p < iw * ih; p += kLabelThreads) {
                    int xo = p / ih, yo = p % ih;
                    int x = wadd(L.ix, xo), y = wadd(L.iy, yo);
                    if (!in_canvas(x, y)) continue;
                    size_t b = occ_index(x, y);
                    atomicOr(&occ[b >> 5], 1u << (b & 31));
                    if (x >= 0 && x < D && y >= 0 && y < D) {
                        LabelPix px;
                        px.alpha = 0.0;
                        px.src = 0x80000000u | (tex0 + (unsigned)(yo * iw + xo));
                        px.pad = 0;
                        plane[(size_t)y * D + x] = px;
                        atomicOr(&pmask[((size_t)y * D + x) >> 5], 1u << (((size_t)y * D + x) & 31));
                    }
                }
            }
            __syncthreads();  // text pixels overwrite icon pixels of the same label (later set_label_pixel wins)
            if (has_text) {
                for (int r = warp; r < R; r += kLabelThreads / 32) {
                    const int y = L.ry0 + r;
                    const int lo = max(kmin[r], max(L.bx0, -D)), hi = min(kmax[r], min(L.bx0 + W - 1, 2 * D - 1));
                    if (lo > hi) continue;
                    const double* row = tot + (size_t)r * W - L.bx0;
                    for (int x = lo + lane; x <= hi; x += 32) {
                        const double total = row[x];
                        if (!(total > 0.0)) continue;
                        size_t b = occ_index(x, y);
                        atomicOr(&occ[b >> 5], 1u << (b & 31));
                        if (x >= 0 && x < D && y >= 0 && y < D) {
                            LabelPix px;
                            px.alpha = total;
                            px.src = 0x40000000u | (L.rgb & 0xffffffu);
                            px.pad = 0;
                            plane[(size_t)y * D + x] = px;
                            atomicOr(&pmask[((size_t)y * D + x) >> 5], 1u << (((size_t)y * D + x) & 31));
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace osmr

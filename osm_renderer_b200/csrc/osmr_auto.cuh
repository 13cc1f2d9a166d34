// osmr_auto.cuh -- SURVEY.md 8(f) row f3: tile -> candidate entities -> ordered styled areas ON THE DEVICE.
//
// Replaces, for the area passes, what the reference does on the CPU before it draws a tile:
//   GeodataReader::get_entities_in_tile_with_neighbors   src/geodata/reader.rs:60-180   (3x3 neighbourhood, dedup)
//   Styler::style_areas / style_entities                  src/mapcss/styler.rs:115-203   (entity -> its styles)
//   compare_styled_entities + the way/multipolygon merge  src/mapcss/styler.rs:172-203,246-272   (painter's order)
// MapCSS selector matching is string work and stays on the host -- exactly like the reference's own style cache
// (src/mapcss/style_cache.rs:68-87, keyed by entity and zoom): per zoom the host hands over, for every way and
// multipolygon, the id of its style list ("class") and per class the list of (style, order) where `order` is the dense
// rank of the sort key (layer, is_foreground_fill, z_index) of styler.rs:246-272.
//
//   auto_bound_kernel    per tile: index records of the 3x3 neighbourhood -> upper bound of the candidate count
//   auto_gather_kernel   per tile: candidates, each exactly once (see "ownership"), minus the ones that cannot touch
//                        the tile (same conservative bbox + reach rule as plan_ops_kernel) -> compact list + #styled areas
//   auto_sort_kernel     per tile: expand to styled areas, sort by (order, global id, mp-before-way, local id, position in
//                        the style list), write osmr_styled_area records in the reference's order
//
// Ownership (dedup without sorting): the importer lists an entity in EVERY z18 tile of the bounding box of its nodes' z18
// tiles (src/geodata/saver.rs:194-226; osmr_set_geodata verifies this for the image it is given).  Inside a query rectangle
// the entity therefore appears in a sub-rectangle of index records, and the record at that sub-rectangle's minimum corner
// (max(entity.min_x, rect.min_x), max(entity.min_y, rect.min_y)) is the only one that emits it.
//
// Dropping areas that cannot touch the tile does not change the image: the reference styles them and then draws nothing.
// The surviving areas keep the reference's relative order, which is all the compositor can see.
#pragma once
#include "osmr_kernels.cuh"

namespace osmr {

struct AutoScene {
    // tile index of the .bin (sorted by (x, y), reader.rs:135-180)
    const uint2* idx_xy;
    const uint2* idx_w;  // (off, len) of the way ids of the record, into ints
    const uint2* idx_m;  // multipolygon ids
    unsigned n_idx;
    const uint2* way_min_tile;  // minimum corner of the entity's z18 tile rectangle
    const uint2* mp_min_tile;
    const unsigned* way_rank;  // rank of (global id, is_way, local id) among all ways and multipolygons
    const unsigned* mp_rank;
    const unsigned* rank_entity;  // rank -> osmr_styled_area.entity
    // per-zoom style tables (osmr_set_zoom_styles)
    const unsigned* way_class;
    const unsigned* mp_class;
    const unsigned* class_begin;
    const osmr_class_style* class_styles;
    unsigned n_classes, n_class_styles;
    int* class_reach;  // per class: -1 no styled area can be visible, else the culling margin in pixels (class_reach_kernel)
    // per call
    unsigned* bound;      // per tile (+1): upper bound of candidates, then its exclusive scan
    unsigned* cand;       // compact candidates (entity codes) of tile t at cand[bound[t] ..]
    unsigned* cand_cnt;   // per tile
    unsigned* inst_cnt;   // per tile (+1): styled areas, then the exclusive scan == area_begin
    unsigned long long* big_keys;  // scratch for tiles with more styled areas than fit in shared memory
    unsigned long long big_cap;
    osmr_styled_area* areas_out;
};

constexpr int kAutoThreads = 256;
constexpr int kAutoWarps = kAutoThreads / 32;
constexpr unsigned kAutoSortCap = 4096;  // keys sorted in shared memory (32 KB)
constexpr unsigned kAutoOrderBits = 20, kAutoWithinBits = 12;

// per class: can any of its styles draw something, and how far beyond the entity's bbox (plan_ops_kernel's rule)
__global__ void class_reach_kernel(Scene s, AutoScene a) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_classes) return;
    int reach = -1;
    for (unsigned i = a.class_begin[c]; i < a.class_begin[c + 1]; ++i) {
        const unsigned sid = a.class_styles[i].style;
        if (sid >= s.n_styles) {
            atomicOr(&s.counters[CNT_BAD_INPUT], 1u);
            continue;
        }
        const osmr_style& st = s.styles[sid];
        if ((st.flags & OSMR_STYLE_FILL_COLOR) || ((st.flags & OSMR_STYLE_FILL_IMAGE) && st.fill_image >= 0)) reach = max(reach, 0);
        for (int pass = 1; pass <= 2; ++pass) {
            LineParams lp;
            if (!line_params(s, st, pass, lp)) continue;
            int r = line_reach(lp.width / 2.0);
            if (is_non_trivial_cap(lp.cap)) r = 2 * r;
            reach = max(reach, r);
        }
    }
    a.class_reach[c] = reach;
}

struct AutoRect {
    long long xa, xb, ya, yb;  // z18 tile range of the 3x3 neighbourhood (reader.rs:63-73, tile.rs:63-73)
};

__device__ __forceinline__ AutoRect auto_rect(const osmr_tile& t) {
    const long long mul = 1ll << (18 - (int)t.zoom);
    AutoRect r;
    r.xa = max(((long long)t.x - 1) * mul, 0ll);  // a neighbour beyond the map edge wraps in u32 and selects nothing
    r.ya = max(((long long)t.y - 1) * mul, 0ll);
    r.xb = ((long long)t.x + 2) * mul - 1;
    r.yb = ((long long)t.y + 2) * mul - 1;
    return r;
}

// first index record with (x, y) >= (col, ya)   (reader.rs:135-180 `find_smallest_feasible_index`)
__device__ __forceinline__ unsigned auto_lower_bound(const AutoScene& a, long long col, long long ya) {
    unsigned lo = 0, hi = a.n_idx;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        const uint2 v = a.idx_xy[mid];
        const bool ge = (long long)v.x > col || ((long long)v.x == col && (long long)v.y >= ya);
        if (ge)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

template <bool kGather>
__device__ __forceinline__ void auto_visit_tile(const Scene& s, const AutoScene& a, unsigned t, unsigned* sh_counts) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const osmr_tile tile = s.tiles[t];
    const AutoRect rc = auto_rect(tile);
    const TileXform xf = make_xform(tile);
    const int D = s.D;
    unsigned local_bound = 0;
    // one warp per column of the rectangle; only columns that exist in the index cost more than a binary search
    for (long long col = rc.xa + warp; col <= rc.xb; col += kAutoWarps) {
        unsigned i = auto_lower_bound(a, col, rc.ya);
        if (i >= a.n_idx) break;  // nothing at or beyond this column
        const uint2 first = a.idx_xy[i];
        if ((long long)first.x != col) {
            // jump to the next populated column this warp owns (reader.rs:160-178 skips empty columns the same way)
            if ((long long)first.x > rc.xb) break;
            const long long nxt = (long long)first.x;
            const long long k = (nxt - rc.xa - warp + kAutoWarps - 1) / kAutoWarps;
            col = rc.xa + warp + (k - 1) * kAutoWarps;  // the loop increment lands on the first owned column >= nxt
            continue;
        }
        for (; i < a.n_idx; ++i) {
            const uint2 xy = a.idx_xy[i];
            if ((long long)xy.x != col || (long long)xy.y > rc.yb) break;
            const uint2 wl = a.idx_w[i], ml = a.idx_m[i];
            if (!kGather) {
                local_bound += wl.y + ml.y;
                continue;
            }
            for (unsigned j = lane; j < wl.y + ml.y; j += 32) {
                const bool is_mp = j >= wl.y;
                const unsigned e = is_mp ? s.ints[ml.x + (j - wl.y)] : s.ints[wl.x + j];
                if (e >= (is_mp ? s.n_mps : s.n_ways)) {
                    atomicOr(&s.counters[CNT_BAD_INPUT], 1u);
                    continue;
                }
                const uint2 mt = is_mp ? a.mp_min_tile[e] : a.way_min_tile[e];
                if ((long long)xy.x != max((long long)mt.x, rc.xa) || (long long)xy.y != max((long long)mt.y, rc.ya)) continue;  // not the owner
                if (is_mp && s.mps[e].y == 0) continue;  // reader.rs:86-93
                const unsigned c = is_mp ? a.mp_class[e] : a.way_class[e];
                if (c >= a.n_classes) continue;  // no style list
                const unsigned n_inst = a.class_begin[c + 1] - a.class_begin[c];
                const int reach = a.class_reach[c];
                if (n_inst == 0 || reach < 0) continue;
                // pixel bbox of all its points, exactly like area_bbox_kernel
                const EntBox eb = (is_mp ? s.mp_box : s.way_box)[e];
                int x0, y0, x1, y1;
                entity_pixel_bbox(eb, xf, x0, y0, x1, y1);
                const unsigned npts = eb.npts;
                if (npts < 2) continue;
                if ((long long)x0 - reach > D - 1 || (long long)x1 + reach < 0 || (long long)y0 - reach > D - 1 || (long long)y1 + reach < 0) continue;
                const unsigned pos = atomicAdd(&sh_counts[0], 1u);
                atomicAdd(&sh_counts[1], n_inst);
                a.cand[a.bound[t] + pos] = is_mp ? (e | OSMR_AREA_MULTIPOLYGON) : e;
            }
        }
    }
    if (!kGather) {
        for (int o = 16; o > 0; o >>= 1) local_bound += __shfl_xor_sync(0xffffffffu, local_bound, o);
        if (lane == 0 && local_bound) atomicAdd(&sh_counts[0], local_bound);
    }
}

__global__ void __launch_bounds__(kAutoThreads) auto_bound_kernel(Scene s, AutoScene a) {
    __shared__ unsigned counts[2];
    if (threadIdx.x < 2) counts[threadIdx.x] = 0;
    __syncthreads();
    auto_visit_tile<false>(s, a, blockIdx.x, counts);
    __syncthreads();
    if (threadIdx.x == 0) a.bound[blockIdx.x] = counts[0];
}

__global__ void __launch_bounds__(kAutoThreads) auto_gather_kernel(Scene s, AutoScene a) {
    __shared__ unsigned counts[2];
    if (threadIdx.x < 2) counts[threadIdx.x] = 0;
    __syncthreads();
    auto_visit_tile<true>(s, a, blockIdx.x, counts);
    __syncthreads();
    if (threadIdx.x == 0) {
        a.cand_cnt[blockIdx.x] = counts[0];
        a.inst_cnt[blockIdx.x] = counts[1];
    }
}

// exclusive scan of v[0..n) in place, v[n] = total (single CTA; n is a tile count); sets *overflow when the total leaves u32
__global__ void __launch_bounds__(1024) auto_scan_kernel(unsigned* v, unsigned n, unsigned* overflow) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n; base += blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const unsigned long long x = i < n ? v[i] : 0u;
        unsigned long long incl = x;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned long long before = carry;
        for (unsigned k = 0; k < warp; ++k) before += warp_tot[k];
        const unsigned long long excl = before + incl - x;
        if (i < n) v[i] = (unsigned)excl;
        if (excl + x > 0xfffffff0ull) atomicOr(overflow, 1u);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) v[n] = (unsigned)carry;
}

__device__ __forceinline__ void bitonic_sort(unsigned long long* keys, unsigned n_pad) {
    for (unsigned k = 2; k <= n_pad; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const unsigned p = i ^ j;
                if (p > i) {
                    const unsigned long long x = keys[i], y = keys[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        keys[i] = y;
                        keys[p] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kAutoThreads) auto_sort_kernel(Scene s, AutoScene a) {
    __shared__ unsigned long long sh_keys[kAutoSortCap];
    __shared__ unsigned cursor;
    __shared__ unsigned long long big_base;
    const unsigned t = blockIdx.x;
    const unsigned out0 = a.inst_cnt[t];
    const unsigned n = a.inst_cnt[t + 1] - out0;
    if (n == 0) return;
    unsigned n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long* keys = sh_keys;
    if (threadIdx.x == 0) {
        cursor = 0;
        big_base = 0;
        if (n_pad > kAutoSortCap) {  // sort in global scratch
            unsigned long long* used = reinterpret_cast<unsigned long long*>(&s.counters[CNT_WALK_ALPHA]);  // free at this stage
            big_base = atomicAdd(used, (unsigned long long)n_pad);
            if (big_base + n_pad > a.big_cap) {
                atomicOr(&s.counters[CNT_OVERFLOW], 8u);
                big_base = ~0ull;
            }
        }
    }
    __syncthreads();
    if (n_pad > kAutoSortCap) {
        if (big_base == ~0ull) return;  // the host grows the scratch and redoes the call
        keys = a.big_keys + big_base;
    }
    for (unsigned i = n + threadIdx.x; i < n_pad; i += blockDim.x) keys[i] = ~0ull;
    // expand the candidates to sort keys: order | rank of (global id, mp-before-way, local id) | position in the style list
    const unsigned* cand = a.cand + a.bound[t];
    const unsigned n_cand = a.cand_cnt[t];
    for (unsigned ci = threadIdx.x; ci < n_cand; ci += blockDim.x) {
        const unsigned code = cand[ci];
        const bool is_mp = (code & OSMR_AREA_MULTIPOLYGON) != 0;
        const unsigned e = code & ~OSMR_AREA_MULTIPOLYGON;
        const unsigned c = is_mp ? a.mp_class[e] : a.way_class[e];
        const unsigned b = a.class_begin[c], cnt = a.class_begin[c + 1] - b;
        const unsigned long long rank = is_mp ? a.mp_rank[e] : a.way_rank[e];
        const unsigned at = atomicAdd(&cursor, cnt);
        for (unsigned w = 0; w < cnt; ++w)
            keys[at + w] = ((unsigned long long)a.class_styles[b + w].order << (32 + kAutoWithinBits)) | (rank << kAutoWithinBits) | w;
    }
    __syncthreads();
    bitonic_sort(keys, n_pad);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = keys[i];
        const unsigned code = a.rank_entity[(unsigned)((k >> kAutoWithinBits) & 0xffffffffull)];
        const bool is_mp = (code & OSMR_AREA_MULTIPOLYGON) != 0;
        const unsigned e = code & ~OSMR_AREA_MULTIPOLYGON;
        const unsigned c = is_mp ? a.mp_class[e] : a.way_class[e];
        osmr_styled_area ar;
        ar.entity = code;
        ar.style = a.class_styles[a.class_begin[c] + (unsigned)(k & ((1u << kAutoWithinBits) - 1u))].style;
        a.areas_out[out0 + i] = ar;
    }
}

}  // namespace osmr

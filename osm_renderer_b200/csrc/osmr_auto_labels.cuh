// osmr_auto_labels.cuh -- SURVEY.md 8(f) row f3 for the LABEL pass: tile -> candidate entities -> ordered label generations
// ON THE DEVICE (the styled-area half is osmr_auto.cuh).
//
// Replaces what the reference does on the CPU between the area passes and draw_labels:
//   GeodataReader::get_entities_in_tile_with_neighbors   src/geodata/reader.rs:60-100    (nodes, ways, multipolygons of the 3x3
//                                                                                          neighbourhood, each once, by local id)
//   Styler::style_areas(.., for_labels = true)            src/mapcss/styler.rs:168-203    (sorted by layer, z_index, global id --
//                                                                                          is_foreground_fill is ignored, :263 --,
//                                                                                          multipolygon first on ties)
//   Styler::style_entities(nodes, zoom, true)             src/mapcss/styler.rs:115-166, src/draw/drawer.rs:106-119
// and hands draw_labels' input (drawer.rs:221-262: the styled areas, then the styled nodes) to label_select_kernel as the
// osmr_label lists osmr_draw_tiles_labeled would have received from the host.
//
// Selector matching stays on the host as per-zoom classes, like the area half (osmr_set_zoom_label_styles): per node, way and
// multipolygon the id of its style list, per class the (label style, order) pairs with `order` the dense rank of (layer, z_index).
//
// Generations that cannot draw or collide -- no icon, and no text because the style has none or the entity lacks the tag -- are
// dropped here (the rule of label_select_kernel): an empty successful generation changes no pixel (labeler.rs:16-37), and the
// survivors keep the reference's relative order.  No geometric culling: a label anywhere on the 3x3 canvas can win a collision
// against one that reaches the tile.
//
//   autol_bound_kernel   per tile: index records of the 3x3 neighbourhood -> upper bound of the candidate count
//   autol_gather_kernel  per tile: entities with at least one live generation, each exactly once (nodes sit in one index record,
//                        ways / multipolygons by the ownership rule of osmr_auto.cuh) + the number of live generations
//   autol_sort_kernel    per tile: expand to generations, sort by (node?, order, global id, mp-before-way, local id, position in
//                        the style list), write osmr_label records
#pragma once
#include "osmr_auto.cuh"
#include "osmr_labels_dev.cuh"

namespace osmr {

struct AutoLabelScene {
    // tile index of the .bin
    const uint2* idx_xy;
    const uint2* idx_n;  // (off, len) of the node ids of the record, into ints
    const uint2* idx_w;
    const uint2* idx_m;
    unsigned n_idx;
    const uint2* way_min_tile;
    const uint2* mp_min_tile;
    const unsigned* way_rank;  // as in AutoScene
    const unsigned* mp_rank;
    const unsigned* rank_entity;
    const unsigned* node_rank;         // rank of (global id, local id) among the nodes
    const unsigned* node_rank_entity;  // rank -> node index
    // per-zoom label classes (osmr_set_zoom_label_styles)
    const unsigned* node_class;
    const unsigned* way_class;
    const unsigned* mp_class;
    const unsigned* class_begin;
    const osmr_class_style* class_styles;  // .style indexes the table of osmr_set_label_styles
    unsigned n_classes;
    // resident label tables (LabelDev): what decides whether a generation is live
    const DevLabelStyle* lstyles;
    unsigned n_lstyles;
    const unsigned* text_id;
    unsigned ent_total, way_base, mp_base;
    // per call
    unsigned* bound;
    unsigned* cand;
    unsigned* cand_cnt;
    unsigned* inst_cnt;  // per tile (+1): live generations, then the exclusive scan == label_begin
    unsigned long long* big_keys;
    unsigned long long big_cap;
    osmr_label* labels_out;
};

constexpr unsigned kAutoLabelNodeOrder = 1u << (kAutoOrderBits - 1);  // nodes follow the areas: top bit of the order field

// label_select_kernel's rule for one (entity, style)
__device__ __forceinline__ bool autol_live(const Scene& s, const AutoLabelScene& a, unsigned sid, unsigned slot) {
    if (sid >= a.n_lstyles) {
        atomicOr(&s.counters[CNT_BAD_INPUT], 1u);
        return false;
    }
    const DevLabelStyle st = a.lstyles[sid];
    if (st.icon >= 0) return true;  // (an icon index beyond the table is label_select_kernel's to report)
    if ((st.flags & OSMR_LSTYLE_TEXT) && (st.flags & OSMR_LSTYLE_FONT_SIZE) && st.key >= 0)
        return a.text_id[(size_t)st.key * a.ent_total + slot] != 0xffffffffu;
    return false;
}

__device__ __forceinline__ unsigned autol_live_count(const Scene& s, const AutoLabelScene& a, unsigned c, unsigned slot) {
    unsigned n = 0;
    for (unsigned i = a.class_begin[c]; i < a.class_begin[c + 1]; ++i) n += autol_live(s, a, a.class_styles[i].style, slot) ? 1u : 0u;
    return n;
}

// entity code (osmr_label.entity) -> class and slot in the text tables; false: the code is not an entity of the dataset
__device__ __forceinline__ bool autol_entity(const Scene& s, const AutoLabelScene& a, unsigned code, unsigned& c, unsigned& slot) {
    const unsigned e = code & ~(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE);
    if (code & OSMR_AREA_MULTIPOLYGON) {
        if (e >= s.n_mps) return false;
        c = a.mp_class[e];
        slot = a.mp_base + e;
    } else if (code & OSMR_LABEL_NODE) {
        if (e >= s.n_nodes) return false;
        c = a.node_class[e];
        slot = e;
    } else {
        if (e >= s.n_ways) return false;
        c = a.way_class[e];
        slot = a.way_base + e;
    }
    return true;
}

template <bool kGather>
__device__ __forceinline__ void autol_visit_tile(const Scene& s, const AutoLabelScene& a, unsigned t, unsigned* sh_counts) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const AutoRect rc = auto_rect(s.tiles[t]);
    AutoScene ix{};  // (auto_lower_bound reads the index through an AutoScene)
    ix.idx_xy = a.idx_xy;
    ix.n_idx = a.n_idx;
    unsigned local_bound = 0;
    for (long long col = rc.xa + warp; col <= rc.xb; col += kAutoWarps) {
        unsigned i = auto_lower_bound(ix, col, rc.ya);
        if (i >= a.n_idx) break;
        const uint2 first = a.idx_xy[i];
        if ((long long)first.x != col) {
            if ((long long)first.x > rc.xb) break;
            const long long nxt = (long long)first.x;
            const long long k = (nxt - rc.xa - warp + kAutoWarps - 1) / kAutoWarps;
            col = rc.xa + warp + (k - 1) * kAutoWarps;
            continue;
        }
        for (; i < a.n_idx; ++i) {
            const uint2 xy = a.idx_xy[i];
            if ((long long)xy.x != col || (long long)xy.y > rc.yb) break;
            const uint2 nl = a.idx_n[i], wl = a.idx_w[i], ml = a.idx_m[i];
            if (!kGather) {
                local_bound += nl.y + wl.y + ml.y;
                continue;
            }
            const unsigned total = nl.y + wl.y + ml.y;
            for (unsigned j = lane; j < total; j += 32) {
                const bool is_node = j < nl.y, is_mp = j >= nl.y + wl.y;
                const unsigned e = is_node ? s.ints[nl.x + j] : (is_mp ? s.ints[ml.x + (j - nl.y - wl.y)] : s.ints[wl.x + (j - nl.y)]);
                const unsigned code = e | (is_node ? OSMR_LABEL_NODE : (is_mp ? OSMR_AREA_MULTIPOLYGON : 0u));
                unsigned c, slot;
                if ((e & (OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE)) || !autol_entity(s, a, code, c, slot)) {
                    atomicOr(&s.counters[CNT_BAD_INPUT], 2u);
                    continue;
                }
                if (!is_node) {
                    const uint2 mt = is_mp ? a.mp_min_tile[e] : a.way_min_tile[e];
                    if ((long long)xy.x != max((long long)mt.x, rc.xa) || (long long)xy.y != max((long long)mt.y, rc.ya)) continue;  // not the owner
                    if (is_mp && s.mps[e].y == 0) continue;  // reader.rs:86-93
                }
                if (c >= a.n_classes) continue;  // no style list
                const unsigned n_live = autol_live_count(s, a, c, slot);
                if (n_live == 0) continue;
                const unsigned pos = atomicAdd(&sh_counts[0], 1u);
                atomicAdd(&sh_counts[1], n_live);
                a.cand[a.bound[t] + pos] = code;
            }
        }
    }
    if (!kGather) {
        for (int o = 16; o > 0; o >>= 1) local_bound += __shfl_xor_sync(0xffffffffu, local_bound, o);
        if (lane == 0 && local_bound) atomicAdd(&sh_counts[0], local_bound);
    }
}

__global__ void __launch_bounds__(kAutoThreads) autol_bound_kernel(Scene s, AutoLabelScene a) {
    __shared__ unsigned counts[2];
    if (threadIdx.x < 2) counts[threadIdx.x] = 0;
    __syncthreads();
    autol_visit_tile<false>(s, a, blockIdx.x, counts);
    __syncthreads();
    if (threadIdx.x == 0) a.bound[blockIdx.x] = counts[0];
}

__global__ void __launch_bounds__(kAutoThreads) autol_gather_kernel(Scene s, AutoLabelScene a) {
    __shared__ unsigned counts[2];
    if (threadIdx.x < 2) counts[threadIdx.x] = 0;
    __syncthreads();
    autol_visit_tile<true>(s, a, blockIdx.x, counts);
    __syncthreads();
    if (threadIdx.x == 0) {
        a.cand_cnt[blockIdx.x] = counts[0];
        a.inst_cnt[blockIdx.x] = counts[1];
    }
}

__global__ void __launch_bounds__(kAutoThreads) autol_sort_kernel(Scene s, AutoLabelScene a) {
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
        if (n_pad > kAutoSortCap) {  // sort in global scratch (handed out through the same counter pair as auto_sort_kernel's)
            unsigned long long* used = reinterpret_cast<unsigned long long*>(&s.counters[CNT_WALK_ALPHA]);
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
    const unsigned* cand = a.cand + a.bound[t];
    const unsigned n_cand = a.cand_cnt[t];
    for (unsigned ci = threadIdx.x; ci < n_cand; ci += blockDim.x) {
        const unsigned code = cand[ci];
        const unsigned e = code & ~(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE);
        unsigned c = 0, slot = 0;
        autol_entity(s, a, code, c, slot);  // (checked by the gather)
        const bool is_node = (code & OSMR_LABEL_NODE) != 0 && !(code & OSMR_AREA_MULTIPOLYGON);
        const unsigned long long rank = is_node ? a.node_rank[e] : ((code & OSMR_AREA_MULTIPOLYGON) ? a.mp_rank[e] : a.way_rank[e]);
        const unsigned b = a.class_begin[c], cnt = a.class_begin[c + 1] - b;
        unsigned at = atomicAdd(&cursor, autol_live_count(s, a, c, slot));
        for (unsigned w = 0; w < cnt; ++w) {
            const osmr_class_style cs = a.class_styles[b + w];
            if (!autol_live(s, a, cs.style, slot)) continue;
            const unsigned long long order = cs.order | (is_node ? kAutoLabelNodeOrder : 0u);
            keys[at++] = (order << (32 + kAutoWithinBits)) | (rank << kAutoWithinBits) | w;
        }
    }
    __syncthreads();
    bitonic_sort(keys, n_pad);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = keys[i];
        const bool is_node = ((k >> (32 + kAutoWithinBits)) & kAutoLabelNodeOrder) != 0;
        const unsigned rank = (unsigned)((k >> kAutoWithinBits) & 0xffffffffull);
        const unsigned code = is_node ? (a.node_rank_entity[rank] | OSMR_LABEL_NODE) : a.rank_entity[rank];
        unsigned c = 0, slot = 0;
        autol_entity(s, a, code, c, slot);
        osmr_label L;
        L.entity = code;
        L.style = a.class_styles[a.class_begin[c] + (unsigned)(k & ((1u << kAutoWithinBits) - 1u))].style;
        a.labels_out[out0 + i] = L;
    }
}

}  // namespace osmr

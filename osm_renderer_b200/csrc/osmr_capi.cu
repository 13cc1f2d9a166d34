// osmr_capi.cu -- C ABI (include/osmr.h) of libosmr_b200.so: context, dataset residency, batch launch.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo ...  (osm_renderer_b200/build.py)
// There is deliberately no CPU code path for any part of the draw path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <stdexcept>
#include <system_error>
#include <string>
#include <vector>

#include "osmr.h"
#include "osmr_kernels.cuh"
#include "osmr_auto.cuh"
#include "osmr_png.cuh"
#include "osmr_labels_host.hpp"
#include "osmr_labels_dev.cuh"
#include "osmr_auto_labels.cuh"
#include <map>
#include <memory>
#include <unordered_map>

using namespace osmr;

namespace {

template <typename T>
struct DevBuf {  // owns its allocation: error paths that return early (CK) do not leak device memory
    T* p = nullptr;
    size_t cap = 0;  // elements
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct PinnedBuf {  // page-locked host staging, grown on demand and kept
    T* p = nullptr;
    size_t cap = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr unsigned kMaxChunks = 16;  // draw chunks of one call (host output is pipelined chunk by chunk)

struct LabelWorkItem {  // a run of consecutive labels of one tile, laid out by one host thread
    uint32_t tile, first, count;
    std::vector<osmr_host::LabelRec> recs;
    std::vector<osmr_host::Seg> segs;
    size_t seg_base;
};

}  // namespace

struct Dataset {
    int device = 0;
    bool has_geo = false;
    unsigned n_nodes = 0, n_ways = 0, n_polys = 0, n_mps = 0, n_ints = 0;
    DevBuf<double2> merc;
    DevBuf<EntBox> way_box, mp_box;
    DevBuf<uint2> ways, polys, mps;
    DevBuf<unsigned> ints;
    std::vector<unsigned> h_way_len, h_mp_pts;  // node counts per entity (host copy, for scratch sizing)
    std::vector<uint8_t> h_bin;                 // host copy of the geodata image: tags and coordinates for the label tables
    osmr_host::BinView h_view;
    std::vector<double2> h_merc;                // host copy of the Mercator factors (direction tables of the label layout)
    // f3: the tile index and what the device-side lookup derives from it
    std::string auto_unavailable;  // why the tile index of the image cannot drive the device-side lookup ("" = it can)
    unsigned n_idx = 0;
    DevBuf<uint2> idx_xy, idx_w, idx_m, way_min_tile, mp_min_tile;
    DevBuf<unsigned> way_rank, mp_rank, rank_entity;
    // f3 for the label pass (osmr_auto_labels.cuh): node lists of the index records, rank of (global id, local id) among the nodes
    std::string auto_labels_unavailable;
    DevBuf<uint2> idx_n;
    DevBuf<unsigned> node_rank, node_rank_entity;
    ~Dataset() { cudaSetDevice(device); }  // (the buffers free themselves right after, on this device)
};

struct osmr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t chunk_done[kMaxChunks] = {};
    cudaEvent_t cev[kMaxChunks][5] = {};   // per draw chunk: start, after cover, after raster, before cover, raster start (after the wait for the labels)
    unsigned chunk_launches[kMaxChunks] = {};
    unsigned last_draw_chunks = 0;  // draw chunks of the last draw (diagnostics)
    PinnedBuf<unsigned> h_cnt;             // per draw chunk: its counters, copied back asynchronously
    cudaEvent_t areas_ready = nullptr;
    bool areas_deferred = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    int num_sms = 0;

    // dataset: immutable once osmr_set_geodata has returned, shared read-only between the contexts made by
    // osmr_ctx_create_shared (reference: one GeodataReader behind an Arc for all worker threads, http_server.rs:42-48)
    std::shared_ptr<Dataset> ds;
    // label pass (host half: osmr_labels_host.hpp)
    osmr_host::TrueType font;
    std::vector<osmr_host::LabelStyleHost> label_styles;
    std::vector<osmr_host::IconDim> label_icon_dims;
    DevBuf<DevIcon> label_icons;
    DevBuf<double4> label_icon_px;
    DevBuf<DevLabel> d_labels;
    DevBuf<DevSeg> d_label_segs;
    DevBuf<unsigned> d_label_begin, label_occ;
    DevBuf<double> label_acc;
    DevBuf<int> label_row_keys;
    DevBuf<unsigned> d_cover_list, d_cover_cursor;
    std::vector<LabelWorkItem> label_items;  // kept between calls: their vectors' capacity is the layout arena
    PinnedBuf<osmr_host::Seg> h_label_segs;
    cudaEvent_t ev_label0 = nullptr, ev_label1 = nullptr;
    // the label pass runs chunk by chunk (the draw's host-output schedule): a draw chunk waits only for the label chunks under it
    unsigned n_lchunks = 0;
    unsigned lchunk_tb[kMaxChunks] = {}, lchunk_tc[kMaxChunks] = {};
    cudaEvent_t lchunk_done[kMaxChunks] = {}, lchunk_up[kMaxChunks] = {}, ev_lcov0[kMaxChunks] = {}, ev_lcov1[kMaxChunks] = {};
    float stats_label_layout_ms = 0.f, stats_label_device_ms = 0.f;
    unsigned label_threads = 32;
    DevBuf<LabelPix> label_plane;
    DevBuf<unsigned> label_pmask;  // per tile D*D bits: the pending pixels of label_plane
    bool label_plane_active = false;
    // label layout on the device (osmr_labels_dev.cuh): resident tables + per-call scratch
    struct LabelResident {
        bool valid = false;
        std::vector<std::string> keys;
        DevBuf<DevLabelStyle> styles;
        DevBuf<unsigned> text_id, text_begin, glyph_vbegin;
        DevBuf<DevGlyphRec> glyphs;
        DevBuf<DevVertex> verts;
        unsigned n_keys = 0, ent_total = 0, way_base = 0, mp_base = 0, n_texts = 0;
        int font_reach = 0;  // font units: how far an outline point can lie from its glyph's pen position (label_precull_kernel)
        std::vector<uint8_t> way_has_text;  // ways whose text may run along the line: they get direction tables
        struct Angle {
            bool set = false;
            int scale = 0;
            DevBuf<unsigned> off;
            DevBuf<double2> sc;
        } angle[19];
    } lres;
    DevBuf<osmr_label> d_label_list;
    DevBuf<ActLabel> l_act;
    DevBuf<unsigned char> l_predead;  // label_precull_kernel's verdict per active label slot
    DevBuf<unsigned> l_act_cnt, l_counters;
    DevBuf<LabelPlace> l_place;
    DevBuf<GlyphPlace> l_gplace;
    DevBuf<unsigned> l_place_vinst, l_vinst_place, l_vcnt, l_curve_list, l_scan_blocks;
    DevBuf<double4> l_vbox;
    DevBuf<unsigned long long> l_curve_codes;
    DevBuf<unsigned char> l_curve_deep;
    size_t l_curves_cap = 0;
    // the capacities the last attempt ran with
    size_t l_places_used_cap = 0, l_segs_used_cap = 0, l_rowrecs_used_cap = 0, l_cells_used_cap = 0, l_ring_used_cap = 0, l_heap_used_slots = 0,
           l_verts_used_cap = 0, l_curves_used_cap = 0;
    DevBuf<CurveRoot> l_curve_root;
    // the second set of the label pass's per-chunk scratch: odd label chunks run on their own stream beside the even ones
    struct LabelScratchB {
        DevBuf<DevSeg> segs;
        DevBuf<double> acc;
        DevBuf<int> row_keys;
        DevBuf<unsigned> cover_list, place_vinst, vinst_place, vcnt, curve_list, scan_blocks;
        DevBuf<GlyphPlace> gplace;
        DevBuf<double4> vbox;
        DevBuf<unsigned long long> curve_codes;
        DevBuf<unsigned char> curve_deep, heap;
        DevBuf<CurveRoot> curve_root;
        DevBuf<double2> ring_pts;
    } lscrB;
    cudaStream_t label_stream2 = nullptr;
    cudaEvent_t label_join = nullptr, label_prep = nullptr;
    size_t l_verts_cap = 0;
    DevBuf<double2> l_ring_pts;
    DevBuf<unsigned char> l_heap;
    PinnedBuf<unsigned> h_lcnt;
    size_t l_places_cap = 0, l_segs_cap = 0, l_rowrecs_cap = 0, l_cells_cap = 0, l_ring_cap = 0, l_heap_slots = 0;
    std::vector<osmr_tile> h_batch_tiles;     // resident labelled batch: host copies for the table look-ups of the label pass
    std::vector<uint32_t> h_label_begin;
    bool batch_has_labels = false;
    bool resident_needs_host_layout = false;  // osmr_batch_draw_labeled's last verdict (osmr_draw_tiles_auto_labeled falls back on it)
    bool label_serial = false;     // debug key "label_serial": the label pass finishes before the area passes start (measurement only)
    bool label_cull = true;        // debug key "label_cull": labels that cannot reach the tile get no outlines and no coverage
    unsigned label_chunks = 0;     // debug key "label_chunks"
    unsigned curve_leaf_cap = 0;   // debug key "curve_leaf_cap" (tests: curves with more leaves are flattened again by one lane)
    bool label_host_only = false;  // debug key "label_host": always lay labels out on the host (round-1 path)
    unsigned stats_label_active = 0, stats_label_poly = 0, stats_label_segs = 0;
    unsigned long long stats_label_cells = 0;
    // styles / icons
    unsigned n_styles = 0, n_dashes = 0, n_icons = 0;
    DevBuf<osmr_style> styles;
    DevBuf<double> dashes;
    DevBuf<DevIcon> icons;
    DevBuf<double4> icon_px;
    // batch
    bool has_batch = false;
    unsigned n_tiles = 0, n_areas = 0;
    int scale = 1;
    DevBuf<osmr_tile> tiles;
    DevBuf<unsigned> area_begin;
    DevBuf<osmr_styled_area> areas;
    std::vector<uint32_t> h_area_begin;
    // scratch
    DevBuf<VisOp> vis;
    DevBuf<RasterOp> rop;
    DevBuf<short4> vis_bbox;
    DevBuf<unsigned> vis_count, work, counters, mask;
    DevBuf<uint2> fill_work, line_work;
    DevBuf<uint4> geom, calc_table;
    DevBuf<double> walk_alpha;        // fragment alphas (line_cover_kernel -> raster_kernel)
    DevBuf<unsigned char> walk_len;   // fragment pixel indices
    DevBuf<BinEntry> bin_entries;     // (block, op) pairs in block order (bin_ops_kernel)
    DevBuf<unsigned> frag_cnt, pair_table;
    DevBuf<uint2> blk_range;          // per (tile, block) of the whole batch
    // Second set of the bump-allocated scratch: the draw chunks of a host-output call alternate between two streams
    // (chunk i+1 is planned while chunk i is still being rastered), so consecutive chunks must not share scratch.
    struct ScratchB {
        DevBuf<uint4> geom;
        DevBuf<unsigned> mask;
        DevBuf<uint2> fill_work, line_work;
        DevBuf<double> walk_alpha;
        DevBuf<unsigned char> walk_len;
        DevBuf<BinEntry> bin_entries;
        DevBuf<unsigned> frag_cnt, pair_table, plan_slice_base, bin_cnt_ent;
        DevBuf<unsigned long long> bin_cnt_cap;
        void release() {
            plan_slice_base.release();
            bin_cnt_ent.release();
            bin_cnt_cap.release();
            bin_entries.release();
            frag_cnt.release();
            pair_table.release();
            geom.release();
            mask.release();
            fill_work.release();
            line_work.release();
            walk_alpha.release();
            walk_len.release();
        }
    } scrB;
    cudaStream_t stream2 = nullptr;
    cudaStream_t label_stream = nullptr;  // the label pass runs beside the area passes; raster_kernel waits for label_done
    cudaEvent_t label_done = nullptr, label_go = nullptr;
    bool label_async = false;             // the label plane of this draw is produced on label_stream
    const osmr_styled_area* tail_src = nullptr;  // styled areas behind the first draw chunk, not yet on their way (flush_area_tail)
    osmr_styled_area* tail_dst = nullptr;
    size_t tail_bytes = 0;
    cudaEvent_t ev_wall0 = nullptr, ev_wall1 = nullptr, join2 = nullptr;  // device wall time of a draw across both streams
    cudaEvent_t prep_done = nullptr;  // upload + style calculators on `stream`: what stream2's first chunk waits for
    bool two_streams = true;          // debug key "two_streams"
    DevBuf<unsigned> plan_slice_base;  // per (tile, pass, slice) of a chunk: plan_count_kernel's counts, then their scan
    DevBuf<unsigned> bin_cnt_ent;      // per (block area, slice, block) of a chunk: bin_count_kernel's counts, then bin_scan_kernel's offsets
    DevBuf<unsigned long long> bin_cnt_cap;
    unsigned plan_slice_areas = 0;     // debug key "plan_slice_areas": styled areas per plan slice (0: the default, 16384)
    size_t geom_cap_units = 0, mask_cap_words = 0, walk_alpha_cap = 0, walk_len_cap = 0, entries_cap = 0, pair_cap = 0;
    DevBuf<unsigned char> out;
    size_t out_bytes = 0;
    // f3: device-side candidate lookup + painter's order (osmr_auto.cuh)
    struct ZoomTable {
        DevBuf<unsigned> way_class, mp_class, class_begin;
        DevBuf<osmr_class_style> class_styles;
        DevBuf<int> class_reach;
        unsigned n_classes = 0, n_class_styles = 0;
        bool set = false;
        // label classes (osmr_set_zoom_label_styles)
        DevBuf<unsigned> l_node_class, l_way_class, l_mp_class, l_class_begin;
        DevBuf<osmr_class_style> l_class_styles;
        unsigned l_n_classes = 0;
        bool l_set = false;
    };
    ZoomTable zoom_tables[19];
    DevBuf<unsigned> auto_bound, auto_cand, auto_cand_cnt, auto_inst;
    DevBuf<unsigned long long> auto_big_keys;
    // f4: PNG encode on the device (osmr_png.cuh)
    DevBuf<unsigned> png_band_words, png_band_bytes, png_band_len, png_tile_off;
    DevBuf<uint2> png_band_adler;
    DevBuf<unsigned char> png_out;
    int fill_cap = kFillCap;
    bool direct_out = false;  // debug key "direct_out": raster_kernel stores the tiles straight into a page-locked `out`
                              // (no D2H stage; measured slower than the staged pipeline: PCIe-bound stores, 26 GB/s)
    bool device_merc = false;  // debug key "device_merc": Mercator factors by project_nodes_kernel instead of the host libm
    unsigned work_items_limit = 0;  // debug key "work_items": pretend the work lists are this short once (exercises their growth)
    unsigned host_chunks = 0;  // 0: tapered default schedule (plan_chunks)
    unsigned resident_chunks = 1;  // debug key "resident_chunks": draw chunks when the output stays in HBM
    unsigned first_chunk = 0;  // tiles in the first draw chunk of the current upload (0: one chunk)
    osmr_stats stats{};

    int fail(int code, const char* what, cudaError_t e = cudaSuccess) noexcept {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        try {
            err = buf;
        } catch (...) {  // the message itself could not be allocated: keep whatever fits without allocating
            err.clear();
        }
        return code;
    }
};

// No exception may cross the C boundary (include/osmr.h): every exported function is a function-try-block ending in one of
// these handlers.  std::bad_alloc (host vectors / strings / thread stacks) -> OSMR_E_NOMEM, std::system_error (thread creation)
// and anything else -> OSMR_E_INVALID with the exception's message as osmr_last_error.
static int osmr_exception_to_code(osmr_ctx* ctx, int code, const char* what) noexcept {
    if (ctx) return ctx->fail(code, what);
    return code;
}
#define OSMR_CATCH_INT(ctxp)                                                                                         \
    catch (const std::bad_alloc&) { return osmr_exception_to_code((ctxp), OSMR_E_NOMEM, "out of host memory"); }       \
    catch (const std::exception& ex_) { return osmr_exception_to_code((ctxp), OSMR_E_INVALID, ex_.what()); }           \
    catch (...) { return osmr_exception_to_code((ctxp), OSMR_E_INVALID, "unknown C++ exception"); }
#define OSMR_CATCH_VAL(v) \
    catch (...) { return (v); }

#define CK(call)                                                       \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return ctx->fail(OSMR_E_CUDA, #call, e_); \
    } while (0)

static inline uint32_t rd_u32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}

extern "C" {

uint32_t osmr_abi_version(void) { return 5; }

int osmr_ctx_create(int device, osmr_ctx** out_ctx) try {
    if (!out_ctx) return OSMR_E_INVALID;
    *out_ctx = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return OSMR_E_CUDA;
    osmr_ctx* ctx = new (std::nothrow) osmr_ctx();
    if (!ctx) return OSMR_E_NOMEM;
    ctx->device = device;
    ctx->ds = std::make_shared<Dataset>();
    ctx->ds->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        // The label pass is the longer of the two and has the latency-bound kernels (one CTA per tile, serial inside): its CTAs go
        // first when both kinds of stream have work pending, the area kernels fill what is left (C2 batch: 13.6 -> 13.2 ms per
        // labelled step).  OSMR_LABEL_PRIORITY=0 in the environment keeps the default priority (A/B).  Six streams per context:
        // the device maps streams onto CUDA_DEVICE_MAX_CONNECTIONS (default 8) hardware queues, and two streams that share a queue
        // wait for each other's launches -- measured: with ten streams a draw chunk started 9 ms late behind the label pass.
        int lo = 0, hi = 0;
        e = cudaDeviceGetStreamPriorityRange(&lo, &hi);  // (hi is the numerically smallest = most urgent)
        const char* lp = getenv("OSMR_LABEL_PRIORITY");
        const int prio = (lp && lp[0] == '0') ? lo : hi;
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->label_stream, cudaStreamNonBlocking, prio);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->label_stream2, cudaStreamNonBlocking, prio);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->label_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->label_prep, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->label_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->label_go, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->prep_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->join2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_wall0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_wall1);
    for (unsigned i = 0; i < kMaxChunks && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->chunk_done[i], cudaEventDisableTiming);
    for (unsigned i = 0; i < kMaxChunks; ++i)
        for (int j = 0; j < 5 && e == cudaSuccess; ++j) e = cudaEventCreate(&ctx->cev[i][j]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->areas_ready, cudaEventDisableTiming);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_label0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_label1);
    for (unsigned i = 0; i < kMaxChunks && e == cudaSuccess; ++i) {
        e = cudaEventCreate(&ctx->ev_lcov0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_lcov1[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lchunk_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lchunk_up[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) {
        double lut[256];
        for (int i = 0; i < 256; ++i) {
            volatile double num = (double)i, den = 255.0;
            lut[i] = num / den;  // IEEE division, identical to the device's `(double)c / 255.0`
        }
        e = cudaMemcpyToSymbol(osmr::kUnitOfU8, lut, sizeof lut);
    }
    if (e != cudaSuccess) {
        osmr_ctx_destroy(ctx);
        return OSMR_E_CUDA;
    }
    *out_ctx = ctx;
    return OSMR_OK;
} OSMR_CATCH_INT(nullptr)

void osmr_ctx_destroy(osmr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->styles.release();
    ctx->dashes.release();
    ctx->icons.release();
    ctx->icon_px.release();
    ctx->tiles.release();
    ctx->area_begin.release();
    ctx->areas.release();
    ctx->vis.release();
    ctx->rop.release();
    ctx->vis_bbox.release();
    ctx->vis_count.release();
    ctx->work.release();
    ctx->fill_work.release();
    for (auto& z : ctx->zoom_tables) {
        z.way_class.release();
        z.mp_class.release();
        z.class_begin.release();
        z.class_styles.release();
        z.class_reach.release();
    }
    ctx->auto_bound.release();
    ctx->auto_cand.release();
    ctx->auto_cand_cnt.release();
    ctx->auto_inst.release();
    ctx->auto_big_keys.release();
    ctx->png_band_words.release();
    ctx->png_band_bytes.release();
    ctx->png_band_len.release();
    ctx->png_tile_off.release();
    ctx->png_band_adler.release();
    ctx->png_out.release();
    ctx->line_work.release();
    ctx->walk_alpha.release();
    ctx->walk_len.release();
    ctx->counters.release();
    ctx->mask.release();
    ctx->label_icons.release();
    ctx->label_icon_px.release();
    ctx->d_labels.release();
    ctx->d_label_segs.release();
    ctx->d_label_begin.release();
    ctx->label_occ.release();
    ctx->label_acc.release();
    ctx->label_row_keys.release();
    ctx->h_label_segs.release();
    ctx->label_plane.release();
    ctx->label_pmask.release();
    ctx->geom.release();
    ctx->scrB.release();
    if (ctx->prep_done) cudaEventDestroy(ctx->prep_done);
    if (ctx->join2) cudaEventDestroy(ctx->join2);
    if (ctx->ev_wall0) cudaEventDestroy(ctx->ev_wall0);
    if (ctx->ev_wall1) cudaEventDestroy(ctx->ev_wall1);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->label_stream) cudaStreamDestroy(ctx->label_stream);
    if (ctx->label_stream2) cudaStreamDestroy(ctx->label_stream2);
    if (ctx->label_join) cudaEventDestroy(ctx->label_join);
    if (ctx->label_prep) cudaEventDestroy(ctx->label_prep);
    if (ctx->label_done) cudaEventDestroy(ctx->label_done);
    if (ctx->label_go) cudaEventDestroy(ctx->label_go);
    ctx->calc_table.release();
    ctx->out.release();
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->chunk_done)
        if (e) cudaEventDestroy(e);
    for (auto& row : ctx->cev)
        for (auto& e : row)
            if (e) cudaEventDestroy(e);
    ctx->h_cnt.release();
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->areas_ready) cudaEventDestroy(ctx->areas_ready);
    if (ctx->ev_label0) cudaEventDestroy(ctx->ev_label0);
    if (ctx->ev_label1) cudaEventDestroy(ctx->ev_label1);
    for (unsigned i = 0; i < kMaxChunks; ++i) {
        if (ctx->ev_lcov0[i]) cudaEventDestroy(ctx->ev_lcov0[i]);
        if (ctx->ev_lcov1[i]) cudaEventDestroy(ctx->ev_lcov1[i]);
        if (ctx->lchunk_done[i]) cudaEventDestroy(ctx->lchunk_done[i]);
        if (ctx->lchunk_up[i]) cudaEventDestroy(ctx->lchunk_up[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// A worker context over the parent's resident dataset (reference: src/http_server.rs:42-48,69-72 -- reader, styler and drawer are
// shared immutably by all worker threads, each thread owns only its TilePixels).  The new context has its own streams, scratch,
// style / icon / label tables and zoom classes; the geodata (Mercator factors, entity tables and boxes, tile index, the host
// copy behind the label tables) is the parent's, by reference.  The dataset is immutable: osmr_set_geodata on either context
// gives THAT context a dataset of its own and leaves the other one untouched.  Contexts may be used from different host threads
// (one thread per context, as before).
int osmr_ctx_create_shared(const osmr_ctx* parent, osmr_ctx** out_ctx) try {
    if (!out_ctx) return OSMR_E_INVALID;
    *out_ctx = nullptr;
    if (!parent) return OSMR_E_INVALID;
    osmr_ctx* ctx = nullptr;
    int rc = osmr_ctx_create(parent->device, &ctx);
    if (rc) return rc;
    ctx->ds = parent->ds;
    *out_ctx = ctx;
    return OSMR_OK;
} OSMR_CATCH_INT(nullptr)

const char* osmr_last_error(const osmr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int osmr_debug_set(osmr_ctx* ctx, const char* key, int value) try {
    if (!ctx || !key) return OSMR_E_INVALID;
    if (strcmp(key, "label_threads") == 0) {  // host threads of the label layout (1: serial)
        if (value < 1 || value > 256) return ctx->fail(OSMR_E_INVALID, "label_threads must be 1..256");
        ctx->label_threads = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "label_host") == 0) {  // 1: label layout on the host for every call (the round-1 path; A/B and tests)
        ctx->label_host_only = value != 0;
        return OSMR_OK;
    }
    if (strcmp(key, "curve_leaf_cap") == 0) {  // tests: leaf codes per curve (0: the build's 128)
        if (value < 0) return ctx->fail(OSMR_E_INVALID, "curve_leaf_cap must be >= 0");
        ctx->curve_leaf_cap = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "label_chunks") == 0) {  // n > 0: the label pass in n equal chunks whatever the batch (tests); 0: the draw's schedule
        if (value < 0 || value > (int)kMaxChunks) return ctx->fail(OSMR_E_INVALID, "label_chunks must be 0..16");
        ctx->label_chunks = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "direct_out") == 0) {  // 0: always stage the tiles in HBM and copy them back (A/B measurements)
        ctx->direct_out = value != 0;
        return OSMR_OK;
    }
    if (strcmp(key, "host_chunks") == 0) {  // equal draw chunks of a staged host-output call (0: tapered default)
        if (value < 0 || value > (int)kMaxChunks) return ctx->fail(OSMR_E_INVALID, "host_chunks must be 0..16");
        ctx->host_chunks = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "device_merc") == 0) {  // takes effect at the next osmr_set_geodata
        ctx->device_merc = value != 0;
        return OSMR_OK;
    }
    if (strcmp(key, "resident_chunks") == 0) {
        if (value < 1 || value > (int)kMaxChunks) return ctx->fail(OSMR_E_INVALID, "resident_chunks must be 1..16");
        ctx->resident_chunks = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "two_streams") == 0) {  // 0: all draw chunks of a host-output call on one stream (A/B measurements)
        ctx->two_streams = value != 0;
        return OSMR_OK;
    }
    if (strcmp(key, "work_items") == 0) {
        if (value < 1) return ctx->fail(OSMR_E_INVALID, "work_items must be positive");
        ctx->work_items_limit = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "label_serial") == 0) {
        ctx->label_serial = value != 0;
        return OSMR_OK;
    }
    if (strcmp(key, "label_cull") == 0) {
        ctx->label_cull = value != 0;
        return OSMR_OK;
    }
    if (strcmp(key, "plan_slice_areas") == 0) {
        if (value < 0) return ctx->fail(OSMR_E_INVALID, "plan_slice_areas must not be negative");
        ctx->plan_slice_areas = (unsigned)value;
        return OSMR_OK;
    }
    if (strcmp(key, "fill_cap") == 0) {
        if (value < 0 || value > kFillCap) return ctx->fail(OSMR_E_INVALID, "fill_cap out of range");
        ctx->fill_cap = value;
        return OSMR_OK;
    }
    if (strcmp(key, "scratch_units") == 0) {  // (re)start with a tiny geometry / mask scratch to exercise the grow-and-redo path
        if (value < 1) return ctx->fail(OSMR_E_INVALID, "scratch_units must be positive");
        ctx->geom.release();
        ctx->mask.release();
        ctx->walk_alpha.release();
        ctx->walk_len.release();
        ctx->scrB.release();
        cudaError_t e1 = ctx->geom.reserve((size_t)value);
        cudaError_t e2 = ctx->mask.reserve((size_t)value);
        if (e1 == cudaSuccess) e1 = ctx->walk_alpha.reserve((size_t)value);
        if (e2 == cudaSuccess) e2 = ctx->walk_len.reserve((size_t)value);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return ctx->fail(OSMR_E_CUDA, "scratch_units", e1 != cudaSuccess ? e1 : e2);
        ctx->geom_cap_units = (size_t)value;
        ctx->mask_cap_words = (size_t)value;
        ctx->walk_alpha_cap = (size_t)value;
        ctx->walk_len_cap = (size_t)value;
        ctx->bin_entries.release();
        ctx->frag_cnt.release();
        ctx->pair_table.release();
        if (ctx->bin_entries.reserve((size_t)value) != cudaSuccess || ctx->frag_cnt.reserve((size_t)value) != cudaSuccess ||
            ctx->pair_table.reserve((size_t)value) != cudaSuccess)
            return ctx->fail(OSMR_E_CUDA, "scratch_units");
        ctx->entries_cap = (size_t)value;
        ctx->pair_cap = (size_t)value;
        return OSMR_OK;
    }
    return ctx->fail(OSMR_E_INVALID, "unknown debug key");
} OSMR_CATCH_INT(ctx)

// ---------------------------------------------------------------------------------------------------------
int osmr_set_geodata(osmr_ctx* ctx, const void* bin, size_t len) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!bin) return ctx->fail(OSMR_E_INVALID, "null geodata image");
    cudaSetDevice(ctx->device);
    const uint8_t* p = (const uint8_t*)bin;
    size_t pos = 0;
    const uint8_t* base[6];
    uint32_t cnt[6];
    const size_t rec[6] = {32, 24, 8, 24, 32, 4};  // nodes, ways, polygons, multipolygons, tiles, ints (reader.rs:301-305)
    for (int i = 0; i < 6; ++i) {
        if (pos + 4 > len) return ctx->fail(OSMR_E_INVALID, "geodata image truncated");
        cnt[i] = rd_u32(p + pos);
        pos += 4;
        if ((size_t)cnt[i] * rec[i] > len - pos) return ctx->fail(OSMR_E_INVALID, "geodata image truncated");
        base[i] = p + pos;
        pos += (size_t)cnt[i] * rec[i];
    }
    const uint32_t n_nodes = cnt[0], n_ways = cnt[1], n_polys = cnt[2], n_mps = cnt[3], n_ints = cnt[5];
    // a fresh dataset object: contexts that share the previous one (osmr_ctx_create_shared) keep it
    {
        std::shared_ptr<Dataset> fresh = std::make_shared<Dataset>();
        fresh->device = ctx->device;
        ctx->ds.swap(fresh);
    }
    std::vector<uint32_t> ints(n_ints);
    if (n_ints) memcpy(ints.data(), base[5], (size_t)n_ints * 4);
    std::vector<uint2> ways(n_ways), polys(n_polys), mps(n_mps);
    ctx->ds->h_way_len.assign(n_ways, 0);
    ctx->ds->h_mp_pts.assign(n_mps, 0);
    auto range_ok = [&](uint32_t off, uint32_t l) { return (uint64_t)off + l <= n_ints; };
    for (uint32_t i = 0; i < n_ways; ++i) {
        ways[i] = make_uint2(rd_u32(base[1] + (size_t)i * 24 + 8), rd_u32(base[1] + (size_t)i * 24 + 12));
        if (!range_ok(ways[i].x, ways[i].y)) return ctx->fail(OSMR_E_INVALID, "way node list out of range");
        for (uint32_t k = 0; k < ways[i].y; ++k)
            if (ints[ways[i].x + k] >= n_nodes) return ctx->fail(OSMR_E_INVALID, "way references a missing node");
        ctx->ds->h_way_len[i] = ways[i].y;
    }
    for (uint32_t i = 0; i < n_polys; ++i) {
        polys[i] = make_uint2(rd_u32(base[2] + (size_t)i * 8), rd_u32(base[2] + (size_t)i * 8 + 4));
        if (!range_ok(polys[i].x, polys[i].y)) return ctx->fail(OSMR_E_INVALID, "polygon node list out of range");
        for (uint32_t k = 0; k < polys[i].y; ++k)
            if (ints[polys[i].x + k] >= n_nodes) return ctx->fail(OSMR_E_INVALID, "polygon references a missing node");
    }
    for (uint32_t i = 0; i < n_mps; ++i) {
        mps[i] = make_uint2(rd_u32(base[3] + (size_t)i * 24 + 8), rd_u32(base[3] + (size_t)i * 24 + 12));
        if (!range_ok(mps[i].x, mps[i].y)) return ctx->fail(OSMR_E_INVALID, "multipolygon polygon list out of range");
        for (uint32_t k = 0; k < mps[i].y; ++k) {
            uint32_t pid = ints[mps[i].x + k];
            if (pid >= n_polys) return ctx->fail(OSMR_E_INVALID, "multipolygon references a missing polygon");
            ctx->ds->h_mp_pts[i] += polys[pid].y;
        }
    }
    ctx->ds->has_geo = false;
    DevBuf<unsigned char> raw_nodes;
    CK(raw_nodes.reserve((size_t)n_nodes * 32 + 16));
    CK(ctx->ds->merc.reserve(n_nodes + 1));
    CK(ctx->ds->ways.reserve(n_ways + 1));
    CK(ctx->ds->polys.reserve(n_polys + 1));
    CK(ctx->ds->mps.reserve(n_mps + 1));
    CK(ctx->ds->ints.reserve(n_ints + 1));
    if (n_nodes) CK(cudaMemcpyAsync(raw_nodes.p, base[0], (size_t)n_nodes * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (n_ways) CK(cudaMemcpyAsync(ctx->ds->ways.p, ways.data(), (size_t)n_ways * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n_polys) CK(cudaMemcpyAsync(ctx->ds->polys.p, polys.data(), (size_t)n_polys * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n_mps) CK(cudaMemcpyAsync(ctx->ds->mps.p, mps.data(), (size_t)n_mps * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n_ints) CK(cudaMemcpyAsync(ctx->ds->ints.p, ints.data(), (size_t)n_ints * 4, cudaMemcpyHostToDevice, ctx->stream));
    // a1, transcendental half (tile.rs:88-101 up to the zoom-independent factor), once per dataset.  By default on the HOST with
    // the platform libm -- glibc's tan / log are what the reference (Rust f64::tan / ln on Linux) calls, so the factors are the
    // reference's to the last bit; the device's tan / log differ from glibc's in the last place for a few arguments, which
    // could flip the integer pixel of a node that sits within ~1e-9 px of an exact .5 tie.  Debug key "device_merc" selects
    // project_nodes_kernel instead (0.02 ms for 1.9 M nodes; the host loop takes ~20 ms on 16 threads).
    std::vector<double2> h_merc;
    ctx->ds->h_merc.clear();
    if (n_nodes && !ctx->device_merc) {
        h_merc.resize(n_nodes);
        const unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        auto work = [&](unsigned t) {
            const double rads_per_deg = kPi / 180.0;  // f64::to_radians
            for (size_t i = t; i < n_nodes; i += n_thr) {
                double lat, lon;
                memcpy(&lat, base[0] + i * 32 + 8, 8);
                memcpy(&lon, base[0] + i * 32 + 16, 8);
                const double lat_rad = lat * rads_per_deg;
                const double lon_rad = lon * rads_per_deg;
                const double x = lon_rad + kPi;
                const double y = kPi - std::log(std::tan((kPi / 4.0) + (lat_rad / 2.0)));
                h_merc[i] = make_double2(x / (2.0 * kPi), y / (2.0 * kPi));
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        CK(cudaMemcpyAsync(ctx->ds->merc.p, h_merc.data(), (size_t)n_nodes * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->ds->h_merc.swap(h_merc);
    } else if (n_nodes) {
        const unsigned char* raw_p = raw_nodes.p;
        project_nodes_kernel<<<(n_nodes + 255) / 256, 256, 0, ctx->stream>>>(raw_p, n_nodes, ctx->ds->merc.p);
        CK(cudaGetLastError());
    }
    CK(ctx->ds->way_box.reserve(n_ways + 1));
    CK(ctx->ds->mp_box.reserve(n_mps + 1));
    if (n_ways + n_mps) {
        Scene gs{};
        gs.merc = ctx->ds->merc.p;
        gs.ways = ctx->ds->ways.p;
        gs.polys = ctx->ds->polys.p;
        gs.mps = ctx->ds->mps.p;
        gs.ints = ctx->ds->ints.p;
        gs.n_ways = n_ways;
        gs.n_mps = n_mps;
        entity_box_kernel<<<(n_ways + n_mps + 127) / 128, 128, 0, ctx->stream>>>(gs, ctx->ds->way_box.p, ctx->ds->mp_box.p);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    raw_nodes.release();
    ctx->ds->h_bin.assign(p, p + len);
    if (!ctx->ds->h_view.parse(ctx->ds->h_bin.data(), ctx->ds->h_bin.size())) return ctx->fail(OSMR_E_INVALID, "geodata image truncated");
    ctx->ds->n_nodes = n_nodes;
    ctx->ds->n_ways = n_ways;
    ctx->ds->n_polys = n_polys;
    ctx->ds->n_mps = n_mps;
    ctx->ds->n_ints = n_ints;
    ctx->ds->has_geo = true;
    ctx->has_batch = false;
    ctx->lres.valid = false;
    for (auto& a : ctx->lres.angle) a.set = false;
    for (auto& z : ctx->zoom_tables) z.set = z.l_set = false;  // classes are per dataset

    // ---- f3: the tile index (reader.rs:135-180) + what the device-side lookup derives from it ----
    {
        const uint32_t n_idx = cnt[4];
        ctx->ds->n_idx = n_idx;
        ctx->ds->auto_unavailable.clear();
        std::vector<uint2> ixy(n_idx), iw(n_idx), im(n_idx), in_(n_idx);
        std::vector<uint8_t> node_seen(n_nodes, 0);
        ctx->ds->auto_labels_unavailable.clear();
        struct Ext {
            uint32_t x0 = 0xffffffffu, y0 = 0xffffffffu, x1 = 0, y1 = 0, n = 0;
        };
        std::vector<Ext> wext(n_ways), mext(n_mps);
        for (uint32_t i = 0; i < n_idx && ctx->ds->auto_unavailable.empty(); ++i) {
            const uint8_t* r = base[4] + (size_t)i * 32;
            ixy[i] = make_uint2(rd_u32(r), rd_u32(r + 4));
            iw[i] = make_uint2(rd_u32(r + 16), rd_u32(r + 20));
            im[i] = make_uint2(rd_u32(r + 24), rd_u32(r + 28));
            in_[i] = make_uint2(rd_u32(r + 8), rd_u32(r + 12));
            // a node belongs to one z18 tile (saver.rs:170-173): the device-side label lookup emits it without a dedup step
            if (!range_ok(in_[i].x, in_[i].y)) {
                ctx->ds->auto_labels_unavailable = "tile index node list out of range";
            } else if (ctx->ds->auto_labels_unavailable.empty()) {
                for (uint32_t k = 0; k < in_[i].y; ++k) {
                    const uint32_t nd = ints[in_[i].x + k];
                    if (nd >= n_nodes) {
                        ctx->ds->auto_labels_unavailable = "tile index references a missing node";
                        break;
                    }
                    if (node_seen[nd]) {
                        ctx->ds->auto_labels_unavailable = "tile index lists a node in more than one record";
                        break;
                    }
                    node_seen[nd] = 1;
                }
            }
            if (i && !(ixy[i - 1].x < ixy[i].x || (ixy[i - 1].x == ixy[i].x && ixy[i - 1].y < ixy[i].y)))
                ctx->ds->auto_unavailable = "tile index is not sorted by (x, y)";
            if (!range_ok(iw[i].x, iw[i].y) || !range_ok(im[i].x, im[i].y)) ctx->ds->auto_unavailable = "tile index id list out of range";
            if (!ctx->ds->auto_unavailable.empty()) break;
            auto note = [&](std::vector<Ext>& ext, uint32_t e) {
                if (e >= ext.size()) {
                    ctx->ds->auto_unavailable = "tile index references a missing entity";
                    return;
                }
                Ext& x = ext[e];
                x.x0 = std::min(x.x0, ixy[i].x);
                x.y0 = std::min(x.y0, ixy[i].y);
                x.x1 = std::max(x.x1, ixy[i].x);
                x.y1 = std::max(x.y1, ixy[i].y);
                ++x.n;
            };
            for (uint32_t k = 0; k < iw[i].y; ++k) note(wext, ints[iw[i].x + k]);
            for (uint32_t k = 0; k < im[i].y; ++k) note(mext, ints[im[i].x + k]);
        }
        // every entity must be listed in the full rectangle of index tiles between its extreme tiles (saver.rs:194-226):
        // the device-side dedup relies on it
        auto rectangular = [](const std::vector<Ext>& ext) {
            for (const Ext& x : ext)
                if (x.n && (uint64_t)(x.x1 - x.x0 + 1) * (uint64_t)(x.y1 - x.y0 + 1) != x.n) return false;
            return true;
        };
        if (ctx->ds->auto_unavailable.empty() && !(rectangular(wext) && rectangular(mext)))
            ctx->ds->auto_unavailable = "tile index does not list every entity in the full rectangle of its z18 tiles";
        if (ctx->ds->auto_unavailable.empty()) {
            std::vector<uint2> wmin(n_ways), mmin(n_mps);
            for (uint32_t i = 0; i < n_ways; ++i) wmin[i] = make_uint2(wext[i].x0, wext[i].y0);
            for (uint32_t i = 0; i < n_mps; ++i) mmin[i] = make_uint2(mext[i].x0, mext[i].y0);
            // rank of (global id, multipolygon before way, local id): the tail of the painter's order (styler.rs:172-203,268-271)
            struct Key {
                uint64_t gid;
                uint32_t is_way, loc;
            };
            std::vector<Key> keys(n_ways + (size_t)n_mps);
            for (uint32_t i = 0; i < n_mps; ++i) {
                uint64_t g;
                memcpy(&g, base[3] + (size_t)i * 24, 8);
                keys[i] = Key{g, 0u, i};
            }
            for (uint32_t i = 0; i < n_ways; ++i) {
                uint64_t g;
                memcpy(&g, base[1] + (size_t)i * 24, 8);
                keys[n_mps + i] = Key{g, 1u, i};
            }
            std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
                if (a.gid != b.gid) return a.gid < b.gid;
                if (a.is_way != b.is_way) return a.is_way < b.is_way;
                return a.loc < b.loc;
            });
            std::vector<unsigned> wrank(n_ways), mrank(n_mps), rank_entity(keys.size());
            for (size_t r = 0; r < keys.size(); ++r) {
                if (keys[r].is_way) {
                    wrank[keys[r].loc] = (unsigned)r;
                    rank_entity[r] = keys[r].loc;
                } else {
                    mrank[keys[r].loc] = (unsigned)r;
                    rank_entity[r] = keys[r].loc | OSMR_AREA_MULTIPOLYGON;
                }
            }
            CK(ctx->ds->idx_xy.reserve(n_idx + 1));
            CK(ctx->ds->idx_w.reserve(n_idx + 1));
            CK(ctx->ds->idx_m.reserve(n_idx + 1));
            CK(ctx->ds->way_min_tile.reserve(n_ways + 1));
            CK(ctx->ds->mp_min_tile.reserve(n_mps + 1));
            CK(ctx->ds->way_rank.reserve(n_ways + 1));
            CK(ctx->ds->mp_rank.reserve(n_mps + 1));
            CK(ctx->ds->rank_entity.reserve(keys.size() + 1));
            auto up = [&](void* d, const void* h, size_t bytes) {
                return bytes ? cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
            };
            CK(up(ctx->ds->idx_xy.p, ixy.data(), (size_t)n_idx * 8));
            CK(up(ctx->ds->idx_w.p, iw.data(), (size_t)n_idx * 8));
            CK(up(ctx->ds->idx_m.p, im.data(), (size_t)n_idx * 8));
            CK(up(ctx->ds->way_min_tile.p, wmin.data(), (size_t)n_ways * 8));
            CK(up(ctx->ds->mp_min_tile.p, mmin.data(), (size_t)n_mps * 8));
            CK(up(ctx->ds->way_rank.p, wrank.data(), (size_t)n_ways * 4));
            CK(up(ctx->ds->mp_rank.p, mrank.data(), (size_t)n_mps * 4));
            CK(up(ctx->ds->rank_entity.p, rank_entity.data(), keys.size() * 4));
            if (n_nodes >= OSMR_LABEL_NODE) ctx->ds->auto_labels_unavailable = "too many nodes for osmr_label.entity";
            if (ctx->ds->auto_labels_unavailable.empty()) {
                // stable order of the styled nodes behind (layer, z_index): global id, then the reader's local-id order
                std::vector<std::pair<uint64_t, uint32_t>> nk(n_nodes);
                for (uint32_t i = 0; i < n_nodes; ++i) {
                    uint64_t g;
                    memcpy(&g, base[0] + (size_t)i * 32, 8);
                    nk[i] = {g, i};
                }
                std::sort(nk.begin(), nk.end());
                std::vector<unsigned> nrank(n_nodes), nrank_entity(n_nodes);
                for (uint32_t r = 0; r < n_nodes; ++r) {
                    nrank[nk[r].second] = r;
                    nrank_entity[r] = nk[r].second;
                }
                CK(ctx->ds->idx_n.reserve(n_idx + 1));
                CK(ctx->ds->node_rank.reserve(n_nodes + 1));
                CK(ctx->ds->node_rank_entity.reserve(n_nodes + 1));
                CK(up(ctx->ds->idx_n.p, in_.data(), (size_t)n_idx * 8));
                CK(up(ctx->ds->node_rank.p, nrank.data(), (size_t)n_nodes * 4));
                CK(up(ctx->ds->node_rank_entity.p, nrank_entity.data(), (size_t)n_nodes * 4));
            }
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

int osmr_set_icons(osmr_ctx* ctx, const osmr_icon* icons, uint32_t n_icons) try {
    if (!ctx) return OSMR_E_INVALID;
    if (n_icons && !icons) return ctx->fail(OSMR_E_INVALID, "null icon table");
    cudaSetDevice(ctx->device);
    std::vector<DevIcon> meta(n_icons);
    std::vector<double4> px;
    for (uint32_t i = 0; i < n_icons; ++i) {
        if (!icons[i].rgba || icons[i].width == 0 || icons[i].height == 0) return ctx->fail(OSMR_E_INVALID, "empty icon");
        meta[i].w = icons[i].width;
        meta[i].h = icons[i].height;
        meta[i].off = (unsigned)px.size();
        meta[i].pad = 0;
        size_t n = (size_t)icons[i].width * icons[i].height;
        for (size_t k = 0; k < n; ++k) {
            const uint8_t* c = icons[i].rgba + 4 * k;
            // RgbaColor::from_components (tile_pixels.rs:24-26): opacity = a/255; channel = opacity * (c/255).
            // Plain IEEE divisions and multiplications: host and device agree bit for bit.
            volatile double opacity = (double)c[3] / 255.0;
            double4 v;
            volatile double r = (double)c[0] / 255.0, g = (double)c[1] / 255.0, b = (double)c[2] / 255.0;
            volatile double pr = opacity * r, pg = opacity * g, pb = opacity * b;
            v.x = pr;
            v.y = pg;
            v.z = pb;
            v.w = opacity;
            px.push_back(v);
        }
    }
    CK(ctx->icons.reserve(n_icons + 1));
    CK(ctx->icon_px.reserve(px.size() + 1));
    if (n_icons) CK(cudaMemcpyAsync(ctx->icons.p, meta.data(), n_icons * sizeof(DevIcon), cudaMemcpyHostToDevice, ctx->stream));
    if (!px.empty()) CK(cudaMemcpyAsync(ctx->icon_px.p, px.data(), px.size() * sizeof(double4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_icons = n_icons;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

int osmr_set_styles(osmr_ctx* ctx, const osmr_style* styles, uint32_t n_styles, const double* dashes, uint32_t n_dashes) try {
    if (!ctx) return OSMR_E_INVALID;
    if ((n_styles && !styles) || (n_dashes && !dashes)) return ctx->fail(OSMR_E_INVALID, "null style table");
    cudaSetDevice(ctx->device);
    for (uint32_t i = 0; i < n_styles; ++i) {
        const osmr_style& s = styles[i];
        if ((s.flags & OSMR_STYLE_DASHES) && ((uint64_t)s.dashes_off + s.dashes_len > n_dashes || s.dashes_len > 32))
            return ctx->fail(OSMR_E_INVALID, "style dashes out of range (at most 32 numbers)");
        if ((s.flags & OSMR_STYLE_CASING_DASHES) &&
            ((uint64_t)s.casing_dashes_off + s.casing_dashes_len > n_dashes || s.casing_dashes_len > 32))
            return ctx->fail(OSMR_E_INVALID, "style casing dashes out of range (at most 32 numbers)");
        if (s.line_cap > OSMR_CAP_SQUARE || s.casing_line_cap > OSMR_CAP_SQUARE) return ctx->fail(OSMR_E_INVALID, "bad line cap");
        // The compositor keeps RGB only: the canvas alpha stays exactly 1.0 as long as every source alpha a satisfies
        // a + fl(1 - a) == 1, which holds for a in [0, 2] (tile_pixels.rs:211-216).  The reference does not clamp `opacity` /
        // `fill-opacity`; a stylesheet with values outside that range would make its canvas alpha drift and its export divide
        // by it -- refuse such a table loudly instead of drawing different pixels.
        if ((s.flags & OSMR_STYLE_OPACITY) && !(s.opacity >= 0.0 && s.opacity <= 2.0))
            return ctx->fail(OSMR_E_INVALID, "style opacity outside [0, 2] is not supported");
        if ((s.flags & OSMR_STYLE_FILL_OPACITY) && !(s.fill_opacity >= 0.0 && s.fill_opacity <= 2.0))
            return ctx->fail(OSMR_E_INVALID, "style fill-opacity outside [0, 2] is not supported");
    }
    CK(ctx->styles.reserve(n_styles + 1));
    CK(ctx->dashes.reserve(n_dashes + 1));
    if (n_styles) CK(cudaMemcpyAsync(ctx->styles.p, styles, (size_t)n_styles * sizeof(osmr_style), cudaMemcpyHostToDevice, ctx->stream));
    if (n_dashes) CK(cudaMemcpyAsync(ctx->dashes.p, dashes, (size_t)n_dashes * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_styles = n_styles;
    ctx->n_dashes = n_dashes;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// ---------------------------------------------------------------------------------------------------------
static int validate_batch(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin) {
    if (!ctx->ds->has_geo) return ctx->fail(OSMR_E_STATE, "osmr_set_geodata has not been called");
    if (n_tiles == 0) return ctx->fail(OSMR_E_INVALID, "empty batch");
    if (!tiles || !area_begin) return ctx->fail(OSMR_E_INVALID, "null batch arrays");
    uint32_t scale = tiles[0].scale;
    if (scale < 1 || scale > 8) return ctx->fail(OSMR_E_INVALID, "scale must be 1..8");
    if (area_begin[0] != 0) return ctx->fail(OSMR_E_INVALID, "area_begin[0] must be 0");
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (tiles[t].scale != scale) return ctx->fail(OSMR_E_INVALID, "all tiles of a batch must share one scale");
        if (tiles[t].zoom > 22) return ctx->fail(OSMR_E_INVALID, "zoom must be <= 22");
        if (area_begin[t + 1] < area_begin[t]) return ctx->fail(OSMR_E_INVALID, "area_begin must be non-decreasing");
    }
    if ((uint64_t)area_begin[n_tiles] * 3ull >= 0xffffffffull) return ctx->fail(OSMR_E_INVALID, "batch too large");
    return OSMR_OK;
}

// defer_tail: copy only the styled areas of the first draw chunk on the compute stream and send the rest on the copy
// stream (event ctx->areas_ready), so that the upload overlaps the drawing of the first chunk (osmr_draw_tiles only:
// the caller's arrays stay valid until that call returns).
// Device-visible alias of a page-locked host buffer (cudaMallocHost / cudaHostRegister memory is mapped under unified
// addressing), or nullptr for pageable memory.
static unsigned char* pinned_device_alias(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost) return nullptr;
    return static_cast<unsigned char*>(a.devicePointer);
}

// Host output is drawn in chunks so that transfers overlap the drawing (all chunks are enqueued back to back):
//   staged (the default): the D2H copy of chunk i runs on its own stream while the later chunks are drawn (it really
//   overlaps only when `out` is page-locked, e.g. from osmr_alloc_pinned).  What stays exposed is the upload of the first
//   chunk's styled areas and the copy of the last chunk's tiles, so the schedule tapers: 1/4, 5/16, 1/4, 1/8, 1/16 of the
//   batch (measured on the C2 batch, B200: 4 equal chunks 11.8 ms, 8: 12.5, 16: 14.7 -- every chunk costs ~0.15 ms of
//   launch tails).  Debug key "host_chunks" = n > 0 forces n equal chunks.
//   direct (debug key "direct_out", page-locked `out`): a small first chunk hides the upload of the remaining styled
//   areas, raster_kernel stores the tiles straight into host memory (PCIe-bound stores, 26 GB/s: slower than staging).
// Counters go home by a store into page-locked host memory, not by a copy: a cudaMemcpyAsync issued early waits in the copy
// engine's queue until its stream gets there and holds up the copies that other streams issue behind it (the tile images, the
// styled-area tail) -- seen as draw chunks that started 10 ms late behind the label pass.
__global__ void export_words_kernel(const unsigned* src, unsigned* dst, unsigned n) {
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
static int export_words(osmr_ctx* ctx, const unsigned* src, unsigned* host_dst, unsigned n, cudaStream_t st) {
    unsigned* alias = reinterpret_cast<unsigned*>(pinned_device_alias(host_dst));
    if (!alias) {
        CK(cudaMemcpyAsync(host_dst, src, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        return OSMR_OK;
    }
    export_words_kernel<<<1, 64, 0, st>>>(src, alias, n);
    CK(cudaGetLastError());
    return OSMR_OK;
}

static unsigned plan_chunks(unsigned n_tiles, bool to_host, bool direct, unsigned host_chunks, unsigned resident_chunks,
                            unsigned sizes[kMaxChunks]) {
    unsigned n = 0;
    if (!to_host && resident_chunks > 1 && n_tiles >= 128 * resident_chunks) {
        // output stays in HBM: equal chunks alternating between the two compute streams (no transfers to hide, but the
        // launch tails of one chunk's kernels overlap the other chunk's work)
        const unsigned c = (n_tiles + resident_chunks - 1) / resident_chunks;
        for (unsigned done = 0; done < n_tiles; done += c) sizes[n++] = std::min(c, n_tiles - done);
    } else if (!to_host || n_tiles < 128) {
        sizes[n++] = n_tiles;
    } else if (direct) {
        sizes[n++] = std::max(32u, n_tiles / 8);
        sizes[n++] = n_tiles - sizes[0];
    } else if (host_chunks) {
        const unsigned c = std::max(64u, (n_tiles + host_chunks - 1) / host_chunks);
        for (unsigned done = 0; done < n_tiles && n < kMaxChunks; done += c) sizes[n++] = std::min(c, n_tiles - done);
        unsigned sum = 0;
        for (unsigned i = 0; i < n; ++i) sum += sizes[i];
        sizes[n - 1] += n_tiles - sum;
    } else {
        static const unsigned sixteenths[5] = {4, 5, 4, 2, 1};
        unsigned done = 0;
        for (unsigned i = 0; i < 5 && done < n_tiles; ++i) {
            unsigned c = (i == 4) ? n_tiles - done : std::min(n_tiles - done, std::max(32u, n_tiles * sixteenths[i] / 16));
            sizes[n++] = c;
            done += c;
        }
        if (done < n_tiles) sizes[n - 1] += n_tiles - done;
    }
    return n;
}

static int flush_area_tail(osmr_ctx* ctx) {
    if (!ctx->tail_src) return OSMR_OK;
    CK(cudaMemcpyAsync(ctx->tail_dst, ctx->tail_src, ctx->tail_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->areas_ready, ctx->copy_stream));
    ctx->tail_src = nullptr;
    return OSMR_OK;
}

static int batch_upload_impl(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                             const osmr_styled_area* areas, bool defer_tail, const void* host_out, bool hold_tail = false) {
    if (!ctx) return OSMR_E_INVALID;
    int rc = validate_batch(ctx, tiles, n_tiles, area_begin);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    uint32_t n_areas = area_begin[n_tiles];
    if (n_areas && !areas) return ctx->fail(OSMR_E_INVALID, "null area list");
    ctx->has_batch = false;
    ctx->batch_has_labels = false;
    CK(ctx->tiles.reserve(n_tiles));
    CK(ctx->area_begin.reserve(n_tiles + 1));
    CK(ctx->areas.reserve(n_areas + 1));
    CK(cudaMemcpyAsync(ctx->tiles.p, tiles, (size_t)n_tiles * sizeof(osmr_tile), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->area_begin.p, area_begin, (size_t)(n_tiles + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->areas_deferred = false;
    size_t head = n_areas;
    ctx->first_chunk = 0;
    if (defer_tail) {
        unsigned sizes[kMaxChunks];
        plan_chunks(n_tiles, true, ctx->direct_out && host_out && pinned_device_alias(host_out), ctx->host_chunks, 1, sizes);
        if (sizes[0] < n_tiles) {
            head = area_begin[sizes[0]];
            ctx->first_chunk = sizes[0];
        }
    }
    if (head) CK(cudaMemcpyAsync(ctx->areas.p, areas, head * sizeof(osmr_styled_area), cudaMemcpyHostToDevice, ctx->stream));
    ctx->tail_src = nullptr;
    if (head < n_areas) {
        if (hold_tail) {  // the caller has something more urgent for the copy engine first (the label lists): flush_area_tail later
            ctx->tail_src = areas + head;
            ctx->tail_dst = ctx->areas.p + head;
            ctx->tail_bytes = (n_areas - head) * sizeof(osmr_styled_area);
        } else {
            CK(cudaMemcpyAsync(ctx->areas.p + head, areas + head, (n_areas - head) * sizeof(osmr_styled_area), cudaMemcpyHostToDevice,
                               ctx->copy_stream));
            CK(cudaEventRecord(ctx->areas_ready, ctx->copy_stream));
        }
        ctx->areas_deferred = true;
    }
    ctx->n_tiles = n_tiles;
    ctx->n_areas = n_areas;
    ctx->scale = (int)tiles[0].scale;
    ctx->h_area_begin.assign(area_begin, area_begin + n_tiles + 1);
    // scratch that scales with the batch
    CK(ctx->vis.reserve(3ull * n_areas + 1));
    CK(ctx->rop.reserve(3ull * n_areas + 1));
    CK(ctx->vis_bbox.reserve(3ull * n_areas + 1));
    CK(ctx->work.reserve(3ull * n_areas + 1));
    CK(ctx->fill_work.reserve((size_t)n_areas + n_areas / 2 + 4096));  // work items; grown when a batch has more
    CK(ctx->line_work.reserve(2ull * n_areas + n_areas / 2 + 4096));
    CK(ctx->vis_count.reserve(3ull * n_tiles));
    CK(ctx->counters.reserve(CNT_COUNT));
    // osmr_batch_upload: the caller's arrays may be reused as soon as we return.  osmr_draw_tiles (defer_tail) keeps them
    // alive until the draw has finished, so the kernels are enqueued right behind the copies without a host round trip.
    if (!defer_tail) CK(cudaStreamSynchronize(ctx->stream));
    ctx->has_batch = true;
    return OSMR_OK;
}

int osmr_batch_upload(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                      const osmr_styled_area* areas) try {
    return batch_upload_impl(ctx, tiles, n_tiles, area_begin, areas, false, nullptr);
} OSMR_CATCH_INT(ctx)

// Enqueues the whole pipeline for tiles [tb, tb+tc) of the uploaded batch on the compute stream (nothing here waits for
// the device): the chunk's counters land in page-locked slot `slot` and are judged by collect_chunk after the stream
// has been synchronised.  dev_out points at tile tb's image.
static int launch_chunk(osmr_ctx* ctx, const uint8_t canvas_rgb[3], uint32_t flags, unsigned char* dev_out, unsigned tb, unsigned tc,
                        unsigned slot, unsigned set) {
    const int D = 256 * ctx->scale;
    const unsigned area_base = ctx->h_area_begin[tb];
    const unsigned n_areas = ctx->h_area_begin[tb + tc] - area_base;
    Scene s{};
    s.merc = ctx->ds->merc.p;
    s.way_box = ctx->ds->way_box.p;
    s.mp_box = ctx->ds->mp_box.p;
    s.ways = ctx->ds->ways.p;
    s.polys = ctx->ds->polys.p;
    s.mps = ctx->ds->mps.p;
    s.ints = ctx->ds->ints.p;
    s.n_nodes = ctx->ds->n_nodes;
    s.n_ways = ctx->ds->n_ways;
    s.n_polys = ctx->ds->n_polys;
    s.n_mps = ctx->ds->n_mps;
    s.n_ints = ctx->ds->n_ints;
    s.styles = ctx->styles.p;
    s.dashes = ctx->dashes.p;
    s.n_styles = ctx->n_styles;
    s.n_dashes = ctx->n_dashes;
    s.icons = ctx->icons.p;
    s.icon_px = ctx->icon_px.p;
    s.n_icons = ctx->n_icons;
    s.tiles = ctx->tiles.p + tb;
    s.area_begin = ctx->area_begin.p + tb;
    s.areas = ctx->areas.p;
    s.n_tiles = tc;
    s.n_areas = n_areas;
    s.area_base = area_base;
    s.D = D;
    s.scale = ctx->scale;
    s.flags = flags;
    if (flags & OSMR_DRAW_HAS_CANVAS_COLOR) memcpy(s.canvas, canvas_rgb, 3);
    s.vis = ctx->vis.p;
    s.rop = ctx->rop.p;
    s.vis_bbox = ctx->vis_bbox.p;
    s.vis_count = ctx->vis_count.p + 3ull * tb;  // per chunk: concurrent chunks must not share the per-tile counts
    s.work = ctx->work.p + 3ull * area_base;  // a chunk's ops fit into the slice of its own styled areas
    auto& fill_work = set ? ctx->scrB.fill_work : ctx->fill_work;
    auto& line_work = set ? ctx->scrB.line_work : ctx->line_work;
    auto& walk_alpha = set ? ctx->scrB.walk_alpha : ctx->walk_alpha;
    auto& walk_len = set ? ctx->scrB.walk_len : ctx->walk_len;
    auto& geom = set ? ctx->scrB.geom : ctx->geom;
    auto& mask = set ? ctx->scrB.mask : ctx->mask;
    s.fill_work = fill_work.p;
    s.line_work = line_work.p;
    s.fill_work_cap = (unsigned)std::min<size_t>(fill_work.cap, 0xffffffffu);
    s.line_work_cap = (unsigned)std::min<size_t>(line_work.cap, 0xffffffffu);
    if (ctx->work_items_limit) {
        s.fill_work_cap = std::min(s.fill_work_cap, ctx->work_items_limit);
        s.line_work_cap = std::min(s.line_work_cap, ctx->work_items_limit);
    }
    auto& bin_entries = set ? ctx->scrB.bin_entries : ctx->bin_entries;
    auto& frag_cnt = set ? ctx->scrB.frag_cnt : ctx->frag_cnt;
    auto& pair_table = set ? ctx->scrB.pair_table : ctx->pair_table;
    s.frag_alpha = walk_alpha.p;
    s.frag_pix = walk_len.p;
    // (16 elements of slack behind the last fragment: the bulk copies of raster_kernel fetch whole 16-byte units)
    {
        const size_t fc = std::min<size_t>(std::min<size_t>(ctx->walk_alpha_cap, walk_alpha.cap), std::min<size_t>(ctx->walk_len_cap, walk_len.cap));
        s.frag_cap = fc > 32 ? fc - 32 : 0;
    }
    s.entries = bin_entries.p;
    s.frag_cnt = frag_cnt.p;
    s.entries_cap = (unsigned)std::min<size_t>(std::min<size_t>(ctx->entries_cap, std::min(bin_entries.cap, frag_cnt.cap)), 0xfffffff0u);
    s.pair = pair_table.p;
    s.pair_cap = (unsigned)std::min<size_t>(std::min<size_t>(ctx->pair_cap, pair_table.cap), 0xfffffff0u);
    s.blk_range = ctx->blk_range.p + (size_t)tb * (size_t)((D / kBW) * (D / kBH));
    s.calc_table = ctx->calc_table.p;
    s.geom = geom.p;
    s.geom_cap = (unsigned)std::min<size_t>(std::min<size_t>(ctx->geom_cap_units, geom.cap), 0xffffffffu);
    s.mask = mask.p;
    s.mask_cap = (unsigned)std::min<size_t>(std::min<size_t>(ctx->mask_cap_words, mask.cap), 0xffffffffu);
    s.counters = ctx->counters.p + (size_t)slot * CNT_COUNT;
    s.fill_cap = ctx->fill_cap;
    s.label_plane = ctx->label_plane_active ? ctx->label_plane.p + (size_t)tb * D * D : nullptr;
    s.label_mask = ctx->label_plane_active ? ctx->label_pmask.p + (size_t)tb * (size_t)(D * D / 32) : nullptr;
    s.label_icon_px = ctx->label_icon_px.p;
    s.out = dev_out;

    cudaStream_t st = set ? ctx->stream2 : ctx->stream;
    cudaEvent_t* ev = ctx->cev[slot];
    CK(cudaEventRecord(ev[0], st));
    CK(cudaMemsetAsync(s.counters, 0, CNT_COUNT * sizeof(unsigned), st));
    unsigned launches = 0;
    if (ctx->n_styles && slot == 0) {  // the calculators depend on (styles, scale, flags) only: once per draw
        style_calc_kernel<<<(2 * ctx->n_styles + 127) / 128, 128, 0, st>>>(s, ctx->calc_table.p);
        ++launches;
    }
    if (slot == 0) CK(cudaEventRecord(ctx->prep_done, st));  // batch description + calculators are in place
    {
        // (tile, pass) lists longer than plan_slice_areas styled areas are cut into slices with a CTA each (low zooms, C4)
        unsigned max_n = 0;
        for (unsigned t = tb; t < tb + tc; ++t) max_n = std::max(max_n, ctx->h_area_begin[t + 1] - ctx->h_area_begin[t]);
        const unsigned per = ctx->plan_slice_areas ? ctx->plan_slice_areas : 16384u;
        const unsigned slices = std::min(256u, std::max(1u, (max_n + per - 1u) / per));
        s.plan_slices = slices;
        s.plan_slice_base = nullptr;
        if (slices > 1u) {
            auto& psb = set ? ctx->scrB.plan_slice_base : ctx->plan_slice_base;
            CK(psb.reserve((size_t)3u * tc * slices + 1));
            s.plan_slice_base = psb.p;
            plan_count_kernel<<<3 * tc * slices, kPlanThreads, 0, st>>>(s);
            plan_scan_kernel<<<(3 * tc + 127) / 128, 128, 0, st>>>(s);
            launches += 2;
        }
        plan_ops_kernel<<<3 * tc * slices, kPlanThreads, 0, st>>>(s);
    }
    build_geometry_kernel<<<ctx->num_sms * 8, kGeomThreads, 0, st>>>(s);
    fill_rows_kernel<<<ctx->num_sms * 16, kFillThreads, 0, st>>>(s);
    {
        const unsigned nblk = (unsigned)((D / kBW) * (D / kBH));
        const unsigned bin_ctas = tc * ((nblk + kBinThreads - 1) / kBinThreads);
        // long op lists (the tiles that got plan slices) are binned slice by slice: count, scan, write
        unsigned bslices = 1u;
        if (s.plan_slices > 1u) {
            unsigned max_n = 0;
            for (unsigned t = tb; t < tb + tc; ++t) max_n = std::max(max_n, ctx->h_area_begin[t + 1] - ctx->h_area_begin[t]);
            const unsigned per = std::max(1u, (ctx->plan_slice_areas ? ctx->plan_slice_areas : 16384u) / 4u);
            bslices = std::min(256u, std::max(1u, (max_n + per - 1u) / per));
            while (bslices > 1u && (size_t)bin_ctas * bslices * kBinThreads * 12u > ((size_t)1 << 30)) bslices /= 2u;
        }
        s.bin_slices = bslices;
        if (bslices > 1u) {
            auto& bce = set ? ctx->scrB.bin_cnt_ent : ctx->bin_cnt_ent;
            auto& bcc = set ? ctx->scrB.bin_cnt_cap : ctx->bin_cnt_cap;
            CK(bce.reserve((size_t)bin_ctas * bslices * kBinThreads + 1));
            CK(bcc.reserve((size_t)bin_ctas * bslices * kBinThreads + 1));
            s.bin_cnt_ent = bce.p;
            s.bin_cnt_cap = bcc.p;
            bin_count_kernel<<<bin_ctas * bslices, kBinThreads, 0, st>>>(s);
            bin_scan_kernel<<<bin_ctas, kBinThreads, 0, st>>>(s);
            bin_write_kernel<<<bin_ctas * bslices, kBinThreads, 0, st>>>(s);
            launches += 2;
        } else {
            bin_ops_kernel<<<bin_ctas, kBinThreads, 0, st>>>(s);
        }
    }
    CK(cudaEventRecord(ev[3], st));
    line_cover_kernel<<<ctx->num_sms * 8, kCoverThreads, 0, st>>>(s);
    launches += 5;
    CK(cudaEventRecord(ev[1], st));
    const unsigned blocks = (unsigned)((D / kBW) * (D / kBH));
    if (ctx->label_async && ctx->label_plane_active)
        for (unsigned i = 0; i < ctx->n_lchunks; ++i)
            if (ctx->lchunk_tb[i] < tb + tc && tb < ctx->lchunk_tb[i] + ctx->lchunk_tc[i]) CK(cudaStreamWaitEvent(st, ctx->lchunk_done[i], 0));
    CK(cudaEventRecord(ev[4], st));  // (after the wait: the stage time of raster_kernel does not include the label pass)
    raster_kernel<<<tc * blocks, kRasterThreads, 0, st>>>(s);
    ++launches;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[2], st));
    {
        int rc = export_words(ctx, s.counters, ctx->h_cnt.p + (size_t)slot * CNT_COUNT, CNT_COUNT, st);
        if (rc) return rc;
        ++launches;
    }
    ctx->chunk_launches[slot] = launches;
    return OSMR_OK;
}

// the second scratch set mirrors the first one's sizes
static int ensure_scrB(osmr_ctx* ctx) {
    CK(ctx->scrB.geom.reserve(ctx->geom.cap));
    CK(ctx->scrB.mask.reserve(ctx->mask.cap));
    CK(ctx->scrB.fill_work.reserve(ctx->fill_work.cap));
    CK(ctx->scrB.line_work.reserve(ctx->line_work.cap));
    CK(ctx->scrB.walk_alpha.reserve(ctx->walk_alpha.cap));
    CK(ctx->scrB.walk_len.reserve(ctx->walk_len.cap));
    CK(ctx->scrB.bin_entries.reserve(ctx->bin_entries.cap));
    CK(ctx->scrB.frag_cnt.reserve(ctx->frag_cnt.cap));
    CK(ctx->scrB.pair_table.reserve(ctx->pair_table.cap));
    return OSMR_OK;
}

// After the compute stream has been synchronised: errors of chunk `slot`, scratch growth on overflow (*redo = true), else
// its statistics.
static int collect_chunk(osmr_ctx* ctx, unsigned slot, unsigned tb, unsigned tc, bool* redo) {
    const unsigned* h_cnt = ctx->h_cnt.p + (size_t)slot * CNT_COUNT;
    const unsigned n_areas = ctx->h_area_begin[tb + tc] - ctx->h_area_begin[tb];
    if (h_cnt[CNT_BAD_INPUT] & 1u) return ctx->fail(OSMR_E_INVALID, "styled area references an entity or style that does not exist");
    if (h_cnt[CNT_BAD_INPUT] & 2u) return ctx->fail(OSMR_E_INVALID, "line wider than 65000 pixels (width * scale): not supported");
    if (h_cnt[CNT_WALK_TRUNC] & 1u) return ctx->fail(OSMR_E_CUDA, "internal error: a perpendicular walk exceeded its proven bound");
    if (h_cnt[CNT_WALK_TRUNC] & 2u) return ctx->fail(OSMR_E_CUDA, "internal error: a line fragment fell into a block the binning had ruled out");
    if (h_cnt[CNT_OVERFLOW] & 256u) return ctx->fail(OSMR_E_CUDA, "internal error: a fragment list outgrew its proven capacity");
    if (h_cnt[CNT_OVERFLOW]) {  // grow the scratch that ran out; the caller redoes the draw
        *redo = true;
        if (h_cnt[CNT_OVERFLOW] & 1u) {
            size_t need = (size_t)h_cnt[CNT_GEOM_USED] + (size_t)h_cnt[CNT_GEOM_USED] / 2 + 1024;
            if (need <= ctx->geom_cap_units) need = ctx->geom_cap_units * 2;
            if (need >= 0xffffffffull) return ctx->fail(OSMR_E_NOMEM, "geometry scratch exceeds 64 GiB; split the batch");
            CK(ctx->geom.reserve(need));
            ctx->geom_cap_units = ctx->geom.cap;
        }
        if (h_cnt[CNT_OVERFLOW] & 4u) {  // fragment storage: the counter holds what the batch asked for (all CTAs add before they check)
            unsigned long long want;
            memcpy(&want, &h_cnt[CNT_WALK_ALPHA], 8);
            if (want >= 0xfffffff0ull) return ctx->fail(OSMR_E_NOMEM, "more than 2^32 line fragment slots; split the batch");
            size_t need = (size_t)want + (size_t)want / 8 + 4096;
            if (need <= ctx->walk_alpha_cap) need = ctx->walk_alpha_cap * 2;
            if (ctx->walk_alpha.reserve(need) != cudaSuccess || ctx->walk_len.reserve(need) != cudaSuccess)
                return ctx->fail(OSMR_E_NOMEM, "line fragment storage does not fit in device memory; split the batch");
            ctx->walk_alpha_cap = ctx->walk_alpha.cap;
            ctx->walk_len_cap = ctx->walk_len.cap;
        }
        if (h_cnt[CNT_OVERFLOW] & 64u) {
            size_t need = (size_t)h_cnt[CNT_BIN_ENTRIES] + (size_t)h_cnt[CNT_BIN_ENTRIES] / 8 + 4096;
            if (need <= ctx->entries_cap) need = ctx->entries_cap * 2;
            CK(ctx->bin_entries.reserve(need));
            CK(ctx->frag_cnt.reserve(need));
            ctx->entries_cap = std::min(ctx->bin_entries.cap, ctx->frag_cnt.cap);
        }
        if (h_cnt[CNT_OVERFLOW] & 128u) {
            size_t need = (size_t)h_cnt[CNT_PAIR_USED] + (size_t)h_cnt[CNT_PAIR_USED] / 8 + 4096;
            if (need <= ctx->pair_cap) need = ctx->pair_cap * 2;
            if (need >= 0xfffffff0ull) return ctx->fail(OSMR_E_NOMEM, "pair tables exceed 2^32 cells; split the batch");
            CK(ctx->pair_table.reserve(need));
            ctx->pair_cap = ctx->pair_table.cap;
        }
        if (h_cnt[CNT_OVERFLOW] & 16u) CK(ctx->fill_work.reserve((size_t)h_cnt[CNT_N_FILL_WORK] + 1024));
        if (h_cnt[CNT_OVERFLOW] & 32u) CK(ctx->line_work.reserve((size_t)h_cnt[CNT_N_LINE_WORK] + 1024));
        if (h_cnt[CNT_OVERFLOW] & 48u) ctx->work_items_limit = 0;
        if (h_cnt[CNT_OVERFLOW] & 2u) {
            size_t need = (size_t)h_cnt[CNT_MASK_USED] + (size_t)h_cnt[CNT_MASK_USED] / 2 + 1024;
            if (need <= ctx->mask_cap_words) need = ctx->mask_cap_words * 2;
            if (need >= 0xffffffffull) return ctx->fail(OSMR_E_NOMEM, "mask scratch exceeds 16 GiB; split the batch");
            CK(ctx->mask.reserve(need));
            ctx->mask_cap_words = ctx->mask.cap;
        }
        return OSMR_OK;
    }
    cudaEvent_t* ev = ctx->cev[slot];
    float ms_plan = 0, ms_cover = 0, ms_raster = 0;
    cudaEventElapsedTime(&ms_plan, ev[0], ev[3]);
    cudaEventElapsedTime(&ms_cover, ev[3], ev[1]);
    cudaEventElapsedTime(&ms_raster, ev[4], ev[2]);
    unsigned long long used_alpha, steps;
    memcpy(&used_alpha, &h_cnt[CNT_WALK_ALPHA], 8);
    memcpy(&steps, &h_cnt[CNT_WALK_STEPS], 8);
    ctx->stats.walk_bytes += used_alpha * 9ull;  // fragment slots handed out (8-byte alpha + pixel byte each)
    ctx->stats.walk_steps += steps;              // fragments stored
    ctx->stats.n_tiles += tc;
    ctx->stats.n_areas += n_areas;
    ctx->stats.n_visible_ops += h_cnt[CNT_VISIBLE];
    ctx->stats.n_node_refs += ((uint64_t)h_cnt[CNT_NODE_REFS_HI] << 32) | h_cnt[CNT_NODE_REFS_LO];
    ctx->stats.kernel_launches += ctx->chunk_launches[slot];
    ctx->stats.geom_bytes += (uint64_t)h_cnt[CNT_GEOM_USED] * 16ull;
    ctx->stats.mask_bytes += (uint64_t)h_cnt[CNT_MASK_USED] * 4ull;
    ctx->stats.ms_plan += ms_plan;
    ctx->stats.ms_raster += ms_raster;
    ctx->stats.ms_cover += ms_cover;
    ctx->stats.ms_total += ms_plan + ms_cover + ms_raster;
    return OSMR_OK;
}

int osmr_batch_draw(osmr_ctx* ctx, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out, float* gpu_ms) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!ctx->has_batch) return ctx->fail(OSMR_E_STATE, "no batch uploaded");
    if ((flags & OSMR_DRAW_HAS_CANVAS_COLOR) && !canvas_rgb) return ctx->fail(OSMR_E_INVALID, "null canvas colour");
    cudaSetDevice(ctx->device);
    const int Di = 256 * ctx->scale;
    const size_t D = (size_t)Di;
    const size_t bytes = (size_t)ctx->n_tiles * D * D * ((flags & OSMR_DRAW_OUT_RGBA) ? 4 : 3);
    const bool to_host = out && !(flags & OSMR_DRAW_OUT_DEVICE);
    unsigned char* alias = (to_host && ctx->direct_out) ? pinned_device_alias(out) : nullptr;
    unsigned char* dev_out;
    if (out && (flags & OSMR_DRAW_OUT_DEVICE)) {
        dev_out = out;
    } else if (alias) {
        dev_out = alias;  // raster_kernel writes into the caller's page-locked buffer
    } else {
        CK(ctx->out.reserve(bytes));
        dev_out = ctx->out.p;
        ctx->out_bytes = bytes;
    }
    const size_t tile_bytes = D * D * ((flags & OSMR_DRAW_OUT_RGBA) ? 4 : 3);
    const bool staged = to_host && !alias;
    unsigned sizes[kMaxChunks];
    const unsigned n_planned = plan_chunks(ctx->n_tiles, to_host, alias != nullptr, ctx->host_chunks, ctx->resident_chunks, sizes);
    if (ctx->areas_deferred && ctx->first_chunk && sizes[0] != ctx->first_chunk)
        return ctx->fail(OSMR_E_STATE, "internal error: draw chunks differ from the upload split");
    CK(ctx->counters.reserve((size_t)kMaxChunks * CNT_COUNT));
    CK(ctx->h_cnt.reserve((size_t)kMaxChunks * CNT_COUNT));
    CK(ctx->calc_table.reserve((size_t)2 * ctx->n_styles * kCalcEntryUnits + 1));
    for (int attempt = 0; attempt < 12; ++attempt) {
        // scratch: first guesses, grown by collect_chunk when a bump allocator overflowed
        if (ctx->geom_cap_units == 0) {
            CK(ctx->geom.reserve((size_t)ctx->n_areas * 8 + (1u << 20)));
            ctx->geom_cap_units = ctx->geom.cap;
        }
        if (ctx->mask_cap_words == 0) {
            CK(ctx->mask.reserve((size_t)ctx->n_tiles * D * (D / 32) * 4 + (1u << 20)));
            ctx->mask_cap_words = ctx->mask.cap;
        }
        if (ctx->walk_alpha_cap == 0 || ctx->walk_len_cap == 0) {  // fragment slots: 1.5 M per tile to start with
            CK(ctx->walk_alpha.reserve((size_t)ctx->n_tiles * (1536u << 10) / 4 + (1u << 20)));
            CK(ctx->walk_len.reserve(ctx->walk_alpha.cap));
            ctx->walk_alpha_cap = ctx->walk_alpha.cap;
            ctx->walk_len_cap = ctx->walk_len.cap;
        }
        if (ctx->entries_cap == 0) {
            CK(ctx->bin_entries.reserve((size_t)ctx->n_areas * 4 + (1u << 20)));
            CK(ctx->frag_cnt.reserve(ctx->bin_entries.cap));
            ctx->entries_cap = std::min(ctx->bin_entries.cap, ctx->frag_cnt.cap);
        }
        if (ctx->pair_cap == 0) {
            CK(ctx->pair_table.reserve((size_t)ctx->n_areas * 4 + (1u << 20)));
            ctx->pair_cap = ctx->pair_table.cap;
        }
        CK(ctx->blk_range.reserve((size_t)ctx->n_tiles * (size_t)((Di / kBW) * (Di / kBH)) + 1));
        ctx->stats = osmr_stats{};
        // every chunk is enqueued without waiting for the previous one; with staged host output the D2H copy of chunk i
        // (third stream) runs while the later chunks are drawn.  The chunks alternate between two compute streams with
        // their own scratch, so the plan / geometry / cover kernels of chunk i+1 fill the SMs that the tail of chunk i's
        // kernels leaves idle (a chunk of a few hundred tiles cannot keep 148 SMs busy through its launch tails).
        const bool use2 = ctx->two_streams && n_planned > 1;
        if (use2) {
            int rcb = ensure_scrB(ctx);
            if (rcb) return rcb;
        }
        unsigned n_chunks = 0;
        unsigned cb[kMaxChunks], cc[kMaxChunks];
        CK(cudaEventRecord(ctx->ev_wall0, ctx->stream));
        for (unsigned tb = 0; tb < ctx->n_tiles;) {
            const unsigned tc = sizes[n_chunks];
            if (n_chunks >= n_planned || tc == 0 || tb + tc > ctx->n_tiles) return ctx->fail(OSMR_E_STATE, "internal error: bad draw chunk plan");
            const unsigned set = use2 ? (n_chunks & 1u) : 0u;
            cudaStream_t cst = set ? ctx->stream2 : ctx->stream;
            if (n_chunks == 1 && set) CK(cudaStreamWaitEvent(cst, ctx->prep_done, 0));
            if (tb > 0 && ctx->areas_deferred) {  // the tail of the styled-area list was uploaded on the copy stream
                CK(cudaStreamWaitEvent(cst, ctx->areas_ready, 0));
                if (!use2 || n_chunks >= 2) ctx->areas_deferred = false;  // both compute streams have waited
            }
            int rc = launch_chunk(ctx, canvas_rgb, flags, dev_out + (size_t)tb * tile_bytes, tb, tc, n_chunks, set);
            if (rc) {
                cudaStreamSynchronize(ctx->stream);
                cudaStreamSynchronize(ctx->stream2);
                cudaStreamSynchronize(ctx->d2h_stream);
                return rc;
            }
            if (staged) {
                CK(cudaEventRecord(ctx->chunk_done[n_chunks], cst));
                CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->chunk_done[n_chunks], 0));
                CK(cudaMemcpyAsync(out + (size_t)tb * tile_bytes, dev_out + (size_t)tb * tile_bytes, (size_t)tc * tile_bytes,
                                   cudaMemcpyDeviceToHost, ctx->d2h_stream));
            }
            cb[n_chunks] = tb;
            cc[n_chunks] = tc;
            ++n_chunks;
            tb += tc;
        }
        if (use2) {  // the end of the draw on the device: both streams
            CK(cudaEventRecord(ctx->join2, ctx->stream2));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->join2, 0));
        }
        CK(cudaEventRecord(ctx->ev_wall1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream2));
        ctx->areas_deferred = false;
        bool redo = false;
        for (unsigned c = 0; c < n_chunks; ++c) {
            int rc = collect_chunk(ctx, c, cb[c], cc[c], &redo);
            if (rc) {
                cudaStreamSynchronize(ctx->d2h_stream);
                return rc;
            }
        }
        if (redo) continue;  // (copies of the incomplete images are simply overwritten by the second round, in stream order)
        if (staged) CK(cudaStreamSynchronize(ctx->d2h_stream));
        ctx->last_draw_chunks = n_chunks;
        if (gpu_ms) {  // device wall time of the draw (with several chunks on two streams the stage times overlap)
            float wall = 0.f;
            cudaEventElapsedTime(&wall, ctx->ev_wall0, ctx->ev_wall1);
            *gpu_ms = wall;
        }
        return OSMR_OK;
    }
    cudaStreamSynchronize(ctx->d2h_stream);
    return ctx->fail(OSMR_E_NOMEM, "scratch kept overflowing");
} OSMR_CATCH_INT(ctx)

int osmr_batch_output(osmr_ctx* ctx, const uint8_t** dev_ptr, size_t* n_bytes) try {
    if (!ctx || !dev_ptr || !n_bytes) return OSMR_E_INVALID;
    *dev_ptr = ctx->out.p;
    *n_bytes = ctx->out_bytes;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

int osmr_draw_tiles(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                    const osmr_styled_area* areas, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!out) return ctx->fail(OSMR_E_INVALID, "null output buffer");
    int rc = batch_upload_impl(ctx, tiles, n_tiles, area_begin, areas, !(flags & OSMR_DRAW_OUT_DEVICE), out);
    if (rc) return rc;
    rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, nullptr);
    if (ctx->areas_deferred) {  // error path before the tail was consumed: do not leave a copy of the caller's memory in flight
        cudaStreamSynchronize(ctx->copy_stream);
        ctx->areas_deferred = false;
    }
    return rc;
} OSMR_CATCH_INT(ctx)

// ---------------------------------------------------------------------------------------------------------
// f3: styles per zoom + device-side candidate lookup and ordering (osmr_auto.cuh)
// ---------------------------------------------------------------------------------------------------------
static int encode_png_from_device(osmr_ctx* ctx, const unsigned char* rgb_dev, uint32_t n_tiles, unsigned scale, uint8_t* png_out,
                                  size_t png_cap, uint64_t* png_offset);
int osmr_set_zoom_styles(osmr_ctx* ctx, uint32_t zoom, const uint32_t* way_class, const uint32_t* mp_class, const uint32_t* class_begin,
                         const osmr_class_style* class_styles, uint32_t n_classes) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!ctx->ds->has_geo) return ctx->fail(OSMR_E_STATE, "osmr_set_geodata has not been called");
    if (zoom > 18) return ctx->fail(OSMR_E_INVALID, "zoom must be <= 18 (tile.rs:5 MAX_ZOOM)");
    if ((ctx->ds->n_ways && !way_class) || (ctx->ds->n_mps && !mp_class) || !class_begin) return ctx->fail(OSMR_E_INVALID, "null class table");
    if (class_begin[0] != 0) return ctx->fail(OSMR_E_INVALID, "class_begin[0] must be 0");
    for (uint32_t c = 0; c < n_classes; ++c) {
        if (class_begin[c + 1] < class_begin[c]) return ctx->fail(OSMR_E_INVALID, "class_begin must be non-decreasing");
        if (class_begin[c + 1] - class_begin[c] >= (1u << kAutoWithinBits)) return ctx->fail(OSMR_E_INVALID, "more than 4095 styles in one class");
    }
    const uint32_t n_cs = class_begin[n_classes];
    if (n_cs && !class_styles) return ctx->fail(OSMR_E_INVALID, "null class style list");
    for (uint32_t i = 0; i < n_cs; ++i)
        if (class_styles[i].order >= (1u << kAutoOrderBits)) return ctx->fail(OSMR_E_INVALID, "style order rank must be below 2^20");
    cudaSetDevice(ctx->device);
    osmr_ctx::ZoomTable& z = ctx->zoom_tables[zoom];
    z.set = false;
    CK(z.way_class.reserve(ctx->ds->n_ways + 1));
    CK(z.mp_class.reserve(ctx->ds->n_mps + 1));
    CK(z.class_begin.reserve(n_classes + 1));
    CK(z.class_styles.reserve(n_cs + 1));
    CK(z.class_reach.reserve(n_classes + 1));
    if (ctx->ds->n_ways) CK(cudaMemcpyAsync(z.way_class.p, way_class, (size_t)ctx->ds->n_ways * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->ds->n_mps) CK(cudaMemcpyAsync(z.mp_class.p, mp_class, (size_t)ctx->ds->n_mps * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(z.class_begin.p, class_begin, (size_t)(n_classes + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n_cs) CK(cudaMemcpyAsync(z.class_styles.p, class_styles, (size_t)n_cs * sizeof(osmr_class_style), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    z.n_classes = n_classes;
    z.n_class_styles = n_cs;
    z.set = true;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// (validation, the per-call arrays and the tile upload: the part both halves of f3 need first)
static int auto_begin(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags) {
    if (!ctx->ds->has_geo) return ctx->fail(OSMR_E_STATE, "osmr_set_geodata has not been called");
    if (!ctx->ds->auto_unavailable.empty()) return ctx->fail(OSMR_E_STATE, ctx->ds->auto_unavailable.c_str());
    if (n_tiles == 0 || !tiles) return ctx->fail(OSMR_E_INVALID, "empty batch");
    const uint32_t zoom = tiles[0].zoom, scale = tiles[0].scale;
    if (scale < 1 || scale > 8) return ctx->fail(OSMR_E_INVALID, "scale must be 1..8");
    if (zoom > 18) return ctx->fail(OSMR_E_INVALID, "zoom must be <= 18 (tile.rs:5 MAX_ZOOM)");
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (tiles[t].zoom != zoom || tiles[t].scale != scale) return ctx->fail(OSMR_E_INVALID, "all tiles of a batch must share one zoom and one scale");
        if (((uint64_t)tiles[t].x + 2) << (18 - zoom) > 0xffffffffull || ((uint64_t)tiles[t].y + 2) << (18 - zoom) > 0xffffffffull)
            return ctx->fail(OSMR_E_INVALID, "tile coordinates out of range for the zoom");
    }
    osmr_ctx::ZoomTable& z = ctx->zoom_tables[zoom];
    if (!z.set) return ctx->fail(OSMR_E_STATE, "osmr_set_zoom_styles has not been called for this zoom");
    if ((flags & OSMR_DRAW_HAS_CANVAS_COLOR) && !canvas_rgb) return ctx->fail(OSMR_E_INVALID, "null canvas colour");
    cudaSetDevice(ctx->device);
    ctx->has_batch = false;
    ctx->areas_deferred = false;
    ctx->first_chunk = 0;
    cudaStream_t st = ctx->stream;
    CK(ctx->tiles.reserve(n_tiles));
    CK(ctx->area_begin.reserve(n_tiles + 1));
    CK(ctx->counters.reserve((size_t)(kMaxChunks + 1) * CNT_COUNT));
    CK(ctx->h_cnt.reserve((size_t)(kMaxChunks + 1) * CNT_COUNT));
    CK(ctx->auto_bound.reserve(n_tiles + 2));
    CK(ctx->auto_cand_cnt.reserve(n_tiles + 1));
    CK(ctx->auto_inst.reserve(n_tiles + 2));
    CK(cudaMemcpyAsync(ctx->tiles.p, tiles, (size_t)n_tiles * sizeof(osmr_tile), cudaMemcpyHostToDevice, st));
    ctx->scale = (int)scale;
    return OSMR_OK;
}

// f3 up to "an ordinary resident batch": candidate lookup, ownership dedup, culling and the painter's order on the device.
// On success ctx holds the batch (tiles, area_begin, areas) exactly as osmr_batch_upload would have left it and the events
// ctx->ev[0] / ev[1] bracket the stage.
static int auto_prepare(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags, bool begun = false) {
    if (!begun) {
        int rc = auto_begin(ctx, tiles, n_tiles, canvas_rgb, flags);
        if (rc) return rc;
    }
    const uint32_t zoom = tiles[0].zoom, scale = tiles[0].scale;
    osmr_ctx::ZoomTable& z = ctx->zoom_tables[zoom];
    cudaStream_t st = ctx->stream;

    Scene s{};
    s.merc = ctx->ds->merc.p;
    s.way_box = ctx->ds->way_box.p;
    s.mp_box = ctx->ds->mp_box.p;
    s.ways = ctx->ds->ways.p;
    s.polys = ctx->ds->polys.p;
    s.mps = ctx->ds->mps.p;
    s.ints = ctx->ds->ints.p;
    s.n_nodes = ctx->ds->n_nodes;
    s.n_ways = ctx->ds->n_ways;
    s.n_polys = ctx->ds->n_polys;
    s.n_mps = ctx->ds->n_mps;
    s.n_ints = ctx->ds->n_ints;
    s.styles = ctx->styles.p;
    s.dashes = ctx->dashes.p;
    s.n_styles = ctx->n_styles;
    s.n_dashes = ctx->n_dashes;
    s.n_icons = ctx->n_icons;
    s.tiles = ctx->tiles.p;
    s.n_tiles = n_tiles;
    s.D = 256 * (int)scale;
    s.scale = (int)scale;
    s.flags = flags;
    unsigned* auto_counters = ctx->counters.p + (size_t)kMaxChunks * CNT_COUNT;
    unsigned* h_auto = ctx->h_cnt.p + (size_t)kMaxChunks * CNT_COUNT;
    s.counters = auto_counters;
    AutoScene a{};
    a.idx_xy = ctx->ds->idx_xy.p;
    a.idx_w = ctx->ds->idx_w.p;
    a.idx_m = ctx->ds->idx_m.p;
    a.n_idx = ctx->ds->n_idx;
    a.way_min_tile = ctx->ds->way_min_tile.p;
    a.mp_min_tile = ctx->ds->mp_min_tile.p;
    a.way_rank = ctx->ds->way_rank.p;
    a.mp_rank = ctx->ds->mp_rank.p;
    a.rank_entity = ctx->ds->rank_entity.p;
    a.way_class = z.way_class.p;
    a.mp_class = z.mp_class.p;
    a.class_begin = z.class_begin.p;
    a.class_styles = z.class_styles.p;
    a.n_classes = z.n_classes;
    a.n_class_styles = z.n_class_styles;
    a.class_reach = z.class_reach.p;
    a.bound = ctx->auto_bound.p;
    a.cand_cnt = ctx->auto_cand_cnt.p;
    a.inst_cnt = ctx->auto_inst.p;

    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaMemsetAsync(auto_counters, 0, CNT_COUNT * sizeof(unsigned), st));
    if (z.n_classes) class_reach_kernel<<<(z.n_classes + 127) / 128, 128, 0, st>>>(s, a);
    // candidates per tile: upper bound -> offsets -> compact list
    auto_bound_kernel<<<n_tiles, kAutoThreads, 0, st>>>(s, a);
    auto_scan_kernel<<<1, 1024, 0, st>>>(ctx->auto_bound.p, n_tiles, &auto_counters[CNT_OVERFLOW]);
    unsigned total_bound = 0;
    CK(cudaMemcpyAsync(&total_bound, ctx->auto_bound.p + n_tiles, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_auto, auto_counters, CNT_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_auto[CNT_OVERFLOW]) return ctx->fail(OSMR_E_NOMEM, "more than 2^32 candidate references; split the batch");
    if (h_auto[CNT_BAD_INPUT]) return ctx->fail(OSMR_E_INVALID, "class style list references a style that does not exist");
    CK(ctx->auto_cand.reserve((size_t)total_bound + 1));
    a.cand = ctx->auto_cand.p;
    auto_gather_kernel<<<n_tiles, kAutoThreads, 0, st>>>(s, a);
    auto_scan_kernel<<<1, 1024, 0, st>>>(ctx->auto_inst.p, n_tiles, &auto_counters[CNT_OVERFLOW]);
    // area_begin = the scanned styled-area counts (device + host copy)
    CK(cudaMemcpyAsync(ctx->area_begin.p, ctx->auto_inst.p, (size_t)(n_tiles + 1) * 4, cudaMemcpyDeviceToDevice, st));
    ctx->h_area_begin.resize(n_tiles + 1);
    CK(cudaMemcpyAsync(ctx->h_area_begin.data(), ctx->auto_inst.p, (size_t)(n_tiles + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_auto, auto_counters, CNT_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_auto[CNT_OVERFLOW]) return ctx->fail(OSMR_E_NOMEM, "more than 2^32 styled areas; split the batch");
    if (h_auto[CNT_BAD_INPUT]) return ctx->fail(OSMR_E_INVALID, "tile index references an entity that does not exist");
    const uint32_t n_areas = ctx->h_area_begin[n_tiles];
    if ((uint64_t)n_areas * 3ull >= 0xffffffffull) return ctx->fail(OSMR_E_INVALID, "batch too large");
    CK(ctx->areas.reserve((size_t)n_areas + 1));
    a.areas_out = ctx->areas.p;
    for (int attempt = 0;; ++attempt) {
        a.big_keys = ctx->auto_big_keys.p;
        a.big_cap = ctx->auto_big_keys.cap;
        CK(cudaMemsetAsync(auto_counters, 0, CNT_COUNT * sizeof(unsigned), st));
        auto_sort_kernel<<<n_tiles, kAutoThreads, 0, st>>>(s, a);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_auto, auto_counters, CNT_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!(h_auto[CNT_OVERFLOW] & 8u)) break;
        if (attempt >= 2) return ctx->fail(OSMR_E_NOMEM, "sort scratch kept overflowing");
        unsigned long long need;
        memcpy(&need, &h_auto[CNT_WALK_ALPHA], 8);
        CK(ctx->auto_big_keys.reserve((size_t)need + 1024));
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    // from here on this is an ordinary resident batch
    ctx->n_tiles = n_tiles;
    ctx->n_areas = n_areas;
    ctx->scale = (int)scale;
    CK(ctx->vis.reserve(3ull * n_areas + 1));
    CK(ctx->rop.reserve(3ull * n_areas + 1));
    CK(ctx->vis_bbox.reserve(3ull * n_areas + 1));
    CK(ctx->work.reserve(3ull * n_areas + 1));
    CK(ctx->fill_work.reserve((size_t)n_areas + n_areas / 2 + 4096));  // work items; grown when a batch has more
    CK(ctx->line_work.reserve(2ull * n_areas + n_areas / 2 + 4096));
    CK(ctx->vis_count.reserve(3ull * n_tiles));
    ctx->has_batch = true;
    return OSMR_OK;
}

int osmr_draw_tiles_auto(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out) try {
    if (!ctx) return OSMR_E_INVALID;
    int rc = auto_prepare(ctx, tiles, n_tiles, canvas_rgb, flags);
    if (rc) return rc;
    float ms_auto = 0.f;  // the stage's events are reused by later stages: read them before the draw
    cudaEventSynchronize(ctx->ev[1]);
    cudaEventElapsedTime(&ms_auto, ctx->ev[0], ctx->ev[1]);
    rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, nullptr);
    ctx->stats.ms_auto = ms_auto;
    ctx->stats.ms_total += ms_auto;
    return rc;
} OSMR_CATCH_INT(ctx)

// Drawer::draw_tile for a tile list (drawer.rs:40-58 as the server calls it, http_server.rs:150-177): f3 in front of the draw
// path, f4 behind it -- only the tile list goes to the device and only PNG files come back.
int osmr_draw_tiles_auto_png(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags,
                             uint8_t* png_out, size_t png_cap, uint64_t* png_offset) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!png_out || !png_offset) return ctx->fail(OSMR_E_INVALID, "null output buffer");
    if (flags & (OSMR_DRAW_OUT_RGBA | OSMR_DRAW_OUT_DEVICE)) return ctx->fail(OSMR_E_INVALID, "osmr_draw_tiles_auto_png encodes RGB into host memory");
    int rc = auto_prepare(ctx, tiles, n_tiles, canvas_rgb, flags);
    if (rc) return rc;
    float ms_auto = 0.f;
    cudaEventSynchronize(ctx->ev[1]);
    cudaEventElapsedTime(&ms_auto, ctx->ev[0], ctx->ev[1]);
    rc = osmr_batch_draw(ctx, canvas_rgb, flags, nullptr, nullptr);  // the RGB tiles stay in HBM (ctx->out)
    if (rc) return rc;
    ctx->stats.ms_auto = ms_auto;
    ctx->stats.ms_total += ms_auto;
    return encode_png_from_device(ctx, ctx->out.p, n_tiles, (unsigned)ctx->scale, png_out, png_cap, png_offset);
} OSMR_CATCH_INT(ctx)

int osmr_auto_readback(osmr_ctx* ctx, uint32_t* area_begin, osmr_styled_area* areas, uint32_t areas_cap) try {
    if (!ctx || !area_begin) return OSMR_E_INVALID;
    if (!ctx->has_batch) return ctx->fail(OSMR_E_STATE, "no batch");
    cudaSetDevice(ctx->device);
    memcpy(area_begin, ctx->h_area_begin.data(), (size_t)(ctx->n_tiles + 1) * 4);
    if (areas) {
        if (areas_cap < ctx->n_areas) return ctx->fail(OSMR_E_INVALID, "areas_cap too small");
        if (ctx->n_areas) CK(cudaMemcpyAsync(areas, ctx->areas.p, (size_t)ctx->n_areas * sizeof(osmr_styled_area), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// ---------------------------------------------------------------------------------------------------------
// f4: PNG files instead of RGB triples (osmr_png.cuh)
// ---------------------------------------------------------------------------------------------------------
static unsigned png_band_cap_words(unsigned scale) {
    const unsigned long long D = 256ull * scale, rows = D / kPngBands, np = 3 * D + 1;
    return (unsigned)((rows * np * 9ull + 128ull) / 32ull + 8ull);  // every byte a 9-bit literal + block framing
}

size_t osmr_png_bound(uint32_t scale) try {
    if (scale < 1 || scale > 8) return 0;
    return (size_t)kPngFixed + (size_t)kPngBands * png_band_cap_words(scale) * 4u;
} OSMR_CATCH_VAL(0)

// n_tiles RGB images of (256 * scale)^2 pixels in HBM -> packed PNG files in host memory
static int encode_png_from_device(osmr_ctx* ctx, const unsigned char* rgb_dev, uint32_t n_tiles, unsigned scale, uint8_t* png_out,
                                  size_t png_cap, uint64_t* png_offset) {
    cudaStream_t st = ctx->stream;
    PngScene ps{};
    ps.rgb = rgb_dev;
    ps.D = 256 * (int)scale;
    ps.n_tiles = n_tiles;
    ps.band_cap_words = png_band_cap_words(scale);
    const size_t n_bands = (size_t)n_tiles * kPngBands;
    CK(ctx->png_band_words.reserve(n_bands * ps.band_cap_words));
    CK(ctx->png_band_bytes.reserve(n_bands));
    CK(ctx->png_band_len.reserve(n_bands));
    CK(ctx->png_band_adler.reserve(n_bands));
    CK(ctx->png_tile_off.reserve(n_tiles + 2));
    CK(ctx->counters.reserve((size_t)(kMaxChunks + 1) * CNT_COUNT));
    ps.band_words = ctx->png_band_words.p;
    ps.band_bytes = ctx->png_band_bytes.p;
    ps.band_len = ctx->png_band_len.p;
    ps.band_adler = ctx->png_band_adler.p;
    ps.tile_off = ctx->png_tile_off.p;
    unsigned* flag = ctx->counters.p + (size_t)kMaxChunks * CNT_COUNT + CNT_OVERFLOW;
    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaMemsetAsync(flag, 0, sizeof(unsigned), st));
    png_encode_kernel<<<n_tiles, kPngThreads, 0, st>>>(ps);
    png_size_kernel<<<(n_tiles + 127) / 128, 128, 0, st>>>(ps);
    auto_scan_kernel<<<1, 1024, 0, st>>>(ps.tile_off, n_tiles, flag);
    std::vector<unsigned> off(n_tiles + 1);
    unsigned h_flag = 0;
    CK(cudaMemcpyAsync(off.data(), ps.tile_off, (size_t)(n_tiles + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_flag) return ctx->fail(OSMR_E_NOMEM, "more than 4 GiB of PNG data; split the batch");
    const size_t total = off[n_tiles];
    for (uint32_t t = 0; t <= n_tiles; ++t) png_offset[t] = off[t];
    if (total > png_cap) return ctx->fail(OSMR_E_NOMEM, "png_out is too small (n_tiles * osmr_png_bound(scale) always suffices)");
    CK(ctx->png_out.reserve(total + 16));
    ps.out = ctx->png_out.p;
    png_finish_kernel<<<n_tiles, kPngThreads, 0, st>>>(ps);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaMemcpyAsync(png_out, ps.out, total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_png = ms;
    ctx->stats.ms_total += ms;
    ctx->stats.kernel_launches += 4;
    return OSMR_OK;
}

int osmr_draw_tiles_png(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin, const osmr_styled_area* areas,
                        const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* png_out, size_t png_cap, uint64_t* png_offset) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!png_out || !png_offset) return ctx->fail(OSMR_E_INVALID, "null output buffer");
    if (flags & (OSMR_DRAW_OUT_RGBA | OSMR_DRAW_OUT_DEVICE)) return ctx->fail(OSMR_E_INVALID, "osmr_draw_tiles_png encodes RGB into host memory");
    int rc = batch_upload_impl(ctx, tiles, n_tiles, area_begin, areas, false, nullptr);
    if (rc) return rc;
    rc = osmr_batch_draw(ctx, canvas_rgb, flags, nullptr, nullptr);  // the RGB tiles stay in HBM (ctx->out)
    if (rc) return rc;
    return encode_png_from_device(ctx, ctx->out.p, n_tiles, (unsigned)ctx->scale, png_out, png_cap, png_offset);
} OSMR_CATCH_INT(ctx)

int osmr_rgb_to_png(osmr_ctx* ctx, const uint8_t* rgb, uint32_t n_images, uint32_t scale, uint8_t* png_out, size_t png_cap,
                    uint64_t* png_offset) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!rgb || !png_out || !png_offset) return ctx->fail(OSMR_E_INVALID, "null argument");
    if (n_images == 0) return ctx->fail(OSMR_E_INVALID, "no images");
    if (scale < 1 || scale > 8) return ctx->fail(OSMR_E_INVALID, "scale must be 1..8");
    cudaSetDevice(ctx->device);
    const size_t D = 256 * (size_t)scale, bytes = (size_t)n_images * D * D * 3;
    CK(ctx->out.reserve(bytes));
    ctx->out_bytes = bytes;
    ctx->stats = osmr_stats{};
    CK(cudaMemcpyAsync(ctx->out.p, rgb, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return encode_png_from_device(ctx, ctx->out.p, n_images, scale, png_out, png_cap, png_offset);
} OSMR_CATCH_INT(ctx)

int osmr_get_stats(osmr_ctx* ctx, osmr_stats* out) try {
    if (!ctx || !out) return OSMR_E_INVALID;
    *out = ctx->stats;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

int osmr_project_nodes(osmr_ctx* ctx, const osmr_tile* tile, int32_t* out_xy) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!ctx->ds->has_geo) return ctx->fail(OSMR_E_STATE, "osmr_set_geodata has not been called");
    if (!tile || !out_xy) return ctx->fail(OSMR_E_INVALID, "null argument");
    if (tile->scale < 1 || tile->scale > 8 || tile->zoom > 22) return ctx->fail(OSMR_E_INVALID, "bad tile");
    cudaSetDevice(ctx->device);
    if (ctx->ds->n_nodes == 0) return OSMR_OK;
    DevBuf<int2> tmp;
    CK(tmp.reserve(ctx->ds->n_nodes));
    int2* tmp_p = tmp.p;
    project_all_kernel<<<(ctx->ds->n_nodes + 255) / 256, 256, 0, ctx->stream>>>(ctx->ds->merc.p, ctx->ds->n_nodes, *tile, tmp_p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_xy, tmp.p, (size_t)ctx->ds->n_nodes * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    tmp.release();
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// ---------------------------------------------------------------------------------------------------------
// label pass
// ---------------------------------------------------------------------------------------------------------
int osmr_set_font(osmr_ctx* ctx, const void* ttf, size_t len) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!ttf || len < 64) return ctx->fail(OSMR_E_INVALID, "null or truncated font");
    ctx->lres.valid = false;
    if (!ctx->font.load((const uint8_t*)ttf, len)) return ctx->fail(OSMR_E_INVALID, "not a TrueType font with a Unicode cmap and glyf outlines");
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

int osmr_set_label_icons(osmr_ctx* ctx, const osmr_icon* icons, uint32_t n_icons) try {
    if (!ctx) return OSMR_E_INVALID;
    if (n_icons && !icons) return ctx->fail(OSMR_E_INVALID, "null icon table");
    cudaSetDevice(ctx->device);
    std::vector<DevIcon> meta(n_icons);
    std::vector<double4> px;
    ctx->label_icon_dims.assign(n_icons, osmr_host::IconDim{0, 0});
    ctx->lres.valid = false;
    for (uint32_t i = 0; i < n_icons; ++i) {
        if (!icons[i].rgba || icons[i].width == 0 || icons[i].height == 0) return ctx->fail(OSMR_E_INVALID, "empty icon");
        meta[i].w = icons[i].width;
        meta[i].h = icons[i].height;
        meta[i].off = (unsigned)px.size();
        meta[i].pad = 0;
        ctx->label_icon_dims[i] = osmr_host::IconDim{icons[i].width, icons[i].height};
        size_t n = (size_t)icons[i].width * icons[i].height;
        for (size_t k = 0; k < n; ++k) {  // RgbaColor::from_components (tile_pixels.rs:24-26)
            const uint8_t* c = icons[i].rgba + 4 * k;
            volatile double opacity = (double)c[3] / 255.0;
            volatile double r = (double)c[0] / 255.0, g = (double)c[1] / 255.0, b = (double)c[2] / 255.0;
            volatile double pr = opacity * r, pg = opacity * g, pb = opacity * b;
            double4 v;
            v.x = pr;
            v.y = pg;
            v.z = pb;
            v.w = opacity;
            px.push_back(v);
        }
    }
    if (px.size() >= 0x3fffffffu) return ctx->fail(OSMR_E_INVALID, "label icon table too large");
    CK(ctx->label_icons.reserve(n_icons + 1));
    CK(ctx->label_icon_px.reserve(px.size() + 1));
    if (n_icons) CK(cudaMemcpyAsync(ctx->label_icons.p, meta.data(), n_icons * sizeof(DevIcon), cudaMemcpyHostToDevice, ctx->stream));
    if (!px.empty()) CK(cudaMemcpyAsync(ctx->label_icon_px.p, px.data(), px.size() * sizeof(double4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

int osmr_set_label_styles(osmr_ctx* ctx, const osmr_label_style* styles, uint32_t n_styles, const char* strings, size_t strings_len) try {
    if (!ctx) return OSMR_E_INVALID;
    if (n_styles && !styles) return ctx->fail(OSMR_E_INVALID, "null label style table");
    std::vector<osmr_host::LabelStyleHost> tmp(n_styles);
    for (uint32_t i = 0; i < n_styles; ++i) {
        tmp[i].s = styles[i];
        if (styles[i].flags & OSMR_LSTYLE_TEXT) {
            if (!strings || (uint64_t)styles[i].text_key_off + styles[i].text_key_len > strings_len)
                return ctx->fail(OSMR_E_INVALID, "label style text key out of range");
            tmp[i].key.assign(strings + styles[i].text_key_off, styles[i].text_key_len);
        }
        if (styles[i].text_position > OSMR_TEXT_POS_LINE) return ctx->fail(OSMR_E_INVALID, "bad text position");
    }
    ctx->label_styles.swap(tmp);
    ctx->lres.valid = false;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// ---- host layout (round-1 path; still the fallback of the device layout and what debug key "label_host" selects) ----
static int labels_via_host(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* label_begin, const osmr_label* labels) {
    int rc = OSMR_OK;
    (void)rc;
    const int D = 256 * ctx->scale, E = 3 * D;
    // ---- host half: layout (string / font / heap work, as the reference does it on the CPU), tiles in parallel ----
    const auto t_host0 = std::chrono::steady_clock::now();
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (label_begin[t + 1] < label_begin[t]) return ctx->fail(OSMR_E_INVALID, "label_begin must be non-decreasing");
        if (label_begin[t + 1] > label_begin[t] && !labels) return ctx->fail(OSMR_E_INVALID, "null label list");
    }
    osmr_host::LayoutEnv env{&ctx->ds->h_view, &ctx->font, &ctx->label_styles, &ctx->label_icon_dims};
    // work items: runs of <= kLabelRun consecutive labels of a tile (the layout of a label does not depend on the others)
    constexpr uint32_t kLabelRun = 96;
    std::vector<LabelWorkItem>& items = ctx->label_items;
    size_t n_items = 0;
    for (uint32_t t = 0; t < n_tiles; ++t)
        for (uint32_t f = label_begin[t]; f < label_begin[t + 1]; f += kLabelRun) ++n_items;
    if (items.size() < n_items) items.resize(n_items);
    {
        size_t k = 0;
        for (uint32_t t = 0; t < n_tiles; ++t)
            for (uint32_t f = label_begin[t]; f < label_begin[t + 1]; f += kLabelRun) {
                items[k].tile = t;
                items[k].first = f;
                items[k].count = std::min(kLabelRun, label_begin[t + 1] - f);
                ++k;
            }
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const unsigned n_threads = (unsigned)std::max<size_t>(1, std::min<size_t>({hw ? hw : 1u, n_items, ctx->label_threads}));
    auto run_parallel = [&](auto&& body) {  // body(item index), items handed out dynamically
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (size_t k; (k = next.fetch_add(1)) < n_items;) body(k);
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < n_threads; ++i) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
    };
    std::atomic<bool> bad{false};
    run_parallel([&](size_t k) {
        LabelWorkItem& it = items[k];
        it.recs.clear();
        it.segs.clear();
        if (bad.load(std::memory_order_relaxed)) return;
        if (!osmr_host::layout_tile(env, tiles[it.tile], labels + it.first, it.count, it.recs, it.segs)) bad.store(true);
    });
    if (bad.load()) return ctx->fail(OSMR_E_INVALID, "label references an entity, style or icon that does not exist");
    // batch assembly: global segment indices, coverage storage and the work list of label_cover_kernel
    std::vector<osmr_host::LabelRec> recs;
    std::vector<unsigned> cover_list;
    std::vector<unsigned> lbegin(n_tiles + 1, 0);
    unsigned long long cells = 0, n_rows = 0;
    size_t n_segs = 0;
    {
        size_t nr = 0;
        for (size_t k = 0; k < n_items; ++k) {
            items[k].seg_base = n_segs;
            nr += items[k].recs.size();
            n_segs += items[k].segs.size();
        }
        if (n_segs >= 0xffffffffull || nr >= 0xffffffffull) return ctx->fail(OSMR_E_NOMEM, "label batch too large; split the batch");
        recs.reserve(nr);
    }
    cudaSetDevice(ctx->device);
    CK(ctx->h_label_segs.reserve(n_segs + 1));
    osmr_host::Seg* const seg_stage = ctx->h_label_segs.p;
    run_parallel([&](size_t k) {
        const LabelWorkItem& it = items[k];
        if (!it.segs.empty()) memcpy(seg_stage + it.seg_base, it.segs.data(), it.segs.size() * sizeof(osmr_host::Seg));
    });
    for (size_t k = 0; k < n_items; ++k) {
        const LabelWorkItem& it = items[k];
        for (osmr_host::LabelRec r : it.recs) {
            r.seg_begin += (unsigned)it.seg_base;
            // rows outside the label canvas cannot collide or draw; columns stay complete (the sweep is a prefix sum)
            r.ry0 = std::max(r.by0, -D);
            long long rows = (long long)std::min(r.by1, 2 * D - 1) - r.ry0 + 1;
            long long cols = (long long)r.bx1 - r.bx0 + 1;
            const bool touches = r.seg_count && rows > 0 && cols > 0 && r.bx1 >= -D && r.bx0 <= 2 * D - 1;
            r.rows = touches ? (int)rows : 0;
            r.width = touches ? (int)cols : 0;
            r.row_first = (unsigned)n_rows;
            r.n_ranges = 0;
            r.range_off = 0;
            r.cell_off = cells;
            if (r.icon < 0 && !touches) continue;  // cannot draw, claim or collide
            if (touches) {
                if (cols > (1 << 20)) return ctx->fail(OSMR_E_NOMEM, "label text wider than 2^20 pixels");
                cover_list.push_back((unsigned)recs.size());
                n_rows += (unsigned long long)rows;
                cells += (unsigned long long)rows * cols;
            }
            recs.push_back(r);
        }
        lbegin[it.tile + 1] = (unsigned)recs.size();
    }
    for (uint32_t t = 0; t < n_tiles; ++t) lbegin[t + 1] = std::max(lbegin[t + 1], lbegin[t]);  // tiles without labels
    if (cells > (1ull << 31) || n_rows >= 0x7fffffffull) return ctx->fail(OSMR_E_NOMEM, "label scratch too large; split the batch");
    ctx->stats_label_layout_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    static_assert(sizeof(osmr_host::LabelRec) == sizeof(DevLabel) && sizeof(osmr_host::Seg) == sizeof(DevSeg), "label wire layout");
    // ---- device half ----
    CK(ctx->d_labels.reserve(recs.size() + 1));
    CK(ctx->d_label_segs.reserve(n_segs + 1));
    CK(ctx->d_cover_list.reserve(cover_list.size() + 1));
    CK(ctx->d_cover_cursor.reserve(2 * kMaxChunks));
    CK(ctx->d_label_begin.reserve(n_tiles + 1));
    CK(ctx->label_occ.reserve((size_t)n_tiles * ((size_t)E * E / 32)));
    CK(ctx->label_acc.reserve(2 * (size_t)cells + 2));
    CK(ctx->label_row_keys.reserve(2 * (size_t)n_rows + 2));
    CK(ctx->label_plane.reserve((size_t)n_tiles * D * D));
    CK(ctx->label_pmask.reserve((size_t)n_tiles * (size_t)(D * D / 32) + 1));
    CK(cudaEventRecord(ctx->ev_label0, ctx->stream));
    if (!recs.empty()) CK(cudaMemcpyAsync(ctx->d_labels.p, recs.data(), recs.size() * sizeof(DevLabel), cudaMemcpyHostToDevice, ctx->stream));
    if (n_segs) CK(cudaMemcpyAsync(ctx->d_label_segs.p, seg_stage, n_segs * sizeof(DevSeg), cudaMemcpyHostToDevice, ctx->stream));
    if (!cover_list.empty())
        CK(cudaMemcpyAsync(ctx->d_cover_list.p, cover_list.data(), cover_list.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_label_begin.p, lbegin.data(), (n_tiles + 1) * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_cover_cursor.p, 0, 2 * sizeof(unsigned), ctx->stream));
    LabelScene ls{};
    ls.labels = ctx->d_labels.p;
    ls.label_begin = ctx->d_label_begin.p;
    ls.segs = ctx->d_label_segs.p;
    ls.cover_list = ctx->d_cover_list.p;
    ls.n_cover = (unsigned)cover_list.size();
    ls.icons = ctx->label_icons.p;
    ls.occ = ctx->label_occ.p;
    ls.acc_a = ctx->label_acc.p;
    ls.acc_s = ctx->label_acc.p + cells;
    ls.kmin = ctx->label_row_keys.p;
    ls.kmax = ctx->label_row_keys.p + n_rows;
    ls.plane = ctx->label_plane.p;
    ls.pmask = ctx->label_pmask.p;
    ls.D = D;
    ls.cover_cursor = ctx->d_cover_cursor.p;
    ls.err_flag = ctx->d_cover_cursor.p + 1;
    if (ls.n_cover) {
        label_cover_kernel<<<std::min<unsigned>(ls.n_cover, (unsigned)ctx->num_sms * kCovCtasPerSm), 32, 0, ctx->stream>>>(ls);
        CK(cudaGetLastError());
    }
    label_commit_kernel<<<n_tiles, kLabelThreads, 0, ctx->stream>>>(ls);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_label1, ctx->stream));
    unsigned cover_err = 0;
    CK(cudaMemcpyAsync(&cover_err, ctx->d_cover_cursor.p + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // recs / segs live on this stack frame
    if (cover_err) return ctx->fail(OSMR_E_CUDA, "internal error: a label coverage term fell outside its proven window");
    CK(cudaEventElapsedTime(&ctx->stats_label_device_ms, ctx->ev_label0, ctx->ev_label1));
    return OSMR_OK;
}

// ---- device layout (osmr_labels_dev.cuh) ----
// Resident tables, built once per (dataset, font, label style table): the text run of every entity under every text key the
// styles use (tag lookup + cmap + hmtx + kern = the string / font-table work of text_placer.rs:24-58,170-209), the glyph
// outlines, the interned label styles.
static int build_label_tables(osmr_ctx* ctx) {
    auto& R = ctx->lres;
    R.valid = false;
    const osmr_host::BinView& g = ctx->ds->h_view;
    const osmr_host::TrueType& font = ctx->font;
    // distinct text keys
    R.keys.clear();
    std::vector<DevLabelStyle> dstyles(ctx->label_styles.size());
    for (size_t i = 0; i < ctx->label_styles.size(); ++i) {
        const osmr_host::LabelStyleHost& st = ctx->label_styles[i];
        DevLabelStyle d{};
        d.icon = st.s.icon;
        d.flags = st.s.flags;
        d.key = -1;
        if (st.s.flags & OSMR_LSTYLE_TEXT) {
            size_t k = 0;
            while (k < R.keys.size() && R.keys[k] != st.key) ++k;
            if (k == R.keys.size()) R.keys.push_back(st.key);
            d.key = (int)k;
        }
        const uint8_t* c = st.s.text_color;
        d.rgb = (st.s.flags & OSMR_LSTYLE_TEXT_COLOR) ? ((unsigned)c[0] | ((unsigned)c[1] << 8) | ((unsigned)c[2] << 16)) : 0u;
        d.text_position = st.s.text_position;
        d.font_size = st.s.font_size;
        dstyles[i] = d;
    }
    const uint32_t n_nodes = g.n_nodes, n_ways = g.n_ways, n_mps = g.n_mps;
    R.n_keys = (unsigned)R.keys.size();
    R.way_base = n_nodes;
    R.mp_base = n_nodes + n_ways;
    R.ent_total = n_nodes + n_ways + n_mps;
    if ((uint64_t)R.n_keys * R.ent_total >= 0xffffffffull) return ctx->fail(OSMR_E_NOMEM, "label text table too large");
    std::vector<unsigned> text_id((size_t)R.n_keys * R.ent_total, 0xffffffffu);
    std::vector<unsigned> text_begin(1, 0u);
    std::vector<DevGlyphRec> glyphs;
    std::vector<unsigned> glyph_vbegin(1, 0u);
    std::vector<DevVertex> verts;
    std::unordered_map<int, int> slot_of_glyph;
    std::unordered_map<std::string, unsigned> text_of_string;
    std::vector<uint32_t> chars;
    R.way_has_text.assign(n_ways, 0);
    auto intern_text = [&](const char* val, size_t len) -> unsigned {
        std::string key(val, len);
        auto it = text_of_string.find(key);
        if (it != text_of_string.end()) return it->second;
        osmr_host::decode_utf8(val, len, chars);
        int prev = -1;
        for (uint32_t cp : chars) {
            const int gi = font.glyph_index(cp);
            DevGlyphRec r;
            auto sit = slot_of_glyph.find(gi);
            if (sit == slot_of_glyph.end()) {
                int slot = -1;
                const std::vector<osmr_host::GlyphVertex>* shape = font.shape(gi);
                if (shape) {
                    slot = (int)glyph_vbegin.size() - 1;
                    for (const osmr_host::GlyphVertex& v : *shape) verts.push_back(DevVertex{v.x, v.y, v.cx, v.cy, (int)v.type});
                    glyph_vbegin.push_back((unsigned)verts.size());
                }
                sit = slot_of_glyph.emplace(gi, slot).first;
            }
            r.slot = sit->second;
            r.advance = font.advance(gi);
            r.kern = prev >= 0 ? font.kerning(prev, gi) : 0;
            r.ws = osmr_host::is_ws(cp) ? 1u : 0u;
            glyphs.push_back(r);
            prev = gi;
        }
        text_begin.push_back((unsigned)glyphs.size());
        const unsigned id = (unsigned)text_begin.size() - 2;
        text_of_string.emplace(std::move(key), id);
        return id;
    };
    for (unsigned k = 0; k < R.n_keys; ++k) {
        const std::string& key = R.keys[k];
        unsigned* row = text_id.data() + (size_t)k * R.ent_total;
        const char* val;
        size_t vlen;
        for (uint32_t i = 0; i < n_nodes; ++i) {
            const uint32_t tl = osmr_host::BinView::u32(g.nodes + (size_t)i * 32 + 28);
            if (!tl) continue;
            if (g.tag(osmr_host::BinView::u32(g.nodes + (size_t)i * 32 + 24), tl, key.data(), key.size(), val, vlen)) row[i] = intern_text(val, vlen);
        }
        for (uint32_t i = 0; i < n_ways; ++i) {
            const uint32_t tl = osmr_host::BinView::u32(g.ways + (size_t)i * 24 + 20);
            if (!tl) continue;
            if (g.tag(osmr_host::BinView::u32(g.ways + (size_t)i * 24 + 16), tl, key.data(), key.size(), val, vlen)) {
                row[R.way_base + i] = intern_text(val, vlen);
                R.way_has_text[i] = 1;
            }
        }
        for (uint32_t i = 0; i < n_mps; ++i) {
            const uint32_t tl = osmr_host::BinView::u32(g.mps + (size_t)i * 24 + 20);
            if (!tl) continue;
            if (g.tag(osmr_host::BinView::u32(g.mps + (size_t)i * 24 + 16), tl, key.data(), key.size(), val, vlen)) row[R.mp_base + i] = intern_text(val, vlen);
        }
    }
    R.n_texts = (unsigned)text_begin.size() - 1;
    cudaStream_t st = ctx->stream;
    CK(R.styles.reserve(dstyles.size() + 1));
    CK(R.text_id.reserve(text_id.size() + 1));
    CK(R.text_begin.reserve(text_begin.size() + 1));
    CK(R.glyphs.reserve(glyphs.size() + 1));
    CK(R.glyph_vbegin.reserve(glyph_vbegin.size() + 1));
    {  // bound of label_precull_kernel: outline extents (twice: a glyph on a way turns around its centre), advances, ascent, descent
        long long max_coord = 0, max_adv = 0;
        for (const DevVertex& v : verts)
            max_coord = std::max<long long>(max_coord, std::max(std::max(std::abs((long long)v.x), std::abs((long long)v.y)),
                                                                std::max(std::abs((long long)v.cx), std::abs((long long)v.cy))));
        for (const DevGlyphRec& g : glyphs) max_adv = std::max<long long>(max_adv, std::abs((long long)g.advance) + std::abs((long long)g.kern));
        const long long reach = 2 * max_coord + max_adv + std::abs((long long)ctx->font.ascent()) + std::abs((long long)ctx->font.descent());
        R.font_reach = (int)std::min<long long>(reach, 0x3fffffff);
    }
    CK(R.verts.reserve(verts.size() + 1));
    auto up = [&](void* d, const void* h, size_t bytes) { return bytes ? cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
    CK(up(R.styles.p, dstyles.data(), dstyles.size() * sizeof(DevLabelStyle)));
    CK(up(R.text_id.p, text_id.data(), text_id.size() * 4));
    CK(up(R.text_begin.p, text_begin.data(), text_begin.size() * 4));
    CK(up(R.glyphs.p, glyphs.data(), glyphs.size() * sizeof(DevGlyphRec)));
    CK(up(R.glyph_vbegin.p, glyph_vbegin.data(), glyph_vbegin.size() * 4));
    CK(up(R.verts.p, verts.data(), verts.size() * sizeof(DevVertex)));
    CK(cudaStreamSynchronize(st));
    for (auto& a : R.angle) a.set = false;
    R.valid = true;
    return OSMR_OK;
}

// Per zoom: sin(-angle) / cos(-angle) of every segment of every way that has a text, angle = atan2(dy, dx) of the way's INTEGER
// pixel differences in its drawing orientation (text_placer.rs:60-99,265-296) -- the platform-libm half of text along a way, from
// glibc like the reference's.  The differences are the same in every tile of the zoom (see osmr_labels_dev.cuh), so tile (0, 0)
// stands for all of them.
static int build_angle_table(osmr_ctx* ctx, unsigned zoom, int scale) {
    auto& R = ctx->lres;
    auto& A = R.angle[zoom];
    A.set = false;
    const uint32_t n_ways = ctx->ds->h_view.n_ways;
    if (ctx->ds->h_merc.size() != ctx->ds->n_nodes) return ctx->fail(OSMR_E_STATE, "no host copy of the Mercator factors (debug key device_merc)");
    std::vector<unsigned> off(n_ways, 0xffffffffu);
    std::vector<double2> sc;
    const double dim = (double)(unsigned)(256u * (1u << zoom)), fscale = (double)scale;
    std::vector<osmr_host::IPoint> pts;
    for (uint32_t w = 0; w < n_ways; ++w) {
        if (!R.way_has_text[w]) continue;
        const uint32_t o = osmr_host::BinView::u32(ctx->ds->h_view.ways + (size_t)w * 24 + 8), len = osmr_host::BinView::u32(ctx->ds->h_view.ways + (size_t)w * 24 + 12);
        if (len < 2) continue;
        pts.clear();
        for (uint32_t i = 0; i < len; ++i) {
            const double2 m = ctx->ds->h_merc[ctx->ds->h_view.int_at(o + i)];
            volatile double x = m.x * dim, y = m.y * dim;  // project_point with the tile origin at 0 (exact IEEE steps)
            volatile double xs = x * fscale, ys = y * fscale;
            pts.push_back(osmr_host::IPoint{osmr_host::f64_as_i32(std::round(xs)), osmr_host::f64_as_i32(std::round(ys))});
        }
        if (pts.front().x > pts.back().x) std::reverse(pts.begin(), pts.end());
        off[w] = (unsigned)sc.size();
        for (uint32_t i = 0; i + 1 < len; ++i) {
            const double angle = std::atan2((double)osmr_host::wsub(pts[i + 1].y, pts[i].y), (double)osmr_host::wsub(pts[i + 1].x, pts[i].x));
            sc.push_back(make_double2(std::sin(-angle), std::cos(-angle)));
        }
    }
    CK(A.off.reserve(off.size() + 1));
    CK(A.sc.reserve(sc.size() + 1));
    if (!off.empty()) CK(cudaMemcpyAsync(A.off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!sc.empty()) CK(cudaMemcpyAsync(A.sc.p, sc.data(), sc.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    A.scale = scale;
    A.set = true;
    return OSMR_OK;
}

static bool label_device_path_allowed(const osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles) {
    if (ctx->label_host_only || ctx->ds->h_merc.size() != ctx->ds->n_nodes) return false;
    const uint32_t scale = tiles[0].scale;
    if (scale != 1 && scale != 2 && scale != 4 && scale != 8) return false;  // `* scale` must be exact (osmr_labels_dev.cuh)
    for (uint32_t t = 0; t < n_tiles; ++t)
        if (tiles[t].zoom > 18) return false;
    return true;
}

// Enqueues the whole label pass on the compute stream (nothing here waits for the device): layout kernels, glyph coverage,
// greedy collisions -> ctx->label_plane.  The counters land in page-locked memory; label_device_judge reads them after the
// draw has synchronised the stream.
static int label_device_enqueue(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* label_begin, const osmr_label* labels,
                                bool resident, bool chunked) {
    auto& R = ctx->lres;
    const int D = 256 * ctx->scale, E = 3 * D;
    if (!R.valid) {
        int rc = build_label_tables(ctx);
        if (rc) return rc;
    }
    for (uint32_t t = 0; t < n_tiles; ++t) {
        auto& A = R.angle[tiles[t].zoom];
        if (!A.set || A.scale != ctx->scale) {
            int rc = build_angle_table(ctx, tiles[t].zoom, ctx->scale);
            if (rc) return rc;
        }
    }
    const uint32_t n_labels = label_begin[n_tiles];
    // The label pass has its own stream: it only meets the area passes at raster_kernel's export, so its kernels (long, latency
    // bound, few warps per SM) run beside plan / geometry / fill rows / bin / line cover instead of in front of them.
    cudaStream_t st = ctx->label_stream;
    CK(cudaEventRecord(ctx->label_go, ctx->stream));  // (the batch description -- tiles -- was uploaded on the compute stream)
    CK(cudaStreamWaitEvent(st, ctx->label_go, 0));
    // first guesses of the bump-allocated scratch (grown by label_device_judge)
    if (!ctx->l_places_cap) ctx->l_places_cap = 1u << 18;
    if (!ctx->l_segs_cap) ctx->l_segs_cap = 1u << 21;
    if (!ctx->l_verts_cap) ctx->l_verts_cap = 1u << 20;
    if (!ctx->l_curves_cap) ctx->l_curves_cap = 1u << 19;
    if (!ctx->l_rowrecs_cap) ctx->l_rowrecs_cap = 1u << 19;
    if (!ctx->l_cells_cap) ctx->l_cells_cap = 1u << 23;
    if (!ctx->l_ring_cap) ctx->l_ring_cap = 1u << 16;
    if (!ctx->l_heap_slots) ctx->l_heap_slots = 256;
    CK(ctx->d_label_list.reserve((size_t)n_labels + 1));
    CK(ctx->d_label_begin.reserve(n_tiles + 1));
    CK(ctx->l_act.reserve((size_t)n_labels + 1));
    CK(ctx->l_predead.reserve((size_t)n_labels + 1));
    CK(ctx->l_place.reserve((size_t)n_labels + 1));
    CK(ctx->d_labels.reserve((size_t)n_labels + 1));
    CK(ctx->l_act_cnt.reserve(n_tiles + 1));
    CK(ctx->l_gplace.reserve(ctx->l_places_cap));
    CK(ctx->l_place_vinst.reserve(ctx->l_places_cap));
    const unsigned n_scan_blocks = (unsigned)((ctx->l_verts_cap + kScanBlock) / kScanBlock) + 1u;
    CK(ctx->l_vinst_place.reserve(ctx->l_verts_cap + 8));
    CK(ctx->l_vcnt.reserve(ctx->l_verts_cap + 8));
    CK(ctx->l_vbox.reserve(ctx->l_verts_cap + 8));
    CK(ctx->l_curve_list.reserve(ctx->l_curves_cap + 8));
    CK(ctx->l_curve_codes.reserve((kCurveLeafCap / 4u) * ctx->l_curves_cap + 8));
    CK(ctx->l_curve_deep.reserve(ctx->l_curves_cap + 8));
    CK(ctx->l_curve_root.reserve(ctx->l_curves_cap + 8));
    CK(ctx->l_scan_blocks.reserve(n_scan_blocks + 8));
    CK(ctx->d_label_segs.reserve(ctx->l_segs_cap));
    CK(ctx->d_cover_list.reserve((size_t)n_labels + 1));
    CK(ctx->d_cover_cursor.reserve(2 * kMaxChunks));
    CK(ctx->label_row_keys.reserve(2 * ctx->l_rowrecs_cap + 2));
    CK(ctx->label_acc.reserve(2 * ctx->l_cells_cap + 2));
    CK(ctx->l_ring_pts.reserve(ctx->l_ring_cap));
    CK(ctx->l_heap.reserve(ctx->l_heap_slots * (size_t)kPolyHeapCap * sizeof(PolyCell)));
    CK(ctx->label_occ.reserve((size_t)n_tiles * ((size_t)E * E / 32)));
    CK(ctx->label_plane.reserve((size_t)n_tiles * D * D));
    CK(ctx->label_pmask.reserve((size_t)n_tiles * (size_t)(D * D / 32) + 1));
    ctx->l_places_used_cap = ctx->l_places_cap;
    ctx->l_segs_used_cap = ctx->l_segs_cap;
    ctx->l_rowrecs_used_cap = ctx->l_rowrecs_cap;
    ctx->l_cells_used_cap = ctx->l_cells_cap;
    ctx->l_ring_used_cap = ctx->l_ring_cap;
    ctx->l_heap_used_slots = ctx->l_heap_slots;
    ctx->l_verts_used_cap = ctx->l_verts_cap;
    ctx->l_curves_used_cap = ctx->l_curves_cap;
    CK(cudaEventRecord(ctx->ev_label0, st));
    CK(ctx->l_counters.reserve((size_t)kMaxChunks * LCNT_COUNT));
    CK(ctx->h_lcnt.reserve((size_t)kMaxChunks * LCNT_COUNT));
    // The label pass of a host-output call runs in TWO halves on two streams with a scratch set each: a draw chunk waits only for
    // the label half above it, so the first tiles travel while the second half is still computed.  More chunks lose: label_select /
    // label_layout / label_commit are one CTA per tile and serial inside a tile (greedy collisions, polylabel), so every chunk
    // pays the latency of its slowest tile, and thirteen kernels per chunk have thirteen tails.  Measured end to end on the C2 batch
    // (B200, ms per call): 1 chunk 16.1, 2 chunks 15.8, 3 chunks 16.4, the draw's own 5-chunk schedule 17.7 (one stream: 20.0).
    // Resident output: one chunk.  Debug key "label_chunks" = n forces n equal chunks.
    {
        unsigned sizes[kMaxChunks];
        unsigned n = 1u;
        sizes[0] = n_tiles;
        if (chunked && !ctx->label_chunks && n_tiles >= 256) {
            n = 2;
            sizes[0] = n_tiles / 2;
            sizes[1] = n_tiles - sizes[0];
        }
        if (ctx->label_chunks) {
            n = std::min<unsigned>(ctx->label_chunks, n_tiles);
            for (unsigned i = 0; i < n; ++i) sizes[i] = n_tiles / n + (i < n_tiles % n ? 1u : 0u);
        }
        ctx->n_lchunks = n;
        unsigned tb = 0;
        for (unsigned i = 0; i < n; ++i) {
            ctx->lchunk_tb[i] = tb;
            ctx->lchunk_tc[i] = sizes[i];
            tb += sizes[i];
        }
        if (tb != n_tiles) return ctx->fail(OSMR_E_STATE, "internal error: bad label chunk plan");
    }
    const bool two = ctx->n_lchunks > 1;
    cudaStream_t st2 = ctx->label_stream2;
    auto& B = ctx->lscrB;
    if (two) {
        CK(B.gplace.reserve(ctx->l_places_cap));
        CK(B.place_vinst.reserve(ctx->l_places_cap));
        CK(B.vinst_place.reserve(ctx->l_verts_cap + 8));
        CK(B.vcnt.reserve(ctx->l_verts_cap + 8));
        CK(B.vbox.reserve(ctx->l_verts_cap + 8));
        CK(B.curve_list.reserve(ctx->l_curves_cap + 8));
        CK(B.curve_codes.reserve((kCurveLeafCap / 4u) * ctx->l_curves_cap + 8));
        CK(B.curve_deep.reserve(ctx->l_curves_cap + 8));
        CK(B.curve_root.reserve(ctx->l_curves_cap + 8));
        CK(B.scan_blocks.reserve(n_scan_blocks + 8));
        CK(B.segs.reserve(ctx->l_segs_cap));
        CK(B.cover_list.reserve((size_t)n_labels + 1));
        CK(B.row_keys.reserve(2 * ctx->l_rowrecs_cap + 2));
        CK(B.acc.reserve(2 * ctx->l_cells_cap + 2));
        CK(B.ring_pts.reserve(ctx->l_ring_cap));
        CK(B.heap.reserve(ctx->l_heap_slots * (size_t)kPolyHeapCap * sizeof(PolyCell)));
    }
    if (!resident) {
        CK(cudaMemcpyAsync(ctx->d_label_begin.p, label_begin, (size_t)(n_tiles + 1) * 4, cudaMemcpyHostToDevice, st));
        // the label lists: the first chunk's right here, the others behind the styled areas on the copy stream
        for (unsigned i = 0; i < ctx->n_lchunks; ++i) {
            const uint32_t l0 = label_begin[ctx->lchunk_tb[i]], l1 = label_begin[ctx->lchunk_tb[i] + ctx->lchunk_tc[i]];
            if (l1 == l0) continue;
            cudaStream_t cs = i == 0 ? st : ctx->copy_stream;
            CK(cudaMemcpyAsync(ctx->d_label_list.p + l0, labels + l0, (size_t)(l1 - l0) * sizeof(osmr_label), cudaMemcpyHostToDevice, cs));
            if (i) CK(cudaEventRecord(ctx->lchunk_up[i], ctx->copy_stream));
        }
    }
    CK(cudaMemsetAsync(ctx->l_counters.p, 0, (size_t)kMaxChunks * LCNT_COUNT * sizeof(unsigned), st));
    CK(cudaMemsetAsync(ctx->d_cover_cursor.p, 0, 2 * sizeof(unsigned) * kMaxChunks, st));
    if (two) {  // the second stream starts behind the batch description and the cleared counters
        CK(cudaEventRecord(ctx->label_prep, st));
        CK(cudaStreamWaitEvent(st2, ctx->label_prep, 0));
    }
    Scene s{};
    s.merc = ctx->ds->merc.p;
    s.ways = ctx->ds->ways.p;
    s.polys = ctx->ds->polys.p;
    s.mps = ctx->ds->mps.p;
    s.ints = ctx->ds->ints.p;
    s.n_nodes = ctx->ds->n_nodes;
    s.n_ways = ctx->ds->n_ways;
    s.n_polys = ctx->ds->n_polys;
    s.n_mps = ctx->ds->n_mps;
    s.n_ints = ctx->ds->n_ints;
    s.way_box = ctx->ds->way_box.p;
    s.mp_box = ctx->ds->mp_box.p;
    s.tiles = ctx->tiles.p;
    s.n_tiles = n_tiles;
    s.D = D;
    s.scale = ctx->scale;
    LabelDev ld{};
    ld.styles = R.styles.p;
    ld.n_styles = (unsigned)ctx->label_styles.size();
    ld.text_id = R.text_id.p;
    ld.n_keys = R.n_keys;
    ld.ent_total = R.ent_total;
    ld.way_base = R.way_base;
    ld.mp_base = R.mp_base;
    ld.text_begin = R.text_begin.p;
    ld.n_texts = R.n_texts;
    ld.glyphs = R.glyphs.p;
    ld.glyph_vbegin = R.glyph_vbegin.p;
    ld.verts = R.verts.p;
    ld.ascent = ctx->font.ascent();
    ld.descent = ctx->font.descent();
    ld.line_gap = ctx->font.line_gap();
    for (int z = 0; z < 19; ++z) {
        ld.way_angle_off[z] = (R.angle[z].set && R.angle[z].scale == ctx->scale) ? R.angle[z].off.p : nullptr;
        ld.sincos[z] = R.angle[z].sc.p;
    }
    ld.icons = ctx->label_icons.p;
    ld.n_icons = (unsigned)ctx->label_icon_dims.size();
    ld.label_begin = ctx->d_label_begin.p;
    ld.labels = ctx->d_label_list.p;
    ld.n_tiles = n_tiles;
    ld.act = ctx->l_act.p;
    ld.act_cnt = ctx->l_act_cnt.p;
    ld.place = ctx->l_place.p;
    ld.gplace = ctx->l_gplace.p;
    ld.place_vinst = ctx->l_place_vinst.p;
    ld.vinst_place = ctx->l_vinst_place.p;
    ld.vcnt = ctx->l_vcnt.p;
    ld.vbox = ctx->l_vbox.p;
    ld.curve_list = ctx->l_curve_list.p;
    ld.curve_codes = ctx->l_curve_codes.p;
    ld.curve_deep = ctx->l_curve_deep.p;
    ld.leaf_cap = ctx->curve_leaf_cap ? std::min(ctx->curve_leaf_cap, kCurveLeafCap) : kCurveLeafCap;
    ld.curves_cap = (unsigned)std::min<size_t>(ctx->l_curves_cap, 0xfffffff0u);
    ld.curve_root = ctx->l_curve_root.p;
    ld.verts_cap = (unsigned)std::min<size_t>(ctx->l_verts_cap, 0xfffffff0u);
    ld.scan_blocks = ctx->l_scan_blocks.p;
    ld.n_scan_blocks = n_scan_blocks;
    ld.gplace_cap = (unsigned)std::min<size_t>(ctx->l_places_cap, 0xfffffff0u);
    ld.segs = ctx->d_label_segs.p;
    ld.segs_cap = (unsigned)std::min<size_t>(ctx->l_segs_cap, 0xfffffff0u);
    ld.out_labels = ctx->d_labels.p;
    ld.cover_list = ctx->d_cover_list.p;
    ld.rowrecs_cap = (unsigned)std::min<size_t>(ctx->l_rowrecs_cap, 0x7ffffff0u);
    ld.cells_cap = ctx->l_cells_cap;
    ld.ring_pts = ctx->l_ring_pts.p;
    ld.ring_cap = (unsigned)std::min<size_t>(ctx->l_ring_cap, 0xfffffff0u);
    ld.heap = ctx->l_heap.p;
    ld.heap_slots = (unsigned)ctx->l_heap_slots;
    ld.counters = ctx->l_counters.p;
    ld.cull = ctx->label_cull ? 1u : 0u;
    ld.predead = ctx->l_predead.p;
    ld.font_reach = R.font_reach;
    const unsigned wide = (unsigned)ctx->num_sms * 8u;
    const Scene s_all = s;
    const LabelDev ld_all = ld;
    cudaStream_t const st_a = st;
    for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch) {
        const unsigned tb = ctx->lchunk_tb[ch], tc = ctx->lchunk_tc[ch];
        const bool setb = two && (ch & 1u);
        st = setb ? st2 : st_a;
        if (!resident && ch && label_begin[tb + tc] > label_begin[tb]) CK(cudaStreamWaitEvent(st, ctx->lchunk_up[ch], 0));
        s = s_all;
        ld = ld_all;
        if (setb) {
            ld.gplace = B.gplace.p;
            ld.place_vinst = B.place_vinst.p;
            ld.vinst_place = B.vinst_place.p;
            ld.vcnt = B.vcnt.p;
            ld.vbox = B.vbox.p;
            ld.curve_list = B.curve_list.p;
            ld.curve_codes = B.curve_codes.p;
            ld.curve_deep = B.curve_deep.p;
            ld.curve_root = B.curve_root.p;
            ld.scan_blocks = B.scan_blocks.p;
            ld.segs = B.segs.p;
            ld.cover_list = B.cover_list.p;
            ld.ring_pts = B.ring_pts.p;
            ld.heap = B.heap.p;
        }
        s.tiles = s_all.tiles + tb;
        s.n_tiles = tc;
        ld.label_begin = ld_all.label_begin + tb;
        ld.act_cnt = ld_all.act_cnt + tb;
        ld.n_tiles = tc;
        ld.counters = ctx->l_counters.p + (size_t)ch * LCNT_COUNT;
        label_select_kernel<<<tc, kLabelSelThreads, 0, st>>>(s, ld);
        label_precull_kernel<<<tc, kCullThreads, 0, st>>>(s, ld);
        label_layout_kernel<<<tc, kLayoutThreads, 0, st>>>(s, ld);
        label_cull_kernel<<<tc, kCullThreads, 0, st>>>(s, ld);
        label_vfill_kernel<<<wide, 128, 0, st>>>(ld);
        label_vline_count_kernel<<<wide, 128, 0, st>>>(ld);
        label_curve_count_kernel<<<wide, 128, 0, st>>>(ld);
        label_scan_sums_kernel<<<n_scan_blocks, 256, 0, st>>>(ld);
        auto_scan_kernel<<<1, 1024, 0, st>>>(ld.scan_blocks, n_scan_blocks, ld.counters + LCNT_SCAN_OVF);
        label_scan_apply_kernel<<<n_scan_blocks, 256, 0, st>>>(ld);
        label_finish_kernel<<<tc, 128, 0, st>>>(s, ld);
        label_vline_write_kernel<<<wide, 128, 0, st>>>(ld);
        label_curve_expand_kernel<<<(unsigned)ctx->num_sms * 8u, 256, 0, st>>>(ld);
        CK(cudaGetLastError());
        LabelScene ls{};
        ls.labels = ctx->d_labels.p;
        ls.label_begin = ctx->d_label_begin.p + tb;
        ls.segs = ld.segs;
        ls.cover_list = ld.cover_list;
        ls.cover_cursor = ctx->d_cover_cursor.p + 2 * ch;
        ls.err_flag = ld.counters + LCNT_COVER_ERR;
        ls.icons = ctx->label_icons.p;
        ls.occ = ctx->label_occ.p + (size_t)tb * ((size_t)E * E / 32);
        ls.acc_a = setb ? B.acc.p : ctx->label_acc.p;
        ls.acc_s = ls.acc_a + ctx->l_cells_cap;
        ls.kmin = setb ? B.row_keys.p : ctx->label_row_keys.p;
        ls.kmax = ls.kmin + ctx->l_rowrecs_cap;
        ls.plane = ctx->label_plane.p + (size_t)tb * D * D;
        ls.pmask = ctx->label_pmask.p + (size_t)tb * (size_t)(D * D / 32);
        ls.D = D;
        ls.n_cover_dev = ld.counters + LCNT_COVER;
        ls.label_cnt = ctx->l_act_cnt.p + tb;
        ls.skip_flags = ld.counters + LCNT_OVERFLOW;
        CK(cudaEventRecord(ctx->ev_lcov0[ch], st));
        label_cover_kernel<<<(unsigned)ctx->num_sms * kCovCtasPerSm, 32, 0, st>>>(ls);
        CK(cudaEventRecord(ctx->ev_lcov1[ch], st));
        label_commit_kernel<<<tc, kLabelThreads, 0, st>>>(ls);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->lchunk_done[ch], st));
    }
    st = st_a;
    if (two) {  // the first stream ends behind the second one: synchronising it is synchronising the label pass
        CK(cudaEventRecord(ctx->label_join, st2));
        CK(cudaStreamWaitEvent(st, ctx->label_join, 0));
    }
    CK(cudaEventRecord(ctx->ev_label1, st));
    {
        int rc = export_words(ctx, ctx->l_counters.p, ctx->h_lcnt.p, ctx->n_lchunks * LCNT_COUNT, st);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->label_done, st));
    ctx->label_async = true;
    return OSMR_OK;
}

// After the stream has been synchronised: 0 = the label plane of this attempt is good, 1 = scratch grown, redo the call,
// 2 = this call needs the host layout, < 0 = error.
static int label_device_judge(osmr_ctx* ctx) {
    auto cells_of = [](const unsigned* c) {
        unsigned long long cells;
        memcpy(&cells, &c[LCNT_CELLS_LO], 8);
        return cells;
    };
    unsigned overflow = 0, fallback = 0;
    for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch) {
        const unsigned* c = ctx->h_lcnt.p + (size_t)ch * LCNT_COUNT;
        if (c[LCNT_BAD]) return ctx->fail(OSMR_E_INVALID, "label references an entity, style or icon that does not exist");
        if (c[LCNT_COVER_ERR]) return ctx->fail(OSMR_E_CUDA, "internal error: a label coverage term fell outside its proven window");
        fallback |= c[LCNT_FALLBACK];
        overflow |= c[LCNT_OVERFLOW];
    }
    if (fallback) return 2;
    if (overflow) {
        auto grow = [](size_t used) { return used + used / 4 + 1024; };
        for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch) {
            const unsigned* c = ctx->h_lcnt.p + (size_t)ch * LCNT_COUNT;
            // (a counter is an upper bound of what the attempt wanted only up to the first overflow: later stages were skipped)
            if (c[LCNT_OVERFLOW] & 1u) ctx->l_places_cap = std::max(grow(c[LCNT_PLACES]), ctx->l_places_cap);
            if (c[LCNT_OVERFLOW] & 2u) ctx->l_segs_cap = std::max(grow(c[LCNT_SEGS]), ctx->l_segs_cap);
            if (c[LCNT_OVERFLOW] & 4u) ctx->l_rowrecs_cap = std::max(grow(c[LCNT_ROWRECS]), ctx->l_rowrecs_cap);
            if (c[LCNT_OVERFLOW] & 8u) ctx->l_cells_cap = std::max(grow((size_t)cells_of(c)), ctx->l_cells_cap);
            if (c[LCNT_OVERFLOW] & 16u) ctx->l_ring_cap = std::max(grow(c[LCNT_RING_PTS]), ctx->l_ring_cap);
            if (c[LCNT_OVERFLOW] & 32u) ctx->l_heap_slots = std::max(grow(c[LCNT_POLY]), ctx->l_heap_slots);
            if (c[LCNT_OVERFLOW] & 64u) ctx->l_verts_cap = std::max(grow(c[LCNT_VERTS]), ctx->l_verts_cap);
            if (c[LCNT_OVERFLOW] & 128u) ctx->l_curves_cap = std::max(grow(c[LCNT_CURVES]), ctx->l_curves_cap);
        }
        // (a counter that stopped short of its real demand -- a stage skipped after an earlier overflow -- still doubles)
        if (overflow & 1u) ctx->l_places_cap = std::max(ctx->l_places_cap, ctx->l_places_used_cap * 2);
        if (overflow & 2u) ctx->l_segs_cap = std::max(ctx->l_segs_cap, ctx->l_segs_used_cap * 2);
        if (overflow & 4u) ctx->l_rowrecs_cap = std::max(ctx->l_rowrecs_cap, ctx->l_rowrecs_used_cap * 2);
        if (overflow & 8u) ctx->l_cells_cap = std::max(ctx->l_cells_cap, ctx->l_cells_used_cap * 2);
        if (overflow & 16u) ctx->l_ring_cap = std::max(ctx->l_ring_cap, ctx->l_ring_used_cap * 2);
        if (overflow & 32u) ctx->l_heap_slots = std::max(ctx->l_heap_slots, ctx->l_heap_used_slots * 2);
        if (overflow & 64u) ctx->l_verts_cap = std::max(ctx->l_verts_cap, ctx->l_verts_used_cap * 2);
        if (overflow & 128u) ctx->l_curves_cap = std::max(ctx->l_curves_cap, ctx->l_curves_used_cap * 2);
        if (ctx->l_places_cap >= 0xfffffff0ull || ctx->l_segs_cap >= 0xfffffff0ull || ctx->l_rowrecs_cap >= 0x7ffffff0ull ||
            ctx->l_cells_cap > (1ull << 33) || ctx->l_ring_cap >= 0xfffffff0ull)
            return ctx->fail(OSMR_E_NOMEM, "label scratch too large; split the batch");
        return 1;
    }
    ctx->stats_label_active = ctx->stats_label_poly = ctx->stats_label_segs = 0;
    ctx->stats_label_cells = 0;
    for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch) {
        const unsigned* c = ctx->h_lcnt.p + (size_t)ch * LCNT_COUNT;
        ctx->stats_label_active += c[LCNT_ACTIVE];
        ctx->stats_label_poly += c[LCNT_POLY];
        ctx->stats_label_segs += c[LCNT_SEGS];
        ctx->stats_label_cells += cells_of(c);
        if (getenv("OSMR_LABEL_DEBUG"))
            fprintf(stderr, "[osmr labels] chunk %u (%u tiles): active %u places %u segs %u rows %u cells %llu ring_pts %u polylabel %u covered %u curves %u\n", ch,
                    ctx->lchunk_tc[ch], c[LCNT_ACTIVE], c[LCNT_PLACES], c[LCNT_SEGS], c[LCNT_ROWRECS], cells_of(c), c[LCNT_RING_PTS], c[LCNT_POLY], c[LCNT_COVER],
                    c[LCNT_CURVES]);
        if (getenv("OSMR_LABEL_DEBUG")) fprintf(stderr, "[osmr labels]   culled %u of %u active labels (%u before the layout)\n", c[LCNT_CULLED], c[LCNT_ACTIVE], c[LCNT_PRECULLED]);
    }
    return 0;
}

// osmr_stats of a resident labelled draw whose label pass was judged good
static void resident_label_stats(osmr_ctx* ctx, uint32_t attempts) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_label0, ctx->ev_label1);
    ctx->stats.ms_label_layout = 0.f;
    ctx->stats.ms_label_device = ms;
    ctx->stats.kernel_launches += 15 * ctx->n_lchunks + 1;
    ctx->stats.label_path = 1;
    ctx->stats.n_labels_active = ctx->stats_label_active;
    ctx->stats.n_labels_polylabel = ctx->stats_label_poly;
    ctx->stats.label_attempts = attempts;
    float mc = 0.f;
    for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch) {
        float m1 = 0.f;
        cudaEventElapsedTime(&m1, ctx->ev_lcov0[ch], ctx->ev_lcov1[ch]);
        mc += m1;
    }
    ctx->stats.ms_label_cover = mc;
    ctx->stats.n_label_segments = ctx->stats_label_segs;
    ctx->stats.n_label_cells = ctx->stats_label_cells;
}

// The resident form of osmr_draw_tiles_labeled (benchmark "value" leg): batch description AND label lists uploaded once ...
int osmr_batch_upload_labeled(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin, const osmr_styled_area* areas,
                              const uint32_t* label_begin, const osmr_label* labels) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!label_begin) return ctx->fail(OSMR_E_INVALID, "null label_begin");
    if (!ctx->font.loaded()) return ctx->fail(OSMR_E_STATE, "osmr_set_font has not been called");
    ctx->batch_has_labels = false;
    int rc = batch_upload_impl(ctx, tiles, n_tiles, area_begin, areas, false, nullptr);
    if (rc) return rc;
    if (label_begin[0] != 0) return ctx->fail(OSMR_E_INVALID, "label_begin[0] must be 0");
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (label_begin[t + 1] < label_begin[t]) return ctx->fail(OSMR_E_INVALID, "label_begin must be non-decreasing");
        if (label_begin[t + 1] > label_begin[t] && !labels) return ctx->fail(OSMR_E_INVALID, "null label list");
    }
    if (!label_device_path_allowed(ctx, tiles, n_tiles))
        return ctx->fail(OSMR_E_STATE, "the resident labelled draw needs the device label layout (scale 1, 2, 4 or 8, zoom <= 18); use osmr_draw_tiles_labeled");
    cudaSetDevice(ctx->device);
    const uint32_t n_labels = label_begin[n_tiles];
    CK(ctx->d_label_list.reserve((size_t)n_labels + 1));
    CK(ctx->d_label_begin.reserve(n_tiles + 1));
    if (n_labels) CK(cudaMemcpyAsync(ctx->d_label_list.p, labels, (size_t)n_labels * sizeof(osmr_label), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_label_begin.p, label_begin, (size_t)(n_tiles + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_batch_tiles.assign(tiles, tiles + n_tiles);
    ctx->h_label_begin.assign(label_begin, label_begin + n_tiles + 1);
    ctx->batch_has_labels = true;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// ... then drawn any number of times: area passes + label pass, everything on the device.  `out` / `gpu_ms` as in osmr_batch_draw.
int osmr_batch_draw_labeled(osmr_ctx* ctx, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out, float* gpu_ms) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!ctx->has_batch || !ctx->batch_has_labels) return ctx->fail(OSMR_E_STATE, "no labelled batch uploaded (osmr_batch_upload_labeled)");
    cudaSetDevice(ctx->device);
    ctx->resident_needs_host_layout = false;
    for (int attempt = 0; attempt < 12; ++attempt) {
        int rc = label_device_enqueue(ctx, ctx->h_batch_tiles.data(), ctx->n_tiles, ctx->h_label_begin.data(), nullptr, true,
                                      out != nullptr && !(flags & OSMR_DRAW_OUT_DEVICE));
        if (rc) {
            cudaStreamSynchronize(ctx->label_stream);
            return rc;
        }
        if (ctx->label_serial) cudaStreamSynchronize(ctx->label_stream);  // (the label kernels' own durations: nothing beside them)
        ctx->label_plane_active = true;
        rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, gpu_ms);
        ctx->label_plane_active = false;
        ctx->label_async = false;
        cudaStreamSynchronize(ctx->label_stream);
        if (rc) return rc;
        const int verdict = label_device_judge(ctx);
        if (verdict < 0) return verdict;
        if (verdict == 0) {
            resident_label_stats(ctx, (uint32_t)attempt + 1);
            return OSMR_OK;
        }
        if (verdict == 2) ctx->resident_needs_host_layout = true;
        if (verdict == 2)
            return ctx->fail(OSMR_E_STATE, "this batch needs the host label layout (a flatness near-tie or an oversized polylabel); use osmr_draw_tiles_labeled");
    }
    return ctx->fail(OSMR_E_NOMEM, "label scratch kept overflowing");
} OSMR_CATCH_INT(ctx)

int osmr_draw_tiles_labeled(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                            const osmr_styled_area* areas, const uint32_t* label_begin, const osmr_label* labels,
                            const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!out) return ctx->fail(OSMR_E_INVALID, "null output buffer");
    if (!label_begin) return ctx->fail(OSMR_E_INVALID, "null label_begin");
    if (!ctx->font.loaded()) return ctx->fail(OSMR_E_STATE, "osmr_set_font has not been called");
    int rc = validate_batch(ctx, tiles, n_tiles, area_begin);
    if (rc) return rc;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (label_begin[t + 1] < label_begin[t]) return ctx->fail(OSMR_E_INVALID, "label_begin must be non-decreasing");
        if (label_begin[t + 1] > label_begin[t] && !labels) return ctx->fail(OSMR_E_INVALID, "null label list");
    }
    if (label_begin[0] != 0) return ctx->fail(OSMR_E_INVALID, "label_begin[0] must be 0");
    // (as osmr_draw_tiles: the styled areas behind the first chunk's travel on the copy stream while the first chunk is drawn;
    // the caller's arrays stay alive until this call returns)
    struct TailGuard {
        osmr_ctx* c;
        ~TailGuard() {  // whatever the way out: no copy of the caller's memory stays in flight
            cudaStreamSynchronize(c->copy_stream);
            cudaStreamSynchronize(c->label_stream);
            c->areas_deferred = false;
            c->tail_src = nullptr;
        }
    } tail_guard{ctx};
    // (copy order: the first draw chunk's styled areas, the label lists -- the label pass is the critical path --, then the rest)
    rc = batch_upload_impl(ctx, tiles, n_tiles, area_begin, areas, !(flags & OSMR_DRAW_OUT_DEVICE), out, true);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    // ---- label layout on the device: everything is enqueued behind the upload, the draw follows without a host round trip ----
    bool on_device = label_device_path_allowed(ctx, tiles, n_tiles);
    for (int attempt = 0; on_device && attempt < 12; ++attempt) {
        const auto t_host0 = std::chrono::steady_clock::now();
        rc = label_device_enqueue(ctx, tiles, n_tiles, label_begin, labels, false, !(flags & OSMR_DRAW_OUT_DEVICE));
        if (rc) {
            cudaStreamSynchronize(ctx->stream);
            return rc;
        }
        const float enqueue_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        rc = flush_area_tail(ctx);
        if (rc) return rc;
        ctx->label_plane_active = true;
        rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, nullptr);  // synchronises the compute streams (which waited for label_done)
        ctx->label_plane_active = false;
        ctx->label_async = false;
        cudaStreamSynchronize(ctx->label_stream);  // the counters' copy is the last thing on it
        if (rc) return rc;
        const int verdict = label_device_judge(ctx);
        if (verdict < 0) return verdict;
        if (verdict == 0) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev_label0, ctx->ev_label1);
            if (getenv("OSMR_TIMELINE")) {  // (diagnostics: when did what happen, relative to the start of the label pass)
                auto at = [&](cudaEvent_t e) {
                    float t = -1.f;
                    cudaEventElapsedTime(&t, ctx->ev_label0, e);
                    return t;
                };
                fprintf(stderr, "[osmr timeline] label pass end %.2f ms\n", at(ctx->ev_label1));
                for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch)
                    fprintf(stderr, "[osmr timeline]   label chunk %u (%u tiles): cover %.2f .. %.2f\n", ch, ctx->lchunk_tc[ch], at(ctx->ev_lcov0[ch]), at(ctx->ev_lcov1[ch]));
                for (unsigned c = 0; c < ctx->last_draw_chunks; ++c)
                    fprintf(stderr, "[osmr timeline]   draw chunk %u: start %.2f, bin done %.2f, cover done %.2f, raster %.2f .. %.2f\n", c, at(ctx->cev[c][0]),
                            at(ctx->cev[c][3]), at(ctx->cev[c][1]), at(ctx->cev[c][4]), at(ctx->cev[c][2]));
            }
            ctx->stats.ms_label_layout = enqueue_ms;  // host time spent on labels: table look-ups and launches only
            ctx->stats.ms_label_device = ms;
            ctx->stats.ms_total += ms;
            ctx->stats.kernel_launches += 15 * ctx->n_lchunks + 1;
            ctx->stats.label_path = 1;
            ctx->stats.n_labels_active = ctx->stats_label_active;
            ctx->stats.n_labels_polylabel = ctx->stats_label_poly;
            ctx->stats.label_attempts = (uint32_t)attempt + 1;
            {
                float mc = 0.f;
                for (unsigned ch = 0; ch < ctx->n_lchunks; ++ch) {
                    float m1 = 0.f;
                    cudaEventElapsedTime(&m1, ctx->ev_lcov0[ch], ctx->ev_lcov1[ch]);
                    mc += m1;
                }
                ctx->stats.ms_label_cover = mc;
                ctx->stats.n_label_segments = ctx->stats_label_segs;
                ctx->stats.n_label_cells = ctx->stats_label_cells;
            }
            return OSMR_OK;
        }
        if (verdict == 2) on_device = false;
    }
    // ---- host layout ----
    rc = flush_area_tail(ctx);
    if (rc) return rc;
    rc = labels_via_host(ctx, tiles, n_tiles, label_begin, labels);
    if (rc) return rc;
    ctx->label_plane_active = true;
    rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, nullptr);
    ctx->label_plane_active = false;
    ctx->stats.ms_label_layout = ctx->stats_label_layout_ms;
    ctx->stats.ms_label_device = ctx->stats_label_device_ms;
    ctx->stats.label_path = 2;
    return rc;
} OSMR_CATCH_INT(ctx)

// ---------------------------------------------------------------------------------------------------------
// f3 for the label pass: label classes per zoom + device-side listing of the label generations (osmr_auto_labels.cuh)
// ---------------------------------------------------------------------------------------------------------
int osmr_set_zoom_label_styles(osmr_ctx* ctx, uint32_t zoom, const uint32_t* node_class, const uint32_t* way_class, const uint32_t* mp_class,
                               const uint32_t* class_begin, const osmr_class_style* class_styles, uint32_t n_classes) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!ctx->ds->has_geo) return ctx->fail(OSMR_E_STATE, "osmr_set_geodata has not been called");
    if (zoom > 18) return ctx->fail(OSMR_E_INVALID, "zoom must be <= 18 (tile.rs:5 MAX_ZOOM)");
    if ((ctx->ds->n_nodes && !node_class) || (ctx->ds->n_ways && !way_class) || (ctx->ds->n_mps && !mp_class) || !class_begin)
        return ctx->fail(OSMR_E_INVALID, "null class table");
    if (class_begin[0] != 0) return ctx->fail(OSMR_E_INVALID, "class_begin[0] must be 0");
    for (uint32_t c = 0; c < n_classes; ++c) {
        if (class_begin[c + 1] < class_begin[c]) return ctx->fail(OSMR_E_INVALID, "class_begin must be non-decreasing");
        if (class_begin[c + 1] - class_begin[c] >= (1u << kAutoWithinBits)) return ctx->fail(OSMR_E_INVALID, "more than 4095 styles in one class");
    }
    const uint32_t n_cs = class_begin[n_classes];
    if (n_cs && !class_styles) return ctx->fail(OSMR_E_INVALID, "null class style list");
    for (uint32_t i = 0; i < n_cs; ++i)
        if (class_styles[i].order >= kAutoLabelNodeOrder) return ctx->fail(OSMR_E_INVALID, "label style order rank must be below 2^19");
    cudaSetDevice(ctx->device);
    osmr_ctx::ZoomTable& z = ctx->zoom_tables[zoom];
    z.l_set = false;
    CK(z.l_node_class.reserve(ctx->ds->n_nodes + 1));
    CK(z.l_way_class.reserve(ctx->ds->n_ways + 1));
    CK(z.l_mp_class.reserve(ctx->ds->n_mps + 1));
    CK(z.l_class_begin.reserve(n_classes + 1));
    CK(z.l_class_styles.reserve(n_cs + 1));
    if (ctx->ds->n_nodes) CK(cudaMemcpyAsync(z.l_node_class.p, node_class, (size_t)ctx->ds->n_nodes * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->ds->n_ways) CK(cudaMemcpyAsync(z.l_way_class.p, way_class, (size_t)ctx->ds->n_ways * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->ds->n_mps) CK(cudaMemcpyAsync(z.l_mp_class.p, mp_class, (size_t)ctx->ds->n_mps * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(z.l_class_begin.p, class_begin, (size_t)(n_classes + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n_cs) CK(cudaMemcpyAsync(z.l_class_styles.p, class_styles, (size_t)n_cs * sizeof(osmr_class_style), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    z.l_n_classes = n_classes;
    z.l_set = true;
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// Behind auto_prepare (the styled areas of the batch are resident): the label lists of the same tiles, built on the device and
// left where osmr_batch_upload_labeled would have put them.  ctx->ev[0] / ev[1] bracket the stage.
static int auto_prepare_labels(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles) {
    if (!ctx->ds->auto_labels_unavailable.empty()) return ctx->fail(OSMR_E_STATE, ctx->ds->auto_labels_unavailable.c_str());
    const uint32_t zoom = tiles[0].zoom;
    osmr_ctx::ZoomTable& z = ctx->zoom_tables[zoom];
    if (!z.l_set) return ctx->fail(OSMR_E_STATE, "osmr_set_zoom_label_styles has not been called for this zoom");
    auto& R = ctx->lres;
    if (!R.valid) {
        int rc = build_label_tables(ctx);
        if (rc) return rc;
    }
    cudaStream_t st = ctx->stream;
    ctx->batch_has_labels = false;
    CK(ctx->auto_bound.reserve(n_tiles + 2));
    CK(ctx->auto_cand_cnt.reserve(n_tiles + 1));
    CK(ctx->auto_inst.reserve(n_tiles + 2));
    CK(ctx->d_label_begin.reserve(n_tiles + 1));
    Scene s{};
    s.mps = ctx->ds->mps.p;
    s.ints = ctx->ds->ints.p;
    s.n_nodes = ctx->ds->n_nodes;
    s.n_ways = ctx->ds->n_ways;
    s.n_mps = ctx->ds->n_mps;
    s.n_ints = ctx->ds->n_ints;
    s.tiles = ctx->tiles.p;
    s.n_tiles = n_tiles;
    unsigned* auto_counters = ctx->counters.p + (size_t)kMaxChunks * CNT_COUNT;
    unsigned* h_auto = ctx->h_cnt.p + (size_t)kMaxChunks * CNT_COUNT;
    s.counters = auto_counters;
    AutoLabelScene a{};
    a.idx_xy = ctx->ds->idx_xy.p;
    a.idx_n = ctx->ds->idx_n.p;
    a.idx_w = ctx->ds->idx_w.p;
    a.idx_m = ctx->ds->idx_m.p;
    a.n_idx = ctx->ds->n_idx;
    a.way_min_tile = ctx->ds->way_min_tile.p;
    a.mp_min_tile = ctx->ds->mp_min_tile.p;
    a.way_rank = ctx->ds->way_rank.p;
    a.mp_rank = ctx->ds->mp_rank.p;
    a.rank_entity = ctx->ds->rank_entity.p;
    a.node_rank = ctx->ds->node_rank.p;
    a.node_rank_entity = ctx->ds->node_rank_entity.p;
    a.node_class = z.l_node_class.p;
    a.way_class = z.l_way_class.p;
    a.mp_class = z.l_mp_class.p;
    a.class_begin = z.l_class_begin.p;
    a.class_styles = z.l_class_styles.p;
    a.n_classes = z.l_n_classes;
    a.lstyles = R.styles.p;
    a.n_lstyles = (unsigned)ctx->label_styles.size();
    a.text_id = R.text_id.p;
    a.ent_total = R.ent_total;
    a.way_base = R.way_base;
    a.mp_base = R.mp_base;
    a.bound = ctx->auto_bound.p;
    a.cand_cnt = ctx->auto_cand_cnt.p;
    a.inst_cnt = ctx->auto_inst.p;

    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaMemsetAsync(auto_counters, 0, CNT_COUNT * sizeof(unsigned), st));
    autol_bound_kernel<<<n_tiles, kAutoThreads, 0, st>>>(s, a);
    auto_scan_kernel<<<1, 1024, 0, st>>>(ctx->auto_bound.p, n_tiles, &auto_counters[CNT_OVERFLOW]);
    unsigned total_bound = 0;
    CK(cudaMemcpyAsync(&total_bound, ctx->auto_bound.p + n_tiles, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_auto, auto_counters, CNT_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_auto[CNT_OVERFLOW]) return ctx->fail(OSMR_E_NOMEM, "more than 2^32 label candidate references; split the batch");
    CK(ctx->auto_cand.reserve((size_t)total_bound + 1));
    a.cand = ctx->auto_cand.p;
    autol_gather_kernel<<<n_tiles, kAutoThreads, 0, st>>>(s, a);
    auto_scan_kernel<<<1, 1024, 0, st>>>(ctx->auto_inst.p, n_tiles, &auto_counters[CNT_OVERFLOW]);
    CK(cudaMemcpyAsync(ctx->d_label_begin.p, ctx->auto_inst.p, (size_t)(n_tiles + 1) * 4, cudaMemcpyDeviceToDevice, st));
    ctx->h_label_begin.resize(n_tiles + 1);
    CK(cudaMemcpyAsync(ctx->h_label_begin.data(), ctx->auto_inst.p, (size_t)(n_tiles + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_auto, auto_counters, CNT_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_auto[CNT_OVERFLOW]) return ctx->fail(OSMR_E_NOMEM, "more than 2^32 label generations; split the batch");
    if (h_auto[CNT_BAD_INPUT] & 1u) return ctx->fail(OSMR_E_INVALID, "label class style list references a label style that does not exist");
    if (h_auto[CNT_BAD_INPUT]) return ctx->fail(OSMR_E_INVALID, "tile index references an entity that does not exist");
    const uint32_t n_labels = ctx->h_label_begin[n_tiles];
    CK(ctx->d_label_list.reserve((size_t)n_labels + 1));
    a.labels_out = ctx->d_label_list.p;
    for (int attempt = 0;; ++attempt) {
        a.big_keys = ctx->auto_big_keys.p;
        a.big_cap = ctx->auto_big_keys.cap;
        CK(cudaMemsetAsync(auto_counters, 0, CNT_COUNT * sizeof(unsigned), st));
        autol_sort_kernel<<<n_tiles, kAutoThreads, 0, st>>>(s, a);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_auto, auto_counters, CNT_COUNT * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!(h_auto[CNT_OVERFLOW] & 8u)) break;
        if (attempt >= 2) return ctx->fail(OSMR_E_NOMEM, "sort scratch kept overflowing");
        unsigned long long need;
        memcpy(&need, &h_auto[CNT_WALK_ALPHA], 8);
        CK(ctx->auto_big_keys.reserve((size_t)need + 1024));
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    ctx->h_batch_tiles.assign(tiles, tiles + n_tiles);
    ctx->batch_has_labels = true;
    return OSMR_OK;
}

// tile list in, finished tiles out: f3 for both halves of draw_to_pixels (drawer.rs:60-131) in front of the draw path
static int auto_labeled_draw(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out) {
    if (!ctx->font.loaded()) return ctx->fail(OSMR_E_STATE, "osmr_set_font has not been called");
    // The label pass is the long pole of a labelled draw and runs on its own stream: its lists are built FIRST, the pass is
    // enqueued, and the styled-area lists (three host round trips) are built beside it.
    int rc = auto_begin(ctx, tiles, n_tiles, canvas_rgb, flags);
    if (rc) return rc;
    float ms_areas = 0.f, ms_labels = 0.f;  // (the stages share their events: read them before the next stage)
    rc = auto_prepare_labels(ctx, tiles, n_tiles);
    if (rc) return rc;
    cudaEventSynchronize(ctx->ev[1]);
    cudaEventElapsedTime(&ms_labels, ctx->ev[0], ctx->ev[1]);
    const uint32_t n_labels = ctx->h_label_begin[n_tiles];
    const bool chunked = out != nullptr && !(flags & OSMR_DRAW_OUT_DEVICE);
    bool host_layout = !label_device_path_allowed(ctx, tiles, n_tiles);
    bool drawn = false;
    ctx->resident_needs_host_layout = false;
    if (!host_layout) {
        rc = label_device_enqueue(ctx, tiles, n_tiles, ctx->h_label_begin.data(), nullptr, true, chunked);
        if (rc) {
            cudaStreamSynchronize(ctx->label_stream);
            return rc;
        }
    }
    rc = auto_prepare(ctx, tiles, n_tiles, canvas_rgb, flags, true);
    if (rc) {
        ctx->label_async = false;
        cudaStreamSynchronize(ctx->label_stream);
        return rc;
    }
    cudaEventSynchronize(ctx->ev[1]);
    cudaEventElapsedTime(&ms_areas, ctx->ev[0], ctx->ev[1]);
    if (!host_layout) {
        ctx->label_plane_active = true;
        rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, nullptr);
        ctx->label_plane_active = false;
        ctx->label_async = false;
        cudaStreamSynchronize(ctx->label_stream);
        if (rc) return rc;
        const int verdict = label_device_judge(ctx);
        if (verdict < 0) return verdict;
        if (verdict == 0) {
            resident_label_stats(ctx, 1);
            drawn = true;
        } else if (verdict == 1) {  // label scratch grown: the plain resident loop redoes the call
            rc = osmr_batch_draw_labeled(ctx, canvas_rgb, flags, out, nullptr);
            if (rc && !ctx->resident_needs_host_layout) return rc;
            drawn = rc == OSMR_OK;
            if (drawn) ctx->stats.label_attempts += 1;
        }
        host_layout = !drawn;
    }
    if (host_layout) {  // the lists come back once and the host lays the labels out (same pixels, osmr_stats.label_path = 2)
        std::vector<osmr_label> h_labels((size_t)n_labels + 1);
        if (n_labels) CK(cudaMemcpyAsync(h_labels.data(), ctx->d_label_list.p, (size_t)n_labels * sizeof(osmr_label), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const std::vector<uint32_t> h_begin = ctx->h_label_begin;
        ctx->batch_has_labels = false;  // (labels_via_host reuses the device label arrays)
        rc = labels_via_host(ctx, tiles, n_tiles, h_begin.data(), h_labels.data());
        if (rc) return rc;
        ctx->label_plane_active = true;
        rc = osmr_batch_draw(ctx, canvas_rgb, flags, out, nullptr);
        ctx->label_plane_active = false;
        if (rc) return rc;
        ctx->stats.ms_label_layout = ctx->stats_label_layout_ms;
        ctx->stats.ms_label_device = ctx->stats_label_device_ms;
        ctx->stats.label_path = 2;
    }
    ctx->stats.ms_auto = ms_areas + ms_labels;
    ctx->stats.ms_total += ms_areas + ms_labels;
    ctx->stats.kernel_launches += 5;
    return OSMR_OK;
}

int osmr_draw_tiles_auto_labeled(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!out && !(flags & OSMR_DRAW_OUT_DEVICE)) return ctx->fail(OSMR_E_INVALID, "null output buffer");
    return auto_labeled_draw(ctx, tiles, n_tiles, canvas_rgb, flags, out);
} OSMR_CATCH_INT(ctx)

// Drawer::draw_tile with its label pass for a tile list (drawer.rs:40-58 behind http_server.rs:150-177): tile list in, PNG files out
int osmr_draw_tiles_auto_labeled_png(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags,
                                     uint8_t* png_out, size_t png_cap, uint64_t* png_offset) try {
    if (!ctx) return OSMR_E_INVALID;
    if (!png_out || !png_offset) return ctx->fail(OSMR_E_INVALID, "null output buffer");
    if (flags & (OSMR_DRAW_OUT_RGBA | OSMR_DRAW_OUT_DEVICE)) return ctx->fail(OSMR_E_INVALID, "osmr_draw_tiles_auto_labeled_png encodes RGB into host memory");
    int rc = auto_labeled_draw(ctx, tiles, n_tiles, canvas_rgb, flags, nullptr);  // the RGB tiles stay in HBM (ctx->out)
    if (rc) return rc;
    return encode_png_from_device(ctx, ctx->out.p, n_tiles, (unsigned)ctx->scale, png_out, png_cap, png_offset);
} OSMR_CATCH_INT(ctx)

int osmr_auto_readback_labels(osmr_ctx* ctx, uint32_t* label_begin, osmr_label* labels, uint32_t labels_cap) try {
    if (!ctx || !label_begin) return OSMR_E_INVALID;
    if (!ctx->has_batch || ctx->h_label_begin.size() != (size_t)ctx->n_tiles + 1) return ctx->fail(OSMR_E_STATE, "no device-built label lists");
    cudaSetDevice(ctx->device);
    memcpy(label_begin, ctx->h_label_begin.data(), (size_t)(ctx->n_tiles + 1) * 4);
    if (labels) {
        const uint32_t n = ctx->h_label_begin[ctx->n_tiles];
        if (!ctx->batch_has_labels) return ctx->fail(OSMR_E_STATE, "the device label lists were replaced by the host layout");
        if (labels_cap < n) return ctx->fail(OSMR_E_INVALID, "labels_cap too small");
        if (n) CK(cudaMemcpyAsync(labels, ctx->d_label_list.p, (size_t)n * sizeof(osmr_label), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return OSMR_OK;
} OSMR_CATCH_INT(ctx)

// pinned host memory for callers that want full-speed transfers (optional)
void* osmr_alloc_pinned(size_t bytes) try {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
} OSMR_CATCH_VAL(nullptr)
void osmr_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"

// osmr_labels_host.hpp -- host half of the label pass of libosmr_b200.so.
//
// The reference does label layout with string, font and heap code on the CPU (labeler.rs, labelable.rs,
// font/text_placer.rs) and only then touches pixels.  This file is that first half for the GPU path: for every
// label generation of a tile it produces what the device needs to rasterise and collide it --
//   * an icon blit rectangle (labeler.rs:91-106: top-left = (centre - dim/2) as i32), and / or
//   * the glyph outlines of the text as a list of f64 line segments in the exact order and with the exact
//     coordinates of the reference's Rasterizer::draw_line calls (text_placer.rs:60-168,211-231; quadratic curves are
//     flattened by rasterizer.rs:86-107's recursive midpoint rule here, because it needs libm hypot).
// All floating point here goes through the same glibc functions the reference's Rust std calls (tan, log, atan2,
// sin, cos, hypot), so the segments are bit-identical to the reference's; the device side (label_kernel) only uses
// IEEE basic operations.  Font access follows stb_truetype 0.3.1 (Cargo.lock), the crate behind
// text_placer.rs:18,49,177-187,201.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include <shared_mutex>
#include <mutex>

#include "osmr.h"

namespace osmr_host {

// ---------------------------------------------------------------------------------------------------------
// records handed to the device (mirrored in osmr_kernels.cuh)
// ---------------------------------------------------------------------------------------------------------
struct LabelRec {
    int icon;                 // label icon table index, -1: none
    int ix, iy;               // icon top-left in tile pixels
    unsigned seg_begin;       // first segment (index into the batch's segment array)
    unsigned seg_count;
    int bx0, by0, bx1, by1;   // inclusive pixel bbox of everything the text can touch; bx0 > bx1: no text geometry
    unsigned rgb;             // text colour 0x00BBGGRR
    // filled by the batch assembler (osmr_draw_tiles_labeled): coverage storage of the rows inside the label canvas
    int ry0;                  // first stored row (by0 clipped to the 3x3 canvas)
    int rows;                 // stored rows (0: the text cannot touch the canvas)
    int width;                // bx1 - bx0 + 1
    unsigned row_first;       // first entry of this label in the per-row key range arrays
    unsigned n_ranges;        // 0: the segments are one range (device layout: glyph ranges)
    unsigned range_off;
    unsigned long long cell_off;  // first cell of this label in the coverage arrays
};
struct Seg {
    double x0, y0, x1, y1;
};

// ---------------------------------------------------------------------------------------------------------
// geodata image views (reader.rs:264-335,351-373)
// ---------------------------------------------------------------------------------------------------------
struct BinView {
    const uint8_t* nodes = nullptr;
    const uint8_t* ways = nullptr;
    const uint8_t* polys = nullptr;
    const uint8_t* mps = nullptr;
    const uint8_t* ints = nullptr;
    const uint8_t* strings = nullptr;
    uint32_t n_nodes = 0, n_ways = 0, n_polys = 0, n_mps = 0, n_ints = 0;
    size_t strings_len = 0;
    static uint32_t u32(const uint8_t* p) {
        uint32_t v;
        memcpy(&v, p, 4);
        return v;
    }
    static double f64(const uint8_t* p) {
        double v;
        memcpy(&v, p, 8);
        return v;
    }
    bool parse(const uint8_t* p, size_t len) {
        size_t pos = 0;
        const size_t rec[6] = {32, 24, 8, 24, 32, 4};
        const uint8_t* base[6];
        uint32_t cnt[6];
        for (int i = 0; i < 6; ++i) {
            if (pos + 4 > len) return false;
            cnt[i] = u32(p + pos);
            pos += 4;
            if ((size_t)cnt[i] * rec[i] > len - pos) return false;
            base[i] = p + pos;
            pos += (size_t)cnt[i] * rec[i];
        }
        nodes = base[0];
        ways = base[1];
        polys = base[2];
        mps = base[3];
        ints = base[5];
        n_nodes = cnt[0];
        n_ways = cnt[1];
        n_polys = cnt[2];
        n_mps = cnt[3];
        n_ints = cnt[5];
        strings = p + pos;
        strings_len = len - pos;
        return true;
    }
    uint32_t int_at(uint32_t i) const { return u32(ints + (size_t)i * 4); }
    double lat(uint32_t n) const { return f64(nodes + (size_t)n * 32 + 8); }
    double lon(uint32_t n) const { return f64(nodes + (size_t)n * 32 + 16); }
    // Tags::get_by_key (reader.rs:351-373): kv refs sorted by key, byte-wise comparison
    bool tag(uint32_t tags_off, uint32_t tags_len, const char* key, size_t key_len, const char*& val, size_t& val_len) const {
        uint32_t n = tags_len / 4;
        if ((uint64_t)tags_off + tags_len > n_ints) return false;
        for (uint32_t lo = 0, hi = n; lo < hi;) {
            uint32_t mid = (lo + hi) / 2;
            uint32_t ko = int_at(tags_off + 4 * mid), kl = int_at(tags_off + 4 * mid + 1);
            if ((uint64_t)ko + kl > strings_len) return false;
            size_t m = std::min<size_t>(kl, key_len);
            int c = memcmp(strings + ko, key, m);
            if (c == 0) c = (kl < key_len) ? -1 : (kl > key_len ? 1 : 0);
            if (c == 0) {
                uint32_t vo = int_at(tags_off + 4 * mid + 2), vl = int_at(tags_off + 4 * mid + 3);
                if ((uint64_t)vo + vl > strings_len) return false;
                val = (const char*)strings + vo;
                val_len = vl;
                return true;
            }
            if (c < 0)
                lo = mid + 1;
            else
                hi = mid;
        }
        return false;
    }
};

// ---------------------------------------------------------------------------------------------------------
// projection as the reference computes it on the CPU (tile.rs:88-106, point.rs:11-19)
// ---------------------------------------------------------------------------------------------------------
static const double kPiH = 3.14159265358979323846264338327950288;
inline int32_t f64_as_i32(double v) {
    if (v != v) return 0;
    if (v <= -2147483648.0) return std::numeric_limits<int32_t>::min();
    if (v >= 2147483647.0) return std::numeric_limits<int32_t>::max();
    return (int32_t)v;
}
inline void coords_rel(const BinView& g, uint32_t node, const osmr_tile& t, double& x, double& y) {
    const double rads_per_deg = kPiH / 180.0;
    double lat_rad = g.lat(node) * rads_per_deg, lon_rad = g.lon(node) * rads_per_deg;
    double mx = lon_rad + kPiH;
    double my = kPiH - std::log(std::tan((kPiH / 4.0) + (lat_rad / 2.0)));
    double dim = (double)(uint32_t)(256u * (1u << t.zoom));
    x = (mx / (2.0 * kPiH)) * dim - (double)(uint32_t)(t.x * 256u);
    y = (my / (2.0 * kPiH)) * dim - (double)(uint32_t)(t.y * 256u);
}
struct IPoint {
    int32_t x, y;
};
inline IPoint point_from_node(const BinView& g, uint32_t node, const osmr_tile& t, double scale) {
    double x, y;
    coords_rel(g, node, t, x, y);
    return IPoint{f64_as_i32(std::round(x * scale)), f64_as_i32(std::round(y * scale))};
}
inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline double ipoint_dist(const IPoint& a, const IPoint& b) {
    double dx = (double)wsub(a.x, b.x), dy = (double)wsub(a.y, b.y);
    return std::sqrt(dx * dx + dy * dy);
}

// ---------------------------------------------------------------------------------------------------------
// TrueType access, behaviour of stb_truetype 0.3.1: FontInfo::new, find_glyph_index, get_glyph_h_metrics,
// get_glyph_kern_advance, get_v_metrics, scale_for_pixel_height, get_glyph_shape
// ---------------------------------------------------------------------------------------------------------
enum { VT_MOVE = 1, VT_LINE = 2, VT_CURVE = 3 };
struct GlyphVertex {
    int16_t x, y, cx, cy;
    uint8_t type;
};

class TrueType {
   public:
    // Every offset read from the file is checked against its length: accessors return 0 for bytes beyond the end, so a
    // truncated or corrupt font yields missing glyphs (or a failed load), never an out-of-bounds read.
    bool load(const uint8_t* data, size_t len) {
        index_map_ = 0;  // a failed load leaves the object unloaded
        d_ = nullptr;
        glyph_cache_.clear();
        bytes_.assign(data, data + len);
        // zero padding behind the file: multi-byte reads that start inside the file never leave the buffer
        bytes_.resize(len + 16, 0);
        len_ = len;
        if (len < 12) return false;
        d_ = bytes_.data();
        if (12 + 16 * (size_t)be16(4) > len) {
            d_ = nullptr;
            return false;
        }
        cmap_ = table("cmap");
        loca_ = table("loca");
        head_ = table("head");
        glyf_ = table("glyf");
        hhea_ = table("hhea");
        hmtx_ = table("hmtx");
        kern_ = table("kern");
        if (!cmap_ || !loca_ || !head_ || !glyf_ || !hhea_ || !hmtx_ || (size_t)head_ + 54 > len || (size_t)hhea_ + 36 > len ||
            (size_t)cmap_ + 4 > len) {
            d_ = nullptr;
            return false;
        }
        uint32_t maxp = table("maxp");
        num_glyphs_ = maxp ? be16(maxp + 4) : 0xffff;
        index_map_ = 0;
        int n = be16(cmap_ + 2);
        for (int i = 0; i < n; ++i) {  // the last Microsoft-Unicode or Unicode-platform record wins
            uint32_t rec = cmap_ + 4 + 8 * i;
            int pid = be16(rec), eid = be16(rec + 2);
            if ((pid == 3 && (eid == 1 || eid == 10)) || pid == 0) index_map_ = cmap_ + be32(rec + 4);
        }
        loc_format_ = be16(head_ + 50);
        glyph_cache_.clear();
        if (index_map_ == 0 || index_map_ >= len) {
            index_map_ = 0;
            d_ = nullptr;
            return false;
        }
        return true;
    }
    bool loaded() const { return d_ != nullptr && index_map_ != 0; }
    int ascent() const { return sbe16(hhea_ + 4); }
    int descent() const { return sbe16(hhea_ + 6); }
    int line_gap() const { return sbe16(hhea_ + 8); }
    float scale_for_pixel_height(float h) const { return h / (float)(ascent() - descent()); }
    int glyph_index(uint32_t cp) const {
        uint32_t m = index_map_;
        int fmt = be16(m);
        if (fmt == 4) {
            if (cp > 0xffff) return 0;
            int segs = be16(m + 6) >> 1;
            uint32_t ends = m + 14, starts = m + 16 + segs * 2, deltas = m + 16 + segs * 4, offs = m + 16 + segs * 6;
            int lo = 0, hi = segs;
            while (lo < hi) {  // first segment whose end code is >= cp
                int mid = (lo + hi) / 2;
                if (be16(ends + 2 * mid) >= cp)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            if (lo >= segs) return 0;
            uint32_t st = be16(starts + 2 * lo);
            if (cp < st) return 0;
            uint32_t ro = be16(offs + 2 * lo);
            if (ro == 0) return (uint16_t)(cp + sbe16(deltas + 2 * lo));
            return be16(offs + 2 * lo + ro + (cp - st) * 2);
        }
        if (fmt == 12 || fmt == 13) {
            uint32_t groups = be32(m + 12);
            int64_t lo = 0, hi = groups;
            while (lo < hi) {
                int64_t mid = lo + ((hi - lo) >> 1);
                uint32_t sc = be32(m + 16 + mid * 12), ec = be32(m + 16 + mid * 12 + 4);
                if (cp < sc)
                    hi = mid;
                else if (cp > ec)
                    lo = mid + 1;
                else
                    return (int)(be32(m + 16 + mid * 12 + 8) + (fmt == 12 ? cp - sc : 0));
            }
            return 0;
        }
        if (fmt == 0) return (int)cp < be16(m + 2) - 6 ? u8(m + 6 + cp) : 0;
        if (fmt == 6) {
            uint32_t first = be16(m + 6), count = be16(m + 8);
            return (cp >= first && cp < first + count) ? be16(m + 10 + (cp - first) * 2) : 0;
        }
        return 0;
    }
    int advance(int g) const {
        int n = be16(hhea_ + 34);
        return g < n ? sbe16(hmtx_ + 4 * g) : sbe16(hmtx_ + 4 * (n - 1));
    }
    int kerning(int g1, int g2) const {
        if (!kern_ || be16(kern_ + 2) < 1 || be16(kern_ + 8) != 1) return 0;
        int l = 0, r = be16(kern_ + 10) - 1;
        uint32_t needle = ((uint32_t)g1 << 16) | (uint32_t)g2;
        while (l <= r) {
            int m = (l + r) >> 1;
            uint32_t straw = be32(kern_ + 18 + m * 6);
            if (needle < straw)
                r = m - 1;
            else if (needle > straw)
                l = m + 1;
            else
                return sbe16(kern_ + 22 + m * 6);
        }
        return 0;
    }
    // outline of a glyph (nullptr: the glyph has none); memoised behind a reader/writer lock (tiles are laid out by
    // several host threads; unordered_map nodes never move, so the returned pointer stays valid)
    const std::vector<GlyphVertex>* shape(int g) const {
        {
            std::shared_lock<std::shared_mutex> rd(glyph_mu_);
            auto it = glyph_cache_.find(g);
            if (it != glyph_cache_.end()) return it->second.empty() ? nullptr : &it->second;
        }
        std::vector<GlyphVertex> v;
        if (!outline(g, v, 0)) v.clear();
        std::unique_lock<std::shared_mutex> wr(glyph_mu_);
        auto it = glyph_cache_.emplace(g, std::move(v)).first;
        return it->second.empty() ? nullptr : &it->second;
    }

   private:
    std::vector<uint8_t> bytes_;
    const uint8_t* d_ = nullptr;
    uint32_t cmap_ = 0, loca_ = 0, head_ = 0, glyf_ = 0, hhea_ = 0, hmtx_ = 0, kern_ = 0, index_map_ = 0;
    int num_glyphs_ = 0, loc_format_ = 0;
    size_t len_ = 0;
    mutable std::unordered_map<int, std::vector<GlyphVertex>> glyph_cache_;
    mutable std::shared_mutex glyph_mu_;

    // (the buffer carries 16 zero bytes behind the file, so a read that STARTS inside the file is always in bounds)
    int u8(uint64_t o) const { return o < len_ ? d_[o] : 0; }
    int be16(uint64_t o) const { return o < len_ ? d_[o] * 256 + d_[o + 1] : 0; }
    int sbe16(uint64_t o) const { return o < len_ ? (int16_t)(d_[o] * 256 + d_[o + 1]) : 0; }
    uint32_t be32(uint64_t o) const { return o < len_ ? (((uint32_t)d_[o] << 24) | (d_[o + 1] << 16) | (d_[o + 2] << 8) | d_[o + 3]) : 0u; }
    uint32_t table(const char* tag) const {
        int n = be16(4);
        for (int i = 0; i < n; ++i)
            if ((size_t)12 + 16 * i + 16 <= len_ && memcmp(d_ + 12 + 16 * i, tag, 4) == 0) return be32(12 + 16 * i + 8);
        return 0;
    }
    int glyf_offset(int g) const {
        if (g >= num_glyphs_ || loc_format_ >= 2) return -1;
        uint32_t a, b;
        if (loc_format_ == 0) {
            a = glyf_ + be16(loca_ + g * 2) * 2;
            b = glyf_ + be16(loca_ + g * 2 + 2) * 2;
        } else {
            a = glyf_ + be32(loca_ + g * 4);
            b = glyf_ + be32(loca_ + g * 4 + 4);
        }
        return a == b ? -1 : (int)a;
    }
    static int16_t f32_to_i16(float f) {  // Rust `as i16`
        if (f != f) return 0;
        if (f <= -32768.0f) return -32768;
        if (f >= 32767.0f) return 32767;
        return (int16_t)f;
    }
    static void close(std::vector<GlyphVertex>& v, bool was_off, bool start_off, int sx, int sy, int scx, int scy, int cx, int cy) {
        if (start_off) {
            if (was_off) v.push_back({(int16_t)((cx + scx) >> 1), (int16_t)((cy + scy) >> 1), (int16_t)cx, (int16_t)cy, VT_CURVE});
            v.push_back({(int16_t)sx, (int16_t)sy, (int16_t)scx, (int16_t)scy, VT_CURVE});
        } else if (was_off) {
            v.push_back({(int16_t)sx, (int16_t)sy, (int16_t)cx, (int16_t)cy, VT_CURVE});
        } else {
            v.push_back({(int16_t)sx, (int16_t)sy, 0, 0, VT_LINE});
        }
    }
    bool outline(int glyph, std::vector<GlyphVertex>& out, int depth) const {
        out.clear();
        int g = glyf_offset(glyph);
        if (g < 0 || depth > 8) return false;
        int contours = sbe16(g);
        if (contours > 0) {
            uint32_t endpts = g + 10;
            int ins = be16(g + 10 + contours * 2);
            uint32_t p = g + 10 + contours * 2 + 2 + ins;
            int n = 1 + be16(endpts + contours * 2 - 2);
            std::vector<uint8_t> fl(n + 1, 1);
            std::vector<int> xs(n + 1, 0), ys(n + 1, 0);
            int repeat = 0;
            uint8_t flags = 0;
            for (int i = 0; i < n; ++i) {
                if (repeat == 0) {
                    flags = (uint8_t)u8(p++);
                    if (flags & 8) repeat = u8(p++);
                } else {
                    --repeat;
                }
                fl[i] = flags;
            }
            int x = 0;
            for (int i = 0; i < n; ++i) {
                if (fl[i] & 2) {
                    int dx = u8(p++);
                    x += (fl[i] & 16) ? dx : -dx;
                } else if (!(fl[i] & 16)) {
                    x += sbe16(p);
                    p += 2;
                }
                xs[i] = (int16_t)x;
            }
            int y = 0;
            for (int i = 0; i < n; ++i) {
                if (fl[i] & 4) {
                    int dy = u8(p++);
                    y += (fl[i] & 32) ? dy : -dy;
                } else if (!(fl[i] & 32)) {
                    y += sbe16(p);
                    p += 2;
                }
                ys[i] = (int16_t)y;
            }
            int next_move = 0, j = 0, sx = 0, sy = 0, cx = 0, cy = 0, scx = 0, scy = 0;
            bool was_off = false, start_off = false;
            for (int i = 0; i < n; ++i) {
                x = xs[i];
                y = ys[i];
                if (next_move == i) {
                    if (i != 0) close(out, was_off, start_off, sx, sy, scx, scy, cx, cy);
                    start_off = !(fl[i] & 1);
                    if (start_off) {
                        scx = x;
                        scy = y;
                        if (!(fl[i + 1] & 1)) {
                            sx = (x + xs[i + 1]) >> 1;
                            sy = (y + ys[i + 1]) >> 1;
                        } else {
                            sx = xs[i + 1];
                            sy = ys[i + 1];
                            ++i;
                        }
                    } else {
                        sx = x;
                        sy = y;
                    }
                    out.push_back({(int16_t)sx, (int16_t)sy, 0, 0, VT_MOVE});
                    was_off = false;
                    next_move = 1 + be16(endpts + j * 2);
                    ++j;
                } else if (!(fl[i] & 1)) {
                    if (was_off) out.push_back({(int16_t)((cx + x) >> 1), (int16_t)((cy + y) >> 1), (int16_t)cx, (int16_t)cy, VT_CURVE});
                    cx = x;
                    cy = y;
                    was_off = true;
                } else {
                    if (was_off)
                        out.push_back({(int16_t)x, (int16_t)y, (int16_t)cx, (int16_t)cy, VT_CURVE});
                    else
                        out.push_back({(int16_t)x, (int16_t)y, 0, 0, VT_LINE});
                    was_off = false;
                }
            }
            close(out, was_off, start_off, sx, sy, scx, scy, cx, cy);
        } else if (contours == -1) {
            uint32_t comp = g + 10;
            for (bool more = true; more;) {
                int flags = be16(comp);
                int gidx = be16(comp + 2);
                comp += 4;
                float mtx[6] = {1, 0, 0, 1, 0, 0};
                if (!(flags & 2)) return false;  // point matching: unsupported by stb_truetype as well
                if (flags & 1) {
                    mtx[4] = (float)sbe16(comp);
                    mtx[5] = (float)sbe16(comp + 2);
                    comp += 4;
                } else {
                    mtx[4] = (float)(int8_t)u8(comp);
                    mtx[5] = (float)(int8_t)u8(comp + 1);
                    comp += 2;
                }
                if (flags & (1 << 3)) {
                    mtx[0] = mtx[3] = (float)sbe16(comp) / 16384.0f;
                    comp += 2;
                } else if (flags & (1 << 6)) {
                    mtx[0] = (float)sbe16(comp) / 16384.0f;
                    mtx[3] = (float)sbe16(comp + 2) / 16384.0f;
                    comp += 4;
                } else if (flags & (1 << 7)) {
                    mtx[0] = (float)sbe16(comp) / 16384.0f;
                    mtx[1] = (float)sbe16(comp + 2) / 16384.0f;
                    mtx[2] = (float)sbe16(comp + 4) / 16384.0f;
                    mtx[3] = (float)sbe16(comp + 6) / 16384.0f;
                    comp += 8;
                }
                float m = std::sqrt(mtx[0] * mtx[0] + mtx[1] * mtx[1]);
                float n = std::sqrt(mtx[2] * mtx[2] + mtx[3] * mtx[3]);
                std::vector<GlyphVertex> part;
                if (outline(gidx, part, depth + 1)) {
                    for (GlyphVertex& v : part) {
                        float px = (float)v.x, py = (float)v.y;
                        v.x = f32_to_i16(m * (mtx[0] * px + mtx[2] * py + mtx[4]));
                        v.y = f32_to_i16(n * (mtx[1] * px + mtx[3] * py + mtx[5]));
                        px = (float)v.cx;
                        py = (float)v.cy;
                        v.cx = f32_to_i16(m * (mtx[0] * px + mtx[2] * py + mtx[4]));
                        v.cy = f32_to_i16(n * (mtx[1] * px + mtx[3] * py + mtx[5]));
                    }
                    out.insert(out.end(), part.begin(), part.end());
                }
                more = (flags & (1 << 5)) != 0;
            }
        } else {
            return false;
        }
        return !out.empty();
    }
};

// ---------------------------------------------------------------------------------------------------------
// polylabel (labelable.rs:125-232), with std::collections::BinaryHeap's sift order
// ---------------------------------------------------------------------------------------------------------
typedef std::pair<double, double> PF;
typedef std::vector<std::vector<PF>> Rings;

inline double seg_dist_sq(const PF& p, const PF& a, const PF& b) {
    double x = a.first, y = a.second, dx = b.first - x, dy = b.second - y;
    if (dx != 0.0 || dy != 0.0) {
        double t = ((p.first - x) * dx + (p.second - y) * dy) / (dx * dx + dy * dy);
        if (t > 1.0) {
            x = b.first;
            y = b.second;
        } else if (t > 0.0) {
            x += dx * t;
            y += dy * t;
        }
    }
    dx = p.first - x;
    dy = p.second - y;
    return dx * dx + dy * dy;
}
inline double signed_dist(const PF& p, const Rings& rings, size_t n_rings) {
    bool inside = false;
    double best = std::numeric_limits<double>::infinity();
    for (size_t k = 0; k < n_rings; ++k)
        for (size_t i = 1; i < rings[k].size(); ++i) {
            const PF &a = rings[k][i], &b = rings[k][i - 1];
            if ((a.second > p.second) != (b.second > p.second) &&
                p.first < (b.first - a.first) * (p.second - a.second) / (b.second - a.second) + a.first)
                inside = !inside;
            best = std::fmin(best, seg_dist_sq(p, a, b));
        }
    return (inside ? 1.0 : -1.0) * std::sqrt(best);
}
struct HeapCell {
    PF c;
    double half, fit, max_fit;
};
class MaxHeapRust {  // push: sift_up; pop: move last to the root, sift to the bottom, sift back up
   public:
    void push(const HeapCell& v) {
        d_.push_back(v);
        up(d_.size() - 1);
    }
    bool pop(HeapCell& out) {
        if (d_.empty()) return false;
        HeapCell last = d_.back();
        d_.pop_back();
        if (d_.empty()) {
            out = last;
            return true;
        }
        out = d_[0];
        size_t end = d_.size(), pos = 0, child = 1;
        while (end >= 2 && child <= end - 2) {
            if (le(d_[child], d_[child + 1])) ++child;
            d_[pos] = d_[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            d_[pos] = d_[child];
            pos = child;
        }
        d_[pos] = last;
        up(pos);
        return true;
    }

   private:
    std::vector<HeapCell> d_;
    static bool le(const HeapCell& a, const HeapCell& b) { return !(a.max_fit > b.max_fit); }  // cmp: Less or Equal
    void up(size_t pos) {
        HeapCell e = d_[pos];
        while (pos > 0) {
            size_t parent = (pos - 1) / 2;
            if (le(e, d_[parent])) break;
            d_[pos] = d_[parent];
            pos = parent;
        }
        d_[pos] = e;
    }
};

inline bool label_anchor_of_rings(Rings rings, double scale, PF& out) {  // get_label_position (labelable.rs:191-204)
    if (rings.empty() || rings[0].empty()) return false;
    auto ring_area = [](const std::vector<PF>& r) {
        double s = 0.0;
        for (size_t i = 1; i < r.size(); ++i) s += r[i].first * r[i - 1].second - r[i - 1].first * r[i].second;
        return std::fabs(s);
    };
    size_t big = 0;
    double big_area = ring_area(rings[0]);
    for (size_t i = 1; i < rings.size(); ++i) {
        double a = ring_area(rings[i]);
        if (a > big_area) {
            big = i;
            big_area = a;
        }
    }
    std::swap(rings[0], rings[big]);
    size_t keep = 1;
    for (size_t i = 1; i < rings.size(); ++i) {
        bool all_inside = true;
        for (const PF& p : rings[i])
            if (!(signed_dist(p, rings, 1) >= 0.0)) {
                all_inside = false;
                break;
            }
        if (all_inside) std::swap(rings[i], rings[keep++]);
    }
    rings.resize(keep);
    double inf = std::numeric_limits<double>::infinity();
    double min_x = inf, max_x = -inf, min_y = inf, max_y = -inf;
    for (const PF& p : rings[0]) {
        min_x = std::fmin(min_x, p.first);
        max_x = std::fmax(max_x, p.first);
        min_y = std::fmin(min_y, p.second);
        max_y = std::fmax(max_y, p.second);
    }
    const double w = max_x - min_x, h = max_y - min_y;
    const double precision = std::fmax(w, h) / 100.0 * scale;
    const double cell = std::fmin(w, h), max_size = std::fmax(w, h);
    if (cell == 0.0) {
        out = PF(min_x, min_y);
        return true;
    }
    PF centroid;
    {
        double area = 0.0, cx = 0.0, cy = 0.0;
        const auto& r = rings[0];
        for (size_t i = 1; i < r.size(); ++i) {
            double c = r[i].first * r[i - 1].second - r[i - 1].first * r[i].second;
            cx += (r[i].first + r[i - 1].first) * c;
            cy += (r[i].second + r[i - 1].second) * c;
            area += c * 3.0;
        }
        centroid = (area == 0.0) ? r[0] : PF(cx / area, cy / area);
    }
    auto fitness = [&](const PF& c, double d) {
        if (d <= 0.0) return d;
        double dx = c.first - centroid.first, dy = c.second - centroid.second;
        return d * (1.0 - std::sqrt(dx * dx + dy * dy) / max_size);
    };
    auto cell_at = [&](const PF& c, double half) {
        double d = signed_dist(c, rings, rings.size());
        return HeapCell{c, half, fitness(c, d), fitness(c, d + half * 1.41421356237309504880168872420969808)};
    };
    MaxHeapRust heap;
    double half = cell / 2.0;
    for (double x = min_x; x < max_x; x += cell)
        for (double y = min_y; y < max_y; y += cell) heap.push(cell_at(PF(x + half, y + half), half));
    HeapCell best = cell_at(centroid, 0.0), cur;
    while (heap.pop(cur)) {
        if (cur.fit > best.fit) best = cur;
        if (cur.max_fit - best.fit <= precision) continue;
        half = cur.half / 2.0;
        for (double dx : {-1.0, 1.0})
            for (double dy : {-1.0, 1.0}) heap.push(cell_at(PF(cur.c.first + dx * half, cur.c.second + dy * half), half));
    }
    out = best.c;
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// layout of one tile's labels
// ---------------------------------------------------------------------------------------------------------
struct LabelStyleHost {
    osmr_label_style s;
    std::string key;
};
struct IconDim {
    uint32_t w, h;
};

class SegSink {  // Rasterizer::draw_line / draw_quad call stream (rasterizer.rs:27-107) -> segments
   public:
    std::vector<Seg>* out;
    double min_x, max_x, min_y, max_y;
    void reset(std::vector<Seg>* o) {
        out = o;
        min_x = min_y = std::numeric_limits<double>::infinity();
        max_x = max_y = -std::numeric_limits<double>::infinity();
    }
    void line(double x0, double y0, double x1, double y1) {
        if (y1 - y0 == 0.0) return;  // draw_line returns before touching anything (rasterizer.rs:30-32)
        out->push_back(Seg{x0, y0, x1, y1});
        // NaN-ignoring min / max without the libm call
        auto mn = [](double a, double b) { return (b < a || a != a) ? b : a; };
        auto mx = [](double a, double b) { return (b > a || a != a) ? b : a; };
        min_x = mn(min_x, mn(x0, x1));
        max_x = mx(max_x, mx(x0, x1));
        min_y = mn(min_y, mn(y0, y1));
        max_y = mx(max_y, mx(y0, y1));
    }
    // The flatness test of draw_quad (rasterizer.rs:86-107) compares platform-libm hypot values.  hypot is accurate to
    // < 1 ulp and sqrt(dx*dx + dy*dy) to < 2 ulp at pixel magnitudes, so the cheap form decides whenever the two sides
    // differ by more than 1e-12 relative; only near-ties pay for the three hypot calls.
    static bool flat_enough(double x0, double y0, double x1, double y1, double x2, double y2) {
        const double ax = std::fabs(x0 - x1), ay = std::fabs(y0 - y1), bx = std::fabs(x1 - x2), by = std::fabs(y1 - y2);
        const double cx = std::fabs(x0 - x2), cy = std::fabs(y0 - y2);
        const double lhs = std::sqrt(ax * ax + ay * ay) + std::sqrt(bx * bx + by * by);
        const double rhs = 1.0001 * std::sqrt(cx * cx + cy * cy);
        const double big = 1e150, tiny = 1e-150;
        const bool safe = lhs < big && rhs < big && lhs > tiny && rhs > tiny;  // also false for NaN
        if (safe && lhs < rhs * (1.0 - 1e-12)) return true;
        if (safe && lhs > rhs * (1.0 + 1e-12)) return false;
        return std::hypot(ax, ay) + std::hypot(bx, by) <= 1.0001 * std::hypot(cx, cy);
    }
    void quad(double x0, double y0, double x1, double y1, double x2, double y2) {
        if (flat_enough(x0, y0, x1, y1, x2, y2)) {
            line(x0, y0, x2, y2);
            return;
        }
        double ax = (x0 + x1) / 2.0, ay = (y0 + y1) / 2.0, bx = (x1 + x2) / 2.0, by = (y1 + y2) / 2.0;
        double mx = (ax + bx) / 2.0, my = (ay + by) / 2.0;
        quad(x0, y0, ax, ay, mx, my);
        quad(mx, my, bx, by, x2, y2);
    }
};

struct LayoutEnv {
    const BinView* bin;
    const TrueType* font;
    const std::vector<LabelStyleHost>* styles;
    const std::vector<IconDim>* icons;
};

inline bool is_ws(uint32_t c) {  // char::is_whitespace
    return c == 0x20 || (c >= 9 && c <= 13) || c == 0x85 || c == 0xa0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200a) || c == 0x2028 ||
           c == 0x2029 || c == 0x202f || c == 0x205f || c == 0x3000;
}
inline void decode_utf8(const char* s, size_t n, std::vector<uint32_t>& out) {
    out.clear();
    for (size_t i = 0; i < n;) {
        uint32_t c = (uint8_t)s[i];
        int extra = c < 0x80 ? 0 : (c >> 5) == 6 ? 1 : (c >> 4) == 14 ? 2 : 3;
        if (extra == 1) c &= 0x1f;
        if (extra == 2) c &= 0x0f;
        if (extra == 3) c &= 0x07;
        for (int k = 0; k < extra && i + 1 + k < n; ++k) c = (c << 6) | ((uint8_t)s[i + 1 + k] & 0x3f);
        out.push_back(c);
        i += 1 + extra;
    }
}

// entity tables
inline bool entity_rings(const BinView& g, uint32_t entity, const osmr_tile& t, double scale, Rings& rings) {
    rings.clear();
    auto ring = [&](uint32_t off, uint32_t len) {
        std::vector<PF> r;
        r.reserve(len);
        for (uint32_t i = 0; i < len; ++i) {
            double x, y;
            coords_rel(g, g.int_at(off + i), t, x, y);
            r.emplace_back(x * scale, y * scale);
        }
        rings.push_back(std::move(r));
    };
    if (entity & OSMR_AREA_MULTIPOLYGON) {
        uint32_t m = entity & ~OSMR_AREA_MULTIPOLYGON;
        if (m >= g.n_mps) return false;
        uint32_t po = BinView::u32(g.mps + (size_t)m * 24 + 8), pl = BinView::u32(g.mps + (size_t)m * 24 + 12);
        for (uint32_t k = 0; k < pl; ++k) {
            uint32_t pid = g.int_at(po + k);
            if (pid >= g.n_polys) return false;
            ring(BinView::u32(g.polys + (size_t)pid * 8), BinView::u32(g.polys + (size_t)pid * 8 + 4));
        }
    } else {
        if (entity >= g.n_ways) return false;
        ring(BinView::u32(g.ways + (size_t)entity * 24 + 8), BinView::u32(g.ways + (size_t)entity * 24 + 12));
    }
    return true;
}

// Lays out the labels of one tile.  Labels that can neither draw nor fail (no icon, no text geometry) are dropped:
// an empty successful generation has no effect on any pixel or on any later collision test.
// Returns false on malformed input (entity / style / icon index out of range).
inline bool layout_tile(const LayoutEnv& env, const osmr_tile& tile, const osmr_label* labels, uint32_t n_labels,
                        std::vector<LabelRec>& recs, std::vector<Seg>& segs) {
    const BinView& g = *env.bin;
    const TrueType& font = *env.font;
    const double gscale = (double)tile.scale;
    std::vector<uint32_t> chars;
    std::vector<int> gids;
    std::vector<double> widths;
    SegSink sink;
    Rings rings;
    for (uint32_t li = 0; li < n_labels; ++li) {
        const osmr_label& L = labels[li];
        if (L.style >= env.styles->size()) return false;
        const LabelStyleHost& st = (*env.styles)[L.style];
        const bool is_node = (L.entity & OSMR_LABEL_NODE) != 0 && !(L.entity & OSMR_AREA_MULTIPOLYGON);
        const bool is_mp = (L.entity & OSMR_AREA_MULTIPOLYGON) != 0;
        const uint32_t idx = L.entity & ~(OSMR_AREA_MULTIPOLYGON | OSMR_LABEL_NODE);
        uint32_t tags_off, tags_len;
        if (is_node) {
            if (idx >= g.n_nodes) return false;
            tags_off = BinView::u32(g.nodes + (size_t)idx * 32 + 24);
            tags_len = BinView::u32(g.nodes + (size_t)idx * 32 + 28);
        } else if (is_mp) {
            if (idx >= g.n_mps) return false;
            tags_off = BinView::u32(g.mps + (size_t)idx * 24 + 16);
            tags_len = BinView::u32(g.mps + (size_t)idx * 24 + 20);
        } else {
            if (idx >= g.n_ways) return false;
            tags_off = BinView::u32(g.ways + (size_t)idx * 24 + 16);
            tags_len = BinView::u32(g.ways + (size_t)idx * 24 + 20);
        }
        // label anchor (labelable.rs): lazily, it is expensive for areas
        bool anchor_done = false, anchor_ok = false;
        PF anchor;
        auto get_anchor = [&]() {
            if (anchor_done) return anchor_ok;
            anchor_done = true;
            if (is_node) {
                IPoint p = point_from_node(g, idx, tile, gscale);
                anchor = PF((double)p.x, (double)p.y);
                anchor_ok = true;
            } else {
                anchor_ok = entity_rings(g, is_mp ? (idx | OSMR_AREA_MULTIPOLYGON) : idx, tile, gscale, rings) &&
                            label_anchor_of_rings(rings, gscale, anchor);
            }
            return anchor_ok;
        };
        LabelRec rec{};
        rec.icon = -1;
        rec.bx0 = 1;
        rec.bx1 = 0;
        rec.by0 = 1;
        rec.by1 = 0;
        rec.seg_begin = (unsigned)segs.size();
        size_t y_offset = 0;
        // label_with_icon (labeler.rs:39-68)
        if (st.s.icon >= 0) {
            if ((size_t)st.s.icon >= env.icons->size()) return false;
            if (get_anchor()) {
                const IconDim& ic = (*env.icons)[(size_t)st.s.icon];
                rec.icon = st.s.icon;
                rec.ix = f64_as_i32(anchor.first - ((double)ic.w / 2.0));
                rec.iy = f64_as_i32(anchor.second - ((double)ic.h / 2.0));
                y_offset = ic.h / 2;
            }
        }
        // label_with_text -> TextPlacer::place (text_placer.rs:24-168)
        const char* text = nullptr;
        size_t text_len = 0;
        if ((st.s.flags & OSMR_LSTYLE_TEXT) && (st.s.flags & OSMR_LSTYLE_FONT_SIZE) &&
            g.tag(tags_off, tags_len, st.key.data(), st.key.size(), text, text_len)) {
            const double font_size = st.s.font_size * gscale;
            const double scale = (double)font.scale_for_pixel_height((float)font_size);
            decode_utf8(text, text_len, chars);
            gids.clear();
            widths.clear();
            double total_width = 0.0;
            int prev = -1;
            for (uint32_t c : chars) {
                int gi = font.glyph_index(c);
                double w = (double)font.advance(gi) * scale;
                if (prev >= 0) w += (double)font.kerning(prev, gi) * scale;
                total_width += w;
                prev = gi;
                gids.push_back(gi);
                widths.push_back(w);
            }
            const double ascent = (double)font.ascent() * scale, descent = (double)font.descent() * scale,
                         line_gap = (double)font.line_gap() * scale;
            sink.reset(&segs);
            auto emit_glyph = [&](int gi, auto tr) {  // Glyph::rasterize (text_placer.rs:211-231)
                const std::vector<GlyphVertex>* shape = font.shape(gi);
                if (!shape) return;
                PF from(0.0, 0.0);
                for (const GlyphVertex& v : *shape) {
                    PF to((double)v.x * scale, (double)v.y * scale);
                    if (v.type == VT_LINE) {
                        PF p1 = tr(from), p0 = tr(to);
                        sink.line(p0.first, p0.second, p1.first, p1.second);
                    } else if (v.type == VT_CURVE) {
                        PF p2 = tr(from), p1 = tr(PF((double)v.cx * scale, (double)v.cy * scale)), p0 = tr(to);
                        sink.quad(p0.first, p0.second, p1.first, p1.second, p2.first, p2.second);
                    }
                    from = to;
                }
            };
            unsigned pos = st.s.text_position ? st.s.text_position : ((is_node || is_mp) ? OSMR_TEXT_POS_CENTER : OSMR_TEXT_POS_LINE);
            if (pos == OSMR_TEXT_POS_LINE) {
                if (!is_node && !is_mp) {  // only ways have waypoints (labelable.rs:33-39)
                    uint32_t off = BinView::u32(g.ways + (size_t)idx * 24 + 8), len = BinView::u32(g.ways + (size_t)idx * 24 + 12);
                    std::vector<IPoint> pts;
                    pts.reserve(len);
                    for (uint32_t i = 0; i < len; ++i) pts.push_back(point_from_node(g, g.int_at(off + i), tile, gscale));
                    if (pts.size() >= 2) {
                        if (pts.front().x > pts.back().x) std::reverse(pts.begin(), pts.end());
                        double way_len = 0.0;
                        for (size_t i = 1; i < pts.size(); ++i) way_len += ipoint_dist(pts[i - 1], pts[i]);
                        if (!(total_width > way_len)) {
                            double cur = (way_len - total_width) / 2.0;
                            const double gcy = (descent + ascent) / 2.0;
                            for (size_t k = 0; k < gids.size(); ++k) {
                                const double gcx = widths[k] / 2.0;
                                double wx, wy, angle;  // compute_way_position (text_placer.rs:265-296)
                                {
                                    size_t i = 0;
                                    double left = cur + gcx;
                                    bool found = false;
                                    while (left > 0.0 && i + 1 < pts.size()) {
                                        double sd = ipoint_dist(pts[i], pts[i + 1]);
                                        if (sd >= left) {
                                            double ratio = left / ipoint_dist(pts[i], pts[i + 1]);
                                            wx = (double)pts[i].x + (double)wsub(pts[i + 1].x, pts[i].x) * ratio;
                                            wy = (double)pts[i].y + (double)wsub(pts[i + 1].y, pts[i].y) * ratio;
                                            angle = std::atan2((double)wsub(pts[i + 1].y, pts[i].y), (double)wsub(pts[i + 1].x, pts[i].x));
                                            found = true;
                                            break;
                                        }
                                        left -= sd;
                                        ++i;
                                    }
                                    if (!found) {
                                        size_t j = pts.size() - 2;
                                        wx = (double)pts.back().x;
                                        wy = (double)pts.back().y;
                                        angle = std::atan2((double)wsub(pts[j + 1].y, pts[j].y), (double)wsub(pts[j + 1].x, pts[j].x));
                                    }
                                }
                                const double sn = std::sin(-angle), cs = std::cos(-angle);
                                emit_glyph(gids[k], [&](const PF& p) {
                                    double tx = p.first - gcx, ty = p.second - gcy;
                                    return PF(wx + (tx * cs - ty * sn), wy - (ty * cs + tx * sn));
                                });
                                cur += widths[k];
                            }
                        }
                    }
                }
            } else if (get_anchor()) {
                struct Row {
                    size_t first, count;
                    double width;
                };
                std::vector<Row> rows;
                Row row{0, 0, 0.0};
                const double max_text_width = 256.0 / 8.0;  // TILE_SIZE / 8, not scaled (text_placer.rs:298)
                for (size_t k = 0; k < gids.size(); ++k) {
                    row.count += 1;
                    row.width += widths[k];
                    bool last = k + 1 == gids.size();
                    bool brk = is_ws(chars[k]) && (row.width + widths[k] > max_text_width);
                    if (brk || last) {
                        rows.push_back(row);
                        row = Row{k + 1, 0, 0.0};
                    }
                }
                const double row_h = ascent - descent + line_gap;
                double cur_y = anchor.second;
                if (y_offset > 0)
                    cur_y += (double)y_offset;
                else
                    cur_y -= (row_h * (double)rows.size()) / 2.0;
                for (const Row& r : rows) {
                    double cur_x = anchor.first - r.width / 2.0;
                    for (size_t k = r.first; k < r.first + r.count; ++k) {
                        const double baseline = cur_y + ascent, xo = cur_x;
                        emit_glyph(gids[k], [&](const PF& p) { return PF(xo + p.first, baseline - p.second); });
                        cur_x += widths[k];
                    }
                    cur_y += row_h;
                }
            }
            rec.seg_count = (unsigned)segs.size() - rec.seg_begin;
            if (rec.seg_count) {
                rec.bx0 = f64_as_i32(std::floor(sink.min_x));
                rec.bx1 = f64_as_i32(std::floor(sink.max_x)) + 1;  // the `s` column is one past the last `a` column
                rec.by0 = f64_as_i32(std::floor(sink.min_y));
                rec.by1 = f64_as_i32(std::floor(sink.max_y));
                const uint8_t* c = st.s.text_color;
                rec.rgb = (st.s.flags & OSMR_LSTYLE_TEXT_COLOR) ? ((unsigned)c[0] | ((unsigned)c[1] << 8) | ((unsigned)c[2] << 16)) : 0u;
            }
        }
        if (rec.icon >= 0 || rec.seg_count) recs.push_back(rec);
    }
    return true;
}

}  // namespace osmr_host

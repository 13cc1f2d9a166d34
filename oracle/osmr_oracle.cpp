// osmr_oracle.cpp -- CPU restatement of the reference's per-tile draw path.
//
// *** TEST INFRASTRUCTURE.  NOT PART OF THE PRODUCT. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library.  libosmr_b200.so never links, loads or calls it, and has no CPU fallback of its own.
//
// What it is: a line-for-line *behavioural* restatement (own code, own containers) of the reference's scalar
// algorithms, f64/i32/i64 arithmetic in the same operation order, compiled with -ffp-contract=off:
//   a1  src/tile.rs:88-106, src/draw/point.rs:11-35           lon/lat -> integer pixel (glibc tan/log, which
//                                                             is what Rust's f64::tan/ln call on Linux)
//   a2  src/draw/fill.rs:51-104                               per-edge all-octant Bresenham -> row spans
//   a3  src/draw/fill.rs:16-47                                per-row even-odd pairing, colour / pattern filler
//   a4  src/draw/line.rs:9-166                                Murphy thick line + perpendiculars + caps
//   a5  src/draw/opacity_calculator.rs:16-185                 feather / dash / cap opacity
//   a6  src/draw/tile_pixels.rs:57-129,150-152,191-199        3x3 canvas, pending pixel per generation
//   a7  src/draw/tile_pixels.rs:13-22,154-181,205-228         alpha-over, RGB export
//   a8  src/draw/drawer.rs:60-104,133-219                     Fill -> Casing -> Stroke ordering, style mapping
//   .bin views: src/geodata/reader.rs:264-335,444-483
// Parity pin: the reference's golden renders tests/rendered/*_expected.png (exact RGB outside the label
// pass, see tests/test_golden_tiles.py) and the doc-test vectors of src/tile.rs:23-29,77-87.
// The label pass (drawer.rs:106-126) lives in osmr_oracle_labels.inc (oracle only; the CUDA library has none yet).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <thread>
#include <vector>

#include "../include/osmr.h"

namespace {

// ---------------------------------------------------------------------------------------------------
// Rust numeric cast semantics
// ---------------------------------------------------------------------------------------------------
inline int32_t f64_as_i32(double v) {  // saturating, NaN -> 0
    if (v != v) return 0;
    if (v <= -2147483648.0) return std::numeric_limits<int32_t>::min();
    if (v >= 2147483647.0) return std::numeric_limits<int32_t>::max();
    return (int32_t)v;
}
inline uint8_t f64_as_u8(double v) {
    if (v != v) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}
inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

// ---------------------------------------------------------------------------------------------------
// .bin views (reader.rs:264-335)
// ---------------------------------------------------------------------------------------------------
inline uint32_t rd_u32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}
inline double rd_f64(const uint8_t* p) {
    double v;
    memcpy(&v, p, 8);
    return v;
}

struct Geodata {
    const uint8_t* nodes = nullptr;
    uint32_t n_nodes = 0;
    const uint8_t* ways = nullptr;
    uint32_t n_ways = 0;
    const uint8_t* polys = nullptr;
    uint32_t n_polys = 0;
    const uint8_t* mps = nullptr;
    uint32_t n_mps = 0;
    const uint8_t* ints = nullptr;
    uint32_t n_ints = 0;

    bool parse(const uint8_t* p, size_t len) {
        size_t pos = 0;
        auto table = [&](const uint8_t*& base, uint32_t& cnt, size_t rec) -> bool {
            if (pos + 4 > len) return false;
            cnt = rd_u32(p + pos);
            pos += 4;
            if (pos + (size_t)cnt * rec > len) return false;
            base = p + pos;
            pos += (size_t)cnt * rec;
            return true;
        };
        const uint8_t* tiles;
        uint32_t n_tiles;
        if (!table(nodes, n_nodes, 32)) return false;
        if (!table(ways, n_ways, 24)) return false;
        if (!table(polys, n_polys, 8)) return false;
        if (!table(mps, n_mps, 24)) return false;
        if (!table(tiles, n_tiles, 32)) return false;
        if (!table(ints, n_ints, 4)) return false;
        return true;
    }
    double lat(uint32_t n) const { return rd_f64(nodes + (size_t)n * 32 + 8); }
    double lon(uint32_t n) const { return rd_f64(nodes + (size_t)n * 32 + 16); }
    uint32_t int_at(uint32_t i) const { return rd_u32(ints + (size_t)i * 4); }
};

// ---------------------------------------------------------------------------------------------------
// a1: projection (tile.rs:88-106, point.rs:11-19)
// ---------------------------------------------------------------------------------------------------
const double PI = 3.14159265358979323846264338327950288;

inline void coords_to_xy(double lat, double lon, uint32_t zoom, double& ox, double& oy) {
    const double rads_per_deg = PI / 180.0;  // f64::to_radians
    double lat_rad = lat * rads_per_deg;
    double lon_rad = lon * rads_per_deg;
    double x = lon_rad + PI;
    double y = PI - std::log(std::tan((PI / 4.0) + (lat_rad / 2.0)));
    double dim = (double)(uint32_t)(256u * (1u << zoom));
    ox = (x / (2.0 * PI)) * dim;
    oy = (y / (2.0 * PI)) * dim;
}

struct Point {
    int32_t x, y;
    bool operator==(const Point& o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point& o) const { return !(*this == o); }
};

inline Point point_from_node(const Geodata& g, uint32_t node, const osmr_tile& t, double scale) {
    double x, y;
    coords_to_xy(g.lat(node), g.lon(node), t.zoom, x, y);
    x = x - (double)(uint32_t)(t.x * 256u);
    y = y - (double)(uint32_t)(t.y * 256u);
    return Point{f64_as_i32(std::round(x * scale)), f64_as_i32(std::round(y * scale))};
}

inline double point_dist(const Point& a, const Point& b) {  // point.rs:21-25
    double dx = (double)wsub(a.x, b.x);
    double dy = (double)wsub(a.y, b.y);
    return std::sqrt(dx * dx + dy * dy);
}

inline Point push_away_from(const Point& self, const Point& other, double by) {  // point.rs:27-35
    double dist = point_dist(self, other);
    double push = by / dist;
    auto coord = [&](int32_t our_c, int32_t other_c) {
        return wadd(our_c, f64_as_i32(std::round((double)wsub(our_c, other_c) * push)));
    };
    return Point{coord(self.x, other.x), coord(self.y, other.y)};
}

// ---------------------------------------------------------------------------------------------------
// a6/a7: compositor (tile_pixels.rs)
// ---------------------------------------------------------------------------------------------------
struct Rgba {
    double r, g, b, a;
};

inline double component_to_opacity(uint8_t c) { return (double)c / 255.0; }

inline Rgba from_color(const uint8_t rgb[3], double opacity) {  // tile_pixels.rs:13-22
    return Rgba{opacity * component_to_opacity(rgb[0]), opacity * component_to_opacity(rgb[1]),
                opacity * component_to_opacity(rgb[2]), opacity};
}
inline Rgba from_components(uint8_t r, uint8_t g, uint8_t b, uint8_t a) {  // tile_pixels.rs:24-26
    uint8_t rgb[3] = {r, g, b};
    return from_color(rgb, component_to_opacity(a));
}

struct NextPixel {
    Rgba color;
    size_t generation;
    bool some;
};

struct TilePixels {
    int32_t bb_min_x, bb_max_x, bb_min_y, bb_max_y;
    int32_t lbb_min_x, lbb_max_x, lbb_min_y, lbb_max_y;
    size_t scaled_tile_size, scaled_ext;
    std::vector<Rgba> pixels;
    std::vector<NextPixel> next_pixels;
    size_t generation = 0;

    explicit TilePixels(size_t scale) {  // tile_pixels.rs:57-87
        scaled_tile_size = 256 * scale;
        int32_t s = (int32_t)scaled_tile_size;
        bb_min_x = 0;
        bb_max_x = s - 1;
        bb_min_y = 0;
        bb_max_y = s - 1;
        lbb_min_x = bb_min_x - s;
        lbb_max_x = bb_max_x + s;
        lbb_min_y = bb_min_y - s;
        lbb_max_y = bb_max_y + s;
        scaled_ext = 768 * scale;
        pixels.assign(scaled_ext * scaled_ext, Rgba{0, 0, 0, 1.0});
        next_pixels.assign(scaled_ext * scaled_ext, NextPixel{Rgba{0, 0, 0, 0}, 0, false});
    }

    void reset(bool has_canvas, const uint8_t rgb[3]) {  // tile_pixels.rs:89-105
        Rgba init = has_canvas ? from_color(rgb, 1.0) : Rgba{0, 0, 0, 1.0};
        for (auto& p : pixels) p = init;
        for (auto& n : next_pixels) n.some = false;
        generation = 0;
    }

    bool idx_of(int32_t x, int32_t y, size_t& idx) const {  // tile_pixels.rs:191-203 (for_labels=false)
        if (x < bb_min_x || x > bb_max_x || y < bb_min_y || y > bb_max_y) return false;
        size_t lx = (size_t)(x - lbb_min_x), ly = (size_t)(y - lbb_min_y);
        idx = ly * scaled_ext + lx;
        return true;
    }

    void blend_pixel(size_t idx) {  // tile_pixels.rs:205-223 (for_labels=false)
        NextPixel& n = next_pixels[idx];
        if (n.some) {
            Rgba& o = pixels[idx];
            double inv = 1.0 - n.color.a;
            Rgba res{n.color.r + inv * o.r, n.color.g + inv * o.g, n.color.b + inv * o.b, n.color.a + inv * o.a};
            o = res;
        }
        n.some = false;
    }

    void set_pixel(int32_t x, int32_t y, const Rgba& c) {  // tile_pixels.rs:107-129
        size_t idx;
        if (!idx_of(x, y, idx)) return;
        bool same = false;
        NextPixel& n = next_pixels[idx];
        if (n.some && n.generation == generation) {
            if (c.a > n.color.a) n.color = c;
            same = true;
        }
        if (!same) {
            blend_pixel(idx);
            next_pixels[idx] = NextPixel{c, generation, true};
        }
    }

    void bump_generation() { generation += 1; }

    void blend_unfinished_pixels() {  // tile_pixels.rs:154-158
        for (size_t i = 0; i < next_pixels.size(); ++i) blend_pixel(i);
    }

    void to_rgb(uint8_t* out) const {  // tile_pixels.rs:164-181
        size_t o = 0;
        for (size_t y = scaled_tile_size; y < 2 * scaled_tile_size; ++y)
            for (size_t x = scaled_tile_size; x < 2 * scaled_tile_size; ++x) {
                const Rgba& p = pixels[y * scaled_ext + x];
                auto post = [&](double v) {
                    double mul = (p.a == 0.0) ? 0.0 : v / p.a;
                    return f64_as_u8(255.0 * mul);
                };
                out[o++] = post(p.r);
                out[o++] = post(p.g);
                out[o++] = post(p.b);
            }
    }
};

// ---------------------------------------------------------------------------------------------------
// icons (icon.rs:29-62)
// ---------------------------------------------------------------------------------------------------
struct Icon {
    size_t width = 0, height = 0;
    std::vector<Rgba> px;
};

// ---------------------------------------------------------------------------------------------------
// a2/a3: fill (fill.rs)
// ---------------------------------------------------------------------------------------------------
struct EdgeRec {
    size_t edge_idx;
    int32_t x_min, x_max;
    bool poisoned;
};
struct RowEdges {
    std::vector<EdgeRec> v;  // insertion order == edge order (IndexMap<usize, Edge>)
};

// fill.rs:51-104.  rows[] covers y in [min_y, max_y]; `touched` lists rows in first-insertion order.
void fill_draw_line(size_t edge_idx, const Point& p1, const Point& p2, std::vector<RowEdges>& rows,
                    std::vector<int32_t>& touched, int32_t min_y, int32_t max_y) {
    int32_t dx = std::abs(wsub(p2.x, p1.x));
    int32_t dy = -std::abs(wsub(p2.y, p1.y));
    int32_t sx = (p1.x < p2.x) ? 1 : -1;
    int32_t sy = (p1.y < p2.y) ? 1 : -1;
    int32_t err = wadd(dx, dy);
    Point cur = p1;
    for (;;) {
        bool is_start = cur == p1;
        bool is_end = cur == p2;
        bool poisoned = is_start ? (p1.y <= p2.y) : (is_end ? (p2.y <= p1.y) : false);
        if (cur.y >= min_y && cur.y <= max_y) {
            RowEdges& r = rows[(size_t)(cur.y - min_y)];
            EdgeRec* e = nullptr;
            // entry(edge_idx): edges are walked one after another, so a hit can only be the last record
            for (size_t k = r.v.size(); k-- > 0;) {
                if (r.v[k].edge_idx == edge_idx) {
                    e = &r.v[k];
                    break;
                }
                if (r.v[k].edge_idx < edge_idx) break;
            }
            if (!e) {
                if (r.v.empty()) touched.push_back(cur.y);
                r.v.push_back(EdgeRec{edge_idx, cur.x, cur.x, poisoned});
                e = &r.v.back();
            }
            e->x_min = std::min(e->x_min, cur.x);
            e->x_max = std::max(e->x_max, cur.x);
            e->poisoned = e->poisoned || poisoned;
        }
        if (is_end) break;
        int32_t e2 = wmul(2, err);
        if (e2 >= dy) {
            err = wadd(err, dy);
            cur.x = wadd(cur.x, sx);
        }
        if (e2 <= dx) {
            err = wadd(err, dx);
            cur.y = wadd(cur.y, sy);
        }
    }
}

struct Filler {
    const uint8_t* color = nullptr;  // Filler::Color
    const Icon* icon = nullptr;      // Filler::Image
};

// fill.rs:16-47
void fill_contour(const std::vector<std::pair<Point, Point>>& pairs, const Filler& filler, double opacity,
                  TilePixels& px) {
    std::vector<RowEdges> rows((size_t)(px.bb_max_y - px.bb_min_y + 1));
    std::vector<int32_t> touched;
    for (size_t idx = 0; idx < pairs.size(); ++idx)
        fill_draw_line(idx, pairs[idx].first, pairs[idx].second, rows, touched, px.bb_min_y, px.bb_max_y);

    std::vector<const EdgeRec*> good;
    for (int32_t y : touched) {
        const RowEdges& r = rows[(size_t)(y - px.bb_min_y)];
        good.clear();
        for (const EdgeRec& e : r.v)
            if (!e.poisoned) good.push_back(&e);
        std::stable_sort(good.begin(), good.end(), [](const EdgeRec* a, const EdgeRec* b) { return a->x_min < b->x_min; });
        size_t idx = 0;
        while (idx + 1 < good.size()) {
            const EdgeRec* e1 = good[idx];
            const EdgeRec* e2 = good[idx + 1];
            int32_t from_x = std::max(e1->x_min, px.bb_min_x);
            int32_t to_x = wadd(std::min(e2->x_max, px.bb_max_x), 1);
            for (int32_t x = from_x; x < to_x; ++x) {
                Rgba c;
                if (filler.color) {
                    c = from_color(filler.color, opacity);
                } else {
                    size_t ix = (size_t)x % filler.icon->width;
                    size_t iy = (size_t)y % filler.icon->height;
                    c = filler.icon->px[iy * filler.icon->width + ix];
                }
                px.set_pixel(x, y, c);
            }
            idx += 2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// a5: opacity calculator (opacity_calculator.rs)
// ---------------------------------------------------------------------------------------------------
struct DashSegment {
    double start_from, start_to, end_from, end_to, opacity_mul;
    bool has_original;
    double orig_a, orig_b;
};

inline bool is_non_trivial_cap(uint8_t cap) { return cap == OSMR_CAP_SQUARE || cap == OSMR_CAP_ROUND; }

struct OpacityCalculator {
    double half_line_width = 0;
    std::vector<DashSegment> dashes;
    double total_dash_len = 0;
    double traveled_distance = 0;

    // opacity_calculator.rs:16-30, 98-143.  dashes == nullptr encodes Option::None.
    OpacityCalculator(double hw, const double* d, size_t n, bool has_dashes, uint8_t cap) : half_line_width(hw) {
        double len_before = 0.0;
        if (has_dashes) {
            // "(0..dashes.len()).chain(0..1)"; dashes[0] on an empty list panics in the reference -- an empty
            // dash list cannot come out of the parser (a Numbers value has >= 1 element).
            for (size_t k = 0; k < n + 1; ++k) {
                size_t idx = (k < n) ? k : 0;
                if (n == 0) break;
                double dash = d[idx];
                double start = len_before;
                if (idx != 0 || dashes.empty()) len_before += dash;
                if (idx % 2 != 0) continue;
                double end = start + dash;
                bool has_orig = cap == OSMR_CAP_ROUND;
                double oa = start, ob = end;
                if (is_non_trivial_cap(cap)) {
                    start -= hw;
                    end += hw;
                }
                double mid = (start + end) / 2.0;
                DashSegment s;
                s.start_from = std::fmin(start - 0.5, mid - 1.0);
                s.start_to = std::fmin(start + 0.5, mid);
                s.end_from = std::fmax(end - 0.5, mid);
                s.end_to = std::fmax(end + 0.5, mid + 1.0);
                s.opacity_mul = std::fmin(end - start, 1.0);
                s.has_original = has_orig;
                s.orig_a = oa;
                s.orig_b = ob;
                dashes.push_back(s);
            }
        }
        total_dash_len = len_before;
    }

    static double by_center_distance(double cd, double hw) {  // :171-185
        double feather_from = std::fmax(hw - 0.5, 0.0);
        double feather_to = std::fmax(hw + 0.5, 1.0);
        double feather_dist = feather_to - feather_from;
        double opacity_mul = std::fmin(2.0 * hw, 1.0);
        double v;
        if (cd < feather_from)
            v = 1.0;
        else if (cd < feather_to)
            v = (feather_to - cd) / feather_dist;
        else
            v = 0.0;
        return opacity_mul * v;
    }

    void calculate(double center_distance, double start_distance, double& opacity, bool& is_in_line) const {  // :32-43
        double sd_opacity;
        bool has_cap = false;
        double cap = 0.0;
        if (dashes.empty()) {  // :50-55
            sd_opacity = 1.0;
        } else {
            double dist_rem = traveled_distance + start_distance;
            if (total_dash_len > 0.0) dist_rem = std::fmod(dist_rem, total_dash_len);
            double op_acc = 0.0;
            for (const DashSegment& s : dashes) {
                // get_opacity_by_segment :145-157
                bool some;
                double base = 0.0;
                if (dist_rem < s.start_from || dist_rem > s.end_to) {
                    some = false;
                } else if (dist_rem <= s.start_to) {
                    some = true;
                    base = (dist_rem - s.start_from) / (s.start_to - s.start_from);
                } else if (dist_rem < s.end_from) {
                    some = true;
                    base = 1.0;
                } else {
                    some = true;
                    base = (s.end_to - dist_rem) / (s.end_to - s.end_from);
                }
                if (some) {
                    double op = s.opacity_mul * base;
                    op_acc = std::fmax(op_acc, op);
                    if (s.has_original) {  // get_distance_in_cap :159-169
                        double dcap;
                        if (dist_rem < s.orig_a)
                            dcap = s.orig_a - dist_rem;
                        else if (dist_rem <= s.orig_b)
                            dcap = 0.0;
                        else
                            dcap = dist_rem - s.orig_b;
                        if (!has_cap || dcap < cap) {
                            has_cap = true;
                            cap = dcap;
                        }
                    }
                }
            }
            sd_opacity = op_acc;
        }
        double cap_dist = has_cap ? cap : 0.0;
        double hw = std::sqrt(half_line_width * half_line_width - cap_dist * cap_dist);
        double cd = by_center_distance(center_distance, hw);
        opacity = std::fmin(sd_opacity, cd);
        is_in_line = cd > 0.0;
    }
};

// ---------------------------------------------------------------------------------------------------
// a4: thick lines (line.rs)
// ---------------------------------------------------------------------------------------------------
void line_draw_line(const Point& p1, const Point& p2, const uint8_t color[3], double initial_opacity,
                    const OpacityCalculator& oc, TilePixels& px) {  // line.rs:65-158
    if (p1 == p2) return;
    auto get_inc = [](int32_t from, int32_t to) { return from <= to ? 1 : -1; };
    int32_t dx = std::abs(wsub(p2.x, p1.x)), dy = std::abs(wsub(p2.y, p1.y));
    bool swap = dx > dy;
    // (mn, mx) are the minor / major running coordinates of the main line
    int32_t mn = swap ? p1.y : p1.x, mx = swap ? p1.x : p1.y;
    int32_t mn_last = swap ? p2.y : p2.x, mx_last = swap ? p2.x : p2.y;
    int32_t mn_delta = swap ? dy : dx, mx_delta = swap ? dx : dy;
    int32_t inc_x = get_inc(p1.x, p2.x), inc_y = get_inc(p1.y, p2.y);
    int32_t mn_inc = swap ? inc_y : inc_x, mx_inc = swap ? inc_x : inc_y;

    int32_t error = 0, p_error = 0;
    auto update_error = [&](int32_t& e) {
        bool corrected = false;
        if (wadd(e, wmul(2, mn_delta)) > mx_delta) {
            e = wsub(e, wmul(2, mx_delta));
            corrected = true;
        }
        e = wadd(e, wmul(2, mn_delta));
        return corrected;
    };

    int64_t numer_const = (int64_t)p2.x * (int64_t)p1.y - (int64_t)p2.y * (int64_t)p1.x;
    int64_t sdx = (int64_t)p2.x - (int64_t)p1.x, sdy = (int64_t)p2.y - (int64_t)p1.y;
    double dxf = (double)dx, dyf = (double)dy;
    double denom = std::sqrt(dyf * dyf + dxf * dxf);

    auto draw_perpendiculars = [&](int32_t mn0, int32_t mx0, int32_t perr) {
        auto one = [&](int32_t mul) {
            int32_t p_mn = mx0;
            int32_t p_mx = mn0;
            int32_t e = wmul(mul, perr);
            for (;;) {
                int32_t perp_x = swap ? p_mn : p_mx;
                int32_t perp_y = swap ? p_mx : p_mn;
                Point cur{perp_x, perp_y};
                int64_t non_const = sdy * (int64_t)perp_x - sdx * (int64_t)perp_y;
                int64_t raw = numer_const + non_const;
                double center_dist = std::fabs((double)raw) / denom;
                double long_start = point_dist(cur, p1);
                double short_start = std::sqrt(std::fmax(long_start * long_start - center_dist * center_dist, 0.0));
                double opacity;
                bool in_line;
                oc.calculate(center_dist, short_start, opacity, in_line);
                if (!in_line) break;
                Rgba c = from_color(color, initial_opacity * opacity);
                px.set_pixel(cur.x, cur.y, c);
                if (update_error(e)) p_mn = wsub(p_mn, wmul(mul, mx_inc));
                p_mx = wadd(p_mx, wmul(mul, mn_inc));
            }
        };
        one(1);
        one(-1);
    };

    for (;;) {
        draw_perpendiculars(mn, mx, p_error);
        if (mn == mn_last && mx == mx_last) break;
        if (update_error(error)) {
            mn = wadd(mn, mn_inc);
            if (update_error(p_error)) draw_perpendiculars(mn, mx, p_error);
        }
        mx = wadd(mx, mx_inc);
    }
}

void draw_lines(const std::vector<std::pair<Point, Point>>& pairs, double width, const uint8_t color[3],
                double opacity, const double* dashes, size_t n_dashes, bool has_dashes, uint8_t line_cap,
                bool use_caps_for_dashes, TilePixels& px) {  // line.rs:9-61
    double half_width = width / 2.0;
    uint8_t cap_for_dashes = use_caps_for_dashes ? line_cap : (uint8_t)OSMR_CAP_NONE;
    OpacityCalculator oc(half_width, dashes, n_dashes, has_dashes, cap_for_dashes);
    const double zero = 0.0;
    OpacityCalculator oc_caps(half_width, &zero, 1, true, line_cap);
    bool has_caps = is_non_trivial_cap(line_cap);
    bool first = true;
    for (size_t i = 0; i < pairs.size(); ++i) {
        const Point& p1 = pairs[i].first;
        const Point& p2 = pairs[i].second;
        line_draw_line(p1, p2, color, opacity, oc, px);
        oc.traveled_distance += point_dist(p1, p2);
        if (p1 != p2 && has_caps) {
            if (first) {
                Point cap_end = push_away_from(p1, p2, half_width);
                line_draw_line(p1, cap_end, color, opacity, oc_caps, px);
            }
            if (i + 1 == pairs.size()) {
                Point cap_end = push_away_from(p2, p1, half_width);
                line_draw_line(p2, cap_end, color, opacity, oc_caps, px);
            }
        }
        first = false;
    }
}

// ---------------------------------------------------------------------------------------------------
// a8: drawer (drawer.rs:60-104,133-219) + point pairs (point_pairs.rs:11-41)
// ---------------------------------------------------------------------------------------------------
struct World {
    Geodata geo;
    const osmr_style* styles;
    uint32_t n_styles;
    const double* dashes;
    uint32_t n_dashes;
    std::vector<Icon> icons;
};

void ring_pairs(const Geodata& g, uint32_t off, uint32_t len, const osmr_tile& t, double scale,
                std::vector<std::pair<Point, Point>>& out) {
    for (uint32_t i = 1; i < len; ++i) {
        // the reference projects both ends of every pair again (point_pairs.rs:13-21)
        Point a = point_from_node(g, g.int_at(off + i - 1), t, scale);
        Point b = point_from_node(g, g.int_at(off + i), t, scale);
        out.emplace_back(a, b);
    }
}

bool area_pairs(const World& w, const osmr_styled_area& a, const osmr_tile& t, double scale,
                std::vector<std::pair<Point, Point>>& out) {
    out.clear();
    const Geodata& g = w.geo;
    if (a.entity & OSMR_AREA_MULTIPOLYGON) {
        uint32_t m = a.entity & ~OSMR_AREA_MULTIPOLYGON;
        if (m >= g.n_mps) return false;
        uint32_t poff = rd_u32(g.mps + (size_t)m * 24 + 8), plen = rd_u32(g.mps + (size_t)m * 24 + 12);
        for (uint32_t k = 0; k < plen; ++k) {
            uint32_t pid = g.int_at(poff + k);
            if (pid >= g.n_polys) return false;
            ring_pairs(g, rd_u32(g.polys + (size_t)pid * 8), rd_u32(g.polys + (size_t)pid * 8 + 4), t, scale, out);
        }
    } else {
        if (a.entity >= g.n_ways) return false;
        ring_pairs(g, rd_u32(g.ways + (size_t)a.entity * 24 + 8), rd_u32(g.ways + (size_t)a.entity * 24 + 12), t, scale, out);
    }
    return true;
}

enum DrawType { FILL = 0, CASING = 1, STROKE = 2 };

void draw_one_area(const World& w, TilePixels& px, const osmr_tile& t, double scale, const osmr_styled_area& a,
                   DrawType dt, bool use_caps_for_dashes, std::vector<std::pair<Point, Point>>& pairs,
                   std::vector<double>& scaled) {  // drawer.rs:156-219
    const osmr_style& s = w.styles[a.style];
    auto scale_dashes = [&](uint32_t off, uint32_t len) {
        scaled.clear();
        for (uint32_t i = 0; i < len; ++i) scaled.push_back(w.dashes[off + i] * scale);
    };
    switch (dt) {
        case FILL: {
            double opacity = (s.flags & OSMR_STYLE_FILL_OPACITY) ? s.fill_opacity : 1.0;
            if (s.flags & OSMR_STYLE_FILL_COLOR) {
                area_pairs(w, a, t, scale, pairs);
                Filler f;
                f.color = s.fill_color;
                fill_contour(pairs, f, opacity, px);
            } else if (s.flags & OSMR_STYLE_FILL_IMAGE) {
                if (s.fill_image >= 0 && (uint32_t)s.fill_image < w.icons.size()) {
                    area_pairs(w, a, t, scale, pairs);
                    Filler f;
                    f.icon = &w.icons[(size_t)s.fill_image];
                    fill_contour(pairs, f, opacity, px);
                }
            }
            break;
        }
        case CASING:
            if ((s.flags & OSMR_STYLE_CASING_COLOR) && (s.flags & OSMR_STYLE_CASING_WIDTH)) {
                area_pairs(w, a, t, scale, pairs);
                bool hd = (s.flags & OSMR_STYLE_CASING_DASHES) != 0;
                if (hd) scale_dashes(s.casing_dashes_off, s.casing_dashes_len);
                draw_lines(pairs, s.casing_width * scale, s.casing_color, 1.0, scaled.data(), hd ? scaled.size() : 0,
                           hd, s.casing_line_cap, use_caps_for_dashes, px);
            }
            break;
        case STROKE:
            if (s.flags & OSMR_STYLE_COLOR) {
                area_pairs(w, a, t, scale, pairs);
                bool hd = (s.flags & OSMR_STYLE_DASHES) != 0;
                if (hd) scale_dashes(s.dashes_off, s.dashes_len);
                double width = scale * ((s.flags & OSMR_STYLE_WIDTH) ? s.width : 1.0);
                double opacity = (s.flags & OSMR_STYLE_OPACITY) ? s.opacity : 1.0;
                draw_lines(pairs, width, s.color, opacity, scaled.data(), hd ? scaled.size() : 0, hd, s.line_cap,
                           use_caps_for_dashes, px);
            }
            break;
    }
    px.bump_generation();
}

void draw_area_passes(const World& w, TilePixels& px, const osmr_tile& t, const osmr_styled_area* areas, uint32_t n_areas,
                      const uint8_t canvas[3], uint32_t flags, uint32_t gen_limit) {
    px.reset((flags & OSMR_DRAW_HAS_CANVAS_COLOR) != 0, canvas);  // drawer.rs:70
    double scale = (double)t.scale;
    bool caps = (flags & OSMR_DRAW_USE_CAPS_FOR_DASHES) != 0;
    std::vector<std::pair<Point, Point>> pairs;
    std::vector<double> scaled;
    uint32_t gen = 0;
    for (int pass = 0; pass < 3; ++pass) {  // drawer.rs:94-100
        for (uint32_t i = 0; i < n_areas; ++i, ++gen) {
            if (gen >= gen_limit) break;
            bool is_mp = (areas[i].entity & OSMR_AREA_MULTIPOLYGON) != 0;
            if (is_mp && pass != FILL) continue;  // draw_areas: multipolygons only in the Fill pass (:144-150)
            draw_one_area(w, px, t, scale, areas[i], (DrawType)pass, caps, pairs, scaled);
        }
    }
    px.blend_unfinished_pixels();  // drawer.rs:104
}

void draw_to_pixels(const World& w, TilePixels& px, const osmr_tile& t, const osmr_styled_area* areas,
                    uint32_t n_areas, const uint8_t canvas[3], uint32_t flags, uint32_t gen_limit, uint8_t* out) {
    draw_area_passes(w, px, t, areas, n_areas, canvas, flags, gen_limit);
    px.to_rgb(out);  // drawer.rs:128 (label pass: see osmr_oracle_draw_tiles_labels)
}

#include "osmr_oracle_labels.inc"

bool build_world(World& w, const void* bin, size_t bin_len, const osmr_style* styles, uint32_t n_styles,
                 const double* dashes, uint32_t n_dashes, const osmr_icon* icons, uint32_t n_icons) {
    if (!w.geo.parse((const uint8_t*)bin, bin_len)) return false;
    w.styles = styles;
    w.n_styles = n_styles;
    w.dashes = dashes;
    w.n_dashes = n_dashes;
    w.icons.resize(n_icons);
    for (uint32_t i = 0; i < n_icons; ++i) {
        Icon& ic = w.icons[i];
        ic.width = icons[i].width;
        ic.height = icons[i].height;
        ic.px.resize(ic.width * ic.height);
        for (size_t k = 0; k < ic.px.size(); ++k) {
            const uint8_t* p = icons[i].rgba + 4 * k;
            ic.px[k] = from_components(p[0], p[1], p[2], p[3]);
        }
    }
    return true;
}

}  // namespace

extern "C" {

// Render n_tiles tiles exactly like n_tiles calls of the reference's Drawer::draw_to_pixels (area passes).
// out: n_tiles * D*D*3 RGB bytes.  n_threads > 1 deals tiles round-robin to worker threads, each with a private
// TilePixels -- the threading model of src/http_server.rs:50-83,105-108.  gen_limit: stop after that many
// generations (debugging aid; UINT32_MAX = all).
int osmr_oracle_draw_tiles(const void* bin, size_t bin_len, const osmr_style* styles, uint32_t n_styles,
                           const double* dashes, uint32_t n_dashes, const osmr_icon* icons, uint32_t n_icons,
                           const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                           const osmr_styled_area* areas, const uint8_t canvas_rgb[3], uint32_t flags,
                           uint32_t gen_limit, int n_threads, uint8_t* out_rgb) {
    World w;
    if (!build_world(w, bin, bin_len, styles, n_styles, dashes, n_dashes, icons, n_icons)) return OSMR_E_INVALID;
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (tiles[t].scale == 0 || tiles[t].scale > 8 || tiles[t].zoom > 24) return OSMR_E_INVALID;
        for (uint32_t i = area_begin[t]; i < area_begin[t + 1]; ++i)
            if (areas[i].style >= n_styles) return OSMR_E_INVALID;
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<size_t> out_off(n_tiles + 1, 0);
    for (uint32_t t = 0; t < n_tiles; ++t) {
        size_t d = 256 * (size_t)tiles[t].scale;
        out_off[t + 1] = out_off[t] + d * d * 3;
    }
    auto worker = [&](int tid) {
        TilePixels* px = nullptr;
        size_t cur_scale = 0;
        for (uint32_t t = (uint32_t)tid; t < n_tiles; t += (uint32_t)n_threads) {
            if (!px || cur_scale != tiles[t].scale) {  // http_server.rs:156-160
                delete px;
                px = new TilePixels(tiles[t].scale);
                cur_scale = tiles[t].scale;
            }
            draw_to_pixels(w, *px, tiles[t], areas + area_begin[t], area_begin[t + 1] - area_begin[t], canvas_rgb,
                           flags, gen_limit, out_rgb + out_off[t]);
        }
        delete px;
    };
    if (n_threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(worker, i);
        for (auto& x : th) x.join();
    }
    return OSMR_OK;
}

// a1 in isolation: Point::from_node for every node of the image.  out_xy: n_nodes*2 int32.
int osmr_oracle_project_nodes(const void* bin, size_t bin_len, const osmr_tile* tile, int32_t* out_xy) {
    Geodata g;
    if (!g.parse((const uint8_t*)bin, bin_len)) return OSMR_E_INVALID;
    for (uint32_t n = 0; n < g.n_nodes; ++n) {
        Point p = point_from_node(g, n, *tile, (double)tile->scale);
        out_xy[2 * n] = p.x;
        out_xy[2 * n + 1] = p.y;
    }
    return OSMR_OK;
}

// tile.rs:88-101 (doc-test vectors tile.rs:77-87)
void osmr_oracle_coords_to_xy(double lat, double lon, uint32_t zoom, double out[2]) {
    coords_to_xy(lat, lon, zoom, out[0], out[1]);
}

// fill.rs:51-104 for ONE edge on rows [min_y, max_y]: out_rows[(y-min_y)*3 + {0,1,2}] = x_min, x_max, state
// with state 0 = row not touched, 1 = span, 2 = poisoned span.  Used to pin the closed form of the CUDA path.
void osmr_oracle_fill_edge_rows(int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t min_y, int32_t max_y,
                                int32_t* out_rows) {
    std::vector<RowEdges> rows((size_t)(max_y - min_y + 1));
    std::vector<int32_t> touched;
    fill_draw_line(0, Point{x1, y1}, Point{x2, y2}, rows, touched, min_y, max_y);
    for (int32_t y = min_y; y <= max_y; ++y) {
        const RowEdges& r = rows[(size_t)(y - min_y)];
        int32_t* o = out_rows + (size_t)(y - min_y) * 3;
        if (r.v.empty()) {
            o[0] = o[1] = o[2] = 0;
        } else {
            o[0] = r.v[0].x_min;
            o[1] = r.v[0].x_max;
            o[2] = r.v[0].poisoned ? 2 : 1;
        }
    }
}

// Full Drawer::draw_to_pixels including the label pass (drawer.rs:106-126).  labels of tile t:
// labels[label_begin[t] .. label_begin[t+1]) in label-generation order.  CPU only; pins the oracle (and the upstream
// restatement) to the reference goldens on EVERY pixel.
int osmr_oracle_draw_tiles_labels(const void* bin, size_t bin_len, const osmr_style* styles, uint32_t n_styles,
                                  const double* dashes, uint32_t n_dashes, const osmr_icon* icons, uint32_t n_icons,
                                  const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                                  const osmr_styled_area* areas, const uint8_t canvas_rgb[3], uint32_t flags,
                                  const void* font, size_t font_len, const osmr_icon* label_icons, uint32_t n_label_icons,
                                  const uint32_t* label_begin, const LabelRec* labels, const char* texts, int n_threads,
                                  uint8_t* out_rgb) {
    World w;
    if (!build_world(w, bin, bin_len, styles, n_styles, dashes, n_dashes, icons, n_icons)) return OSMR_E_INVALID;
    LabelWorld lw;
    lw.w = &w;
    lw.texts = texts;
    if (!lw.font.init((const uint8_t*)font, font_len)) return OSMR_E_INVALID;
    lw.icons.resize(n_label_icons);
    for (uint32_t i = 0; i < n_label_icons; ++i) {
        Icon& ic = lw.icons[i];
        ic.width = label_icons[i].width;
        ic.height = label_icons[i].height;
        ic.px.resize(ic.width * ic.height);
        for (size_t k = 0; k < ic.px.size(); ++k) {
            const uint8_t* p = label_icons[i].rgba + 4 * k;
            ic.px[k] = from_components(p[0], p[1], p[2], p[3]);
        }
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<size_t> out_off(n_tiles + 1, 0);
    for (uint32_t t = 0; t < n_tiles; ++t) {
        size_t d = 256 * (size_t)tiles[t].scale;
        out_off[t + 1] = out_off[t] + d * d * 3;
    }
    auto worker = [&](int tid) {
        TilePixels* px = nullptr;
        size_t cur_scale = 0;
        for (uint32_t t = (uint32_t)tid; t < n_tiles; t += (uint32_t)n_threads) {
            if (!px || cur_scale != tiles[t].scale) {
                delete px;
                px = new TilePixels(tiles[t].scale);
                cur_scale = tiles[t].scale;
            }
            draw_area_passes(w, *px, tiles[t], areas + area_begin[t], area_begin[t + 1] - area_begin[t], canvas_rgb, flags, 0xffffffffu);
            draw_labels(lw, *px, tiles[t], labels + label_begin[t], label_begin[t + 1] - label_begin[t]);
            px->to_rgb(out_rgb + out_off[t]);
        }
        delete px;
    };
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(worker, i);
    for (auto& x : th) x.join();
    return OSMR_OK;
}

uint32_t osmr_oracle_abi_version(void) { return 2; }

}  // extern "C"

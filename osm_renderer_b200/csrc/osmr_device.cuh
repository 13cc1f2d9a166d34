// osmr_device.cuh -- device-side arithmetic shared by the kernels of libosmr_b200.so.
//
// Everything here reproduces the reference's scalar arithmetic bit-for-bit (f64 / i32 / i64, IEEE basic
// operations in the reference's order).  The translation unit MUST be compiled with -fmad=false: the
// reference has no fused multiply-adds, and e.g. `src + (1-a)*dst` contracted to an FMA changes the
// truncated 8-bit result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "osmr.h"

#ifndef OSMR_COUNT
#define OSMR_COUNT(name, n) ((void)0)  // event counters exist only in the host emulator build (tests/emu/)
#endif

namespace osmr {

constexpr double kPi = 3.14159265358979323846264338327950288;

// component_to_opacity (tile_pixels.rs:226-228): f64::from(c) / 255.0 for every u8, filled by the host with the same
// IEEE division (osmr_ctx_create).  A table look-up instead of an inline f64 division keeps ~4 KB of code out of
// raster_kernel.
__constant__ double kUnitOfU8[256];
__device__ __forceinline__ double unit_of_u8(unsigned char c) { return kUnitOfU8[c]; }

// ------------------------------------------------------------------------------------------------------
// Rust cast semantics
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int f64_as_i32(double v) {
    // `as i32`: round toward zero, saturate, NaN -> 0.  cvt.rzi.s32.f64 has exactly these semantics.
    return __double2int_rz(v);
}
__device__ __forceinline__ unsigned f64_as_u8(double v) {
    // `as u8`
    if (!(v > 0.0)) return 0u;  // negatives, -0, NaN
    if (v >= 255.0) return 255u;
    return (unsigned)__double2int_rz(v);
}
// sqrt / division with the operand classes that send CUDA's inline f64 sequences to their (long, out-of-line)
// special-case subroutines peeled off first: a zero or negative radicand, a zero dividend.  Those are everyday
// inputs here (pixels exactly on a centre line, the first pixel of a segment, points beyond a round cap), and
// results are identical: sqrt(+-0) = +-0, sqrt(x<0) = NaN, +-0 / positive = +-0.
// (The special operand is replaced by 1.0 *before* the operation and the result selected afterwards: a plain
// `if (x == 0) return x;` is if-converted by the compiler, which then still sends those lanes down the slow path.)
// OSMR_OPAQUE_F64 hides the substituted operand from the optimiser, which otherwise proves that the substitution never
// changes the result, folds it away again and sends the special operands back down the slow path (seen in the SASS of
// line_cover_kernel: MUFU.RSQ64H on the raw operand + CALL ..dsqrt_rn_f64_mediumpath for every zero).
#ifdef OSMR_EMULATED
#define OSMR_OPAQUE_F64(v) ((void)0)
#else
#define OSMR_OPAQUE_F64(v) asm volatile("" : "+d"(v))
#endif
__device__ __forceinline__ double sqrt_peeled(double x) {
    const bool special = !(x > 0.0);  // zero, negative or NaN
    double arg = special ? 1.0 : x;
    OSMR_OPAQUE_F64(arg);
    const double r = sqrt(arg);
    if (!special) return r;
    return (x == 0.0) ? x : __longlong_as_double(0x7ff8000000000000LL);
}
__device__ __forceinline__ double div_pos_peeled(double a, double b) {  // b > 0 and finite
    const bool zero = (a == 0.0);
    double num = zero ? 1.0 : a;
    OSMR_OPAQUE_F64(num);
    const double q = num / b;
    return zero ? a : q;
}

__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// ------------------------------------------------------------------------------------------------------
// a1: projection.  reference src/tile.rs:88-106 + src/draw/point.rs:11-19.
//
// coords_to_xy(z) = factor * (256 * 2^z) where factor = x/(2*pi) depends only on the node.  The second
// factor is a power of two, so the product is exact and `merc` (the per-node factor pair, computed once per
// dataset by project_nodes_kernel) reproduces coords_to_xy for every zoom with one multiplication.
// ------------------------------------------------------------------------------------------------------
struct TileXform {
    double dim;    // 256 * 2^zoom
    double tx256;  // f64::from(tile.x * TILE_SIZE) (u32 multiply, wrapping)
    double ty256;
    double scale;
};

__device__ __forceinline__ TileXform make_xform(const osmr_tile& t) {
    TileXform x;
    x.dim = (double)(unsigned)(256u * (1u << (t.zoom & 31u)));
    x.tx256 = (double)(unsigned)(t.x * 256u);
    x.ty256 = (double)(unsigned)(t.y * 256u);
    x.scale = (double)t.scale;
    return x;
}

__device__ __forceinline__ int2 project_point(const double2 m, const TileXform& t) {
    double x = m.x * t.dim - t.tx256;  // -fmad=false keeps mul and sub separate
    double y = m.y * t.dim - t.ty256;
    int2 p;
    p.x = f64_as_i32(round(x * t.scale));  // f64::round: half away from zero == CUDA round()
    p.y = f64_as_i32(round(y * t.scale));
    return p;
}

// Point::dist (point.rs:21-25)
__device__ __forceinline__ double point_dist(int ax, int ay, int bx, int by) {
    double dx = (double)wsub(ax, bx);
    double dy = (double)wsub(ay, by);
    return sqrt_peeled(dx * dx + dy * dy);
}

// Point::push_away_from (point.rs:27-35)
__device__ __forceinline__ int2 push_away_from(int sx, int sy, int ox, int oy, double by) {
    double dist = point_dist(sx, sy, ox, oy);
    double push = by / dist;
    int2 r;
    r.x = wadd(sx, f64_as_i32(round((double)wsub(sx, ox) * push)));
    r.y = wadd(sy, f64_as_i32(round((double)wsub(sy, oy) * push)));
    return r;
}

// ------------------------------------------------------------------------------------------------------
// a2: fill edge -> row span in O(1).  Closed form of the all-octant Bresenham walk of fill.rs:51-104
// (pinned against the step-by-step oracle by tests/test_closed_forms.py).
// Returns false when row y is not visited by the edge.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ long long ceil_div(long long a, long long b) {  // b > 0
    long long q = a / b;
    return (a % b != 0 && a > 0) ? q + 1 : q;
}

__device__ __forceinline__ long long fill_hi(long long jj, long long a, long long b) {
    if (jj == b) return a;
    long long Y = ceil_div(a - 2 * b + 2 * jj * a, 2 * b);
    Y = Y < 0 ? 0 : Y;
    return Y < a ? Y : a;
}

template <typename I>
__device__ __forceinline__ I fill_hi_t(I jj, I a, I b) {
    if (jj == b) return a;
    I num = a - 2 * b + 2 * jj * a, den = 2 * b;
    I q = num / den;
    if (num % den != 0 && num > 0) q += 1;  // ceil
    q = q < 0 ? 0 : q;
    return q < a ? q : a;
}

template <typename I>
__device__ __forceinline__ void fill_span_t(I a, I b, I j, I& lo, I& e) {
    if (j == 0) {
        lo = 0;
    } else {
        I h = fill_hi_t<I>(j - 1, a, b);
        I num = 2 * a - b + 2 * (j - 1) * a, den = 2 * b;
        I X = num / den;
        if (num % den != 0 && num < 0) X -= 1;  // floor
        lo = h + ((h <= X) ? 1 : 0);
        lo = lo < a ? lo : a;
    }
    I h = fill_hi_t<I>(j, a, b);
    e = lo > h ? lo : h;
}

__device__ __forceinline__ bool fill_edge_row_span(int x1, int y1, int x2, int y2, int y, int& xmin, int& xmax,
                                                   bool& poisoned) {
    long long a = llabs((long long)x2 - (long long)x1);
    long long b = llabs((long long)y2 - (long long)y1);
    int sx = (x1 < x2) ? 1 : -1;
    int sy = (y1 < y2) ? 1 : -1;
    long long j = ((long long)y - (long long)y1) * sy;
    if (j < 0 || j > b) return false;
    if (b == 0) {
        xmin = min(x1, x2);
        xmax = max(x1, x2);
        poisoned = true;
        return true;
    }
    long long lo, e;
    if (a < 16384 && b < 16384) {  // every product below 2^30: 32-bit divisions (4x cheaper than 64-bit ones)
        int lo32, e32;
        fill_span_t<int>((int)a, (int)b, (int)j, lo32, e32);
        lo = lo32;
        e = e32;
    } else {
        fill_span_t<long long>(a, b, j, lo, e);
    }
    int xa = x1 + sx * (int)lo;
    int xb = x1 + sx * (int)e;
    xmin = min(xa, xb);
    xmax = max(xa, xb);
    poisoned = (j == 0 && y1 <= y2) || (j == b && y2 <= y1);
    return true;
}

// ------------------------------------------------------------------------------------------------------
// a4: thick-line traversal with random access (closed form of line.rs:65-158).
// Number of Bresenham corrections after n calls of `update_error` starting from error e0.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ncorr(long long e0, long long n, long long mn_d, long long mx_d) {
    long long num = e0 + 2 * mn_d * n - mx_d;
    if (num <= 0) return 0;
    long long den = 2 * mx_d;
    if (num < 0x7fffffffLL && den < 0x7fffffffLL) return (long long)(((unsigned)num + (unsigned)den - 1u) / (unsigned)den);
    return (num + den - 1) / den;
}

// ------------------------------------------------------------------------------------------------------
// a5: opacity calculator (opacity_calculator.rs).  Built once per line op in shared memory.
// ------------------------------------------------------------------------------------------------------
constexpr int kMaxDashSegs = 17;  // dash lists of up to 32 numbers -> <= 17 "on" segments

struct DashSeg {  // 64 bytes
    double start_from, start_to, end_from, end_to, opacity_mul;
    double orig_a, orig_b;  // valid when OpacityCalc::round_caps
    double unit_ramps;      // 1.0 when both ramp lengths (start_to-start_from, end_to-end_from) are exactly 1.0
};

struct alignas(16) OpacityCalc {  // 64-byte header + 64 bytes per dash segment (same layout in HBM and smem)
    double half_line_width;
    double total_dash_len;
    // per-op constants of get_opacity_by_center_distance for cap_dist == 0 (the only case unless round_caps)
    double feather_from, feather_to, feather_dist, opacity_mul;
    int n_segs;       // 0 == "dashes: None"
    int round_caps;   // original_endpoints is Some(..)  (LineCap::Round)
    double inv_total;  // 1 / total_dash_len: quotient estimate of the exact fmod below
    DashSeg segs[kMaxDashSegs];
};
static_assert(sizeof(DashSeg) == 64 && sizeof(OpacityCalc) == 64 + 64 * kMaxDashSegs, "calculator wire layout");

__device__ __forceinline__ bool is_non_trivial_cap(unsigned cap) { return cap == OSMR_CAP_SQUARE || cap == OSMR_CAP_ROUND; }

__device__ __forceinline__ void center_feather(double hw, double& from, double& to, double& dist, double& mul) {
    from = fmax(hw - 0.5, 0.0);  // f64::max ignores NaN, like fmax
    to = fmax(hw + 0.5, 1.0);
    dist = to - from;
    mul = fmin(2.0 * hw, 1.0);
}

// OpacityCalculator::new + compute_segments (opacity_calculator.rs:16-30, 98-143)
__device__ inline void build_calc(OpacityCalc& c, double hw, const double* dashes, int n, double dash_scale, bool has_dashes,
                                  unsigned cap, int max_segs) {
    c.half_line_width = hw;
    c.n_segs = 0;
    c.round_caps = (cap == OSMR_CAP_ROUND) ? 1 : 0;
    double len_before = 0.0;
    if (has_dashes && n > 0) {
        for (int k = 0; k < n + 1; ++k) {
            int idx = (k < n) ? k : 0;
            double dash = dashes[idx] * dash_scale;  // drawer.rs:163-164 scale_dashes
            double start = len_before;
            if (idx != 0 || k == 0) len_before += dash;  // `segments.is_empty()` <=> first iteration
            if (idx % 2 != 0) continue;
            double end = start + dash;
            double oa = start, ob = end;
            if (is_non_trivial_cap(cap)) {
                start -= hw;
                end += hw;
            }
            double mid = (start + end) / 2.0;
            if (c.n_segs < max_segs) {
                DashSeg& s = c.segs[c.n_segs++];
                s.start_from = fmin(start - 0.5, mid - 1.0);
                s.start_to = fmin(start + 0.5, mid);
                s.end_from = fmax(end - 0.5, mid);
                s.end_to = fmax(end + 0.5, mid + 1.0);
                s.opacity_mul = fmin(end - start, 1.0);
                s.orig_a = oa;
                s.orig_b = ob;
                s.unit_ramps = (s.start_to - s.start_from == 1.0 && s.end_to - s.end_from == 1.0) ? 1.0 : 0.0;
            }
        }
    }
    c.total_dash_len = len_before;
    c.inv_total = (len_before > 0.0) ? 1.0 / len_before : 0.0;
    double hw0 = sqrt(hw * hw - 0.0 * 0.0);  // calculate() with cap_dist == 0
    center_feather(hw0, c.feather_from, c.feather_to, c.feather_dist, c.opacity_mul);
}

// f64 `%` for x >= 0, y > 0.  fmod's result x - n*y (n = floor(x/y)) is always representable, so one FMA with the
// right integer n yields it exactly; a quotient that rounded across an integer is repaired by the sign test.
// Much smaller than libdevice's general fmod, which matters for the instruction cache of raster_kernel.
__device__ __noinline__ double fmod_general(double x, double y) {
    OSMR_COUNT("fmod_general", 1);
    return fmod(x, y);
}
__device__ __noinline__ double div_general(double a, double b) { return a / b; }
__device__ __forceinline__ double fmod_exact(double x, double y, double inv_y) {
    OSMR_COUNT("fmod_exact", 1);
    if (x >= 0.0 && x < y) return x;
    double qf = x * inv_y;  // estimate of x / y, off by far less than one for quotients below 2^50
    if (!(x >= 0.0) || !(qf < 1.0e15)) return fmod_general(x, y);
    double q = floor(qf);
    double r = __fma_rn(-q, y, x);  // exact: x - q*y is representable for q within one of floor(x/y)
    if (r < 0.0) {
        q -= 1.0;
        r = __fma_rn(-q, y, x);
    } else if (r >= y) {
        q += 1.0;
        r = __fma_rn(-q, y, x);
    }
    if (!(r >= 0.0 && r < y)) return fmod_general(x, y);  // cannot happen for sane dash lengths; stay exact anyway
    return r;
}

// OpacityCalculator::calculate (opacity_calculator.rs:32-43) with traveled distance supplied per segment.
__device__ __forceinline__ void calc_opacity(const OpacityCalc& c, double traveled, double center_distance, double start_distance,
                                             double& opacity, bool& is_in_line) {
    double sd_opacity = 1.0;
    double ff = c.feather_from, ft = c.feather_to, fd = c.feather_dist, fm = c.opacity_mul;
    if (c.n_segs != 0) {
        double dist_rem = traveled + start_distance;
        if (c.total_dash_len > 0.0) dist_rem = fmod_exact(dist_rem, c.total_dash_len, c.inv_total);  // `%=` (opacity_calculator.rs:59)
        double acc = 0.0;
        bool has_cap = false;
        double cap = 0.0;
        for (int i = 0; i < c.n_segs; ++i) {
            const DashSeg& s = c.segs[i];
            if (dist_rem < s.start_from || dist_rem > s.end_to) continue;  // NaN falls through to the last branch, as in the reference
            double base;
            if (dist_rem <= s.start_to)
                base = (s.unit_ramps != 0.0) ? (dist_rem - s.start_from) : div_general(dist_rem - s.start_from, s.start_to - s.start_from);  // x / 1.0 == x
            else if (dist_rem < s.end_from)
                base = 1.0;
            else
                base = (s.unit_ramps != 0.0) ? (s.end_to - dist_rem) : div_general(s.end_to - dist_rem, s.end_to - s.end_from);
            acc = fmax(acc, s.opacity_mul * base);
            if (c.round_caps) {
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
        sd_opacity = acc;
        if (has_cap && cap != 0.0) {
            double hw = sqrt_peeled(c.half_line_width * c.half_line_width - cap * cap);  // may be NaN on purpose
            center_feather(hw, ff, ft, fd, fm);
        }
    }
    double v;
    if (center_distance < ff)
        v = 1.0;
    else if (center_distance < ft)
        v = (fd == 1.0) ? (ft - center_distance) : div_general(ft - center_distance, fd);
    else
        v = 0.0;
    double cd = fm * v;
    opacity = fmin(sd_opacity, cd);
    is_in_line = cd > 0.0;
}

}  // namespace osmr

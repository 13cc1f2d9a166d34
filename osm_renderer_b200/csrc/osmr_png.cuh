// osmr_png.cuh -- SURVEY.md 8(f) row f4: PNG encode of the RGB tiles on the device.
//
// Replaces rgb_triples_to_png (reference src/draw/png_writer.rs:4-21; png crate, 8-bit RGB, one IDAT) for
// Drawer::draw_tile (src/draw/drawer.rs:40-58).  A PNG encoder is free to choose filters and the deflate stream, so the
// contract is the decoded image (lossless), not the byte stream of the `png` crate:
//   * per row the filter (None / Sub / Up / Paeth) with the smallest sum of absolute residuals (libpng's heuristic);
//   * deflate with the FIXED Huffman code and distance-1 matches (run-length): filtered map tiles are mostly runs of zeros;
//   * a tile is cut into 8 bands of rows, one warp per band: each band is one deflate block followed by an empty stored
//     block (the classic sync flush), so the bands are byte aligned and can be produced independently; matches never
//     cross a row, so no band refers back to another one;
//   * Adler-32 of the filtered bytes and CRC-32 of the IDAT chunk are computed in parallel (per-band sums; per-thread CRC
//     slices combined with x^(8n) mod P multiplications, zlib's crc32_combine arithmetic).
//
//   png_encode_kernel   CTA per tile, warp per band   RGB rows -> filter -> tokens -> bit stream in the band's scratch
//   png_size_kernel     thread per tile                size of the tile's PNG file from its band sizes
//   (auto_scan_kernel)                                 offsets of the files in the packed output
//   png_finish_kernel   CTA per tile                   signature, IHDR, IDAT (zlib header, bands, Adler-32), CRCs, IEND
#pragma once
#include "osmr_kernels.cuh"

namespace osmr {

constexpr int kPngThreads = 256;
constexpr int kPngBands = kPngThreads / 32;
constexpr unsigned kPngFixed = 8 + 25 + 12 + 2 + 4 + 12;  // signature, IHDR, IDAT framing, zlib header, Adler-32, IEND
constexpr unsigned kAdlerMod = 65521u;
constexpr unsigned kCrcPoly = 0xedb88320u;

struct PngScene {
    const unsigned char* rgb;  // n_tiles images of D*D*3 bytes
    int D;
    unsigned n_tiles;
    unsigned* band_words;      // scratch: band b of tile t at band_words + (t * kPngBands + b) * band_cap_words
    unsigned band_cap_words;
    unsigned* band_bytes;      // [n_tiles * kPngBands]
    uint2* band_adler;         // [n_tiles * kPngBands]: (sum of bytes, sum of (n - i) * byte) mod 65521 over the band, and
    unsigned* band_len;        // the number of filtered bytes of the band
    unsigned* tile_off;        // [n_tiles + 1]: PNG sizes, then their exclusive scan
    unsigned char* out;        // packed files
};

__device__ __forceinline__ int png_paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// filtered byte i (0 = the filter type byte) of a row; `cur` / `up` point at the raw bytes of the row and of the one above
// (up == nullptr for the first row: zeros)
__device__ __forceinline__ unsigned png_filtered(const unsigned char* __restrict__ cur, const unsigned char* __restrict__ up, int filter, int i) {
    if (i == 0) return (unsigned)filter;
    const int j = i - 1;
    const int x = cur[j];
    const int a = j >= 3 ? cur[j - 3] : 0;
    const int b = up ? up[j] : 0;
    const int c = (up && j >= 3) ? up[j - 3] : 0;
    int pred = 0;
    if (filter == 1)
        pred = a;
    else if (filter == 2)
        pred = b;
    else if (filter == 4)
        pred = png_paeth(a, b, c);
    return (unsigned)(x - pred) & 0xffu;
}

// one deflate token in stream bit order (first bit = bit 0): fixed Huffman codes go out MSB first, extra bits LSB first
__device__ __forceinline__ void png_literal(unsigned v, unsigned& bits, unsigned& nb) {
    if (v < 144u) {
        bits = __brev(0x30u + v) >> 24;
        nb = 8;
    } else {
        bits = __brev(0x190u + (v - 144u)) >> 23;
        nb = 9;
    }
}
__device__ __forceinline__ void png_match_dist1(unsigned len, unsigned& bits, unsigned& nb) {  // 3 <= len <= 258, distance 1
    unsigned code, eb = 0, extra = 0;
    if (len <= 10u) {
        code = 254u + len;
    } else if (len == 258u) {
        code = 285u;
    } else {
        const unsigned l = len - 3u;
        const unsigned hb = 31u - (unsigned)__clz(l);
        eb = hb - 2u;
        code = 265u + 4u * (eb - 1u) + ((l >> eb) & 3u);
        extra = l & ((1u << eb) - 1u);
    }
    if (code < 280u) {
        bits = __brev(code - 256u) >> 25;  // 7 bits: 0000000 .. 0010111
        nb = 7;
    } else {
        bits = __brev(0xc0u + (code - 280u)) >> 24;  // 8 bits: 11000000 ..
        nb = 8;
    }
    bits |= extra << nb;
    nb += eb;
    nb += 5;  // distance code 0 (distance 1): five zero bits, no extra bits
}

#ifndef OSMR_PNG_ROW_CACHE
#define OSMR_PNG_ROW_CACHE 3080
#endif
constexpr int kPngRowCache = OSMR_PNG_ROW_CACHE;  // filtered bytes of a row kept in shared memory (scale <= 4; larger rows are recomputed)

struct PngWarpSmem {
    unsigned char frow[kPngRowCache];
    unsigned stage[64];   // bit staging: stage[0] holds the stream's partial word
    unsigned emask[200];  // e[i] = filtered[i] == filtered[i-1], one bit per position of the row (<= 6145 positions)
    int lastz[200];       // last position with e == 0 at or before the end of word g
    int firstz[201];      // first position with e == 0 at or after the start of word g
};

__global__ void __launch_bounds__(kPngThreads) png_encode_kernel(PngScene ps) {
    __shared__ PngWarpSmem smem[kPngBands];
    const unsigned lane = lane_id(), band = threadIdx.x >> 5, tile = blockIdx.x;
    PngWarpSmem& sm = smem[band];
    const int D = ps.D, W = 3 * D, NP = W + 1;  // NP filtered bytes per row
    const int n_words = (NP + 31) / 32;
    const int rows = D / kPngBands;
    const unsigned char* img = ps.rgb + (size_t)tile * D * D * 3;
    unsigned* out_words = ps.band_words + ((size_t)tile * kPngBands + band) * ps.band_cap_words;
    unsigned long long cursor = 0;  // bits produced by this band
    unsigned flushed = 0;           // words already written to out_words
    for (unsigned i = lane; i < 64; i += 32) sm.stage[i] = 0;
    __syncwarp();
    // appends (bits, nb) of every lane in lane order; flushes the complete words of the staging buffer
    auto emit = [&](unsigned bits, unsigned nb) {
        unsigned incl = nb;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += v;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (nb) {
            const unsigned long long pos = cursor + (incl - nb) - 32ull * flushed;  // bit position inside the staging buffer
            const unsigned w = (unsigned)(pos >> 5), sh = (unsigned)(pos & 31);
            atomicOr(&sm.stage[w], bits << sh);
            if (sh + nb > 32u) atomicOr(&sm.stage[w + 1], bits >> (32u - sh));
        }
        cursor += total;
        __syncwarp();
        const unsigned full = (unsigned)(cursor >> 5) - flushed;  // complete words in the staging buffer (<= 19)
        if (full) {
            const unsigned carry = sm.stage[full];  // every lane reads the partial word before it is moved
            const unsigned mine = lane < full ? sm.stage[lane] : 0u;
            __syncwarp();
            if (lane < full) out_words[flushed + lane] = mine;
            if (lane <= full) sm.stage[lane] = (lane == 0) ? carry : 0u;
            flushed += full;
            __syncwarp();
        }
    };
    const bool last_band = band == kPngBands - 1;
    emit(lane == 0 ? ((last_band ? 1u : 0u) | (1u << 1)) : 0u, lane == 0 ? 3u : 0u);  // BFINAL, BTYPE = 01 (fixed Huffman)
    unsigned long long s1 = 0, s2 = 0;  // Adler-32 sums of this band relative to its start
    unsigned long long n_done = 0;
    for (int r = 0; r < rows; ++r) {
        const int y = (int)band * rows + r;
        const unsigned char* cur = img + (size_t)y * W;
        const unsigned char* up = y > 0 ? cur - W : nullptr;
        // ---- filter choice: smallest sum of |residual as signed byte| (ties: the lowest filter number) ----
        unsigned sum[4] = {0, 0, 0, 0};
        // four bytes per lane and iteration: the row and the one above as 32-bit words (rows are 4-byte aligned: 3 * D bytes)
        const unsigned* cw = reinterpret_cast<const unsigned*>(cur);
        const unsigned* uw = reinterpret_cast<const unsigned*>(up);
        for (int wd = (int)lane; wd < W / 4; wd += 32) {
            const unsigned x4 = cw[wd], xp4 = wd ? cw[wd - 1] : 0u;
            const unsigned u4 = up ? uw[wd] : 0u, up4 = (up && wd) ? uw[wd - 1] : 0u;
            const unsigned a4 = (xp4 >> 8) | (x4 << 24);  // the bytes three positions to the left (zeros before the row)
            const unsigned c4 = (up4 >> 8) | (u4 << 24);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int x = (int)((x4 >> (8 * q)) & 0xffu), a = (int)((a4 >> (8 * q)) & 0xffu);
                const int b = (int)((u4 >> (8 * q)) & 0xffu), c = (int)((c4 >> (8 * q)) & 0xffu);
                const int res[4] = {x, x - a, x - b, x - png_paeth(a, b, c)};
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const int v = res[f] & 0xff;
                    sum[f] += (unsigned)(v < 128 ? v : 256 - v);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < 4; ++f)
            for (int o = 16; o > 0; o >>= 1) sum[f] += __shfl_xor_sync(0xffffffffu, sum[f], o);
        int filter = 0;
        unsigned best = sum[0];
        if (sum[1] < best) {
            best = sum[1];
            filter = 1;
        }
        if (sum[2] < best) {
            best = sum[2];
            filter = 2;
        }
        if (sum[3] < best) {
            best = sum[3];
            filter = 4;
        }
        // ---- equality mask of the filtered row + Adler-32 sums ----
        unsigned long long r1 = 0, r2 = 0;
        const bool cached = NP <= kPngRowCache;
        if (cached) {
            if (lane == 0) {
                sm.frow[0] = (unsigned char)filter;
                r1 += (unsigned)filter;
                r2 += (unsigned long long)NP * (unsigned)filter;
            }
            for (int wd = (int)lane; wd < W / 4; wd += 32) {
                const unsigned x4 = cw[wd], xp4 = wd ? cw[wd - 1] : 0u;
                const unsigned u4 = up ? uw[wd] : 0u, up4 = (up && wd) ? uw[wd - 1] : 0u;
                const unsigned a4 = (xp4 >> 8) | (x4 << 24);
                const unsigned c4 = (up4 >> 8) | (u4 << 24);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int x = (int)((x4 >> (8 * q)) & 0xffu), a = (int)((a4 >> (8 * q)) & 0xffu);
                    const int b = (int)((u4 >> (8 * q)) & 0xffu), c = (int)((c4 >> (8 * q)) & 0xffu);
                    const int pred = filter == 1 ? a : (filter == 2 ? b : (filter == 4 ? png_paeth(a, b, c) : 0));
                    const unsigned f = (unsigned)(x - pred) & 0xffu;
                    const int i = 4 * wd + q + 1;
                    sm.frow[i] = (unsigned char)f;
                    r1 += f;
                    r2 += (unsigned long long)(NP - i) * f;
                }
            }
            __syncwarp();
        }
        for (int g = 0; g < n_words; ++g) {
            const int i = 32 * g + (int)lane;
            bool e = false;
            if (i < NP) {
                if (cached) {
                    e = i >= 1 && sm.frow[i] == sm.frow[i - 1];
                } else {
                    const unsigned f = png_filtered(cur, up, filter, i);
                    r1 += f;
                    r2 += (unsigned long long)(NP - i) * f;
                    e = i >= 1 && f == png_filtered(cur, up, filter, i - 1);
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, e);
            if (lane == 0) sm.emask[g] = m;
        }
        for (int o = 16; o > 0; o >>= 1) {
            r1 += __shfl_xor_sync(0xffffffffu, r1, o);
            r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
        // band-relative Adler: A += r1, B += NP * A_before + r2
        s2 = (s2 + (unsigned long long)NP * s1 + r2) % kAdlerMod;
        s1 = (s1 + r1) % kAdlerMod;
        n_done += (unsigned long long)NP;
        __syncwarp();
        // ---- last / first zero of the mask per word (positions beyond the row count as zeros) ----
        // (one mask word per lane: inclusive max-scan from the left / min-scan from the right, 32 words at a time)
        {
            int carry = 0;  // e[0] == 0 always
            for (int g0 = 0; g0 < n_words; g0 += 32) {
                const int g = g0 + (int)lane;
                const unsigned z = g < n_words ? ~sm.emask[g] : 0u;
                int v = z ? 32 * g + 31 - __clz(z) : -1;
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, v, o);
                    if ((int)lane >= o) v = max(v, u);
                }
                v = max(v, carry);
                if (g < n_words) sm.lastz[g] = v;  // only lastz[g - 1] of a word g inside the row is ever read: a valid position
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
            carry = NP;
            if (lane == 0) sm.firstz[n_words] = NP;
            for (int g0 = ((n_words - 1) / 32) * 32; g0 >= 0; g0 -= 32) {
                const int g = g0 + (int)lane;
                const unsigned z = g < n_words ? ~sm.emask[g] : 0u;
                int v = z ? min(32 * g + __ffs(z) - 1, NP) : NP;
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_down_sync(0xffffffffu, v, o);
                    if ((int)lane + o < 32) v = min(v, u);
                }
                v = min(v, carry);
                if (g < n_words) sm.firstz[g] = v;
                carry = __shfl_sync(0xffffffffu, v, 0);
            }
        }
        __syncwarp();
        // ---- tokens, 32 positions at a time in stream order ----
        for (int g = 0; g < n_words; ++g) {
            const int i = 32 * g + (int)lane;
            unsigned bits = 0, nb = 0;
            if (i < NP) {
                const unsigned m = sm.emask[g];
                bool literal = true;
                if ((m >> lane) & 1u) {
                    // run of equal bytes [s, end): s - 1 is the last position before i whose byte differs from its predecessor
                    const unsigned below = ~m & ((1u << lane) - 1u);
                    const int z = below ? 32 * g + 31 - __clz(below) : sm.lastz[g - 1];
                    const unsigned above = lane == 31 ? 0u : (~m & ~((2u << lane) - 1u));
                    const int end = above ? min(32 * g + __ffs(above) - 1, NP) : sm.firstz[g + 1];
                    const int s = z + 1, L = end - s, off = i - s;
                    const int chunk = off / 258, Lc = min(258, L - 258 * chunk);
                    if (Lc >= 3) {
                        literal = false;
                        if (off - 258 * chunk == 0) png_match_dist1((unsigned)Lc, bits, nb);
                    }
                }
                if (literal) png_literal(cached ? (unsigned)sm.frow[i] : png_filtered(cur, up, filter, i), bits, nb);
            }
            if (!__any_sync(0xffffffffu, nb != 0)) continue;  // 32 positions inside one long run: nothing to emit
            emit(bits, nb);
        }
    }
    // end of block (code 256: seven zero bits); every band but the last: empty stored block = byte alignment (sync flush)
    emit(0u, lane == 0 ? 7u : 0u);
    if (!last_band) {
        emit(0u, lane == 0 ? 3u : 0u);  // BFINAL = 0, BTYPE = 00
        const unsigned pad = (unsigned)((8 - (cursor & 7)) & 7);
        emit(0u, lane == 0 ? pad : 0u);
        emit(lane == 0 ? 0xffff0000u : 0u, lane == 0 ? 32u : 0u);  // LEN = 0, NLEN = 0xffff
    } else {
        const unsigned pad = (unsigned)((8 - (cursor & 7)) & 7);
        emit(0u, lane == 0 ? pad : 0u);
    }
    // the partial last word
    if (lane == 0) {
        if (cursor & 31) out_words[flushed] = sm.stage[0];
        const size_t bi = (size_t)tile * kPngBands + band;
        ps.band_bytes[bi] = (unsigned)(cursor >> 3);
        ps.band_adler[bi] = make_uint2((unsigned)s1, (unsigned)s2);
        ps.band_len[bi] = (unsigned)n_done;
    }
}

__global__ void png_size_kernel(PngScene ps) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ps.n_tiles) return;
    unsigned z = 0;
    for (int b = 0; b < kPngBands; ++b) z += ps.band_bytes[(size_t)t * kPngBands + b];
    ps.tile_off[t] = kPngFixed + z;
}

// zlib's crc32_combine arithmetic: polynomials over GF(2) modulo P in the reflected representation
__device__ __forceinline__ unsigned crc_multmodp(unsigned a, unsigned b) {
    unsigned m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}

__global__ void __launch_bounds__(kPngThreads) png_finish_kernel(PngScene ps) {
    __shared__ unsigned table[256];
    __shared__ unsigned x2n[32];     // x^(2^i) mod P
    __shared__ unsigned part[kPngThreads];
    __shared__ unsigned band_off[kPngBands + 1];
    __shared__ unsigned mul_level;
    const unsigned tile = blockIdx.x, tid = threadIdx.x;
    const int D = ps.D;
    unsigned char* out = ps.out + ps.tile_off[tile];
    {  // CRC byte table
        unsigned c = tid;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
        table[tid] = c;
    }
    if (tid == 0) {
        x2n[0] = 1u << 30;  // x^1
        for (int i = 1; i < 32; ++i) x2n[i] = crc_multmodp(x2n[i - 1], x2n[i - 1]);
        unsigned o = 0;
        for (int b = 0; b < kPngBands; ++b) {
            band_off[b] = o;
            o += ps.band_bytes[(size_t)tile * kPngBands + b];
        }
        band_off[kPngBands] = o;
    }
    __syncthreads();
    auto x2nmodp = [&](unsigned long long n, unsigned k) {  // x^(n * 2^k) mod P
        unsigned p = 1u << 31;
        while (n) {
            if (n & 1ull) p = crc_multmodp(x2n[k & 31u], p);
            n >>= 1;
            ++k;
        }
        return p;
    };
    auto crc_std = [&](const unsigned char* p, unsigned n) {  // the standard CRC-32 of a short byte string
        unsigned c = 0xffffffffu;
        for (unsigned i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xffu] ^ (c >> 8);
        return c ^ 0xffffffffu;
    };
    auto put_be32 = [&](unsigned char* p, unsigned v) {
        p[0] = (unsigned char)(v >> 24);
        p[1] = (unsigned char)(v >> 16);
        p[2] = (unsigned char)(v >> 8);
        p[3] = (unsigned char)v;
    };
    const unsigned zbytes = band_off[kPngBands];
    const unsigned zlen = 2u + zbytes + 4u;  // zlib stream
    unsigned char* idat = out + 8 + 25;      // length, type, data, crc
    if (tid == 0) {
        const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        for (int i = 0; i < 8; ++i) out[i] = sig[i];
        unsigned char* ih = out + 8;
        put_be32(ih, 13u);
        ih[4] = 'I';
        ih[5] = 'H';
        ih[6] = 'D';
        ih[7] = 'R';
        put_be32(ih + 8, (unsigned)D);
        put_be32(ih + 12, (unsigned)D);
        ih[16] = 8;  // bit depth
        ih[17] = 2;  // colour type RGB (png_writer.rs:9)
        ih[18] = 0;
        ih[19] = 0;
        ih[20] = 0;
        put_be32(ih + 21, crc_std(ih + 4, 17));
        put_be32(idat, zlen);
        idat[4] = 'I';
        idat[5] = 'D';
        idat[6] = 'A';
        idat[7] = 'T';
        idat[8] = 0x78;  // deflate, 32 K window
        idat[9] = 0x01;  // no preset dictionary, fastest; (0x78 << 8 | 0x01) % 31 == 0
        // Adler-32 of all filtered bytes: the bands' relative sums applied in order
        unsigned long long A = 1, B = 0;
        for (int b = 0; b < kPngBands; ++b) {
            const size_t bi = (size_t)tile * kPngBands + b;
            const uint2 s = ps.band_adler[bi];
            const unsigned long long n = ps.band_len[bi];
            B = (B + (n % kAdlerMod) * A + s.y) % kAdlerMod;
            A = (A + s.x) % kAdlerMod;
        }
        put_be32(idat + 10 + zbytes, (unsigned)((B << 16) | A));
        unsigned char* iend = idat + 8 + zlen + 4;
        const unsigned char tail[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
        for (int i = 0; i < 12; ++i) iend[i] = tail[i];
    }
    // the bands' bytes
    for (int b = 0; b < kPngBands; ++b) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(ps.band_words + ((size_t)tile * kPngBands + b) * ps.band_cap_words);
        unsigned char* dst = idat + 10 + band_off[b];
        const unsigned n = band_off[b + 1] - band_off[b];
        for (unsigned i = tid; i < n; i += kPngThreads) dst[i] = src[i];
    }
    __syncthreads();
    // CRC-32 over type + data: every thread a slice of Lc bytes of the stream right-aligned in kPngThreads * Lc bytes (a raw
    // CRC -- zero initial value, no final inversion -- ignores leading zeros), then a combine tree with x^(8 Lc 2^j)
    const unsigned N = 4u + zlen;
    const unsigned char* reg = idat + 4;
    const unsigned Lc = (N + kPngThreads - 1) / kPngThreads;
    const unsigned padn = Lc * kPngThreads - N;
    {
        unsigned c = 0;
        const unsigned v0 = tid * Lc;
        for (unsigned k = 0; k < Lc; ++k) {
            const unsigned v = v0 + k;
            if (v >= padn) c = table[(c ^ reg[v - padn]) & 0xffu] ^ (c >> 8);
        }
        part[tid] = c;
    }
    if (tid == 0) mul_level = x2nmodp(Lc, 3);  // x^(8 Lc)
    __syncthreads();
    for (unsigned stride = 1; stride < kPngThreads; stride <<= 1) {
        const unsigned m = mul_level;
        unsigned v = 0;
        const bool active = (tid % (2 * stride)) == 0;
        if (active) v = crc_multmodp(m, part[tid]) ^ part[tid + stride];
        __syncthreads();
        if (active) part[tid] = v;
        if (tid == 0) mul_level = crc_multmodp(m, m);
        __syncthreads();
    }
    if (tid == 0) {
        // the standard CRC = raw CRC ^ (the all-ones initial value pushed through N bytes) ^ all ones
        const unsigned crc = part[0] ^ crc_multmodp(x2nmodp(N, 3), 0xffffffffu) ^ 0xffffffffu;
        put_be32(idat + 8 + zlen, crc);
    }
}

}  // namespace osmr

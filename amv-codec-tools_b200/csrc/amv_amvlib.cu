// amv_amvlib.cu -- back half of the amvlib-flavoured decoder (sm_100a): the arithmetic of
// C-AMVDecoder/amvlib/AmvJpeg.c behind AmvVideoDecode (AMVDec.c:259-286).
//
//   k_idct_bgr : one WARP per group of 5 macroblocks (30 blocks on 30 lanes, like the encoder).
//     A  every lane scatters its block's tokens (raster position, coefficient * quantiser as a
//        full int -- IQtIZzBlock :1012-1048) into a conflict-free shared column, runs Fast_IDCT
//        (idctrow :1078-1125, idctcol :1127-1175; Chen-Wang on 32-bit ints, results clamped to
//        -256..255 by iclp :1069-1076) in registers and leaves the 64 results, +128 for luma, in a
//        [sample][lane] int16 exchange tile;
//     B  the warp turns the tile into pixels: each lane takes four horizontally adjacent pixels at
//        a time, reads 4 Y + 2 Cb + 2 Cr, applies StoreBuffer's fixed-point conversion (:808-810),
//        and writes 12 bytes, so one pixel row of the group is one contiguous 240-byte run of the
//        bottom-up BGR24 bitmap (row h-1-y, :801).
//
// The row/column "all AC zero" shortcuts of the reference are identities of the general path
// (row: ((b0<<11)+128)>>8 == b0<<3; column: ((b0<<8)+8192)>>14 == (b0+32)>>6, every other term
// rounds to 0), so the general path alone is bit-exact and stays divergence-free.
#include "amv_common.cuh"
#include "amv_tables.cuh"
#include "amv_dct.cuh"
#include "amv_kernels.h"

namespace amv {

constexpr int kBgrWarps = 4;
constexpr int kBgrThreads = kBgrWarps * 32;
constexpr int kBgrSegMB = 5;

struct CwC { enum { W1 = 2841, W2 = 2676, W3 = 2408, W5 = 1609, W6 = 1108, W7 = 565 }; };

__device__ __forceinline__ int iclp(int v) { return max(-256, min(255, v)); }

// idctrow (AmvJpeg.c:1078-1125), general path; 32-bit wrap-around arithmetic like the reference's ints
__device__ __forceinline__ void cw_row(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
    int x0 = (int)(((uint32_t)b0 << 11) + 128u), x1 = (int)((uint32_t)b4 << 11), x2 = b6, x3 = b2, x4 = b1, x5 = b7, x6 = b5, x7 = b3, x8;
    x8 = CwC::W7 * (x4 + x5);
    x4 = x8 + (CwC::W1 - CwC::W7) * x4;
    x5 = x8 - (CwC::W1 + CwC::W7) * x5;
    x8 = CwC::W3 * (x6 + x7);
    x6 = x8 - (CwC::W3 - CwC::W5) * x6;
    x7 = x8 - (CwC::W3 + CwC::W5) * x7;
    x8 = x0 + x1; x0 -= x1;
    x1 = CwC::W6 * (x3 + x2);
    x2 = x1 - (CwC::W2 + CwC::W6) * x2;
    x3 = x1 + (CwC::W2 - CwC::W6) * x3;
    x1 = x4 + x6; x4 -= x6;
    x6 = x5 + x7; x5 -= x7;
    x7 = x8 + x3; x8 -= x3;
    x3 = x0 + x2; x0 -= x2;
    x2 = (181 * (x4 + x5) + 128) >> 8;
    x4 = (181 * (x4 - x5) + 128) >> 8;
    b0 = (x7 + x1) >> 8; b1 = (x3 + x2) >> 8; b2 = (x0 + x4) >> 8; b3 = (x8 + x6) >> 8;
    b4 = (x8 - x6) >> 8; b5 = (x0 - x4) >> 8; b6 = (x3 - x2) >> 8; b7 = (x7 - x1) >> 8;
}

// idctcol (AmvJpeg.c:1127-1175), general path, clamped by iclp
__device__ __forceinline__ void cw_col(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
    int x0 = (int)(((uint32_t)b0 << 8) + 8192u), x1 = (int)((uint32_t)b4 << 8), x2 = b6, x3 = b2, x4 = b1, x5 = b7, x6 = b5, x7 = b3, x8;
    x8 = CwC::W7 * (x4 + x5) + 4;
    x4 = (x8 + (CwC::W1 - CwC::W7) * x4) >> 3;
    x5 = (x8 - (CwC::W1 + CwC::W7) * x5) >> 3;
    x8 = CwC::W3 * (x6 + x7) + 4;
    x6 = (x8 - (CwC::W3 - CwC::W5) * x6) >> 3;
    x7 = (x8 - (CwC::W3 + CwC::W5) * x7) >> 3;
    x8 = x0 + x1; x0 -= x1;
    x1 = CwC::W6 * (x3 + x2) + 4;
    x2 = (x1 - (CwC::W2 + CwC::W6) * x2) >> 3;
    x3 = (x1 + (CwC::W2 - CwC::W6) * x3) >> 3;
    x1 = x4 + x6; x4 -= x6;
    x6 = x5 + x7; x5 -= x7;
    x7 = x8 + x3; x8 -= x3;
    x3 = x0 + x2; x0 -= x2;
    x2 = (181 * (x4 + x5) + 128) >> 8;
    x4 = (181 * (x4 - x5) + 128) >> 8;
    b0 = iclp((x7 + x1) >> 14); b1 = iclp((x3 + x2) >> 14); b2 = iclp((x0 + x4) >> 14); b3 = iclp((x8 + x6) >> 14);
    b4 = iclp((x8 - x6) >> 14); b5 = iclp((x0 - x4) >> 14); b6 = iclp((x3 - x2) >> 14); b7 = iclp((x7 - x1) >> 14);
}

// StoreBuffer's conversion (AmvJpeg.c:808-827): y carries its +128, u / v are centred; clamp to 0..255
__device__ __forceinline__ uint32_t bgr_of(int y, int u, int v) {
    const int y8 = y << 8;
    const int r = __vimin_s32_relu((y8 + 18 * u + 367 * v) >> 8, 255);
    const int g = __vimin_s32_relu((y8 - 159 * u - 220 * v) >> 8, 255);
    const int b = __vimin_s32_relu((y8 + 411 * u - 29 * v) >> 8, 255);
    return (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16);
}

template <bool FAST>
__global__ void __launch_bounds__(kBgrThreads)
k_idct_bgr(const uint32_t *__restrict__ tokens, const uint32_t *__restrict__ blk_off, const uint64_t *__restrict__ slot_off,
           const uint32_t *__restrict__ scan_len, int n, Geom g, int nseg, uint8_t *__restrict__ bgr, int line_bytes,
           uint64_t frame_stride) {
    // per warp: 64 x 32 words.  Phase A uses it as the coefficient column [position][lane]; phase B
    // reuses the first half as the int16 sample tile [sample][lane]
    __shared__ uint32_t tile[kBgrWarps][64 * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t item = (int64_t)blockIdx.x * kBgrWarps + wid;
    if (item >= (int64_t)n * nseg) return;
    const int f = (int)(item / nseg);
    if (scan_len[f] == 0) return;                       // rejected by k_unstuff; the picture is left untouched
    const int m0 = (int)(item - (int64_t)f * nseg) * kBgrSegMB;
    const int total_mb = g.mbw * g.mbh;
    const int nmb = min(kBgrSegMB, total_mb - m0);
    // block role of the lane: lanes 0..19 luma (4 per MB), 20..24 Cb, 25..29 Cr
    int mi, b;
    if (lane < 20) { mi = lane >> 2; b = lane & 3; }
    else if (lane < 25) { mi = lane - 20; b = 4; }
    else { mi = lane - 25; b = 5; }
    const bool active = lane < 30 && mi < nmb;
    uint32_t *col = &tile[wid][lane];
    const uint32_t col_s = smem_addr(col);

    // ---------------- A: tokens -> coefficients -> Fast_IDCT -> sample tile
    int v[64];
    if (active) {
#pragma unroll
        for (int k = 0; k < 64; k++) col[k * 32] = 0;
        const uint32_t bo = blk_off[(uint64_t)f * g.nblk + (uint32_t)(m0 + mi) * 6 + b];
        const uint32_t *tp = tokens + slot_off[f] * 4 + (bo & ((1u << kTokCountShift) - 1u));
        const uint32_t nac = bo >> kTokCountShift;
        uint32_t t = __ldg(tp);
        for (uint32_t a = 0; a <= nac; a++) {
            const uint32_t nx = __ldg(tp + a + 1);
            sts32(col_s + (t >> 26) * 128u, (uint32_t)((int32_t)(t << 6) >> 6));       // 26-bit two's complement value
            t = nx;
        }
#pragma unroll
        for (int k = 0; k < 64; k++) v[k] = (int)col[k * 32];
#pragma unroll
        for (int r = 0; r < 8; r++)
            cw_row(v[8 * r], v[8 * r + 1], v[8 * r + 2], v[8 * r + 3], v[8 * r + 4], v[8 * r + 5], v[8 * r + 6], v[8 * r + 7]);
#pragma unroll
        for (int c = 0; c < 8; c++)
            cw_col(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
    }
    __syncwarp();
    if (active) {
        const int bias = b < 4 ? 128 : 0;               // IQtIZzBlock's offset (:1025-1038)
        uint16_t *st16 = reinterpret_cast<uint16_t *>(&tile[wid][0]) + lane;
#pragma unroll
        for (int k = 0; k < 64; k++) st16[k * 32] = (uint16_t)(v[k] + bias);
    }
    __syncwarp();

    // ---------------- B: 16 pixel rows x (nmb * 4) quads of four pixels
    const int16_t *smp = reinterpret_cast<const int16_t *>(&tile[wid][0]);
    uint8_t *img = bgr + (uint64_t)f * frame_stride;
    const int nq = nmb * 4, ntask = 16 * nq;
    for (int t = lane; t < ntask; t += 32) {
        const int r = t / nq, q = t - r * nq;
        const int qmi = q >> 2, j0 = (q & 3) * 4;
        const int mb = m0 + qmi;
        const int my = mb / g.mbw, mx = mb - my * g.mbw;
        const int Y = my * 16 + r, X0 = mx * 16 + j0;
        if (Y >= g.h || X0 >= g.w) continue;
        const int ylane = qmi * 4 + (r >> 3) * 2 + (j0 >> 3);
        const int yi = (r & 7) * 8 + (j0 & 7);
        const int ci = (r >> 1) * 8 + (j0 >> 1);
        const int y0 = smp[(yi + 0) * 32 + ylane], y1 = smp[(yi + 1) * 32 + ylane], y2 = smp[(yi + 2) * 32 + ylane],
                  y3 = smp[(yi + 3) * 32 + ylane];
        const int u0 = smp[ci * 32 + 20 + qmi], u1 = smp[(ci + 1) * 32 + 20 + qmi];
        const int v0 = smp[ci * 32 + 25 + qmi], v1 = smp[(ci + 1) * 32 + 25 + qmi];
        const uint32_t p0 = bgr_of(y0, u0, v0), p1 = bgr_of(y1, u0, v0), p2 = bgr_of(y2, u1, v1), p3 = bgr_of(y3, u1, v1);
        uint8_t *d = img + (uint64_t)(g.h - 1 - Y) * line_bytes + 3 * X0;
        if (FAST && X0 + 4 <= g.w) {
            uint32_t *d32 = reinterpret_cast<uint32_t *>(d);
            d32[0] = p0 | (p1 << 24);
            d32[1] = (p1 >> 8) | (p2 << 16);
            d32[2] = (p2 >> 16) | (p3 << 8);
        } else {
            const uint32_t px[4] = { p0, p1, p2, p3 };
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (X0 + k < g.w) { d[3 * k] = (uint8_t)px[k]; d[3 * k + 1] = (uint8_t)(px[k] >> 8); d[3 * k + 2] = (uint8_t)(px[k] >> 16); }
        }
    }
}

void launch_idct_bgr(const uint32_t *tokens, const uint32_t *blk_off, const uint64_t *slot_off, const uint32_t *scan_len,
                     int n, const Geom &g, uint8_t *bgr, int line_bytes, uint64_t frame_stride, cudaStream_t s) {
    const int total_mb = g.mbw * g.mbh;
    const int nseg = (total_mb + kBgrSegMB - 1) / kBgrSegMB;
    const int64_t items = (int64_t)n * nseg;
    const unsigned grid = (unsigned)((items + kBgrWarps - 1) / kBgrWarps);
    const bool fast = (((uintptr_t)bgr | (uintptr_t)line_bytes | frame_stride) & 3) == 0;
    if (fast) AMV_LAUNCH(k_idct_bgr<true>, grid, kBgrThreads, 0, s, tokens, blk_off, slot_off, scan_len, n, g, nseg, bgr, line_bytes, frame_stride);
    else      AMV_LAUNCH(k_idct_bgr<false>, grid, kBgrThreads, 0, s, tokens, blk_off, slot_off, scan_len, n, g, nseg, bgr, line_bytes, frame_stride);
}

}  // namespace amv

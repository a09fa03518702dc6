// amv_dct.cuh -- the two integer 8x8 transforms of the AMV path as straight-line
// register code (one thread = one block).  Both are ring arithmetic mod 2^32
// (adds, multiplies) up to the right shifts, so any regrouping of the sums is
// bit-identical to the reference's evaluation order.
#pragma once
#include "amv_common.cuh"

namespace amv {

AMV_HD int sext16(int v) { return (int)(int16_t)v; }

AMV_HD int clamp_u8(int v) {
#if defined(__CUDA_ARCH__)
    return __vimin_s32_relu(v, 255);          // max(min(v,255),0): one VIMNMX on sm_90+
#else
    return v < 0 ? 0 : (v > 255 ? 255 : v);
#endif
}

// ---------------------------------------------------------------- simple_idct
// simple_idct_put (simple_idct.c:390-398): idctRowCondDC (:78-181) on each row,
// results truncated to int16, then idctSparseColPut (:183-253) with clamp.
// W4 = 16383 (:50), row shift 11, column shift 20, column bias W4*32.
struct IdctC { enum { W1 = 22725, W2 = 21407, W3 = 19266, W4 = 16383, W5 = 12873, W6 = 8867, W7 = 4520 }; };

// 1-D butterfly shared by both passes: e0 is the pre-biased DC term.  Every product is taken with
// K * W: the sums are linear mod 2^32, so the outputs are K times the reference's sums (mod 2^32).
// NZ: inputs x[NZ..7] are known to be zero (the folded column pass of idct_put_block<2>); their terms are not evaluated.
template <int K, int NZ = 8>
AMV_HD void idct_1d(int e0, int x1, int x2, int x3, int x4, int x5, int x6, int x7, int (&s)[4], int (&d)[4]) {
    constexpr int W1 = IdctC::W1 * K, W2 = IdctC::W2 * K, W3 = IdctC::W3 * K, W4 = IdctC::W4 * K, W5 = IdctC::W5 * K,
                  W6 = IdctC::W6 * K, W7 = IdctC::W7 * K;
    // Every sum is ONE chain of multiply-adds that starts from the term below it (the even part from e0, the outputs
    // s from the even part), and every difference is 2 * (what the chain started from) - (its end): 36 instead of 44
    // instructions per pass, 28 of them on the FMA-heavy pipe -- the ALU pipe is the busier one in the IDCT kernels.
    auto t1 = [&](int w, int c) { return NZ > 1 ? mad_lo(w, x1, c) : c; };
    auto t2 = [&](int w, int c) { return NZ > 2 ? mad_lo(w, x2, c) : c; };
    auto t3 = [&](int w, int c) { return NZ > 3 ? mad_lo(w, x3, c) : c; };
    auto t4 = [&](int w, int c) { return NZ > 4 ? mad_lo(w, x4, c) : c; };
    auto t5 = [&](int w, int c) { return NZ > 5 ? mad_lo(w, x5, c) : c; };
    auto t6 = [&](int w, int c) { return NZ > 6 ? mad_lo(w, x6, c) : c; };
    auto t7 = [&](int w, int c) { return NZ > 7 ? mad_lo(w, x7, c) : c; };
    auto twice_minus = [](int a, int e) { return (int)(2u * (uint32_t)a - (uint32_t)e); };      // 2a - e mod 2^32
    const int ea = t4(W4, e0), eb = twice_minus(e0, ea);
    const int a0 = t6(W6, t2(W2, ea)), a3 = twice_minus(ea, a0);
    const int a1 = t6(-W2, t2(W6, eb)), a2 = twice_minus(eb, a1);
    s[0] = t7(W7, t5(W5, t3(W3, t1(W1, a0))));
    s[1] = t7(-W5, t5(-W1, t3(-W7, t1(W3, a1))));
    s[2] = t7(W3, t5(W7, t3(-W1, t1(W5, a2))));
    s[3] = t7(-W1, t5(W3, t3(-W5, t1(W7, a3))));
    d[0] = twice_minus(a0, s[0]); d[1] = twice_minus(a1, s[1]); d[2] = twice_minus(a2, s[2]); d[3] = twice_minus(a3, s[3]);
}

// four results of the column pass -> four clamped pixels in one word (x0 in the lowest byte)
AMV_HD uint32_t pack_pixels(int p0, int p1, int p2, int p3) {
#if defined(__CUDA_ARCH__)
    // cvt.pack.sat.u8.s32: d = { c[15:0], sat_u8(a), sat_u8(b) } -- clamp (ff_cropTbl) and pack, two pixels per instruction
    uint32_t hi, w;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(p3), "r"(p2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(p1), "r"(p0), "r"(hi));
    return w;
#else
    return (uint32_t)clamp_u8(p0) | ((uint32_t)clamp_u8(p1) << 8) | ((uint32_t)clamp_u8(p2) << 16) | ((uint32_t)clamp_u8(p3) << 24);
#endif
}

// in : c[32], word 4*r+i = coefficient (r,2i) in the low half, (r,2i+1) in the high half
// out: o[16], words 2*r and 2*r+1 = the 8 pixels of row r, column 0 in the lowest byte
// ROWS: the caller guarantees that coefficient rows ROWS..7 are all zero.  Such a row leaves the row
// pass as zeros (idctRowCondDC's DC-only shortcut yields row[0] << 3 = 0), so it is not evaluated and the
// column pass folds its terms away at compile time -- the result is identical to the full transform.
template <int ROWS = 8>
AMV_HD void idct_put_block(const uint32_t (&c)[32], uint32_t (&o)[16]) {
    int m[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r >= ROWS) {
#pragma unroll
            for (int i = 0; i < 8; i++) m[8 * r + i] = 0;
            continue;
        }
        const uint32_t w0 = c[4 * r], w1 = c[4 * r + 1], w2 = c[4 * r + 2], w3 = c[4 * r + 3];
        const int x0 = sext16((int)w0), x1 = (int)w0 >> 16;
        const int x2 = sext16((int)w1), x3 = (int)w1 >> 16;
        const int x4 = sext16((int)w2), x5 = (int)w2 >> 16;
        const int x6 = sext16((int)w3), x7 = (int)w3 >> 16;
        // The row pass runs scaled by 32: the reference keeps (sum >> 11) truncated to int16, i.e. bits
        // 11..26 of the sum -- which are the top half of 32 * sum, one arithmetic shift away, whatever
        // the sum's upper bits do.  DC-only rows take (row[0] << 3) & 0xffff for every output (:98-103);
        // feeding that value, pre-shifted, as the DC term makes the general path produce it.
        const bool dc_only = ((w0 >> 16) | w1 | w2 | w3) == 0;
        const int e0 = dc_only ? (int)((uint32_t)x0 << 19) : (IdctC::W4 * 32) * x0 + (1 << 15);
        int s[4], d[4];
        idct_1d<32>(e0, x1, x2, x3, x4, x5, x6, x7, s, d);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            m[8 * r + i]     = s[i] >> 16;
            m[8 * r + 7 - i] = d[i] >> 16;
        }
    }
    int px[64];
#pragma unroll
    for (int col = 0; col < 8; col++) {
        int s[4], d[4];
        idct_1d<1, ROWS>(IdctC::W4 * (m[col] + 32), m[8 + col], m[16 + col], m[24 + col], m[32 + col], m[40 + col],
                   m[48 + col], m[56 + col], s, d);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            px[8 * i + col]       = s[i] >> 20;
            px[8 * (7 - i) + col] = d[i] >> 20;
        }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        o[2 * r]     = pack_pixels(px[8 * r], px[8 * r + 1], px[8 * r + 2], px[8 * r + 3]);
        o[2 * r + 1] = pack_pixels(px[8 * r + 4], px[8 * r + 5], px[8 * r + 6], px[8 * r + 7]);
    }
}

// ------------------------------------------------------------------ fdct_islow
// ff_jpeg_fdct_islow (jfdctint.c:184-341): CONST_BITS 13, PASS1_BITS 4.  The
// reference evaluates in 64-bit temporaries and stores int16 between the passes;
// for 8-bit pixel input every stored value fits int16 and every pre-shift sum fits
// int32 (checked exhaustively over the extremal patterns in tests/), so 32-bit
// wrap-around arithmetic reproduces it exactly.
struct FdctC {
    enum { C0_298 = 2446, C0_390 = 3196, C0_541 = 4433, C0_765 = 6270, C0_899 = 7373, C1_175 = 9633,
           C1_501 = 12299, C1_847 = 15137, C1_961 = 16069, C2_053 = 16819, C2_562 = 20995, C3_072 = 25172 };
};

// One 8-point pass.  ROW: outputs 0/4 are shifted up by 4, the others descaled by 9;
// column pass: 0/4 descaled by 4, the others by 17.
template <bool ROW>
AMV_HD void fdct_1d(int &v0, int &v1, int &v2, int &v3, int &v4, int &v5, int &v6, int &v7) {
    constexpr int SH = ROW ? 9 : 17;
    constexpr int RND = 1 << (SH - 1);
    const int p0 = v0 + v7, m0 = v0 - v7, p1 = v1 + v6, m1 = v1 - v6;
    const int p2 = v2 + v5, m2 = v2 - v5, p3 = v3 + v4, m3 = v3 - v4;
    const int q0 = p0 + p3, q3 = p0 - p3, q1 = p1 + p2, q2 = p1 - p2;
    if (ROW) { v0 = (q0 + q1) << 4; v4 = (q0 - q1) << 4; }
    else     { v0 = (q0 + q1 + 8) >> 4; v4 = (q0 - q1 + 8) >> 4; }
    const int r = (q2 + q3) * FdctC::C0_541 + RND;
    v2 = (r + q3 * FdctC::C0_765) >> SH;
    v6 = (r - q2 * FdctC::C1_847) >> SH;
    const int sc = m3 + m1, sd = m2 + m0;
    const int z = (sc + sd) * FdctC::C1_175 + RND;
    const int a = (m3 + m0) * -FdctC::C0_899, b = (m2 + m1) * -FdctC::C2_562;
    const int c = sc * -FdctC::C1_961 + z, d = sd * -FdctC::C0_390 + z;
    v7 = (m3 * FdctC::C0_298 + a + c) >> SH;
    v5 = (m2 * FdctC::C2_053 + b + d) >> SH;
    v3 = (m1 * FdctC::C3_072 + b + c) >> SH;
    v1 = (m0 * FdctC::C1_501 + a + d) >> SH;
}

// in place on 64 ints (raster order), pixels in -> coefficients out
AMV_HD void fdct_block(int (&b)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++)
        fdct_1d<true>(b[8 * r], b[8 * r + 1], b[8 * r + 2], b[8 * r + 3], b[8 * r + 4], b[8 * r + 5], b[8 * r + 6], b[8 * r + 7]);
#pragma unroll
    for (int c = 0; c < 8; c++)
        fdct_1d<false>(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
}

// ------------------------------------------------- fdct_islow, regrouped for the two integer pipes
// Up to its descale an 8-point pass is an exact integer linear map, so any regrouping of its sums gives the same
// bits (mod 2^32, and every true sum fits).  LL&M's factorisation minimises multiplies, which is the wrong economy on
// an SM whose adds / shifts / byte extractions (ALU pipe) and multiply-adds / dot products (FMA pipe) each issue every
// second cycle per scheduler: the factorised pass is 12 multiply-adds against 32 adds and shifts, plus 8 byte
// extractions per pixel row.  The forms below move work to the multiply-add side:
//  * row pass on the PACKED pixel bytes as loaded: output k = sum_i C[k][i] * d[i] is four two-way dot products
//    (IDP.2A: two signed 16-bit constants x two unsigned bytes, accumulating; the rounding constant is the initial
//    accumulator), outputs 0 and 4 two four-way dot products with +-16 (IDP.4A).  No byte is ever extracted.
//  * column pass with the odd half written out (four multiply-adds per output on the differences v[i] - v[7-i]).
// fdct_lin(k, i): coefficient of input i in output k of the factorised pass before rounding and shift (outputs 0 / 4:
// before their << 4 or >> 4), derived at compile time from the factorised form itself.
constexpr int fdct_lin(int k, int i) {
    int v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    v[i] = 1;
    const int p0 = v[0] + v[7], m0 = v[0] - v[7], p1 = v[1] + v[6], m1 = v[1] - v[6];
    const int p2 = v[2] + v[5], m2 = v[2] - v[5], p3 = v[3] + v[4], m3 = v[3] - v[4];
    const int q0 = p0 + p3, q3 = p0 - p3, q1 = p1 + p2, q2 = p1 - p2;
    const int r = (q2 + q3) * FdctC::C0_541;
    const int sc = m3 + m1, sd = m2 + m0;
    const int z = (sc + sd) * FdctC::C1_175;
    const int a = (m3 + m0) * -FdctC::C0_899, b = (m2 + m1) * -FdctC::C2_562;
    const int c = sc * -FdctC::C1_961 + z, d = sd * -FdctC::C0_390 + z;
    switch (k) {
    case 0: return q0 + q1;
    case 4: return q0 - q1;
    case 2: return r + q3 * FdctC::C0_765;
    case 6: return r - q2 * FdctC::C1_847;
    case 7: return m3 * FdctC::C0_298 + a + c;
    case 5: return m2 * FdctC::C2_053 + b + d;
    case 3: return m1 * FdctC::C3_072 + b + c;
    default: return m0 * FdctC::C1_501 + a + d;
    }
}
constexpr uint32_t pack_s16x2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
constexpr uint32_t pack_s8x4(int b0, int b1, int b2, int b3) {
    return ((uint32_t)b0 & 0xffu) | (((uint32_t)b1 & 0xffu) << 8) | (((uint32_t)b2 & 0xffu) << 16) | ((uint32_t)b3 << 24);
}

// acc + c.lo16 * byte0(d) + c.hi16 * byte1(d)   (HI: bytes 2 and 3); constants signed, pixel bytes unsigned
template <bool HI>
AMV_HD int dot2_s16_u8(uint32_t c, uint32_t d, int acc) {
#if defined(__CUDA_ARCH__)
    int r;
    if (HI) asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(c), "r"(d), "r"(acc));
    else    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(c), "r"(d), "r"(acc));
    return r;
#else
    const uint32_t dd = HI ? d >> 16 : d;
    return (int)((uint32_t)acc + (uint32_t)((int)(int16_t)(c & 0xffffu) * (int)(dd & 0xffu)) +
                 (uint32_t)((int)(int16_t)(c >> 16) * (int)((dd >> 8) & 0xffu)));
#endif
}
// acc + sum of four signed constant bytes x unsigned pixel bytes
AMV_HD int dot4_s8_u8(uint32_t c, uint32_t d, int acc) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(c), "r"(d), "r"(acc));
    return r;
#else
    uint32_t s = (uint32_t)acc;
    for (int j = 0; j < 4; j++) s += (uint32_t)((int)(int8_t)(c >> (8 * j)) * (int)((d >> (8 * j)) & 0xffu));
    return (int)s;
#endif
}

// row pass, output K, from the row's eight pixels as two little-endian words
template <int K>
AMV_HD int fdct_row_dot(uint32_t lo, uint32_t hi) {
    if (K == 0 || K == 4) {
        constexpr uint32_t c0 = pack_s8x4(16 * fdct_lin(K, 0), 16 * fdct_lin(K, 1), 16 * fdct_lin(K, 2), 16 * fdct_lin(K, 3));
        constexpr uint32_t c1 = pack_s8x4(16 * fdct_lin(K, 4), 16 * fdct_lin(K, 5), 16 * fdct_lin(K, 6), 16 * fdct_lin(K, 7));
        return dot4_s8_u8(c1, hi, dot4_s8_u8(c0, lo, 0));
    } else {
        constexpr uint32_t c01 = pack_s16x2(fdct_lin(K, 0), fdct_lin(K, 1)), c23 = pack_s16x2(fdct_lin(K, 2), fdct_lin(K, 3));
        constexpr uint32_t c45 = pack_s16x2(fdct_lin(K, 4), fdct_lin(K, 5)), c67 = pack_s16x2(fdct_lin(K, 6), fdct_lin(K, 7));
        int acc = 1 << 8;
        acc = dot2_s16_u8<false>(c01, lo, acc);
        acc = dot2_s16_u8<true>(c23, lo, acc);
        acc = dot2_s16_u8<false>(c45, hi, acc);
        acc = dot2_s16_u8<true>(c67, hi, acc);
        return acc >> 9;
    }
}

// column pass with the odd outputs written out; EVEN_DIRECT also writes outputs 2 and 6 as two multiply-adds each
template <int K>
AMV_HD int fdct_col_odd(int m0, int m1, int m2, int m3) {
    constexpr int c0 = fdct_lin(K, 0), c1 = fdct_lin(K, 1), c2 = fdct_lin(K, 2), c3 = fdct_lin(K, 3);
    return (m0 * c0 + m1 * c1 + m2 * c2 + m3 * c3 + (1 << 16)) >> 17;
}
template <bool EVEN_DIRECT>
AMV_HD void fdct_col_direct(int &v0, int &v1, int &v2, int &v3, int &v4, int &v5, int &v6, int &v7) {
    const int p0 = v0 + v7, m0 = v0 - v7, p1 = v1 + v6, m1 = v1 - v6;
    const int p2 = v2 + v5, m2 = v2 - v5, p3 = v3 + v4, m3 = v3 - v4;
    const int q0 = p0 + p3, q3 = p0 - p3, q1 = p1 + p2, q2 = p1 - p2;
    v0 = (q0 + q1 + 8) >> 4;
    v4 = (q0 - q1 + 8) >> 4;
    if (EVEN_DIRECT) {
        v2 = (q3 * (FdctC::C0_541 + FdctC::C0_765) + q2 * FdctC::C0_541 + (1 << 16)) >> 17;
        v6 = (q3 * FdctC::C0_541 + q2 * (FdctC::C0_541 - FdctC::C1_847) + (1 << 16)) >> 17;
    } else {
        const int r = (q2 + q3) * FdctC::C0_541 + (1 << 16);
        v2 = (r + q3 * FdctC::C0_765) >> 17;
        v6 = (r - q2 * FdctC::C1_847) >> 17;
    }
    v1 = fdct_col_odd<1>(m0, m1, m2, m3);
    v3 = fdct_col_odd<3>(m0, m1, m2, m3);
    v5 = fdct_col_odd<5>(m0, m1, m2, m3);
    v7 = fdct_col_odd<7>(m0, m1, m2, m3);
}

// px: the block's pixel rows as loaded, words 2r and 2r + 1 = row r (column 0 in the lowest byte); b: coefficients, raster order.
// FORM bit 0: row pass by dot products on the packed bytes (else bytes extracted, factorised pass);
// FORM bit 1: column pass with the odd half written out; bit 2: outputs 2 / 6 written out as well.
template <int FORM>
AMV_HD void fdct_block_px(const uint32_t (&px)[16], int (&b)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t lo = px[2 * r], hi = px[2 * r + 1];
        if (FORM & 1) {
            b[8 * r]     = fdct_row_dot<0>(lo, hi);
            b[8 * r + 1] = fdct_row_dot<1>(lo, hi);
            b[8 * r + 2] = fdct_row_dot<2>(lo, hi);
            b[8 * r + 3] = fdct_row_dot<3>(lo, hi);
            b[8 * r + 4] = fdct_row_dot<4>(lo, hi);
            b[8 * r + 5] = fdct_row_dot<5>(lo, hi);
            b[8 * r + 6] = fdct_row_dot<6>(lo, hi);
            b[8 * r + 7] = fdct_row_dot<7>(lo, hi);
        } else {
#pragma unroll
            for (int x = 0; x < 4; x++) {
                b[8 * r + x]     = (int)((lo >> (8 * x)) & 0xffu);
                b[8 * r + 4 + x] = (int)((hi >> (8 * x)) & 0xffu);
            }
            fdct_1d<true>(b[8 * r], b[8 * r + 1], b[8 * r + 2], b[8 * r + 3], b[8 * r + 4], b[8 * r + 5], b[8 * r + 6], b[8 * r + 7]);
        }
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
        if (FORM & 2) {
            if (FORM & 4) fdct_col_direct<true>(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
            else          fdct_col_direct<false>(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
        } else {
            fdct_1d<false>(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
        }
    }
}

// ------------------------------------------------------------------ quantiser
// dct_quantize_c intra, bias 0 (mpegvideo_enc.c:3647-3725, :492-496):
//   DC: (b + 32) / 64 (C division);  AC: sign(b) * ((|b| * qmat) >> 22), clipped to +-1023
// with qmat = (1<<22) / (8*M).  qm10 = qmat << 10 turns the shift into a high-half multiply.
AMV_HD int quant_dc(int b) { return (b + 32) / 64; }
AMV_HD int quant_ac(int b, uint32_t qm10) {
    const uint32_t a = (uint32_t)(b < 0 ? -b : b);
#if defined(__CUDA_ARCH__)
    int q = (int)__umulhi(a, qm10);
#else
    int q = (int)(((uint64_t)a * qm10) >> 32);
#endif
    // clip_coeffs' +-1023 (mpegvideo_enc.c:1403-1432) cannot trigger: for qscale >= 2 the matrix
    // entries are >= 4, so |q| <= 16320 / 32 = 510 for any 8-bit picture (checked in tests/).
    return b < 0 ? -q : q;
}

}  // namespace amv

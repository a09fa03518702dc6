/*
 * amv_oracle.c -- CPU restatement of the AMV intra-frame codec path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (amv-codec-tools_b200/,
 * libamvcuda) includes, links or calls this file.  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg have a bit-exact,
 * single-threaded, scalar statement of what the reference computes.
 *
 * PARITY PINNED: tests/test_oracle_vs_ref.py checks every function below
 * against the unmodified reference compiled in place (oracle/build_ref.sh ->
 * oracle/_ref/libamvref.so) and against the in-tree fixture
 * C-AMVDecoder/bin/AMV1.amv; tests/golden/ holds vectors produced by that
 * reference build so the pin also holds where /root/reference is absent.
 *
 * Every function cites the reference lines it restates
 * (paths relative to AMVmuxer/ffmpeg/libavcodec/).  The code is written from
 * the arithmetic, not transcribed: loops, table layouts and bit I/O are ours.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define AMVO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ tables */

/* Decoder quantiser: the two fixed tables the reference wraps every AMV packet
 * with (sp5xdec.c:40,60-61 -> sp5x.h "index 5, Q60", rows 10 and 11), listed
 * in the order they appear in a DQT segment, i.e. zigzag order. */
static const uint8_t kDecQuantZZ[2][64] = {
  { 13,  9, 10, 11, 10,  8, 13, 11, 10, 11, 14, 14, 13, 15, 19, 32,
    21, 19, 18, 18, 19, 39, 28, 30, 23, 32, 46, 41, 49, 48, 46, 41,
    45, 44, 51, 58, 74, 62, 51, 54, 70, 55, 44, 45, 64, 87, 65, 70,
    76, 78, 82, 83, 82, 50, 62, 90, 97, 90, 80, 96, 74, 81, 82, 79 },
  { 14, 14, 14, 19, 17, 19, 38, 21, 21, 38, 79, 53, 45, 53, 79, 79,
    79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79,
    79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79,
    79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79 },
};

/* Encoder quantiser base: MPEG-1 default intra matrix, raster order
 * (mpeg12data.c:30-39; used by mpegvideo_enc.c:2866-2873). */
static const uint8_t kEncBase[64] = {
   8, 16, 19, 22, 26, 27, 29, 34,  16, 16, 22, 24, 27, 29, 34, 37,
  19, 22, 26, 27, 29, 34, 34, 38,  22, 22, 26, 27, 29, 34, 37, 40,
  22, 26, 27, 29, 32, 35, 40, 48,  26, 27, 29, 32, 35, 40, 48, 58,
  26, 27, 29, 34, 38, 46, 56, 69,  27, 29, 35, 38, 46, 56, 69, 83,
};

/* JPEG Annex K.3 Huffman specifications: count of codes per length 1..16,
 * then the symbols in code order (mjpeg.c:65-126; the same bytes sit in the
 * DHT template sp5x.h:76-132 the decoder parses). */
static const uint8_t kCnt[4][16] = {
  /* DC luma   */ { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 },
  /* DC chroma */ { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 },
  /* AC luma   */ { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d },
  /* AC chroma */ { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 },
};
static const uint8_t kSymDC[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t kSymACL[162] = {
  0x01,0x02,0x03,0x00,0x04,0x11,0x05,0x12,0x21,0x31,0x41,0x06,0x13,0x51,0x61,0x07,
  0x22,0x71,0x14,0x32,0x81,0x91,0xa1,0x08,0x23,0x42,0xb1,0xc1,0x15,0x52,0xd1,0xf0,
  0x24,0x33,0x62,0x72,0x82,0x09,0x0a,0x16,0x17,0x18,0x19,0x1a,0x25,0x26,0x27,0x28,
  0x29,0x2a,0x34,0x35,0x36,0x37,0x38,0x39,0x3a,0x43,0x44,0x45,0x46,0x47,0x48,0x49,
  0x4a,0x53,0x54,0x55,0x56,0x57,0x58,0x59,0x5a,0x63,0x64,0x65,0x66,0x67,0x68,0x69,
  0x6a,0x73,0x74,0x75,0x76,0x77,0x78,0x79,0x7a,0x83,0x84,0x85,0x86,0x87,0x88,0x89,
  0x8a,0x92,0x93,0x94,0x95,0x96,0x97,0x98,0x99,0x9a,0xa2,0xa3,0xa4,0xa5,0xa6,0xa7,
  0xa8,0xa9,0xaa,0xb2,0xb3,0xb4,0xb5,0xb6,0xb7,0xb8,0xb9,0xba,0xc2,0xc3,0xc4,0xc5,
  0xc6,0xc7,0xc8,0xc9,0xca,0xd2,0xd3,0xd4,0xd5,0xd6,0xd7,0xd8,0xd9,0xda,0xe1,0xe2,
  0xe3,0xe4,0xe5,0xe6,0xe7,0xe8,0xe9,0xea,0xf1,0xf2,0xf3,0xf4,0xf5,0xf6,0xf7,0xf8,
  0xf9,0xfa };
static const uint8_t kSymACC[162] = {
  0x00,0x01,0x02,0x03,0x11,0x04,0x05,0x21,0x31,0x06,0x12,0x41,0x51,0x07,0x61,0x71,
  0x13,0x22,0x32,0x81,0x08,0x14,0x42,0x91,0xa1,0xb1,0xc1,0x09,0x23,0x33,0x52,0xf0,
  0x15,0x62,0x72,0xd1,0x0a,0x16,0x24,0x34,0xe1,0x25,0xf1,0x17,0x18,0x19,0x1a,0x26,
  0x27,0x28,0x29,0x2a,0x35,0x36,0x37,0x38,0x39,0x3a,0x43,0x44,0x45,0x46,0x47,0x48,
  0x49,0x4a,0x53,0x54,0x55,0x56,0x57,0x58,0x59,0x5a,0x63,0x64,0x65,0x66,0x67,0x68,
  0x69,0x6a,0x73,0x74,0x75,0x76,0x77,0x78,0x79,0x7a,0x82,0x83,0x84,0x85,0x86,0x87,
  0x88,0x89,0x8a,0x92,0x93,0x94,0x95,0x96,0x97,0x98,0x99,0x9a,0xa2,0xa3,0xa4,0xa5,
  0xa6,0xa7,0xa8,0xa9,0xaa,0xb2,0xb3,0xb4,0xb5,0xb6,0xb7,0xb8,0xb9,0xba,0xc2,0xc3,
  0xc4,0xc5,0xc6,0xc7,0xc8,0xc9,0xca,0xd2,0xd3,0xd4,0xd5,0xd6,0xd7,0xd8,0xd9,0xda,
  0xe2,0xe3,0xe4,0xe5,0xe6,0xe7,0xe8,0xe9,0xea,0xf2,0xf3,0xf4,0xf5,0xf6,0xf7,0xf8,
  0xf9,0xfa };

/* IMA ADPCM tables (adpcm.c:56-75). */
static const int16_t kImaStep[89] = {
      7,     8,     9,    10,    11,    12,    13,    14,    16,    17,
     19,    21,    23,    25,    28,    31,    34,    37,    41,    45,
     50,    55,    60,    66,    73,    80,    88,    97,   107,   118,
    130,   143,   157,   173,   190,   209,   230,   253,   279,   307,
    337,   371,   408,   449,   494,   544,   598,   658,   724,   796,
    876,   963,  1060,  1166,  1282,  1411,  1552,  1707,  1878,  2066,
   2272,  2499,  2749,  3024,  3327,  3660,  4026,  4428,  4871,  5358,
   5894,  6484,  7132,  7845,  8630,  9493, 10442, 11487, 12635, 13899,
  15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767 };
static const int8_t kImaIdxAdj[8] = { -1, -1, -1, -1, 2, 4, 6, 8 };

/* derived tables, filled once */
static uint8_t  g_zz[64];            /* zigzag scan position -> raster index (dsputil.c:50-59) */
static uint8_t  g_hlen[4][256];      /* code length per symbol (mjpeg.c:129-147) */
static uint16_t g_hcode[4][256];     /* right-aligned code per symbol */
static int g_ready;

static void build_tables(void)
{
    if (g_ready) return;
    /* zigzag: walk the anti-diagonals of the 8x8 grid, alternating direction */
    int n = 0;
    for (int d = 0; d < 15; d++) {
        for (int t = 0; t <= d; t++) {
            int r = (d & 1) ? t : d - t, c = d - r;
            if (r < 8 && c < 8) g_zz[n++] = (uint8_t)(r * 8 + c);
        }
    }
    /* canonical Huffman codes: consecutive integers within a length, doubled
     * when the length grows (mjpeg.c:129-147) */
    for (int t = 0; t < 4; t++) {
        const uint8_t *sym = t < 2 ? kSymDC : (t == 2 ? kSymACL : kSymACC);
        unsigned next = 0; int k = 0;
        memset(g_hlen[t], 0, 256);
        for (int len = 1; len <= 16; len++) {
            for (int j = 0; j < kCnt[t][len - 1]; j++, k++) {
                g_hlen[t][sym[k]] = (uint8_t)len;
                g_hcode[t][sym[k]] = (uint16_t)next++;
            }
            next <<= 1;
        }
    }
    g_ready = 1;
}

AMVO_API void amvo_get_zigzag(uint8_t out[64]) { build_tables(); memcpy(out, g_zz, 64); }
AMVO_API void amvo_get_huff(int table, uint8_t len[256], uint16_t code[256])
{
    build_tables(); memcpy(len, g_hlen[table], 256); memcpy(code, g_hcode[table], 512);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------- inverse transform (dec) */

/* simple_idct_put (simple_idct.c:390-398): rows in place on int16 with the
 * DC-only shortcut (:93-127), then columns with clamp-and-store (:183-253).
 * W4 is 16383 (:50); sums are 32-bit two's complement.  We keep the sums in
 * uint32_t so wrap-around is defined, and convert at the shifts. */
enum { W1 = 22725, W2 = 21407, W3 = 19266, W4 = 16383, W5 = 12873, W6 = 8867, W7 = 4520 };

static inline int32_t asr32(uint32_t v, int s) { return (int32_t)v >> s; }

static void idct_row(int16_t *r)
{
    if (!(r[1] | r[2] | r[3] | r[4] | r[5] | r[6] | r[7])) {
        /* (row[0] << 3) & 0xffff replicated (:98-103) */
        int16_t dc = (int16_t)(uint16_t)((uint32_t)(int32_t)r[0] << 3);
        for (int i = 0; i < 8; i++) r[i] = dc;
        return;
    }
    uint32_t e0 = (uint32_t)(W4 * r[0]) + (1u << 10);
    uint32_t e[4] = { e0 + (uint32_t)(W2 * r[2]), e0 + (uint32_t)(W6 * r[2]),
                      e0 - (uint32_t)(W6 * r[2]), e0 - (uint32_t)(W2 * r[2]) };
    uint32_t o[4] = { (uint32_t)(W1 * r[1]) + (uint32_t)(W3 * r[3]),
                      (uint32_t)(W3 * r[1]) - (uint32_t)(W7 * r[3]),
                      (uint32_t)(W5 * r[1]) - (uint32_t)(W1 * r[3]),
                      (uint32_t)(W7 * r[1]) - (uint32_t)(W5 * r[3]) };
    /* rows 4..7 contribute nothing when zero, so the :147 test is only a speed-up */
    e[0] += (uint32_t)(W4 * r[4]) + (uint32_t)(W6 * r[6]);
    e[1] -= (uint32_t)(W4 * r[4]) + (uint32_t)(W2 * r[6]);
    e[2] += (uint32_t)(W2 * r[6]) - (uint32_t)(W4 * r[4]);
    e[3] += (uint32_t)(W4 * r[4]) - (uint32_t)(W6 * r[6]);
    o[0] += (uint32_t)(W5 * r[5]) + (uint32_t)(W7 * r[7]);
    o[1] -= (uint32_t)(W1 * r[5]) + (uint32_t)(W5 * r[7]);
    o[2] += (uint32_t)(W7 * r[5]) + (uint32_t)(W3 * r[7]);
    o[3] += (uint32_t)(W3 * r[5]) - (uint32_t)(W1 * r[7]);
    for (int i = 0; i < 4; i++) {
        r[i]     = (int16_t)asr32(e[i] + o[i], 11);
        r[7 - i] = (int16_t)asr32(e[i] - o[i], 11);
    }
}

static void idct_col_put(uint8_t *dst, int ls, const int16_t *c, uint8_t *undef)
{
    /* a0 = W4 * (col[0] + 32)  (:190; (1<<19)/W4 == 32) */
    uint32_t e0 = (uint32_t)(W4 * (c[0] + 32));
    uint32_t e[4] = { e0 + (uint32_t)(W2 * c[16]), e0 + (uint32_t)(W6 * c[16]),
                      e0 - (uint32_t)(W6 * c[16]), e0 - (uint32_t)(W2 * c[16]) };
    uint32_t o[4] = { (uint32_t)(W1 * c[8]) + (uint32_t)(W3 * c[24]),
                      (uint32_t)(W3 * c[8]) - (uint32_t)(W7 * c[24]),
                      (uint32_t)(W5 * c[8]) - (uint32_t)(W1 * c[24]),
                      (uint32_t)(W7 * c[8]) - (uint32_t)(W5 * c[24]) };
    e[0] += (uint32_t)(W4 * c[32]) + (uint32_t)(W6 * c[48]);
    e[1] -= (uint32_t)(W4 * c[32]) + (uint32_t)(W2 * c[48]);
    e[2] += (uint32_t)(W2 * c[48]) - (uint32_t)(W4 * c[32]);
    e[3] += (uint32_t)(W4 * c[32]) - (uint32_t)(W6 * c[48]);
    o[0] += (uint32_t)(W5 * c[40]) + (uint32_t)(W7 * c[56]);
    o[1] -= (uint32_t)(W1 * c[40]) + (uint32_t)(W5 * c[56]);
    o[2] += (uint32_t)(W7 * c[40]) + (uint32_t)(W3 * c[56]);
    o[3] += (uint32_t)(W3 * c[40]) - (uint32_t)(W1 * c[56]);
    for (int i = 0; i < 4; i++) {
        /* ff_cropTbl (dsputil.c:3812-3820) is a clamp to 0..255 over its domain;
         * outside it the reference is undefined, we clamp */
        int hi = asr32(e[i] + o[i], 20), lo = asr32(e[i] - o[i], 20);
        dst[i * ls]       = (uint8_t)clampi(hi, 0, 255);
        dst[(7 - i) * ls] = (uint8_t)clampi(lo, 0, 255);
        if (undef) {   /* index outside ff_cropTbl's -1024..1279: the reference reads foreign memory */
            undef[i * ls]       = hi < -1024 || hi > 1279;
            undef[(7 - i) * ls] = lo < -1024 || lo > 1279;
        }
    }
}

/* out[64]: the 8x8 pixels, row-major */
AMVO_API void amvo_idct_put_ex(const int16_t in[64], uint8_t out[64], uint8_t *undef /* 64 or NULL */)
{
    int16_t b[64];
    memcpy(b, in, sizeof(b));
    for (int r = 0; r < 8; r++) idct_row(b + 8 * r);
    for (int c = 0; c < 8; c++) idct_col_put(out + c, 8, b + c, undef ? undef + c : NULL);
}
AMVO_API void amvo_idct_put(const int16_t in[64], uint8_t out[64]) { amvo_idct_put_ex(in, out, NULL); }

/* -------------------------------------------------- forward transform (enc) */

/* ff_jpeg_fdct_islow (jfdctint.c:184-341): LL&M 1-D DCT, CONST_BITS 13,
 * PASS1_BITS 4 (:127-128); rows first (outputs kept <<4, stored as int16),
 * then columns (descaled by 4 resp. 13+4). */
/* the reference's temporaries are int_fast32_t, i.e. 64-bit on the x86-64 build
 * it is pinned against; for 8-bit pixel input nothing exceeds 32 bits anyway
 * (8 + CONST_BITS + PASS1_BITS = 25 <= 26, jfdctint.c:120-124) */
typedef int64_t fdct_t;
static inline fdct_t descale(fdct_t v, int n) { return (v + ((fdct_t)1 << (n - 1))) >> n; }

static void fdct_1d(const fdct_t s[8], fdct_t out[8], int sh_even, int sh_odd, int row_pass)
{
    enum { C0_298 = 2446, C0_390 = 3196, C0_541 = 4433, C0_765 = 6270, C0_899 = 7373,
           C1_175 = 9633, C1_501 = 12299, C1_847 = 15137, C1_961 = 16069,
           C2_053 = 16819, C2_562 = 20995, C3_072 = 25172 };
    fdct_t p0 = s[0] + s[7], m0 = s[0] - s[7];
    fdct_t p1 = s[1] + s[6], m1 = s[1] - s[6];
    fdct_t p2 = s[2] + s[5], m2 = s[2] - s[5];
    fdct_t p3 = s[3] + s[4], m3 = s[3] - s[4];
    /* even half */
    fdct_t q0 = p0 + p3, q3 = p0 - p3, q1 = p1 + p2, q2 = p1 - p2;
    if (row_pass) { out[0] = (q0 + q1) * (1 << sh_even); out[4] = (q0 - q1) * (1 << sh_even); }
    else          { out[0] = descale(q0 + q1, sh_even); out[4] = descale(q0 - q1, sh_even); }
    fdct_t r = (q2 + q3) * C0_541;
    out[2] = descale(r + q3 * C0_765, sh_odd);
    out[6] = descale(r - q2 * C1_847, sh_odd);
    /* odd half: m3,m2,m1,m0 are the paper's i0..i3 */
    fdct_t a = m3 + m0, b = m2 + m1, c = m3 + m1, d = m2 + m0;
    fdct_t z = (c + d) * C1_175;
    fdct_t t3 = m3 * C0_298, t2 = m2 * C2_053, t1 = m1 * C3_072, t0 = m0 * C1_501;
    a *= -C0_899; b *= -C2_562;
    c = c * -C1_961 + z;
    d = d * -C0_390 + z;
    out[7] = descale(t3 + a + c, sh_odd);
    out[5] = descale(t2 + b + d, sh_odd);
    out[3] = descale(t1 + b + c, sh_odd);
    out[1] = descale(t0 + a + d, sh_odd);
}

AMVO_API void amvo_fdct_islow(int16_t blk[64])
{
    fdct_t in[8], out[8];
    for (int r = 0; r < 8; r++) {
        for (int i = 0; i < 8; i++) in[i] = blk[8 * r + i];
        fdct_1d(in, out, 4, 13 - 4, 1);
        for (int i = 0; i < 8; i++) blk[8 * r + i] = (int16_t)out[i];
    }
    for (int c = 0; c < 8; c++) {
        for (int i = 0; i < 8; i++) in[i] = blk[8 * i + c];
        fdct_1d(in, out, 4, 13 + 4, 0);
        for (int i = 0; i < 8; i++) blk[8 * i + c] = (int16_t)out[i];
    }
}

/* ---------------------------------------------------------------- quantiser */

/* qscale from AVFrame.quality / lambda (mpegvideo_enc.c:143-148, qmin 2 qmax 31
 * utils.c:497-498) */
AMVO_API int amvo_qscale_from_lambda(int lambda, int qmin, int qmax)
{
    return clampi((lambda * 139 + 128 * 64) >> 14, qmin, qmax);
}

/* Per-frame matrix: M[0]=8, M[i]=clip_u8(base[i]*qscale>>3); multiplier
 * (1<<22)/(8*M[i]) (mpegvideo_enc.c:2866-2877, ff_convert_matrix :69-91). */
AMVO_API void amvo_enc_qmat(int qscale, int32_t qmat[64])
{
    for (int i = 0; i < 64; i++) {
        int m = i ? clampi((kEncBase[i] * qscale) >> 3, 0, 255) : kEncBase[0];
        qmat[i] = (int32_t)((1u << 22) / (unsigned)(8 * (m ? m : 1)));
        if (!m) qmat[i] = 0; /* unreachable for qscale >= 2 */
    }
}

/* dct_quantize_c intra path (mpegvideo_enc.c:3647-3725) after the FDCT:
 * DC (b+32)/64 ; AC sign * (|b*qmat| >> 22) (bias 0 :492-496) ; returns the last
 * zigzag position holding a non-zero AC (0 if none); AC clipped to +-1023
 * (clip_coeffs :1403-1432 with mjpegenc.c:55-56). */
AMVO_API int amvo_quantize(int16_t blk[64], const int32_t qmat[64])
{
    build_tables();
    int last = 0;
    blk[0] = (int16_t)((blk[0] + 32) / 64);
    for (int k = 1; k < 64; k++) {
        int j = g_zz[k];
        int32_t lv = (int32_t)((uint32_t)(int32_t)blk[j] * (uint32_t)qmat[j]);
        int q = 0;
        if (lv >= (1 << 22))        q =  (lv >> 22);
        else if (lv <= -(1 << 22))  q = -(int)((uint32_t)(-(int64_t)lv) >> 22);
        q = clampi(q, -1023, 1023);
        blk[j] = (int16_t)q;
        if (q) last = k;
    }
    return last;
}

/* ---------------------------------------------------------- MSB-first writer */

typedef struct { uint8_t *p; size_t cap, nbytes; uint32_t acc; int nacc; int overflow; } bitw;

static void bw_put(bitw *w, int n, uint32_t v)       /* n <= 24 (bitstream.h:213-250 semantics) */
{
    w->acc = (w->acc << n) | (v & ((1u << n) - 1));
    w->nacc += n;
    while (w->nacc >= 8) {
        uint8_t b = (uint8_t)(w->acc >> (w->nacc - 8));
        if (w->nbytes < w->cap) w->p[w->nbytes] = b; else w->overflow = 1;
        w->nbytes++;
        w->nacc -= 8;
    }
}

static int bit_width(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

/* magnitude category + low bits (ff_mjpeg_encode_dc mjpegenc.c:357-377 and the
 * AC branch :413-426): negative values send (v-1) masked to the category. */
static void put_coef(bitw *w, int tbl, int run, int v)
{
    int mag = v < 0 ? -v : v;
    int nb = bit_width((unsigned)mag);
    int sym = (run << 4) | nb;
    bw_put(w, g_hlen[tbl][sym], g_hcode[tbl][sym]);
    if (nb) bw_put(w, nb, (uint32_t)(v < 0 ? v - 1 : v));
}

/* ------------------------------------------------------------------ encoder */

/* Row the reference starts reading a plane from before walking upwards
 * (mjpegenc.c:467-470): vs*(8*mb_h - ((h/2)&7)) - 1, vs = 2 luma / 1 chroma. */
static int flip_start_row(int h, int vs)
{
    int mb_h = (h + 15) / 16;
    return vs * (8 * mb_h - ((h / 2) & 7)) - 1;
}

/* Encode one frame into out[0..cap). Returns packet size, or
 *  -1 unsupported geometry (reference would read outside the picture),
 *  -2 packet does not fit.
 * amv_encode_picture (mjpegenc.c:454-472) -> MPV_encode_picture
 * (mpegvideo_enc.c:1205-1352) -> encode_thread (:2004-2631) ->
 * encode_mb_internal (:1457-1750) -> ff_mjpeg_encode_mb (mjpegenc.c:437-450)
 * -> trailer (mjpegenc.c:345-355). */
AMVO_API int amvo_encode_frame(const uint8_t *py, const uint8_t *pu, const uint8_t *pv,
                               int ls_y, int ls_c, int w, int h, int qscale,
                               uint8_t *out, uint32_t cap)
{
    build_tables();
    if (w < 2 || h < 2 || qscale < 1 || qscale > 31) return -1;
    const int mbw = (w + 15) / 16, mbh = (h + 15) / 16;
    const int cw = w >> 1, chh = h >> 1;                      /* chroma extent the reference copies (:866-867) */
    const int y0 = flip_start_row(h, 2), c0 = flip_start_row(h, 1);
    if (y0 != h - 1 || c0 > ((h + 1) >> 1) - 1 || c0 - (chh - 1) < 0) return -1;

    int32_t qmat[64];
    amvo_enc_qmat(qscale, qmat);

    /* entropy-coded segment is produced un-stuffed first, stuffing afterwards
     * (escape_FF mjpegenc.c:282-336 runs once over the finished frame) */
    size_t rawcap = (size_t)mbw * mbh * 6 * 256 + 64;
    uint8_t *raw = (uint8_t *)malloc(rawcap);
    bitw bw = { raw, rawcap, 0, 0, 0, 0 };
    int pred[3] = { 128, 128, 128 };                           /* mpegvideo_enc.c:2033-2036 */

    for (int my = 0; my < mbh; my++)
    for (int mx = 0; mx < mbw; mx++)
    for (int b = 0; b < 6; b++) {
        int16_t blk[64];
        const int comp = b < 4 ? 0 : b - 3;
        const uint8_t *pl = comp == 0 ? py : (comp == 1 ? pu : pv);
        const int ls = comp ? ls_c : ls_y;
        const int vw = comp ? cw : w, vh = comp ? chh : h, r0 = comp ? c0 : y0;
        const int bx = comp ? mx * 8 : mx * 16 + (b & 1) * 8;
        const int by = comp ? my * 8 : my * 16 + (b >> 1) * 8;
        /* get_pixels (dsputil.c:399-416) on the flipped, edge-replicated picture
         * (load_input_picture :857-879 + ff_emulated_edge_mc mpegvideo.c:1416-1470) */
        for (int yy = 0; yy < 8; yy++)
            for (int xx = 0; xx < 8; xx++) {
                int X = bx + xx, Y = by + yy;
                if (X > vw - 1) X = vw - 1;
                if (Y > vh - 1) Y = vh - 1;
                blk[yy * 8 + xx] = pl[(ptrdiff_t)(r0 - Y) * ls + X];
            }
        amvo_fdct_islow(blk);
        int last = amvo_quantize(blk, qmat);
        /* encode_block (mjpegenc.c:379-435) */
        int diff = blk[0] - pred[comp];
        pred[comp] = blk[0];
        put_coef(&bw, comp ? 1 : 0, 0, diff);
        const int ac = comp ? 3 : 2;
        int run = 0;
        for (int k = 1; k <= last; k++) {
            int v = blk[g_zz[k]];
            if (!v) { run++; continue; }
            for (; run >= 16; run -= 16) bw_put(&bw, g_hlen[ac][0xf0], g_hcode[ac][0xf0]);
            put_coef(&bw, ac, run, v);
            run = 0;
        }
        if (last < 63) bw_put(&bw, g_hlen[ac][0], g_hcode[ac][0]);
    }
    /* pad to a byte with ones (ff_mjpeg_encode_stuffing mjpegenc.c:338-343) */
    if (bw.nacc) bw_put(&bw, 8 - bw.nacc, 0xff);

    /* SOI, stuffed segment, EOI */
    size_t o = 0; int fits = 1;
#define EMIT(b) do { if (o < cap) out[o] = (uint8_t)(b); else fits = 0; o++; } while (0)
    EMIT(0xff); EMIT(0xd8);
    for (size_t i = 0; i < bw.nbytes; i++) { EMIT(raw[i]); if (raw[i] == 0xff) EMIT(0x00); }
    EMIT(0xff); EMIT(0xd9);
#undef EMIT
    free(raw);
    return fits ? (int)o : -2;
}

/* ------------------------------------------------------------------ decoder */

typedef struct { const uint8_t *p; size_t n, pos; uint64_t acc; int nacc; int overrun; } bitr;

static inline void br_fill(bitr *r)
{
    while (r->nacc <= 56) {
        uint64_t b = 0;
        if (r->pos < r->n) b = r->p[r->pos]; else if (r->pos >= r->n + 8) r->overrun = 1;
        r->pos++;
        r->acc |= b << (56 - r->nacc);
        r->nacc += 8;
    }
}
static inline uint32_t br_peek(bitr *r, int n) { br_fill(r); return n ? (uint32_t)(r->acc >> (64 - n)) : 0; }
static inline void br_skip(bitr *r, int n) { r->acc <<= n; r->nacc -= n; }
static inline size_t br_bits_used(const bitr *r) { return r->pos * 8 - (size_t)r->nacc; }

/* the tables a scan is decoded with: the fixed AMV/SP5X ones (sp5x.h templates) or what a JPEG's own DQT /
 * DHT segments carry.  Huffman slot t: 0 DC of component 0, 1 DC of components 1/2, 2/3 the AC tables. */
typedef struct {
    uint8_t cnt[4][16];      /* codes per length 1..16 */
    uint8_t sym[4][256];     /* symbols in code order */
    uint8_t qzz[2][64];      /* quantisers in zigzag (DQT) order: component 0, components 1/2 */
    uint8_t hs[3], vs[3];    /* sampling factors of the three components (SOF0); all zero = 2x2, 1x1, 1x1 */
    int restart_interval;    /* DRI (mjpeg_decode_dri mjpegdec.c:858-867); 0 = none */
} scan_tables;

static const scan_tables *fixed_tables(void)
{
    static scan_tables T; static int ready;
    if (!ready) {
        memcpy(T.cnt, kCnt, sizeof(T.cnt));
        memcpy(T.sym[0], kSymDC, 12); memcpy(T.sym[1], kSymDC, 12);
        memcpy(T.sym[2], kSymACL, 162); memcpy(T.sym[3], kSymACC, 162);
        memcpy(T.qzz, kDecQuantZZ, sizeof(T.qzz));
        T.hs[0] = T.vs[0] = 2; T.hs[1] = T.vs[1] = T.hs[2] = T.vs[2] = 1;
        T.restart_interval = 0;
        ready = 1;
    }
    return &T;
}

/* canonical decode: find the shortest length whose code range contains the
 * next bits (what the 9-bit two-level VLC of bitstream.h:813-839 resolves to) */
static int huff_get(bitr *r, const scan_tables *T, int t)
{
    uint32_t v = br_peek(r, 16);
    unsigned first = 0; int k = 0;
    const uint8_t *sym = T->sym[t];
    for (int len = 1; len <= 16; len++) {
        unsigned c = v >> (16 - len);
        int cnt = T->cnt[t][len - 1];
        if (c >= first && c < first + cnt) { br_skip(r, len); return sym[k + (c - first)]; }
        k += cnt;
        first = (first + cnt) << 1;
    }
    return -1;
}

/* get_xbits (bitstream.h:629-639): n-bit two's-complement-ish JPEG extend */
static int get_extend(bitr *r, int n)
{
    if (!n) return 0;
    int v = (int)br_peek(r, n);
    br_skip(r, n);
    return (v >> (n - 1)) ? v : v - ((1 << n) - 1);
}

/* status bits returned by amvo_decode_frame */
enum { AMVO_E_SHORT = 1, AMVO_E_BADCODE = 2, AMVO_E_COEFIDX = 4, AMVO_E_MARKER = 8, AMVO_E_OVERRUN = 16 };

/* Remove byte stuffing exactly like the SOS branch of ff_mjpeg_decode_frame
 * (mjpegdec.c:1137-1160) does on `payload ++ FF D9` (sp5xdec.c:75-88):
 * FF followed by 00 -> FF; FF D0..D7 kept; runs of FF collapse; FF + any other
 * non-zero byte ends the scan data. Returns unstuffed length. */
AMVO_API size_t amvo_unstuff(const uint8_t *pkt, uint32_t size, uint8_t *dst, int *flags)
{
    size_t n = 0, i = 2, end = size >= 4 ? size - 2 : 2;   /* payload = pkt[2 .. size-2) */
    int stop = 0;
    if (size < 4) { if (flags) *flags |= AMVO_E_SHORT; }
    while (i < end && !stop) {
        uint8_t x = pkt[i++];
        dst[n++] = x;
        if (x != 0xff) continue;
        uint8_t z = 0xff;
        while (i < end && z == 0xff) z = pkt[i++];
        if (z == 0xff) { z = 0xd9; }                       /* ran into the appended EOI */
        if (z >= 0xd0 && z <= 0xd7) dst[n++] = z;
        else if (z) { stop = 1; if (z != 0xd9 && flags) *flags |= AMVO_E_MARKER; }
    }
    if (!stop) dst[n++] = 0xff;                            /* the FF of the appended EOI is copied too */
    else if (i < end && flags) *flags |= AMVO_E_MARKER;     /* scan cut short by a marker inside the payload */
    return n;
}

/* the scan decoder shared by AMV and SP5X: `scan` is the un-stuffed entropy-coded segment; flip = AMV's
 * bottom-up placement (mjpegdec.c:672-677), 0 for SP5X */
static int decode_scan(uint8_t *scan, size_t nscan, int flags, int w, int h, int flip,
                       uint8_t *py, uint8_t *pu, uint8_t *pv, int ls_y, int ls_c,
                       int16_t *coef_dump, uint8_t *uy, uint8_t *uu, uint8_t *uv, const scan_tables *T)
{
    build_tables();
    if (!T) T = fixed_tables();
    /* interleaved scan (mjpeg_decode_scan mjpegdec.c:660-736): per MCU, per component, v x h blocks in raster
     * order (:700-722); the MCU covers 8*h_max x 8*v_max pixels (ff_mjpeg_decode_sos :808-811); component c's
     * plane is ceil(w * h_c / h_max) x ceil(h * v_c / v_max) */
    const int hmax = T->hs[0] > T->hs[1] ? T->hs[0] : T->hs[1], vmax = T->vs[0] > T->vs[1] ? T->vs[0] : T->vs[1];
    const int mbw = (w + 8 * hmax - 1) / (8 * hmax), mbh = (h + 8 * vmax - 1) / (8 * vmax);
    int pw[3], ph[3];
    for (int c = 0; c < 3; c++) {
        pw[c] = (w * T->hs[c] + hmax - 1) / hmax;
        ph[c] = (h * T->vs[c] + vmax - 1) / vmax;
    }
    int16_t q[2][64];                                       /* raster order (mjpegdec.c:131-134) */
    for (int t = 0; t < 2; t++) for (int k = 0; k < 64; k++) q[t][g_zz[k]] = T->qzz[t][k];
    bitr br = { scan, nscan, 0, 0, 0, 0 };
    int pred[3] = { 1024, 1024, 1024 };                     /* mjpegdec.c:805-806 */
    const int y0 = flip_start_row(h, 2), c0 = flip_start_row(h, 1);
    int nblk = 0;

    int restart_count = 0;
    for (int my = 0; my < mbh && !(flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)); my++)
    for (int mx = 0; mx < mbw && !(flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)); mx++) {
    if (T->restart_interval && !restart_count) restart_count = T->restart_interval;          /* :682-683 */
    for (int comp = 0; comp < 3 && !(flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)); comp++)
    for (int b = 0; b < T->hs[comp] * T->vs[comp]; b++, nblk++) {
        const int tq = comp ? 1 : 0;
        int16_t blk[64] = { 0 };
        int s = huff_get(&br, T, tq);
        if (s < 0) { flags |= AMVO_E_BADCODE; break; }
        pred[comp] += get_extend(&br, s) * q[tq][0];
        blk[0] = (int16_t)pred[comp];
        for (int k = 0;;) {
            int rs = huff_get(&br, T, 2 + tq);
            if (rs < 0) { flags |= AMVO_E_BADCODE; break; }
            if (rs == 0x00) break;                           /* EOB */
            if (rs == 0xf0) { k += 16; continue; }           /* ZRL: no range check (:400-401) */
            k += (rs >> 4) + 1;
            int lv = get_extend(&br, rs & 15);
            if (k > 63) { flags |= AMVO_E_COEFIDX; break; }
            blk[g_zz[k]] = (int16_t)(lv * q[tq][g_zz[k]]);
            if (k == 63) break;
        }
        if (flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)) break;
        if (coef_dump) memcpy(coef_dump + (size_t)nblk * 64, blk, sizeof(blk));

        uint8_t px[64], ud[64];
        amvo_idct_put_ex(blk, px, ud);
        /* placement (mjpegdec.c:672-677,710-716; bottom-up for AMV), clipped to the picture */
        uint8_t *pl = comp == 0 ? py : (comp == 1 ? pu : pv);
        uint8_t *um = comp == 0 ? uy : (comp == 1 ? uu : uv);
        const int ls = comp ? ls_c : ls_y, vw = pw[comp], vh = ph[comp];
        const int r0 = comp ? c0 : y0;
        const int bx = (T->hs[comp] * mx + b % T->hs[comp]) * 8;
        const int by = (T->vs[comp] * my + b / T->hs[comp]) * 8;
        for (int yy = 0; yy < 8; yy++) {
            int row = flip ? r0 - (by + yy) : by + yy;
            if (row < 0 || row >= vh) continue;
            for (int xx = 0; xx < 8; xx++)
                if (bx + xx < vw) {
                    pl[(ptrdiff_t)row * ls + bx + xx] = px[yy * 8 + xx];
                    if (um) um[(ptrdiff_t)row * ls + bx + xx] = ud[yy * 8 + xx];
                }
        }
    }
    /* restart: byte-align, skip the RSTn marker (un-stuffing keeps FF Dn, :1151-1152), reset the predictors
     * (:726-732, including the reference's "< 1350" condition: larger intervals are never honoured) */
    if (T->restart_interval && T->restart_interval < 1350 && !--restart_count &&
        !(flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)) && !(my == mbh - 1 && mx == mbw - 1) /* nothing follows the last MCU */) {
        br_fill(&br);
        br_skip(&br, (int)((8 - (br_bits_used(&br) & 7)) & 7));
        br_fill(&br);
        br_skip(&br, 16);
        pred[0] = pred[1] = pred[2] = 1024;
    }
    }
    if (br_bits_used(&br) > nscan * 8) flags |= AMVO_E_OVERRUN;
    return flags;
}

/* Decode one packet into three planes (strides in bytes).  Returns 0 or a mask
 * of AMVO_E_*; on error the picture content is unspecified.
 * sp5x_decode_frame (sp5xdec.c:33-93) -> ff_mjpeg_decode_frame
 * (mjpegdec.c:1106-1340) -> mjpeg_decode_scan (:660-736) -> decode_block
 * (:376-430) -> simple_idct_put. */
AMVO_API int amvo_decode_frame_ex(const uint8_t *pkt, uint32_t size, int w, int h,
                                  uint8_t *py, uint8_t *pu, uint8_t *pv, int ls_y, int ls_c,
                                  int16_t *coef_dump /* optional: 64*6*mbw*mbh dequantised coefs */,
                                  uint8_t *uy, uint8_t *uu, uint8_t *uv /* optional: 1 where the reference is undefined */)
{
    int flags = 0;
    uint8_t *scan = (uint8_t *)malloc((size_t)size + 16);
    size_t nscan = amvo_unstuff(pkt, size, scan, &flags);
    flags = decode_scan(scan, nscan, flags, w, h, 1, py, pu, pv, ls_y, ls_c, coef_dump, uy, uu, uv, NULL);
    free(scan);
    return flags;
}

/* SP5X (sp5x_decoder, sp5xdec.c:33-188 with codec_id != CODEC_ID_AMV): the scan is the packet from byte 14
 * to its end; the reference stuffs every FF itself (:78-84) before the MJPEG decoder un-stuffs it again, so
 * the bytes are literal, followed by the FF of the appended EOI (:87-88); no vertical flip. */
AMVO_API int amvo_sp5x_decode_frame(const uint8_t *pkt, uint32_t size, int w, int h,
                                    uint8_t *py, uint8_t *pu, uint8_t *pv, int ls_y, int ls_c,
                                    uint8_t *uy, uint8_t *uu, uint8_t *uv /* optional undefined-domain masks */)
{
    int flags = size < 14 ? AMVO_E_SHORT : 0;
    const size_t npay = size >= 14 ? size - 14 : 0;
    uint8_t *scan = (uint8_t *)malloc(npay + 16);
    if (npay) memcpy(scan, pkt + 14, npay);
    scan[npay] = 0xff;
    flags = decode_scan(scan, npay + 1, flags, w, h, 0, py, pu, pv, ls_y, ls_c, NULL, uy, uu, uv, NULL);
    free(scan);
    return flags;
}

AMVO_API int amvo_sp5x_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                                     int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v, int *status,
                                     uint8_t *uy, uint8_t *uu, uint8_t *uv)
{
    const int cw = (w + 1) >> 1, chh = (h + 1) >> 1;
    for (int i = 0; i < n; i++) {
        size_t yo = (size_t)i * w * h, co = (size_t)i * cw * chh;
        int st = amvo_sp5x_decode_frame(pkts + off[i], size[i], w, h, y + yo, u + co, v + co, w, cw,
                                        uy ? uy + yo : NULL, uu ? uu + co : NULL, uv ? uv + co : NULL);
        if (status) status[i] = st;
    }
    return n;
}

/* ---- plain baseline MJPEG: the `mjpeg_decoder` of the same source (ff_mjpeg_decode_frame mjpegdec.c:1106-1340)
 * on a full JPEG whose tables travel in the stream.  Header walk: find_marker (:1076-1104) + the segment parsers
 * ff_mjpeg_decode_dqt (:113-145, 8-bit tables, zigzag order), ff_mjpeg_decode_dht (:148-192), ff_mjpeg_decode_sof
 * (:194-345, 8 bits, pix_fmt_id 0x221111 = 4:2:0), ff_mjpeg_decode_sos (:738-856; last_dc = 1024 :805-806,
 * 16x16 MCUs :808-811), mjpeg_decode_dri (:858-867).  Supported: SOF0, three components sampled 2x2 / 1x1 / 1x1,
 * components 1 and 2 sharing their quantiser and Huffman tables, no restart interval; anything else is
 * AMVO_E_HEADER.  The scan is un-stuffed and decoded exactly like AMV's (same functions), top-down. */
enum { AMVO_E_HEADER = 128 };
typedef struct { int w, h; uint32_t scan_start; scan_tables T; } mjpeg_header;

static int mjpeg_parse(const uint8_t *p, uint32_t size, mjpeg_header *H)
{
    uint8_t q[4][64]; int have_q[4] = { 0, 0, 0, 0 };
    uint8_t hc[2][4][16], hs[2][4][256]; int have_h[2][4] = { { 0 } };
    int comp_id[3], comp_q[3], have_sof = 0;
    uint32_t i = 0;
    H->T.restart_interval = 0;                                          /* reset at SOI (:1225) */
    if (size < 4 || p[0] != 0xff || p[1] != 0xd8) return -1;
    i = 2;
    while (i + 4 <= size) {
        if (p[i] != 0xff) { i++; continue; }
        const int m = p[i + 1];
        if (m < 0xc0 || m == 0xff) { i++; continue; }                   /* find_marker: FF followed by C0..FE */
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7)) { i += 2; continue; }
        if (m == 0xd9) return -1;                                        /* EOI before any scan */
        const uint32_t len = ((uint32_t)p[i + 2] << 8) | p[i + 3];
        if (len < 2 || i + 2 + len > size) return -1;
        const uint8_t *d = p + i + 4; uint32_t n = len - 2;
        if (m == 0xdb) {                                                 /* DQT */
            while (n >= 65) {
                if (d[0] >> 4) return -1;                                /* 16-bit tables (:121-125) */
                const int id = d[0] & 15; if (id >= 4) return -1;
                memcpy(q[id], d + 1, 64); have_q[id] = 1;
                d += 65; n -= 65;
            }
        } else if (m == 0xc4) {                                          /* DHT */
            while (n > 0) {
                if (n < 17) return -1;
                const int cls = d[0] >> 4, id = d[0] & 15; int tot = 0;
                if (cls >= 2 || id >= 4) return -1;
                for (int k = 0; k < 16; k++) tot += d[1 + k];
                if (tot > 256 || n < 17u + (uint32_t)tot) return -1;
                memcpy(hc[cls][id], d + 1, 16); memset(hs[cls][id], 0, 256); memcpy(hs[cls][id], d + 17, tot);
                have_h[cls][id] = 1;
                d += 17 + tot; n -= 17 + tot;
            }
        } else if (m == 0xc0) {                                          /* SOF0 */
            if (n < 15 || d[0] != 8 || d[5] != 3) return -1;
            H->h = (d[1] << 8) | d[2]; H->w = (d[3] << 8) | d[4];
            for (int c = 0; c < 3; c++) {
                comp_id[c] = d[6 + 3 * c];
                H->T.hs[c] = d[7 + 3 * c] >> 4; H->T.vs[c] = d[7 + 3 * c] & 15;
                comp_q[c] = d[8 + 3 * c]; if (comp_q[c] >= 4) return -1;
            }
            {   /* 4:2:0 (0x221111), 4:2:2 as 0x211111 or the encoder's 0x221212, 4:4:4 (0x111111): ff_mjpeg_decode_sof :283-311 */
                const unsigned id = (d[7] << 16) | (d[10] << 8) | d[13];
                if (id != 0x221111 && id != 0x211111 && id != 0x221212 && id != 0x111111) return -1;
            }
            if (comp_q[1] != comp_q[2]) return -1;
            have_sof = 1;
        } else if (m >= 0xc1 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
            return -1;                                                   /* other SOFn: not baseline */
        } else if (m == 0xdd) {                                          /* DRI */
            if (len != 4) return -1;
            H->T.restart_interval = (d[0] << 8) | d[1];
        } else if (m == 0xda) {                                          /* SOS */
            if (!have_sof || n < 10 || d[0] != 3 || len != 6 + 2 * 3) return -1;
            int td[3], ta[3];
            for (int c = 0; c < 3; c++) {
                if (d[1 + 2 * c] != comp_id[c]) return -1;               /* components in frame order */
                td[c] = d[2 + 2 * c] >> 4; ta[c] = d[2 + 2 * c] & 15;
                if (td[c] >= 4 || ta[c] >= 4 || !have_h[0][td[c]] || !have_h[1][ta[c]]) return -1;
            }
            if (td[1] != td[2] || ta[1] != ta[2]) return -1;
            if (d[7] != 0 || d[8] != 63 || d[9] != 0) return -1;         /* Ss, Se, Ah/Al of a sequential scan */
            if (!have_q[comp_q[0]] || !have_q[comp_q[1]]) return -1;
            for (int c = 0; c < 2; c++) {
                memcpy(H->T.cnt[c], hc[0][td[c]], 16);     memcpy(H->T.sym[c], hs[0][td[c]], 256);
                memcpy(H->T.cnt[2 + c], hc[1][ta[c]], 16); memcpy(H->T.sym[2 + c], hs[1][ta[c]], 256);
                memcpy(H->T.qzz[c], q[comp_q[c]], 64);
            }
            H->scan_start = i + 2 + len;
            return 0;
        }
        i += 2 + len;
    }
    return -1;
}

static void mjpeg_chroma_dims(const mjpeg_header *H, int *cw, int *ch)
{
    const int hmax = H->T.hs[0], vmax = H->T.vs[0];
    *cw = (H->w * H->T.hs[1] + hmax - 1) / hmax;
    *ch = (H->h * H->T.vs[1] + vmax - 1) / vmax;
}

/* -> picture size, where the scan starts, and the chroma plane size the sampling gives (any pointer may be NULL) */
AMVO_API int amvo_mjpeg_header(const uint8_t *pkt, uint32_t size, int *w, int *h, uint32_t *scan_start, int *cw, int *ch)
{
    mjpeg_header H;
    int a, b;
    if (mjpeg_parse(pkt, size, &H)) return -1;
    mjpeg_chroma_dims(&H, &a, &b);
    if (w) *w = H.w;
    if (h) *h = H.h;
    if (scan_start) *scan_start = H.scan_start;
    if (cw) *cw = a;
    if (ch) *ch = b;
    return 0;
}

AMVO_API int amvo_mjpeg_decode_frame(const uint8_t *pkt, uint32_t size, int w, int h,
                                     uint8_t *py, uint8_t *pu, uint8_t *pv, int ls_y, int ls_c,
                                     uint8_t *uy, uint8_t *uu, uint8_t *uv /* optional undefined-domain masks */)
{
    mjpeg_header H;
    if (mjpeg_parse(pkt, size, &H) || H.w != w || H.h != h) return AMVO_E_HEADER;
    /* the scan runs from the end of the SOS header to the next marker (:1137-1160); amvo_unstuff works on
     * pkt[2 .. size-2) ++ FF D9, so hand it the packet from two bytes before the scan */
    int flags = 0;
    uint8_t *scan = (uint8_t *)malloc((size_t)size + 16);
    size_t nscan = amvo_unstuff(pkt + H.scan_start - 2, size - (H.scan_start - 2), scan, &flags);
    flags = decode_scan(scan, nscan, flags, w, h, 0, py, pu, pv, ls_y, ls_c, NULL, uy, uu, uv, &H.T);
    free(scan);
    return flags;
}

/* cw x chh: the chroma plane size of the frames' sampling (amvo_mjpeg_header) */
AMVO_API int amvo_mjpeg_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                                      int n, int w, int h, int cw, int chh, uint8_t *y, uint8_t *u, uint8_t *v, int *status,
                                      uint8_t *uy, uint8_t *uu, uint8_t *uv)
{
    for (int i = 0; i < n; i++) {
        size_t yo = (size_t)i * w * h, co = (size_t)i * cw * chh;
        int st = amvo_mjpeg_decode_frame(pkts + off[i], size[i], w, h, y + yo, u + co, v + co, w, cw,
                                         uy ? uy + yo : NULL, uu ? uu + co : NULL, uv ? uv + co : NULL);
        if (status) status[i] = st;
    }
    return n;
}

AMVO_API int amvo_decode_frame(const uint8_t *pkt, uint32_t size, int w, int h,
                               uint8_t *py, uint8_t *pu, uint8_t *pv, int ls_y, int ls_c,
                               int16_t *coef_dump)
{
    return amvo_decode_frame_ex(pkt, size, w, h, py, pu, pv, ls_y, ls_c, coef_dump, NULL, NULL, NULL);
}

/* -------------------------------------------------------------------- ADPCM */

/* adpcm_decode_frame, CODEC_ID_ADPCM_IMA_AMV (adpcm.c:1268-1292) with
 * adpcm_ima_expand_nibble(...,3) (:716-742).  Returns samples written
 * (2*(size-8)), or -1 if the chunk is shorter than its header / the header's
 * step index is outside 0..88 (the reference indexes out of bounds there). */
AMVO_API int amvo_adpcm_decode_chunk(const uint8_t *c, uint32_t size, int16_t *pcm)
{
    if (size < 8) return -1;
    int pred = (int16_t)(c[0] | (c[1] << 8));
    int idx = (int16_t)(c[2] | (c[3] << 8));
    if (idx < 0 || idx > 88) return -1;
    int n = 0;
    for (uint32_t i = 8; i < size; i++)
        for (int half = 0; half < 2; half++) {
            int nib = half ? (c[i] & 15) : (c[i] >> 4);       /* high nibble first (:1281-1282) */
            int step = kImaStep[idx];
            idx = clampi(idx + kImaIdxAdj[nib & 7], 0, 88);
            int diff = ((2 * (nib & 7) + 1) * step) >> 3;
            pred = clampi(nib & 8 ? pred - diff : pred + diff, -32768, 32767);
            pcm[n++] = (int16_t)pred;
        }
    return n;
}

/* adpcm_encode_frame, CODEC_ID_ADPCM_IMA_AMV (adpcm.c:461-496, non-trellis) for
 * one chunk of `nsamples` (even) samples with the step index carried in.
 * adpcm_ima_compress_sample (:219-227).  Returns bytes written (8+nsamples/2). */
AMVO_API int amvo_adpcm_encode_chunk(const int16_t *pcm, uint32_t nsamples, int step_index_in,
                                     uint8_t *out, int *step_index_out)
{
    if ((nsamples & 1) || step_index_in < 0 || step_index_in > 88) return -1;
    int prev = nsamples ? pcm[0] : 0, idx = step_index_in;
    if (!nsamples) prev = 0;
    out[0] = (uint8_t)prev; out[1] = (uint8_t)(prev >> 8);
    out[2] = (uint8_t)idx;  out[3] = (uint8_t)(idx >> 8);
    out[4] = (uint8_t)nsamples; out[5] = (uint8_t)(nsamples >> 8);
    out[6] = (uint8_t)(nsamples >> 16); out[7] = (uint8_t)(nsamples >> 24);
    for (uint32_t i = 0; i < nsamples; i++) {
        int step = kImaStep[idx];
        int delta = pcm[i] - prev;
        int mag = delta < 0 ? -delta : delta;
        int q = mag * 4 / step;
        if (q > 7) q = 7;
        int nib = q + (delta < 0 ? 8 : 0);
        int mv = (step * (2 * q + 1)) / 8;                    /* yamaha_difflookup (:124-127), C division */
        prev = clampi(delta < 0 ? prev - mv : prev + mv, -32768, 32767);
        idx = clampi(idx + kImaIdxAdj[q], 0, 88);
        if (i & 1) out[8 + (i >> 1)] |= (uint8_t)nib;
        else       out[8 + (i >> 1)]  = (uint8_t)(nib << 4);
    }
    if (step_index_out) *step_index_out = idx;
    return 8 + (int)(nsamples >> 1);
}

/* The `-trellis N` path of the same encoder: adpcm_compress_trellis (adpcm.c:287-443, IMA branch :383-395 with
 * STORE_NODE :340-377) -- a beam search over the decoder states with a frontier of 2^N nodes kept sorted by
 * accumulated squared error, states with the same decoded sample collapsed, the best path frozen every 128 samples.
 * Restated with explicit arrays: `ord` is the sorted frontier (indices into the node pool of the current step), the
 * pool of a step has `frontier` slots and evicted nodes are recycled together with their path slot, exactly like
 * the reference recycles `nodes_next[frontier-1]`.  trellis 1..5. */
typedef struct { uint32_t ssd; int path, sample1, step; } tnode;
AMVO_API int amvo_adpcm_encode_chunk_trellis(const int16_t *pcm, uint32_t nsamples, int step_index_in, int trellis,
                                             uint8_t *out, int *step_index_out)
{
    static const int8_t difflookup[16] = { 1, 3, 5, 7, 9, 11, 13, 15, -1, -3, -5, -7, -9, -11, -13, -15 };
    if ((nsamples & 1) || step_index_in < 0 || step_index_in > 88 || trellis < 1 || trellis > 5) return -1;
    enum { FREEZE = 128, FMAX = 32 };
    const int frontier = 1 << trellis;
    const int prev0 = nsamples ? pcm[0] : 0;
    out[0] = (uint8_t)prev0; out[1] = (uint8_t)(prev0 >> 8);
    out[2] = (uint8_t)step_index_in; out[3] = (uint8_t)(step_index_in >> 8);
    out[4] = (uint8_t)nsamples; out[5] = (uint8_t)(nsamples >> 8); out[6] = (uint8_t)(nsamples >> 16); out[7] = (uint8_t)(nsamples >> 24);
    uint8_t *nib = (uint8_t *)malloc(nsamples + 1);
    static tnode pool[2][FMAX];
    static struct { uint8_t nibble; int prev; } paths[FMAX * FREEZE];
    int cur[FMAX], nxt[FMAX], ncur = 1, pathn = 0, froze = -1;         /* sorted frontiers: slots of pool[(i-1)&1] / pool[i&1] */
    pool[1][0].ssd = 0; pool[1][0].path = 0; pool[1][0].step = step_index_in; pool[1][0].sample1 = prev0;
    cur[0] = 0;
    for (int i = 0; i < (int)nsamples; i++) {
        tnode *from = pool[(i & 1) ^ 1], *to = pool[i & 1];
        const int sample = pcm[i];
        int nnxt = 0, nalloc = 0;
        for (int j = 0; j < ncur; j++) {
            const tnode *p = &from[cur[j]];
            const int range = j < frontier / 2 ? 1 : 0;
            const int st = kImaStep[p->step];
            const int div = (sample - p->sample1) * 4 / st;
            int nmin = clampi(div - range, -7, 6), nmax = clampi(div + range, -6, 7);
            if (nmin <= 0) nmin--;                                   /* distinguish -0 from +0 */
            if (nmax < 0) nmax--;
            for (int nidx = nmin; nidx <= nmax; nidx++) {
                const int nibble = nidx < 0 ? 7 - nidx : nidx;
                const int dec = clampi(p->sample1 + (st * difflookup[nibble]) / 8, -32768, 32767);
                const int d = sample - dec;
                const uint32_t ssd = p->ssd + (uint32_t)d * (uint32_t)d;
                if (nnxt == frontier && ssd >= to[nxt[frontier - 1]].ssd) continue;
                int k, dup = 0;
                for (k = 0; k < nnxt; k++) if (dec == to[nxt[k]].sample1) { dup = 1; break; }
                if (dup) continue;
                for (k = 0; k < frontier; k++) {
                    if (k >= nnxt || ssd < to[nxt[k]].ssd) {
                        int slot;
                        if (nnxt == frontier) slot = nxt[frontier - 1];              /* recycle the evicted node and its path slot */
                        else { slot = nalloc++; to[slot].path = pathn++; }
                        to[slot].ssd = ssd;
                        to[slot].step = clampi(p->step + kImaIdxAdj[nibble & 7], 0, 88);
                        to[slot].sample1 = dec;
                        paths[to[slot].path].nibble = (uint8_t)nibble;
                        paths[to[slot].path].prev = p->path;
                        const int last = nnxt == frontier ? frontier - 1 : nnxt;    /* memmove(&next[k+1], &next[k], ...) */
                        for (int m = last; m > k; m--) nxt[m] = nxt[m - 1];
                        nxt[k] = slot;
                        if (nnxt < frontier) nnxt++;
                        break;
                    }
                }
            }
        }
        memcpy(cur, nxt, sizeof(cur)); ncur = nnxt;
        tnode *now = to;
        if (now[cur[0]].ssd > (1u << 28)) {                               /* prevent overflow */
            for (int j = 1; j < ncur; j++) now[cur[j]].ssd -= now[cur[0]].ssd;
            now[cur[0]].ssd = 0;
        }
        if (i == froze + FREEZE) {                                        /* merge old paths to save memory */
            int pp = now[cur[0]].path;
            for (int k = i; k > froze; k--) { nib[k] = paths[pp].nibble; pp = paths[pp].prev; }
            froze = i; pathn = 0;
            ncur = 1;                                                     /* "just kill them all" */
        }
    }
    const tnode *best = &pool[(nsamples - 1) & 1][cur[0]];
    if (nsamples) {
        int pp = best->path;
        for (int i = (int)nsamples - 1; i > froze; i--) { nib[i] = paths[pp].nibble; pp = paths[pp].prev; }
    }
    for (uint32_t i = 0; i < nsamples / 2; i++) out[8 + i] = (uint8_t)((nib[2 * i] << 4) | nib[2 * i + 1]);
    if (step_index_out) *step_index_out = nsamples ? best->step : step_index_in;
    free(nib);
    return 8 + (int)(nsamples >> 1);
}

/* How many samples (2n) the reference encoder puts in the next chunk
 * (adpcm.c:468-477): host-side bookkeeping of the AVCodec shim. */
AMVO_API uint32_t amvo_adpcm_next_chunk_samples(int frame_size, int sample_rate,
                                                uint64_t samples_written, int *extra_carry)
{
    int n = frame_size >> 1;
    *extra_carry += frame_size & 1;
    n += *extra_carry >> 1;
    *extra_carry &= 1;
    int i = (int)((samples_written + 2 * (uint64_t)n) % (uint64_t)sample_rate);
    if (i && i + frame_size > sample_rate) n += (sample_rate - i) >> 1;
    return (uint32_t)n * 2;
}

/* ------------------------------------------------------------ batch drivers */

AMVO_API int amvo_encode_frames(const uint8_t *y, const uint8_t *u, const uint8_t *v,
                                int n, int w, int h, int qscale,
                                uint8_t *out, uint64_t *off, uint32_t *size, uint64_t cap)
{
    const int cw = (w + 1) >> 1, chh = (h + 1) >> 1;
    uint64_t pos = 0;
    for (int i = 0; i < n; i++) {
        uint64_t room = cap - pos;
        int sz = amvo_encode_frame(y + (size_t)i * w * h, u + (size_t)i * cw * chh, v + (size_t)i * cw * chh,
                                   w, cw, w, h, qscale, out + pos, room > 0xffffffffu ? 0xffffffffu : (uint32_t)room);
        if (sz < 0) return sz;
        off[i] = pos; size[i] = (uint32_t)sz; pos += (uint32_t)sz;
    }
    return n;
}

AMVO_API int amvo_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                                int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v, int *status,
                                uint8_t *uy, uint8_t *uu, uint8_t *uv)
{
    const int cw = (w + 1) >> 1, chh = (h + 1) >> 1;
    for (int i = 0; i < n; i++) {
        size_t yo = (size_t)i * w * h, co = (size_t)i * cw * chh;
        int st = amvo_decode_frame_ex(pkts + off[i], size[i], w, h, y + yo, u + co, v + co, w, cw, NULL,
                                      uy ? uy + yo : NULL, uu ? uu + co : NULL, uv ? uv + co : NULL);
        if (status) status[i] = st;
    }
    return n;
}

AMVO_API int amvo_adpcm_decode_chunks(const uint8_t *chunks, const uint64_t *off, const uint32_t *size,
                                      int n, int16_t *pcm, const uint64_t *pcm_off, int *status)
{
    for (int i = 0; i < n; i++) {
        int r = amvo_adpcm_decode_chunk(chunks + off[i], size[i], pcm + pcm_off[i]);
        if (status) status[i] = r < 0 ? r : 0;
    }
    return n;
}

AMVO_API int amvo_adpcm_encode_chunks(const int16_t *pcm, const uint64_t *pcm_off, const uint32_t *nsamples,
                                      const int16_t *step_in, int16_t *step_out, int n,
                                      uint8_t *out, const uint64_t *out_off)
{
    for (int i = 0; i < n; i++) {
        int so = 0;
        int r = amvo_adpcm_encode_chunk(pcm + pcm_off[i], nsamples[i], step_in ? step_in[i] : 0,
                                        out + out_off[i], &so);
        if (r < 0) return r;
        if (step_out) step_out[i] = (int16_t)so;
    }
    return n;
}

/* ========================================================================== *
 * amvlib flavour: C-AMVDecoder/amvlib decodes the same packets with
 * different arithmetic -- its own quantiser tables, a zigzag table with a typo, a
 * Chen-Wang IDCT on 32-bit ints and a fixed-point YUV->BGR24 bottom-up store.
 * Paths below are relative to C-AMVDecoder/amvlib/.
 * PARITY PINNED by tests/test_oracle_vs_ref.py against oracle/_ref/libamvlibref.so
 * (the unmodified amvlib compiled in place) on the fixture bin/AMV1.amv and on
 * reference-encoded synthetic frames; vectors in tests/golden/amvlib_golden.npz.
 * ========================================================================== */

/* amv_luminance_quant_tbl / amv_chrominance_quant_tbl (AmvJpeg.c:30-39,52-61), indexed by
 * zigzag position (IQtIZzBlock uses pQt[tag], tag = zigzag index, :1041-1046) */
static const uint8_t kAmvlibQuant[2][64] = {
  { 0x08, 0x06, 0x06, 0x07, 0x06, 0x05, 0x08, 0x07, 0x07, 0x07, 0x09, 0x09, 0x08, 0x0A, 0x0C, 0x14,
    0x0D, 0x0C, 0x0B, 0x0B, 0x0C, 0x19, 0x12, 0x13, 0x0F, 0x14, 0x1D, 0x1A, 0x1F, 0x1E, 0x1D, 0x1A,
    0x1C, 0x1C, 0x20, 0x24, 0x2E, 0x27, 0x20, 0x22, 0x2C, 0x27, 0x1C, 0x1C, 0x28, 0x37, 0x29, 0x2C,
    0x30, 0x31, 0x34, 0x34, 0x34, 0x1F, 0x27, 0x39, 0x3D, 0x38, 0x32, 0x3C, 0x2E, 0x33, 0x34, 0x32 },
  { 0x09, 0x09, 0x09, 0x0C, 0x0B, 0x0C, 0x18, 0x0D, 0x0D, 0x18, 0x32, 0x21, 0x1C, 0x21, 0x32, 0x32,
    0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32,
    0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32,
    0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32, 0x32 },
};

/* Raster position -> zigzag index as amvlib has it (AmvJpeg.c:131-141): the standard scan except
 * that raster (3,4) reads index 37 instead of 31 -- coefficient 31 is never used and coefficient 37
 * lands in two places. */
AMVO_API void amvo_amvlib_zigzag(uint8_t raster_to_zz[64])
{
    build_tables();
    for (int k = 0; k < 64; k++) raster_to_zz[g_zz[k]] = (uint8_t)k;
    raster_to_zz[3 * 8 + 4] = 37;
}

static inline int32_t wr32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }     /* 32-bit wrap */
static inline int32_t shl32(int32_t v, int s) { return (int32_t)((uint32_t)v << s); }
/* Initialize_Fast_IDCT :1069-1076: the clamp table covers -512..511; beyond that the reference reads
 * past it (undefined), we saturate and report it through *ud */
static inline int iclp_u(int32_t v, uint8_t *ud) { if (ud && (v < -512 || v > 511)) *ud = 1; return v < -256 ? -256 : (v > 255 ? 255 : v); }

#define AW1 2841
#define AW2 2676
#define AW3 2408
#define AW5 1609
#define AW6 1108
#define AW7 565

/* idctrow (AmvJpeg.c:1078-1125) */
static void amvlib_idct_row(int32_t *b)
{
    int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8;
    x1 = shl32(b[4], 11); x2 = b[6]; x3 = b[2]; x4 = b[1]; x5 = b[7]; x6 = b[5]; x7 = b[3];
    if (!(x1 | x2 | x3 | x4 | x5 | x6 | x7)) { int32_t v = shl32(b[0], 3); for (int i = 0; i < 8; i++) b[i] = v; return; }
    x0 = wr32((int64_t)shl32(b[0], 11) + 128);
    x8 = wr32((int64_t)AW7 * wr32((int64_t)x4 + x5));
    x4 = wr32(x8 + (int64_t)(AW1 - AW7) * x4);
    x5 = wr32(x8 - (int64_t)(AW1 + AW7) * x5);
    x8 = wr32((int64_t)AW3 * wr32((int64_t)x6 + x7));
    x6 = wr32(x8 - (int64_t)(AW3 - AW5) * x6);
    x7 = wr32(x8 - (int64_t)(AW3 + AW5) * x7);
    x8 = wr32((int64_t)x0 + x1); x0 = wr32((int64_t)x0 - x1);
    x1 = wr32((int64_t)AW6 * wr32((int64_t)x3 + x2));
    x2 = wr32(x1 - (int64_t)(AW2 + AW6) * x2);
    x3 = wr32(x1 + (int64_t)(AW2 - AW6) * x3);
    x1 = wr32((int64_t)x4 + x6); x4 = wr32((int64_t)x4 - x6);
    x6 = wr32((int64_t)x5 + x7); x5 = wr32((int64_t)x5 - x7);
    x7 = wr32((int64_t)x8 + x3); x8 = wr32((int64_t)x8 - x3);
    x3 = wr32((int64_t)x0 + x2); x0 = wr32((int64_t)x0 - x2);
    x2 = wr32((int64_t)181 * wr32((int64_t)x4 + x5) + 128) >> 8;
    x4 = wr32((int64_t)181 * wr32((int64_t)x4 - x5) + 128) >> 8;
    b[0] = wr32((int64_t)x7 + x1) >> 8; b[1] = wr32((int64_t)x3 + x2) >> 8;
    b[2] = wr32((int64_t)x0 + x4) >> 8; b[3] = wr32((int64_t)x8 + x6) >> 8;
    b[4] = wr32((int64_t)x8 - x6) >> 8; b[5] = wr32((int64_t)x0 - x4) >> 8;
    b[6] = wr32((int64_t)x3 - x2) >> 8; b[7] = wr32((int64_t)x7 - x1) >> 8;
}

/* idctcol (AmvJpeg.c:1127-1175) */
static void amvlib_idct_col(int32_t *b, uint8_t *ud /* 8 flags with stride 8, or NULL */)
{
#define iclp(v, i) iclp_u((v), ud ? ud + 8 * (i) : NULL)
    int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8;
    x1 = shl32(b[8 * 4], 8); x2 = b[8 * 6]; x3 = b[8 * 2]; x4 = b[8 * 1]; x5 = b[8 * 7]; x6 = b[8 * 5]; x7 = b[8 * 3];
    if (!(x1 | x2 | x3 | x4 | x5 | x6 | x7)) {
        const int32_t t = wr32((int64_t)b[0] + 32) >> 6;
        for (int i = 0; i < 8; i++) b[8 * i] = iclp(t, i);
        return;
    }
    x0 = wr32((int64_t)shl32(b[0], 8) + 8192);
    x8 = wr32((int64_t)AW7 * wr32((int64_t)x4 + x5) + 4);
    x4 = wr32(x8 + (int64_t)(AW1 - AW7) * x4) >> 3;
    x5 = wr32(x8 - (int64_t)(AW1 + AW7) * x5) >> 3;
    x8 = wr32((int64_t)AW3 * wr32((int64_t)x6 + x7) + 4);
    x6 = wr32(x8 - (int64_t)(AW3 - AW5) * x6) >> 3;
    x7 = wr32(x8 - (int64_t)(AW3 + AW5) * x7) >> 3;
    x8 = wr32((int64_t)x0 + x1); x0 = wr32((int64_t)x0 - x1);
    x1 = wr32((int64_t)AW6 * wr32((int64_t)x3 + x2) + 4);
    x2 = wr32(x1 - (int64_t)(AW2 + AW6) * x2) >> 3;
    x3 = wr32(x1 + (int64_t)(AW2 - AW6) * x3) >> 3;
    x1 = wr32((int64_t)x4 + x6); x4 = wr32((int64_t)x4 - x6);
    x6 = wr32((int64_t)x5 + x7); x5 = wr32((int64_t)x5 - x7);
    x7 = wr32((int64_t)x8 + x3); x8 = wr32((int64_t)x8 - x3);
    x3 = wr32((int64_t)x0 + x2); x0 = wr32((int64_t)x0 - x2);
    x2 = wr32((int64_t)181 * wr32((int64_t)x4 + x5) + 128) >> 8;
    x4 = wr32((int64_t)181 * wr32((int64_t)x4 - x5) + 128) >> 8;
    b[8 * 0] = iclp(wr32((int64_t)x7 + x1) >> 14, 0); b[8 * 1] = iclp(wr32((int64_t)x3 + x2) >> 14, 1);
    b[8 * 2] = iclp(wr32((int64_t)x0 + x4) >> 14, 2); b[8 * 3] = iclp(wr32((int64_t)x8 + x6) >> 14, 3);
    b[8 * 4] = iclp(wr32((int64_t)x8 - x6) >> 14, 4); b[8 * 5] = iclp(wr32((int64_t)x0 - x4) >> 14, 5);
    b[8 * 6] = iclp(wr32((int64_t)x3 - x2) >> 14, 6); b[8 * 7] = iclp(wr32((int64_t)x7 - x1) >> 14, 7);
#undef iclp
}

/* Fast_IDCT (AmvJpeg.c:1050-1059) on dequantised raster coefficients, in place; values -256..255 */
AMVO_API void amvo_amvlib_idct_ex(int32_t blk[64], uint8_t ud[64] /* or NULL: 1 where the reference is undefined */)
{
    if (ud) memset(ud, 0, 64);
    for (int i = 0; i < 8; i++) amvlib_idct_row(blk + 8 * i);
    for (int i = 0; i < 8; i++) amvlib_idct_col(blk + i, ud ? ud + i : NULL);
}
AMVO_API void amvo_amvlib_idct(int32_t blk[64]) { amvo_amvlib_idct_ex(blk, NULL); }

static inline uint8_t clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* StoreBuffer's colour conversion (AmvJpeg.c:808-810): y already carries its +128, u/v are centred */
AMVO_API void amvo_amvlib_yuv_to_bgr(int y, int u, int v, uint8_t bgr[3])
{
    bgr[2] = clip_u8(((y << 8) + 18 * u + 367 * v) >> 8);
    bgr[1] = clip_u8(((y << 8) - 159 * u - 220 * v) >> 8);
    bgr[0] = clip_u8(((y << 8) + 411 * u - 29 * v) >> 8);
}

/* AmvVideoDecode -> AmvJpegDecode (AMVDec.c:259-286, AmvJpeg.c:1515-1539): one packet -> one
 * bottom-up BGR24 bitmap with rows of line_bytes bytes (the reference: WIDTHBYTES(w*24)).
 * Entropy layer: ReadByte (:1061-1069) drops the byte after every FF, DecodeElement (:842-936) is a
 * canonical Huffman decode over the K.3 tables + JPEG sign extension, HufBlock (:938-976) fills 64
 * coefficients in zigzag order, DecodeMCUBlock (:1177-1242) chains the DC per component from 0 in
 * 16-bit arithmetic.  Pixels outside w x h are not stored; bytes of the bitmap that no pixel covers
 * are left as the caller provided them (the reference memsets its buffer to 0 first).
 * Returns 0 or a mask of AMVO_E_*; coef_dump (optional) receives the dequantised raster blocks. */
AMVO_API int amvo_amvlib_decode_frame_ex(const uint8_t *pkt, uint32_t size, int w, int h,
                                         uint8_t *bgr, int line_bytes, int32_t *coef_dump,
                                         uint8_t *undef /* optional, laid out like bgr: 1 where the reference is undefined */)
{
    build_tables();
    int flags = 0;
    if (size < 4) flags |= AMVO_E_SHORT;
    uint8_t r2z[64];
    amvo_amvlib_zigzag(r2z);
    /* the byte reader: everything after SOI, dropping the byte that follows an FF */
    uint8_t *scan = (uint8_t *)malloc((size_t)size + 16);
    size_t nscan = 0;
    for (size_t i = 2; i < size; i++) { scan[nscan++] = pkt[i]; if (pkt[i] == 0xff) i++; }
    bitr br = { scan, nscan, 0, 0, 0, 0 };
    const scan_tables *T = fixed_tables();
    int16_t pred[3] = { 0, 0, 0 };
    const int mbw = (w + 15) / 16, mbh = (h + 15) / 16;
    int nblk = 0;
    for (int my = 0; my < mbh && !(flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)); my++)
    for (int mx = 0; mx < mbw && !(flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)); mx++) {
        int32_t px[6][64];
        uint8_t ud[6][64];
        for (int b = 0; b < 6; b++, nblk++) {
            const int comp = b < 4 ? 0 : b - 3, tq = comp ? 1 : 0;
            int16_t zz[64] = { 0 };
            int s = huff_get(&br, T, tq);
            if (s < 0) { flags |= AMVO_E_BADCODE; break; }
            pred[comp] = (int16_t)(pred[comp] + (int16_t)get_extend(&br, s));
            zz[0] = pred[comp];
            for (int k = 1; k < 64;) {
                int rs = huff_get(&br, T, 2 + tq);
                if (rs < 0) { flags |= AMVO_E_BADCODE; break; }
                if (rs == 0x00) break;
                k += rs >> 4;                                   /* run of zeros, then one value (0 for ZRL) */
                int lv = get_extend(&br, rs & 15);
                if (k > 63) { flags |= AMVO_E_COEFIDX; break; }  /* the reference writes past BlockBuffer here */
                zz[k++] = (int16_t)lv;
            }
            if (flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)) break;
            for (int i = 0; i < 64; i++) px[b][i] = (int32_t)zz[r2z[i]] * kAmvlibQuant[tq][r2z[i]];
            if (coef_dump) memcpy(coef_dump + (size_t)nblk * 64, px[b], sizeof(px[b]));
            amvo_amvlib_idct_ex(px[b], ud[b]);
        }
        if (flags & (AMVO_E_BADCODE | AMVO_E_COEFIDX)) break;
        for (int i = 0; i < 16; i++) {
            const int Y = my * 16 + i;
            if (Y >= h) break;
            uint8_t *row = bgr + (size_t)(h - 1 - Y) * line_bytes;
            for (int j = 0; j < 16; j++) {
                const int X = mx * 16 + j;
                if (X >= w) break;
                const int yv = px[(i >> 3) * 2 + (j >> 3)][(i & 7) * 8 + (j & 7)] + 128;
                const int ci = (i >> 1) * 8 + (j >> 1);
                amvo_amvlib_yuv_to_bgr(yv, px[4][ci], px[5][ci], row + 3 * X);
                if (undef) memset(undef + (size_t)(h - 1 - Y) * line_bytes + 3 * X,
                                  ud[(i >> 3) * 2 + (j >> 3)][(i & 7) * 8 + (j & 7)] | ud[4][ci] | ud[5][ci], 3);
            }
        }
    }
    if (br_bits_used(&br) > nscan * 8) flags |= AMVO_E_OVERRUN;
    free(scan);
    return flags;
}

AMVO_API int amvo_amvlib_decode_frame(const uint8_t *pkt, uint32_t size, int w, int h,
                                      uint8_t *bgr, int line_bytes, int32_t *coef_dump)
{
    return amvo_amvlib_decode_frame_ex(pkt, size, w, h, bgr, line_bytes, coef_dump, NULL);
}

AMVO_API int amvo_amvlib_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size, int n,
                                       int w, int h, uint8_t *bgr, int line_bytes, uint64_t frame_stride, int *status,
                                       uint8_t *undef)
{
    for (int i = 0; i < n; i++) {
        int st = amvo_amvlib_decode_frame_ex(pkts + off[i], size[i], w, h, bgr + (size_t)i * frame_stride, line_bytes, NULL,
                                             undef ? undef + (size_t)i * frame_stride : NULL);
        if (status) status[i] = st;
    }
    return n;
}

/* AmvAudioDecode (AMVDec.c:288-340) + AdpcmImaDecodeFrame (AdpcmIma.c:206-242): predictor = le16,
 * step index = ONE byte (chunk[2]), data decoded in groups of four bytes, high nibble first -- so a
 * chunk whose data length is not a multiple of 4 yields up to 6 samples from bytes past its end
 * (taken as zero here, which is what the harness feeds the reference).  Returns samples written
 * (8 * ceil((size-8)/4)) or -1. */
AMVO_API int amvo_amvlib_audio_decode_chunk(const uint8_t *c, uint32_t size, int16_t *pcm)
{
    if (size <= 8) return -1;
    int pred = (int16_t)(c[0] | (c[1] << 8)), idx = c[2];
    if (idx > 88) return -1;                                   /* the reference indexes past step_table */
    const uint32_t nd = size - 8, ng = (nd + 3) / 4;
    int n = 0;
    for (uint32_t i = 0; i < ng * 4; i++) {
        const int byte = i < nd ? c[8 + i] : 0;
        for (int half = 0; half < 2; half++) {
            int nib = half ? (byte & 15) : (byte >> 4);
            int step = kImaStep[idx];
            idx = clampi(idx + kImaIdxAdj[nib & 7], 0, 88);
            int diff = ((2 * (nib & 7) + 1) * step) >> 3;
            pred = clampi(nib & 8 ? pred - diff : pred + diff, -32768, 32767);
            pcm[n++] = (int16_t)pred;
        }
    }
    return n;
}

/* ========================================================================== *
 * Range conversion next to the codec: what img_convert does between yuv420p (CCIR 601
 * range) and yuvj420p (full range): img_apply_table (imgconvert.c:1236-1260, 2492-2510) with the
 * tables of img_convert_init (:1221-1233) = colorspace.h:69-84, SCALEBITS 10, FIX(x) = (int)(x*1024+0.5).
 * dir 0: CCIR -> JPEG, dir 1: JPEG -> CCIR.  In place allowed.
 * ========================================================================== */
static inline uint8_t range_y(int y, int dir)
{
    if (dir) return (uint8_t)((y * 879 + (512 + (16 << 10))) >> 10);                       /* Y_JPEG_TO_CCIR */
    return clip_u8((y * 1192 + (512 - 16 * 1192)) >> 10);                                   /* Y_CCIR_TO_JPEG, clamped by cm[] */
}
static inline uint8_t range_c(int c, int dir)
{
    if (dir) { int v = ((c - 128) * 903 + (512 + (128 << 10))) >> 10; return (uint8_t)(v < 16 ? 16 : v); }   /* C_JPEG_TO_CCIR */
    return clip_u8(((c - 128) * 1161 + (512 + (128 << 10))) >> 10);                         /* C_CCIR_TO_JPEG */
}
AMVO_API void amvo_convert_range(const uint8_t *y, const uint8_t *u, const uint8_t *v, size_t ny, size_t nc, int dir,
                                 uint8_t *oy, uint8_t *ou, uint8_t *ov)
{
    for (size_t i = 0; i < ny; i++) oy[i] = range_y(y[i], dir);
    for (size_t i = 0; i < nc; i++) { ou[i] = range_c(u[i], dir); ov[i] = range_c(v[i], dir); }
}

/* ========================================================================== *
 * Pre stages next to the codec (second part): the picture scaler and the audio resampler
 * ffmpeg.c runs in front of the encoders.  Both stand on one polyphase bank builder.
 * ========================================================================== */

/* av_build_filter (resample2.c:93-141).  type 0: the cubic with first derivative -0.5 the scaler asks for
 * (:108-113); type >= 2: Kaiser-windowed sinc, beta = type (:118-121, bessel :77-87).  Every phase is
 * normalised to `scale` and rounded through lrintf, i.e. after a conversion to float (:131).  The floating
 * point expressions keep the reference's operand order: the coefficients must come out bit-identical. */
static double bessel_i0(double x)
{
    double v = 1, t = 1;
    x = x * x / 4;
    for (int i = 1; i < 50; i++) { t *= x / (i * i); v += t; }
    return v;
}
static void build_bank(int16_t *bank, double factor, int taps, int phases, int scale, int type)
{
    const int center = (taps - 1) / 2;
    double *tab = (double *)malloc(sizeof(double) * (size_t)taps);
    if (factor > 1.0) factor = 1.0;
    for (int ph = 0; ph < phases; ph++) {
        double norm = 0;
        for (int i = 0; i < taps; i++) {
            double x = M_PI * ((double)(i - center) - (double)ph / phases) * factor, y, w;
            if (type == 0) {
                const float d = -0.5;
                x = fabs(((double)(i - center) - (double)ph / phases) * factor);
                if (x < 1.0) y = 1 - 3 * x * x + 2 * x * x * x + d * (-x * x + x * x * x);
                else         y = d * (-4 + 8 * x - 5 * x * x + x * x * x);
            } else {
                y = x == 0 ? 1.0 : sin(x) / x;
                w = 2.0 * x / (factor * taps * M_PI);
                y *= bessel_i0(type * sqrt(1 - w * w > 0 ? 1 - w * w : 0));
            }
            tab[i] = y;
            norm += y;
        }
        for (int i = 0; i < taps; i++) {
            long c = lrintf(tab[i] * scale / norm);
            bank[ph * taps + i] = (int16_t)(c < -32768 ? -32768 : c > 32767 ? 32767 : c);
        }
    }
    free(tab);
}

/* ---- picture scaler: img_resample_init (imgresample.c:433-485: 4 taps, 16 phases, 8-bit coefficients,
 * h_incr / v_incr in 16.16 from the LUMA sizes, cubic banks for factor (float)out/(float)in) and img_resample
 * (:487-507: planes 1 and 2 at sizes >> 1) -> component_resample (:362-431).  Per output row the filter's
 * bottom tap sits at source row (2*65536 + y*v_incr) >> 16, per output column the first tap at
 * (-65536 + x*h_incr) >> 16; taps outside the plane repeat the edge (h_resample_slow :289-323, row clamp
 * :381-386); the horizontal result is clamped and kept as a byte (the line buffer) before the vertical pass. */
static void scale_plane(const uint8_t *in, int iw, int ih, int ils, uint8_t *out, int ow, int oh, int ols,
                        int h_incr, int v_incr, const int16_t *hf, const int16_t *vf)
{
    for (int y = 0; y < oh; y++) {
        const int sy = 2 * 65536 + y * v_incr, bottom = sy >> 16;
        const int16_t *fv = vf + 4 * ((sy >> 12) & 15);
        for (int x = 0; x < ow; x++) {
            const int sx = -65536 + x * h_incr, left = sx >> 16;
            const int16_t *fh = hf + 4 * ((sx >> 12) & 15);
            int acc = 0;
            for (int j = 0; j < 4; j++) {
                const uint8_t *row = in + (size_t)ils * (size_t)clampi(bottom - 3 + j, 0, ih - 1);
                int hs = 0;
                for (int t = 0; t < 4; t++) hs += row[clampi(left + t, 0, iw - 1)] * fh[t];
                acc += clip_u8(hs >> 8) * fv[j];
            }
            out[(size_t)ols * y + x] = clip_u8(acc >> 8);
        }
    }
}
/* tight planes [n][h][w], chroma stored (w+1)/2 x (h+1)/2; the reference writes (ow>>1) x (oh>>1) of them */
AMVO_API int amvo_scale_frames(const uint8_t *y, const uint8_t *u, const uint8_t *v, int n, int iw, int ih, int ow, int oh,
                               uint8_t *oy, uint8_t *ou, uint8_t *ov)
{
    if (iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0) return -1;
    int16_t hf[64], vf[64];
    build_bank(hf, (float)ow / (float)iw, 4, 16, 256, 0);
    build_bank(vf, (float)oh / (float)ih, 4, 16, 256, 0);
    const int h_incr = (iw * 65536) / ow, v_incr = (ih * 65536) / oh;
    const int icw = (iw + 1) >> 1, ich = (ih + 1) >> 1, ocw = (ow + 1) >> 1, och = (oh + 1) >> 1;
    for (int i = 0; i < n; i++) {
        scale_plane(y + (size_t)i * iw * ih, iw, ih, iw, oy + (size_t)i * ow * oh, ow, oh, ow, h_incr, v_incr, hf, vf);
        scale_plane(u + (size_t)i * icw * ich, iw >> 1, ih >> 1, icw, ou + (size_t)i * ocw * och, ow >> 1, oh >> 1, ocw, h_incr, v_incr, hf, vf);
        scale_plane(v + (size_t)i * icw * ich, iw >> 1, ih >> 1, icw, ov + (size_t)i * ocw * och, ow >> 1, oh >> 1, ocw, h_incr, v_incr, hf, vf);
    }
    return n;
}
AMVO_API void amvo_scale_banks(int iw, int ih, int ow, int oh, int16_t *hf, int16_t *vf)
{
    build_bank(hf, (float)ow / (float)iw, 4, 16, 256, 0);
    build_bank(vf, (float)oh / (float)ih, 4, 16, 256, 0);
}

/* ---- audio resampler: audio_resample_init(1 output channel, in_ch, out_rate, in_rate) (resample.c:93-129:
 * 16 taps, 1024 phases, not linear, cutoff 0.8) -> av_resample_init (resample2.c:185-206: filter_length =
 * ceil(16 / min(out*0.8/in, 1)), Kaiser beta 9, coefficients scaled to 1 << 15, first output centred on sample
 * 0: index = -1024 * ((len - 1) / 2)); audio_resample (resample.c:131-235): two input channels are averaged
 * (stereo_to_mono :53-75, (l + r) >> 1) before av_resample (resample2.c:234-323).
 * The reference walks `index` by in_rate*1024 / out_rate per output with the remainder kept in `frac`
 * (:290-296), carries unconsumed samples from call to call (resample.c:214-216) and never flushes; over the
 * whole stream output k therefore starts at index0 + floor(k * in_rate * 1024 / out_rate), which this
 * restatement evaluates directly (64-bit), for any split of the stream into calls (calls of at least
 * filter_length samples; the taps left of sample 0 mirror: src[|i| % src_size], :256-258).
 * The sum is the reference's 32-bit accumulator (wrapping), rounded >> 15 and saturated (:274-275).
 * Returns the samples written (all outputs whose taps lie inside the n_in samples), -1 if out_cap is short. */
AMVO_API int amvo_resample_filter_length(int in_rate, int out_rate)
{
    double factor = out_rate * 0.8 / in_rate;
    if (factor > 1.0) factor = 1.0;
    int len = (int)ceil(16 / factor);
    return len < 1 ? 1 : len;
}
AMVO_API int amvo_resample_bank(int in_rate, int out_rate, int16_t *bank /* len * 1024 */)
{
    double factor = out_rate * 0.8 / in_rate;
    if (factor > 1.0) factor = 1.0;
    const int len = amvo_resample_filter_length(in_rate, out_rate);
    build_bank(bank, factor, len, 1024, 1 << 15, 9);
    return len;
}
AMVO_API int64_t amvo_audio_resample(const int16_t *in, int64_t n_in, int in_ch, int in_rate, int out_rate,
                                     int16_t *out, int64_t out_cap)
{
    if (n_in <= 0 || in_rate <= 0 || out_rate <= 0 || (in_ch != 1 && in_ch != 2)) return -1;
    const int len = amvo_resample_filter_length(in_rate, out_rate);
    int16_t *bank = (int16_t *)malloc(sizeof(int16_t) * (size_t)len * 1024);
    int16_t *mono = (int16_t *)malloc(sizeof(int16_t) * (size_t)n_in);
    amvo_resample_bank(in_rate, out_rate, bank);
    for (int64_t i = 0; i < n_in; i++) mono[i] = in_ch == 2 ? (int16_t)((in[2 * i] + in[2 * i + 1]) >> 1) : in[i];
    const int64_t index0 = -1024 * (int64_t)((len - 1) / 2);
    int64_t k = 0;
    for (;; k++) {
        const int64_t index = index0 + (k * (int64_t)in_rate * 1024) / out_rate;
        const int64_t first = index >> 10;
        const int16_t *f = bank + (size_t)len * (size_t)(index & 1023);
        uint32_t acc = 0;
        if (first < 0) {
            for (int i = 0; i < len; i++) {
                int64_t a = first + i; if (a < 0) a = -a;
                acc += (uint32_t)(mono[a % n_in] * f[i]);
            }
        } else if (first + len > n_in) {
            break;
        } else {
            for (int i = 0; i < len; i++) acc += (uint32_t)(mono[first + i] * f[i]);
        }
        if (k >= out_cap) { k = -1; break; }
        int32_t val = ((int32_t)(acc + (1u << 14))) >> 15;
        out[k] = (int16_t)(val < -32768 ? -32768 : val > 32767 ? 32767 : val);
    }
    free(bank); free(mono);
    return k;
}

// amv_tables.cuh -- the fixed tables of the AMV codec path and the host-side
// builders that turn them into the lookup structures the kernels stage in
// shared memory.  Values are the format's (JPEG Annex K.3 Huffman specs, the
// AMV decoder's two quantiser tables, MPEG-1 intra matrix, IMA step sizes);
// reference locations are cited per table.  Layouts are ours.
#pragma once
#include "amv_common.cuh"
#include <string.h>

namespace amv {

// zigzag scan position -> raster index (dsputil.c:50-59)
static const uint8_t kZigzag[64] = {
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

// same table, usable with compile-time indices inside fully unrolled device loops
AMV_HD constexpr int zigzag_at(int k) {
    constexpr uint8_t t[64] = {
         0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5,
        12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
        35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
        58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
    return t[k];
}

// raster index -> zigzag scan position
AMV_HD constexpr int zigzag_inv_at(int j) {
    constexpr uint8_t t[64] = {
         0,  1,  5,  6, 14, 15, 27, 28,  2,  4,  7, 13, 16, 26, 29, 42,
         3,  8, 12, 17, 25, 30, 41, 43,  9, 11, 18, 24, 31, 40, 44, 53,
        10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
        21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };
    return t[j];
}

// Decoder quantisers in zigzag order: sp5x_quant_table[10] / [11] (sp5x.h:187-195),
// selected by sp5xdec.c:40,60-61.
static const uint8_t kDecQuant[2][64] = {
  { 13,  9, 10, 11, 10,  8, 13, 11, 10, 11, 14, 14, 13, 15, 19, 32,
    21, 19, 18, 18, 19, 39, 28, 30, 23, 32, 46, 41, 49, 48, 46, 41,
    45, 44, 51, 58, 74, 62, 51, 54, 70, 55, 44, 45, 64, 87, 65, 70,
    76, 78, 82, 83, 82, 50, 62, 90, 97, 90, 80, 96, 74, 81, 82, 79 },
  { 14, 14, 14, 19, 17, 19, 38, 21, 21, 38, 79, 53, 45, 53, 79, 79,
    79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79,
    79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79,
    79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79, 79 } };

// amvlib's decoder quantisers, zigzag order: amv_luminance_quant_tbl / amv_chrominance_quant_tbl
// (C-AMVDecoder/amvlib/AmvJpeg.c:30-39,52-61)
static const uint8_t kAmvlibQuant[2][64] = {
  {  8,  6,  6,  7,  6,  5,  8,  7,  7,  7,  9,  9,  8, 10, 12, 20,
    13, 12, 11, 11, 12, 25, 18, 19, 15, 20, 29, 26, 31, 30, 29, 26,
    28, 28, 32, 36, 46, 39, 32, 34, 44, 39, 28, 28, 40, 55, 41, 44,
    48, 49, 52, 52, 52, 31, 39, 57, 61, 56, 50, 60, 46, 51, 52, 50 },
  {  9,  9,  9, 12, 11, 12, 24, 13, 13, 24, 50, 33, 28, 33, 50, 50,
    50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50,
    50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50,
    50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50 } };

// Encoder matrix base, raster order: ff_mpeg1_default_intra_matrix (mpeg12data.c:30-39)
static const uint8_t kEncIntraBase[64] = {
     8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83 };

// Huffman specs (mjpeg.c:65-126): table order DC-luma, DC-chroma, AC-luma, AC-chroma
static const uint8_t kHuffCount[4][16] = {
    { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 },
    { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 },
    { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125 },
    { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119 } };
static const uint8_t kHuffSymDC[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t kHuffSymACLuma[162] = {
    0x01,0x02,0x03,0x00,0x04,0x11,0x05,0x12,0x21,0x31,0x41,0x06,0x13,0x51,0x61,0x07,0x22,0x71,
    0x14,0x32,0x81,0x91,0xa1,0x08,0x23,0x42,0xb1,0xc1,0x15,0x52,0xd1,0xf0,0x24,0x33,0x62,0x72,
    0x82,0x09,0x0a,0x16,0x17,0x18,0x19,0x1a,0x25,0x26,0x27,0x28,0x29,0x2a,0x34,0x35,0x36,0x37,
    0x38,0x39,0x3a,0x43,0x44,0x45,0x46,0x47,0x48,0x49,0x4a,0x53,0x54,0x55,0x56,0x57,0x58,0x59,
    0x5a,0x63,0x64,0x65,0x66,0x67,0x68,0x69,0x6a,0x73,0x74,0x75,0x76,0x77,0x78,0x79,0x7a,0x83,
    0x84,0x85,0x86,0x87,0x88,0x89,0x8a,0x92,0x93,0x94,0x95,0x96,0x97,0x98,0x99,0x9a,0xa2,0xa3,
    0xa4,0xa5,0xa6,0xa7,0xa8,0xa9,0xaa,0xb2,0xb3,0xb4,0xb5,0xb6,0xb7,0xb8,0xb9,0xba,0xc2,0xc3,
    0xc4,0xc5,0xc6,0xc7,0xc8,0xc9,0xca,0xd2,0xd3,0xd4,0xd5,0xd6,0xd7,0xd8,0xd9,0xda,0xe1,0xe2,
    0xe3,0xe4,0xe5,0xe6,0xe7,0xe8,0xe9,0xea,0xf1,0xf2,0xf3,0xf4,0xf5,0xf6,0xf7,0xf8,0xf9,0xfa };
static const uint8_t kHuffSymACChroma[162] = {
    0x00,0x01,0x02,0x03,0x11,0x04,0x05,0x21,0x31,0x06,0x12,0x41,0x51,0x07,0x61,0x71,0x13,0x22,
    0x32,0x81,0x08,0x14,0x42,0x91,0xa1,0xb1,0xc1,0x09,0x23,0x33,0x52,0xf0,0x15,0x62,0x72,0xd1,
    0x0a,0x16,0x24,0x34,0xe1,0x25,0xf1,0x17,0x18,0x19,0x1a,0x26,0x27,0x28,0x29,0x2a,0x35,0x36,
    0x37,0x38,0x39,0x3a,0x43,0x44,0x45,0x46,0x47,0x48,0x49,0x4a,0x53,0x54,0x55,0x56,0x57,0x58,
    0x59,0x5a,0x63,0x64,0x65,0x66,0x67,0x68,0x69,0x6a,0x73,0x74,0x75,0x76,0x77,0x78,0x79,0x7a,
    0x82,0x83,0x84,0x85,0x86,0x87,0x88,0x89,0x8a,0x92,0x93,0x94,0x95,0x96,0x97,0x98,0x99,0x9a,
    0xa2,0xa3,0xa4,0xa5,0xa6,0xa7,0xa8,0xa9,0xaa,0xb2,0xb3,0xb4,0xb5,0xb6,0xb7,0xb8,0xb9,0xba,
    0xc2,0xc3,0xc4,0xc5,0xc6,0xc7,0xc8,0xc9,0xca,0xd2,0xd3,0xd4,0xd5,0xd6,0xd7,0xd8,0xd9,0xda,
    0xe2,0xe3,0xe4,0xe5,0xe6,0xe7,0xe8,0xe9,0xea,0xf2,0xf3,0xf4,0xf5,0xf6,0xf7,0xf8,0xf9,0xfa };

// IMA ADPCM step sizes (adpcm.c:64-75)
static const uint16_t kImaStep[89] = {
        7,     8,     9,    10,    11,    12,    13,    14,    16,    17,    19,    21,
       23,    25,    28,    31,    34,    37,    41,    45,    50,    55,    60,    66,
       73,    80,    88,    97,   107,   118,   130,   143,   157,   173,   190,   209,
      230,   253,   279,   307,   337,   371,   408,   449,   494,   544,   598,   658,
      724,   796,   876,   963,  1060,  1166,  1282,  1411,  1552,  1707,  1878,  2066,
     2272,  2499,  2749,  3024,  3327,  3660,  4026,  4428,  4871,  5358,  5894,  6484,
     7132,  7845,  8630,  9493, 10442, 11487, 12635, 13899, 15289, 16818, 18500, 20350,
    22385, 24623, 27086, 29794, 32767 };

// ----------------------------------------------------------------------------
// Decoder lookup structure.
//
// One uint32 array holds, for each of the four Huffman tables, a 1024-entry first
// level indexed by the next 10 bits, followed by 64-entry second-level tables for
// the few 10-bit prefixes that continue into longer codes (canonical JPEG codes put
// every code longer than 10 bits behind a handful of all-ones prefixes).
//
// entry: [4:0]  code length            [8:5] size (magnitude bits that follow)
//        [12:9] run                    bit 13 = pointer: [12:0] = index of the second-level table
//        bit 14 = no such code (length field 1 so a garbage lane still advances)
//        bit 15 = resolved: code AND magnitude bits fit in the 10 index bits, so
//        [31:16] already holds the finished value -- for AC tables the token
//        (run << 12 | level & 0xfff), for DC tables the signed difference.
// EOB is (run 0, size 0); ZRL is (run 15, size 0).
constexpr int kVlcFirstBits = 10;
constexpr int kVlcSecondBits = 6;
constexpr uint32_t kVlcPtr = 1u << 13;
constexpr uint32_t kVlcBad = 1u << 14;
constexpr uint32_t kVlcResolved = 1u << 15;
constexpr int kVlcMaxEntries = 4 * 1024 + 20 * 64;

// Tokens handed from the Huffman kernel to the IDCT kernel (32 bit each).  A block is its DC token
// (low half = absolute dequantised DC as int16) followed by one token per non-zero AC coefficient:
//   [31:16] byte offset of the coefficient inside the consumer's shared-memory column
//           ((j >> 1) * 128 + (j & 1) * 2 for raster index j), [15:0] level * quantiser as int16.
// The block-offset table entry is (AC token count << 24) | index of the DC token in the frame's region.
constexpr uint32_t kTokCountShift = 24;

struct VlcTables {
    uint32_t e[kVlcMaxEntries];
    int      base[4];        // first-level base of DC-luma, DC-chroma, AC-luma, AC-chroma
    int      count;          // entries used
};

// zigzag position -> (raster index | quantiser << 8), per component class (luma, chroma);
// tz: the same for the token producer: (column byte offset << 16) | quantiser
struct DequantTables { uint32_t zq[2][64]; uint32_t tz[2][64]; };

// amvlib flavour of the token producer's table.  amvlib's raster->zigzag table has a
// typo (AmvJpeg.c:131-141: raster (3,4) reads index 37 instead of 31), so zigzag coefficient 31 is
// never used and coefficient 37 lands at two raster positions.  Entry per zigzag position k:
//   [7:0] quantiser   bit 8 = coefficient is dropped   bit 9 = second position valid
//   [15:10] second raster position   [31:26] first raster position
// amvlib tokens: [31:26] raster position, [25:0] coefficient * quantiser (two's complement, 26 bits:
// the product of a 16-bit level and an 8-bit quantiser always fits).
struct AmvlibDequantTables { uint32_t tz[2][64]; };
constexpr uint32_t kAmvlibTokSkip = 1u << 8, kAmvlibTokDup = 1u << 9;

// Encoder: symbol -> (code << 5 | length).  Index: DC-luma 0..15, DC-chroma 16..31,
// AC-luma 32..287, AC-chroma 288..543.
constexpr int kEncDcLuma = 0, kEncDcChroma = 16, kEncAcLuma = 32, kEncAcChroma = 288, kEncHuffEntries = 544;
struct EncHuffTables { uint32_t e[kEncHuffEntries]; };

inline const uint8_t *huff_symbols(int t) {
    return t < 2 ? kHuffSymDC : (t == 2 ? kHuffSymACLuma : kHuffSymACChroma);
}

// canonical code assignment (ff_mjpeg_build_huffman_codes, mjpeg.c:129-147) from a code specification:
// codes per length 1..16 and the symbols in code order (what a DHT segment carries)
inline void huff_codes_from(const uint8_t counts[16], const uint8_t *sym, uint8_t len[256], uint16_t code[256]) {
    memset(len, 0, 256);
    memset(code, 0, 512);
    unsigned next = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int j = 0; j < counts[l - 1] && k < 256; j++, k++) {
            len[sym[k]] = (uint8_t)l;
            code[sym[k]] = (uint16_t)next++;
        }
        next <<= 1;
    }
}
// a specification is usable if its codes fit their lengths (no over-subscription) and it has at most 256 symbols
inline bool huff_counts_valid(const uint8_t counts[16]) {
    unsigned next = 0, total = 0;
    for (int l = 1; l <= 16; l++) {
        next += counts[l - 1];
        total += counts[l - 1];
        if (next > (1u << l)) return false;
        next <<= 1;
    }
    return total <= 256;
}
inline void huff_codes(int t, uint8_t len[256], uint16_t code[256]) { huff_codes_from(kHuffCount[t], huff_symbols(t), len, code); }

// the code specification of the four tables a scan uses: DC of component 0, DC of components 1/2, then the AC tables
struct HuffSpec { uint8_t counts[4][16]; uint8_t syms[4][256]; };
inline void fixed_huff_spec(HuffSpec &H) {
    memset(&H, 0, sizeof(H));
    for (int t = 0; t < 4; t++) {
        memcpy(H.counts[t], kHuffCount[t], 16);
        memcpy(H.syms[t], huff_symbols(t), t < 2 ? 12 : 162);
    }
}

inline int jpeg_extend(int v, int size) { return size && v < (1 << (size - 1)) ? v - ((1 << size) - 1) : v; }

// false: more long-code prefixes than the second-level area holds (never for the fixed tables)
inline bool build_vlc_tables_from(VlcTables &T, const HuffSpec &H) {
    for (int i = 0; i < kVlcMaxEntries; i++) T.e[i] = kVlcBad | 1;
    int used = 0;
    for (int t = 0; t < 4; t++) {
        uint8_t len[256]; uint16_t code[256];
        huff_codes_from(H.counts[t], H.syms[t], len, code);
        T.base[t] = used;
        used += 1 << kVlcFirstBits;
        for (int s = 0; s < 256; s++) {
            if (!len[s]) continue;
            const int run = t < 2 ? 0 : (s >> 4), size = t < 2 ? s : (s & 15);
            if (size > 15) return false;                   // the entry's size field has four bits
            const uint32_t ent = (uint32_t)(len[s] | (size << 5) | (run << 9));
            if (len[s] <= kVlcFirstBits) {
                const int spare = kVlcFirstBits - len[s];
                const int lo = code[s] << spare;
                for (int i = 0; i < (1 << spare); i++) {
                    uint32_t e = ent;
                    if (size <= spare) {
                        // the magnitude bits are the top `size` bits of i: finish the value here
                        const int v = jpeg_extend(size ? i >> (spare - size) : 0, size);
                        const uint32_t val = t < 2 ? (uint32_t)(v & 0xffff) : (uint32_t)((run << 12) | (v & 0xfff));
                        e |= kVlcResolved | (val << 16);
                    }
                    T.e[T.base[t] + lo + i] = e;
                }
            } else {
                const int pre = code[s] >> (len[s] - kVlcFirstBits);
                uint32_t &slot = T.e[T.base[t] + pre];
                if (!(slot & kVlcPtr)) {
                    if (used + (1 << kVlcSecondBits) > kVlcMaxEntries) return false;
                    slot = kVlcPtr | (uint32_t)used;
                    used += 1 << kVlcSecondBits;
                }
                const int sub = slot & 0x1fff;
                const int rest = len[s] - kVlcFirstBits;               // 1..6 bits left
                const int lo = (code[s] & ((1 << rest) - 1)) << (kVlcSecondBits - rest);
                for (int i = 0; i < (1 << (kVlcSecondBits - rest)); i++) T.e[sub + lo + i] = ent;
            }
        }
    }
    T.count = used;
    return true;
}
inline void build_vlc_tables(VlcTables &T) { HuffSpec H; fixed_huff_spec(H); build_vlc_tables_from(T, H); }

// Second view of the same codes, for the flat (one symbol per iteration, no DC/AC branch) token
// kernel.  DC and AC symbols go through the same code: a block starts with kb = 0 (kb = zigzag
// position + 1), every entry carries the amount kb advances by and whether the symbol yields a
// token, and the end of a block shows as kb >= 64.
//   direct entry : [4:0] code length (>= 1)  [7] yields a 16-bit token  [15:8] bits to consume  [20:16] size
//                  [30:23] advance of kb  [31] symbol yields a token
//                  advance: DC 1, AC coefficient run + 1, ZRL 16, EOB 128
//   [4:0] == 0   : bit 5 = no such code, else [23:8] = index of the second-level table, [31:24] its index bits
// First level: kFlatDcBits / kFlatAcBits index bits; second level: the remaining bits up to 16.
constexpr int kFlatDcBits = 10, kFlatAcBits = 12;
constexpr int kFlatSecondCap = 1024;
constexpr int kFlatMaxEntries = 2 * (1 << kFlatDcBits) + 2 * (1 << kFlatAcBits) + kFlatSecondCap;
constexpr uint32_t kFlatBad = 1u << 5;
constexpr uint32_t kFlatTok16 = 1u << 7;        // direct entries: the symbol yields a 16-bit token (coefficients, DC and ZRL)
constexpr uint32_t kFlatAdvEob = 128;
struct FlatVlcTables {
    uint32_t e[kFlatMaxEntries];
    int      base[4];        // DC-luma, DC-chroma, AC-luma, AC-chroma
    int      count;          // entries used; > kFlatMaxEntries: the codes do not fit (never for the fixed tables)
};

// generic over the code specification (counts per length 1..16, symbols in code order), so tables
// read from a DHT segment build the same way as the fixed ones
inline bool build_flat_vlc_table(FlatVlcTables &F, int t, const uint8_t counts[16], const uint8_t *syms, int &used) {
    const bool dc = t < 2;
    const int fb = dc ? kFlatDcBits : kFlatAcBits, sb = 16 - fb;
    F.base[t] = used;
    used += 1 << fb;
    if (used > kFlatMaxEntries) return false;
    for (int i = 0; i < (1 << fb); i++) F.e[F.base[t] + i] = kFlatBad;
    unsigned next = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
        for (int j = 0; j < counts[l - 1]; j++, k++) {
            const unsigned code = next++;
            const int s = syms[k];
            const int run = dc ? 0 : (s >> 4), size = dc ? s : (s & 15);
            if (dc && size > 16) return false;
            uint32_t adv = dc ? 1u : (uint32_t)run + 1u;
            if (!dc && size == 0) adv = run == 15 ? 16u : kFlatAdvEob;          // ZRL / EOB (other run,0 symbols: treated as EOB)
            const uint32_t emit = (dc || size) ? 1u : 0u;
            const uint32_t tok16 = (emit || (!dc && size == 0 && run == 15)) ? kFlatTok16 : 0u;
            const uint32_t ent = (uint32_t)l | tok16 | ((uint32_t)(l + size) << 8) | ((uint32_t)size << 16) | (adv << 23) | (emit << 31);
            if (l <= fb) {
                const int spare = fb - l;
                for (int i = 0; i < (1 << spare); i++) F.e[F.base[t] + (code << spare) + i] = ent;
            } else {
                const unsigned pre = code >> (l - fb);
                uint32_t &slot = F.e[F.base[t] + pre];
                if (slot == kFlatBad) {
                    if (used + (1 << sb) > kFlatMaxEntries) return false;
                    slot = ((uint32_t)used << 8) | ((uint32_t)sb << 24);
                    for (int i = 0; i < (1 << sb); i++) F.e[used + i] = kFlatBad;
                    used += 1 << sb;
                }
                const int sub = (slot >> 8) & 0xffff;
                const int rest = l - fb;
                const int lo = (code & ((1u << rest) - 1u)) << (sb - rest);
                for (int i = 0; i < (1 << (sb - rest)); i++) F.e[sub + lo + i] = ent;
            }
        }
        next <<= 1;
    }
    return true;
}

inline bool build_flat_vlc_tables_from(FlatVlcTables &F, const HuffSpec &H) {
    int used = 0;
    bool ok = true;
    for (int t = 0; t < 4; t++) ok = ok && build_flat_vlc_table(F, t, H.counts[t], H.syms[t], used);
    F.count = ok ? used : kFlatMaxEntries + 1;
    return ok;
}
inline void build_flat_vlc_tables(FlatVlcTables &F) { HuffSpec H; fixed_huff_spec(H); build_flat_vlc_tables_from(F, H); }

inline void build_dequant_tables_from(DequantTables &D, const uint8_t qzz[2][64]) {
    for (int c = 0; c < 2; c++)
        for (int k = 0; k < 64; k++) {
            const uint32_t j = kZigzag[k];
            D.zq[c][k] = j | ((uint32_t)qzz[c][k] << 8);
            D.tz[c][k] = (((j >> 1) * 128u + (j & 1u) * 2u) << 16) | (uint32_t)qzz[c][k];
        }
}
inline void build_dequant_tables(DequantTables &D) { build_dequant_tables_from(D, kDecQuant); }

inline void build_amvlib_dequant_tables(AmvlibDequantTables &D) {
    for (int c = 0; c < 2; c++)
        for (int k = 0; k < 64; k++) {
            uint32_t e = (uint32_t)kAmvlibQuant[c][k] | ((uint32_t)kZigzag[k] << 26);
            if (k == 31) e |= kAmvlibTokSkip;                                // no raster position reads index 31
            if (k == 37) e |= kAmvlibTokDup | ((uint32_t)(3 * 8 + 4) << 10);  // (3,4) reads 37 as well
            D.tz[c][k] = e;
        }
}

inline void build_enc_huff_tables(EncHuffTables &E) {
    memset(E.e, 0, sizeof(E.e));
    const int base[4] = { kEncDcLuma, kEncDcChroma, kEncAcLuma, kEncAcChroma };
    for (int t = 0; t < 4; t++) {
        uint8_t len[256]; uint16_t code[256];
        huff_codes(t, len, code);
        for (int s = 0; s < (t < 2 ? 16 : 256); s++)
            E.e[base[t] + s] = ((uint32_t)code[s] << 5) | len[s];
    }
}

}  // namespace amv

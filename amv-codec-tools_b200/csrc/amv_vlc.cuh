// amv_vlc.cuh -- per-lane Huffman decoding of the AMV scan (mjpegdec.c:358-430).
// A "lane" is one CUDA thread walking one subsequence of one frame's un-stuffed
// bitstream.  The same code runs in two modes:
//   EMIT = false : synchronisation / counting pass (no coefficients written)
//   EMIT = true  : output pass (dequantised coefficients go to the lane's slot)
#pragma once
#include "amv_common.cuh"
#include "amv_tables.cuh"

namespace amv {

// MSB-first reader over big-endian bytes stored in 32-bit words (the un-stuffed
// scratch slot is 16-byte aligned and zero padded).  Keeps >32 valid bits in
// `acc` before every symbol and one word of look-ahead in `nxt` so the load
// latency of the next word hides behind ~5 symbols of work.
struct BitReader {
    const uint32_t *base;
    uint32_t nwords;      // words that may be read; beyond that zeros are supplied
    uint32_t widx;        // words merged into acc so far
    uint64_t acc;
    int      nb;          // valid bits in acc (counted from the MSB)
    uint32_t nxt;

    AMV_HD uint32_t load(uint32_t i) const {
        if (i >= nwords) return 0;
#if defined(__CUDA_ARCH__)
        return bswap32(__ldg(base + i));
#else
        return bswap32(base[i]);
#endif
    }
    AMV_HD void init(const uint32_t *b, uint32_t nw, uint32_t bitpos) {
        base = b; nwords = nw;
        widx = bitpos >> 5;
        const uint32_t sh = bitpos & 31;
        acc = (((uint64_t)load(widx) << 32) | load(widx + 1)) << sh;
        nb = 64 - (int)sh;
        widx += 2;
        nxt = load(widx);
    }
    AMV_HD void refill() {
        if (nb <= 32) {
            acc |= (uint64_t)nxt << (32 - nb);
            nb += 32;
            widx++;
            nxt = load(widx);
        }
    }
    AMV_HD uint32_t peek32() const { return (uint32_t)(acc >> 32); }
    AMV_HD void skip(int n) { acc <<= n; nb -= n; }
    AMV_HD uint32_t bitpos() const { return widx * 32u - (uint32_t)nb; }
};

// value of the next `size` bits of p (already aligned so that they are the top
// bits), JPEG-extended (get_xbits, bitstream.h:629-639)
AMV_HD int extend_bits(uint32_t top, int size) {
#if defined(__CUDA_ARCH__)
    const uint32_t v = __funnelshift_l(top, 0u, size);      // top >> (32-size), 0 for size 0
#else
    const uint32_t v = size ? top >> (32 - size) : 0u;
#endif
    const uint32_t half = (1u << size) >> 1;
    return v < half ? (int)v - (int)((1u << size) - 1u) : (int)v;
}

AMV_HD uint32_t vlc_lookup(const uint16_t *lut, int base, uint32_t p) {
    uint32_t e = lut[base + (p >> (32 - kVlcFirstBits))];
    if (e & kVlcPtr) e = lut[(e & 0x1fffu) + ((p >> (32 - kVlcFirstBits - kVlcSecondBits)) & ((1u << kVlcSecondBits) - 1u))];
    return e;
}

// Decode one 8x8 block.
//   tq       : 0 luma tables, 1 chroma tables
//   dc_diff  : receives the DC difference (caller owns the predictor chain)
//   put(k,v) : EMIT only -- called for every AC coefficient, k = zigzag position 1..63
// Returns status bits (0 = clean).
template <bool EMIT, class Put>
AMV_HD uint32_t decode_block(BitReader &br, const uint16_t *lut, const int *tbl_base, int tq, int &dc_diff, Put put) {
    uint32_t st = 0;
    br.refill();
    uint32_t p = br.peek32();
    uint32_t e = vlc_lookup(lut, tbl_base[tq], p);
    int len = e & 31, size = (e >> 5) & 15;
    if (e & kVlcBad) st |= AMV_ST_BADCODE;
    dc_diff = extend_bits(p << len, size);
    br.skip(len + size);
    int k = 0;
    for (;;) {
        br.refill();
        p = br.peek32();
        e = vlc_lookup(lut, tbl_base[2 + tq], p);
        len = e & 31; size = (e >> 5) & 15;
        const int run = (e >> 9) & 15;
        br.skip(len + size);
        if (e & kVlcBad) { st |= AMV_ST_BADCODE; break; }
        if (size == 0) {
            if (run != 15) break;          // EOB
            k += 16;                       // ZRL: no range check in the reference (mjpegdec.c:400-401)
            if (k > 1024) break;           // only garbage lanes get here; keeps them bounded
            continue;
        }
        k += run + 1;
        if (k > 63) { st |= AMV_ST_COEFIDX; break; }
        if (EMIT) put(k, extend_bits(p << len, size));
        if (k == 63) break;
    }
    return st;
}

// What a lane knows when it hands over to its right neighbour: where the first
// block that starts at or after the subsequence boundary begins, which block of
// the 6-block MCU that is, and what it accumulated on the way.
struct LaneExit {
    uint32_t bitpos;     // absolute bit position in the un-stuffed scan
    uint32_t phase;      // block-in-MCU index 0..5 of the next block
    uint32_t nblocks;    // blocks decoded by this lane
    int      dc[3];      // sum of DC differences per component over those blocks
};

}  // namespace amv

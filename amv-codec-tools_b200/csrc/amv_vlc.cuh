// amv_vlc.cuh -- per-lane Huffman decoding of the AMV scan (mjpegdec.c:358-430).
// A "lane" is one CUDA thread walking one subsequence of one frame's un-stuffed
// bitstream.  The same walker serves the synchronisation/counting pass and the
// token-emitting pass; what differs is the Sink it reports to.
#pragma once
#include "amv_common.cuh"
#include "amv_tables.cuh"

namespace amv {

// MSB-first reader over big-endian bytes stored in 32-bit words (the un-stuffed
// scratch slot is 16-byte aligned and zero padded).  Keeps >32 valid bits in
// `acc` before every symbol.  `raw` is the NEXT word, already requested from
// memory but not yet touched, so its latency hides behind ~5 symbols of work.
struct BitReader {
    const uint32_t *base;
    uint32_t nwords;      // words that may be read; beyond that zeros are supplied
    uint32_t widx;        // words merged into acc so far
    uint64_t acc;
    int      nb;          // valid bits in acc (counted from the MSB)
    uint32_t raw;         // word widx, as loaded (little-endian view of big-endian bytes)

    AMV_HD uint32_t load(uint32_t i) const {
        if (i >= nwords) return 0;
#if defined(__CUDA_ARCH__)
        return __ldg(base + i);
#else
        return base[i];
#endif
    }
    AMV_HD void init(const uint32_t *b, uint32_t nw, uint32_t bitpos) {
        base = b; nwords = nw;
        widx = bitpos >> 5;
        const uint32_t sh = bitpos & 31;
        acc = (((uint64_t)bswap32(load(widx)) << 32) | bswap32(load(widx + 1))) << sh;
        nb = 64 - (int)sh;
        widx += 2;
        raw = load(widx);
    }
    AMV_HD void refill() {
        if (nb <= 32) {
            acc |= (uint64_t)bswap32(raw) << (32 - nb);
            nb += 32;
            widx++;
            raw = load(widx);
        }
    }
    AMV_HD uint32_t peek32() const { return (uint32_t)(acc >> 32); }
    AMV_HD void skip(int n) { acc <<= n; nb -= n; }
    AMV_HD uint32_t bitpos() const { return widx * 32u - (uint32_t)nb; }
};

// value of the next `size` bits (given left-aligned in `top`), JPEG-extended
// (get_xbits, bitstream.h:629-639)
AMV_HD int extend_bits(uint32_t top, int size) {
#if defined(__CUDA_ARCH__)
    const uint32_t v = __funnelshift_l(top, 0u, size);      // top >> (32-size), 0 for size 0
#else
    const uint32_t v = size ? top >> (32 - size) : 0u;
#endif
    const uint32_t half = (1u << size) >> 1;
    return v < half ? (int)v - (int)((1u << size) - 1u) : (int)v;
}

AMV_HD uint32_t vlc_lookup(const uint32_t *lut, int base, uint32_t p) {
    uint32_t e = lut[base + (p >> (32 - kVlcFirstBits))];
    if (e & kVlcPtr) e = lut[(e & 0x1fffu) + ((p >> (32 - kVlcFirstBits - kVlcSecondBits)) & ((1u << kVlcSecondBits) - 1u))];
    return e;
}

// Walk one 8x8 block.  Sink interface:
//   void dc(int diff)          -- DC difference (caller owns the predictor chain)
//   void ac(int k, int level)  -- every non-zero AC coefficient: zigzag position 1..63 and its level
// Returns status bits (0 = clean).  Mirrors decode_block (mjpegdec.c:376-430): ZRL advances
// without a range check, a coefficient index > 63 is the "error count" condition.
template <class Sink>
AMV_HD uint32_t walk_block(BitReader &br, const uint32_t *lut, const int *tbl_base, int tq, Sink &sink) {
    uint32_t st = 0;
    br.refill();
    uint32_t p = br.peek32();
    uint32_t e = vlc_lookup(lut, tbl_base[tq], p);
    int len = e & 31, size = (e >> 5) & 15;
    if (e & kVlcBad) st |= AMV_ST_BADCODE;
    sink.dc((e & kVlcResolved) ? (int)(int16_t)(e >> 16) : extend_bits(p << len, size));
    br.skip(len + size);
    int k = 0;
    for (;;) {
        br.refill();
        p = br.peek32();
        e = vlc_lookup(lut, tbl_base[2 + tq], p);
        len = e & 31; size = (e >> 5) & 15;
        const uint32_t run = (e >> 9) & 15;
        br.skip(len + size);
        if (e & kVlcBad) { st |= AMV_ST_BADCODE; break; }
        if (size == 0) {
            if (run != 15) break;            // EOB
            k += 16;                         // ZRL
            if (k > 1024) break;             // only garbage lanes get here; keeps them bounded
            continue;
        }
        k += (int)run + 1;
        if (k > 63) { st |= AMV_ST_COEFIDX; break; }
        sink.ac(k, (e & kVlcResolved) ? (int)((int32_t)(e << 4) >> 20) : extend_bits(p << len, size));
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

// amv_dec.cu -- AMV video decode kernels (sm_100a).
//
//   k_unstuff   : one CTA per packet; strips FF00 byte stuffing (mjpegdec.c:1137-1160 applied to
//                 sp5xdec.c:75-88's `payload ++ FF D9`) into a 16-byte aligned scratch slot.
//                 128-bit coalesced loads, ballot/popc + CTA scan for the compaction, shared-memory
//                 staging so the scratch is written in 128-bit units.
//   k_vlc_sync  : self-synchronising subsequence decode.  P lanes of a warp share a frame, each
//                 walks its own subsequence from a guessed state; lanes hand their exit state to
//                 the right neighbour with __shfl_up and the warp iterates until a __ballot shows
//                 no lane's entry state changed.  A segmented shuffle scan then gives every lane
//                 its first block index and DC predictors.
//   k_vlc_tokens: every lane re-walks its (now exactly delimited) subsequence and turns the
//                 variable-length codes into fixed-width 16-bit tokens + per-block offsets.
//   k_idct      : one thread per 8x8 block in plane raster order (coalesced 256-byte row stores):
//                 tokens -> dequantised coefficients in a conflict-free shared-memory column ->
//                 simple_idct in registers -> bottom-up store (mjpegdec.c:672-677,710-716).
#include <type_traits>
#include <stdlib.h>
#include "amv_common.cuh"
#include "amv_tables.cuh"
#include "amv_dct.cuh"
#include "amv_vlc.cuh"
#include "amv_kernels.h"

namespace amv {

// ------------------------------------------------------------------------------------------------
// exclusive scan of (aligned) sizes -> offsets, single CTA (n is at most a few million; the scan
// is a vanishing fraction of the work it feeds)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan_sizes(const uint32_t *__restrict__ size, int n, uint32_t align_mask,
                                                     uint32_t pad, uint64_t *__restrict__ off,
                                                     uint64_t *__restrict__ carry_io /* running total in/out, may be NULL */) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t running;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = carry_io ? *carry_io : 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        uint64_t v = 0;
        if (i < n) v = (uint64_t)((size[i] + align_mask) & ~align_mask) + pad;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += t;
            }
            warp_tot[lane] = w;          // inclusive over warps
        }
        __syncthreads();
        const uint64_t before = running + (wid ? warp_tot[wid - 1] : 0) + (inc - v);
        if (i < n) off[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) running = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && carry_io) *carry_io = running;
}

// ------------------------------------------------------------------------------------------------
// k_unstuff
//
// One CTA per packet, tiles of one 16-byte unit per thread.  Byte flags are computed four at a
// time with SIMD-in-register tricks (an FF byte shows up in ~6 % of the units, so any per-byte or
// per-unit branch would be taken by almost every warp); the kept bytes of a unit are closed up in
// registers (one short loop iteration per removed byte) and land in the shared-memory stage as at
// most five word-wide `red.shared.or` into a zeroed tile, which is then written out in 128-bit
// units.  Only the units that touch the ends of the payload take the byte-serial path.
// ------------------------------------------------------------------------------------------------

// 0x80 in every byte of v that is 00 (exact, no cross-byte carries)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) {
    return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v | 0x7f7f7f7fu);
}
// the four 0x80 byte flags of f as bits 0..3
__device__ __forceinline__ uint32_t flag_bits(uint32_t f) { return (((f >> 7) * 0x00204081u) >> 21) & 0xfu; }

template <int kUnstuffThreads>
__global__ void __launch_bounds__(kUnstuffThreads, kUnstuffThreads == 32 ? 32 : 1)
k_unstuff(const uint8_t *__restrict__ pkts, uint64_t pkts_bytes, const uint64_t *__restrict__ pkt_off,
          const uint32_t *__restrict__ pkt_size, int n, uint8_t *__restrict__ scratch,
          const uint64_t *__restrict__ slot_off, uint64_t scratch_bytes, uint32_t *__restrict__ scan_len,
          int32_t *__restrict__ status, uint32_t head, int sp5x) {
    // framing: AMV = FF D8 | stuffed scan | FF D9 (sp5xdec.c:75-77), head 2; plain JPEG = the same with the marker
    // segments up to the end of the SOS header in front, head = their length; SP5X = 14 header bytes | scan with
    // LITERAL FF bytes to the end of the packet (the reference stuffs them itself before handing over, :78-84)
    const uint32_t framing = sp5x ? head : head + 2u;
    constexpr int kUnstuffTile = kUnstuffThreads * 16;
    constexpr int kUnstuffStage = kUnstuffTile + 64;
    __shared__ __align__(16) uint8_t stage[kUnstuffStage];
    __shared__ uint32_t warp_cnt[kUnstuffThreads / 32];
    __shared__ uint32_t s_first_term;
    // a one-warp CTA (the form that runs) needs no CTA barrier: warp-level synchronisation and votes do
    auto cta_sync = [&]() { if (kUnstuffThreads == 32) __syncwarp(); else __syncthreads(); };
    auto cta_or = [&](bool v) -> bool { return kUnstuffThreads == 32 ? __any_sync(0xffffffffu, v) != 0 : __syncthreads_or(v) != 0; };
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t stage_s = smem_addr(stage);
    // invariant at the start of every tile: stage[0..carry) holds the carried bytes, the rest is zero
    for (int i = tid * 16; i < kUnstuffStage; i += kUnstuffThreads * 16) *reinterpret_cast<uint4 *>(stage + i) = make_uint4(0, 0, 0, 0);
    cta_sync();

    for (int f = blockIdx.x; f < n; f += gridDim.x) {
        const uint64_t off = pkt_off[f];
        const uint32_t size = pkt_size[f];
        const uint64_t slot = slot_off[f];
        int32_t st = 0;
        if (!range_ok(off, size, pkts_bytes) || !range_ok(slot, (uint64_t)((size + 15u) & ~15u) + kSlotPad, scratch_bytes)) {
            if (tid == 0) { scan_len[f] = 0; status[f] = AMV_ST_RANGE; }
            continue;
        }
        if (size < framing) st |= AMV_ST_SHORT;
        // virtual stream: payload bytes pkt[head .. size-tail) followed by FF D9
        const uint32_t npay = size >= framing ? size - framing : 0;
        const uint32_t V = npay + 2;
        const uint8_t *pay = pkts + off + head;
        const uint32_t mis = (uint32_t)((uintptr_t)pay & 15);       // bytes before the payload in its first 16 B unit
        const uint8_t *abase = pay - mis;
        uint8_t *dst = scratch + slot;
        uint32_t carry = 0;       // bytes waiting in stage[0..carry)
        uint32_t written = 0;     // bytes already flushed to dst
        bool done = false;

        // the unit of the NEXT tile is requested before the current tile is processed, so its
        // latency hides behind the flags / scan / compaction (and the barriers) of this one
        auto load_unit = [&](uint32_t t0) -> uint4 {
            const int64_t i0 = (int64_t)t0 + tid * 16 - mis;
            if (i0 + 16 > 0 && i0 < (int64_t)npay) return __ldg(reinterpret_cast<const uint4 *>(abase + t0 + tid * 16));
            return make_uint4(0, 0, 0, 0);
        };
        uint4 qn = load_unit(0);
        uint32_t tail_w = 0;      // last word of the tile before (lane 31's)

        for (uint32_t t0 = 0; !done; t0 += kUnstuffTile) {
            // this thread's 16 bytes: virtual indices i0 .. i0+15, i = (t0 + tid*16 + b) - mis
            const int64_t i0 = (int64_t)t0 + tid * 16 - mis;
            uint32_t wv[4] = { qn.x, qn.y, qn.z, qn.w };
            qn = load_unit(t0 + kUnstuffTile);
            uint32_t keep = 0, term = 0;
            // the byte in front of this thread's unit: the left neighbour's last one, for lane 0 the last lane's of the tile
            // before (a unit inside the payload always has a loaded unit on its left).  One-warp CTAs only.
            uint32_t left = 0;
            if (kUnstuffThreads == 32) {
                left = __shfl_up_sync(0xffffffffu, wv[3], 1);
                if (lane == 0) left = tail_w;
                tail_w = __shfl_sync(0xffffffffu, wv[3], 31);
            }
            if (sp5x && i0 >= 0 && i0 + 16 <= (int64_t)npay) {
                keep = 0xffffu;                   // literal bytes: nothing to drop, nothing terminates
            } else if (!sp5x && i0 >= 1 && i0 + 16 <= (int64_t)npay) {
                // ---- unit inside the payload: SIMD byte flags.  after_ff = the byte before is FF;
                // drop = after_ff and (00 or FF); terminator = after_ff and not (00, FF, RSTn)
                const uint32_t f0 = zero_bytes(~wv[0]), f1 = zero_bytes(~wv[1]), f2 = zero_bytes(~wv[2]), f3 = zero_bytes(~wv[3]);
                const uint32_t pf = kUnstuffThreads == 32 ? (left >= 0xff000000u ? 0x80000000u : 0u)
                                                          : (pay[i0 - 1] == 0xff ? 0x80000000u : 0u);
                const uint32_t a0 = __funnelshift_l(pf, f0, 8), a1 = __funnelshift_l(f0, f1, 8),
                               a2 = __funnelshift_l(f1, f2, 8), a3 = __funnelshift_l(f2, f3, 8);
                uint32_t dropb = 0;
                if (a0 | a1 | a2 | a3) {
                    const uint32_t n0 = zero_bytes(wv[0]) | f0, n1 = zero_bytes(wv[1]) | f1, n2 = zero_bytes(wv[2]) | f2,
                                   n3 = zero_bytes(wv[3]) | f3;
                    dropb = flag_bits(a0 & n0) | (flag_bits(a1 & n1) << 4) | (flag_bits(a2 & n2) << 8) | (flag_bits(a3 & n3) << 12);
                    // a byte behind an FF that is neither 00 nor FF: RSTn or a marker -- none in a sound AMV scan up to its
                    // appended EOI (which the byte-serial path below sees), so the RSTn flags are built only when there is one
                    const uint32_t m0 = a0 & ~n0, m1 = a1 & ~n1, m2 = a2 & ~n2, m3 = a3 & ~n3;
                    if (m0 | m1 | m2 | m3) {
                        const uint32_t r0 = zero_bytes((wv[0] & 0xf8f8f8f8u) ^ 0xd0d0d0d0u), r1 = zero_bytes((wv[1] & 0xf8f8f8f8u) ^ 0xd0d0d0d0u),
                                       r2 = zero_bytes((wv[2] & 0xf8f8f8f8u) ^ 0xd0d0d0d0u), r3 = zero_bytes((wv[3] & 0xf8f8f8f8u) ^ 0xd0d0d0d0u);
                        term = flag_bits(m0 & ~r0) | (flag_bits(m1 & ~r1) << 4) | (flag_bits(m2 & ~r2) << 8) | (flag_bits(m3 & ~r3) << 12);
                    }
                }
                keep = ~dropb & 0xffffu;
            } else if (i0 + 16 > 0 && i0 < (int64_t)V) {
                // ---- unit at an end of the payload: byte-serial, with the appended FF D9
                uint32_t prev = 0;
                if (i0 - 1 >= 0 && i0 - 1 < (int64_t)npay) prev = pay[i0 - 1];
                else if (i0 - 1 == (int64_t)npay) prev = 0xff;
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    const int64_t i = i0 + b;
                    uint32_t x = (wv[b >> 2] >> (8 * (b & 3))) & 0xff;
                    if (i == (int64_t)npay) x = 0xff;              // appended EOI
                    else if (i == (int64_t)npay + 1) x = 0xd9;
                    const bool valid = i >= 0 && i < (int64_t)V;
                    // the byte before the payload is the SOS header's 00; SP5X payload bytes are literal
                    const bool after_ff = prev == 0xff && i > 0 && !(sp5x && i <= (int64_t)npay);
                    const bool drop = after_ff && (x == 0x00 || x == 0xff);
                    const bool is_term = after_ff && !(x == 0x00 || x == 0xff || (x >= 0xd0 && x <= 0xd7));
                    if (valid && !drop) keep |= 1u << b;
                    if (valid && is_term) term |= 1u << b;
                    // patch the byte in place so the compaction below sees the virtual FF
                    if (i == (int64_t)npay) wv[b >> 2] |= 0xffu << (8 * (b & 3));
                    prev = x;
                }
            }
            // first terminator in this tile (only the last tile of a sound packet has one)
            if (cta_or(term != 0)) {
                if (tid == 0) s_first_term = 0xffffffffu;
                cta_sync();
                if (term) atomicMin(&s_first_term, (uint32_t)(tid * 16 + (__ffs(term) - 1)));
                cta_sync();
                const uint32_t ft = s_first_term;
                // keep only bytes strictly before the terminator
                const int rel = (int)ft - tid * 16;
                if (rel <= 0) keep = 0; else if (rel < 16) keep &= (1u << rel) - 1u;
                done = true;
                // the terminator must be the appended D9, anything earlier is a marker inside the scan
                if ((int64_t)t0 + ft - mis != (int64_t)V - 1) st |= AMV_ST_MARKER;
            } else if ((int64_t)t0 + kUnstuffTile - mis >= (int64_t)V) {
                done = true;   // unreachable (D9 always terminates), kept for safety
            }
            // CTA exclusive scan of kept counts
            const uint32_t cnt = __popc(keep);
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) warp_cnt[wid] = inc;
            cta_sync();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kUnstuffThreads / 32; w++) {
                const uint32_t c = warp_cnt[w];
                if (w < wid) wbase += c;
                total += c;
            }
            const uint32_t pos = carry + wbase + inc - cnt;
            if (cnt) {
                // close the unit up in registers: remove the holes below its highest kept byte, top one first
                uint64_t lo = (uint64_t)wv[0] | ((uint64_t)wv[1] << 32), hi = (uint64_t)wv[2] | ((uint64_t)wv[3] << 32);
                uint32_t holes = ~keep & ((2u << (31 - __clz(keep))) - 1u);
                while (holes) {
                    const int p = 31 - __clz(holes);
                    holes ^= 1u << p;
                    if (p < 8) {
                        const uint64_t m = (1ull << (8 * p)) - 1ull;
                        lo = (lo & m) | ((lo >> 8) & ~m) | (hi << 56);
                        hi >>= 8;
                    } else {
                        const uint64_t m = (1ull << (8 * (p - 8))) - 1ull;
                        hi = (hi & m) | ((hi >> 8) & ~m);
                    }
                }
                if (cnt <= 8) { hi = 0; if (cnt < 8) lo &= (1ull << (8 * cnt)) - 1ull; }
                else if (cnt < 16) hi &= (1ull << (8 * (cnt - 8))) - 1ull;
                // OR the (at most) five words it covers into the zeroed stage
                const uint32_t b0 = (uint32_t)lo, b1 = (uint32_t)(lo >> 32), b2 = (uint32_t)hi, b3 = (uint32_t)(hi >> 32);
                const uint32_t sh = (pos & 3u) * 8u;
                const uint32_t wa = stage_s + (pos & ~3u);
                const uint32_t o0 = b0 << sh, o1 = __funnelshift_l(b0, b1, sh), o2 = __funnelshift_l(b1, b2, sh),
                               o3 = __funnelshift_l(b2, b3, sh), o4 = __funnelshift_l(b3, 0u, sh);
                if (o0) red_or_shared(wa, o0);
                if (o1) red_or_shared(wa + 4, o1);
                if (o2) red_or_shared(wa + 8, o2);
                if (o3) red_or_shared(wa + 12, o3);
                if (o4) red_or_shared(wa + 16, o4);
            }
            cta_sync();
            const uint32_t have = carry + total;
            // the stage is zero past `have`, so the final flush carries its own zero padding
            const uint32_t flush = done ? (((have + 15u) & ~15u) + 16u) : (have & ~15u);
            // have <= tile + 15: a thread moves at most two units (its own, the first threads one more behind the tile) --
            // written out, the compiler's general loop cost a tenth of the kernel's instructions
            const uint32_t u0 = tid * 16, u1 = u0 + kUnstuffTile;
            {
                uint8_t *d = dst + written;
                if (u0 < flush) *reinterpret_cast<uint4 *>(d + u0) = lds128(stage_s + u0);
                if (u1 < flush) *reinterpret_cast<uint4 *>(d + u1) = lds128(stage_s + u1);
            }
            const uint32_t rem = done ? 0u : have - flush;
            uint8_t keepb = 0;
            if (tid < rem) keepb = stage[flush + tid];
            cta_sync();
            const uint32_t zend = ((have + 15u) & ~15u) + 16u;
            if (u0 < zend) sts128(stage_s + u0, make_uint4(0, 0, 0, 0));
            if (u1 < zend) sts128(stage_s + u1, make_uint4(0, 0, 0, 0));
            cta_sync();
            if (tid < rem) stage[tid] = keepb;
            carry = rem;
            written += done ? have : flush;       // at the end: the logical length
            cta_sync();
        }
        if (tid == 0) { scan_len[f] = written; status[f] = st; }
    }
}

// ------------------------------------------------------------------------------------------------
// shared-memory tables for the VLC kernels
// ------------------------------------------------------------------------------------------------
// Everything a scan needs besides its bits, in device memory: the fixed AMV / SP5X set, the amvlib
// flavour's (same codes, amvlib's quantisers and zigzag), or one built from a JPEG's own DQT / DHT.
struct DecTableSet {
    FlatVlcTables flat;          // k_vlc_sync, k_vlc_tokens
    uint32_t tz[2][64];          // zigzag position -> token fields | quantiser (DequantTables::tz / AmvlibDequantTables::tz)
    int q0[2];                   // DC quantisers of component 0 / components 1, 2
};
// which reference decoder the arithmetic follows: the ffmpeg fork's (sp5xdec/mjpegdec/simple_idct)
// or amvlib's (AmvJpeg.c) -- same bitstream, different quantisers, DC chain, zigzag and IDCT
// kFlavorJpeg: ffmpeg arithmetic with PER-FRAME quantisers (a plain JPEG's own DQT segment), see k_mjpeg_check
// kFlavorJpegDri: the same with a restart interval (RSTn markers in the scan)
enum { kFlavorFfmpeg = 0, kFlavorAmvlib = 1, kFlavorJpeg = 2, kFlavorJpegDri = 3 };
__device__ DecTableSet g_dec_sets[2];

// ------------------------------------------------------------------------------------------------
// k_vlc_sync: self-synchronising subsequence decode.  The P lanes of a frame (P = 2..32 consecutive
// lanes of a warp) each walk one P-th of the scan from a guessed state (the subsequence's first bit
// taken as a block boundary, block 0 of a macroblock), hand their exit state -- where the first block
// at or behind the subsequence's end starts, and which block of the macroblock it is -- to the right
// neighbour with __shfl_up, and walk again from the state they are handed until a __ballot shows that
// no entry state changed (JPEG's Huffman codes re-synchronise within a few symbols, so this is two
// walks for almost every lane).  A segmented shuffle scan then gives every lane its first block index
// and DC predictors.
//
// The walk is the flat symbol loop of k_vlc_tokens without the token side: one iteration = one symbol
// for every lane that is still inside its subsequence, bits through the same shared-memory ring with
// service points every kTokPeriod symbols, DC and AC through the same predicated code; per block only
// the block count, the next block's tables and the rotation of the three DC sums.
// ------------------------------------------------------------------------------------------------
constexpr int kTokThreads = 256;
constexpr int kTokWarps = kTokThreads / 32;
constexpr int kRingWords = 16;
constexpr int kTokPeriod = 4;         // symbols between two service points

struct SyncSmem {
    uint32_t ring[kTokWarps][kRingWords * 32];     // 2 KB per warp, 2 KB aligned
    uint2    bstate[8];                            // per block-in-MCU: DC table, AC table | component change on entering it
    uint32_t lut[kFlatMaxEntries];
};
constexpr size_t kSyncSmemBytes = sizeof(SyncSmem) + 2048;

__global__ void __launch_bounds__(kTokThreads)
k_vlc_sync(const uint8_t *__restrict__ scratch, const uint64_t *__restrict__ slot_off,
           const uint32_t *__restrict__ scan_len, int n, int log2p, LaneStart *__restrict__ starts,
           uint32_t *__restrict__ rounds_out /* optional: max rounds per warp, for profiling */, int flavor,
           const DecTableSet *__restrict__ tabs, const uint8_t *__restrict__ qtab /* kFlavorJpeg: 2 x 64 quantisers per frame */,
           int nl, int nc /* blocks per MCU: luma, one chroma component */) {
    AMV_EXTERN_SHARED(uint8_t, sync_smem_raw, 16);
    const uint32_t raw_s = smem_addr(sync_smem_raw);
    SyncSmem &S = *reinterpret_cast<SyncSmem *>(sync_smem_raw + (((raw_s + 2047u) & ~2047u) - raw_s));
    const int nlut = tabs->flat.count;
    for (int i = threadIdx.x; i < nlut; i += blockDim.x) S.lut[i] = tabs->flat.e[i];
    const uint32_t lut_s = smem_addr(S.lut);
    const uint32_t nbm = (uint32_t)(nl + 2 * nc);               // blocks per MCU (<= 8): nl luma, nc Cb, nc Cr
    if (threadIdx.x < nbm) {
        const uint32_t bq = threadIdx.x, tq = bq >= (uint32_t)nl ? 1 : 0;
        const bool enters = bq == 0 || bq == (uint32_t)nl || bq == (uint32_t)(nl + nc);   // first block of a component
        uint2 bs;
        bs.x = ((lut_s + (uint32_t)tabs->flat.base[tq] * 4u) << 8) | (32u - kFlatDcBits);
        bs.y = ((lut_s + (uint32_t)tabs->flat.base[2 + tq] * 4u) << 8) | (32u - kFlatAcBits) | (enters ? 0x80u : 0u);
        S.bstate[bq] = bs;
    }
    __syncthreads();
    const int P = 1 << log2p;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (int)(gt >> log2p);
    const int p = (int)(gt & (P - 1));
    const bool active = f < n;
    const uint32_t ring_s = smem_addr(&S.ring[wid][lane]);      // word w of this lane: | (w & 15) << 7
    const uint32_t bstate_s = smem_addr(&S.bstate[0]);

    const uint32_t *words = nullptr;
    uint32_t cap_words = 0, end_bit = 0, total_bits = 0, L = 0;
    if (active) {
        const uint32_t U = scan_len[f];
        words = reinterpret_cast<const uint32_t *>(scratch + slot_off[f]);
        cap_words = (((U + 15u) & ~15u) + kSlotPad) >> 2;       // inside the zero-padded slot (U <= packet size)
        total_bits = U * 8u;
        L = (((total_bits + P - 1) >> log2p) + 31u) & ~31u;
        const uint64_t e = (uint64_t)(p + 1) * L;
        end_bit = e < total_bits ? (uint32_t)e : total_bits;
    }
    uint32_t start_bit = active ? (uint32_t)min((uint64_t)p * L, (uint64_t)total_bits) : 0u, start_phase = 0;
    LaneExit ex = { 0, 0, 0, { 0, 0, 0 } };
    auto load_group = [&](uint32_t w) -> uint4 {        // words [w, w+4) of the scan, zeros past the slot
        if (w + 4 <= cap_words) return __ldg(reinterpret_cast<const uint4 *>(words + w));
        return make_uint4(0, 0, 0, 0);
    };
    auto ring_put = [&](uint32_t w, const uint4 &q) {   // w is a multiple of 4
        const uint32_t a = ring_s | ((w << 7) & 0x780u);
        sts32(a, bswap32(q.x)); sts32(a + 128, bswap32(q.y)); sts32(a + 256, bswap32(q.z)); sts32(a + 384, bswap32(q.w));
    };

    uint32_t rounds = 0;
    bool need = active;                     // this lane walks in the coming round
    for (int r = 0; r <= P; r++) {
        if (r > 0) {
            // entry state = left neighbour's exit state (lane 0 of a frame starts the scan)
            uint32_t nbit = __shfl_up_sync(0xffffffffu, ex.bitpos, 1);
            uint32_t nph = __shfl_up_sync(0xffffffffu, ex.phase, 1);
            if (p == 0) { nbit = 0; nph = 0; }
            need = active && (nbit != start_bit || nph != start_phase);
            if (!__ballot_sync(0xffffffffu, need)) break;
            if (need) { start_bit = nbit; start_phase = nph; }
        }
        rounds++;
        // ---- walk the subsequence from (start_bit, start_phase): blocks that START before end_bit
        uint32_t bp = start_bit;
        uint32_t wr = (start_bit >> 5) & ~3u;           // next word the ring receives; ring = words [wr-16, wr)
        uint4 pend = make_uint4(0, 0, 0, 0);
        bool on = need && start_bit < end_bit;
        uint32_t kb = 0;                                // zigzag position of the last symbol + 1; 0: the DC comes next
        uint32_t b = start_phase, nb = 0;
        int dA = 0, dB = 0, dC = 0;                     // sums of DC differences; dA: the current block's component
        uint32_t desc = 0, acd = 0;
        {
            const uint2 bs = S.bstate[b];
            desc = bs.x; acd = bs.y;
        }
        if (on) {
            ring_put(wr, load_group(wr));
            ring_put(wr + 4, load_group(wr + 4));
            wr += 8;
            pend = load_group(wr);
        }
        while (__any_sync(0xffffffffu, on)) {
            if (on) {       // service point: ring top-up
                const uint32_t rd = bp >> 5;
                if ((int)(wr + 4 - rd) <= kRingWords) { ring_put(wr, pend); wr += 4; pend = load_group(wr); }
            }
#pragma unroll
            for (int u = 0; u < kTokPeriod; u++) {
                const uint32_t x = bp << 2;
                const uint32_t wa = lds32(ring_s | (x & 0x780u)), wc = lds32(ring_s | ((x + 128u) & 0x780u));
                const uint32_t hi = __funnelshift_l(wc, wa, bp);
                uint32_t e = lds32((desc >> 8) + (__funnelshift_r(hi, 0u, desc) << 2));
                if ((e & 31u) == 0) {
                    if (!(e & kFlatBad)) {
                        const uint32_t fb = 32u - (desc & 31u), sb = e >> 24;
                        e = lds32(lut_s + ((((e >> 8) & 0xffffu) + ((hi << fb) >> (32u - sb))) << 2));
                    }
                    if ((e & 31u) == 0)     // no such code: a bad DC reads as difference 0, a bad AC ends the block
                        e = 1u | (1u << 8) | (kb ? (kFlatAdvEob << 23) : ((1u << 23) | (1u << 31)));
                }
                const uint32_t top = hi << (e & 31u);
                if (on) { bp += __byte_perm(e, 0, 0x4441); kb += (e >> 23) & 0xffu; }
                const int sg = (int)(~top) >> 31;
                const int lvl = (int)((__funnelshift_l(top ^ (uint32_t)sg, 0u, e >> 16) ^ (uint32_t)sg) - (uint32_t)sg);
                const bool isdc = kb == 1u;
                dA += (isdc && on) ? lvl : 0;
                const bool nz = (int)e < 0;             // the symbol carries a value (for AC symbols: size != 0)
                const bool endp = on && kb >= 64u && (nz || kb >= 128u);
                b = endp ? (b + 1u == nbm ? 0u : b + 1u) : b;
                nb += endp ? 1u : 0u;
                const uint2 bs = S.bstate[b];            // after the step: the next block's tables at a block end
                const bool rot = endp && (bs.y & 0x80u);
                const int tA = rot ? dB : dA, tB = rot ? dC : dB, tC = rot ? dA : dC;
                dA = tA; dB = tB; dC = tC;
                desc = endp ? bs.x : ((isdc && on) ? acd : desc);
                acd = endp ? bs.y : acd;
                kb = endp ? 0u : kb;
                on = on && !(endp && bp >= end_bit);
            }
        }
        if (need) {
            // the sums sit rotated to the component of the block that comes next
            const int c = b < (uint32_t)nl ? 0 : (b < (uint32_t)(nl + nc) ? 1 : 2);
            ex.bitpos = bp; ex.phase = b; ex.nblocks = nb;
            ex.dc[0] = c == 0 ? dA : (c == 1 ? dC : dB);
            ex.dc[1] = c == 0 ? dB : (c == 1 ? dA : dC);
            ex.dc[2] = c == 0 ? dC : (c == 1 ? dB : dA);
        }
    }
    // segmented (width P) exclusive scans: first block index and DC difference sums
    uint32_t nb_inc = ex.nblocks;
    int d0 = ex.dc[0], d1 = ex.dc[1], d2 = ex.dc[2];
    for (int d = 1; d < P; d <<= 1) {
        const uint32_t tn = __shfl_up_sync(0xffffffffu, nb_inc, d);
        const int t0 = __shfl_up_sync(0xffffffffu, d0, d), t1 = __shfl_up_sync(0xffffffffu, d1, d),
                  t2 = __shfl_up_sync(0xffffffffu, d2, d);
        if (p >= d) { nb_inc += tn; d0 += t0; d1 += t1; d2 += t2; }
    }
    if (active) {
        LaneStart s;
        s.bitpos = start_bit;
        s.first_block = nb_inc - ex.nblocks;
        s.nblocks = ex.nblocks;
        if (flavor == kFlavorAmvlib) {                         // quantised units, 16-bit chain from 0 (AmvJpeg.c:1194-1196)
            s.pred[0] = sext16(d0 - ex.dc[0]); s.pred[1] = sext16(d1 - ex.dc[1]); s.pred[2] = sext16(d2 - ex.dc[2]);
        } else {
            const int q0l = flavor == kFlavorJpeg ? (int)qtab[(size_t)f * 128] : tabs->q0[0];
            const int q0c = flavor == kFlavorJpeg ? (int)qtab[(size_t)f * 128 + 64] : tabs->q0[1];
            s.pred[0] = 1024 + q0l * (d0 - ex.dc[0]);          // last_dc starts at 1024 (mjpegdec.c:805-806)
            s.pred[1] = 1024 + q0c * (d1 - ex.dc[1]);
            s.pred[2] = 1024 + q0c * (d2 - ex.dc[2]);
        }
        starts[gt] = s;
    }
    if (rounds_out && lane == 0) atomicMax(rounds_out, rounds);
}

// ------------------------------------------------------------------------------------------------
// k_vlc_sync_lean: the same self-synchronising rounds with the walk of k_vlc_tokens_lean (further down: AC-only symbol
// body; a lane that ends a block parks until the next service point, where the parked lanes count their block, rotate
// the DC sums, switch tables and read the next block's DC together).  For the fixed AMV / SP5X tables; the flat kernel
// above stays for custom tables and the amvlib flavour.
// ------------------------------------------------------------------------------------------------
template <int NW>
struct SyncLeanSmemT {
    uint32_t ring[NW][kRingWords * 32];            // 2 KB per warp, 2 KB aligned
    uint4    bstate[8];                            // per block-in-MCU: DC table, AC table, component change on entering it, next index
    uint32_t lut[kFlatMaxEntries];                 // AC entries with the token flag cleared: the top 9 bits are the advance
};
template <int NW> constexpr size_t sync_lean_smem_bytes() { return sizeof(SyncLeanSmemT<NW>) + 2048; }
struct SyncCheckpoint {     // while a walk runs: what it had at the boundary; afterwards: what it added from there to its exit
    uint32_t bitpos, phase, nblocks;
    int dc[3];
    uint32_t exit_bitpos, exit_phase;   // where that walk left the subsequence
    bool valid, fresh;
};

template <int NW>           // warps per CTA, chosen per launch (pick_vlc_warps)
__global__ void __launch_bounds__(NW * 32)
k_vlc_sync_lean(const uint8_t *__restrict__ scratch, const uint64_t *__restrict__ slot_off,
                const uint32_t *__restrict__ scan_len, int n, int log2p, LaneStart *__restrict__ starts,
                uint32_t *__restrict__ rounds_out /* optional: max rounds per warp, for profiling */,
                const DecTableSet *__restrict__ tabs, int nl, int nc /* blocks per MCU: luma, one chroma component */) {
    AMV_EXTERN_SHARED(uint8_t, synclean_smem_raw, 16);
    const uint32_t raw_s = smem_addr(synclean_smem_raw);
    SyncLeanSmemT<NW> &S = *reinterpret_cast<SyncLeanSmemT<NW> *>(synclean_smem_raw + (((raw_s + 2047u) & ~2047u) - raw_s));
    const int nlut = tabs->flat.count, ac0 = tabs->flat.base[2];
    for (int i = threadIdx.x; i < nlut; i += blockDim.x) {
        const uint32_t e = tabs->flat.e[i];
        S.lut[i] = (i >= ac0 && (e & 31u)) ? e & 0x7fffffffu : e;
    }
    const uint32_t lut_s = smem_addr(S.lut);
    const uint32_t nbm = (uint32_t)(nl + 2 * nc);               // blocks per MCU (<= 8): nl luma, nc Cb, nc Cr
    if (threadIdx.x < nbm) {
        const uint32_t bq = threadIdx.x, tq = bq >= (uint32_t)nl ? 1 : 0;
        const bool enters = bq == 0 || bq == (uint32_t)nl || bq == (uint32_t)(nl + nc);   // first block of a component
        uint4 bs;
        bs.x = lut_s + (uint32_t)tabs->flat.base[tq] * 4u;
        bs.y = lut_s + (uint32_t)tabs->flat.base[2 + tq] * 4u;
        bs.z = enters ? 1u : 0u;
        bs.w = bq + 1u == nbm ? 0u : bq + 1u;
        S.bstate[bq] = bs;
    }
    __syncthreads();
    const int P = 1 << log2p;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (int)(gt >> log2p);
    const int p = (int)(gt & (P - 1));
    const bool active = f < n;
    const uint32_t ring_s = smem_addr(&S.ring[wid][lane]);      // word w of this lane: | (w & 15) << 7
    const uint32_t bstate_s = smem_addr(&S.bstate[0]);

    const uint32_t *words = nullptr;
    uint32_t cap_words = 0, end_bit = 0, total_bits = 0, L = 0;
    if (active) {
        const uint32_t U = scan_len[f];
        words = reinterpret_cast<const uint32_t *>(scratch + slot_off[f]);
        cap_words = (((U + 15u) & ~15u) + kSlotPad) >> 2;       // inside the zero-padded slot (U <= packet size)
        total_bits = U * 8u;
        L = (((total_bits + P - 1) >> log2p) + 31u) & ~31u;
        const uint64_t e = (uint64_t)(p + 1) * L;
        end_bit = e < total_bits ? (uint32_t)e : total_bits;
    }
    uint32_t start_bit = active ? (uint32_t)min((uint64_t)p * L, (uint64_t)total_bits) : 0u, start_phase = 0;
    LaneExit ex = { 0, 0, 0, { 0, 0, 0 } };
    auto load_group = [&](uint32_t w) -> uint4 {        // words [w, w+4) of the scan, zeros past the slot
        if (w + 4 <= cap_words) return __ldg(reinterpret_cast<const uint4 *>(words + w));
        return make_uint4(0, 0, 0, 0);
    };
    auto ring_put = [&](uint32_t w, const uint4 &q) {   // w is a multiple of 4
        const uint32_t a = ring_s | ((w << 7) & 0x780u);
        sts32(a, bswap32(q.x)); sts32(a + 128, bswap32(q.y)); sts32(a + 256, bswap32(q.z)); sts32(a + 384, bswap32(q.w));
    };

    // Checkpoints: the first block boundary a walk reaches at or behind a quarter and five eighths of the subsequence is
    // remembered as (bit, block-in-MCU) with the blocks and DC sums the walk collects from there to its exit.  A later
    // walk that arrives at the same boundary would repeat the earlier one bit for bit from there on, so it stops and
    // adds the remembered remainder.  The second walk of a lane differs from its first only up to the point where the
    // guessed state had re-synchronised (a few blocks in, profiles/README.md), so it ends at the first checkpoint.
    SyncCheckpoint cp0 = { 0, 0, 0, { 0, 0, 0 }, 0, 0, false, false }, cp1 = cp0;
    const uint32_t cp0_pos = start_bit + (L >> 2), cp1_pos = start_bit + (L >> 1) + (L >> 3);
    auto by_component = [&](uint32_t blk, int a, int bb, int c3, int *out) {       // the rotated sums, per component
        const int c = blk < (uint32_t)nl ? 0 : (blk < (uint32_t)(nl + nc) ? 1 : 2);
        out[0] = c == 0 ? a : (c == 1 ? c3 : bb);
        out[1] = c == 0 ? bb : (c == 1 ? a : c3);
        out[2] = c == 0 ? c3 : (c == 1 ? bb : a);
    };

    uint32_t rounds = 0;
    bool need = active;                     // this lane walks in the coming round
    for (int r = 0; r <= P; r++) {
        if (r > 0) {
            // entry state = left neighbour's exit state (lane 0 of a frame starts the scan)
            uint32_t nbit = __shfl_up_sync(0xffffffffu, ex.bitpos, 1);
            uint32_t nph = __shfl_up_sync(0xffffffffu, ex.phase, 1);
            if (p == 0) { nbit = 0; nph = 0; }
            need = active && (nbit != start_bit || nph != start_phase);
            if (!__ballot_sync(0xffffffffu, need)) break;
            if (need) { start_bit = nbit; start_phase = nph; }
        }
        rounds++;
        // ---- walk the subsequence from (start_bit, start_phase): blocks that START before end_bit
        uint32_t bp = start_bit;
        uint32_t wr = (start_bit >> 5) & ~3u;           // next word the ring receives; ring = words [wr-16, wr)
        uint4 pend = make_uint4(0, 0, 0, 0);
        bool live = need && start_bit < end_bit;        // the lane has a block to walk
        bool joined = false;                            // the walk met a checkpoint of an earlier one: ex is complete
        int cpn = 0;                                    // the checkpoint this walk passes next
        bool on = false, ended = false;                 // inside a block (its DC is read) / parked at its end
        uint32_t kb = 0;                                // zigzag position of the last symbol + 1
        uint32_t b = start_phase, nb = 0;
        int dA = 0, dB = 0, dC = 0;                     // sums of DC differences; dA: the current block's component
        uint32_t dct_s, act_s, bnext;
        {
            const uint4 bs = S.bstate[b];
            dct_s = bs.x; act_s = bs.y; bnext = bs.w;
        }
        auto window = [&]() -> uint32_t {               // the 32 bits at bp
            const uint32_t x = bp << 2;
            const uint32_t wa = lds32(ring_s | (x & 0x780u)), wc = lds32(ring_s | ((x + 128u) & 0x780u));
            return __funnelshift_l(wc, wa, bp);
        };
        if (live) {
            ring_put(wr, load_group(wr));
            ring_put(wr + 4, load_group(wr + 4));
            wr += 8;
            pend = load_group(wr);
        }
        while (__any_sync(0xffffffffu, live)) {
            // ---- service point: block ends of the parked lanes, ring top-up, the DC of every lane that starts a block
            if (ended) {
                nb++;
                b = bnext;
                const uint4 bs = lds128(bstate_s + b * 16u);
                if (bs.z) { const int t = dA; dA = dB; dB = dC; dC = t; }       // the coming block opens another component
                dct_s = bs.x; act_s = bs.y; bnext = bs.w;
                ended = false;
                live = bp < end_bit;                    // a block that starts at or behind the boundary is the neighbour's
                if (live && cpn < 2 && bp >= (cpn == 0 ? cp0_pos : cp1_pos)) {
                    SyncCheckpoint &cp = cpn == 0 ? cp0 : cp1;
                    int sums[3];
                    by_component(b, dA, dB, dC, sums);
                    if (cp.valid && cp.bitpos == bp && cp.phase == b) {
                        // the earlier walk went on from this very state: its exit is this walk's exit
                        ex.bitpos = cp.exit_bitpos; ex.phase = cp.exit_phase;
                        ex.nblocks = nb + cp.nblocks;
                        ex.dc[0] = sums[0] + cp.dc[0]; ex.dc[1] = sums[1] + cp.dc[1]; ex.dc[2] = sums[2] + cp.dc[2];
                        joined = true;
                        live = false;
                    } else {
                        cp.bitpos = bp; cp.phase = b; cp.nblocks = nb;
                        cp.dc[0] = sums[0]; cp.dc[1] = sums[1]; cp.dc[2] = sums[2];
                        cp.valid = false; cp.fresh = true;
                    }
                    cpn++;
                }
            }
            if (live) {
                const uint32_t rd = bp >> 5;
                if ((int)(wr + 4 - rd) <= kRingWords) { ring_put(wr, pend); wr += 4; pend = load_group(wr); }
                if (!on) {      // the block's DC (mjpeg_decode_dc, mjpegdec.c:358-373): only its difference counts here
                    const uint32_t hi = window();
                    uint32_t e = lds32(dct_s + ((hi >> (32 - kFlatDcBits)) << 2));
                    if ((e & 31u) == 0) {
                        if (!(e & kFlatBad)) e = lds32(lut_s + ((((e >> 8) & 0xffffu) + ((hi << kFlatDcBits) >> (32u - (e >> 24)))) << 2));
                        if ((e & 31u) == 0) e = 1u | (1u << 8);                 // no such code: reads as difference 0
                    }
                    const uint32_t top = __funnelshift_l(0u, hi, e);
                    const int sg = (int)(~top) >> 31;
                    dA += (int)((__funnelshift_l(top ^ (uint32_t)sg, 0u, e >> 16) ^ (uint32_t)sg) - (uint32_t)sg);
                    bp += (e >> 8) & 0xffu;
                    kb = 1;
                    on = true;
                }
            }
            // ---- kTokPeriod AC symbols (decode_block, mjpegdec.c:391-428), predicated; parked lanes keep their state
#pragma unroll
            for (int u = 0; u < kTokPeriod; u++) {
                const uint32_t hi = window();
                uint32_t e = lds32(act_s + ((hi >> (32 - kFlatAcBits)) << 2));
                if ((e & 31u) == 0) {           // (keeping parked lanes out of this side path, as the token pass does, measured slower here: 9.16 -> 9.62 ms)
                    if (!(e & kFlatBad)) e = lds32(lut_s + ((((e >> 8) & 0xffffu) + ((hi << kFlatAcBits) >> (32u - (e >> 24)))) << 2)) & 0x7fffffffu;
                    if ((e & 31u) == 0) e = 1u | (1u << 8) | (kFlatAdvEob << 23);        // no such code: ends the block
                }
                if (on) { bp += (e >> 8) & 0xffu; kb += e >> 23; }      // (a PRMT here, as in the token pass, measured slower: 9.15 -> 9.58 ms)
                // a coefficient at position 63 or behind it, or EOB (advance 128), ends the block; ZRL alone does not
                const bool nz = ((e >> 16) & 31u) != 0;
                const bool fin = (kb & 0xc0u) != 0 && (nz || (kb & 0x80u) != 0);
                if (on && fin) ended = true;
                on = on && !fin;
            }
        }
        if (need) {
            if (!joined) {      // the sums sit rotated to the component of the block that comes next
                ex.bitpos = bp; ex.phase = b; ex.nblocks = nb;
                by_component(b, dA, dB, dC, ex.dc);
            }
            // what this walk noted at its checkpoints becomes the remainder from there to the exit
            if (cp0.fresh) {
                cp0.nblocks = ex.nblocks - cp0.nblocks;
                cp0.dc[0] = ex.dc[0] - cp0.dc[0]; cp0.dc[1] = ex.dc[1] - cp0.dc[1]; cp0.dc[2] = ex.dc[2] - cp0.dc[2];
                cp0.exit_bitpos = ex.bitpos; cp0.exit_phase = ex.phase;
                cp0.fresh = false; cp0.valid = true;
            }
            if (cp1.fresh) {
                cp1.nblocks = ex.nblocks - cp1.nblocks;
                cp1.dc[0] = ex.dc[0] - cp1.dc[0]; cp1.dc[1] = ex.dc[1] - cp1.dc[1]; cp1.dc[2] = ex.dc[2] - cp1.dc[2];
                cp1.exit_bitpos = ex.bitpos; cp1.exit_phase = ex.phase;
                cp1.fresh = false; cp1.valid = true;
            }
        }
    }
    // segmented (width P) exclusive scans: first block index and DC difference sums
    uint32_t nb_inc = ex.nblocks;
    int d0 = ex.dc[0], d1 = ex.dc[1], d2 = ex.dc[2];
    for (int d = 1; d < P; d <<= 1) {
        const uint32_t tn = __shfl_up_sync(0xffffffffu, nb_inc, d);
        const int t0 = __shfl_up_sync(0xffffffffu, d0, d), t1 = __shfl_up_sync(0xffffffffu, d1, d),
                  t2 = __shfl_up_sync(0xffffffffu, d2, d);
        if (p >= d) { nb_inc += tn; d0 += t0; d1 += t1; d2 += t2; }
    }
    if (active) {
        LaneStart s;
        s.bitpos = start_bit;
        s.first_block = nb_inc - ex.nblocks;
        s.nblocks = ex.nblocks;
        s.pred[0] = 1024 + tabs->q0[0] * (d0 - ex.dc[0]);          // last_dc starts at 1024 (mjpegdec.c:805-806)
        s.pred[1] = 1024 + tabs->q0[1] * (d1 - ex.dc[1]);
        s.pred[2] = 1024 + tabs->q0[1] * (d2 - ex.dc[2]);
        starts[gt] = s;
    }
    if (rounds_out && lane == 0) atomicMax(rounds_out, rounds);
}

// ------------------------------------------------------------------------------------------------
// k_vlc_tokens: Huffman -> fixed-width tokens.  Every lane re-walks its (now exactly delimited)
// subsequence and writes, per block, a DC token (absolute dequantised DC) followed by one 32-bit
// token per non-zero AC coefficient -- already de-zigzagged and dequantised, so the consumer only
// scatters them -- plus the block's (count, offset) entry.  A block of b bits yields at most b/2
// tokens (a coefficient costs at least 3 bits, DC + EOB at least 4), so the region is 16 bytes per
// scan byte and a lane that starts at bit s writes from token s/2 (+4 per lane of slack for 16-byte
// alignment) without ever meeting its neighbour.
//
// The symbol loop is the serial heart of the decoder.  It is FLAT: one iteration = one Huffman
// symbol of whatever kind, for every lane that still has blocks left, so lanes never wait for the
// longest block of the warp (the block-structured first version ran with 20 of 32 lanes active):
//  * DC and AC symbols share the code path.  A block starts with k = -1; every table entry says by
//    how much k advances (DC 1, coefficient run+1, ZRL 16, EOB 128), the dequant table at k = 0
//    holds the DC quantiser, the DC predictor is added under a select, and "k >= 63" ends the block.
//    The only per-block work is a short tail (offset entry, next block's tables from a 6-entry
//    shared table, predictor rotation Y -> Cb -> Cr -> Y).
//  * bits come from a per-lane 16-word ring in shared memory ([word][lane], conflict-free, 2 KB
//    aligned per warp so a slot address is one LOP3) and are looked at through one funnel shift of
//    two ring words; the ring is topped up with one 128-bit global load per lane at warp-uniform
//    service points every kTokPeriod symbols (<= 124 bits), the load issued at one service point
//    being stored at the next, so its latency never sits inside the symbol code and the window
//    needs no "ring ran dry" branch;
//  * one 12-bit (AC) / 10-bit (DC) first-level lookup per symbol, byte-field entries; longer codes
//    (2 % of the symbols with 10 bits, < 0.3 % with 12) take a second lookup;
//  * tokens are staged in an 8-slot per-lane shared ring and leave as 16-byte stores at the
//    service points.
// ------------------------------------------------------------------------------------------------

constexpr int kTokStage = 8;          // staged tokens per lane
// flag on the dequant entries of zigzag positions past 63, in a bit that reaches neither the product nor the token
constexpr uint32_t kTzErr = 1u << 15, kTzErrAmvlib = 1u << 16;

struct TokSmem {
    uint32_t ring[kTokWarps][kRingWords * 32];     // 2 KB per warp, 2 KB aligned
    uint32_t tstage[kTokWarps][kTokStage * 32];    // 1 KB per warp, 1 KB aligned
    uint32_t tz[2][128];                           // kb = zigzag position + 1 -> (consumer column byte offset << 16) | quantiser; 512 B aligned
    uint4    bstate[8];                            // per block-in-MCU: DC table, AC table, dequant table, next entry | predictor rotation
    uint32_t lut[kFlatMaxEntries];
};
constexpr size_t kTokSmemBytes = sizeof(TokSmem) + 2048;

template <int FLAVOR>
__global__ void __launch_bounds__(kTokThreads)
k_vlc_tokens(const uint8_t *__restrict__ scratch, const uint64_t *__restrict__ slot_off,
             const uint32_t *__restrict__ scan_len, const uint32_t *__restrict__ pkt_size, int n, int log2p,
             const LaneStart *__restrict__ starts, int nblk, uint32_t *__restrict__ tokens,
             uint32_t *__restrict__ blk_off, int32_t *__restrict__ status, const DecTableSet *__restrict__ tabs,
             const uint8_t *__restrict__ qtab /* kFlavorJpeg: 2 x 64 quantisers per frame */,
             int nl, int nc /* blocks per MCU: luma, one chroma component */, int restart /* kFlavorJpegDri: MCUs per interval */) {
    constexpr bool kPerFrameQ = FLAVOR == kFlavorJpeg || FLAVOR == kFlavorJpegDri;
    // a restart can add 23 bits (alignment + the marker) to what the symbols of a period consume: service twice as often
    constexpr int kPeriod = FLAVOR == kFlavorJpegDri ? 2 : kTokPeriod;
    AMV_EXTERN_SHARED(uint8_t, tok_smem_raw, 16);
    const uint32_t raw_s = smem_addr(tok_smem_raw);
    TokSmem &S = *reinterpret_cast<TokSmem *>(tok_smem_raw + (((raw_s + 2047u) & ~2047u) - raw_s));
    const int nlut = tabs->flat.count;
    for (int i = threadIdx.x; i < nlut; i += blockDim.x) S.lut[i] = tabs->flat.e[i];
    {   // dequant table indexed by kb; positions past 64 (only broken streams get there) alias the last one
        const int c = threadIdx.x >> 7, kb = threadIdx.x & 127, kk = kb == 0 ? 0 : (kb > 64 ? 63 : kb - 1);
        const uint32_t z = tabs->tz[c][kk];
        S.tz[c][kb] = z | (kb > 64 ? (FLAVOR == kFlavorAmvlib ? kTzErrAmvlib : kTzErr) : 0u);
    }
    const uint32_t lut_s = smem_addr(S.lut);
    const uint32_t bstate_s = smem_addr(&S.bstate[0]);
    const uint32_t nbm = (uint32_t)(nl + 2 * nc);               // blocks per MCU (<= 8): nl luma, nc Cb, nc Cr
    if (threadIdx.x < nbm) {
        // block b of the MCU (Y.. Cb.. Cr..; Y Y Y Y Cb Cr for AMV): its tables, the entry of the block after it, and
        // whether the component changes on entering it
        const uint32_t bq = threadIdx.x, tq = bq >= (uint32_t)nl ? 1 : 0;
        const bool enters = bq == 0 || bq == (uint32_t)nl || bq == (uint32_t)(nl + nc);
        uint4 bs;
        bs.x = ((lut_s + (uint32_t)tabs->flat.base[tq] * 4u) << 8) | (32u - kFlatDcBits);
        bs.y = ((lut_s + (uint32_t)tabs->flat.base[2 + tq] * 4u) << 8) | (32u - kFlatAcBits) | (enters ? 0x80u : 0u);
        bs.z = smem_addr(&S.tz[tq][0]);
        bs.w = bstate_s + (bq + 1u == nbm ? 0u : bq + 1u) * 16u;
        S.bstate[bq] = bs;
    }
    __syncthreads();
    const int P = 1 << log2p;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (int)(gt >> log2p);
    const int p = (int)(gt & (P - 1));
    const uint32_t ring_s = smem_addr(&S.ring[wid][lane]);      // word w of this lane: | (w & 15) << 7
    const uint32_t tst_s = smem_addr(&S.tstage[wid][lane]);     // staged token c of this lane: | (c & 7) << 7
    const uint32_t tz_s = smem_addr(&S.tz[0][0]);
    const uint8_t *qf = kPerFrameQ ? qtab + (size_t)(f < n ? f : 0) * 128 : nullptr;   // this frame's quantisers
    int rst_count = restart;                                     // MCUs until the next restart (mjpegdec.c:682-683,726-732)
    constexpr int kPred0 = FLAVOR == kFlavorAmvlib ? 0 : 1024;   // last_dc (mjpegdec.c:805-806) / ycoef.. (AmvJpeg.c:1510)

    // ---- lane set-up (inactive lanes keep count = 0 and never enter the symbol code)
    uint32_t count = 0, first = 0, bit = 0, U = 0, st = 0;
    int pred0 = kPred0, pred1 = kPred0, pred2 = kPred0;
    const uint32_t *words = nullptr;
    uint32_t cap_words = 0;
    uint64_t so = 0;
    if (f < n) {
        U = scan_len[f];
        so = slot_off[f];
        words = reinterpret_cast<const uint32_t *>(scratch + so);
        cap_words = (((pkt_size[f] + 15u) & ~15u) + kSlotPad) >> 2;
        if (U) {       // U == 0: rejected by k_unstuff (status already says why)
            count = (uint32_t)nblk;
            if (log2p) {
                const LaneStart s = starts[gt];
                first = s.first_block; bit = s.bitpos; count = s.nblocks;
                pred0 = s.pred[0]; pred1 = s.pred[1]; pred2 = s.pred[2];
                if (first >= (uint32_t)nblk) count = 0;
                else if (first + count > (uint32_t)nblk) count = nblk - first;
                if (p == P - 1 && first + s.nblocks < (uint32_t)nblk) {
                    // the scan ran out before the picture was complete: keep decoding (zeros) like a
                    // sequential reader would, and say so
                    count = nblk - first;
                    st |= AMV_ST_OVERRUN;
                }
            }
        }
    }
    // token output: the lane's tokens are numbered from 0 and live at tok_first + number inside the
    // frame's region (16 B per scan byte); they leave in groups of four (16-byte stores)
    const uint32_t tok_cap = cap_words * 16u;
    uint32_t tok_first = ((bit >> 1) + 4u * (uint32_t)p + 3u) & ~3u;
    uint32_t *tok_frame = tokens + so * 4;
    uint32_t flushed = 0, blk0 = 0;
    uint32_t *boff = blk_off + (uint64_t)(f < n ? f : 0) * nblk + first;       // the lane's (count, offset) entries
    uint32_t bi = 0;                                                           // blocks finished
    // Bounds are checked once per block, not per store: a block yields at most 66 tokens, so a lane whose
    // next token lies within kTokBlockRoom of the end of its frame's region (only streams that are already
    // broken get there: a sound block of b bits has at most b/2 tokens) is flagged and parked on the
    // region's tail, where its remaining blocks overwrite each other -- memory-safe, and every block still
    // gets a valid (count, offset) entry for the consumer.
    constexpr uint32_t kTokBlockRoom = 80;
    const uint32_t tpark = (tok_cap - kTokBlockRoom) & ~3u;    // tok_cap >= 640 (kSlotPad)

    // ---- bit source
    uint32_t bp = bit;                      // next unread bit of the scan
    uint32_t wr = (bit >> 5) & ~3u;         // next word the ring receives (16-byte groups); ring = words [wr-16, wr)
    uint4 pend = make_uint4(0, 0, 0, 0);    // the group at wr, requested at the previous service point
    auto load_group = [&](uint32_t w) -> uint4 {        // words [w, w+4) of the scan, zeros past the slot
        if (w + 4 <= cap_words) return __ldg(reinterpret_cast<const uint4 *>(words + w));
        return make_uint4(0, 0, 0, 0);
    };
    auto ring_put = [&](uint32_t w, const uint4 &q) {   // w is a multiple of 4
        const uint32_t a = ring_s | ((w << 7) & 0x780u);
        sts32(a, bswap32(q.x)); sts32(a + 128, bswap32(q.y)); sts32(a + 256, bswap32(q.z)); sts32(a + 384, bswap32(q.w));
    };
    auto flush_group = [&]() {
        const uint32_t a = tst_s | ((flushed << 7) & 0x200u);
        const uint4 q = make_uint4(lds32(a), lds32(a + 128), lds32(a + 256), lds32(a + 384));
        *reinterpret_cast<uint4 *>(tok_frame + (tok_first + flushed)) = q;
        flushed += 4;
    };

    bool on = count > 0;
    // kt = tokens emitted by this lane << 8 | kb, kb = zigzag position of the last symbol + 1 (0: the DC comes
    // next).  One add per symbol moves both: table entries hold (token? << 8 | advance) in their top 9 bits.
    // kb <= 127 before a symbol (else the block has ended) and the advance is <= 128: no carry into the count.
    uint32_t kt = 0;
    uint32_t desc = 0, acd = 0, tzp = 0, nxt_s = bstate_s;
    uint32_t zacc = 0;                                           // dequant entries seen at block ends (error flag)
    int predA = pred0, predB = pred1, predC = pred2;             // predA: the current block's component
    if (on) {
        ring_put(wr, load_group(wr));
        ring_put(wr + 4, load_group(wr + 4));
        wr += 8;
        pend = load_group(wr);
    }
    {   // every lane gets valid tables, also the ones without work: they run the symbol code with frozen state
        const uint32_t b = first % nbm;
        const uint4 bs = S.bstate[b];
        desc = bs.x; acd = bs.y; tzp = bs.z; nxt_s = bs.w;
        if (b >= (uint32_t)nl && b < (uint32_t)(nl + nc)) { predA = pred1; predB = pred2; predC = pred0; }
        if (b >= (uint32_t)(nl + nc)) { predA = pred2; predB = pred0; predC = pred1; }
    }

    while (__any_sync(0xffffffffu, on)) {
        // ---- service point (warp-uniform): ring top-up, token flush, room check
        if (on) {
            const uint32_t rd = bp >> 5;                                       // words below this one are dead
            if ((int)(wr + 4 - rd) <= kRingWords) { ring_put(wr, pend); wr += 4; pend = load_group(wr); }
            if (FLAVOR == kFlavorAmvlib) { while ((kt >> 8) - flushed >= 4) flush_group(); }
            else if ((kt >> 8) - flushed >= 4) flush_group();
            if (tok_first + (kt >> 8) > tpark) { st |= AMV_ST_OVERRUN; tok_first = (tpark - (kt >> 8)) & ~3u; }   // see kTokBlockRoom
        }
        // ---- kTokPeriod symbols, straight-line: everything but the long-code lookup is predicated, so the
        // four symbol bodies interleave freely.  A lane that has finished keeps executing with its state frozen
        // (its loads stay inside the tables and its own ring, its stores are predicated off).
#pragma unroll
        for (int u = 0; u < kPeriod; u++) {
            // ---- the 32 bits at bp
            const uint32_t x = bp << 2;
            const uint32_t wa = lds32(ring_s | (x & 0x780u)), wc = lds32(ring_s | ((x + 128u) & 0x780u));
            const uint32_t hi = __funnelshift_l(wc, wa, bp);
            // ---- symbol (mjpeg_decode_dc mjpegdec.c:358-373 / decode_block :391-428)
            uint32_t e = lds32((desc >> 8) + (__funnelshift_r(hi, 0u, desc) << 2));
            if ((e & 31u) == 0) {
                if (!(e & kFlatBad)) {
                    const uint32_t fb = 32u - (desc & 31u), sb = e >> 24;
                    e = lds32(lut_s + ((((e >> 8) & 0xffffu) + ((hi << fb) >> (32u - sb))) << 2));
                }
                if ((e & 31u) == 0) {       // no such code: a bad DC reads as difference 0, a bad AC ends the block
                    if (on) st |= AMV_ST_BADCODE;
                    e = 1u | (1u << 8) | ((kt & 0xffu) ? (kFlatAdvEob << 23) : ((1u << 23) | (1u << 31)));
                }
            }
            const uint32_t slot = tst_s | ((kt >> 1) & 0x380u);                // staging slot of the token, if one comes
            const uint32_t top = hi << (e & 31u);
            if (on) { bp += __byte_perm(e, 0, 0x4441); kt += e >> 23; }
            const int sg = (int)(~top) >> 31;                                  // get_xbits: -1 if the first bit is 0
            // the size bits under the code, sign-extended; the shift count is the entry's bits [20:16] (wrap mode ignores the rest)
            const int lvl = (int)((__funnelshift_l(top ^ (uint32_t)sg, 0u, e >> 16) ^ (uint32_t)sg) - (uint32_t)sg);
            const uint32_t z = lds32(tzp | ((kt << 2) & 0x1fcu));
            const uint32_t kb = kt & 0xffu;
            const bool isdc = kb == 1u;
            uint32_t tok;
            if (FLAVOR == kFlavorAmvlib) {
                // the DC chain runs in quantised units in a 16-bit variable (AmvJpeg.c:1194-1196); the product
                // with the quantiser is a full int (IQtIZzBlock :1041-1046)
                const int dcv = sext16(predA + lvl);
                const int c = isdc ? dcv : lvl;
                predA = isdc ? dcv : predA;
                tok = (z & 0xfc000000u) | ((uint32_t)(c * (int)(z & 0xffu)) & 0x03ffffffu);
                sts32(slot, tok);
                if ((z & (kAmvlibTokSkip | kAmvlibTokDup)) && (int)e < 0 && on) {      // the zigzag typo (AmvJpeg.c:131-141)
                    if (z & kAmvlibTokSkip) kt -= 256u;                        // coefficient 31 is never read
                    else { sts32(tst_s | ((kt >> 1) & 0x380u), ((z << 16) & 0xfc000000u) | (tok & 0x03ffffffu)); kt += 256u; }
                }
            } else {
                // level * quant_matrix[j] as int16 (:420,428); block[0] = (int16)(last_dc += diff * q0) (:387-389)
                int q = (int)(z & 0xffu);
                if (kPerFrameQ) q = (int)__ldg(qf + (((tzp - tz_s) >> 3) + ((kb - 1u) & 63u)));   // tables are 512 B apart
                const int prod = lvl * q;
                const int val = prod + (isdc ? predA : 0);
                predA = isdc ? val : predA;
                tok = __byte_perm((uint32_t)val, z, 0x7610);
                sts32(slot, tok);
            }
            // ---- end of block: EOB, coefficient 63, or a coefficient index > 63 ("error count", :423-424; those
            // positions carry kTzErr in the dequant table)
            const bool nz = lvl != 0;
            const bool endp = on && kb >= 64u && (nz || kb >= 128u);
            const uint32_t tc = kt >> 8;
            // (AC token count << 24) + index of the DC token; the sum is < 2^24 whatever tok_first wrapped to
            if (endp) boff[bi] = ((tc - blk0 - 1u) << kTokCountShift) + (tok_first + blk0);
            const uint4 bs = lds128(nxt_s);
            const bool next_is_first = nxt_s == bstate_s;                      // the coming block opens an MCU
            zacc |= (endp && nz) ? z : 0u;
            bi += endp ? 1u : 0u;
            const bool rot = endp && (bs.y & 0x80u);
            const int nA = rot ? predB : predA, nB = rot ? predC : predB, nC = rot ? predA : predC;
            predA = nA; predB = nB; predC = nC;
            desc = endp ? bs.x : (isdc ? acd : desc);
            acd = endp ? bs.y : acd;
            tzp = endp ? bs.z : tzp;
            nxt_s = endp ? bs.w : nxt_s;
            kt = endp ? (kt & ~0xffu) : kt;
            blk0 = endp ? tc : blk0;
            on = on && bi != count;
            if (FLAVOR == kFlavorJpegDri) {
                // end of an MCU (the coming block is block 0): count it; at the end of a restart interval the reader
                // skips to the next byte boundary and over the RSTn marker (un-stuffing keeps FF Dn), and the DC
                // predictors start over (mjpegdec.c:726-732).  Nothing follows the last MCU.
                const bool mcu_end = endp && next_is_first;
                rst_count -= mcu_end ? 1 : 0;
                if (mcu_end && rst_count == 0) {
                    rst_count = restart;
                    if (on) { bp = ((bp + 7u) & ~7u) + 16u; predA = 1024; predB = 1024; predC = 1024; }
                }
            }
        }
    }
    if (zacc & (FLAVOR == kFlavorAmvlib ? kTzErrAmvlib : kTzErr)) st |= AMV_ST_COEFIDX;
    // flush what is still staged (unused upper tokens of the last group are don't-care, the group is ours alone)
    while (flushed < (kt >> 8)) flush_group();
    // a lane that owns no block just passes through; otherwise it must end inside the scan
    if (count && bp > U * 8u) st |= AMV_ST_OVERRUN;
    if (st && f < n) atomicOr(&status[f], (int32_t)st);
}

// ------------------------------------------------------------------------------------------------
// k_vlc_tokens16: the token pass for the fixed AMV / SP5X tables, restructured around what the flat loop above
// spends its instructions on.  There, everything a block BOUNDARY needs (offset entry, next block's tables, predictor
// rotation, the DC's own arithmetic, counters) is predicated code inside every symbol -- well over a third of its
// ~78 instructions -- because 32 lanes end their blocks at 32 different symbols.  Here a lane that ends a block parks
// until the next service point (every kTokPeriod symbols, warp-uniform), where the parked lanes do their block ends
// AND decode the next block's DC together under one branch: the symbol body is AC-only and shrinks to what an AC
// symbol needs, at the price of ~1.5 idle symbol slots per block.
// Tokens are 16 bit: the DC token is the block's absolute dequantised DC, every other token is
// (run << 12) | (level & 0xfff) -- the Huffman symbol in fixed width, ZRL included (run 15, level 0) -- and the
// consumer (k_idct16) accumulates positions and multiplies by the quantiser.  The fixed tables code at most 10
// magnitude bits, so a level always fits its 12-bit field; custom tables (plain JPEG) keep the 32-bit pass.
// Region: 4 tokens per slot byte (8 B per scan byte, half of the 32-bit pass), lane p of a frame starts at token
// (bit / 2 + 16 p) rounded up to 8, so groups of eight leave as aligned 16-byte stores.
// Bits: per service period a lane consumes at most one DC (11 + 11 bits) and kTokPeriod AC symbols (16 + 10 each),
// 126 bits, and receives up to 128.
// ------------------------------------------------------------------------------------------------
constexpr int kTok16Stage = 16;       // staged 16-bit tokens per lane
// CTA size of the lean token pass, chosen per launch (tok_lean_warps): three CTAs of 8 warps fit an SM (72 KB of shared memory
// each) or two of 11 warps (82 KB).  When a launch has between 16 and 22 warps per SM -- 100 000 frames on one lane each are
// 3 125 warps -- the 8-warp CTAs put 24 warps on some SMs and 16 on others, the 11-warp CTAs at most 22 on any: 8.39 -> 8.13 ms.
// Outside that window the 8-warp CTAs stay (8 192 frames of 1280x720 on 8 lanes each: 9.1 ms against 12.2 ms with 11 warps).
template <int NW>
struct Tok16SmemT {
    uint32_t ring[NW][kRingWords * 32];         // 2 KB per warp, 2 KB aligned
    // 1 KB per warp, 1 KB aligned.  16-bit tokens: halfword c*32 + lane = staged token c (of 16) of the lane;
    // 32-bit tokens: word c*32 + lane = staged token c (of 8)
    uint32_t tstage[NW][8 * 32];
    uint32_t tz[2][128];                                   // 32-bit tokens: kb = zigzag position + 1 -> (consumer column byte offset << 16) | quantiser; 512 B aligned
    uint4    bstate[8];                                    // per block-in-MCU: DC table, AC table, DC quantiser | dequant table, component change | next index
    uint32_t lut[kFlatMaxEntries];                         // AC entries rewritten: advance - 1, bit 31 = yields a token
};
template <int NW> constexpr size_t tok16_smem_bytes() { return sizeof(Tok16SmemT<NW>) + 2048; }

// T16: 16-bit (run << 12 | level) tokens for k_idct16; else 32-bit (column offset << 16 | level x quantiser) tokens, DC included
// (offset 0, absolute DC), for k_idct.  Measured (ncu, 100 000 frames): 6.36 G warp instructions against 7.71 G -- the table
// look-up, the product, the predicated store and twice the flushes -- on a pass that is bound by the ALU pipe (77 % / 67 %
// active): 9.2 ms against 12.4 ms, which the 32-bit tokens' cheaper consumer (8.0 against 9.3 ms) does not win back.
template <bool T16, int NW>
__global__ void __launch_bounds__(NW * 32)
k_vlc_tokens_lean(const uint8_t *__restrict__ scratch, const uint64_t *__restrict__ slot_off,
                  const uint32_t *__restrict__ scan_len, const uint32_t *__restrict__ pkt_size, int n, int log2p,
                  const LaneStart *__restrict__ starts, int nblk, void *__restrict__ tokens_v,
                  uint32_t *__restrict__ blk_off, int32_t *__restrict__ status, const DecTableSet *__restrict__ tabs,
                  int nl, int nc /* blocks per MCU: luma, one chroma component */) {
    AMV_EXTERN_SHARED(uint8_t, tok16_smem_raw, 16);
    const uint32_t raw_s = smem_addr(tok16_smem_raw);
    Tok16SmemT<NW> &S = *reinterpret_cast<Tok16SmemT<NW> *>(tok16_smem_raw + (((raw_s + 2047u) & ~2047u) - raw_s));
    const int nlut = tabs->flat.count, ac0 = tabs->flat.base[2];
    for (int i = threadIdx.x; i < nlut; i += blockDim.x) {
        uint32_t e = tabs->flat.e[i];
        // everything from the first AC table on (the AC first levels and their second levels) feeds the symbol loop
        // (16-bit tokens carry ZRL as a token, 32-bit tokens only what has a value: bit 31 as it is)
        // (advance - 1, not the advance: for a coefficient that is its run, which the 16-bit token takes from these bits as they are)
        if (i >= ac0 && (e & 31u)) e = T16 ? ((e - (1u << 23)) & 0x7fffffffu) | ((e & kFlatTok16) << 24) : e - (1u << 23);
        S.lut[i] = e;
    }
    if (!T16 && threadIdx.x < 256) {     // dequant table indexed by kb; positions past 64 (only broken streams get there) alias the last one
        const int c = threadIdx.x >> 7, kb = threadIdx.x & 127, kk = kb == 0 ? 0 : (kb > 64 ? 63 : kb - 1);
        S.tz[c][kb] = tabs->tz[c][kk];
    }
    const uint32_t lut_s = smem_addr(S.lut);
    const uint32_t nbm = (uint32_t)(nl + 2 * nc);               // blocks per MCU (<= 8): nl luma, nc Cb, nc Cr
    if (threadIdx.x < nbm) {
        const uint32_t bq = threadIdx.x, tq = bq >= (uint32_t)nl ? 1 : 0;
        const bool enters = bq == 0 || bq == (uint32_t)nl || bq == (uint32_t)(nl + nc);   // first block of a component
        uint4 bs;
        bs.x = lut_s + (uint32_t)tabs->flat.base[tq] * 4u;
        bs.y = lut_s + (uint32_t)tabs->flat.base[2 + tq] * 4u;
        bs.z = (uint32_t)tabs->q0[tq] | (smem_addr(&S.tz[tq][0]) << 8);
        bs.w = (enters ? 0x80u : 0u) | (bq + 1u == nbm ? 0u : bq + 1u);
        S.bstate[bq] = bs;
    }
    __syncthreads();
    const int P = 1 << log2p;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (int)(gt >> log2p);
    const int p = (int)(gt & (P - 1));
    const uint32_t ring_s = smem_addr(&S.ring[wid][lane]);      // word w of this lane: | (w & 15) << 7
    // staged token c of this lane: 16-bit | (c & 15) << 6, 32-bit | (c & 7) << 7
    const uint32_t tst_s = T16 ? smem_addr(reinterpret_cast<uint16_t *>(&S.tstage[wid][0]) + lane) : smem_addr(&S.tstage[wid][lane]);
    const uint32_t bstate_s = smem_addr(&S.bstate[0]);
    constexpr uint32_t kGroup = T16 ? 8u : 4u;                  // tokens per 16-byte store

    // ---- lane set-up (inactive lanes keep count = 0 and never enter the symbol code), as in k_vlc_tokens
    uint32_t count = 0, first = 0, bit = 0, U = 0, st = 0;
    int pred0 = 1024, pred1 = 1024, pred2 = 1024;               // last_dc (mjpegdec.c:805-806)
    const uint32_t *words = nullptr;
    uint32_t cap_words = 0;
    uint64_t so = 0;
    if (f < n) {
        U = scan_len[f];
        so = slot_off[f];
        words = reinterpret_cast<const uint32_t *>(scratch + so);
        cap_words = (((pkt_size[f] + 15u) & ~15u) + kSlotPad) >> 2;
        if (U) {       // U == 0: rejected by k_unstuff (status already says why)
            count = (uint32_t)nblk;
            if (log2p) {
                const LaneStart s = starts[gt];
                first = s.first_block; bit = s.bitpos; count = s.nblocks;
                pred0 = s.pred[0]; pred1 = s.pred[1]; pred2 = s.pred[2];
                if (first >= (uint32_t)nblk) count = 0;
                else if (first + count > (uint32_t)nblk) count = nblk - first;
                if (p == P - 1 && first + s.nblocks < (uint32_t)nblk) {
                    // the scan ran out before the picture was complete: keep decoding (zeros) like a
                    // sequential reader would, and say so
                    count = nblk - first;
                    st |= AMV_ST_OVERRUN;
                }
            }
        }
    }
    const uint32_t tok_cap = cap_words * 16u;                   // 4 tokens per slot byte
    // 16 tokens of slack per lane: the lane's first group is aligned up (<= 7) and its last group is written whole (<= 7
    // don't-care tokens), and neither may reach the next lane's first group however short the subsequence is
    uint32_t tok_first = ((bit >> 1) + 2u * kGroup * (uint32_t)p + kGroup - 1u) & ~(kGroup - 1u);
    uint16_t *tok_frame16 = static_cast<uint16_t *>(tokens_v) + so * 4;
    uint32_t *tok_frame32 = static_cast<uint32_t *>(tokens_v) + so * 4;
    uint32_t flushed = 0, blk0 = 0;
    uint32_t *boff = blk_off + (uint64_t)(f < n ? f : 0) * nblk + first;       // the lane's (count, offset) entries
    uint32_t bi = 0;                                                           // blocks finished
    // room is checked once per service period (at most kTokPeriod + 1 new tokens, 7 more staged): a lane whose next token
    // lies within kTokBlockRoom of the end of its frame's region (only streams that are already broken get there) is
    // flagged and parked on the region's tail, where its remaining blocks overwrite each other -- memory-safe (the
    // consumer's reads of such a block stay inside the token workspace), and every block still gets a (count, offset) entry
    constexpr uint32_t kTokBlockRoom = 48;
    const uint32_t tpark = (tok_cap - kTokBlockRoom) & ~(kGroup - 1u);    // tok_cap >= 640 (kSlotPad)

    // ---- bit source (as in k_vlc_tokens)
    uint32_t bp = bit;                      // next unread bit of the scan
    uint32_t wr = (bit >> 5) & ~3u;         // next word the ring receives (16-byte groups); ring = words [wr-16, wr)
    uint4 pend = make_uint4(0, 0, 0, 0);    // the group at wr, requested at the previous service point
    auto load_group = [&](uint32_t w) -> uint4 {        // words [w, w+4) of the scan, zeros past the slot
        if (w + 4 <= cap_words) return __ldg(reinterpret_cast<const uint4 *>(words + w));
        return make_uint4(0, 0, 0, 0);
    };
    auto ring_put = [&](uint32_t w, const uint4 &q) {   // w is a multiple of 4
        const uint32_t a = ring_s | ((w << 7) & 0x780u);
        sts32(a, bswap32(q.x)); sts32(a + 128, bswap32(q.y)); sts32(a + 256, bswap32(q.z)); sts32(a + 384, bswap32(q.w));
    };
    auto window = [&]() -> uint32_t {                   // the 32 bits at bp
        const uint32_t x = bp << 2;
        const uint32_t wa = lds32(ring_s | (x & 0x780u)), wc = lds32(ring_s | ((x + 128u) & 0x780u));
        return __funnelshift_l(wc, wa, bp);
    };
    // kt = tokens emitted by this lane << 8 | kb, kb = zigzag position of the last symbol + 1.  One add per symbol moves
    // both: the (rewritten) AC entries hold (token? << 8 | advance - 1) in their top 9 bits.  kb < 128 before a symbol
    // (else the block has ended) and the advance is <= 128: no carry into the count.
    uint32_t kt = 0;
    auto stage_addr = [&](uint32_t c) -> uint32_t { return T16 ? tst_s | ((c << 6) & 0x3c0u) : tst_s | ((c << 7) & 0x380u); };
    auto flush_group = [&]() {              // staged tokens [flushed, flushed + kGroup)
        uint32_t w[4];
        // flushed is a multiple of the group size: one address, the group's slots at fixed offsets from it
        const uint32_t fb = stage_addr(flushed);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (T16) {
                const uint32_t lo = lds_u16(fb + 128u * j), hi16 = lds_u16(fb + 128u * j + 64u);
                w[j] = lo | (hi16 << 16);
            } else w[j] = lds32(fb + 128u * j);
        }
        if (T16) *reinterpret_cast<uint4 *>(tok_frame16 + (tok_first + flushed)) = make_uint4(w[0], w[1], w[2], w[3]);
        else     *reinterpret_cast<uint4 *>(tok_frame32 + (tok_first + flushed)) = make_uint4(w[0], w[1], w[2], w[3]);
        flushed += kGroup;
    };

    bool live = count > 0;                  // the lane has blocks left
    bool on = false;                        // ... and is inside one (its DC is read; not parked at a block end)
    bool ended = false;                     // parked at a block end
    int endlvl = 0;                         // ... whose last symbol carried this level (non-zero: a coefficient)
    uint32_t dct_s = 0, act_s = 0, bnext = 0, tz_s = 0;
    int q0 = 0;
    uint32_t cerr = 0;                      // a coefficient index ran past 63 ("error count", mjpegdec.c:423-424)
    int predA = pred0, predB = pred1, predC = pred2;             // predA: the current block's component
    if (live) {
        ring_put(wr, load_group(wr));
        ring_put(wr + 4, load_group(wr + 4));
        wr += 8;
        pend = load_group(wr);
    }
    {   // every lane gets valid tables, also the ones without work: they run the symbol code with frozen state
        const uint32_t b = first % nbm;
        const uint4 bs = S.bstate[b];
        dct_s = bs.x; act_s = bs.y; q0 = (int)(bs.z & 0xffu); tz_s = bs.z >> 8; bnext = bs.w & 0x7fu;
        if (b >= (uint32_t)nl && b < (uint32_t)(nl + nc)) { predA = pred1; predB = pred2; predC = pred0; }
        if (b >= (uint32_t)(nl + nc)) { predA = pred2; predB = pred0; predC = pred1; }
    }

    while (__any_sync(0xffffffffu, live)) {
        // ---- service point (warp-uniform): block ends of the parked lanes, ring top-up, token flush, room check,
        // and the DC of every lane that starts a block
        if (ended) {
            const uint32_t ntok = kt >> 8;
            // (symbol token count << 24) + index of the DC token; the sum is < 2^24 whatever tok_first wrapped to
            boff[bi] = ((ntok - blk0 - 1u) << kTokCountShift) + (tok_first + blk0);
            bi++;
            blk0 = ntok;
            if (endlvl != 0 && (kt & 0xffu) != 64u) cerr = 1;            // the last coefficient sat past position 63
            const uint4 bs = lds128(bstate_s + bnext * 16u);
            if (bs.w & 0x80u) { const int t = predA; predA = predB; predB = predC; predC = t; }    // Y -> Cb -> Cr -> Y
            dct_s = bs.x; act_s = bs.y; q0 = (int)(bs.z & 0xffu); tz_s = bs.z >> 8; bnext = bs.w & 0x7fu;
            ended = false;
            live = bi != count;
        }
        if (live) {
            const uint32_t rd = bp >> 5;                                       // words below this one are dead
            if ((int)(wr + 4 - rd) <= kRingWords) { ring_put(wr, pend); wr += 4; pend = load_group(wr); }
            if (T16) { if ((kt >> 8) - flushed >= kGroup) flush_group(); }
            else { while ((kt >> 8) - flushed >= kGroup) flush_group(); }       // at most 3 + 1 + kTokPeriod staged: twice at most
            if (tok_first + (kt >> 8) > tpark) { st |= AMV_ST_OVERRUN; tok_first = (tpark - (kt >> 8)) & ~(kGroup - 1u); }   // see kTokBlockRoom
            if (!on) {
                // ---- the block's DC (mjpeg_decode_dc, mjpegdec.c:358-373): block[0] = (int16)(last_dc += diff * q0) (:387-389)
                const uint32_t hi = window();
                uint32_t e = lds32(dct_s + ((hi >> (32 - kFlatDcBits)) << 2));
                if ((e & 31u) == 0) {
                    if (!(e & kFlatBad)) e = lds32(lut_s + ((((e >> 8) & 0xffffu) + ((hi << kFlatDcBits) >> (32u - (e >> 24)))) << 2));
                    if ((e & 31u) == 0) { st |= AMV_ST_BADCODE; e = 1u | (1u << 8); }        // no such code: reads as difference 0
                }
                const uint32_t top = __funnelshift_l(0u, hi, e);               // hi << code length
                const int sg = (int)(~top) >> 31;                              // get_xbits: -1 if the first bit is 0
                const int diff = (int)((__funnelshift_l(top ^ (uint32_t)sg, 0u, e >> 16) ^ (uint32_t)sg) - (uint32_t)sg);
                bp += byte1(e);
                predA += diff * q0;
                if (T16) sts16(stage_addr(kt >> 8), (uint32_t)predA);
                else     sts32(stage_addr(kt >> 8), (uint32_t)predA & 0xffffu);                 // column offset 0
                kt = (kt & ~0xffu) + 0x101u;                                   // one token, position 0 done
                on = true;
            }
        }
        // ---- kTokPeriod AC symbols, straight-line and predicated (decode_block, mjpegdec.c:391-428): a lane that is parked
        // or has finished keeps executing with its state frozen (its loads stay inside the tables and its own ring, its
        // token lands in a dead staging slot)
#pragma unroll
        for (int u = 0; u < kTokPeriod; u++) {
            const uint32_t hi = window();
            uint32_t e = lds32(act_s + ((hi >> (32 - kFlatAcBits)) << 2));
            // (a parked lane looks at the next block's DC code through the AC table and would take this side path -- the long
            // codes -- in most iterations of the warp: ncu counted it with 2.3 threads per instruction at 5 % of the kernel's
            // instructions.  Its entry is never used, so it stays out.)
            if ((e & 31u) == 0 && on) {
                if (!(e & kFlatBad)) e = lds32(lut_s + ((((e >> 8) & 0xffffu) + ((hi << kFlatAcBits) >> (32u - (e >> 24)))) << 2));
                if ((e & 31u) == 0) {       // no such code: ends the block
                    st |= AMV_ST_BADCODE;
                    e = 1u | (1u << 8) | ((kFlatAdvEob - 1u) << 23);
                }
            }
            const uint32_t top = __funnelshift_l(0u, hi, e);                   // hi << code length
            const int sg = (int)(~top) >> 31;                                  // get_xbits: -1 if the first bit is 0
            // the size bits under the code, sign-extended; the shift count is the entry's bits [20:16] (wrap mode ignores the rest)
            const int lvl = (int)((__funnelshift_l(top ^ (uint32_t)sg, 0u, e >> 16) ^ (uint32_t)sg) - (uint32_t)sg);
            if (T16) {
                // the symbol in fixed width: run = advance - 1 (15 for ZRL) over the level
                // (one bit-select: the halfword store drops whatever the level carries above bit 15)
                sts16(stage_addr(kt >> 8), bitselect(e >> 11, (uint32_t)lvl, 0xf000u));
                if (on) { bp += byte1(e); kt += (e >> 23) + 1u; }
            } else {
                // level * quant_matrix[j] as int16 over the consumer's column offset (:420,428); the token goes where the count
                // stood before the symbol, and only if the symbol has a value (a ring of eight has no dead slot to spare)
                const uint32_t slot = stage_addr(kt >> 8);
                const bool emit = on && (int)e < 0;
                if (on) { bp += byte1(e); kt += (e >> 23) + 1u; }
                const uint32_t z = lds32(tz_s | ((kt << 2) & 0x1fcu));
                if (emit) sts32(slot, __byte_perm((uint32_t)(lvl * (int)(z & 0xffu)), z, 0x7610));
            }
            // ---- end of block: EOB (advance 128), coefficient 63, or a coefficient index > 63
            const bool nz = lvl != 0;
            const bool fin = (kt & 0xc0u) != 0 && (nz || (kt & 0x80u) != 0);
            if (on && fin) { ended = true; endlvl = lvl; }
            on = on && !fin;
        }
    }
    if (cerr) st |= AMV_ST_COEFIDX;
    // flush what is still staged (unused upper tokens of the last group are don't-care, the group is ours alone)
    while (flushed < (kt >> 8)) flush_group();
    // a lane that owns no block just passes through; otherwise it must end inside the scan
    if (count && bp > U * 8u) st |= AMV_ST_OVERRUN;
    if (st && f < n) atomicOr(&status[f], (int32_t)st);
}

// ------------------------------------------------------------------------------------------------
// k_idct: one thread per 8x8 block, blocks enumerated in PLANE raster order so that the 32 lanes
// of a warp own 32 horizontally adjacent blocks: every pixel-row store of the warp is one
// contiguous 256-byte run.  Tokens -> dequantised coefficients in a conflict-free shared-memory
// column -> simple_idct in registers -> bottom-up store (mjpegdec.c:672-677,710-716).
// ------------------------------------------------------------------------------------------------
constexpr int kIdctThreads = 128;

template <bool FAST>
__global__ void __launch_bounds__(kIdctThreads)
k_idct(const uint32_t *__restrict__ tokens, const uint32_t *__restrict__ blk_off, const uint64_t *__restrict__ slot_off,
       const uint32_t *__restrict__ scan_len, int n, Geom g, uint8_t *__restrict__ py, uint8_t *__restrict__ pu,
       uint8_t *__restrict__ pv, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
       uint32_t mg_nblk, uint32_t mg_rowl, uint32_t mg_rowc /* div_magic of nblk and of the luma / chroma blocks per row */) {
    __shared__ uint32_t tile[kIdctThreads / 32][32 * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t iu;        // divisions by launch constants: one high multiply and a correction each (div_by_magic)
    const int f = (gt >> 32) == 0 ? (int)div_by_magic((uint32_t)gt, (uint32_t)g.nblk, mg_nblk, iu) : (int)(gt / g.nblk);
    if (f >= n) return;
    if (scan_len[f] == 0) return;
    const int i = (gt >> 32) == 0 ? (int)iu : (int)(gt - (int64_t)f * g.nblk);
    // blocks are enumerated in plane raster order (all of Y, then Cb, then Cr); blk = the block's index in
    // bitstream order: MCU by MCU, inside an MCU component by component, v x h blocks in raster order
    const int nluma = g.nl * g.mbw * g.mbh, nchroma = g.nc * g.mbw * g.mbh;
    int comp, bx, by, blk;
    if (i < nluma) {
        comp = 0;
        const int rowb = g.mbw << g.llh;
        uint32_t rem;
        by = (int)div_by_magic((uint32_t)i, (uint32_t)rowb, mg_rowl, rem); bx = (int)rem;
        blk = ((by >> g.llv) * g.mbw + (bx >> g.llh)) * g.nb + ((by & ((1 << g.llv) - 1)) << g.llh) + (bx & ((1 << g.llh) - 1));
    } else {
        const int j = i - nluma;
        comp = j < nchroma ? 1 : 2;
        const int jj = comp == 1 ? j : j - nchroma;
        const int rowb = g.mbw << g.lch;
        uint32_t rem;
        by = (int)div_by_magic((uint32_t)jj, (uint32_t)rowb, mg_rowc, rem); bx = (int)rem;
        blk = ((by >> g.lcv) * g.mbw + (bx >> g.lch)) * g.nb + g.nl + (comp - 1) * g.nc +
              ((by & ((1 << g.lcv) - 1)) << g.lch) + (bx & ((1 << g.lch) - 1));
    }
    const uint32_t bo = blk_off[(uint64_t)f * g.nblk + blk];
    const uint32_t *tf = tokens + slot_off[f] * 4;                         // the frame's token region (16-byte aligned)
    const uint32_t first = bo & ((1u << kTokCountShift) - 1u), last = first + (bo >> kTokCountShift);   // DC token .. last AC token
    uint32_t *slot = &tile[wid][lane];
    const uint32_t slot_s = smem_addr(slot);
#pragma unroll
    for (int k = 0; k < 32; k++) slot[k * 32] = 0;
    // the tokens are already (column offset, value): scatter them.  They are fetched as aligned groups of four
    // (one 128-bit load instead of four dependent 32-bit ones), one group ahead in flight; the tokens of a group
    // that belong to the neighbouring blocks are skipped.  The look-ahead stays inside the region's slack.
    uint32_t gi = first & ~3u;
    uint4 q = __ldg(reinterpret_cast<const uint4 *>(tf + gi));
    for (; gi <= last; gi += 4) {
        const uint4 nq = __ldg(reinterpret_cast<const uint4 *>(tf + gi + 4));
        const uint32_t tk[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (gi + j >= first && gi + j <= last)
                sts16(slot_s + (tk[j] >> 16), tk[j]);
        q = nq;
    }

    uint32_t c[32], o[16];
#pragma unroll
    for (int k = 0; k < 32; k++) c[k] = slot[k * 32];
    // Smooth content (chroma planes almost always) has nothing below the second coefficient row: when that holds
    // for the whole warp, take the transform specialised for it (same results, a quarter of the arithmetic).
    uint32_t lower = 0;
#pragma unroll
    for (int k = 8; k < 32; k++) lower |= c[k];
    if (__all_sync(__activemask(), lower == 0)) idct_put_block<2>(c, o);
    else idct_put_block<8>(c, o);

    uint8_t *pl = comp == 0 ? py + (uint64_t)f * fs_y : (comp == 1 ? pu : pv) + (uint64_t)f * fs_c;
    const int ls = comp ? ls_c : ls_y;
    const int vw = comp ? g.cw : g.w, vh = comp ? g.ch : g.h, r0 = comp ? g.c0 : g.y0;
    const int x0 = bx * 8, y0 = by * 8;
#pragma unroll
    for (int yy = 0; yy < 8; yy++) {
        const int row = g.flip ? r0 - (y0 + yy) : y0 + yy;     // AMV pictures are stored bottom-up (mjpegdec.c:672-677), SP5X is not
        if (row < 0 || row >= vh) continue;
        uint8_t *d = pl + (int64_t)row * ls + x0;
        if (FAST) {
            *reinterpret_cast<uint2 *>(d) = make_uint2(o[2 * yy], o[2 * yy + 1]);
        } else {
#pragma unroll
            for (int xx = 0; xx < 8; xx++)
                if (x0 + xx < vw) d[xx] = (uint8_t)(o[2 * yy + (xx >> 2)] >> (8 * (xx & 3)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_idct16: k_idct for the 16-bit tokens of k_vlc_tokens16.  Same thread-per-block layout in plane raster order and
// the same transform; the scatter loop walks the block's symbols -- position = runs so far + symbols so far, value =
// level x quantiser truncated to int16 (decode_block, mjpegdec.c:420,428) -- eight tokens per 128-bit load.
// ------------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(kIdctThreads, 8)     // 8 CTAs per SM (64 registers, 92 bytes of spills) with the multiply-add chain transform: 7.94 ms; 6 / 7 CTAs 8.11, 9 (56 registers) 8.26, 10 (48) 9.47
k_idct16(const uint16_t *__restrict__ tokens, const uint32_t *__restrict__ blk_off, const uint64_t *__restrict__ slot_off,
         const uint32_t *__restrict__ scan_len, int n, Geom g, const DecTableSet *__restrict__ tabs, uint8_t *__restrict__ py,
         uint8_t *__restrict__ pu, uint8_t *__restrict__ pv, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
         uint32_t mg_nblk, uint32_t mg_rowl, uint32_t mg_rowc /* div_magic of nblk and of the luma / chroma blocks per row */) {
    __shared__ uint32_t tile[kIdctThreads / 32][32 * 32];
    // zigzag position -> (column byte offset << 16) | quantiser.  The scatter loop looks a symbol up at
    // (sum of the block's runs so far, mod 64) + (the symbol's ordinal in the block + 1): the ordinal is a constant of the
    // unrolled slot plus a per-group base, the sum wraps by itself in the top six bits of a register.  In a sound stream that
    // index is the position (<= 63); a block has at most 144 symbols (the producer ends it once the position passes 63 / 127),
    // so 256 entries -- 64..255 repeat 0..63 -- keep every look-up of a flagged stream inside the table as well.
    // (8 words in front of each table: the slots of a block's first group that lie in front of the block look up ordinals -7..0)
    __shared__ uint32_t tzs[2][8 + 256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < 128) {
        const uint32_t z = tabs->tz[threadIdx.x >> 6][threadIdx.x & 63];
        uint32_t *t = &tzs[threadIdx.x >> 6][8 + (threadIdx.x & 63)];
        t[0] = z; t[64] = z; t[128] = z; t[192] = z;
    }
    __syncthreads();
    // block-stride loop (the launcher covers every block with one pass; see launch_idct16): a thread only ever touches its own
    // column of the tile, so iterations would need no synchronisation
    const int64_t total = (int64_t)n * g.nblk;
    for (int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gt < total; gt += (int64_t)gridDim.x * blockDim.x) {
    // the block index fits 32 bits for every batch the workspace can hold at once (warp-uniform choice of the divide)
    // (the three divisions of this prologue -- frame, block row in the plane -- cost a tenth of the kernel's instructions as
    // emulated divides: they are one high multiply and a correction each, div_by_magic)
    uint32_t iu;
    const int f = (gt >> 32) == 0 ? (int)div_by_magic((uint32_t)gt, (uint32_t)g.nblk, mg_nblk, iu) : (int)(gt / g.nblk);
    if (scan_len[f] == 0) continue;
    const int i = (gt >> 32) == 0 ? (int)iu : (int)(gt - (int64_t)f * g.nblk);
    // blocks are enumerated in plane raster order (all of Y, then Cb, then Cr); blk = the block's index in
    // bitstream order: MCU by MCU, inside an MCU component by component, v x h blocks in raster order
    const int nluma = g.nl * g.mbw * g.mbh, nchroma = g.nc * g.mbw * g.mbh;
    int comp, bx, by, blk;
    if (i < nluma) {
        comp = 0;
        const int rowb = g.mbw << g.llh;
        uint32_t rem;
        by = (int)div_by_magic((uint32_t)i, (uint32_t)rowb, mg_rowl, rem); bx = (int)rem;
        blk = ((by >> g.llv) * g.mbw + (bx >> g.llh)) * g.nb + ((by & ((1 << g.llv) - 1)) << g.llh) + (bx & ((1 << g.llh) - 1));
    } else {
        const int j = i - nluma;
        comp = j < nchroma ? 1 : 2;
        const int jj = comp == 1 ? j : j - nchroma;
        const int rowb = g.mbw << g.lch;
        uint32_t rem;
        by = (int)div_by_magic((uint32_t)jj, (uint32_t)rowb, mg_rowc, rem); bx = (int)rem;
        blk = ((by >> g.lcv) * g.mbw + (bx >> g.lch)) * g.nb + g.nl + (comp - 1) * g.nc +
              ((by & ((1 << g.lcv) - 1)) << g.lch) + (bx & ((1 << g.lch) - 1));
    }
    const uint32_t bo = blk_off[(uint64_t)f * g.nblk + blk];
    const uint16_t *tf = tokens + slot_off[f] * 4;                         // the frame's token region (16-byte aligned)
    const uint32_t first = bo & ((1u << kTokCountShift) - 1u), last = first + (bo >> kTokCountShift);   // DC token .. last symbol token
    uint32_t *slot = &tile[wid][lane];
    const uint32_t slot_s = smem_addr(slot);
    const uint32_t tz_s = smem_addr(&tzs[comp ? 1 : 0][8]);
#pragma unroll
    for (int k = 0; k < 32; k++) slot[k * 32] = 0;
    // The DC token is absolute and already dequantised.  The symbol tokens behind it are fetched as aligned groups of eight
    // (one 128-bit load), one group ahead in flight; the tokens of a group that belong to the neighbouring blocks are
    // skipped by predicate (no branches: the slots of a group are straight-line code).  The look-ahead stays inside the
    // region's slack.
    sts16(slot_s, (uint32_t)__ldg(tf + first));
    const uint32_t cnt = bo >> kTokCountShift;                             // symbol tokens: indices first + 1 .. first + cnt
    uint32_t gi = (first + 1u) & ~7u;
    uint4 q = __ldg(reinterpret_cast<const uint4 *>(tf + gi));
    {   // the tokens of the first group that lie in front of the block are cleared (run 0): the sum below takes every token
        const int lead = 16 * (int)(first + 1u - gi);                      // bits in front of the block's first symbol token
        q.x &= __funnelshift_lc(0u, ~0u, max(lead, 0));
        q.y &= __funnelshift_lc(0u, ~0u, max(lead - 32, 0));
        q.z &= __funnelshift_lc(0u, ~0u, max(lead - 64, 0));
        q.w &= __funnelshift_lc(0u, ~0u, max(lead - 96, 0));
    }
    uint32_t k26 = 0;                                                      // (sum of the runs of the symbols so far) << 26
    uint32_t rel = gi - first - 1u;                                        // index of the group's first token among the block's symbols (may be "negative")
    uint32_t tzg = tz_s + 4u * rel;                                        // table address of position = ordinal of the group's first token
    for (; gi <= last; gi += 8, rel += 8, tzg += 32u) {
        const uint4 nq = __ldg(reinterpret_cast<const uint4 *>(tf + gi + 8));
        const uint32_t tw[4] = { q.x, q.y, q.z, q.w };
        // which of the group's eight tokens are the block's: 0 <= rel + j < cnt, as one bit mask per group (a bit test per
        // token instead of an add and a compare)
        const int srel = (int)rel;
        const uint32_t vm = (0xffu >> (8 - min((int)cnt - srel, 8))) & (0xffu << max(-srel, 0));
        // one slot of the group; its index is a compile-time constant (it rides in the table load's immediate field)
        auto slot = [&](auto J) {
            constexpr int j = decltype(J)::value;
            const uint32_t w = tw[j >> 1];
            const bool valid = (vm >> j) & 1u;
            const int lvl = (j & 1) ? (int)(w << 4) >> 20 : sext12(w);
            // the run nibble, moved to bit 26, joins the sum: one mask, one multiply-add that shifts and adds (tokens behind
            // the block's last only move a sum nobody reads any more)
            k26 = (j & 1) ? mad_hi_u32(w & 0xf0000000u, 1u << 30, k26) : (uint32_t)mad_lo((int)(w & 0xf000u), 1 << 14, (int)k26);
            const uint32_t z = lds32_at<4 * (j + 1)>(tzg + (k26 >> 24));
            // z = column offset << 16 | quantiser: the halfword stored is the low half of level x z, which is level x quantiser
            // truncated to int16 whatever the offset adds above bit 15 -- no mask needed
            if (valid) sts16(slot_s + (z >> 16), (uint32_t)lvl * z);
        };
        slot(std::integral_constant<int, 0>{}); slot(std::integral_constant<int, 1>{}); slot(std::integral_constant<int, 2>{});
        slot(std::integral_constant<int, 3>{}); slot(std::integral_constant<int, 4>{}); slot(std::integral_constant<int, 5>{});
        slot(std::integral_constant<int, 6>{}); slot(std::integral_constant<int, 7>{});
        q = nq;
    }

    uint32_t c[32], o[16];
#pragma unroll
    for (int kk = 0; kk < 32; kk++) c[kk] = slot[kk * 32];
    // Smooth content (chroma planes almost always) has nothing below the second coefficient row: when that holds
    // for the whole warp, take the transform specialised for it (same results, a quarter of the arithmetic).
    uint32_t lower = 0;
#pragma unroll
    for (int kk = 8; kk < 32; kk++) lower |= c[kk];
    if (__all_sync(__activemask(), lower == 0)) idct_put_block<2>(c, o);
    else idct_put_block<8>(c, o);

    uint8_t *pl = comp == 0 ? py + (uint64_t)f * fs_y : (comp == 1 ? pu : pv) + (uint64_t)f * fs_c;
    const int ls = comp ? ls_c : ls_y;
    const int vw = comp ? g.cw : g.w, vh = comp ? g.ch : g.h, r0 = comp ? g.c0 : g.y0;
    const int x0 = bx * 8, y0 = by * 8;
    // AMV pictures are stored bottom-up (mjpegdec.c:672-677), SP5X is not: the block's first row and the step to the next
    const int row0 = g.flip ? r0 - y0 : y0, rstep = g.flip ? -1 : 1;
    const int64_t dstep = g.flip ? -(int64_t)ls : (int64_t)ls;
    uint8_t *d = pl + (int64_t)row0 * ls + x0;
    const int row7 = row0 + 7 * rstep;
    if (FAST && row0 >= 0 && row0 < vh && row7 >= 0 && row7 < vh) {
        // the whole block lies inside the picture (every block of a picture whose height is a multiple of 16): eight stores,
        // no per-row tests
#pragma unroll
        for (int yy = 0; yy < 8; yy++, d += dstep) *reinterpret_cast<uint2 *>(d) = make_uint2(o[2 * yy], o[2 * yy + 1]);
        continue;
    }
#pragma unroll
    for (int yy = 0; yy < 8; yy++, d += dstep) {
        const int row = row0 + rstep * yy;
        if (row < 0 || row >= vh) continue;
        if (FAST) {
            *reinterpret_cast<uint2 *>(d) = make_uint2(o[2 * yy], o[2 * yy + 1]);
        } else {
#pragma unroll
            for (int xx = 0; xx < 8; xx++)
                if (x0 + xx < vw) d[xx] = (uint8_t)(o[2 * yy + (xx >> 2)] >> (8 * (xx & 3)));
        }
    }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static bool fill_table_set(DecTableSet &T, const HuffSpec &H, const uint8_t qzz[2][64], bool *sync_ok) {
    if (!build_flat_vlc_tables_from(T.flat, H)) return false;
    if (sync_ok) *sync_ok = true;          // the synchronisation pass uses the same table
    DequantTables dq;
    build_dequant_tables_from(dq, qzz);
    memcpy(T.tz, dq.tz, sizeof(T.tz));
    T.q0[0] = qzz[0][0]; T.q0[1] = qzz[1][0];
    return true;
}

cudaError_t upload_dec_tables(cudaStream_t s) {
    static DecTableSet h[2];      // built once; identical for every context
    static bool built = false;
    if (!built) {
        HuffSpec H;
        fixed_huff_spec(H);
        if (!fill_table_set(h[0], H, kDecQuant, nullptr)) return cudaErrorInvalidValue;
        h[1] = h[0];
        AmvlibDequantTables adq;
        build_amvlib_dequant_tables(adq);
        memcpy(h[1].tz, adq.tz, sizeof(h[1].tz));
        h[1].q0[0] = kAmvlibQuant[0][0]; h[1].q0[1] = kAmvlibQuant[1][0];
        built = true;
    }
    return cudaMemcpyToSymbolAsync(g_dec_sets, h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
}

const DecTableSet *fixed_dec_tables(bool amvlib) {
    void *p = nullptr;
    if (cudaGetSymbolAddress(&p, g_dec_sets) != cudaSuccess) return nullptr;
    return static_cast<const DecTableSet *>(p) + (amvlib ? 1 : 0);
}

size_t dec_table_set_bytes() { return sizeof(DecTableSet); }

bool build_dec_table_set(void *host_buf, const uint8_t counts[4][16], const uint8_t syms[4][256], const uint8_t qzz[2][64],
                         bool *sync_ok) {
    HuffSpec H;
    for (int t = 0; t < 4; t++) {
        if (!huff_counts_valid(counts[t])) return false;
        memcpy(H.counts[t], counts[t], 16);
        memcpy(H.syms[t], syms[t], 256);
    }
    for (int t = 0; t < 2; t++)                        // DC categories above 16 bits cannot be coefficients of this path
        for (int k = 0; k < 256; k++) if (H.syms[t][k] > 16) return false;
    return fill_table_set(*static_cast<DecTableSet *>(host_buf), H, qzz, sync_ok);
}

// opt-in to more than 48 KB of dynamic shared memory: per device, called from amv_create
cudaError_t decode_setup_device() {
    cudaError_t e = cudaFuncSetAttribute(k_vlc_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyncSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_sync_lean<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sync_lean_smem_bytes<8>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_sync_lean<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sync_lean_smem_bytes<7>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens<kFlavorJpeg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTokSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens<kFlavorJpegDri>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTokSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens<kFlavorAmvlib>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTokSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens<kFlavorFfmpeg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTokSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens_lean<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok16_smem_bytes<7>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens_lean<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok16_smem_bytes<7>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens_lean<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok16_smem_bytes<8>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens_lean<true, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok16_smem_bytes<11>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens_lean<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok16_smem_bytes<8>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vlc_tokens_lean<false, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tok16_smem_bytes<11>());
    return e;
}

void launch_scan_sizes(const uint32_t *size, int n, uint32_t align_mask, uint32_t pad, uint64_t *off,
                       uint64_t *carry_io, cudaStream_t s) {
    AMV_LAUNCH(k_scan_sizes, 1, 1024, 0, s, size, n, align_mask, pad, off, carry_io);
}

void launch_unstuff(const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                    uint8_t *scratch, const uint64_t *slot_off, uint64_t scratch_bytes, uint32_t *scan_len,
                    int32_t *status, uint32_t head, bool literal, cudaStream_t s) {
    // One-warp CTAs (512-byte tiles): the tile loop is a chain of synchronisations, small CTAs keep more independent
    // chains per SM, and one warp needs no CTA barrier at all (measured per 100k frames: 2.6 / 2.1 / 2.06 ms at 256 /
    // 128 / 64 threads, 1.94 ms at 32 threads with barriers, see the kernel for the warp-synchronous form)
    constexpr int kThreads = 32, kPerSM = 32;
    const int grid = n < kNumSMs * kPerSM ? n : kNumSMs * kPerSM;
    AMV_LAUNCH(k_unstuff<kThreads>, grid, kThreads, 0, s, pkts, pkts_bytes, pkt_off, pkt_size, n, scratch, slot_off, scratch_bytes,
                                                  scan_len, status, head, literal ? 1 : 0);
}

// Plain JPEG frames are decoded with the Huffman tables and frame geometry of ONE header (amv_mjpeg_configure),
// but each with the quantisers of its own DQT segment (an encoder under rate control rewrites them per frame):
// the frame's marker segments must equal the sample's byte for byte outside the two 64-byte quantiser fields,
// which are copied to qtab for the token kernel.  A frame that differs is not decoded and says so.
__global__ void k_mjpeg_check(const uint8_t *__restrict__ pkts, uint64_t pkts_bytes, const uint64_t *__restrict__ pkt_off,
                              const uint32_t *__restrict__ pkt_size, int n, const uint8_t *__restrict__ hdr, uint32_t hdr_len,
                              uint32_t qpos0, uint32_t qpos1, uint8_t *__restrict__ qtab, uint32_t *__restrict__ scan_len,
                              int32_t *__restrict__ status) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint64_t off = pkt_off[f];
    const uint32_t size = pkt_size[f];
    uint8_t *q = qtab + (size_t)f * 128;
    const bool inside = range_ok(off, size, pkts_bytes);
    if (!inside || size < hdr_len + 2u) {      // out of range: k_unstuff reported it; too short: no scan
        for (int i = lane; i < 128; i += 32) q[i] = 1;
        if (inside && lane == 0) { scan_len[f] = 0; atomicOr(&status[f], AMV_ST_HEADER); }
        return;
    }
    bool diff = false;
    for (uint32_t i = lane; i < hdr_len; i += 32) {
        const bool quant = (i - qpos0 < 64u) || (i - qpos1 < 64u);
        diff |= !quant && pkts[off + i] != hdr[i];
    }
    for (int i = lane; i < 64; i += 32) { q[i] = pkts[off + qpos0 + i]; q[64 + i] = pkts[off + qpos1 + i]; }
    if (__any_sync(0xffffffffu, diff) && lane == 0) { scan_len[f] = 0; atomicOr(&status[f], AMV_ST_HEADER); }
}

void launch_mjpeg_check(const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                        const uint8_t *hdr, uint32_t hdr_len, uint32_t qpos0, uint32_t qpos1, uint8_t *qtab,
                        uint32_t *scan_len, int32_t *status, cudaStream_t s) {
    AMV_LAUNCH(k_mjpeg_check, (n + 7) / 8, 256, 0, s, pkts, pkts_bytes, pkt_off, pkt_size, n, hdr, hdr_len, qpos0, qpos1, qtab, scan_len,
                                              status);
}

// Warps per CTA of the lean VLC kernels for a launch of `lanes` lanes.  Their shared memory (45 KB of tables + 2..3 KB per
// warp) lets three CTAs of 7 or 8 warps or two of 11 warps live on an SM; the CTA scheduler deals a one-wave grid evenly, so the
// busiest SM carries ceil(CTAs / 148) CTAs.  Among the sizes that fit the launch into one wave the one with the fewest warps on
// the busiest SM wins (3 125 warps: 11 -> 22 warps instead of 24 with 8; 2 048 warps: 7 -> 14 instead of 16); launches of
// several waves keep 8.
static int pick_vlc_warps(int64_t lanes, bool allow11) {
    const int64_t warps = (lanes + 31) / 32;
    const int cand[3] = { 8, 7, 11 }, resident[3] = { 3, 3, 2 };
    int best = 8;
    int64_t best_max = INT64_MAX;
    for (int i = 0; i < (allow11 ? 3 : 2); i++) {
        const int64_t ctas = (warps + cand[i] - 1) / cand[i];
        if (ctas > (int64_t)kNumSMs * resident[i]) continue;           // more than one wave
        const int64_t busiest = ((ctas + kNumSMs - 1) / kNumSMs) * cand[i];
        if (busiest < best_max) { best_max = busiest; best = cand[i]; }
    }
    return best;
}

void launch_vlc_sync(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, int n, int log2p,
                     LaneStart *starts, uint32_t *rounds_out, bool amvlib, const DecTableSet *tabs, const uint8_t *qtab,
                     int nl, int nc, bool lean, cudaStream_t s) {
    const int64_t lanes = (int64_t)n << log2p;
    const int grid = (int)((lanes + kTokThreads - 1) / kTokThreads);
    if (lean && !amvlib && !qtab) {         // fixed AMV / SP5X tables
        if (pick_vlc_warps(lanes, false) == 7)
            AMV_LAUNCH(k_vlc_sync_lean<7>, (int)((lanes + 223) / 224), 224, sync_lean_smem_bytes<7>(), s, scratch, slot_off, scan_len, n, log2p, starts, rounds_out, tabs, nl, nc);
        else
            AMV_LAUNCH(k_vlc_sync_lean<8>, grid, kTokThreads, sync_lean_smem_bytes<8>(), s, scratch, slot_off, scan_len, n, log2p, starts, rounds_out, tabs, nl, nc);
        return;
    }
    AMV_LAUNCH(k_vlc_sync, grid, kTokThreads, kSyncSmemBytes, s, scratch, slot_off, scan_len, n, log2p, starts, rounds_out,
                                            amvlib ? kFlavorAmvlib : (qtab ? kFlavorJpeg : kFlavorFfmpeg), tabs, qtab, nl, nc);
}

void launch_vlc_tokens(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, const uint32_t *pkt_size,
                       int n, int log2p, const LaneStart *starts, int nblk, uint32_t *tokens, uint32_t *blk_off,
                       int32_t *status, bool amvlib, const DecTableSet *tabs, const uint8_t *qtab, int nl, int nc, int restart,
                       cudaStream_t s) {
    const int64_t lanes = (int64_t)n << log2p;
    const int grid = (int)((lanes + kTokThreads - 1) / kTokThreads);
    if (amvlib)
        AMV_LAUNCH(k_vlc_tokens<kFlavorAmvlib>, grid, kTokThreads, kTokSmemBytes, s, scratch, slot_off, scan_len, pkt_size, n, log2p, starts, nblk,
                                                                 tokens, blk_off, status, tabs, nullptr, nl, nc, 0);
    else if (qtab && restart)
        AMV_LAUNCH(k_vlc_tokens<kFlavorJpegDri>, grid, kTokThreads, kTokSmemBytes, s, scratch, slot_off, scan_len, pkt_size, n, log2p, starts,
                                                                  nblk, tokens, blk_off, status, tabs, qtab, nl, nc, restart);
    else if (qtab)
        AMV_LAUNCH(k_vlc_tokens<kFlavorJpeg>, grid, kTokThreads, kTokSmemBytes, s, scratch, slot_off, scan_len, pkt_size, n, log2p, starts, nblk,
                                                               tokens, blk_off, status, tabs, qtab, nl, nc, 0);
    else
        AMV_LAUNCH(k_vlc_tokens<kFlavorFfmpeg>, grid, kTokThreads, kTokSmemBytes, s, scratch, slot_off, scan_len, pkt_size, n, log2p, starts, nblk,
                                                                 tokens, blk_off, status, tabs, nullptr, nl, nc, 0);
}

void launch_vlc_tokens16(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, const uint32_t *pkt_size,
                         int n, int log2p, const LaneStart *starts, int nblk, uint16_t *tokens, uint32_t *blk_off,
                         int32_t *status, const DecTableSet *tabs, int nl, int nc, cudaStream_t s) {
    const int64_t lanes = (int64_t)n << log2p;
    const int nw = pick_vlc_warps(lanes, true);
    if (nw == 11)
        AMV_LAUNCH((k_vlc_tokens_lean<true, 11>), (int)((lanes + 351) / 352), 352, tok16_smem_bytes<11>(), s, scratch, slot_off, scan_len, pkt_size, n,
                   log2p, starts, nblk, tokens, blk_off, status, tabs, nl, nc);
    else if (nw == 7)
        AMV_LAUNCH((k_vlc_tokens_lean<true, 7>), (int)((lanes + 223) / 224), 224, tok16_smem_bytes<7>(), s, scratch, slot_off, scan_len, pkt_size, n,
                   log2p, starts, nblk, tokens, blk_off, status, tabs, nl, nc);
    else
        AMV_LAUNCH((k_vlc_tokens_lean<true, 8>), (int)((lanes + 255) / 256), 256, tok16_smem_bytes<8>(), s, scratch, slot_off, scan_len, pkt_size, n,
                   log2p, starts, nblk, tokens, blk_off, status, tabs, nl, nc);
}

void launch_vlc_tokens_lean(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, const uint32_t *pkt_size,
                            int n, int log2p, const LaneStart *starts, int nblk, uint32_t *tokens, uint32_t *blk_off,
                            int32_t *status, const DecTableSet *tabs, int nl, int nc, cudaStream_t s) {
    const int64_t lanes = (int64_t)n << log2p;
    const int nw = pick_vlc_warps(lanes, true);
    if (nw == 11)
        AMV_LAUNCH((k_vlc_tokens_lean<false, 11>), (int)((lanes + 351) / 352), 352, tok16_smem_bytes<11>(), s, scratch, slot_off, scan_len, pkt_size, n,
                   log2p, starts, nblk, tokens, blk_off, status, tabs, nl, nc);
    else if (nw == 7)
        AMV_LAUNCH((k_vlc_tokens_lean<false, 7>), (int)((lanes + 223) / 224), 224, tok16_smem_bytes<7>(), s, scratch, slot_off, scan_len, pkt_size, n,
                   log2p, starts, nblk, tokens, blk_off, status, tabs, nl, nc);
    else
        AMV_LAUNCH((k_vlc_tokens_lean<false, 8>), (int)((lanes + 255) / 256), 256, tok16_smem_bytes<8>(), s, scratch, slot_off, scan_len, pkt_size, n,
                   log2p, starts, nblk, tokens, blk_off, status, tabs, nl, nc);
}

void launch_idct16(const uint16_t *tokens, const uint32_t *blk_off, const uint64_t *slot_off, const uint32_t *scan_len, int n,
                   const Geom &g, const DecTableSet *tabs, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c, uint64_t fs_y,
                   uint64_t fs_c, cudaStream_t s) {
    const int64_t threads = (int64_t)n * g.nblk;
    // one CTA per 128 blocks.  The kernel is written as a block-stride loop, but a resident grid measured slower (6 CTAs per
    // SM 9.53 ms, 12 CTAs 8.77 ms, one CTA per 128 blocks 8.49 ms per 100 000 frames, profiles/r6l_*): the hardware's CTA
    // scheduler balances the SMs better than a static stride does
    const int64_t grid = (threads + kIdctThreads - 1) / kIdctThreads;
    const bool fast = (g.w % 16 == 0) &&
                      ((((uintptr_t)y | (uintptr_t)u | (uintptr_t)v | (uintptr_t)ls_y | (uintptr_t)ls_c | fs_y | fs_c) & 7) == 0);
    const uint32_t mg_nblk = div_magic((uint32_t)g.nblk), mg_rowl = div_magic((uint32_t)(g.mbw << g.llh)), mg_rowc = div_magic((uint32_t)(g.mbw << g.lch));
    if (fast) AMV_LAUNCH(k_idct16<true>, (unsigned)grid, kIdctThreads, 0, s, tokens, blk_off, slot_off, scan_len, n, g, tabs, y, u, v, ls_y, ls_c, fs_y, fs_c,
                         mg_nblk, mg_rowl, mg_rowc);
    else      AMV_LAUNCH(k_idct16<false>, (unsigned)grid, kIdctThreads, 0, s, tokens, blk_off, slot_off, scan_len, n, g, tabs, y, u, v, ls_y, ls_c, fs_y, fs_c,
                         mg_nblk, mg_rowl, mg_rowc);
}

void launch_idct(const uint32_t *tokens, const uint32_t *blk_off, const uint64_t *slot_off, const uint32_t *scan_len, int n,
                 const Geom &g, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                 cudaStream_t s) {
    const int64_t threads = (int64_t)n * g.nblk;
    const int64_t grid = (threads + kIdctThreads - 1) / kIdctThreads;
    const bool fast = (g.w % 16 == 0) &&
                      ((((uintptr_t)y | (uintptr_t)u | (uintptr_t)v | (uintptr_t)ls_y | (uintptr_t)ls_c | fs_y | fs_c) & 7) == 0);
    const uint32_t mg_nblk = div_magic((uint32_t)g.nblk), mg_rowl = div_magic((uint32_t)(g.mbw << g.llh)), mg_rowc = div_magic((uint32_t)(g.mbw << g.lch));
    if (fast)
        AMV_LAUNCH(k_idct<true>, (unsigned)grid, kIdctThreads, 0, s, tokens, blk_off, slot_off, scan_len, n, g, y, u, v, ls_y, ls_c,
                                                             fs_y, fs_c, mg_nblk, mg_rowl, mg_rowc);
    else
        AMV_LAUNCH(k_idct<false>, (unsigned)grid, kIdctThreads, 0, s, tokens, blk_off, slot_off, scan_len, n, g, y, u, v, ls_y, ls_c,
                                                              fs_y, fs_c, mg_nblk, mg_rowl, mg_rowc);
}

}  // namespace amv

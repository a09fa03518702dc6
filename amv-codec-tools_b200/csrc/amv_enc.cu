// amv_enc.cu -- AMV video encode kernels (sm_100a).
//
// One WARP encodes one frame at a time (persistent, warp-stride over the batch) so that the two
// serial quantities of a frame -- the running bit position of the entropy-coded segment and the
// count of stuffed FF bytes -- never leave the warp: no look-back, no scratch traffic, no CTA
// barriers, the packet is written once.  Three kernels share that plan:
//
//   k_encode16v2  the one that runs (option encode_rounds = 4; 2 = the same with the factorised transform).  A
//                 frame is cut into segments of 16 macroblocks coded in three homogeneous rounds (32 chroma
//                 blocks, then twice 32 luma blocks), every lane one block per round:
//     A  coalesced 64-bit row loads of the bottom-up picture (amv_encode_picture mjpegenc.c:454-472
//        + edge replication mpegvideo.c:1416-1470), fdct_islow in registers (jfdctint.c:261) regrouped for
//        the two integer pipes -- row pass as IDP.2A / IDP.4A dot products on the packed pixel bytes, column
//        pass with its odd half written out (amv_dct.cuh) --, the bit-reversed zigzag mask of the
//        coefficients that survive the quantiser (mpegvideo_enc.c:3647)
//     B  ONE Huffman pass per block into a lane-private bit string (encode_block mjpegenc.c:379-435),
//        quantising only the survivors, 37 instructions per coded coefficient, branch-free bit writer
//     C  prefix scan of the string lengths in bitstream order (warp shuffles)
//     D  warp-cooperative bit packer: every lane shifts its string to its scanned bit offset and ORs
//        it into the half segment's shared-memory bit buffer
//     E  FF00 stuffing (escape_FF mjpegenc.c:282-336) as a second scan over FF counts, bytes staged
//        in shared memory and stored to the packet slot as aligned 128-bit units; SOI/EOI framing and
//        1-bit padding (ff_mjpeg_encode_stuffing :338-343, trailer :345-355) at frame start / end
//   k_encode16    round 1's version of the same rounds (encode_rounds = 1; quantises every coefficient,
//                 stores packet bytes one per lane)
//   k_encode      the first encoder: segments of 5 macroblocks = 30 blocks on 30 lanes (20 luma, 5 Cb,
//                 5 Cr), a code-LENGTH pass before the scan and a second Huffman pass that packs.  It
//                 has no limit on a block's code length, so it takes the frames the rounds kernels
//                 hand back (a string that outgrows its staging column) and runs alone at
//                 encode_rounds = 0.
#include "amv_common.cuh"
#include "amv_tables.cuh"
#include "amv_dct.cuh"
#include "amv_kernels.h"

namespace amv {

constexpr int kEncWarps = 4;                    // warps per CTA, each encoding its own frames
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kSegMB = 5;                       // macroblocks per segment: 30 blocks on 30 lanes
constexpr int kSegBlocks = kSegMB * 6;
// worst case per block: 20-bit DC + 63 x 26-bit AC = 1658 bits = 52 words
constexpr int kStageWords = (20 + 63 * 26 + 31) / 32;
constexpr int kSegWords = kSegBlocks * kStageWords + 8;

struct EncTablesDev {
    EncHuffTables huff;
    uint8_t zigzag[64];
    uint8_t intra_base[64];
};
__device__ EncTablesDev g_enc_tables;

// per-warp working set (dynamic shared memory, one instance per warp)
struct EncWarpSmem {
    union {                                 // never live at the same time (A,B use coef; D,E use seg)
        uint16_t coef[64 * 32];             // halfword k*32 + lane : zigzag coefficient k of the lane's block
        uint32_t seg[kSegWords];            // the segment's contiguous bit string
    } u;
    uint32_t stage[kStageWords * 32];       // word w*32 + lane : the lane's private bit string
    uint32_t qm10[64];
    uint32_t lens[32];                      // bit length per block in bitstream order -> exclusive offsets
    int      dcq[32];                       // quantised DC per block in bitstream order
    int      carry_dc[4];                   // last DC of each component from the previous segment
};
struct EncSmem {
    uint32_t huff[kEncHuffEntries];
    EncWarpSmem w[kEncWarps];
};

__device__ __forceinline__ int bit_width(uint32_t v) { return 32 - __clz(v); }

// exact per-byte "== 0xFF" detector: 0x80 in every byte of v that is FF
__device__ __forceinline__ uint32_t ff_bytes(uint32_t v) {
    const uint32_t t = ~v;                                       // FF bytes become 00
    return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

// One WARP encodes one frame at a time (persistent, warp-stride over the batch), so the frame's two
// serial quantities -- running bit position and count of stuffed FF bytes -- never leave the warp
// and every synchronisation is a __syncwarp: warps of an SM run fully independently of each other.
template <bool FAST>
__global__ void __launch_bounds__(kEncThreads)
k_encode(const uint8_t *__restrict__ py, const uint8_t *__restrict__ pu, const uint8_t *__restrict__ pv,
         int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int n, Geom g, const int32_t *__restrict__ qscale,
         uint8_t *__restrict__ slots, uint64_t slot_stride, uint32_t pkt_cap, uint32_t *__restrict__ out_size,
         int32_t *__restrict__ status, const int32_t *__restrict__ only /* optional: encode frame f only if only[f] != 0 */) {
    AMV_EXTERN_SHARED(uint8_t, smem_raw, 16);
    EncSmem &S = *reinterpret_cast<EncSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kEncHuffEntries; i += kEncThreads) S.huff[i] = g_enc_tables.huff.e[i];
    __syncthreads();
    EncWarpSmem &W = S.w[wid];

    // block role of this lane inside a segment: lanes 0..19 luma (4 per MB), 20..24 Cb, 25..29 Cr
    int mi, b;
    if (lane < 20) { mi = lane >> 2; b = lane & 3; }
    else if (lane < 25) { mi = lane - 20; b = 4; }
    else { mi = lane - 25; b = 5; }
    const bool has_role = lane < kSegBlocks;
    const int comp = b < 4 ? 0 : b - 3;
    const int sigma = has_role ? mi * 6 + b : 31;        // position in bitstream order (0..29)
    const int total_mb = g.mbw * g.mbh;
    // valid source extent: the reference copies w x h luma and (w>>1) x (h>>1) chroma (mpegvideo_enc.c:866-867)
    const int vw = comp ? (g.w >> 1) : g.w, vh = comp ? (g.h >> 1) : g.h, r0 = comp ? g.c0 : g.y0;
    const int ls = comp ? ls_c : ls_y;
    // explicit shared addresses for the hot loops
    const uint32_t huff_dc_s = smem_addr(&S.huff[comp ? kEncDcChroma : kEncDcLuma]);
    const uint32_t huff_ac_s = smem_addr(&S.huff[comp ? kEncAcChroma : kEncAcLuma]);
    const uint32_t coef_s = smem_addr(&W.u.coef[lane]);          // coefficient k: + k*64
    const uint32_t stage_s = smem_addr(&W.stage[lane]);          // word w: + w*128
    const uint32_t seg_s = smem_addr(&W.u.seg[0]);
    const int gw = blockIdx.x * kEncWarps + wid, nw_total = gridDim.x * kEncWarps;

    for (int f = gw; f < n; f += nw_total) {
        if (only && !only[f]) continue;
        const int qs = qscale ? qscale[f] : 2;
        __syncwarp();
        for (int t = lane; t < 64; t += 32) {
            // intra_matrix / q_intra_matrix for this frame (mpegvideo_enc.c:2866-2877, ff_convert_matrix :69-91)
            int m = 8;
            if (t) m = min(max(((int)g_enc_tables.intra_base[t] * qs) >> 3, 1), 255);
            W.qm10[t] = ((1u << 22) / (uint32_t)(8 * m)) << 10;
        }
        if (lane < 3) W.carry_dc[lane] = 128;            // last_dc init (mpegvideo_enc.c:2033-2036)
        uint32_t overflow = (qs < 2 || qs > 31) ? AMV_ST_RANGE : 0;      // qscale domain: SURVEY 9.13
        uint32_t carry_word = 0;                        // partial word carried into the next segment
        const uint8_t *pl = (comp == 0 ? py + (uint64_t)f * fs_y : (comp == 1 ? pu : pv) + (uint64_t)f * fs_c);
        uint8_t *pkt = slots + (uint64_t)f * slot_stride;
        uint32_t G = 2;                                 // bytes written so far (SOI)
        uint32_t r = 0;                                 // carried bits (sit in carry_word, MSB side)
        if (lane == 0 && pkt_cap >= 2) { pkt[0] = 0xff; pkt[1] = 0xd8; }
        __syncwarp();

        for (int m0 = 0; m0 < total_mb; m0 += kSegMB) {
            const int nmb = min(kSegMB, total_mb - m0);
            const bool active = has_role && mi < nmb;
            const bool last_seg = m0 + kSegMB >= total_mb;
            uint32_t mask_lo = 0, mask_hi = 0;          // non-zero AC positions (zigzag) of this block
            int dc = 0;

            // ---------------- A: load, FDCT, quantise
            if (active) {
                const int mb = m0 + mi;
                const int mx = mb % g.mbw, my = mb / g.mbw;
                const int bx = comp ? mx * 8 : mx * 16 + (b & 1) * 8;
                const int by = comp ? my * 8 : my * 16 + (b >> 1) * 8;
                int v[64];
#pragma unroll
                for (int yy = 0; yy < 8; yy++) {
                    const int Y = min(by + yy, vh - 1);                 // bottom edge replication
                    const uint8_t *row = pl + (int64_t)(r0 - Y) * ls;
                    if (FAST) {
                        const uint2 q = *reinterpret_cast<const uint2 *>(row + bx);
#pragma unroll
                        for (int xx = 0; xx < 4; xx++) {
                            v[yy * 8 + xx]     = (q.x >> (8 * xx)) & 0xff;
                            v[yy * 8 + 4 + xx] = (q.y >> (8 * xx)) & 0xff;
                        }
                    } else {
#pragma unroll
                        for (int xx = 0; xx < 8; xx++) v[yy * 8 + xx] = row[min(bx + xx, vw - 1)];   // right edge replication
                    }
                }
                fdct_block(v);
                dc = quant_dc(v[0]);
                // raster order so the multiplier loads vectorise; mask bit = zigzag position
#pragma unroll
                for (int j = 1; j < 64; j++) {
                    const int q = quant_ac(v[j], W.qm10[j]);
                    v[j] = q;
                    const int k = zigzag_inv_at(j);
                    if (q) { if (k < 32) mask_lo |= 1u << k; else mask_hi |= 1u << (k - 32); }
                }
#pragma unroll
                for (int k = 1; k < 64; k++) W.u.coef[k * 32 + lane] = (uint16_t)v[zigzag_at(k)];
                W.dcq[sigma] = dc;
            }
            __syncwarp();

            // ---------------- B: Huffman-code the block into the lane's private bit string
            uint32_t len = 0;
            if (active) {
                int pred;
                if (comp == 0) pred = (b > 0) ? W.dcq[sigma - 1] : (mi > 0 ? W.dcq[sigma - 3] : W.carry_dc[0]);
                else           pred = mi > 0 ? W.dcq[sigma - 6] : W.carry_dc[comp];
                uint32_t acc = 0;                   // MSB-first accumulator: the top `fill` bits are valid
                uint32_t fill = 0;
                uint32_t wp = stage_s;              // next private word
                auto put = [&](uint32_t code, uint32_t nbits) {  // 1 <= nbits <= 27
                    const uint32_t t = code << (32 - nbits);     // left-aligned
                    acc |= t >> fill;
                    const uint32_t nf = fill + nbits;
                    if (nf >= 32) { sts32(wp, acc); wp += 128; acc = t << (32 - fill); fill = nf - 32; }   // fill >= 5 here
                    else fill = nf;
                };
                {   // DC (ff_mjpeg_encode_dc, mjpegenc.c:357-377)
                    const int diff = dc - pred;
                    const int nb = bit_width((uint32_t)(diff < 0 ? -diff : diff));
                    const uint32_t e = lds32(huff_dc_s + nb * 4);
                    const uint32_t mant = (uint32_t)(diff + (diff >> 31)) & ((1u << nb) - 1u);
                    put(((e >> 5) << nb) | mant, (e & 31) + (uint32_t)nb);
                }
                const uint32_t ezrl = lds32(huff_ac_s + 0xf0 * 4), eeob = lds32(huff_ac_s);
                int prevk = 0;
                auto ac_run = [&](uint32_t m, int base) {        // encode_block's AC loop (mjpegenc.c:403-430)
                    while (m) {
                        const int k = base + __ffs((int)m) - 1;
                        m &= m - 1;
                        int run = k - prevk - 1;
                        prevk = k;
                        const int cv = lds_s16(coef_s + (uint32_t)k * 64);
                        const int cb = bit_width((uint32_t)(cv < 0 ? -cv : cv));
                        for (; run >= 16; run -= 16) put(ezrl >> 5, ezrl & 31);
                        const uint32_t e = lds32(huff_ac_s + (uint32_t)((run << 4) | cb) * 4);
                        const uint32_t mant = (uint32_t)(cv + (cv >> 31)) & ((1u << cb) - 1u);
                        put(((e >> 5) << cb) | mant, (e & 31) + (uint32_t)cb);
                    }
                };
                ac_run(mask_lo, 0);
                ac_run(mask_hi, 32);
                if (prevk != 63) put(eeob >> 5, eeob & 31);                 // EOB only if last_index < 63 (:432-434)
                if (fill > 0) sts32(wp, acc);
                len = ((wp - stage_s) >> 7) * 32u + fill;
            }
            W.lens[sigma] = len;               // lanes 30/31 and inactive blocks write 0 (slot 31 is a dummy)
            __syncwarp();

            // ---------------- C: exclusive scan of the block lengths in bitstream order (lane = sigma)
            const uint32_t mylen_sigma = lane < kSegBlocks ? W.lens[lane] : 0;
            const uint32_t inc_sigma = warp_incl_scan(mylen_sigma, lane);
            const uint32_t T = __shfl_sync(0xffffffffu, inc_sigma, 31);
            __syncwarp();
            W.lens[lane] = inc_sigma - mylen_sigma;
            // DC predictors for the next segment (everyone read the old ones before the last __syncwarp)
            if (active && mi == nmb - 1 && (b == 3 || b >= 4)) W.carry_dc[comp] = dc;
            uint32_t R = r + T;                              // bits in the buffer after this segment
            // clear the words this segment will OR into; word 0 starts with the carried bits
            const uint32_t used_words = (R + 7 + 31) >> 5;
            for (uint32_t i = 1 + lane; i <= used_words; i += 32) W.u.seg[i] = 0;
            if (lane == 0) W.u.seg[0] = carry_word;
            __syncwarp();

            // ---------------- D: bit packer -- shift the private string to its scanned bit offset
            if (active && len) {
                const uint32_t o = r + W.lens[sigma];
                const uint32_t sh = o & 31;
                uint32_t dst = seg_s + (o >> 5) * 4, src = stage_s;
                const uint32_t nsrc = (len + 31) >> 5;
                uint32_t prev = 0;
                for (uint32_t j = 0; j < nsrc; j++) {
                    const uint32_t v = lds32(src);
                    red_or_shared(dst, __funnelshift_r(v, prev, sh));      // (prev:v) >> sh
                    prev = v; src += 128; dst += 4;
                }
                const uint32_t tail = sh ? prev << (32 - sh) : 0u;
                if (tail) red_or_shared(dst, tail);
            }
            __syncwarp();
            if (last_seg) {
                // pad to a byte with ones (ff_mjpeg_encode_stuffing, mjpegenc.c:338-343)
                const uint32_t pad = (0u - R) & 7u;
                if (lane == 0 && pad) W.u.seg[R >> 5] |= ((1u << pad) - 1u) << (32 - (R & 31) - pad);
                R += pad;
                __syncwarp();
            }

            // ---------------- E: FF00 stuffing + output of the complete bytes
            const uint32_t B = last_seg ? (R >> 3) : ((R >> 5) << 2);     // bytes leaving the buffer now
            const uint32_t nw = (B + 3) >> 2;
            const uint32_t per = (nw + 31) >> 5;
            const uint32_t w0 = min((uint32_t)lane * per, nw), w1 = min(w0 + per, nw);
            // bytes past B in the last word are zero bits, never FF: no masking needed for the count
            uint32_t ffc = 0;
            for (uint32_t w = w0; w < w1; w++) ffc += __popc(ff_bytes(W.u.seg[w]));
            const uint32_t inc = warp_incl_scan(ffc, lane);
            const uint32_t ff_total = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t ff_before = inc - ffc;
            const bool fits = (uint64_t)G + B + ff_total + 2 <= pkt_cap && !overflow;
            if (fits) {
                uint8_t *o = pkt + G + w0 * 4 + ff_before;
                for (uint32_t w = w0; w < w1; w++) {
                    const uint32_t v = W.u.seg[w];
                    const uint32_t nvalid = min(4u, B - w * 4);
                    if (nvalid == 4 && ff_bytes(v) == 0) {
                        o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v;
                        o += 4;
                    } else {
                        for (uint32_t k = 0; k < nvalid; k++) {
                            const uint8_t by = (uint8_t)(v >> (24 - 8 * k));
                            *o++ = by;
                            if (by == 0xff) *o++ = 0;
                        }
                    }
                }
            } else if (!overflow) overflow = AMV_ST_NOSPACE;
            carry_word = last_seg ? 0 : W.u.seg[R >> 5];                  // carry the partial word (same for all lanes)
            G += B + ff_total;
            r = last_seg ? 0 : (R & 31);
            __syncwarp();
        }
        if (lane == 0) {
            if (!overflow) { pkt[G] = 0xff; pkt[G + 1] = 0xd9; }      // EOI (mjpegenc.c:354)
            out_size[f] = overflow ? 0 : G + 2;
            status[f] = (int32_t)overflow;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_encode16: the same encoder with HOMOGENEOUS rounds.  k_encode puts 20 luma and 10 chroma blocks of
// 5 macroblocks on one warp; a luma block of ordinary content carries ~17 non-zero coefficients, a chroma
// block ~2, so its Huffman loop runs with 13 of 32 lanes busy.  Here a segment is 16 macroblocks done in
// three rounds of 32 like blocks: all 32 chroma blocks (16 Cb, 16 Cr), then the 32 luma blocks of
// macroblocks 0..7, then those of 8..15; each luma round scans, packs and outputs its half of the segment
// together with the chroma strings of the same macroblocks, which wait in their own staging columns.
// The strings are staged at sizes that fit ordinary content (768 bits per luma block, 384 per chroma
// block; the worst case is 1658) so the working set stays what it was: a frame with a longer block is
// not finished here -- it is flagged in `redo` and encoded by k_encode right after (launch_encode).
// ------------------------------------------------------------------------------------------------
constexpr int kWL = 24, kWC = 12;                               // staging words per luma / chroma block
constexpr int kSeg16Words = 32 * kWL + 16 * kWC + 8;            // half a segment: 32 luma + 16 chroma strings

struct Enc16WarpSmem {
    union {                                 // never live at the same time (A,B use coef; D,E use seg)
        uint16_t coef[64 * 32];             // halfword k*32 + lane : zigzag coefficient k of the lane's block
        uint32_t seg[kSeg16Words];          // the half segment's contiguous bit string
    } u;
    uint32_t stageL[kWL * 32];              // word w*32 + lane : the lane's luma string of this round
    uint32_t stageC[kWC * 32];              // the lane's chroma string of this segment
    uint32_t qm10[64];
    uint32_t lenC[32];                      // chroma string lengths: Cb of macroblock j at j, Cr at 16 + j
    uint32_t cpre[20];                      // exclusive prefix over the macroblocks of lenCb + lenCr
    int      carry_dc[4];                   // last DC of each component so far
};
struct Enc16Smem {
    uint32_t huff[kEncHuffEntries];
    Enc16WarpSmem w[kEncWarps];
};

template <bool FAST>
__global__ void __launch_bounds__(kEncThreads)
k_encode16(const uint8_t *__restrict__ py, const uint8_t *__restrict__ pu, const uint8_t *__restrict__ pv,
           int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int n, Geom g, const int32_t *__restrict__ qscale,
           uint8_t *__restrict__ slots, uint64_t slot_stride, uint32_t pkt_cap, uint32_t *__restrict__ out_size,
           int32_t *__restrict__ status, int32_t *__restrict__ redo) {
    AMV_EXTERN_SHARED(uint8_t, smem_raw, 16);
    Enc16Smem &S = *reinterpret_cast<Enc16Smem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kEncHuffEntries; i += kEncThreads) S.huff[i] = g_enc_tables.huff.e[i];
    __syncthreads();
    Enc16WarpSmem &W = S.w[wid];
    const int total_mb = g.mbw * g.mbh;
    const uint32_t coef_s = smem_addr(&W.u.coef[lane]);          // coefficient k: + k*64
    const uint32_t stageL_s = smem_addr(&W.stageL[lane]), stageC_s = smem_addr(&W.stageC[lane]);   // word w: + w*128
    const uint32_t seg_s = smem_addr(&W.u.seg[0]);
    const int gw = blockIdx.x * kEncWarps + wid, nw_total = gridDim.x * kEncWarps;

    for (int f = gw; f < n; f += nw_total) {
        const int qs = qscale ? qscale[f] : 2;
        __syncwarp();
        for (int t = lane; t < 64; t += 32) {
            // intra_matrix / q_intra_matrix for this frame (mpegvideo_enc.c:2866-2877, ff_convert_matrix :69-91)
            int m = 8;
            if (t) m = min(max(((int)g_enc_tables.intra_base[t] * qs) >> 3, 1), 255);
            W.qm10[t] = ((1u << 22) / (uint32_t)(8 * m)) << 10;
        }
        if (lane < 3) W.carry_dc[lane] = 128;            // last_dc init (mpegvideo_enc.c:2033-2036)
        uint32_t overflow = (qs < 2 || qs > 31) ? AMV_ST_RANGE : 0;      // qscale domain: SURVEY 9.13
        uint32_t carry_word = 0;                        // partial word carried into the next half segment
        uint8_t *pkt = slots + (uint64_t)f * slot_stride;
        uint32_t G = 2;                                 // bytes written so far (SOI)
        uint32_t r = 0;                                 // carried bits (sit in carry_word, MSB side)
        bool too_big = false;                           // a string outgrew its staging column (warp-uniform)
        if (lane == 0 && pkt_cap >= 2) { pkt[0] = 0xff; pkt[1] = 0xd8; }
        __syncwarp();

        for (int m0 = 0; m0 < total_mb && !too_big; m0 += 16) {
            const int nmb = min(16, total_mb - m0);
            const bool last_seg = m0 + 16 >= total_mb;
            uint32_t lenC_own = 0;                       // this lane's chroma string of the segment

            for (int rd = 0; rd < 3 && !too_big; rd++) {
                const int hh = rd - 1;                   // which half of the segment a luma round covers
                if (rd && 8 * hh >= nmb) break;
                // ---- the lane's block in this round
                const int comp = rd ? 0 : 1 + (lane >> 4);
                const int mbi = rd ? 8 * hh + (lane >> 2) : (lane & 15);
                const int b = rd ? (lane & 3) : 0;
                const bool active = mbi < nmb;
                const int vw = comp ? (g.w >> 1) : g.w, vh = comp ? (g.h >> 1) : g.h, r0 = comp ? g.c0 : g.y0;
                const int ls = comp ? ls_c : ls_y;
                const uint8_t *pl = (comp == 0 ? py + (uint64_t)f * fs_y : (comp == 1 ? pu : pv) + (uint64_t)f * fs_c);
                const uint32_t huff_dc_s = smem_addr(&S.huff[comp ? kEncDcChroma : kEncDcLuma]);
                const uint32_t huff_ac_s = smem_addr(&S.huff[comp ? kEncAcChroma : kEncAcLuma]);
                const uint32_t stage_s = rd ? stageL_s : stageC_s;
                const uint32_t cap_words = rd ? kWL : kWC;
                uint32_t mask_lo = 0, mask_hi = 0;          // non-zero AC positions (zigzag) of this block
                int dc = 0;

                // ---------------- A: load, FDCT, quantise
                if (active) {
                    const int mb = m0 + mbi;
                    const int mx = mb % g.mbw, my = mb / g.mbw;
                    const int bx = comp ? mx * 8 : mx * 16 + (b & 1) * 8;
                    const int by = comp ? my * 8 : my * 16 + (b >> 1) * 8;
                    int v[64];
                    uint32_t px[16];                                        // the pixel rows as packed bytes
#pragma unroll
                    for (int yy = 0; yy < 8; yy++) {
                        const int Y = min(by + yy, vh - 1);                 // bottom edge replication
                        const uint8_t *row = pl + (int64_t)(r0 - Y) * ls;
                        if (FAST) {
                            const uint2 q = *reinterpret_cast<const uint2 *>(row + bx);
                            px[2 * yy] = q.x; px[2 * yy + 1] = q.y;
                        } else {
                            uint32_t q[2] = { 0u, 0u };
#pragma unroll
                            for (int xx = 0; xx < 8; xx++) q[xx >> 2] |= (uint32_t)row[min(bx + xx, vw - 1)] << (8 * (xx & 3));   // right edge replication
                            px[2 * yy] = q[0]; px[2 * yy + 1] = q[1];
                        }
                    }
                    fdct_block_px<0>(px, v);
                    dc = quant_dc(v[0]);
                    // raster order so the multiplier loads vectorise; mask bit = zigzag position
#pragma unroll
                    for (int j = 1; j < 64; j++) {
                        const int q = quant_ac(v[j], W.qm10[j]);
                        v[j] = q;
                        const int k = zigzag_inv_at(j);
                        if (q) { if (k < 32) mask_lo |= 1u << k; else mask_hi |= 1u << (k - 32); }
                    }
#pragma unroll
                    for (int k = 1; k < 64; k++) W.u.coef[k * 32 + lane] = (uint16_t)v[zigzag_at(k)];
                }
                // DC predictor: the previous block of the component is the previous lane (Y0..Y3 of a macroblock and
                // the macroblocks themselves are consecutive lanes; Cb and Cr each fill 16 consecutive lanes)
                const int dc_prev = __shfl_up_sync(0xffffffffu, dc, 1);
                const bool first_of_comp = rd ? lane == 0 : (lane & 15) == 0;
                const int pred = first_of_comp ? W.carry_dc[comp] : dc_prev;
                __syncwarp();

                // ---------------- B: Huffman-code the block into the lane's private bit string
                uint32_t len = 0;
                if (active) {
                    uint32_t acc = 0;                   // MSB-first accumulator: the top `fill` bits are valid
                    uint32_t fill = 0;
                    uint32_t wp = stage_s;              // next private word
                    const uint32_t wend = stage_s + cap_words * 128u;
                    auto put = [&](uint32_t code, uint32_t nbits) {  // 1 <= nbits <= 27
                        const uint32_t t = code << (32 - nbits);     // left-aligned
                        acc |= t >> fill;
                        const uint32_t nf = fill + nbits;
                        if (nf >= 32) { if (wp < wend) sts32(wp, acc); wp += 128; acc = t << (32 - fill); fill = nf - 32; }   // fill >= 5 here
                        else fill = nf;
                    };
                    {   // DC (ff_mjpeg_encode_dc, mjpegenc.c:357-377)
                        const int diff = dc - pred;
                        const int nb = bit_width((uint32_t)(diff < 0 ? -diff : diff));
                        const uint32_t e = lds32(huff_dc_s + nb * 4);
                        const uint32_t mant = (uint32_t)(diff + (diff >> 31)) & ((1u << nb) - 1u);
                        put(((e >> 5) << nb) | mant, (e & 31) + (uint32_t)nb);
                    }
                    const uint32_t ezrl = lds32(huff_ac_s + 0xf0 * 4), eeob = lds32(huff_ac_s);
                    int prevk = 0;
                    auto ac_run = [&](uint32_t m, int base) {        // encode_block's AC loop (mjpegenc.c:403-430)
                        while (m) {
                            const int k = base + __ffs((int)m) - 1;
                            m &= m - 1;
                            int run = k - prevk - 1;
                            prevk = k;
                            const int cv = lds_s16(coef_s + (uint32_t)k * 64);
                            const int cb = bit_width((uint32_t)(cv < 0 ? -cv : cv));
                            for (; run >= 16; run -= 16) put(ezrl >> 5, ezrl & 31);
                            const uint32_t e = lds32(huff_ac_s + (uint32_t)((run << 4) | cb) * 4);
                            const uint32_t mant = (uint32_t)(cv + (cv >> 31)) & ((1u << cb) - 1u);
                            put(((e >> 5) << cb) | mant, (e & 31) + (uint32_t)cb);
                        }
                    };
                    ac_run(mask_lo, 0);
                    ac_run(mask_hi, 32);
                    if (prevk != 63) put(eeob >> 5, eeob & 31);                 // EOB only if last_index < 63 (:432-434)
                    if (fill > 0 && wp < wend) sts32(wp, acc);
                    len = ((wp - stage_s) >> 7) * 32u + fill;
                }
                if (__any_sync(0xffffffffu, len > cap_words * 32u)) { too_big = true; break; }
                // DC predictors for what follows: the last active block of each component in this round
                {
                    const int na = rd ? min(32, 4 * (nmb - 8 * hh)) : min(16, nmb);       // active lanes (per component)
                    if (active && (rd ? lane : (lane & 15)) == na - 1) W.carry_dc[comp] = dc;
                }
                if (rd == 0) {
                    // chroma round: keep the strings, publish their lengths and the per-macroblock prefix
                    lenC_own = len;
                    W.lenC[lane] = len;
                    __syncwarp();
                    const uint32_t cj = lane < 16 ? W.lenC[lane] + W.lenC[16 + lane] : 0u;
                    const uint32_t cinc = warp_incl_scan(cj, lane);
                    if (lane < 16) W.cpre[lane + 1] = cinc;
                    if (lane == 0) W.cpre[0] = 0;
                    __syncwarp();
                    continue;
                }

                // ---------------- C: bit offsets of the half segment's strings in bitstream order
                // (per macroblock: Y0 Y1 Y2 Y3 Cb Cr)
                const uint32_t linc = warp_incl_scan(len, lane);
                const uint32_t cbase = W.cpre[8 * hh];
                const uint32_t luma_off = (linc - len) + (W.cpre[8 * hh + (lane >> 2)] - cbase);
                // chroma strings of this half are packed by the lanes that made them: Cb of macroblock 8*hh + j by
                // lane 8*hh + j, Cr by lane 16 + 8*hh + j
                const int cj = lane & 7;
                const uint32_t luma_end = __shfl_sync(0xffffffffu, linc, 4 * cj + 3);
                const bool cpack = (((lane & 15) >> 3) == hh) && ((lane & 15) < nmb);
                const uint32_t chroma_off = luma_end + (W.cpre[8 * hh + cj] - cbase) + (lane >= 16 ? W.lenC[lane - 16] : 0u);
                const uint32_t T = __shfl_sync(0xffffffffu, linc, 31) + (W.cpre[min(8 * hh + 8, 16)] - cbase);
                const bool last_half = last_seg && (hh == 1 || nmb <= 8);
                __syncwarp();
                uint32_t R = r + T;                              // bits in the buffer after this half
                // clear the words this half will OR into; word 0 starts with the carried bits
                const uint32_t used_words = (R + 7 + 31) >> 5;
                for (uint32_t i = 1 + lane; i <= used_words; i += 32) W.u.seg[i] = 0;
                if (lane == 0) W.u.seg[0] = carry_word;
                __syncwarp();

                // ---------------- D: bit packer -- shift the private strings to their scanned bit offsets
                {
                    auto pack = [&](uint32_t src, uint32_t nbits, uint32_t o) {
                        const uint32_t sh = o & 31;
                        uint32_t dst = seg_s + (o >> 5) * 4;
                        const uint32_t nsrc = (nbits + 31) >> 5;
                        uint32_t prev = 0;
                        for (uint32_t j = 0; j < nsrc; j++) {
                            const uint32_t v = lds32(src);
                            red_or_shared(dst, __funnelshift_r(v, prev, sh));      // (prev:v) >> sh
                            prev = v; src += 128; dst += 4;
                        }
                        const uint32_t tail = sh ? prev << (32 - sh) : 0u;
                        if (tail) red_or_shared(dst, tail);
                    };
                    if (len) pack(stageL_s, len, r + luma_off);
                    if (cpack && lenC_own) pack(stageC_s, lenC_own, r + chroma_off);
                }
                __syncwarp();
                if (last_half) {
                    // pad to a byte with ones (ff_mjpeg_encode_stuffing, mjpegenc.c:338-343)
                    const uint32_t pad = (0u - R) & 7u;
                    if (lane == 0 && pad) W.u.seg[R >> 5] |= ((1u << pad) - 1u) << (32 - (R & 31) - pad);
                    R += pad;
                    __syncwarp();
                }

                // ---------------- E: FF00 stuffing + output of the complete bytes
                const uint32_t B = last_half ? (R >> 3) : ((R >> 5) << 2);     // bytes leaving the buffer now
                const uint32_t nw = (B + 3) >> 2;
                const uint32_t per = (nw + 31) >> 5;
                const uint32_t w0 = min((uint32_t)lane * per, nw), w1 = min(w0 + per, nw);
                // bytes past B in the last word are zero bits, never FF: no masking needed for the count
                uint32_t ffc = 0;
                for (uint32_t w = w0; w < w1; w++) ffc += __popc(ff_bytes(W.u.seg[w]));
                const uint32_t inc = warp_incl_scan(ffc, lane);
                const uint32_t ff_total = __shfl_sync(0xffffffffu, inc, 31);
                const uint32_t ff_before = inc - ffc;
                const bool fits = (uint64_t)G + B + ff_total + 2 <= pkt_cap && !overflow;
                if (fits) {
                    uint8_t *o = pkt + G + w0 * 4 + ff_before;
                    for (uint32_t w = w0; w < w1; w++) {
                        const uint32_t v = W.u.seg[w];
                        const uint32_t nvalid = min(4u, B - w * 4);
                        if (nvalid == 4 && ff_bytes(v) == 0) {
                            o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v;
                            o += 4;
                        } else {
                            for (uint32_t k = 0; k < nvalid; k++) {
                                const uint8_t by = (uint8_t)(v >> (24 - 8 * k));
                                *o++ = by;
                                if (by == 0xff) *o++ = 0;
                            }
                        }
                    }
                } else if (!overflow) overflow = AMV_ST_NOSPACE;
                carry_word = last_half ? 0 : W.u.seg[R >> 5];                 // carry the partial word (same for all lanes)
                G += B + ff_total;
                r = last_half ? 0 : (R & 31);
                __syncwarp();
            }
        }
        if (lane == 0) {
            redo[f] = too_big ? 1 : 0;
            if (!too_big) {
                if (!overflow) { pkt[G] = 0xff; pkt[G + 1] = 0xd9; }      // EOI (mjpegenc.c:354)
                out_size[f] = overflow ? 0 : G + 2;
                status[f] = (int32_t)overflow;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_encode16v2: k_encode16 with three cuts in the instruction stream (same rounds, same strings, same bytes).
//  * The quantiser is not evaluated for the 63 AC coefficients of every block, only for the ones that survive it:
//    a coefficient quantises to non-zero iff |c| * qmat >= 2^22, i.e. iff c * c >= T * T with T = ceil(2^22 / qmat),
//    so the transform stage spends one multiply-add (c * c - T^2) and one funnel shift per coefficient to build the
//    zigzag mask of survivors, stores the RAW coefficient, and the Huffman loop quantises what it is about to code
//    (dct_quantize_c, mpegvideo_enc.c:3686-3716: level = |c| * qmat >> 22, sign restored).
//  * The bit writer of the Huffman loop has no branch: the word under construction is stored every time and the
//    pointer advances by a select.
//  * Stuffed bytes are staged in shared memory (the luma staging columns are free by then) and leave as aligned
//    128-bit stores; the bytes short of a 16-byte unit wait in registers (one per lane) for the next half segment.
//    The packet slot must be 16-byte aligned (launch_encode checks; else the one-kernel path runs).
// ------------------------------------------------------------------------------------------------
struct Enc16v2WarpSmem {
    union {                                 // never live at the same time (A,B use coef; D,E use seg)
        uint16_t coef[64 * 32];             // halfword k*32 + lane : RAW zigzag coefficient k of the lane's block
        uint32_t seg[kSeg16Words];          // the half segment's contiguous bit string
    } u;
    union {
        uint32_t stageL[kWL * 32];          // word w*32 + lane : the lane's luma string of this round
        uint8_t  obuf[kWL * 32 * 4];        // stage E: the stuffed bytes of the half segment, in packet order
    } s;
    uint32_t stageC[kWC * 32];              // the lane's chroma string of this segment
    int32_t  nthr2z[64];                    // zigzag k: -(T*T), T = smallest |coefficient| that quantises to non-zero
    uint32_t qm10z[64];                     // zigzag k: qmat << 10
    uint32_t lenC[32];                      // chroma string lengths: Cb of macroblock j at j, Cr at 16 + j
    uint32_t cpre[20];                      // exclusive prefix over the macroblocks of lenCb + lenCr
    int      carry_dc[4];                   // last DC of each component so far
};
struct Enc16v2Smem {
    uint32_t huff[kEncHuffEntries];
    Enc16v2WarpSmem w[kEncWarps];
};

template <bool FAST, int MINB, int FD>
__global__ void __launch_bounds__(kEncThreads, MINB)
k_encode16v2(const uint8_t *__restrict__ py, const uint8_t *__restrict__ pu, const uint8_t *__restrict__ pv,
             int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int n, Geom g, const int32_t *__restrict__ qscale,
             uint8_t *__restrict__ slots, uint64_t slot_stride, uint32_t pkt_cap, uint32_t *__restrict__ out_size,
             int32_t *__restrict__ status, int32_t *__restrict__ redo) {
    AMV_EXTERN_SHARED(uint8_t, smem_raw, 16);
    Enc16v2Smem &S = *reinterpret_cast<Enc16v2Smem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kEncHuffEntries; i += kEncThreads) {
        // (code << 5 | length) -> code left-aligned in the word, length in the low five bits (codes have at most 16 bits)
        const uint32_t e = g_enc_tables.huff.e[i], l = e & 31u;
        S.huff[i] = l ? ((e >> 5) << (32u - l)) | l : 0u;
    }
    __syncthreads();
    Enc16v2WarpSmem &W = S.w[wid];
    const int total_mb = g.mbw * g.mbh;
    const uint32_t coef_s = smem_addr(&W.u.coef[lane]);          // coefficient k: + k*64
    const uint32_t stageL_s = smem_addr(&W.s.stageL[lane]), stageC_s = smem_addr(&W.stageC[lane]);   // word w: + w*128
    const uint32_t seg_s = smem_addr(&W.u.seg[0]);
    const uint32_t qm_s = smem_addr(&W.qm10z[0]);
    const uint32_t nthr_s = smem_addr(&W.nthr2z[0]);
    const uint32_t obuf_s = smem_addr(&W.s.obuf[0]);
    const int gw = blockIdx.x * kEncWarps + wid, nw_total = gridDim.x * kEncWarps;
    int cur_qs = -1;
    uint32_t px[16];                                              // the block's pixel rows as packed bytes
    // macroblock index -> row by one high multiply: ceil(2^32 / mbw) is exact for mb * mbw < 2^32 (mbw == 1: row = mb)
    const uint32_t mbw_magic = g.mbw > 1 ? 0xffffffffu / (uint32_t)g.mbw + 1u : 0u;
    // The pixel rows of this lane's block in round rrd of the segment at macroblock mm0 of frame ff, as loaded (two words per
    // row, column 0 in the lowest byte): bottom-up addressing (amv_encode_picture mjpegenc.c:454-472) and edge replication by
    // clamped indices (mpegvideo.c:1416-1470).  Lanes without a block in that round leave q alone.
    auto fetch = [&](int ff, int mm0, int rrd, uint32_t (&q)[16]) {
        const int nmb_ = min(16, total_mb - mm0);
        const int comp_ = rrd ? 0 : 1 + (lane >> 4);
        const int mbi_ = rrd ? 8 * (rrd - 1) + (lane >> 2) : (lane & 15);
        const int b_ = rrd ? (lane & 3) : 0;
        if (mbi_ >= nmb_) return;
        const int vw_ = comp_ ? (g.w >> 1) : g.w, vh_ = comp_ ? (g.h >> 1) : g.h, r0_ = comp_ ? g.c0 : g.y0;
        const int ls_ = comp_ ? ls_c : ls_y;
        const uint8_t *pl_ = (comp_ == 0 ? py + (uint64_t)ff * fs_y : (comp_ == 1 ? pu : pv) + (uint64_t)ff * fs_c);
        const int mb = mm0 + mbi_;
        const int my = g.mbw > 1 ? (int)__umulhi((uint32_t)mb, mbw_magic) : mb, mx = mb - my * g.mbw;
        const int bx = comp_ ? mx * 8 : mx * 16 + (b_ & 1) * 8;
        const int by = comp_ ? my * 8 : my * 16 + (b_ >> 1) * 8;
#pragma unroll
        for (int yy = 0; yy < 8; yy++) {
            const int Y = min(by + yy, vh_ - 1);                 // bottom edge replication
            const uint8_t *row = pl_ + (int64_t)(r0_ - Y) * ls_;
            if (FAST) {
                const uint2 w2 = *reinterpret_cast<const uint2 *>(row + bx);
                q[2 * yy] = w2.x; q[2 * yy + 1] = w2.y;
            } else {
                uint32_t w2[2] = { 0u, 0u };
#pragma unroll
                for (int xx = 0; xx < 8; xx++) w2[xx >> 2] |= (uint32_t)row[min(bx + xx, vw_ - 1)] << (8 * (xx & 3));   // right edge replication
                q[2 * yy] = w2[0]; q[2 * yy + 1] = w2[1];
            }
        }
    };

    for (int f = gw; f < n; f += nw_total) {
        const int qs = qscale ? qscale[f] : 2;
        __syncwarp();
        if (qs != cur_qs) {
            for (int k = lane; k < 64; k += 32) {
                // intra_matrix / q_intra_matrix for this frame (mpegvideo_enc.c:2866-2877, ff_convert_matrix :69-91)
                int m = 8;
                if (k) m = min(max(((int)g_enc_tables.intra_base[g_enc_tables.zigzag[k]] * qs) >> 3, 1), 255);
                const uint32_t qmat = (1u << 22) / (uint32_t)(8 * m);
                const uint32_t T = ((1u << 22) + qmat - 1u) / qmat;
                W.qm10z[k] = qmat << 10;
                W.nthr2z[k] = -(int32_t)(T * T);
            }
            cur_qs = qs;
        }
        if (lane < 3) W.carry_dc[lane] = 128;            // last_dc init (mpegvideo_enc.c:2033-2036)
        uint32_t overflow = (qs < 2 || qs > 31) ? AMV_ST_RANGE : 0;      // qscale domain: SURVEY 9.13
        uint32_t carry_word = 0;                        // partial word carried into the next half segment
        uint8_t *pkt = slots + (uint64_t)f * slot_stride;
        uint32_t G = 2;                                 // bytes produced so far (SOI); the last `oc` of them wait in `cbyte`
        uint32_t oc = 2;                                // bytes short of a 16-byte unit, lane i holds byte i
        uint32_t cbyte = lane == 0 ? 0xffu : 0xd8u;     // SOI (mjpegenc.c:197-204)
        uint32_t r = 0;                                 // carried bits (sit in carry_word, MSB side)
        bool too_big = false;                           // a string outgrew its staging column (warp-uniform)
        __syncwarp();

        for (int m0 = 0; m0 < total_mb && !too_big; m0 += 16) {
            const int nmb = min(16, total_mb - m0);
            const bool last_seg = m0 + 16 >= total_mb;
            uint32_t lenC_own = 0;                       // this lane's chroma string of the segment

            for (int rd = 0; rd < 3 && !too_big; rd++) {
                const int hh = rd - 1;                   // which half of the segment a luma round covers
                if (rd && 8 * hh >= nmb) break;
                // ---- the lane's block in this round
                const int comp = rd ? 0 : 1 + (lane >> 4);
                const int mbi = rd ? 8 * hh + (lane >> 2) : (lane & 15);
                const int b = rd ? (lane & 3) : 0;
                const bool active = mbi < nmb;
                const uint32_t huff_dc_s = smem_addr(&S.huff[comp ? kEncDcChroma : kEncDcLuma]);
                const uint32_t huff_ac_s = smem_addr(&S.huff[comp ? kEncAcChroma : kEncAcLuma]);
                const uint32_t stage_s = rd ? stageL_s : stageC_s;
                const uint32_t cap_words = rd ? kWL : kWC;
                uint32_t mask_lo = 0, mask_hi = 0;          // AC positions (zigzag) that quantise to non-zero
                int dc = 0;

                // ---------------- A: load, FDCT, survivor mask
                // (the rows are fetched, and the transform is closed, outside the branch the mask is built in: in this shape ptxas
                // keeps the stage free of spills at 96 registers -- worth 0.8 ms per 100 000 frames)
                fetch(f, m0, rd, px);
                int v[64];
                if (active) fdct_block_px<FD>(px, v);
                if (active) {
                    dc = quant_dc(v[0]);
                    // c * c - T * T is negative exactly for the coefficients that quantise to zero: its sign bit is shifted
                    // into the (inverted) mask, lowest zigzag position first, so that position k ends up at bit 31 - (k & 31): the
                    // Huffman loop finds its next coefficient with one count of leading zeros
                    uint32_t inv_lo = 0, inv_hi = 0;
#pragma unroll
                    for (int k4 = 0; k4 < 16; k4++) {
                        const uint4 t4 = lds128(nthr_s + 16 * k4);       // -(T*T) of positions 4*k4 .. 4*k4 + 3
                        const int nt[4] = { (int)t4.x, (int)t4.y, (int)t4.z, (int)t4.w };
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            const int k = 4 * k4 + kk;
                            if (k == 0) continue;                         // the DC
                            const int c = v[zigzag_at(k)];
                            if (k >= 32) inv_hi = __funnelshift_l((uint32_t)(c * c + nt[kk]), inv_hi, 1);
                            else         inv_lo = __funnelshift_l((uint32_t)(c * c + nt[kk]), inv_lo, 1);
                            W.u.coef[k * 32 + lane] = (uint16_t)c;
                        }
                    }
                    mask_lo = ~inv_lo & 0x7fffffffu;            // 31 positions went in: bit 31 (position 0, the DC) stays clear
                    mask_hi = ~inv_hi;
                }
                // DC predictor: the previous block of the component is the previous lane (Y0..Y3 of a macroblock and
                // the macroblocks themselves are consecutive lanes; Cb and Cr each fill 16 consecutive lanes)
                const int dc_prev = __shfl_up_sync(0xffffffffu, dc, 1);
                const bool first_of_comp = rd ? lane == 0 : (lane & 15) == 0;
                const int pred = first_of_comp ? W.carry_dc[comp] : dc_prev;
                __syncwarp();

                // ---------------- B: Huffman-code the block into the lane's private bit string
                uint32_t len = 0;
                if (active) {
                    uint32_t acc = 0;                   // MSB-first accumulator: the top `fill` bits are valid
                    uint32_t fill = 0;
                    uint32_t wp = stage_s;              // the word under construction
                    const uint32_t wlast = stage_s + (cap_words - 1u) * 128u;
                    // branch-free: the word under construction is stored every time (a string that outgrows its column
                    // keeps rewriting the last word and is caught by its length), the pointer moves by a select.
                    // t: the code and its mantissa, left-aligned; 1 <= nbits <= 27
                    // nm32 = nbits - 32, so that "the word is full" is a sign test and the caller's (code length - leading zeros) needs no + 32
                    auto put = [&](uint32_t t, int nm32) {
                        acc |= t >> fill;
                        const int nf = (int)fill + nm32;             // bits in the word after this symbol, - 32
                        sts32(min(wp, wlast), acc);
                        if (nf >= 0) { acc = __funnelshift_r(0u, t, fill); wp += 128u; }          // full, then fill >= 5: t << (32 - fill) as the low word of (t : 0) >> fill
                        fill = (uint32_t)nf & 31u;
                    };
                    // table entries: code left-aligned | length.  The mantissa (low `size` bits of x) goes right under the code:
                    // x << (32 - size) drops whatever sits above it, the funnel shift by the entry's low five bits moves it down
                    {   // DC (ff_mjpeg_encode_dc, mjpegenc.c:357-377): negative differences send diff - 1
                        const int diff = dc - pred;
                        const int nb = bit_width((uint32_t)(diff < 0 ? -diff : diff));
                        const uint32_t e = lds32(huff_dc_s + nb * 4);
                        const uint32_t y = __funnelshift_l(0u, (uint32_t)(diff + (diff >> 31)), 0u - (uint32_t)nb);    // nb == 0: diff == 0
                        put((e & ~31u) | __funnelshift_r(y, 0u, e), (int)(e & 31u) + nb - 32);
                    }
                    const uint32_t ezrl = lds32(huff_ac_s + 0xf0 * 4), eeob = lds32(huff_ac_s);
                    const uint32_t huff_ac_last_s = huff_ac_s + 4u + 31u * 4u;     // size = 32 - cz: entry index + 31 - cz
                    int nprev = -1;                                  // -(position of the last coded coefficient) - 1
                    auto ac_run = [&](uint32_t m, int base) {        // encode_block's AC loop (mjpegenc.c:403-430)
                        int k = base - 1;
                        while (m) {
                            const int z = clz_nz(m);                 // masks are bit-reversed: leading zeros = zero coefficients skipped
                            m = (m << z) << 1;
                            k += z + 1;
                            int run = k + nprev;
                            nprev = ~k;
                            const int raw = lds_s16(coef_s + (uint32_t)k * 64);
                            // level = |c| * qmat >> 22 (>= 1 here), sign restored (dct_quantize_c, mpegvideo_enc.c:3686-3716)
                            const uint32_t q = __umulhi((uint32_t)(raw < 0 ? -raw : raw), lds32(qm_s + (uint32_t)k * 4));
                            const int cz = clz_nz(q);                                         // 32 - size
                            for (; run >= 16; run -= 16) put(ezrl & ~31u, (int)(ezrl & 31u) - 32);
                            const uint32_t e = lds32(huff_ac_last_s + ((uint32_t)run << 6) - ((uint32_t)cz << 2));    // entry (run << 4) + size - 1
                            const uint32_t y = __funnelshift_l(0u, q ^ (uint32_t)(raw >> 31), (uint32_t)cz);   // negative: level - 1 = ~|level|; << (32 - size)
                            put((e & ~31u) | __funnelshift_r(y, 0u, e), (int)(e & 31u) - cz);
                        }
                    };
                    ac_run(mask_lo, 0);
                    ac_run(mask_hi, 32);
                    if (nprev != -64) put(eeob & ~31u, (int)(eeob & 31u) - 32);    // EOB only if last_index < 63 (:432-434)
                    if (fill > 0) sts32(min(wp, wlast), acc);
                    len = ((wp - stage_s) >> 7) * 32u + fill;
                }
                if (__any_sync(0xffffffffu, len > cap_words * 32u)) { too_big = true; break; }
                // DC predictors for what follows: the last active block of each component in this round
                {
                    const int na = rd ? min(32, 4 * (nmb - 8 * hh)) : min(16, nmb);       // active lanes (per component)
                    if (active && (rd ? lane : (lane & 15)) == na - 1) W.carry_dc[comp] = dc;
                }
                if (rd == 0) {
                    // chroma round: keep the strings, publish their lengths and the per-macroblock prefix
                    lenC_own = len;
                    W.lenC[lane] = len;
                    __syncwarp();
                    const uint32_t cj = lane < 16 ? W.lenC[lane] + W.lenC[16 + lane] : 0u;
                    const uint32_t cinc = warp_incl_scan(cj, lane);
                    if (lane < 16) W.cpre[lane + 1] = cinc;
                    if (lane == 0) W.cpre[0] = 0;
                    __syncwarp();
                    continue;
                }

                // ---------------- C: bit offsets of the half segment's strings in bitstream order
                // (per macroblock: Y0 Y1 Y2 Y3 Cb Cr)
                const uint32_t linc = warp_incl_scan(len, lane);
                const uint32_t cbase = W.cpre[8 * hh];
                const uint32_t luma_off = (linc - len) + (W.cpre[8 * hh + (lane >> 2)] - cbase);
                // chroma strings of this half are packed by the lanes that made them: Cb of macroblock 8*hh + j by
                // lane 8*hh + j, Cr by lane 16 + 8*hh + j
                const int cj = lane & 7;
                const uint32_t luma_end = __shfl_sync(0xffffffffu, linc, 4 * cj + 3);
                const bool cpack = (((lane & 15) >> 3) == hh) && ((lane & 15) < nmb);
                const uint32_t chroma_off = luma_end + (W.cpre[8 * hh + cj] - cbase) + (lane >= 16 ? W.lenC[lane - 16] : 0u);
                const uint32_t T = __shfl_sync(0xffffffffu, linc, 31) + (W.cpre[min(8 * hh + 8, 16)] - cbase);
                const bool last_half = last_seg && (hh == 1 || nmb <= 8);
                __syncwarp();
                uint32_t R = r + T;                              // bits in the buffer after this half
                // clear the words this half will OR into; word 0 starts with the carried bits
                const uint32_t used_words = (R + 7 + 31) >> 5;
                for (uint32_t i = 1 + lane; i <= used_words; i += 32) W.u.seg[i] = 0;
                if (lane == 0) W.u.seg[0] = carry_word;
                __syncwarp();

                // ---------------- D: bit packer -- shift the private strings to their scanned bit offsets
                {
                    auto pack = [&](uint32_t src, uint32_t nbits, uint32_t o) {
                        const uint32_t sh = o & 31;
                        uint32_t dst = seg_s + (o >> 5) * 4;
                        const uint32_t nsrc = (nbits + 31) >> 5;
                        uint32_t prev = 0;
                        for (uint32_t j = 0; j < nsrc; j++) {
                            const uint32_t v = lds32(src);
                            red_or_shared(dst, __funnelshift_r(v, prev, sh));      // (prev:v) >> sh
                            prev = v; src += 128; dst += 4;
                        }
                        const uint32_t tail = __funnelshift_r(0u, prev, sh);      // prev << (32 - sh), 0 for sh == 0: the low word of (prev : 0) >> sh
                        if (tail) red_or_shared(dst, tail);
                    };
                    if (len) pack(stageL_s, len, r + luma_off);
                    if (cpack && lenC_own) pack(stageC_s, lenC_own, r + chroma_off);
                }
                __syncwarp();
                if (last_half) {
                    // pad to a byte with ones (ff_mjpeg_encode_stuffing, mjpegenc.c:338-343)
                    const uint32_t pad = (0u - R) & 7u;
                    if (lane == 0 && pad) W.u.seg[R >> 5] |= ((1u << pad) - 1u) << (32 - (R & 31) - pad);
                    R += pad;
                    __syncwarp();
                }

                // ---------------- E: FF00 stuffing (escape_FF, mjpegenc.c:282-336) into the byte stage, 128-bit stores
                const uint32_t B = last_half ? (R >> 3) : ((R >> 5) << 2);     // bytes leaving the bit buffer now
                const uint32_t nw = (B + 3) >> 2;
                const uint32_t per = (nw + 31) >> 5;
                const uint32_t w0 = min((uint32_t)lane * per, nw), w1 = min(w0 + per, nw);
                // bytes past B in the last word are zero bits, never FF: no masking needed for the count
                uint32_t ffc = 0;
                for (uint32_t w = w0; w < w1; w++) ffc += __popc(ff_bytes(W.u.seg[w]));
                const uint32_t inc = warp_incl_scan(ffc, lane);
                const uint32_t ff_total = __shfl_sync(0xffffffffu, inc, 31);
                const uint32_t ff_before = inc - ffc;
                const uint32_t Bo = B + ff_total;                             // stuffed bytes of this half segment
                const uint32_t tail_bytes = last_half ? 2u : 0u;              // EOI
                if (oc + Bo + tail_bytes > sizeof(W.s.obuf)) { too_big = true; break; }      // never for ordinary content: k_encode takes the frame
                const bool fits = (uint64_t)G + Bo + 2 <= pkt_cap && !overflow;
                if (fits) {
                    // the luma strings have been packed: their columns now stage the output bytes
                    if ((uint32_t)lane < oc) W.s.obuf[lane] = (uint8_t)cbyte;
                    // branch-free per word: a word with an FF byte turns up in some lane in most iterations, so a byte-serial
                    // side path would run, one lane wide, almost every time (ncu: 74 % of the iterations in the first version)
                    uint32_t o = obuf_s + oc + w0 * 4 + ff_before;
                    for (uint32_t w = w0; w < w1; w++) {
                        const uint32_t v = W.u.seg[w];
                        const uint32_t nvalid = B - w * 4;                   // >= 4 except in the segment's last word
                        const uint32_t f = ff_bytes(v);                      // 0x80 in every byte that is FF
                        const uint32_t f0 = f >> 31, f1 = (f >> 23) & 1u, f2 = (f >> 15) & 1u, f3 = (f >> 7) & 1u;
                        const uint32_t p1 = 1u + f0, p2 = p1 + 1u + f1, p3 = p2 + 1u + f2;
                        sts8(o, v >> 24);
                        if (f0) sts8(o + 1u, 0u);
                        if (nvalid > 1u) { sts8(o + p1, v >> 16); if (f1) sts8(o + p1 + 1u, 0u); }
                        if (nvalid > 2u) { sts8(o + p2, v >> 8); if (f2) sts8(o + p2 + 1u, 0u); }
                        if (nvalid > 3u) { sts8(o + p3, v); if (f3) sts8(o + p3 + 1u, 0u); }
                        o += p3 + 1u + f3;                                   // only the last word can be short, and nothing follows it
                    }
                    uint32_t total = oc + Bo;
                    if (last_half && lane == 0) { W.s.obuf[total] = 0xff; W.s.obuf[total + 1] = 0xd9; }     // EOI (mjpegenc.c:354)
                    total += tail_bytes;
                    __syncwarp();
                    uint8_t *g0 = pkt + (G - oc);                             // 16-byte aligned: where obuf[0] goes
                    const uint32_t nfull = total >> 4;
                    for (uint32_t u = lane; u < nfull; u += 32)
                        *reinterpret_cast<uint4 *>(g0 + 16 * u) = lds128(obuf_s + 16 * u);
                    const uint32_t rem = total & 15u;
                    if (last_half) {
                        if ((uint32_t)lane < rem) g0[16 * nfull + lane] = W.s.obuf[16 * nfull + lane];
                        oc = 0;
                    } else {
                        cbyte = (uint32_t)lane < rem ? W.s.obuf[16 * nfull + lane] : 0u;
                        oc = rem;
                    }
                } else if (!overflow) overflow = AMV_ST_NOSPACE;
                carry_word = last_half ? 0 : W.u.seg[R >> 5];                 // carry the partial word (same for all lanes)
                G += Bo;
                r = last_half ? 0 : (R & 31);
                __syncwarp();
            }
        }
        if (lane == 0) {
            redo[f] = too_big ? 1 : 0;
            if (!too_big) {
                out_size[f] = overflow ? 0 : G + 2;
                status[f] = (int32_t)overflow;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed layout: copy each packet from its slot to its scanned offset
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_compact(const uint8_t *__restrict__ slots, uint64_t slot_stride, uint32_t *__restrict__ size,
          const uint64_t *__restrict__ off, int n, uint8_t *__restrict__ out, uint64_t out_cap,
          int32_t *__restrict__ status) {
    for (int f = blockIdx.x; f < n; f += gridDim.x) {
        const uint32_t sz = size[f];
        const uint64_t o = off[f];
        if (!range_ok(o, sz, out_cap)) {        // amvcuda.h: out_size is 0 wherever status is not (every thread has read size[f])
            __syncthreads();
            if (threadIdx.x == 0) { atomicOr(&status[f], AMV_ST_NOSPACE); size[f] = 0; }
            continue;
        }
        const uint8_t *src = slots + (uint64_t)f * slot_stride;
        uint8_t *dst = out + o;
        // head bytes up to 16-byte alignment of dst, then 128-bit stores fed by two aligned
        // source loads realigned with byte permutes, then the tail
        const uint32_t head = min(sz, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
        if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
        const uint32_t body = (sz - head) & ~15u;
        const uint32_t sh = head & 3;            // src is slot-aligned (>= 16 B): misalignment of src+head to 4 B
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(src + (head & ~3u));
        for (uint32_t i = threadIdx.x * 16; i < body; i += blockDim.x * 16) {
            const uint32_t *p = sw + (i >> 2);
            const uint32_t a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3], a4 = sh ? p[4] : 0;
            uint4 q;
            q.x = __funnelshift_r(a0, a1, 8 * sh); q.y = __funnelshift_r(a1, a2, 8 * sh);
            q.z = __funnelshift_r(a2, a3, 8 * sh); q.w = __funnelshift_r(a3, a4, 8 * sh);
            *reinterpret_cast<uint4 *>(dst + head + i) = q;
        }
        for (uint32_t i = head + body + threadIdx.x; i < sz; i += blockDim.x) dst[i] = src[i];
    }
}

// copies per-frame metadata to (mapped, pinned) host memory without touching a copy engine
__global__ void k_export_meta(const uint64_t *__restrict__ off, const uint32_t *__restrict__ sz, const int32_t *__restrict__ st,
                              uint64_t *__restrict__ hoff, uint32_t *__restrict__ hsz, int32_t *__restrict__ hst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (off) hoff[i] = off[i];
    if (sz) hsz[i] = sz[i];
    if (st) hst[i] = st[i];
}

__global__ void k_slot_offsets(uint64_t *off, int n, uint64_t stride, uint64_t base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] = base + (uint64_t)i * stride;
}

// ------------------------------------------------------------------------------------------------
cudaError_t upload_enc_tables(cudaStream_t s) {
    static EncTablesDev h;
    static bool built = false;
    if (!built) {
        build_enc_huff_tables(h.huff);
        memcpy(h.zigzag, kZigzag, 64);
        memcpy(h.intra_base, kEncIntraBase, 64);
        built = true;
    }
    return cudaMemcpyToSymbolAsync(g_enc_tables, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
}

// opt-in to more than 48 KB of dynamic shared memory: per device, called from amv_create
cudaError_t encode_setup_device() {
    cudaError_t e = cudaFuncSetAttribute(k_encode<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_encode<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_encode16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Enc16Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_encode16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Enc16Smem));
#define AMV_ENC16V2_ATTR(MINB, FD) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_encode16v2<true, MINB, FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Enc16v2Smem)); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_encode16v2<false, MINB, FD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Enc16v2Smem));
    AMV_ENC16V2_ATTR(5, 0) AMV_ENC16V2_ATTR(4, 0) AMV_ENC16V2_ATTR(5, 3) AMV_ENC16V2_ATTR(4, 3)
#undef AMV_ENC16V2_ATTR
    return e;
}

int encode_grid(int n, int per_sm) {
    // CTAs per SM: k_encode 4 (shared memory and registers: 16 independent warps), k_encode16 5 (registers)
    const int cap = kNumSMs * per_sm;
    const int need = (n + kEncWarps - 1) / kEncWarps;  // one frame per warp at a time
    return need < 1 ? 1 : (need < cap ? need : cap);
}

void launch_encode(const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                   int n, const Geom &g, const int32_t *qscale, uint8_t *slots, uint64_t slot_stride, uint32_t pkt_cap,
                   uint32_t *out_size, int32_t *status, int32_t *redo, int form, cudaStream_t s) {
    const bool fast = (g.w % 16 == 0) &&
                      ((((uintptr_t)y | (uintptr_t)u | (uintptr_t)v | (uintptr_t)ls_y | (uintptr_t)ls_c | fs_y | fs_c) & 7) == 0);
    // form 0 (or redo == nullptr): the plain one-kernel path.  Else a rounds kernel first (homogeneous rounds, strings staged at
    // ordinary-content sizes) -- form 2: k_encode16v2, needs 16-byte aligned packet slots; form 1: k_encode16 -- then
    // k_encode for the frames it flagged because a block's string outgrew its column.
    if (!redo) form = 0;
    if (form >= 2 && ((((uintptr_t)slots | slot_stride) & 15) != 0)) form = 1;
    if (form >= 2) {
        // form 4 (default): five CTAs per SM (96 registers), the transform regrouped for the two integer pipes (amv_dct.cuh:
        // dot-product rows on the packed pixel bytes, written-out odd columns); 2: the transform in its factorised form;
        // 3 / 8: forms 2 / 4 at four CTAs per SM (128 registers)
#define AMV_ENC16V2_GO(MINB, FD) do { \
        if (fast) AMV_LAUNCH((k_encode16v2<true, MINB, FD>), encode_grid(n, MINB), kEncThreads, sizeof(Enc16v2Smem), s, y, u, v, ls_y, ls_c, fs_y, fs_c, n, g, \
                             qscale, slots, slot_stride, pkt_cap, out_size, status, redo); \
        else      AMV_LAUNCH((k_encode16v2<false, MINB, FD>), encode_grid(n, MINB), kEncThreads, sizeof(Enc16v2Smem), s, y, u, v, ls_y, ls_c, fs_y, fs_c, n, g, \
                             qscale, slots, slot_stride, pkt_cap, out_size, status, redo); } while (0)
        switch (form) {
        case 2: AMV_ENC16V2_GO(5, 0); break;
        case 3: AMV_ENC16V2_GO(4, 0); break;
        case 8: AMV_ENC16V2_GO(4, 3); break;
        default: AMV_ENC16V2_GO(5, 3); break;
        }
#undef AMV_ENC16V2_GO
    } else if (form == 1) {
        if (fast) AMV_LAUNCH(k_encode16<true>, encode_grid(n, 5), kEncThreads, sizeof(Enc16Smem), s, y, u, v, ls_y, ls_c, fs_y, fs_c, n, g,
                             qscale, slots, slot_stride, pkt_cap, out_size, status, redo);
        else      AMV_LAUNCH(k_encode16<false>, encode_grid(n, 5), kEncThreads, sizeof(Enc16Smem), s, y, u, v, ls_y, ls_c, fs_y, fs_c, n, g,
                             qscale, slots, slot_stride, pkt_cap, out_size, status, redo);
    }
    if (fast) AMV_LAUNCH(k_encode<true>, encode_grid(n, 4), kEncThreads, sizeof(EncSmem), s, y, u, v, ls_y, ls_c, fs_y, fs_c, n, g, qscale,
                         slots, slot_stride, pkt_cap, out_size, status, form ? redo : nullptr);
    else      AMV_LAUNCH(k_encode<false>, encode_grid(n, 4), kEncThreads, sizeof(EncSmem), s, y, u, v, ls_y, ls_c, fs_y, fs_c, n, g, qscale,
                         slots, slot_stride, pkt_cap, out_size, status, form ? redo : nullptr);
}

void launch_compact(const uint8_t *slots, uint64_t slot_stride, uint32_t *size, const uint64_t *off, int n,
                    uint8_t *out, uint64_t out_cap, int32_t *status, cudaStream_t s) {
    const int grid = n < kNumSMs * 8 ? (n < 1 ? 1 : n) : kNumSMs * 8;
    AMV_LAUNCH(k_compact, grid, 256, 0, s, slots, slot_stride, size, off, n, out, out_cap, status);
}

void launch_export_meta(const uint64_t *off, const uint32_t *sz, const int32_t *st, uint64_t *hoff, uint32_t *hsz, int32_t *hst,
                        int n, cudaStream_t s) {
    AMV_LAUNCH(k_export_meta, (n + 255) / 256, 256, 0, s, off, sz, st, hoff, hsz, hst, n);
}

void launch_slot_offsets(uint64_t *off, int n, uint64_t stride, uint64_t base, cudaStream_t s) {
    AMV_LAUNCH(k_slot_offsets, (n + 255) / 256, 256, 0, s, off, n, stride, base);
}

}  // namespace amv

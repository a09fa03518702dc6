// amv_adpcm.cu -- IMA-ADPCM-AMV chunk decode / encode (sm_100a).
//
// The sample recurrence (predictor, step index) is strictly serial inside a chunk, and chunks
// are independent (decode: state is in the 8-byte header, adpcm.c:1270-1271; encode: given the
// step index carried in, adpcm.c:466).  So: one thread per chunk (or per chained stream), 32
// chunks per warp, and the warp moves data for its 32 chunks cooperatively through shared
// memory so that global traffic is coalesced although each thread walks its own chunk:
//
//   decode tile : 32 nibble bytes in  -> 64 samples (128 B) out per chunk
//   encode tile : 64 samples (128 B) in -> 32 nibble bytes out per chunk
//
// Input rows are fetched by the whole warp one chunk at a time (32 consecutive bytes / 64
// consecutive samples per instruction), transposed through a padded shared-memory tile
// (pitch odd in words => conflict-free column access), and results go back the same way.
#include "amv_common.cuh"
#include "amv_tables.cuh"
#include "amv_kernels.h"

namespace amv {

constexpr int kAdpcmWarps = 8;
constexpr int kAdpcmThreads = kAdpcmWarps * 32;
constexpr int kTileBytes = 32;                 // nibble bytes per chunk per tile
constexpr int kTileSamples = 2 * kTileBytes;   // 64
constexpr int kNibPitch = kTileBytes / 4 + 1;  // 9 words
constexpr int kPcmPitch = kTileSamples / 2 + 1;  // 33 words

__device__ uint16_t g_ima_step[96];

struct AdpcmSmem {
    uint16_t step[96];
    uint32_t nib[kAdpcmWarps][32 * kNibPitch];
    uint32_t pcm[kAdpcmWarps][32 * kPcmPitch];
};

__device__ __forceinline__ int ima_index_adjust(int q /* 0..7 */) { return (q & 4) ? ((q & 3) + 1) * 2 : -1; }

// adpcm_ima_expand_nibble(c, nibble, 3)  (adpcm.c:716-742)
__device__ __forceinline__ int ima_expand(int nib, int &pred, int &idx, const uint16_t *step_tab) {
    const int step = step_tab[idx];
    const int q = nib & 7;
    idx = min(max(idx + ima_index_adjust(q), 0), 88);
    const int diff = ((2 * q + 1) * step) >> 3;
    pred = (nib & 8) ? pred - diff : pred + diff;
    pred = min(max(pred, -32768), 32767);
    return pred;
}

// adpcm_ima_compress_sample (adpcm.c:219-227): q = min(7, |delta|*4/step) found by three
// compare/subtract steps against 4*step, 2*step, step (exact integer quotient, no division).
__device__ __forceinline__ int ima_compress(int sample, int &prev, int &idx, const uint16_t *step_tab) {
    const int step = step_tab[idx];
    const int delta = sample - prev;
    int t = (delta < 0 ? -delta : delta) * 4;
    int q = 0;
    if (t >= 4 * step) { q = 4; t -= 4 * step; }
    if (t >= 2 * step) { q |= 2; t -= 2 * step; }
    if (t >= step) q |= 1;
    const int mv = (step * (2 * q + 1)) >> 3;        // (step * difflookup) / 8, magnitude part
    prev = delta < 0 ? prev - mv : prev + mv;
    prev = min(max(prev, -32768), 32767);
    idx = min(max(idx + ima_index_adjust(q), 0), 88);
    return q | (delta < 0 ? 8 : 0);
}

// the same for the asynchronous kernel, with the quotient by reciprocal: the table holds (step, ceil(2^34 / step)) per index
// and q = min(7, umulhi(|delta|, reciprocal)).  Exact: the product overshoots 4*|delta|/step by less than |delta| / 2^32
// <= 2^-16, and a quotient that is not an integer lies at least 1/step >= 2^-15 below the next one (checked for all 89 steps
// x 65 536 differences by the CPU test suite).
__device__ __forceinline__ int ima_compress_s(int sample, int &prev, int &idx, uint32_t tab_s) {
    const uint2 sm = lds64(tab_s + 8u * (uint32_t)idx);
    const int step = (int)sm.x;
    const int delta = sample - prev;
    const int q = (int)min(__umulhi((uint32_t)(delta < 0 ? -delta : delta), sm.y), 7u);
    const int mv = (step * (2 * q + 1)) >> 3;        // (step * difflookup) / 8, magnitude part
    prev = delta < 0 ? prev - mv : prev + mv;
    prev = min(max(prev, -32768), 32767);
    idx = min(max(idx + ima_index_adjust(q), 0), 88);
    return q | (delta < 0 ? 8 : 0);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAdpcmThreads)
k_adpcm_decode(const uint8_t *__restrict__ chunks, uint64_t chunks_bytes, const uint64_t *__restrict__ off,
               const uint32_t *__restrict__ size, int n, int16_t *__restrict__ pcm, uint64_t pcm_samples,
               const uint64_t *__restrict__ pcm_off, int32_t *__restrict__ status) {
    __shared__ AdpcmSmem S;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) S.step[i] = g_ima_step[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t *nib = S.nib[wid], *out = S.pcm[wid];
    const int nwarps = (n + 31) >> 5;

    for (int wg = blockIdx.x * kAdpcmWarps + wid; wg < nwarps; wg += gridDim.x * kAdpcmWarps) {
        const int c = wg * 32 + lane;
        uint64_t src = 0, dsts = 0;
        uint32_t nbytes = 0;      // nibble bytes of this lane's chunk
        int pred = 0, idx = 0, st = 0;
        if (c < n) {
            const uint64_t o = off[c];
            const uint32_t sz = size[c];
            dsts = pcm_off[c];
            if (sz < 8) st = AMV_ST_SHORT;
            else if (!range_ok(o, sz, chunks_bytes) || !range_ok(dsts, 2ull * (sz - 8), pcm_samples)) st = AMV_ST_RANGE;
            else {
                const uint8_t *h = chunks + o;
                pred = (int)(int16_t)(h[0] | (h[1] << 8));
                idx = (int)(int16_t)(h[2] | (h[3] << 8));
                if (idx < 0 || idx > 88) st = AMV_ST_RANGE;   // the reference indexes step_table out of bounds here
                else { src = o + 8; nbytes = sz - 8; }
            }
            status[c] = st;
        }
        uint32_t maxb = nbytes;
#pragma unroll
        for (int d = 16; d; d >>= 1) maxb = max(maxb, __shfl_xor_sync(0xffffffffu, maxb, d));

        for (uint32_t t0 = 0; t0 < maxb; t0 += kTileBytes) {
            // warp-cooperative load: chunk j's 32 tile bytes with one byte per lane
            for (int j = 0; j < 32; j++) {
                const uint64_t sj = __shfl_sync(0xffffffffu, src, j);
                const uint32_t nj = __shfl_sync(0xffffffffu, nbytes, j);
                if (t0 >= nj) continue;
                uint8_t b = 0;
                if (t0 + lane < nj) b = chunks[sj + t0 + lane];
                reinterpret_cast<uint8_t *>(nib + j * kNibPitch)[lane] = b;
            }
            __syncwarp();
            if (t0 < nbytes) {
                const uint32_t *row = nib + lane * kNibPitch;
                uint32_t *orow = out + lane * kPcmPitch;
#pragma unroll
                for (int w = 0; w < kTileBytes / 4; w++) {
                    const uint32_t v = row[w];
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int byte = (v >> (8 * b)) & 0xff;
                        const int s0 = ima_expand(byte >> 4, pred, idx, S.step);      // high nibble first (:1281-1282)
                        const int s1 = ima_expand(byte & 15, pred, idx, S.step);
                        orow[w * 4 + b] = (uint32_t)(s0 & 0xffff) | ((uint32_t)s1 << 16);
                    }
                }
            }
            __syncwarp();
            // warp-cooperative store: chunk j's 64 samples, 32-bit per lane when aligned
            for (int j = 0; j < 32; j++) {
                const uint64_t dj = __shfl_sync(0xffffffffu, dsts, j);
                const uint32_t nj = __shfl_sync(0xffffffffu, nbytes, j);
                if (t0 >= nj) continue;
                const uint32_t v = out[j * kPcmPitch + lane];
                const uint64_t s0 = dj + 2ull * t0 + 2 * lane;           // sample index of this lane's pair
                if (t0 + lane < nj) {
                    if ((reinterpret_cast<uintptr_t>(pcm + s0) & 3) == 0) *reinterpret_cast<uint32_t *>(pcm + s0) = v;
                    else { pcm[s0] = (int16_t)(v & 0xffff); pcm[s0 + 1] = (int16_t)(v >> 16); }
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_adpcm_decode_async: the decoder with per-lane ASYNCHRONOUS staging of its input.  k_adpcm_decode above fetches a
// tile of 32 chunks x 32 nibble bytes with 32 warp-wide loads of one BYTE per lane (two shuffles each) -- as many
// instructions as the 768 the tile's arithmetic needs.  Here every lane asks for its own chunk's next 64 bytes -- the
// 16-byte aligned span around them, at most 80 bytes -- one tile ahead, decodes the previous tile out of its own
// shared-memory row meanwhile, and only the PCM goes out cooperatively (one chunk row per store instruction: 128
// contiguous bytes).  Every 16-byte unit the copy touches holds at least one byte of the chunk, so nothing outside the
// caller's buffer's last unit is read.  Two ways of asking, same kernel otherwise:
//   BULK   cp.async.bulk.shared.global on the warp's mbarrier (the TMA unit's 1-D copy, SASS UBLKCP).  Its operands are
//          warp-uniform, so 32 lanes with 32 addresses issue one after the other (an ELECT / R2UR loop, ~10 instructions
//          per lane and tile);
//   else   cp.async 16-byte copies (SASS LDGSTS): per-thread addressing, five instructions per lane and tile.
// ------------------------------------------------------------------------------------------------
constexpr int kBulkWarps = 4;
constexpr int kBulkTile = 64;                        // nibble bytes per chunk per tile
constexpr int kBulkRow = kBulkTile + 16;             // the aligned span of a tile
constexpr int kBulkOutPitch = kBulkTile + 1;         // words: 64 sample pairs + 1 (conflict-free columns)

struct AdpcmBulkWarp {
    __align__(16) uint8_t in[2][32][kBulkRow];
    uint32_t out[32 * kBulkOutPitch];
    uint4 row[32];              // per chunk of the warp: where its samples go (pointer lo, hi), nibble bytes, 32-bit aligned?
    __align__(8) uint64_t bar[2];
};
struct AdpcmBulkSmem {
    uint16_t step[96];
    __align__(16) AdpcmBulkWarp w[kBulkWarps];
};

template <bool BULK>
__global__ void __launch_bounds__(kBulkWarps * 32)
k_adpcm_decode_async(const uint8_t *__restrict__ chunks, uint64_t chunks_bytes, const uint64_t *__restrict__ off,
                     const uint32_t *__restrict__ size, int n, int16_t *__restrict__ pcm, uint64_t pcm_samples,
                     const uint64_t *__restrict__ pcm_off, int32_t *__restrict__ status) {
    AMV_EXTERN_SHARED(uint8_t, adpcm_smem_raw, 16);
    AdpcmBulkSmem &S = *reinterpret_cast<AdpcmBulkSmem *>(adpcm_smem_raw);
    for (int i = threadIdx.x; i < 96; i += blockDim.x) S.step[i] = g_ima_step[i];
    const uint32_t step_s = smem_addr(&S.step[0]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    AdpcmBulkWarp &W = S.w[wid];
    const uint32_t bar0_s = smem_addr(&W.bar[0]);                       // stage b: + 8 b
    const uint32_t row0_s = smem_addr(&W.in[0][lane][0]);               // stage b: + b * sizeof(W.in[0])
    if (BULK && lane == 0) { mbar_init(bar0_s, 32); mbar_init(bar0_s + 8, 32); }
    __syncthreads();
    const int nwarps = (n + 31) >> 5;
    uint32_t phases = 0;                                                // bit b: parity stage b's barrier completes next

    for (int wg = blockIdx.x * kBulkWarps + wid; wg < nwarps; wg += gridDim.x * kBulkWarps) {
        const int c = wg * 32 + lane;
        uint64_t src = 0, dsts = 0;
        uint32_t nbytes = 0;      // nibble bytes of this lane's chunk
        int pred = 0, idx = 0, st = 0;
        if (c < n) {
            const uint64_t o = off[c];
            const uint32_t sz = size[c];
            dsts = pcm_off[c];
            if (sz < 8) st = AMV_ST_SHORT;
            else if (!range_ok(o, sz, chunks_bytes) || !range_ok(dsts, 2ull * (sz - 8), pcm_samples)) st = AMV_ST_RANGE;
            else {
                const uint8_t *h = chunks + o;
                pred = (int)(int16_t)(h[0] | (h[1] << 8));
                idx = (int)(int16_t)(h[2] | (h[3] << 8));
                if (idx < 0 || idx > 88) st = AMV_ST_RANGE;   // the reference indexes step_table out of bounds here
                else { src = o + 8; nbytes = sz - 8; }
            }
            status[c] = st;
        }
        {
            const uint64_t p = reinterpret_cast<uint64_t>(pcm + dsts);
            W.row[lane] = make_uint4((uint32_t)p, (uint32_t)(p >> 32), nbytes, (p & 3) == 0 ? 1u : 0u);
        }
        uint32_t maxb = nbytes;
#pragma unroll
        for (int d = 16; d; d >>= 1) maxb = max(maxb, __shfl_xor_sync(0xffffffffu, maxb, d));
        const uint32_t ntiles = (maxb + kBulkTile - 1) / kBulkTile;
        const uint8_t *base = chunks + src;
        // the tile's aligned span goes to the lane's row of stage b; every lane takes part, with or without bytes
        auto issue = [&](uint32_t t, uint32_t b) {
            const uint32_t t0 = t * kBulkTile, row_s = row0_s + b * (uint32_t)sizeof(W.in[0]);
            uint32_t bytes = 0;
            uintptr_t a16 = 0;
            if (t0 < nbytes) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(base + t0), e = a + min((uint32_t)kBulkTile, nbytes - t0);
                a16 = a & ~uintptr_t(15);
                bytes = (uint32_t)(((e + 15) & ~uintptr_t(15)) - a16);
            }
            if (BULK) {
                if (bytes) {
                    fence_proxy_async();               // the row was read (two tiles ago) through the generic proxy
                    mbar_arrive_expect_tx(bar0_s + 8 * b, bytes);
                    bulk_g2s(row_s, reinterpret_cast<const void *>(a16), bytes, bar0_s + 8 * b);
                } else mbar_arrive_expect_tx(bar0_s + 8 * b, 0);
            } else {
                for (uint32_t k = 0; k < bytes; k += 16) cp_async16(row_s + k, reinterpret_cast<const void *>(a16 + k));
                cp_async_commit();
            }
        };
        __syncwarp();
        if (ntiles) issue(0, 0);
        for (uint32_t t = 0; t < ntiles; t++) {
            const uint32_t b = t & 1u;
            if (t + 1 < ntiles) issue(t + 1, b ^ 1u);
            if (BULK) { mbar_wait(bar0_s + 8 * b, (phases >> b) & 1u); phases ^= 1u << b; }
            else { if (t + 1 < ntiles) cp_async_wait<1>(); else cp_async_wait<0>(); }
            const uint32_t t0 = t * kBulkTile;
            if (t0 < nbytes) {
                const uint32_t len = min((uint32_t)kBulkTile, nbytes - t0);
                const uint32_t ra = row0_s + b * (uint32_t)sizeof(W.in[0]) + (uint32_t)(reinterpret_cast<uintptr_t>(base + t0) & 15);
                uint32_t *orow = W.out + lane * kBulkOutPitch;
                const uint32_t wa = ra & ~3u, sh = (ra & 3u) * 8u;
                uint32_t lo = lds32(wa);
                // one nibble (adpcm_ima_expand_nibble, adpcm.c:716-742).  A single look-up per (step index, nibble) was
                // measured too (a 5.7 KB table of difference | next index): 2.51 ms per 1 M chunks against 1.88 ms for this
                // arithmetic -- the table's dependent shared-memory load sits in the chain from nibble to nibble.
                auto expand = [&](int nib) -> uint32_t {
                    const int step = (int)lds_u16(step_s + 2u * (uint32_t)idx);
                    const int q = nib & 7;
                    idx = min(max(idx + ima_index_adjust(q), 0), 88);
                    const int diff = ((2 * q + 1) * step) >> 3;
                    pred = (nib & 8) ? pred - diff : pred + diff;
                    pred = min(max(pred, -32768), 32767);
                    return (uint32_t)pred;
                };
                for (uint32_t w = 0; w * 4 < len; w++) {
                    const uint32_t hi = lds32(wa + 4 * w + 4);
                    const uint32_t v = __funnelshift_r(lo, hi, sh);
                    lo = hi;
                    if (w * 4 + 4 <= len) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {                          // high nibble first (:1281-1282)
                            const uint32_t s0 = expand((int)(v >> (8 * k + 4)) & 15);
                            const uint32_t s1 = expand((int)(v >> (8 * k)) & 15);
                            orow[w * 4 + k] = __byte_perm(s0, s1, 0x5410);
                        }
                    } else {
                        for (uint32_t k = 0; w * 4 + k < len; k++) {
                            const uint32_t byte = (v >> (8 * k)) & 0xffu;
                            const uint32_t s0 = expand((int)(byte >> 4));
                            const uint32_t s1 = expand((int)(byte & 15u));
                            orow[w * 4 + k] = __byte_perm(s0, s1, 0x5410);
                        }
                    }
                }
            }
            __syncwarp();
            // warp-cooperative store: chunk j's sample pairs (one per nibble byte), two 32-bit words per lane, predicated
            {
                const uint32_t w0 = t0 + (uint32_t)lane;                       // the lane's first pair of the tile, in the chunk
                const uint32_t *orow = W.out + lane;
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const uint4 r = W.row[j];
                    const uint32_t v0 = orow[j * kBulkOutPitch], v1 = orow[j * kBulkOutPitch + 32];
                    const uint64_t p = ((uint64_t)r.y << 32) | r.x;
                    if (r.w) {
                        uint32_t *d = reinterpret_cast<uint32_t *>(p);
                        if (w0 < r.z) d[w0] = v0;
                        if (w0 + 32u < r.z) d[w0 + 32u] = v1;
                    } else {
                        int16_t *d = reinterpret_cast<int16_t *>(p);
                        if (w0 < r.z) { d[2 * w0] = (int16_t)(v0 & 0xffff); d[2 * w0 + 1] = (int16_t)(v0 >> 16); }
                        if (w0 + 32u < r.z) { d[2 * w0 + 64] = (int16_t)(v1 & 0xffff); d[2 * w0 + 65] = (int16_t)(v1 >> 16); }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One lane per stream; a stream is a run of chunks whose step index is chained.  With
// first_chunk == NULL every chunk is its own stream.
__global__ void __launch_bounds__(kAdpcmThreads)
k_adpcm_encode(const int16_t *__restrict__ pcm, uint64_t pcm_samples, const uint64_t *__restrict__ pcm_off,
               const uint32_t *__restrict__ nsamples, const uint32_t *__restrict__ first_chunk, int nstreams, int nchunks,
               const int16_t *__restrict__ step_in, int16_t *__restrict__ step_out, uint8_t *__restrict__ outb,
               uint64_t out_bytes, const uint64_t *__restrict__ out_off, int32_t *__restrict__ status) {
    __shared__ AdpcmSmem S;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) S.step[i] = g_ima_step[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t *nib = S.nib[wid], *in = S.pcm[wid];
    const int nwarps = (nstreams + 31) >> 5;

    for (int wg = blockIdx.x * kAdpcmWarps + wid; wg < nwarps; wg += gridDim.x * kAdpcmWarps) {
        const int s = wg * 32 + lane;
        uint32_t c0 = 0, c1 = 0;
        int idx = 0;
        bool dead = false;          // a bad chunk stops its stream
        if (s < nstreams) {
            c0 = first_chunk ? first_chunk[s] : (uint32_t)s;
            c1 = first_chunk ? first_chunk[s + 1] : (uint32_t)s + 1;
            // a table that is not monotonic or runs past the chunk arrays: the stream owns nothing (the host-memory
            // entry point rejects such a table outright)
            if (c1 > (uint32_t)nchunks) c1 = (uint32_t)nchunks;
            if (c0 > c1) c0 = c1;
            idx = step_in ? step_in[s] : 0;
            if (idx < 0 || idx > 88) { dead = true; for (uint32_t c = c0; c < c1; c++) status[c] = AMV_ST_RANGE; }
        }
        uint32_t maxc = c1 - c0;
#pragma unroll
        for (int d = 16; d; d >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, d));

        for (uint32_t k = 0; k < maxc; k++) {
            const uint32_t c = c0 + k;
            const bool have = !dead && c < c1;
            if (dead && c < c1 && s < nstreams) status[c] = AMV_ST_RANGE;     // rest of a broken stream
            uint64_t src = 0, dst = 0;
            uint32_t ns = 0;
            int prev = 0;
            if (have) {
                ns = nsamples[c]; src = pcm_off[c]; dst = out_off[c];
                int st = 0;
                if (ns & 1) st = AMV_ST_RANGE;
                else if (!range_ok(src, ns, pcm_samples) || !range_ok(dst, 8ull + ns / 2, out_bytes)) st = AMV_ST_RANGE;
                status[c] = st;
                if (st) { dead = true; ns = 0; }
                else {
                    // header: first sample, step index carried in, sample count (adpcm.c:464-479)
                    prev = ns ? pcm[src] : 0;
                    uint8_t *h = outb + dst;
                    h[0] = (uint8_t)prev; h[1] = (uint8_t)(prev >> 8);
                    h[2] = (uint8_t)idx;  h[3] = (uint8_t)(idx >> 8);
                    h[4] = (uint8_t)ns; h[5] = (uint8_t)(ns >> 8); h[6] = (uint8_t)(ns >> 16); h[7] = (uint8_t)(ns >> 24);
                }
            }
            uint32_t maxs = ns;
#pragma unroll
            for (int d = 16; d; d >>= 1) maxs = max(maxs, __shfl_xor_sync(0xffffffffu, maxs, d));

            for (uint32_t t0 = 0; t0 < maxs; t0 += kTileSamples) {
                for (int j = 0; j < 32; j++) {
                    const uint64_t sj = __shfl_sync(0xffffffffu, src, j);
                    const uint32_t nj = __shfl_sync(0xffffffffu, ns, j);
                    if (t0 >= nj) continue;
                    // 64 samples of chunk j: lanes take samples lane and lane+32
                    int16_t a = 0, b = 0;
                    if (t0 + lane < nj) a = pcm[sj + t0 + lane];
                    if (t0 + 32 + lane < nj) b = pcm[sj + t0 + 32 + lane];
                    int16_t *row = reinterpret_cast<int16_t *>(in + j * kPcmPitch);
                    row[lane] = a; row[32 + lane] = b;
                }
                __syncwarp();
                if (t0 < ns) {
                    const uint32_t *row = in + lane * kPcmPitch;
                    uint32_t *orow = nib + lane * kNibPitch;
#pragma unroll
                    for (int w = 0; w < kTileBytes / 4; w++) {
                        uint32_t packed = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const uint32_t v = row[w * 4 + b];
                            // sample pairs past the end of the chunk are not encoded (keeps the carried state exact)
                            if (t0 + 2 * (w * 4 + b) < ns) {
                                const int n0 = ima_compress((int)(int16_t)(v & 0xffff), prev, idx, S.step);
                                const int n1 = ima_compress((int)(int16_t)(v >> 16), prev, idx, S.step);
                                packed |= (uint32_t)((n0 << 4) | n1) << (8 * b);
                            }
                        }
                        orow[w] = packed;
                    }
                }
                __syncwarp();
                for (int j = 0; j < 32; j++) {
                    const uint64_t dj = __shfl_sync(0xffffffffu, dst, j);
                    const uint32_t nj = __shfl_sync(0xffffffffu, ns, j);
                    if (t0 >= nj) continue;
                    const uint32_t bytes_left = (nj - t0 + 1) / 2;
                    if ((uint32_t)lane < bytes_left)
                        outb[dj + 8 + t0 / 2 + lane] = reinterpret_cast<const uint8_t *>(nib + j * kNibPitch)[lane];
                }
                __syncwarp();
            }
        }
        if (s < nstreams && step_out) step_out[s] = (int16_t)idx;
    }
}

// ------------------------------------------------------------------------------------------------
// k_adpcm_encode_async: the encoder with the same per-lane asynchronous input staging as k_adpcm_decode_async (cp.async,
// 16-byte copies: the aligned span around the lane's next 64 samples, at most 144 bytes, one tile ahead) instead of 32
// warp-wide loads of two samples per lane; nibble bytes still leave cooperatively, one chunk row per store instruction,
// with the chunks' output offsets and sizes read from shared memory instead of shuffled.
// ------------------------------------------------------------------------------------------------
constexpr int kEncRow = 2 * kTileSamples + 16;       // bytes: the aligned span of 64 samples
struct AdpcmEncWarp {
    __align__(16) uint8_t in[2][32][kEncRow];
    uint32_t nib[32 * kNibPitch];
    uint4 row[32];                                  // per chunk of the warp: where its nibble bytes go (pointer lo, hi), how many
};
struct AdpcmEncSmem {
    uint2 step[96];                                 // (step, ceil(2^34 / step)) per step index
    __align__(16) AdpcmEncWarp w[kBulkWarps];
};

__global__ void __launch_bounds__(kBulkWarps * 32)
k_adpcm_encode_async(const int16_t *__restrict__ pcm, uint64_t pcm_samples, const uint64_t *__restrict__ pcm_off,
                     const uint32_t *__restrict__ nsamples, const uint32_t *__restrict__ first_chunk, int nstreams, int nchunks,
                     const int16_t *__restrict__ step_in, int16_t *__restrict__ step_out, uint8_t *__restrict__ outb,
                     uint64_t out_bytes, const uint64_t *__restrict__ out_off, int32_t *__restrict__ status) {
    AMV_EXTERN_SHARED(uint8_t, adpcm_enc_smem_raw, 16);
    AdpcmEncSmem &S = *reinterpret_cast<AdpcmEncSmem *>(adpcm_enc_smem_raw);
    for (int i = threadIdx.x; i < 96; i += blockDim.x) {
        const uint32_t st = g_ima_step[i];
        S.step[i] = make_uint2(st, st ? (uint32_t)(((1ull << 34) + st - 1u) / st) : 0u);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    AdpcmEncWarp &W = S.w[wid];
    const uint32_t row0_s = smem_addr(&W.in[0][lane][0]);               // stage b: + b * sizeof(W.in[0])
    const uint32_t step_s = smem_addr(&S.step[0]);
    const int nwarps = (nstreams + 31) >> 5;

    for (int wg = blockIdx.x * kBulkWarps + wid; wg < nwarps; wg += gridDim.x * kBulkWarps) {
        const int s = wg * 32 + lane;
        uint32_t c0 = 0, c1 = 0;
        int idx = 0;
        bool dead = false;          // a bad chunk stops its stream
        if (s < nstreams) {
            c0 = first_chunk ? first_chunk[s] : (uint32_t)s;
            c1 = first_chunk ? first_chunk[s + 1] : (uint32_t)s + 1;
            if (c1 > (uint32_t)nchunks) c1 = (uint32_t)nchunks;
            if (c0 > c1) c0 = c1;
            idx = step_in ? step_in[s] : 0;
            if (idx < 0 || idx > 88) { dead = true; for (uint32_t c = c0; c < c1; c++) status[c] = AMV_ST_RANGE; }
        }
        uint32_t maxc = c1 - c0;
#pragma unroll
        for (int d = 16; d; d >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, d));

        for (uint32_t k = 0; k < maxc; k++) {
            const uint32_t c = c0 + k;
            const bool have = !dead && c < c1;
            if (dead && c < c1 && s < nstreams) status[c] = AMV_ST_RANGE;     // rest of a broken stream
            uint64_t src = 0, dst = 0;
            uint32_t ns = 0;
            int prev = 0;
            if (have) {
                ns = nsamples[c]; src = pcm_off[c]; dst = out_off[c];
                int st = 0;
                if (ns & 1) st = AMV_ST_RANGE;
                else if (!range_ok(src, ns, pcm_samples) || !range_ok(dst, 8ull + ns / 2, out_bytes)) st = AMV_ST_RANGE;
                status[c] = st;
                if (st) { dead = true; ns = 0; }
                else {
                    // header: first sample, step index carried in, sample count (adpcm.c:464-479)
                    prev = ns ? pcm[src] : 0;
                    uint8_t *h = outb + dst;
                    h[0] = (uint8_t)prev; h[1] = (uint8_t)(prev >> 8);
                    h[2] = (uint8_t)idx;  h[3] = (uint8_t)(idx >> 8);
                    h[4] = (uint8_t)ns; h[5] = (uint8_t)(ns >> 8); h[6] = (uint8_t)(ns >> 16); h[7] = (uint8_t)(ns >> 24);
                }
            }
            {
                const uint64_t p = reinterpret_cast<uint64_t>(outb + dst + 8);
                W.row[lane] = make_uint4((uint32_t)p, (uint32_t)(p >> 32), ns / 2, 0u);
            }
            uint32_t maxs = ns;
#pragma unroll
            for (int d = 16; d; d >>= 1) maxs = max(maxs, __shfl_xor_sync(0xffffffffu, maxs, d));
            const uint32_t ntiles = (maxs + kTileSamples - 1) / kTileSamples;
            const int16_t *base = pcm + src;
            auto issue = [&](uint32_t t, uint32_t b) {
                const uint32_t t0 = t * kTileSamples, row_s = row0_s + b * (uint32_t)sizeof(W.in[0]);
                if (t0 < ns) {
                    const uintptr_t a = reinterpret_cast<uintptr_t>(base + t0), e = a + 2u * min((uint32_t)kTileSamples, ns - t0);
                    const uintptr_t a16 = a & ~uintptr_t(15);
                    const uint32_t bytes = (uint32_t)(((e + 15) & ~uintptr_t(15)) - a16);
                    for (uint32_t q = 0; q < bytes; q += 16) cp_async16(row_s + q, reinterpret_cast<const void *>(a16 + q));
                }
                cp_async_commit();
            };
            __syncwarp();
            if (ntiles) issue(0, 0);
            for (uint32_t t = 0; t < ntiles; t++) {
                const uint32_t b = t & 1u, t0 = t * kTileSamples;
                if (t + 1 < ntiles) { issue(t + 1, b ^ 1u); cp_async_wait<1>(); } else cp_async_wait<0>();
                if (t0 < ns) {
                    // the lane's samples sit in its row from the byte the source address has inside its 16-byte unit (even)
                    const uint32_t ra = row0_s + b * (uint32_t)sizeof(W.in[0]) + (uint32_t)(reinterpret_cast<uintptr_t>(base + t0) & 15);
                    const uint32_t wa = ra & ~3u, sh = (ra & 3u) * 8u;
                    uint32_t *orow = W.nib + lane * kNibPitch;
                    uint32_t lo = lds32(wa);
                    // eight samples per trip, not the whole tile unrolled: 3 300 SASS instructions walked by 20 warps at
                    // different phases miss the instruction cache (ncu: no_instruction 1.0 per issue)
                    if (t0 + kTileSamples <= ns) {          // a whole tile: no end-of-chunk test per pair
#pragma unroll 1
                        for (int w = 0; w < kTileBytes / 4; w++) {
                            uint32_t packed = 0;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const uint32_t hi = lds32(wa + 4u * (w * 4 + q) + 4u);
                                const uint32_t v = __funnelshift_r(lo, hi, sh);
                                lo = hi;
                                const int n0 = ima_compress_s((int)(int16_t)(v & 0xffff), prev, idx, step_s);
                                const int n1 = ima_compress_s((int)(int16_t)(v >> 16), prev, idx, step_s);
                                packed |= (uint32_t)((n0 << 4) | n1) << (8 * q);
                            }
                            orow[w] = packed;
                        }
                    } else {
#pragma unroll 1
                        for (int w = 0; w < kTileBytes / 4; w++) {
                            uint32_t packed = 0;
#pragma unroll 1
                            for (int q = 0; q < 4; q++) {
                                const uint32_t hi = lds32(wa + 4u * (w * 4 + q) + 4u);
                                const uint32_t v = __funnelshift_r(lo, hi, sh);
                                lo = hi;
                                // sample pairs past the end of the chunk are not encoded (keeps the carried state exact)
                                if (t0 + 2 * (w * 4 + q) < ns) {
                                    const int n0 = ima_compress_s((int)(int16_t)(v & 0xffff), prev, idx, step_s);
                                    const int n1 = ima_compress_s((int)(int16_t)(v >> 16), prev, idx, step_s);
                                    packed |= (uint32_t)((n0 << 4) | n1) << (8 * q);
                                }
                            }
                            orow[w] = packed;
                        }
                    }
                }
                __syncwarp();
                {   // the 32 nibble bytes of every chunk of the warp: lane i stores byte i, one chunk per trip, predicated
                    const uint32_t boff = t0 / 2 + (uint32_t)lane;
                    const uint8_t *nb = reinterpret_cast<const uint8_t *>(W.nib) + lane;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const uint4 r = W.row[j];
                        if (boff < r.z) reinterpret_cast<uint8_t *>(((uint64_t)r.y << 32) | r.x)[boff] = nb[j * kNibPitch * 4];
                    }
                }
                __syncwarp();
            }
        }
        if (s < nstreams && step_out) step_out[s] = (int16_t)idx;
    }
}

// ------------------------------------------------------------------------------------------------
// The `-trellis N` path of the reference encoder (adpcm_compress_trellis, adpcm.c:287-443, IMA branch
// :383-395 with STORE_NODE :340-377): a beam search over decoder states with a frontier of F = 2^N nodes
// kept sorted by accumulated squared error; candidates that decode to a sample already in the next
// frontier are dropped, a full frontier recycles its worst node together with its path slot, and the
// best path is frozen every 128 samples.  The search of one chunk is a chain of data-dependent
// insertions, so the unit of parallelism stays the chunk (one thread per stream, chunks of a stream
// chained through the step index); node pools, the sorted index lists and the F x 128 path table live
// in per-thread local memory.  Samples are read and nibble bytes written straight from / to global
// memory: the search, not the 3 KB of I/O per chunk, is what this kernel spends its time on.
template <int F>
__global__ void __launch_bounds__(128)
k_adpcm_encode_trellis(const int16_t *__restrict__ pcm, uint64_t pcm_samples, const uint64_t *__restrict__ pcm_off,
                       const uint32_t *__restrict__ nsamples, const uint32_t *__restrict__ first_chunk, int nstreams, int nchunks,
                       const int16_t *__restrict__ step_in, int16_t *__restrict__ step_out, uint8_t *__restrict__ outb,
                       uint64_t out_bytes, const uint64_t *__restrict__ out_off, int32_t *__restrict__ status) {
    constexpr int kFreeze = 128;
    __shared__ uint16_t step_tab[96];
    for (int i = threadIdx.x; i < 96; i += blockDim.x) step_tab[i] = g_ima_step[i];
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    uint32_t c0 = first_chunk ? first_chunk[s] : (uint32_t)s, c1 = first_chunk ? first_chunk[s + 1] : (uint32_t)s + 1;
    if (c1 > (uint32_t)nchunks) c1 = (uint32_t)nchunks;
    if (c0 > c1) c0 = c1;
    int idx = step_in ? step_in[s] : 0;
    bool dead = idx < 0 || idx > 88;

    struct Node { uint32_t ssd; int sample1; uint16_t path; uint8_t step; };
    Node pool[2][F];
    uint8_t cur[F], nxt[F];
    uint16_t paths[F * kFreeze];            // [15:12] nibble, [11:0] previous path slot
    uint8_t nibbuf[kFreeze];

    for (uint32_t c = c0; c < c1; c++) {
        if (dead) { status[c] = AMV_ST_RANGE; continue; }
        const uint32_t ns = nsamples[c];
        const uint64_t src = pcm_off[c], dst = out_off[c];
        if ((ns & 1) || !range_ok(src, ns, pcm_samples) || !range_ok(dst, 8ull + ns / 2, out_bytes)) { status[c] = AMV_ST_RANGE; dead = true; continue; }
        status[c] = 0;
        const int prev0 = ns ? pcm[src] : 0;
        uint8_t *h = outb + dst;
        h[0] = (uint8_t)prev0; h[1] = (uint8_t)(prev0 >> 8);
        h[2] = (uint8_t)idx;   h[3] = (uint8_t)(idx >> 8);
        h[4] = (uint8_t)ns; h[5] = (uint8_t)(ns >> 8); h[6] = (uint8_t)(ns >> 16); h[7] = (uint8_t)(ns >> 24);
        if (!ns) continue;
        uint8_t *body = h + 8;

        int ncur = 1, pathn = 0, froze = -1;
        pool[1][0].ssd = 0; pool[1][0].path = 0; pool[1][0].step = (uint8_t)idx; pool[1][0].sample1 = prev0;
        cur[0] = 0;
        for (int i = 0; i < (int)ns; i++) {
            Node *from = pool[(i & 1) ^ 1], *to = pool[i & 1];
            const int sample = pcm[src + i];
            int nnxt = 0, nalloc = 0;
            for (int j = 0; j < ncur; j++) {
                const Node p = from[cur[j]];
                const int range = j < F / 2 ? 1 : 0;
                const int st = step_tab[p.step];
                const int div = (sample - p.sample1) * 4 / st;
                int nmin = min(max(div - range, -7), 6), nmax = min(max(div + range, -6), 7);
                if (nmin <= 0) nmin--;                                   // distinguish -0 from +0
                if (nmax < 0) nmax--;
                for (int nidx = nmin; nidx <= nmax; nidx++) {
                    const int nibble = nidx < 0 ? 7 - nidx : nidx;
                    const int mag = (st * (2 * (nibble & 7) + 1)) >> 3;  // (step * yamaha_difflookup[nibble]) / 8, magnitude part
                    const int dec = min(max((nibble & 8) ? p.sample1 - mag : p.sample1 + mag, -32768), 32767);
                    const int d = sample - dec;
                    const uint32_t ssd = p.ssd + (uint32_t)d * (uint32_t)d;
                    if (nnxt == F && ssd >= to[nxt[F - 1]].ssd) continue;
                    bool dup = false;
                    for (int k = 0; k < nnxt; k++) dup |= dec == to[nxt[k]].sample1;
                    if (dup) continue;
                    int k = 0;
                    while (k < nnxt && ssd >= to[nxt[k]].ssd) k++;       // first slot whose node is worse
                    int slot;
                    if (nnxt == F) slot = nxt[F - 1];                    // recycle the evicted node and its path slot
                    else { slot = nalloc++; to[slot].path = (uint16_t)pathn++; }
                    to[slot].ssd = ssd;
                    to[slot].step = (uint8_t)min(max((int)p.step + ima_index_adjust(nibble & 7), 0), 88);
                    to[slot].sample1 = dec;
                    paths[to[slot].path] = (uint16_t)((nibble << 12) | p.path);
                    const int last = nnxt == F ? F - 1 : nnxt;
                    for (int m = last; m > k; m--) nxt[m] = nxt[m - 1];
                    nxt[k] = (uint8_t)slot;
                    if (nnxt < F) nnxt++;
                }
            }
            for (int k = 0; k < nnxt; k++) cur[k] = nxt[k];
            ncur = nnxt;
            if (to[cur[0]].ssd > (1u << 28)) {                           // prevent overflow
                const uint32_t base = to[cur[0]].ssd;
                for (int j = 1; j < ncur; j++) to[cur[j]].ssd -= base;
                to[cur[0]].ssd = 0;
            }
            if (i == froze + kFreeze) {                                  // merge old paths to save memory
                int pp = to[cur[0]].path;
                for (int k = kFreeze - 1; k >= 0; k--) { nibbuf[k] = (uint8_t)(paths[pp] >> 12); pp = paths[pp] & 0xfff; }
                uint8_t *o = body + ((froze + 1) >> 1);                  // froze + 1 is a multiple of 128
                for (int k = 0; k < kFreeze / 2; k++) o[k] = (uint8_t)((nibbuf[2 * k] << 4) | nibbuf[2 * k + 1]);
                froze = i; pathn = 0;
                ncur = 1;                                                // "just kill them all"
            }
        }
        const Node best = pool[(ns - 1) & 1][cur[0]];
        const int rest = (int)ns - 1 - froze;                            // even: ns is even, froze + 1 a multiple of 128
        int pp = best.path;
        for (int k = rest - 1; k >= 0; k--) { nibbuf[k] = (uint8_t)(paths[pp] >> 12); pp = paths[pp] & 0xfff; }
        uint8_t *o = body + ((froze + 1) >> 1);
        for (int k = 0; k < rest / 2; k++) o[k] = (uint8_t)((nibbuf[2 * k] << 4) | nibbuf[2 * k + 1]);
        idx = best.step;
    }
    if (step_out) step_out[s] = (int16_t)idx;
}

// ------------------------------------------------------------------------------------------------
cudaError_t upload_adpcm_tables(cudaStream_t s) {
    static uint16_t h[96];
    for (int i = 0; i < 96; i++) h[i] = kImaStep[i < 89 ? i : 88];
    return cudaMemcpyToSymbolAsync(g_ima_step, h, sizeof(h), 0, cudaMemcpyHostToDevice, s);
}

static int adpcm_grid(int units) {
    const int warps = (units + 31) / 32;
    const int ctas = (warps + kAdpcmWarps - 1) / kAdpcmWarps;
    const int cap = kNumSMs * 5;
    return ctas < 1 ? 1 : (ctas < cap ? ctas : cap);
}

cudaError_t adpcm_setup_device() {
    cudaError_t e = cudaFuncSetAttribute(k_adpcm_decode_async<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdpcmBulkSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_adpcm_decode_async<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdpcmBulkSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_adpcm_encode_async, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AdpcmEncSmem));
    return e;
}

void launch_adpcm_decode(const uint8_t *chunks, uint64_t chunks_bytes, const uint64_t *off, const uint32_t *size, int n,
                         int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off, int32_t *status, int form, cudaStream_t s) {
    if (form) {             // asynchronous input staging: 2 = cp.async.bulk + mbarrier (UBLKCP), 1 = cp.async (LDGSTS)
        const int warps = (n + 31) / 32, ctas = (warps + kBulkWarps - 1) / kBulkWarps, cap = kNumSMs * 4;
        const int grid = ctas < 1 ? 1 : (ctas < cap ? ctas : cap);
        if (form == 2) AMV_LAUNCH(k_adpcm_decode_async<true>, grid, kBulkWarps * 32, sizeof(AdpcmBulkSmem), s, chunks, chunks_bytes, off, size, n,
                                  pcm, pcm_samples, pcm_off, status);
        else           AMV_LAUNCH(k_adpcm_decode_async<false>, grid, kBulkWarps * 32, sizeof(AdpcmBulkSmem), s, chunks, chunks_bytes, off, size, n,
                                  pcm, pcm_samples, pcm_off, status);
        return;
    }
    AMV_LAUNCH(k_adpcm_decode, adpcm_grid(n), kAdpcmThreads, 0, s, chunks, chunks_bytes, off, size, n, pcm, pcm_samples, pcm_off,
                                                           status);
}

void launch_adpcm_encode(const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off, const uint32_t *nsamples,
                         const uint32_t *first_chunk, int nstreams, int nchunks, const int16_t *step_in, int16_t *step_out,
                         uint8_t *out, uint64_t out_bytes, const uint64_t *out_off, int32_t *status, int trellis, int form, cudaStream_t s) {
    if (trellis > 0) {
        const int grid = (nstreams + 127) / 128;
#define AMV_TRELLIS_CASE(T)                                                                                              \
        case T: AMV_LAUNCH((k_adpcm_encode_trellis<(1 << T)>), grid, 128, 0, s, pcm, pcm_samples, pcm_off, nsamples, first_chunk,   \
                           nstreams, nchunks, step_in, step_out, out, out_bytes, out_off, status); break;
        switch (trellis) {
            AMV_TRELLIS_CASE(1) AMV_TRELLIS_CASE(2) AMV_TRELLIS_CASE(3) AMV_TRELLIS_CASE(4) AMV_TRELLIS_CASE(5)
            default: break;
        }
#undef AMV_TRELLIS_CASE
        return;
    }
    if (form) {             // asynchronous input staging (cp.async)
        const int warps = (nstreams + 31) / 32, ctas = (warps + kBulkWarps - 1) / kBulkWarps, cap = kNumSMs * 5;
        AMV_LAUNCH(k_adpcm_encode_async, ctas < 1 ? 1 : (ctas < cap ? ctas : cap), kBulkWarps * 32, sizeof(AdpcmEncSmem), s, pcm, pcm_samples,
                   pcm_off, nsamples, first_chunk, nstreams, nchunks, step_in, step_out, out, out_bytes, out_off, status);
        return;
    }
    AMV_LAUNCH(k_adpcm_encode, adpcm_grid(nstreams), kAdpcmThreads, 0, s, pcm, pcm_samples, pcm_off, nsamples, first_chunk, nstreams,
               nchunks, step_in, step_out, out, out_bytes, out_off, status);
}

}  // namespace amv

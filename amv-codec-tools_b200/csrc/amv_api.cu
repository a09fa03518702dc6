// amv_api.cu -- the C ABI of libamvcuda (include/amvcuda.h): context, workspaces, host<->device
// staging, and the kernel sequences behind each entry point.  No arithmetic of the codec lives
// here and there is no CPU path: without a usable sm_100 device amv_create fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/amvcuda.h"
#include "amv_common.cuh"
#include "amv_kernels.h"

using namespace amv;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

enum WsId {
    WS_SLOT_OFF, WS_SCAN_LEN, WS_STATUS, WS_STARTS, WS_SCRATCH, WS_SLOTS, WS_CARRY, WS_ROUNDS, WS_TOKENS, WS_BLKOFF, WS_QTAB, WS_REDO,
    // device mirrors of host arguments (AMV_MEM_HOST calls)
    WS_H_A, WS_H_B, WS_H_C, WS_H_D, WS_H_E, WS_H_F, WS_H_G, WS_H_H, WS_H_I,
    WS_RS_BANK,     // the audio resampler's polyphase bank
    WS_COUNT
};

}  // namespace

// optional per-kernel device timing (option "profile_events"): CUDA events recorded on the
// context's stream right around the launch of each hot kernel
enum KernelKind { KK_ENCODE, KK_IDCT, KK_UNSTUFF, KK_SYNC, KK_ADPCM_DEC, KK_ADPCM_ENC, KK_COMPACT, KK_TOKENS, KK_IDCT_BGR, KK_COUNT };
static const char *const kKernelKindName[KK_COUNT] = { "encode", "idct", "unstuff", "sync", "adpcm_dec", "adpcm_enc", "compact", "tokens", "idct_bgr" };
struct EvPair { cudaEvent_t a, b; int kind; };

struct amv_ctx {
    bool opt_profile = false;
    std::vector<EvPair> evs;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    int opt_log2p = -1;                 // decode lanes per frame, -1 = choose from the batch size
    uint64_t opt_slot_ws_bytes = 4ull << 30;
    uint32_t last_rounds = 0;
    std::string err;
    DevBuf ws[WS_COUNT];
    // host-memory calls: copy-in / copy-out streams next to the compute stream, and a pinned
    // bounce buffer for the per-frame metadata (the caller's arrays may be pageable)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    void *pinned_meta = nullptr;
    size_t pinned_meta_cap = 0;
    void *pinned_small = nullptr;       // staging of small host calls (the AVCodec callbacks: one frame per call)
    size_t pinned_small_cap = 0;
    int opt_host_chunk = 0;             // frames per pipeline stage, 0 = choose
    int opt_resample_form = 2;          // audio resampler: 2 = phase rows (M outputs per coefficient row), 1 = tiles, 0 = direct form
    int opt_scale_form = 1;             // scaler: 2 = staged tiles (source rows staged in shared memory), 1 = tiles, 0 = direct form
    int opt_encode_rounds = 4;          // encoder: 4 = k_encode16v2 with the regrouped transform, 2 = with the factorised one, 1 = k_encode16 (each + k_encode for the frames it hands back), 0 = k_encode alone
    int opt_trellis = 0;                // ADPCM encoder: 0 = adpcm_ima_compress_sample, 1..5 = -trellis N beam search
    // decode (AMV / SP5X, fixed tables): 2 = lean pass with 16-bit tokens + k_idct16, 1 = lean pass with 32-bit tokens + k_idct,
    // 0 = k_vlc_tokens (flat loop, 32-bit tokens) + k_idct.  Measured per 100 000 frames 320x240 (profiles/r5b_*): tokens + idct
    // 9.16 + 9.31 / 12.38 + 8.01 / 12.33 + 8.01 ms -- the token pass is ALU-pipe bound and the 32-bit token's offset and
    // product cost it more than they save the consumer.
    int opt_token_pass = 2;
    int opt_adpcm_form = 1;             // ADPCM decode input staging: 1 = cp.async (LDGSTS), 2 = cp.async.bulk + mbarrier (UBLKCP), 0 = cooperative loads
    bool opt_small_calls = true;        // host calls of a few frames: one pinned block, one stream, one synchronisation
    bool opt_zero_copy_packets = true;  // decode: kernels read pinned packets in place (else: DMA into a device copy)
    // plain JPEG (amv_mjpeg_configure): the table set and the header bytes every frame must start with
    void *mj_tables = nullptr;          // DecTableSet in device memory
    uint8_t *mj_hdr = nullptr;          // device copy of the sample's bytes up to the end of the SOS header
    uint32_t mj_hdr_len = 0;
    int mj_restart = 0;                 // restart interval of the configured header (0: none, or one the reference ignores)
    int mj_samp[4] = { 1, 1, 0, 0 };    // log2 sampling of the configured header: luma h, v; chroma h, v
    uint32_t mj_qpos[2] = { 0, 0 };     // where the 64 quantisers of component 0 / components 1, 2 sit in a frame
    int mj_w = 0, mj_h = 0;
    int rs_in_rate = 0, rs_out_rate = 0; // rate pair WS_RS_BANK was built for
    bool mj_sync_ok = false;            // the lane-synchronisation table could be built (else one lane per frame)
};

// how a packet is framed and which tables its scan uses
struct DecMode {
    uint32_t head = 2;                  // bytes in front of the scan
    bool literal = false;               // SP5X: scan bytes are literal and run to the end of the packet
    bool flip = true;                   // AMV stores the picture bottom-up
    const amv::DecTableSet *tables = nullptr;   // nullptr: the fixed AMV / SP5X (or amvlib) set
    const uint8_t *hdr = nullptr;       // plain JPEG: required header bytes (device)
    uint32_t hdr_len = 0;
    uint32_t qpos[2] = { 0, 0 };        // plain JPEG: offsets of the per-frame quantisers
    int samp[4] = { 1, 1, 0, 0 };       // log2 sampling: luma h, v; chroma h, v (AMV / SP5X: 4:2:0)
    int restart = 0;                    // plain JPEG: MCUs per restart interval
    Geom geom(int w, int h) const { Geom g = make_geom_sampled(w, h, samp[0], samp[1], samp[2], samp[3]); if (!flip) g.flip = 0; return g; }
    bool allow_sync = true;
};
static const DecMode kModeAmv;
static DecMode mode_sp5x() { DecMode m; m.head = 14; m.literal = true; m.flip = false; return m; }

namespace {

int fail(amv_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

// Entry points run on the context's device and put the caller's current device back when they return.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        // always set: on a host thread that has made no CUDA call yet this is what binds the device's context to the
        // thread (without it cudaPointerGetAttributes reports no device pointer for pinned memory)
        err = cudaSetDevice(dev);
        changed = err == cudaSuccess && prev != dev;
    }
    ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(ctx)                                                            \
    DeviceGuard dev_guard_((ctx)->device);                                        \
    if (dev_guard_.err != cudaSuccess) return fail((ctx), AMV_ERR_CUDA, "cudaSetDevice", dev_guard_.err)

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return fail(ctx, AMV_ERR_CUDA, #call, e_);         \
    } while (0)

int ensure(amv_ctx *ctx, WsId id, size_t bytes, void **out) {
    DevBuf &b = ctx->ws[id];
    if (bytes > b.cap) {
        if (b.p) {
            CK(cudaStreamSynchronize(ctx->stream));     // nothing in flight may still use the old block
            CK(cudaFree(b.p));
            b.p = nullptr; b.cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) { b.p = nullptr; return fail(ctx, AMV_ERR_NOMEM, "cudaMalloc workspace", e); }
        b.cap = want;
    }
    *out = b.p;
    return AMV_OK;
}

#define ENSURE(id, bytes, ptr)                                                    \
    do {                                                                          \
        void *p_ = nullptr;                                                       \
        int r_ = ensure(ctx, id, (bytes), &p_);                                   \
        if (r_ != AMV_OK) return r_;                                              \
        ptr = reinterpret_cast<decltype(ptr)>(p_);                                \
    } while (0)

int check_launch(amv_ctx *ctx, const char *what, int count = 1) {
    ctx->launches += count;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(ctx, AMV_ERR_CUDA, what, e);
    return AMV_OK;
}

constexpr size_t kMaxProfileEvents = 4096;
struct ScopedTimer {
    amv_ctx *c; EvPair e; bool on;
    ScopedTimer(amv_ctx *ctx, int kind) : c(ctx), on(ctx->opt_profile) {
        if (!on) return;
        e.kind = kind;
        if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) { on = false; return; }
        cudaEventRecord(e.a, c->stream);
    }
    ~ScopedTimer() {
        if (!on) return;
        cudaEventRecord(e.b, c->stream);
        // nobody polls: keep the newest kMaxProfileEvents samples instead of growing without bound
        if (c->evs.size() >= kMaxProfileEvents) {
            cudaEventDestroy(c->evs.front().a); cudaEventDestroy(c->evs.front().b);
            c->evs.erase(c->evs.begin());
        }
        c->evs.push_back(e);
    }
};

int pick_log2p(const amv_ctx *ctx, int n) {
    if (ctx->opt_log2p >= 0) return ctx->opt_log2p > 5 ? 5 : ctx->opt_log2p;
    // Splitting a frame into P lanes costs the synchronisation pass: two walks of every subsequence for a token pass
    // that gets P times shorter, so P = 2 never pays and P >= 4 pays only while one lane per frame leaves the GPU
    // short of warps.  Measured on 320x240 (decode ms, 1 / 4 / 8 / 16 lanes): 4096 frames 5.5 / 3.9 / 2.5 / 2.3,
    // 8192: 6.0 / 4.3 / 3.3 / 3.8, 16384: 6.8 / 6.2 / 6.8 / 6.4, 25000: 7.7 / 9.3 / 8.9 / 9.3 -- i.e. aim at ~64 k lanes
    // below ~20 k frames, one lane per frame above.
    if (n >= 20000) return 0;
    int l = 2;
    while (l < 5 && ((int64_t)n << l) < 65536) l++;
    return l;
}

// ---------------------------------------------------------------------------------- device paths
// Front half shared by both decoder flavours: slot offsets, un-stuffing, (lane synchronisation,)
// Huffman -> tokens.  payload_bytes: upper bound of the bytes of the n packets; sizes the scratch.
struct DecodeFront {
    uint64_t *slot_off; uint32_t *scan_len; uint32_t *tokens; uint32_t *blk_off; int32_t *st; int launches;
    bool tok16; const DecTableSet *tabs;
};

int decode_front(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size,
                 int n, const Geom &g, int32_t *status, uint64_t payload_bytes, bool amvlib, DecodeFront &F,
                 const DecMode &mode = kModeAmv) {
    const int log2p = mode.allow_sync ? pick_log2p(ctx, n) : 0;
    const DecTableSet *tabs = mode.tables ? mode.tables : fixed_dec_tables(amvlib);
    if (!tabs) return fail(ctx, AMV_ERR_CUDA, "decoder tables are not on the device");
    uint64_t *slot_off; uint32_t *scan_len; int32_t *st = status; LaneStart *starts = nullptr; uint8_t *scratch;
    uint32_t *rounds; uint32_t *tokens; uint32_t *blk_off;
    const uint64_t scratch_bytes = payload_bytes + (uint64_t)(kSlotPad + 16) * n + 64;
    ENSURE(WS_SLOT_OFF, sizeof(uint64_t) * n, slot_off);
    ENSURE(WS_SCAN_LEN, sizeof(uint32_t) * n, scan_len);
    ENSURE(WS_SCRATCH, scratch_bytes, scratch);
    ENSURE(WS_ROUNDS, sizeof(uint32_t), rounds);
    // at most one token per 2 scan bits: 16 bytes of 32-bit tokens per scan byte, 8 of 16-bit ones (the pass chosen below)
    const bool tok16_ws = !amvlib && !mode.tables && !mode.hdr && ctx->opt_token_pass == 2;
    ENSURE(WS_TOKENS, scratch_bytes * (tok16_ws ? 8 : 16) + 1024, tokens);
    ENSURE(WS_BLKOFF, sizeof(uint32_t) * (size_t)n * g.nblk, blk_off);
    if (!st) ENSURE(WS_STATUS, sizeof(int32_t) * n, st);
    if (log2p) ENSURE(WS_STARTS, sizeof(LaneStart) * ((size_t)n << log2p), starts);

    launch_scan_sizes(pkt_size, n, 15u, kSlotPad, slot_off, nullptr, ctx->stream);
    { ScopedTimer tm(ctx, KK_UNSTUFF);
      launch_unstuff(pkts, pkts_bytes, pkt_off, pkt_size, n, scratch, slot_off, scratch_bytes, scan_len, st, mode.head,
                     mode.literal, ctx->stream); }
    int lc = 2;
    uint8_t *qtab = nullptr;
    if (mode.hdr) {
        ENSURE(WS_QTAB, (size_t)n * 128, qtab);
        launch_mjpeg_check(pkts, pkts_bytes, pkt_off, pkt_size, n, mode.hdr, mode.hdr_len, mode.qpos[0], mode.qpos[1], qtab,
                           scan_len, st, ctx->stream);
        lc++;
    }
    if (log2p) {
        CK(cudaMemsetAsync(rounds, 0, sizeof(uint32_t), ctx->stream));
        { ScopedTimer tm(ctx, KK_SYNC);
          launch_vlc_sync(scratch, slot_off, scan_len, n, log2p, starts, rounds, amvlib, tabs, qtab, g.nl, g.nc,
                          !mode.tables && ctx->opt_token_pass != 0, ctx->stream); }
        lc++;
    }
    // fixed AMV / SP5X tables: the 16-bit token pass; amvlib flavour and plain JPEG (own tables, per-frame quantisers): 32-bit tokens
    const bool fixed = !amvlib && !mode.tables && !mode.hdr;
    const bool tok16 = fixed && ctx->opt_token_pass == 2;
    { ScopedTimer tm(ctx, KK_TOKENS);
      if (fixed && ctx->opt_token_pass == 1) launch_vlc_tokens_lean(scratch, slot_off, scan_len, pkt_size, n, log2p, starts, g.nblk, tokens, blk_off, st, tabs, g.nl, g.nc, ctx->stream);
      else if (tok16) launch_vlc_tokens16(scratch, slot_off, scan_len, pkt_size, n, log2p, starts, g.nblk, reinterpret_cast<uint16_t *>(tokens), blk_off, st, tabs, g.nl, g.nc, ctx->stream);
      else launch_vlc_tokens(scratch, slot_off, scan_len, pkt_size, n, log2p, starts, g.nblk, tokens, blk_off, st, amvlib, tabs, qtab, g.nl, g.nc, mode.restart, ctx->stream); }
    lc++;
    F.slot_off = slot_off; F.scan_len = scan_len; F.tokens = tokens; F.blk_off = blk_off; F.st = st; F.launches = lc;
    F.tok16 = tok16; F.tabs = tabs;
    return AMV_OK;
}

int decode_device(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                  const uint32_t *pkt_size, int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c,
                  uint64_t fs_y, uint64_t fs_c, int32_t *status, uint64_t payload_bytes, const DecMode &mode = kModeAmv) {
    const Geom g = mode.geom(w, h);
    DecodeFront F;
    int r = decode_front(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, g, status, payload_bytes, false, F, mode);
    if (r != AMV_OK) return r;
    { ScopedTimer tm(ctx, KK_IDCT);
      if (F.tok16) launch_idct16(reinterpret_cast<const uint16_t *>(F.tokens), F.blk_off, F.slot_off, F.scan_len, n, g, F.tabs, y, u, v, ls_y, ls_c, fs_y, fs_c, ctx->stream);
      else launch_idct(F.tokens, F.blk_off, F.slot_off, F.scan_len, n, g, y, u, v, ls_y, ls_c, fs_y, fs_c, ctx->stream); }
    return check_launch(ctx, "decode kernels", F.launches + 1);
}

// amvlib flavour: same front half with amvlib's quantisers / DC chain / zigzag, then Chen-Wang IDCT + BGR24
int decode_bgr_device(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                      const uint32_t *pkt_size, int n, int w, int h, uint8_t *bgr, int line_bytes, uint64_t frame_stride,
                      int32_t *status, uint64_t payload_bytes) {
    const Geom g = make_geom(w, h);
    DecodeFront F;
    int r = decode_front(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, g, status, payload_bytes, true, F);
    if (r != AMV_OK) return r;
    { ScopedTimer tm(ctx, KK_IDCT_BGR);
      launch_idct_bgr(F.tokens, F.blk_off, F.slot_off, F.scan_len, n, g, bgr, line_bytes, frame_stride, ctx->stream); }
    return check_launch(ctx, "amvlib decode kernels", F.launches + 1);
}

bool encode_geometry_ok(int w, int h) {
    if (w < 2 || h < 2) return false;
    const Geom g = make_geom(w, h);
    // the reference starts reading each plane at rows y0 / c0 and walks up h resp. h>>1 rows
    // (mjpegenc.c:467-470, mpegvideo_enc.c:857-879); outside this set it reads out of the picture
    return g.y0 == h - 1 && g.c0 <= g.ch - 1 && g.c0 - ((h >> 1) - 1) >= 0;
}

int encode_device(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y,
                  uint64_t fs_c, int n, int w, int h, const int32_t *qscale, uint8_t *out, uint64_t out_cap,
                  uint32_t pkt_cap, int layout, uint64_t *out_off, uint32_t *out_size, int32_t *status,
                  bool reset_carry = true, uint64_t slot_base = 0) {
    const Geom g = make_geom(w, h);
    int32_t *st = status;
    if (!st) ENSURE(WS_STATUS, sizeof(int32_t) * n, st);
    int32_t *redo = nullptr;
    if (ctx->opt_encode_rounds) ENSURE(WS_REDO, sizeof(int32_t) * n, redo);
    const int enc_launches = redo ? 2 : 1;
    if (layout == AMV_LAYOUT_SLOTS) {
        if (slot_base + (uint64_t)pkt_cap * n > out_cap) return fail(ctx, AMV_ERR_ARG, "out_cap < n * pkt_cap");
        { ScopedTimer tm(ctx, KK_ENCODE);
          launch_encode(y, u, v, ls_y, ls_c, fs_y, fs_c, n, g, qscale, out + slot_base, pkt_cap, pkt_cap, out_size, st, redo, ctx->opt_encode_rounds, ctx->stream); }
        launch_slot_offsets(out_off, n, pkt_cap, slot_base, ctx->stream);
        return check_launch(ctx, "encode kernels", 1 + enc_launches);
    }
    // packed: encode into 16-byte aligned workspace slots, scan the sizes, compact
    const uint64_t stride = ((uint64_t)pkt_cap + 15) & ~15ull;
    int sub = (int)(ctx->opt_slot_ws_bytes / stride);
    if (sub < 1) sub = 1;
    if (sub > n) sub = n;
    uint8_t *slots; uint64_t *carry;
    ENSURE(WS_SLOTS, stride * sub + 64, slots);
    ENSURE(WS_CARRY, sizeof(uint64_t), carry);
    if (reset_carry) CK(cudaMemsetAsync(carry, 0, sizeof(uint64_t), ctx->stream));
    int lc = 0;
    for (int f0 = 0; f0 < n; f0 += sub) {
        const int m = n - f0 < sub ? n - f0 : sub;
        { ScopedTimer tm(ctx, KK_ENCODE);
          launch_encode(y + fs_y * f0, u + fs_c * f0, v + fs_c * f0, ls_y, ls_c, fs_y, fs_c, m, g, qscale ? qscale + f0 : nullptr,
                        slots, stride, pkt_cap, out_size + f0, st + f0, redo ? redo + f0 : nullptr, ctx->opt_encode_rounds, ctx->stream); }
        launch_scan_sizes(out_size + f0, m, 0u, 0u, out_off + f0, carry, ctx->stream);
        { ScopedTimer tm(ctx, KK_COMPACT);
          launch_compact(slots, stride, out_size + f0, out_off + f0, m, out, out_cap, st + f0, ctx->stream); }
        lc += 2 + enc_launches;
    }
    return check_launch(ctx, "encode kernels", lc);
}

// ------------------------------------------------------------------------------------ host staging
// Copies a host array into a device mirror (workspace id) and returns the device pointer.
int to_device(amv_ctx *ctx, WsId id, const void *host, size_t bytes, void **dev) {
    int r = ensure(ctx, id, bytes ? bytes : 1, dev);
    if (r != AMV_OK) return r;
    if (bytes) CK(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return AMV_OK;
}

#define TO_DEVICE(id, host, bytes, ptr)                                           \
    do {                                                                          \
        void *p_ = nullptr;                                                       \
        int r_ = to_device(ctx, id, (host), (bytes), &p_);                        \
        if (r_ != AMV_OK) return r_;                                              \
        ptr = reinterpret_cast<decltype(ptr)>(p_);                                \
    } while (0)

// planes: host (ls, fs) layout <-> tight device layout (ls = width, fs = width*height)
int copy_planes(amv_ctx *ctx, uint8_t *dev, uint8_t *host, int width, int height, int ls, uint64_t fs, int n, bool to_host) {
    const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    const uint64_t tight = (uint64_t)width * height;
    if (ls == width && fs == tight) {
        if (to_host) CK(cudaMemcpyAsync(host, dev, tight * n, kind, ctx->stream));
        else         CK(cudaMemcpyAsync(dev, host, tight * n, kind, ctx->stream));
        return AMV_OK;
    }
    for (int i = 0; i < n; i++) {
        if (to_host) CK(cudaMemcpy2DAsync(host + fs * i, ls, dev + tight * i, width, width, height, kind, ctx->stream));
        else         CK(cudaMemcpy2DAsync(dev + tight * i, width, host + fs * i, ls, width, height, kind, ctx->stream));
    }
    return AMV_OK;
}


// ------------------------------------------------------------------------------------ host pipeline
// AMV_MEM_HOST calls run as a three-stage pipeline over chunks of frames: copy-in of chunk i+1
// (stream s_in), kernels of chunk i (the context's stream), copy-out of chunk i-1 (stream s_out),
// so PCIe traffic in both directions overlaps the compute.  Chunk buffers are rings of 3.
constexpr int kRing = 3;

int ensure_pipeline(amv_ctx *ctx, size_t meta_bytes) {
    if (!ctx->s_in) CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    if (meta_bytes > ctx->pinned_meta_cap) {
        CK(cudaStreamSynchronize(ctx->s_out));
        if (ctx->pinned_meta) cudaFreeHost(ctx->pinned_meta);
        ctx->pinned_meta = nullptr; ctx->pinned_meta_cap = 0;
        CK(cudaMallocHost(&ctx->pinned_meta, meta_bytes + 4096));
        ctx->pinned_meta_cap = meta_bytes + 4096;
    }
    return AMV_OK;
}

int host_chunk_frames(const amv_ctx *ctx, int n, size_t frame_bytes) {
    if (ctx->opt_host_chunk > 0) return ctx->opt_host_chunk < n ? ctx->opt_host_chunk : n;
    // ~8 stages per call, between 96 MB and 512 MB of frames per stage: a stage costs ~10 launches
    // and one host wake-up, which must stay small against its PCIe time (~2 ms per 100 MB)
    int c = (n + 7) / 8;
    const int lo = (int)((96u << 20) / frame_bytes) + 1, hi = (int)((512u << 20) / frame_bytes) + 1;
    if (c < lo) c = lo;
    if (c > hi) c = hi;
    return c < n ? c : n;
}

// Large transfers are issued in pieces: a copy engine serves its queue in submission order, so a
// short copy of another stream (metadata, the other direction's packets) would otherwise wait
// behind a 100+ MB transfer and stall that stream's whole pipeline.
constexpr size_t kCopyPiece = 16u << 20;
cudaError_t copy_pieces(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
    for (size_t o = 0; o < bytes; o += kCopyPiece) {
        const size_t m = bytes - o < kCopyPiece ? bytes - o : kCopyPiece;
        cudaError_t e = cudaMemcpyAsync((uint8_t *)dst + o, (const uint8_t *)src + o, m, kind, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// planes of frames [f0, f0+m): host (ls, fs) layout <-> tight device layout, on stream s
int copy_planes_on(amv_ctx *ctx, cudaStream_t s, uint8_t *dev, uint8_t *host, int width, int height, int ls, uint64_t fs,
                   int m, bool to_host) {
    const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    const uint64_t tight = (uint64_t)width * height;
    if (ls == width && fs == tight) {
        if (to_host) CK(copy_pieces(host, dev, tight * m, kind, s));
        else         CK(copy_pieces(dev, host, tight * m, kind, s));
        return AMV_OK;
    }
    for (int i = 0; i < m; i++) {
        if (to_host) CK(cudaMemcpy2DAsync(host + fs * i, ls, dev + tight * i, width, width, height, kind, s));
        else         CK(cudaMemcpy2DAsync(dev + tight * i, width, host + fs * i, ls, width, height, kind, s));
    }
    return AMV_OK;
}

struct EventRing {
    cudaEvent_t e[kRing] = { nullptr, nullptr, nullptr };
    bool used[kRing] = { false, false, false };
    ~EventRing() { for (int i = 0; i < kRing; i++) if (e[i]) cudaEventDestroy(e[i]); }
    cudaError_t init() {
        for (int i = 0; i < kRing; i++) {
            cudaError_t r = cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming);
            if (r != cudaSuccess) return r;
        }
        return cudaSuccess;
    }
};

// Device-accessible alias of a pinned (page-locked) host buffer, or NULL if the buffer is pageable.
// Under UVA every cudaMallocHost / cudaHostRegister allocation is mapped into the device's address
// space, so kernels can read and write it directly over PCIe.
void *device_view(const void *host_ptr) {
    if (!host_ptr) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type == cudaMemoryTypeHost && a.devicePointer) return a.devicePointer;
    return nullptr;
}

// ------------------------------------------------------------------------------------ small host calls
// A call that moves a few frames (the AVCodec callbacks hand over ONE frame or packet per call, ffmpeg.c:1083,814) does not
// need the three-stream copy pipeline above, and cannot afford it: event rings, three stream synchronisations and
// per-plane pageable copies cost ~2 ms per call.  Here everything rides the context's one stream: the caller's bytes are
// gathered into one pinned block (a CPU memcpy), one H2D copy, the kernels, one D2H copy, ONE synchronisation, and a CPU
// memcpy into the caller's (strided) buffers.
constexpr size_t kSmallCallBytes = 8u << 20;

int ensure_small(amv_ctx *ctx, size_t bytes) {
    if (bytes > ctx->pinned_small_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned_small) cudaFreeHost(ctx->pinned_small);
        ctx->pinned_small = nullptr; ctx->pinned_small_cap = 0;
        const size_t want = bytes + bytes / 4 + 65536;
        CK(cudaMallocHost(&ctx->pinned_small, want));
        ctx->pinned_small_cap = want;
    }
    return AMV_OK;
}

inline size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

void copy_rows(uint8_t *dst, size_t dst_ls, const uint8_t *src, size_t src_ls, int width, int rows) {
    if (dst_ls == (size_t)width && src_ls == (size_t)width) { memcpy(dst, src, (size_t)width * rows); return; }
    for (int r = 0; r < rows; r++) memcpy(dst + r * dst_ls, src + r * src_ls, width);
}

int decode_host_small(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size,
                      int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                      int32_t *status, const DecMode &mode) {
    const Geom gm = mode.geom(w, h);
    const int cw = gm.cw, ch = gm.ch;
    const uint64_t ty = (uint64_t)w * h, tc = (uint64_t)cw * ch;
    // the byte range the packets span (frames whose range is outside the buffer keep their offsets: the kernels flag them)
    uint64_t lo = UINT64_MAX, hi = 0, payload = 0;
    for (int f = 0; f < n; f++) {
        const uint64_t a = pkt_off[f], b = a + pkt_size[f];
        payload += ((uint64_t)pkt_size[f] + 15) & ~15ull;
        if (b > pkts_bytes || b < a) continue;
        if (a < lo) lo = a;
        if (b > hi) hi = b;
    }
    if (hi < lo) { lo = 0; hi = 0; }
    const size_t span = (size_t)(hi - lo);
    // pinned block: [packets][offsets][sizes] in, [status][planes] out
    const size_t o_off = al256(span), o_sz = o_off + al256(8 * (size_t)n), in_bytes = o_sz + al256(4 * (size_t)n);
    const size_t o_st = in_bytes, o_y = o_st + al256(4 * (size_t)n), o_u = o_y + ty * n, o_v = o_u + tc * n, all_bytes = al256(o_v + tc * n);
    int r = ensure_small(ctx, all_bytes);
    if (r != AMV_OK) return r;
    uint8_t *pin = static_cast<uint8_t *>(ctx->pinned_small), *dev;
    ENSURE(WS_H_A, all_bytes, dev);
    if (span) memcpy(pin, pkts + lo, span);
    uint64_t *p_off = reinterpret_cast<uint64_t *>(pin + o_off);
    for (int f = 0; f < n; f++) {
        const uint64_t a = pkt_off[f], b = a + pkt_size[f];
        p_off[f] = (b > pkts_bytes || b < a) ? UINT64_MAX - 0xffff : a - lo;     // out of range stays out of range
    }
    memcpy(pin + o_sz, pkt_size, 4 * (size_t)n);
    CK(cudaMemcpyAsync(dev, pin, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    r = decode_device(ctx, dev, span, reinterpret_cast<uint64_t *>(dev + o_off), reinterpret_cast<uint32_t *>(dev + o_sz), n, w, h,
                      dev + o_y, dev + o_u, dev + o_v, w, cw, ty, tc, reinterpret_cast<int32_t *>(dev + o_st), payload, mode);
    if (r != AMV_OK) return r;
    CK(cudaMemcpyAsync(pin + o_st, dev + o_st, all_bytes - o_st, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int32_t *p_st = reinterpret_cast<const int32_t *>(pin + o_st);
    for (int f = 0; f < n; f++) {
        if (p_st[f] & (AMV_ST_RANGE | AMV_ST_HEADER)) continue;        // nothing was decoded for this frame: the caller's planes stay
        copy_rows(y + fs_y * f, ls_y, pin + o_y + ty * f, w, w, h);
        copy_rows(u + fs_c * f, ls_c, pin + o_u + tc * f, cw, cw, ch);
        copy_rows(v + fs_c * f, ls_c, pin + o_v + tc * f, cw, cw, ch);
    }
    if (status) memcpy(status, p_st, 4 * (size_t)n);
    return AMV_OK;
}

int encode_host_small(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y,
                      uint64_t fs_c, int n, int w, int h, const int32_t *qscale, uint8_t *out, uint64_t out_cap, uint32_t pkt_cap,
                      int layout, uint64_t *out_off, uint32_t *out_size, int32_t *status) {
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    const uint64_t ty = (uint64_t)w * h, tc = (uint64_t)cw * ch;
    // a packet never exceeds 3000 bytes per macroblock in the reference (MAX_MB_BYTES, mpegvideo.h:75); the device-side slot
    // is the smaller of the caller's pkt_cap and that bound, so a 1 MB bit buffer does not become 1 MB of D2H
    const uint64_t mb_bound = 3000ull * (uint64_t)((w + 15) / 16) * ((h + 15) / 16) + 1024;
    const uint32_t dcap = (uint32_t)((pkt_cap < mb_bound ? pkt_cap : mb_bound) + 15) & ~15u;
    if (layout == AMV_LAYOUT_SLOTS && (uint64_t)pkt_cap * n > out_cap) return fail(ctx, AMV_ERR_ARG, "out_cap < n * pkt_cap");
    const size_t o_u = ty * n, o_v = o_u + tc * n, o_q = al256(o_v + tc * n), in_bytes = o_q + al256(4 * (size_t)n);
    const size_t o_off = in_bytes, o_sz = o_off + al256(8 * (size_t)n), o_st = o_sz + al256(4 * (size_t)n), o_pk = o_st + al256(4 * (size_t)n);
    const size_t all_bytes = al256(o_pk + (size_t)dcap * n);
    int r = ensure_small(ctx, all_bytes);
    if (r != AMV_OK) return r;
    uint8_t *pin = static_cast<uint8_t *>(ctx->pinned_small), *dev;
    ENSURE(WS_H_A, all_bytes, dev);
    for (int f = 0; f < n; f++) {
        copy_rows(pin + ty * f, w, y + fs_y * f, ls_y, w, h);
        copy_rows(pin + o_u + tc * f, cw, u + fs_c * f, ls_c, cw, ch);
        copy_rows(pin + o_v + tc * f, cw, v + fs_c * f, ls_c, cw, ch);
    }
    if (qscale) memcpy(pin + o_q, qscale, 4 * (size_t)n);
    CK(cudaMemcpyAsync(dev, pin, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    r = encode_device(ctx, dev, dev + o_u, dev + o_v, w, cw, ty, tc, n, w, h, qscale ? reinterpret_cast<int32_t *>(dev + o_q) : nullptr,
                      dev + o_pk, (uint64_t)dcap * n, dcap, AMV_LAYOUT_SLOTS, reinterpret_cast<uint64_t *>(dev + o_off),
                      reinterpret_cast<uint32_t *>(dev + o_sz), reinterpret_cast<int32_t *>(dev + o_st));
    if (r != AMV_OK) return r;
    // sizes and status first; then exactly the bytes that were produced
    CK(cudaMemcpyAsync(pin + o_sz, dev + o_sz, o_pk - o_sz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t *p_sz = reinterpret_cast<const uint32_t *>(pin + o_sz);
    int32_t *p_st = reinterpret_cast<int32_t *>(pin + o_st);
    for (int f = 0; f < n; f++)
        if (p_sz[f]) CK(cudaMemcpyAsync(pin + o_pk + (size_t)dcap * f, dev + o_pk + (size_t)dcap * f, p_sz[f], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint64_t pos = 0;
    for (int f = 0; f < n; f++) {
        uint32_t sz = p_sz[f];
        // a packet that fits the device slot but not the caller's: "encoded frame too large"
        const uint64_t dst = layout == AMV_LAYOUT_SLOTS ? (uint64_t)pkt_cap * f : pos;
        if (p_st[f] == 0 && (sz > pkt_cap || dst + sz > out_cap)) { p_st[f] = AMV_ST_NOSPACE; sz = 0; }
        if (sz) memcpy(out + dst, pin + o_pk + (size_t)dcap * f, sz);
        out_off[f] = dst; out_size[f] = sz;
        pos += sz;
    }
    if (status) memcpy(status, p_st, 4 * (size_t)n);
    return AMV_OK;
}

// How the host path moves data:
//  * bulk planes travel by DMA (cudaMemcpyAsync, in <= 16 MB pieces) on their own copy streams, in a
//    3-deep ring of chunks that overlaps copy-in, kernels and copy-out;
//  * everything small and latency-critical -- per-frame metadata, the decoder's input packets, the
//    encoder's packed output -- never enters a copy-engine queue: the kernels read / write the
//    caller's pinned buffers (or the context's own pinned bounce buffer) in place.  A copy engine
//    serves its queue strictly in submission order, so a 16-byte status copy queued behind another
//    call's gigabyte of planes would otherwise stall a whole pipeline (measured: two concurrent
//    calls, one H2D-bound and one D2H-bound, ran at the SUM of their times).

int decode_host(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size,
                int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                int32_t *status, const DecMode &mode = kModeAmv) {
    const Geom gm = mode.geom(w, h);
    const int cw = gm.cw, ch = gm.ch;
    const uint64_t ty = (uint64_t)w * h, tc = (uint64_t)cw * ch;
    if (ctx->opt_small_calls && (ty + 2 * tc) * n + pkts_bytes <= kSmallCallBytes)
        return decode_host_small(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, mode);
    const int C = host_chunk_frames(ctx, n, (size_t)(ty + 2 * tc));
    // pinned bounce: status (4) + offsets (8) + sizes (4) per frame
    int r = ensure_pipeline(ctx, 16 * (size_t)n);
    if (r != AMV_OK) return r;
    uint64_t *p_off = reinterpret_cast<uint64_t *>(ctx->pinned_meta);
    uint32_t *p_sz = reinterpret_cast<uint32_t *>(p_off + n);
    int32_t *p_st = reinterpret_cast<int32_t *>(p_sz + n);
    // offsets / sizes: read in place if the caller's arrays are pinned, else from our pinned copy
    const uint64_t *v_off = static_cast<const uint64_t *>(device_view(pkt_off));
    const uint32_t *v_sz = static_cast<const uint32_t *>(device_view(pkt_size));
    if (!v_off) { memcpy(p_off, pkt_off, sizeof(uint64_t) * n); v_off = static_cast<const uint64_t *>(device_view(p_off)); }
    if (!v_sz) { memcpy(p_sz, pkt_size, sizeof(uint32_t) * n); v_sz = static_cast<const uint32_t *>(device_view(p_sz)); }
    int32_t *v_st = static_cast<int32_t *>(device_view(p_st));
    if (!v_off || !v_sz || !v_st) return fail(ctx, AMV_ERR_CUDA, "pinned bounce buffer is not device-mapped");
    const uint8_t *v_pk = ctx->opt_zero_copy_packets ? static_cast<const uint8_t *>(device_view(pkts)) : nullptr;   // zero-copy packets if pinned
    uint8_t *d_pk = nullptr, *d_y, *d_u, *d_v; int32_t *d_st;
    if (!v_pk) ENSURE(WS_H_A, pkts_bytes ? pkts_bytes : 1, d_pk);
    ENSURE(WS_H_D, ty * C * kRing, d_y);
    ENSURE(WS_H_E, tc * C * kRing, d_u);
    ENSURE(WS_H_F, tc * C * kRing, d_v);
    ENSURE(WS_H_G, sizeof(int32_t) * n, d_st);
    EventRing e_in, e_cmp, e_out;
    if (e_in.init() != cudaSuccess || e_cmp.init() != cudaSuccess || e_out.init() != cudaSuccess)
        return fail(ctx, AMV_ERR_CUDA, "cudaEventCreate");
    int rc = AMV_OK;
    const int nchunks = (n + C - 1) / C;
    for (int i = 0; i < nchunks && rc == AMV_OK; i++) {
        const int f0 = i * C, m = n - f0 < C ? n - f0 : C, slot = i % kRing;
        uint64_t lo = UINT64_MAX, hi = 0, payload = 0;
        for (int f = f0; f < f0 + m; f++) {
            const uint64_t a = pkt_off[f], b = a + pkt_size[f];
            payload += ((uint64_t)pkt_size[f] + 15) & ~15ull;
            if (b > pkts_bytes || b < a) continue;        // reported per frame by the kernels (range_ok: no wrap for any offset)
            if (a < lo) lo = a;
            if (b > hi) hi = b;
        }
        if (!v_pk) {       // pageable packets: staged copy of this chunk's byte range
            if (hi > lo) CK(copy_pieces(d_pk + lo, pkts + lo, hi - lo, cudaMemcpyHostToDevice, ctx->s_in));
            CK(cudaEventRecord(e_in.e[slot], ctx->s_in));
            CK(cudaStreamWaitEvent(ctx->stream, e_in.e[slot], 0));
        }
        if (e_out.used[slot]) CK(cudaStreamWaitEvent(ctx->stream, e_out.e[slot], 0));      // ring slot drained?
        rc = decode_device(ctx, v_pk ? v_pk : d_pk, pkts_bytes, v_off + f0, v_sz + f0, m, w, h, d_y + ty * C * slot,
                           d_u + tc * C * slot, d_v + tc * C * slot, w, cw, ty, tc, d_st + f0, payload, mode);
        if (rc != AMV_OK) break;
        CK(cudaEventRecord(e_cmp.e[slot], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_out, e_cmp.e[slot], 0));
        if ((rc = copy_planes_on(ctx, ctx->s_out, d_y + ty * C * slot, y + fs_y * f0, w, h, ls_y, fs_y, m, true)) != AMV_OK) break;
        if ((rc = copy_planes_on(ctx, ctx->s_out, d_u + tc * C * slot, u + fs_c * f0, cw, ch, ls_c, fs_c, m, true)) != AMV_OK) break;
        if ((rc = copy_planes_on(ctx, ctx->s_out, d_v + tc * C * slot, v + fs_c * f0, cw, ch, ls_c, fs_c, m, true)) != AMV_OK) break;
        CK(cudaEventRecord(e_out.e[slot], ctx->s_out));
        e_out.used[slot] = true;
    }
    if (rc == AMV_OK) { launch_export_meta(nullptr, nullptr, d_st, nullptr, nullptr, v_st, n, ctx->stream); ctx->launches++; }
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->s_out);
    if (rc != AMV_OK) return rc;
    if (e != cudaSuccess) return fail(ctx, AMV_ERR_CUDA, "decode pipeline", e);
    if (status) memcpy(status, p_st, sizeof(int32_t) * n);
    return AMV_OK;
}

int encode_host(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y,
                uint64_t fs_c, int n, int w, int h, const int32_t *qscale, uint8_t *out, uint64_t out_cap, uint32_t pkt_cap,
                int layout, uint64_t *out_off, uint32_t *out_size, int32_t *status) {
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    const uint64_t ty = (uint64_t)w * h, tc = (uint64_t)cw * ch;
    if (ctx->opt_small_calls && (ty + 2 * tc) * n <= kSmallCallBytes / 2)
        return encode_host_small(ctx, y, u, v, ls_y, ls_c, fs_y, fs_c, n, w, h, qscale, out, out_cap, pkt_cap, layout, out_off, out_size, status);
    const int C = host_chunk_frames(ctx, n, (size_t)(ty + 2 * tc));
    // pinned metadata: offsets (8), sizes (4), status (4), qscale (4) per frame
    int r = ensure_pipeline(ctx, 20 * (size_t)n);
    if (r != AMV_OK) return r;
    uint64_t *p_off = reinterpret_cast<uint64_t *>(ctx->pinned_meta);
    uint32_t *p_sz = reinterpret_cast<uint32_t *>(p_off + n);
    int32_t *p_st = reinterpret_cast<int32_t *>(p_sz + n);
    int32_t *p_q = p_st + n;
    uint64_t *v_off = static_cast<uint64_t *>(device_view(p_off));
    uint32_t *v_sz = static_cast<uint32_t *>(device_view(p_sz));
    int32_t *v_st = static_cast<int32_t *>(device_view(p_st));
    const int32_t *v_q = nullptr;
    if (qscale) {
        v_q = static_cast<const int32_t *>(device_view(qscale));
        if (!v_q) { memcpy(p_q, qscale, sizeof(int32_t) * n); v_q = static_cast<const int32_t *>(device_view(p_q)); }
    }
    if (!v_off || !v_sz || !v_st || (qscale && !v_q)) return fail(ctx, AMV_ERR_CUDA, "pinned bounce buffer is not device-mapped");
    // packed output into a pinned caller buffer is written in place by k_compact (128-bit stores over
    // PCIe); slot layout (byte stores) and pageable buffers are staged in device memory and copied
    uint8_t *v_out = layout == AMV_LAYOUT_PACKED ? static_cast<uint8_t *>(device_view(out)) : nullptr;
    uint8_t *d_y, *d_u, *d_v, *d_out = nullptr; uint64_t *d_off; uint32_t *d_sz; int32_t *d_st;
    ENSURE(WS_H_A, ty * C * kRing, d_y);
    ENSURE(WS_H_B, tc * C * kRing, d_u);
    ENSURE(WS_H_C, tc * C * kRing, d_v);
    const uint64_t dcap = layout == AMV_LAYOUT_SLOTS ? (uint64_t)pkt_cap * n : out_cap;
    if (layout == AMV_LAYOUT_SLOTS && dcap > out_cap) return fail(ctx, AMV_ERR_ARG, "out_cap < n * pkt_cap");
    if (!v_out) ENSURE(WS_H_E, dcap ? dcap : 1, d_out);
    ENSURE(WS_H_F, sizeof(uint64_t) * n, d_off);
    ENSURE(WS_H_G, sizeof(uint32_t) * n, d_sz);
    ENSURE(WS_H_H, sizeof(int32_t) * n, d_st);
    const int nchunks = (n + C - 1) / C;
    EventRing e_in;
    if (e_in.init() != cudaSuccess) return fail(ctx, AMV_ERR_CUDA, "cudaEventCreate");
    std::vector<cudaEvent_t> e_cmp(nchunks, nullptr);
    for (int i = 0; i < nchunks; i++)
        if (cudaEventCreateWithFlags(&e_cmp[i], cudaEventDisableTiming) != cudaSuccess) return fail(ctx, AMV_ERR_CUDA, "cudaEventCreate");
    int rc = AMV_OK;
    for (int i = 0; i < nchunks && rc == AMV_OK; i++) {
        const int f0 = i * C, m = n - f0 < C ? n - f0 : C, slot = i % kRing;
        // at most kRing - 1 chunks in flight: bounds what this call keeps queued on the H2D engine and
        // makes ring slot `slot` (last used by chunk i - kRing) free
        if (i >= kRing - 1) CK(cudaEventSynchronize(e_cmp[i - (kRing - 1)]));
        if ((rc = copy_planes_on(ctx, ctx->s_in, d_y + ty * C * slot, const_cast<uint8_t *>(y) + fs_y * f0, w, h, ls_y, fs_y, m, false)) != AMV_OK) break;
        if ((rc = copy_planes_on(ctx, ctx->s_in, d_u + tc * C * slot, const_cast<uint8_t *>(u) + fs_c * f0, cw, ch, ls_c, fs_c, m, false)) != AMV_OK) break;
        if ((rc = copy_planes_on(ctx, ctx->s_in, d_v + tc * C * slot, const_cast<uint8_t *>(v) + fs_c * f0, cw, ch, ls_c, fs_c, m, false)) != AMV_OK) break;
        CK(cudaEventRecord(e_in.e[slot], ctx->s_in));
        CK(cudaStreamWaitEvent(ctx->stream, e_in.e[slot], 0));
        rc = encode_device(ctx, d_y + ty * C * slot, d_u + tc * C * slot, d_v + tc * C * slot, w, cw, ty, tc, m, w, h,
                           v_q ? v_q + f0 : nullptr, v_out ? v_out : d_out, dcap, pkt_cap, layout, d_off + f0, d_sz + f0,
                           d_st + f0, /*reset_carry=*/i == 0, /*slot_base=*/(uint64_t)pkt_cap * f0);
        if (rc != AMV_OK) break;
        launch_export_meta(d_off + f0, d_sz + f0, d_st + f0, v_off + f0, v_sz + f0, v_st + f0, m, ctx->stream);
        ctx->launches++;
        CK(cudaEventRecord(e_cmp[i], ctx->stream));
    }
    cudaStreamSynchronize(ctx->s_in);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < nchunks; i++) cudaEventDestroy(e_cmp[i]);
    if (rc != AMV_OK) return rc;
    if (e != cudaSuccess) return fail(ctx, AMV_ERR_CUDA, "encode pipeline", e);
    if (!v_out) {      // staged output: the metadata is home, send the bytes that were produced
        uint64_t used = 0;
        if (layout == AMV_LAYOUT_SLOTS) used = (uint64_t)pkt_cap * n;
        else for (int f = 0; f < n; f++) if (p_off[f] + p_sz[f] > used) used = p_off[f] + p_sz[f];
        if (used > out_cap) used = out_cap;
        if (used) CK(copy_pieces(out, d_out, used, cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaStreamSynchronize(ctx->s_out));
    }
    memcpy(out_off, p_off, sizeof(uint64_t) * n);
    memcpy(out_size, p_sz, sizeof(uint32_t) * n);
    if (status) memcpy(status, p_st, sizeof(int32_t) * n);
    return AMV_OK;
}


bool bad_mem(int mem) { return mem != AMV_MEM_HOST && mem != AMV_MEM_DEVICE; }

}  // namespace

// ================================================================================== public ABI
extern "C" {

AMV_API int amv_version(void) { return AMVCUDA_VERSION; }

AMV_API const char *amv_strerror(int err) {
    switch (err) {
    case AMV_OK: return "ok";
    case AMV_ERR_ARG: return "invalid argument";
    case AMV_ERR_NODEVICE: return "no usable CUDA device (libamvcuda has no CPU path)";
    case AMV_ERR_CUDA: return "CUDA runtime error";
    case AMV_ERR_NOMEM: return "out of device memory";
    case AMV_ERR_UNSUPPORTED: return "unsupported option";
    default: return "unknown error";
    }
}

AMV_API const char *amv_last_error(const amv_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
AMV_API uint64_t amv_launch_count(const amv_ctx *ctx) { return ctx ? ctx->launches : 0; }

AMV_API int amv_qscale_from_quality(int quality, int qmin, int qmax) {
    int q = (quality * 139 + 128 * 64) >> 14;          // update_qscale, mpegvideo_enc.c:143-148
    return q < qmin ? qmin : (q > qmax ? qmax : q);
}

AMV_API void *amv_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
AMV_API void amv_host_free(void *p) { if (p) cudaFreeHost(p); }

AMV_API int amv_create(const amv_params *params, amv_ctx **out_ctx) {
    if (!out_ctx) return AMV_ERR_ARG;
    *out_ctx = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return AMV_ERR_NODEVICE; }
    int dev = params ? params->device : -1;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return AMV_ERR_NODEVICE; }
    if (dev >= ndev) return AMV_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return AMV_ERR_NODEVICE;
    if (prop.major != 10) return AMV_ERR_NODEVICE;     // kernels are built for sm_100a only
    DeviceGuard guard(dev);
    if (guard.err != cudaSuccess) return AMV_ERR_NODEVICE;
    amv_ctx *ctx = new (std::nothrow) amv_ctx();
    if (!ctx) return AMV_ERR_NOMEM;
    ctx->device = dev;
    if (params && params->stream) ctx->stream = (cudaStream_t)params->stream;
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return AMV_ERR_CUDA; }
        ctx->own_stream = true;
    }
    // tables and function attributes are per device: set them up for this context's device (a process-wide
    // "done" flag left every device after the first without the shared-memory opt-in)
    cudaError_t e = encode_setup_device();
    if (e == cudaSuccess) e = decode_setup_device();
    if (e == cudaSuccess) e = adpcm_setup_device();
    if (e == cudaSuccess) e = upload_dec_tables(ctx->stream);
    if (e == cudaSuccess) e = upload_enc_tables(ctx->stream);
    if (e == cudaSuccess) e = upload_adpcm_tables(ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { amv_destroy(ctx); return AMV_ERR_CUDA; }
    *out_ctx = ctx;
    return AMV_OK;
}

AMV_API void amv_destroy(amv_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < WS_COUNT; i++) if (ctx->ws[i].p) cudaFree(ctx->ws[i].p);
    for (size_t i = 0; i < ctx->evs.size(); i++) { cudaEventDestroy(ctx->evs[i].a); cudaEventDestroy(ctx->evs[i].b); }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->pinned_meta) cudaFreeHost(ctx->pinned_meta);
    if (ctx->pinned_small) cudaFreeHost(ctx->pinned_small);
    if (ctx->mj_tables) cudaFree(ctx->mj_tables);
    if (ctx->mj_hdr) cudaFree(ctx->mj_hdr);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

AMV_API int amv_set_stream(amv_ctx *ctx, void *cuda_stream) {
    if (!ctx) return AMV_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return AMV_OK;
}

AMV_API int amv_sync(amv_ctx *ctx) {
    if (!ctx) return AMV_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return AMV_OK;
}

AMV_API int amv_set_option(amv_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return AMV_ERR_ARG;
    if (!strcmp(key, "decode_log2_lanes")) { ctx->opt_log2p = (int)value; return AMV_OK; }
    if (!strcmp(key, "encode_slot_workspace_bytes")) { ctx->opt_slot_ws_bytes = (uint64_t)value; return AMV_OK; }
    if (!strcmp(key, "profile_events")) { ctx->opt_profile = value != 0; return AMV_OK; }
    if (!strcmp(key, "host_chunk_frames")) { ctx->opt_host_chunk = (int)value; return AMV_OK; }
    if (!strcmp(key, "encode_rounds")) {
        if (value < 0 || value > 8 || (value >= 5 && value <= 7)) return AMV_ERR_UNSUPPORTED;
        ctx->opt_encode_rounds = (int)value;
        return AMV_OK;
    }
    if (!strcmp(key, "resample_form")) {
        if (value < 0 || value > 2) return AMV_ERR_UNSUPPORTED;
        ctx->opt_resample_form = (int)value;
        return AMV_OK;
    }
    if (!strcmp(key, "scale_form")) {
        if (value < 0 || value > 2) return AMV_ERR_UNSUPPORTED;
        ctx->opt_scale_form = (int)value;
        return AMV_OK;
    }
    if (!strcmp(key, "adpcm_trellis")) {
        if (value < 0 || value > 5) return AMV_ERR_UNSUPPORTED;
        ctx->opt_trellis = (int)value;
        return AMV_OK;
    }
    if (!strcmp(key, "host_zero_copy_packets")) { ctx->opt_zero_copy_packets = value != 0; return AMV_OK; }
    if (!strcmp(key, "adpcm_form")) {
        if (value < 0 || value > 2) return AMV_ERR_UNSUPPORTED;
        ctx->opt_adpcm_form = (int)value;
        return AMV_OK;
    }
    if (!strcmp(key, "host_small_calls")) { ctx->opt_small_calls = value != 0; return AMV_OK; }
    if (!strcmp(key, "decode_token_pass")) {
        if (value < 0 || value > 2) return AMV_ERR_UNSUPPORTED;
        ctx->opt_token_pass = (int)value;
        return AMV_OK;
    }
    return AMV_ERR_UNSUPPORTED;
}

AMV_API int64_t amv_get_stat(amv_ctx *ctx, const char *key) {
    if (!ctx || !key) return -1;
    // what amv_mjpeg_configure read: bytes in front of the scan, where the per-frame quantisers sit
    if (!strcmp(key, "mjpeg_header_bytes")) return ctx->mj_hdr_len;
    if (!strcmp(key, "mjpeg_quant_offset_0")) return ctx->mj_qpos[0];
    if (!strcmp(key, "mjpeg_quant_offset_1")) return ctx->mj_qpos[1];
    if (!strcmp(key, "mjpeg_sync_table")) return ctx->mj_sync_ok ? 1 : 0;
    if (!strcmp(key, "mjpeg_chroma_width") || !strcmp(key, "mjpeg_chroma_height")) {
        const Geom g = make_geom_sampled(ctx->mj_w, ctx->mj_h, ctx->mj_samp[0], ctx->mj_samp[1], ctx->mj_samp[2], ctx->mj_samp[3]);
        return key[13] == 'w' ? g.cw : g.ch;
    }
    // "<kernel>_kernel_ns" / "<kernel>_kernel_launches": device time and launch count of one hot kernel
    // accumulated since the last query of that key pair (needs option profile_events = 1)
    for (int k = 0; k < KK_COUNT; k++) {
        char kn[64], kl[64];
        snprintf(kn, sizeof kn, "%s_kernel_ns", kKernelKindName[k]);
        snprintf(kl, sizeof kl, "%s_kernel_launches", kKernelKindName[k]);
        const bool want_ns = !strcmp(key, kn), want_l = !strcmp(key, kl);
        if (!want_ns && !want_l) continue;
        cudaStreamSynchronize(ctx->stream);
        double ns = 0; int64_t cnt = 0;
        for (size_t i = 0; i < ctx->evs.size(); i++) {
            if (ctx->evs[i].kind != k) continue;
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->evs[i].a, ctx->evs[i].b) == cudaSuccess) ns += (double)ms * 1e6;
            cnt++;
        }
        if (want_l) return cnt;
        // the ns query consumes the samples
        std::vector<EvPair> keep;
        for (size_t i = 0; i < ctx->evs.size(); i++) {
            if (ctx->evs[i].kind == k) { cudaEventDestroy(ctx->evs[i].a); cudaEventDestroy(ctx->evs[i].b); }
            else keep.push_back(ctx->evs[i]);
        }
        ctx->evs.swap(keep);
        return (int64_t)ns;
    }
    if (!strcmp(key, "decode_sync_rounds")) {
        uint32_t r = 0;
        if (!ctx->ws[WS_ROUNDS].p) return 0;
        if (cudaMemcpyAsync(&r, ctx->ws[WS_ROUNDS].p, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
        cudaStreamSynchronize(ctx->stream);
        return r;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------- decode
static int decode_frames_common(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                                const uint32_t *pkt_size, int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v,
                                int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int32_t *status, int mem, const DecMode &mode) {
    if (!ctx) return AMV_ERR_ARG;
    if (n < 0 || w <= 0 || h <= 0 || w > 16384 || h > 16384 || bad_mem(mem)) return fail(ctx, AMV_ERR_ARG, "bad n / dimensions / mem");
    if (n == 0) return AMV_OK;
    if (!pkts || !pkt_off || !pkt_size || !y || !u || !v) return fail(ctx, AMV_ERR_ARG, "null buffer");
    const Geom gm = mode.geom(w, h);
    const int cw = gm.cw, ch = gm.ch;
    if (ls_y < w || ls_c < cw || fs_y < (uint64_t)ls_y * (h - 1) + w || fs_c < (uint64_t)ls_c * (ch - 1) + cw)
        return fail(ctx, AMV_ERR_ARG, "strides smaller than the picture");
    ON_DEVICE(ctx);
    if (mem == AMV_MEM_DEVICE)
        return decode_device(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, pkts_bytes, mode);
    return decode_host(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, mode);
}

AMV_API int amv_decode_frames(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                              const uint32_t *pkt_size, int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v,
                              int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int32_t *status, int mem) {
    return decode_frames_common(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, mem, kModeAmv);
}

AMV_API int amv_decode_frames_sp5x(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                                   const uint32_t *pkt_size, int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v,
                                   int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int32_t *status, int mem) {
    return decode_frames_common(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, mem, mode_sp5x());
}

// ------------------------------------------------------------------------------ decode, plain JPEG
// Host-side walk over the marker segments of the sample frame (find_marker + the segment parsers of
// mjpegdec.c, see include/amvcuda.h).  No pixel arithmetic here: the tables go to the device, the scan to the kernels.
AMV_API int amv_mjpeg_configure(amv_ctx *ctx, const uint8_t *p, uint32_t size, int *w_out, int *h_out) {
    if (!ctx) return AMV_ERR_ARG;
    if (!p || size < 4 || p[0] != 0xff || p[1] != 0xd8) return fail(ctx, AMV_ERR_ARG, "not a JPEG (no SOI)");
    uint8_t q[4][64], hc[2][4][16], hs[2][4][256];
    bool have_q[4] = { false, false, false, false }, have_h[2][4] = { { false } };
    int comp_id[3] = { 0, 0, 0 }, comp_q[3] = { 0, 0, 0 }, w = 0, h = 0, samp[4] = { 1, 1, 0, 0 }, restart = 0;
    bool have_sof = false;
    uint32_t i = 2, scan_start = 0, qofs[4] = { 0, 0, 0, 0 };
    int td[3] = { 0, 0, 0 }, ta[3] = { 0, 0, 0 };
    auto bad = [&](const char *what) { return fail(ctx, AMV_ERR_UNSUPPORTED, what); };
    while (i + 4 <= size && !scan_start) {
        if (p[i] != 0xff) { i++; continue; }
        const int m = p[i + 1];
        if (m < 0xc0 || m == 0xff) { i++; continue; }                       // find_marker: FF followed by C0..FE
        if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7)) { i += 2; continue; }
        if (m == 0xd9) return bad("EOI before the scan");
        const uint32_t len = ((uint32_t)p[i + 2] << 8) | p[i + 3];
        if (len < 2 || i + 2 + len > size) return bad("marker segment runs past the frame");
        const uint8_t *d = p + i + 4;
        uint32_t n = len - 2;
        if (m == 0xdb) {                                                     // DQT
            while (n >= 65) {
                if (d[0] >> 4) return bad("16-bit quantiser table");
                const int id = d[0] & 15;
                if (id >= 4) return bad("quantiser table index");
                memcpy(q[id], d + 1, 64); have_q[id] = true;
                qofs[id] = (uint32_t)(d + 1 - p);
                d += 65; n -= 65;
            }
        } else if (m == 0xc4) {                                              // DHT
            while (n > 0) {
                if (n < 17) return bad("short DHT");
                const int cls = d[0] >> 4, id = d[0] & 15;
                uint32_t tot = 0;
                if (cls >= 2 || id >= 4) return bad("Huffman table class / index");
                for (int k = 0; k < 16; k++) tot += d[1 + k];
                if (tot > 256 || n < 17 + tot) return bad("short DHT");
                memcpy(hc[cls][id], d + 1, 16); memset(hs[cls][id], 0, 256); memcpy(hs[cls][id], d + 17, tot);
                have_h[cls][id] = true;
                d += 17 + tot; n -= 17 + tot;
            }
        } else if (m == 0xc0) {                                              // SOF0
            if (n < 15 || d[0] != 8 || d[5] != 3) return bad("not 8-bit, three components");
            h = (d[1] << 8) | d[2]; w = (d[3] << 8) | d[4];
            for (int c = 0; c < 3; c++) {
                comp_id[c] = d[6 + 3 * c];
                comp_q[c] = d[8 + 3 * c];
                if (comp_q[c] >= 4) return bad("quantiser table index");
            }
            // pix_fmt_id of ff_mjpeg_decode_sof (:283-311): 4:2:0, 4:2:2 in both layouts (the reference's own
            // encoder writes 2x2 / 1x2 / 1x2), 4:4:4
            switch ((d[7] << 16) | (d[10] << 8) | d[13]) {
                case 0x221111: samp[0] = 1; samp[1] = 1; samp[2] = 0; samp[3] = 0; break;
                case 0x211111: samp[0] = 1; samp[1] = 0; samp[2] = 0; samp[3] = 0; break;
                case 0x221212: samp[0] = 1; samp[1] = 1; samp[2] = 0; samp[3] = 1; break;
                case 0x111111: samp[0] = 0; samp[1] = 0; samp[2] = 0; samp[3] = 0; break;
                default: return bad("sampling other than 4:2:0, 4:2:2, 4:4:4");
            }
            if (comp_q[1] != comp_q[2]) return bad("Cb and Cr use different quantiser tables");
            have_sof = true;
        } else if (m >= 0xc1 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
            return bad("not a baseline (SOF0) frame");
        } else if (m == 0xdd) {                                              // DRI (mjpeg_decode_dri :858-867)
            if (len != 4) return bad("DRI segment");
            restart = (d[0] << 8) | d[1];
        } else if (m == 0xda) {                                              // SOS
            if (!have_sof || n < 10 || d[0] != 3 || len != 6 + 2 * 3) return bad("scan header");
            for (int c = 0; c < 3; c++) {
                if (d[1 + 2 * c] != comp_id[c]) return bad("scan components out of frame order");
                td[c] = d[2 + 2 * c] >> 4; ta[c] = d[2 + 2 * c] & 15;
                if (td[c] >= 4 || ta[c] >= 4 || !have_h[0][td[c]] || !have_h[1][ta[c]]) return bad("missing Huffman table");
            }
            if (td[1] != td[2] || ta[1] != ta[2]) return bad("Cb and Cr use different Huffman tables");
            if (d[7] != 0 || d[8] != 63 || d[9] != 0) return bad("not a sequential scan");
            if (!have_q[comp_q[0]] || !have_q[comp_q[1]]) return bad("missing quantiser table");
            scan_start = i + 2 + len;
        }
        i += 2 + len;
    }
    if (!scan_start) return bad("no scan");
    if (w <= 0 || h <= 0 || w > 16384 || h > 16384) return bad("picture size");
    uint8_t counts[4][16], syms[4][256], qzz[2][64];
    for (int c = 0; c < 2; c++) {
        memcpy(counts[c], hc[0][td[c]], 16);     memcpy(syms[c], hs[0][td[c]], 256);
        memcpy(counts[2 + c], hc[1][ta[c]], 16); memcpy(syms[2 + c], hs[1][ta[c]], 256);
        memcpy(qzz[c], q[comp_q[c]], 64);
    }
    std::vector<uint8_t> host(dec_table_set_bytes());
    bool sync_ok = false;
    if (!build_dec_table_set(host.data(), counts, syms, qzz, &sync_ok)) return bad("Huffman codes do not fit the lookup tables");
    ON_DEVICE(ctx);
    CK(cudaStreamSynchronize(ctx->stream));                // nothing in flight may still use the previous configuration
    if (!ctx->mj_tables) CK(cudaMalloc(&ctx->mj_tables, dec_table_set_bytes()));
    if (ctx->mj_hdr) { cudaFree(ctx->mj_hdr); ctx->mj_hdr = nullptr; }
    CK(cudaMalloc(reinterpret_cast<void **>(&ctx->mj_hdr), scan_start));
    CK(cudaMemcpy(ctx->mj_tables, host.data(), host.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->mj_hdr, p, scan_start, cudaMemcpyHostToDevice));
    ctx->mj_hdr_len = scan_start; ctx->mj_w = w; ctx->mj_h = h; ctx->mj_sync_ok = sync_ok;
    ctx->mj_qpos[0] = qofs[comp_q[0]]; ctx->mj_qpos[1] = qofs[comp_q[1]];
    for (int k = 0; k < 4; k++) ctx->mj_samp[k] = samp[k];
    // the reference honours a restart interval only below 1350 MCUs (mjpegdec.c:726 "buggy workaround"): a larger one
    // leaves the markers in the scan as data, there as here
    ctx->mj_restart = restart < 1350 ? restart : 0;
    if (w_out) *w_out = w;
    if (h_out) *h_out = h;
    return AMV_OK;
}

AMV_API int amv_decode_frames_mjpeg(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                                    const uint32_t *pkt_size, int n, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v,
                                    int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int32_t *status, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (!ctx->mj_tables || !ctx->mj_hdr) return fail(ctx, AMV_ERR_ARG, "amv_mjpeg_configure has not been called");
    if (w != ctx->mj_w || h != ctx->mj_h) return fail(ctx, AMV_ERR_ARG, "dimensions differ from the configured header");
    DecMode m;
    m.head = ctx->mj_hdr_len; m.flip = false;
    m.tables = static_cast<const DecTableSet *>(ctx->mj_tables);
    m.hdr = ctx->mj_hdr; m.hdr_len = ctx->mj_hdr_len;
    m.qpos[0] = ctx->mj_qpos[0]; m.qpos[1] = ctx->mj_qpos[1];
    for (int k = 0; k < 4; k++) m.samp[k] = ctx->mj_samp[k];
    m.restart = ctx->mj_restart;
    m.allow_sync = ctx->mj_sync_ok && !ctx->mj_restart;      // restart intervals: one lane per frame (the lane hand-over carries no restart state)
    return decode_frames_common(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, mem, m);
}

// -------------------------------------------------------------------------- decode, amvlib flavour
AMV_API int amv_decode_frames_bgr24(amv_ctx *ctx, const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off,
                                    const uint32_t *pkt_size, int n, int w, int h, uint8_t *bgr, int line_bytes,
                                    uint64_t frame_stride, int32_t *status, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (n < 0 || w <= 0 || h <= 0 || w > 16384 || h > 16384 || bad_mem(mem)) return fail(ctx, AMV_ERR_ARG, "bad n / dimensions / mem");
    if (n == 0) return AMV_OK;
    if (!pkts || !pkt_off || !pkt_size || !bgr) return fail(ctx, AMV_ERR_ARG, "null buffer");
    if (line_bytes < 3 * w || frame_stride < (uint64_t)line_bytes * (h - 1) + 3 * (uint64_t)w)
        return fail(ctx, AMV_ERR_ARG, "strides smaller than the bitmap");
    ON_DEVICE(ctx);
    if (mem == AMV_MEM_DEVICE)
        return decode_bgr_device(ctx, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, bgr, line_bytes, frame_stride, status, pkts_bytes);
    // host buffers: staged copy in, kernels, copy out (bytes of the bitmaps that no pixel covers -- row
    // padding -- are copied back as they came in, like the reference leaves them untouched)
    uint8_t *d_pk, *d_bgr; uint64_t *d_off; uint32_t *d_sz; int32_t *d_st;
    const uint64_t bgr_bytes = frame_stride * (uint64_t)(n - 1) + (uint64_t)line_bytes * (h - 1) + 3 * (uint64_t)w;
    TO_DEVICE(WS_H_A, pkts, pkts_bytes, d_pk);
    TO_DEVICE(WS_H_B, pkt_off, sizeof(uint64_t) * n, d_off);
    TO_DEVICE(WS_H_C, pkt_size, sizeof(uint32_t) * n, d_sz);
    TO_DEVICE(WS_H_D, bgr, bgr_bytes, d_bgr);
    ENSURE(WS_H_E, sizeof(int32_t) * n, d_st);
    int r = decode_bgr_device(ctx, d_pk, pkts_bytes, d_off, d_sz, n, w, h, d_bgr, line_bytes, frame_stride, d_st, pkts_bytes);
    if (r != AMV_OK) return r;
    CK(cudaMemcpyAsync(bgr, d_bgr, bgr_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CK(cudaMemcpyAsync(status, d_st, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return AMV_OK;
}

// ------------------------------------------------------------------------------- range conversion
AMV_API int amv_convert_range(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c,
                              uint64_t fs_y, uint64_t fs_c, int n, int w, int h, int dir, uint8_t *oy, uint8_t *ou,
                              uint8_t *ov, int ols_y, int ols_c, uint64_t ofs_y, uint64_t ofs_c, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (n < 0 || w <= 0 || h <= 0 || w > 16384 || h > 16384 || bad_mem(mem) || (dir != 0 && dir != 1))
        return fail(ctx, AMV_ERR_ARG, "bad n / dimensions / mem / dir");
    if (n == 0) return AMV_OK;
    if (!y || !u || !v || !oy || !ou || !ov) return fail(ctx, AMV_ERR_ARG, "null buffer");
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    if (ls_y < w || ls_c < cw || ols_y < w || ols_c < cw || fs_y < (uint64_t)ls_y * (h - 1) + w || fs_c < (uint64_t)ls_c * (ch - 1) + cw ||
        ofs_y < (uint64_t)ols_y * (h - 1) + w || ofs_c < (uint64_t)ols_c * (ch - 1) + cw)
        return fail(ctx, AMV_ERR_ARG, "strides smaller than the picture");
    ON_DEVICE(ctx);
    if (mem == AMV_MEM_DEVICE) {
        launch_convert_range(y, u, v, oy, ou, ov, n, w, h, ls_y, ls_c, fs_y, fs_c, ols_y, ols_c, ofs_y, ofs_c, dir, ctx->stream);
        return check_launch(ctx, "range conversion kernels", 3);
    }
    // host buffers: tight device copies in, kernels, tight copies out
    const uint64_t ty = (uint64_t)w * h, tc = (uint64_t)cw * ch;
    uint8_t *d_y, *d_u, *d_v;
    ENSURE(WS_H_A, ty * n, d_y);
    ENSURE(WS_H_B, tc * n, d_u);
    ENSURE(WS_H_C, tc * n, d_v);
    int r;
    if ((r = copy_planes(ctx, d_y, const_cast<uint8_t *>(y), w, h, ls_y, fs_y, n, false)) != AMV_OK) return r;
    if ((r = copy_planes(ctx, d_u, const_cast<uint8_t *>(u), cw, ch, ls_c, fs_c, n, false)) != AMV_OK) return r;
    if ((r = copy_planes(ctx, d_v, const_cast<uint8_t *>(v), cw, ch, ls_c, fs_c, n, false)) != AMV_OK) return r;
    launch_convert_range(d_y, d_u, d_v, d_y, d_u, d_v, n, w, h, w, cw, ty, tc, w, cw, ty, tc, dir, ctx->stream);
    if ((r = check_launch(ctx, "range conversion kernels", 3)) != AMV_OK) return r;
    if ((r = copy_planes(ctx, d_y, oy, w, h, ols_y, ofs_y, n, true)) != AMV_OK) return r;
    if ((r = copy_planes(ctx, d_u, ou, cw, ch, ols_c, ofs_c, n, true)) != AMV_OK) return r;
    if ((r = copy_planes(ctx, d_v, ov, cw, ch, ols_c, ofs_c, n, true)) != AMV_OK) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return AMV_OK;
}

// ------------------------------------------------------------------------------- picture scaler
AMV_API int amv_scale_frames_ex(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c,
                                uint64_t fs_y, uint64_t fs_c, int n, int iw, int ih, uint8_t *oy, uint8_t *ou, uint8_t *ov,
                                int ols_y, int ols_c, uint64_t ofs_y, uint64_t ofs_c, int ow, int oh, int flags, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (n < 0 || iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0 || iw > 16384 || ih > 16384 || ow > 16384 || oh > 16384 || bad_mem(mem) ||
        (flags & ~(AMV_SCALE_IN_JPEG_RANGE | AMV_SCALE_OUT_JPEG_RANGE)))
        return fail(ctx, AMV_ERR_ARG, "bad n / dimensions / flags / mem");
    if (n == 0) return AMV_OK;
    if (!y || !u || !v || !oy || !ou || !ov) return fail(ctx, AMV_ERR_ARG, "null buffer");
    // the reference scales planes 1 and 2 at (w >> 1) x (h >> 1) (imgresample.c:494-505)
    const int icw = iw >> 1, ich = ih >> 1, ocw = ow >> 1, och = oh >> 1;
    if (ocw > 0 && och > 0 && (icw == 0 || ich == 0))
        return fail(ctx, AMV_ERR_UNSUPPORTED, "a 1-pixel-wide or -high source has no chroma to scale (the reference reads outside the plane)");
    if (ls_y < iw || ls_c < icw || ols_y < ow || ols_c < ocw || fs_y < (uint64_t)ls_y * (ih - 1) + iw ||
        (ich > 0 && fs_c < (uint64_t)ls_c * (ich - 1) + icw) || ofs_y < (uint64_t)ols_y * (oh - 1) + ow ||
        (och > 0 && ofs_c < (uint64_t)ols_c * (och - 1) + ocw))
        return fail(ctx, AMV_ERR_ARG, "strides smaller than the picture");
    ScaleBanks banks;
    build_scale_banks(iw, ih, ow, oh, &banks);
    ON_DEVICE(ctx);
    const bool chroma = icw > 0 && ich > 0 && ocw > 0 && och > 0;
    const bool host = mem == AMV_MEM_HOST, pre = (flags & AMV_SCALE_IN_JPEG_RANGE) != 0, post = (flags & AMV_SCALE_OUT_JPEG_RANGE) != 0;
    const uint64_t ty = (uint64_t)iw * ih, tc = (uint64_t)icw * ich, oty = (uint64_t)ow * oh, otc = (uint64_t)ocw * och;
    // what the scaler reads and writes: the caller's device planes, or tight device copies (host buffers; a
    // range-converted input never overwrites the caller's planes)
    const uint8_t *sy = y, *su = u, *sv = v;
    int sls_y = ls_y, sls_c = ls_c;
    uint64_t sfs_y = fs_y, sfs_c = fs_c;
    uint8_t *dy = oy, *du = ou, *dv = ov;
    int dls_y = ols_y, dls_c = ols_c;
    uint64_t dfs_y = ofs_y, dfs_c = ofs_c;
    int launches = 0, r;
    if (host || pre) {
        uint8_t *t_y, *t_u, *t_v;
        ENSURE(WS_H_A, ty * n, t_y);
        ENSURE(WS_H_B, tc * n + 1, t_u);
        ENSURE(WS_H_C, tc * n + 1, t_v);
        if (host) {
            if ((r = copy_planes(ctx, t_y, const_cast<uint8_t *>(y), iw, ih, ls_y, fs_y, n, false)) != AMV_OK) return r;
            if (chroma) {
                if ((r = copy_planes(ctx, t_u, const_cast<uint8_t *>(u), icw, ich, ls_c, fs_c, n, false)) != AMV_OK) return r;
                if ((r = copy_planes(ctx, t_v, const_cast<uint8_t *>(v), icw, ich, ls_c, fs_c, n, false)) != AMV_OK) return r;
            }
            sy = t_y; su = t_u; sv = t_v; sls_y = iw; sls_c = icw; sfs_y = ty; sfs_c = tc;
        }
        if (pre) {      // img_convert YUVJ420P -> YUV420P in front of the scaler (imgresample.c:617-636)
            launch_convert_range_plane(sy, t_y, iw, ih, n, sls_y, iw, sfs_y, ty, 1, false, ctx->stream);
            launches++;
            if (chroma) {
                launch_convert_range_plane(su, t_u, icw, ich, n, sls_c, icw, sfs_c, tc, 1, true, ctx->stream);
                launch_convert_range_plane(sv, t_v, icw, ich, n, sls_c, icw, sfs_c, tc, 1, true, ctx->stream);
                launches += 2;
            }
            sy = t_y; su = t_u; sv = t_v; sls_y = iw; sls_c = icw; sfs_y = ty; sfs_c = tc;
        }
    }
    if (host) {
        ENSURE(WS_H_D, oty * n, dy);
        ENSURE(WS_H_E, otc * n + 4, du);
        ENSURE(WS_H_F, otc * n + 4, dv);
        dls_y = ow; dls_c = ocw; dfs_y = oty; dfs_c = otc;
    }
    launches += launch_scale_frames(sy, su, sv, sls_y, sls_c, sfs_y, sfs_c, n, iw, ih, dy, du, dv, dls_y, dls_c, dfs_y, dfs_c, ow, oh,
                                    banks, ctx->opt_scale_form, ctx->stream);
    if (post) {         // img_convert YUV420P -> YUVJ420P behind it (:671-682), over the area the scaler wrote
        launch_convert_range_plane(dy, dy, ow, oh, n, dls_y, dls_y, dfs_y, dfs_y, 0, false, ctx->stream);
        launches++;
        if (chroma) {
            launch_convert_range_plane(du, du, ocw, och, n, dls_c, dls_c, dfs_c, dfs_c, 0, true, ctx->stream);
            launch_convert_range_plane(dv, dv, ocw, och, n, dls_c, dls_c, dfs_c, dfs_c, 0, true, ctx->stream);
            launches += 2;
        }
    }
    if ((r = check_launch(ctx, "scaler kernels", launches)) != AMV_OK) return r;
    if (!host) return AMV_OK;
    if ((r = copy_planes(ctx, dy, oy, ow, oh, ols_y, ofs_y, n, true)) != AMV_OK) return r;
    if (chroma) {
        if ((r = copy_planes(ctx, du, ou, ocw, och, ols_c, ofs_c, n, true)) != AMV_OK) return r;
        if ((r = copy_planes(ctx, dv, ov, ocw, och, ols_c, ofs_c, n, true)) != AMV_OK) return r;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return AMV_OK;
}

AMV_API int amv_scale_frames(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c,
                             uint64_t fs_y, uint64_t fs_c, int n, int iw, int ih, uint8_t *oy, uint8_t *ou, uint8_t *ov,
                             int ols_y, int ols_c, uint64_t ofs_y, uint64_t ofs_c, int ow, int oh, int mem) {
    return amv_scale_frames_ex(ctx, y, u, v, ls_y, ls_c, fs_y, fs_c, n, iw, ih, oy, ou, ov, ols_y, ols_c, ofs_y, ofs_c, ow, oh, 0, mem);
}

// the filter banks both stages run on, as the host builds them (init-time work, no device involved)
AMV_API int amv_scale_banks(int iw, int ih, int ow, int oh, int16_t *h_bank, int16_t *v_bank, int32_t *h_incr, int32_t *v_incr) {
    if (iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0 || iw > 16384 || ih > 16384 || ow > 16384 || oh > 16384 || !h_bank || !v_bank)
        return AMV_ERR_ARG;
    ScaleBanks b;
    build_scale_banks(iw, ih, ow, oh, &b);
    memcpy(h_bank, b.h, sizeof(b.h));
    memcpy(v_bank, b.v, sizeof(b.v));
    if (h_incr) *h_incr = b.h_incr;
    if (v_incr) *v_incr = b.v_incr;
    return AMV_OK;
}
AMV_API int amv_audio_resample_bank(int in_rate, int out_rate, int16_t *bank, uint64_t bank_cap) {
    if (in_rate <= 0 || out_rate <= 0 || in_rate > (1 << 21) || out_rate > (1 << 21)) return AMV_ERR_ARG;
    const int len = resample_filter_length(in_rate, out_rate);
    if (len > 4096) return AMV_ERR_UNSUPPORTED;
    if (bank) {
        if (bank_cap < (uint64_t)len * 1024) return AMV_ERR_ARG;
        build_resample_bank(in_rate, out_rate, bank);
    }
    return len;
}

// ------------------------------------------------------------------------------- audio resampler
AMV_API uint64_t amv_audio_resample_count(uint64_t n_in, int in_rate, int out_rate) {
    if (in_rate <= 0 || out_rate <= 0 || n_in == 0 || n_in > (1ull << 40)) return 0;
    return (uint64_t)resample_output_count((int64_t)n_in, in_rate, out_rate);
}

AMV_API int amv_audio_resample_from(amv_ctx *ctx, const int16_t *in, uint64_t in_base, uint64_t n_in, int in_channels, int in_rate,
                                    int out_rate, uint64_t k_start, int16_t *out, uint64_t out_cap, uint64_t *n_out, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (bad_mem(mem) || (in_channels != 1 && in_channels != 2) || in_rate <= 0 || out_rate <= 0 || in_rate > (1 << 21) ||
        out_rate > (1 << 21) || n_in > (1ull << 40) || in_base > (1ull << 40) || k_start > (1ull << 40))
        return fail(ctx, AMV_ERR_ARG, "bad mem / channels / rates / length");
    if (!n_out) return fail(ctx, AMV_ERR_ARG, "null n_out");
    *n_out = 0;
    if (n_in == 0) return AMV_OK;
    if (!in || !out) return fail(ctx, AMV_ERR_ARG, "null buffer");
    const int len = resample_filter_length(in_rate, out_rate);
    if (len > 4096) return fail(ctx, AMV_ERR_UNSUPPORTED, "rate ratio needs more than 4096 filter taps");
    const int64_t total = resample_output_count((int64_t)(in_base + n_in), in_rate, out_rate);
    int64_t k = total > (int64_t)k_start ? total - (int64_t)k_start : 0;
    if ((uint64_t)k > out_cap) k = (int64_t)out_cap;          // like av_resample's dst_size: the rest comes with a later call
    if (k > 0) {
        const int64_t first = resample_first_tap((int64_t)k_start, in_rate, out_rate);
        if (first < 0 ? in_base != 0 : first < (int64_t)in_base)
            return fail(ctx, AMV_ERR_ARG, "output k_start needs samples in front of in_base");
    }
    if (mem == AMV_MEM_DEVICE && (((uintptr_t)in & (in_channels == 2 ? 3 : 1)) || ((uintptr_t)out & 1)))
        return fail(ctx, AMV_ERR_ARG, "misaligned sample pointer");
    ON_DEVICE(ctx);
    // the polyphase bank of this rate pair (kept until the rates change)
    if (ctx->rs_in_rate != in_rate || ctx->rs_out_rate != out_rate) {
        std::vector<int16_t> rows((size_t)len * 1024);
        build_resample_bank(in_rate, out_rate, rows.data());
        const int len8 = (len + 7) & ~7;             // device rows are zero-padded to 8 coefficients (amv_resample.cu)
        std::vector<int16_t> bank((size_t)len8 * 1024, 0);
        for (int ph = 0; ph < 1024; ph++) memcpy(&bank[(size_t)ph * len8], &rows[(size_t)ph * len], sizeof(int16_t) * len);
        void *p = nullptr;
        int r = ensure(ctx, WS_RS_BANK, bank.size() * sizeof(int16_t), &p);
        if (r != AMV_OK) return r;
        ctx->rs_in_rate = ctx->rs_out_rate = 0;
        CK(cudaMemcpyAsync(p, bank.data(), bank.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->rs_in_rate = in_rate; ctx->rs_out_rate = out_rate;
    }
    const int16_t *d_bank = reinterpret_cast<const int16_t *>(ctx->ws[WS_RS_BANK].p);
    if (mem == AMV_MEM_DEVICE) {
        launch_audio_resample(in, (int64_t)n_in, (int64_t)in_base, in_channels, d_bank, len, in_rate, out_rate, (int64_t)k_start, out, k,
                              ctx->opt_resample_form, ctx->stream);
        *n_out = (uint64_t)k;
        return k > 0 ? check_launch(ctx, "audio resampler kernel", 1) : AMV_OK;
    }
    int16_t *d_in, *d_out;
    TO_DEVICE(WS_H_A, in, sizeof(int16_t) * n_in * in_channels, d_in);
    ENSURE(WS_H_B, sizeof(int16_t) * (k > 0 ? k : 1), d_out);
    if (k > 0) {
        launch_audio_resample(d_in, (int64_t)n_in, (int64_t)in_base, in_channels, d_bank, len, in_rate, out_rate, (int64_t)k_start, d_out,
                              k, ctx->opt_resample_form, ctx->stream);
        int r = check_launch(ctx, "audio resampler kernel", 1);
        if (r != AMV_OK) return r;
        CK(cudaMemcpyAsync(out, d_out, sizeof(int16_t) * k, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    *n_out = (uint64_t)k;
    return AMV_OK;
}

AMV_API int amv_audio_resample(amv_ctx *ctx, const int16_t *in, uint64_t n_in, int in_channels, int in_rate, int out_rate,
                               int16_t *out, uint64_t out_cap, uint64_t *n_out, int mem) {
    if (ctx && in_rate > 0 && out_rate > 0 && n_in <= (1ull << 40) && n_in > 0 &&
        (uint64_t)resample_output_count((int64_t)n_in, in_rate, out_rate) > out_cap) {
        if (n_out) *n_out = 0;
        return fail(ctx, AMV_ERR_ARG, "out_cap smaller than the resampled stream (see amv_audio_resample_count)");
    }
    return amv_audio_resample_from(ctx, in, 0, n_in, in_channels, in_rate, out_rate, 0, out, out_cap, n_out, mem);
}

AMV_API int64_t amv_audio_resample_first_tap(uint64_t k, int in_rate, int out_rate) {
    if (in_rate <= 0 || out_rate <= 0 || k > (1ull << 40)) return 0;
    return resample_first_tap((int64_t)k, in_rate, out_rate);
}

// ---------------------------------------------------------------------------------------- encode
AMV_API int amv_encode_frames(amv_ctx *ctx, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c,
                              uint64_t fs_y, uint64_t fs_c, int n, int w, int h, const int32_t *qscale, uint8_t *out,
                              uint64_t out_cap, uint32_t pkt_cap, int layout, uint64_t *out_off, uint32_t *out_size,
                              int32_t *status, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (n < 0 || w <= 0 || h <= 0 || w > 16384 || h > 16384 || bad_mem(mem)) return fail(ctx, AMV_ERR_ARG, "bad n / dimensions / mem");
    if (layout != AMV_LAYOUT_PACKED && layout != AMV_LAYOUT_SLOTS) return fail(ctx, AMV_ERR_ARG, "bad layout");
    if (!encode_geometry_ok(w, h))
        return fail(ctx, AMV_ERR_UNSUPPORTED, "height outside the reference encoder's defined domain ((h/2)%8 must be 0 or 4)");
    if (n == 0) return AMV_OK;
    if (!y || !u || !v || !out || !out_off || !out_size) return fail(ctx, AMV_ERR_ARG, "null buffer");
    if (pkt_cap < 4) return fail(ctx, AMV_ERR_ARG, "pkt_cap too small");
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    if (ls_y < w || ls_c < cw || fs_y < (uint64_t)ls_y * (h - 1) + w || fs_c < (uint64_t)ls_c * (ch - 1) + cw)
        return fail(ctx, AMV_ERR_ARG, "strides smaller than the picture");
    ON_DEVICE(ctx);
    if (mem == AMV_MEM_DEVICE)
        return encode_device(ctx, y, u, v, ls_y, ls_c, fs_y, fs_c, n, w, h, qscale, out, out_cap, pkt_cap, layout, out_off,
                             out_size, status);

    if (qscale) for (int i = 0; i < n; i++) if (qscale[i] < 2 || qscale[i] > 31) return fail(ctx, AMV_ERR_UNSUPPORTED, "qscale outside 2..31");
    return encode_host(ctx, y, u, v, ls_y, ls_c, fs_y, fs_c, n, w, h, qscale, out, out_cap, pkt_cap, layout, out_off, out_size, status);
}

// ----------------------------------------------------------------------------------------- adpcm
AMV_API int amv_adpcm_dec_chunks(amv_ctx *ctx, const uint8_t *chunks, uint64_t chunks_bytes, const uint64_t *chunk_off,
                                 const uint32_t *chunk_size, int n, int16_t *pcm, uint64_t pcm_samples,
                                 const uint64_t *pcm_off, int32_t *status, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (n < 0 || bad_mem(mem)) return fail(ctx, AMV_ERR_ARG, "bad n / mem");
    if (n == 0) return AMV_OK;
    if (!chunks || !chunk_off || !chunk_size || !pcm || !pcm_off) return fail(ctx, AMV_ERR_ARG, "null buffer");
    ON_DEVICE(ctx);
    if (mem == AMV_MEM_DEVICE) {
        int32_t *st = status;
        if (!st) ENSURE(WS_STATUS, sizeof(int32_t) * n, st);
        { ScopedTimer tm(ctx, KK_ADPCM_DEC);
          launch_adpcm_decode(chunks, chunks_bytes, chunk_off, chunk_size, n, pcm, pcm_samples, pcm_off, st, ctx->opt_adpcm_form, ctx->stream); }
        return check_launch(ctx, "adpcm decode kernel");
    }
    uint8_t *d_c; uint64_t *d_off, *d_poff; uint32_t *d_sz; int16_t *d_pcm; int32_t *d_st;
    TO_DEVICE(WS_H_A, chunks, chunks_bytes, d_c);
    TO_DEVICE(WS_H_B, chunk_off, sizeof(uint64_t) * n, d_off);
    TO_DEVICE(WS_H_C, chunk_size, sizeof(uint32_t) * n, d_sz);
    TO_DEVICE(WS_H_D, pcm_off, sizeof(uint64_t) * n, d_poff);
    ENSURE(WS_H_E, sizeof(int16_t) * pcm_samples, d_pcm);
    ENSURE(WS_H_F, sizeof(int32_t) * n, d_st);
    launch_adpcm_decode(d_c, chunks_bytes, d_off, d_sz, n, d_pcm, pcm_samples, d_poff, d_st, ctx->opt_adpcm_form, ctx->stream);
    int r = check_launch(ctx, "adpcm decode kernel");
    if (r != AMV_OK) return r;
    // Only what the kernel wrote goes home: samples no chunk covers, and the regions of rejected chunks, stay as the
    // caller had them (like the device-memory path leaves them).  Neighbouring chunks travel as one copy.
    std::vector<int32_t> hst((size_t)n);
    CK(cudaMemcpyAsync(hst.data(), d_st, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i <= n; i++) {
        const bool wrote = i < n && hst[i] == 0 && chunk_size[i] > 8;
        const uint64_t a = wrote ? pcm_off[i] : 0, len = wrote ? 2ull * (chunk_size[i] - 8) : 0;
        if (wrote && hi > lo && a == hi) { hi += len; continue; }
        if (hi > lo) CK(cudaMemcpyAsync(pcm + lo, d_pcm + lo, sizeof(int16_t) * (hi - lo), cudaMemcpyDeviceToHost, ctx->stream));
        lo = a; hi = a + len;
    }
    if (status) memcpy(status, hst.data(), sizeof(int32_t) * n);
    CK(cudaStreamSynchronize(ctx->stream));
    return AMV_OK;
}

static int adpcm_encode_common(amv_ctx *ctx, const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off,
                               const uint32_t *nsamples, const uint32_t *first_chunk, int nstreams, int nchunks,
                               const int16_t *step_in, int16_t *step_out, uint8_t *out, uint64_t out_bytes,
                               const uint64_t *out_off, int32_t *status, int mem) {
    if (!ctx) return AMV_ERR_ARG;
    if (nstreams < 0 || nchunks < 0 || bad_mem(mem)) return fail(ctx, AMV_ERR_ARG, "bad n / mem");
    if (nstreams == 0 || nchunks == 0) return AMV_OK;
    if (!pcm || !pcm_off || !nsamples || !out || !out_off) return fail(ctx, AMV_ERR_ARG, "null buffer");
    ON_DEVICE(ctx);
    if (mem == AMV_MEM_DEVICE) {
        int32_t *st = status;
        if (!st) ENSURE(WS_STATUS, sizeof(int32_t) * nchunks, st);
        { ScopedTimer tm(ctx, KK_ADPCM_ENC);
          launch_adpcm_encode(pcm, pcm_samples, pcm_off, nsamples, first_chunk, nstreams, nchunks, step_in, step_out, out, out_bytes,
                              out_off, st, ctx->opt_trellis, ctx->opt_adpcm_form, ctx->stream); }
        return check_launch(ctx, "adpcm encode kernel");
    }
    if (first_chunk) {      // the stream table is on the host here: it must be monotonic and end inside the chunk arrays
        for (int s_ = 0; s_ < nstreams; s_++)
            if (first_chunk[s_] > first_chunk[s_ + 1]) return fail(ctx, AMV_ERR_ARG, "first_chunk is not monotonic");
        if (first_chunk[nstreams] > (uint32_t)nchunks) return fail(ctx, AMV_ERR_ARG, "first_chunk runs past nchunks");
    }
    int16_t *d_pcm; uint64_t *d_poff, *d_ooff; uint32_t *d_ns, *d_fc = nullptr; int16_t *d_si = nullptr, *d_so; uint8_t *d_out;
    int32_t *d_st;
    TO_DEVICE(WS_H_A, pcm, sizeof(int16_t) * pcm_samples, d_pcm);
    TO_DEVICE(WS_H_B, pcm_off, sizeof(uint64_t) * nchunks, d_poff);
    TO_DEVICE(WS_H_C, nsamples, sizeof(uint32_t) * nchunks, d_ns);
    TO_DEVICE(WS_H_D, out_off, sizeof(uint64_t) * nchunks, d_ooff);
    if (first_chunk) TO_DEVICE(WS_H_E, first_chunk, sizeof(uint32_t) * (nstreams + 1), d_fc);
    if (step_in) TO_DEVICE(WS_H_F, step_in, sizeof(int16_t) * nstreams, d_si);
    ENSURE(WS_H_G, sizeof(int16_t) * nstreams, d_so);
    ENSURE(WS_H_H, out_bytes ? out_bytes : 1, d_out);
    ENSURE(WS_H_I, sizeof(int32_t) * nchunks, d_st);
    CK(cudaMemsetAsync(d_st, 0, sizeof(int32_t) * nchunks, ctx->stream));
    launch_adpcm_encode(d_pcm, pcm_samples, d_poff, d_ns, d_fc, nstreams, nchunks, d_si, d_so, d_out, out_bytes, d_ooff, d_st,
                        ctx->opt_trellis, ctx->opt_adpcm_form, ctx->stream);
    int r = check_launch(ctx, "adpcm encode kernel");
    if (r != AMV_OK) return r;
    // as in amv_adpcm_dec_chunks: only the chunks that were written travel back, neighbours as one copy
    std::vector<int32_t> hst((size_t)nchunks);
    CK(cudaMemcpyAsync(hst.data(), d_st, sizeof(int32_t) * nchunks, cudaMemcpyDeviceToHost, ctx->stream));
    if (step_out) CK(cudaMemcpyAsync(step_out, d_so, sizeof(int16_t) * nstreams, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint64_t lo = 0, hi = 0;
    for (int c = 0; c <= nchunks; c++) {
        const bool wrote = c < nchunks && hst[c] == 0;
        const uint64_t a = wrote ? out_off[c] : 0, len = wrote ? 8ull + nsamples[c] / 2 : 0;
        if (wrote && hi > lo && a == hi) { hi += len; continue; }
        if (hi > lo) CK(cudaMemcpyAsync(out + lo, d_out + lo, hi - lo, cudaMemcpyDeviceToHost, ctx->stream));
        lo = a; hi = a + len;
    }
    if (status) memcpy(status, hst.data(), sizeof(int32_t) * nchunks);
    CK(cudaStreamSynchronize(ctx->stream));
    return AMV_OK;
}

AMV_API int amv_adpcm_enc_chunks(amv_ctx *ctx, const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off,
                                 const uint32_t *nsamples, const int16_t *step_in, int16_t *step_out, int n, uint8_t *out,
                                 uint64_t out_bytes, const uint64_t *out_off, int32_t *status, int mem) {
    return adpcm_encode_common(ctx, pcm, pcm_samples, pcm_off, nsamples, nullptr, n, n, step_in, step_out, out, out_bytes,
                               out_off, status, mem);
}

AMV_API int amv_adpcm_enc_streams(amv_ctx *ctx, const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off,
                                  const uint32_t *nsamples, const uint32_t *first_chunk, int nstreams, int nchunks,
                                  const int16_t *step_in, int16_t *step_out, uint8_t *out, uint64_t out_bytes,
                                  const uint64_t *out_off, int32_t *status, int mem) {
    if (!first_chunk) return fail(ctx, AMV_ERR_ARG, "first_chunk is NULL");
    return adpcm_encode_common(ctx, pcm, pcm_samples, pcm_off, nsamples, first_chunk, nstreams, nchunks, step_in, step_out, out,
                               out_bytes, out_off, status, mem);
}

}  // extern "C"

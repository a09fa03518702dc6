// amv_kernels.h -- internal launcher interface between the kernels (*.cu) and the C ABI (amv_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "amv_common.cuh"

namespace amv {

// entry state of one decode lane, produced by k_vlc_sync and consumed by k_decode
struct LaneStart {
    uint32_t bitpos;        // where the lane's first block starts in the un-stuffed scan
    uint32_t first_block;   // index of that block in bitstream order
    uint32_t nblocks;       // blocks the lane owns
    int      pred[3];       // DC predictors (Y, Cb, Cr) before that block
};

// ---- decode
cudaError_t upload_dec_tables(cudaStream_t s);
// The tables a scan is decoded with live in device memory as one DecTableSet (amv_dec.cu): the two fixed
// sets (AMV / SP5X, amvlib flavour), or one built on the host from a JPEG's DQT / DHT content
// (build_dec_table_set: false if the codes are malformed or do not fit the lookup tables; *sync_ok says whether
// the lane-synchronisation kernel's smaller table could be built too) and copied to the device by the caller.
struct DecTableSet;
const DecTableSet *fixed_dec_tables(bool amvlib);
size_t dec_table_set_bytes();
bool build_dec_table_set(void *host_buf, const uint8_t counts[4][16], const uint8_t syms[4][256], const uint8_t qzz[2][64],
                         bool *sync_ok);
void launch_scan_sizes(const uint32_t *size, int n, uint32_t align_mask, uint32_t pad, uint64_t *off,
                       uint64_t *carry_io, cudaStream_t s);
// head: bytes in front of the scan (AMV 2, SP5X 14, plain JPEG: everything up to the end of the SOS header);
// literal: SP5X's un-stuffed scan that runs to the end of the packet
void launch_unstuff(const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                    uint8_t *scratch, const uint64_t *slot_off, uint64_t scratch_bytes, uint32_t *scan_len,
                    int32_t *status, uint32_t head, bool literal, cudaStream_t s);
// plain JPEG: compares each frame's marker segments with the configured header outside the two quantiser fields
// (at qpos0 / qpos1) and copies those 2 x 64 quantisers to qtab[f]
void launch_mjpeg_check(const uint8_t *pkts, uint64_t pkts_bytes, const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                        const uint8_t *hdr, uint32_t hdr_len, uint32_t qpos0, uint32_t qpos1, uint8_t *qtab,
                        uint32_t *scan_len, int32_t *status, cudaStream_t s);
// qtab != nullptr: per-frame quantisers (plain JPEG) instead of the table set's
void launch_vlc_sync(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, int n, int log2p,
                     LaneStart *starts, uint32_t *rounds_out, bool amvlib, const DecTableSet *tabs, const uint8_t *qtab,
                     int nl, int nc /* blocks per MCU: luma, one chroma component */,
                     bool lean /* fixed tables: the lean walk (k_vlc_sync_lean) */, cudaStream_t s);
void launch_vlc_tokens(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, const uint32_t *pkt_size,
                       int n, int log2p, const LaneStart *starts, int nblk, uint32_t *tokens, uint32_t *blk_off,
                       int32_t *status, bool amvlib, const DecTableSet *tabs, const uint8_t *qtab, int nl, int nc,
                       int restart /* plain JPEG with DRI: MCUs per restart interval (one lane per frame only), else 0 */,
                       cudaStream_t s);
// the fixed-table (AMV / SP5X) pass with 16-bit tokens: (run << 12 | level), the consumer dequantises
void launch_vlc_tokens16(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, const uint32_t *pkt_size,
                         int n, int log2p, const LaneStart *starts, int nblk, uint16_t *tokens, uint32_t *blk_off,
                         int32_t *status, const DecTableSet *tabs, int nl, int nc, cudaStream_t s);
// the same lean pass with k_vlc_tokens' 32-bit (column offset, dequantised value) tokens: the consumer is k_idct
void launch_vlc_tokens_lean(const uint8_t *scratch, const uint64_t *slot_off, const uint32_t *scan_len, const uint32_t *pkt_size,
                            int n, int log2p, const LaneStart *starts, int nblk, uint32_t *tokens, uint32_t *blk_off,
                            int32_t *status, const DecTableSet *tabs, int nl, int nc, cudaStream_t s);
void launch_idct16(const uint16_t *tokens, const uint32_t *blk_off, const uint64_t *slot_off, const uint32_t *scan_len, int n,
                   const Geom &g, const DecTableSet *tabs, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c, uint64_t fs_y,
                   uint64_t fs_c, cudaStream_t s);
void launch_idct(const uint32_t *tokens, const uint32_t *blk_off, const uint64_t *slot_off, const uint32_t *scan_len, int n,
                 const Geom &g, uint8_t *y, uint8_t *u, uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                 cudaStream_t s);

// amvlib flavour (C-AMVDecoder/amvlib/AmvJpeg.c): Chen-Wang IDCT + fixed-point YUV->BGR24, bottom-up rows
void launch_idct_bgr(const uint32_t *tokens, const uint32_t *blk_off, const uint64_t *slot_off, const uint32_t *scan_len,
                     int n, const Geom &g, uint8_t *bgr, int line_bytes, uint64_t frame_stride, cudaStream_t s);

// range conversion yuv420p <-> yuvj420p (imgconvert.c img_apply_table); dir 0: CCIR -> JPEG, 1: JPEG -> CCIR
void launch_convert_range(const uint8_t *y, const uint8_t *u, const uint8_t *v, uint8_t *oy, uint8_t *ou, uint8_t *ov, int n,
                          int w, int h, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int ols_y, int ols_c, uint64_t ofs_y,
                          uint64_t ofs_c, int dir, cudaStream_t s);

void launch_convert_range_plane(const uint8_t *src, uint8_t *dst, int width, int rows, int n, int ls_in, int ls_out, uint64_t fs_in,
                                uint64_t fs_out, int dir, bool chroma, cudaStream_t s);

// picture scaler (imgresample.c img_resample) and audio resampler (resample.c audio_resample -> resample2.c av_resample);
// the banks are built on the host exactly as av_build_filter does and passed as data
struct ScaleBanks { int16_t h[64]; int16_t v[64]; int h_incr, v_incr; };     // 16 phases x 4 taps each; 16.16 source steps
void build_scale_banks(int iw, int ih, int ow, int oh, ScaleBanks *b);
int  launch_scale_frames(const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                         int n, int iw, int ih, uint8_t *oy, uint8_t *ou, uint8_t *ov, int ols_y, int ols_c, uint64_t ofs_y,
                         uint64_t ofs_c, int ow, int oh, const ScaleBanks &b, int form /* 2 staged tiles, 1 tiles, 0 direct */,
                         cudaStream_t s);      // returns the launches made
int  resample_filter_length(int in_rate, int out_rate);
void build_resample_bank(int in_rate, int out_rate, int16_t *bank /* filter_length * 1024 */);
int64_t resample_output_count(int64_t n_in, int in_rate, int out_rate);
int64_t resample_first_tap(int64_t k, int in_rate, int out_rate);
// in holds the stream's samples [in_base, in_base + n_in); outputs k_base .. k_base + n_out - 1 go to out[0 ..]
void launch_audio_resample(const int16_t *in, int64_t n_in, int64_t in_base, int in_ch, const int16_t *bank, int len, int in_rate,
                           int out_rate, int64_t k_base, int16_t *out, int64_t n_out, int form /* 2 phase rows, 1 tiles, 0 direct */,
                           cudaStream_t s);

// ---- encode
cudaError_t upload_enc_tables(cudaStream_t s);
cudaError_t encode_setup_device();      // per-device function attributes (dynamic shared memory opt-in)
cudaError_t decode_setup_device();
int  encode_grid(int n, int per_sm);
void launch_encode(const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                   int n, const Geom &g, const int32_t *qscale, uint8_t *slots, uint64_t slot_stride, uint32_t pkt_cap,
                   uint32_t *out_size, int32_t *status, int32_t *redo /* n flags of scratch, or nullptr: one-kernel path */,
                   int form /* 0 one kernel, 1 k_encode16 + k_encode, 2 k_encode16v2 + k_encode */, cudaStream_t s);
void launch_compact(const uint8_t *slots, uint64_t slot_stride, uint32_t *size /* zeroed where the packet does not fit */, const uint64_t *off, int n,
                    uint8_t *out, uint64_t out_cap, int32_t *status, cudaStream_t s);
void launch_export_meta(const uint64_t *off, const uint32_t *sz, const int32_t *st, uint64_t *hoff, uint32_t *hsz, int32_t *hst,
                        int n, cudaStream_t s);
void launch_slot_offsets(uint64_t *off, int n, uint64_t stride, uint64_t base, cudaStream_t s);

// ---- adpcm
cudaError_t upload_adpcm_tables(cudaStream_t s);
cudaError_t adpcm_setup_device();
void launch_adpcm_decode(const uint8_t *chunks, uint64_t chunks_bytes, const uint64_t *off, const uint32_t *size, int n,
                         int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off, int32_t *status,
                         int form /* input staging: 2 cp.async.bulk + mbarrier, 1 cp.async, 0 cooperative byte loads */, cudaStream_t s);
void launch_adpcm_encode(const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off, const uint32_t *nsamples,
                         const uint32_t *first_chunk, int nstreams, int nchunks, const int16_t *step_in, int16_t *step_out,
                         uint8_t *out, uint64_t out_bytes, const uint64_t *out_off, int32_t *status, int trellis,
                         int form /* 1 (or 2) cp.async input staging, 0 cooperative loads */, cudaStream_t s);

}  // namespace amv

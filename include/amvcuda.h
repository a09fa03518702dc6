/*
 * amvcuda.h -- C ABI of libamvcuda: the AMV intra-frame codec path on B200 (sm_100a).
 *
 * This is the drop-in boundary.  Each entry point replaces the per-frame /
 * per-chunk body of one reference codec callback (paths relative to
 * /root/reference/AMVmuxer/ffmpeg/libavcodec/ unless noted); the reference-side
 * AVCodec and amvlib bindings that call them are in INTEGRATION.md and glue/.
 *
 *   amv_decode_frames     <- AVCodec amv_decoder.decode   = sp5x_decode_frame  sp5xdec.c:33-188,203-212
 *                            (-> ff_mjpeg_decode_frame mjpegdec.c:1106, mjpeg_decode_scan :660,
 *                                decode_block :376, simple_idct_put simple_idct.c:390)
 *   amv_encode_frames     <- AVCodec amv_encoder.encode   = amv_encode_picture mjpegenc.c:454-472,485-494
 *                            (-> MPV_encode_picture mpegvideo_enc.c:1205, encode_mb_internal :1457,
 *                                dct_quantize_c :3647, ff_jpeg_fdct_islow jfdctint.c:261,
 *                                encode_block mjpegenc.c:379, escape_FF :282, trailer :345)
 *   amv_adpcm_dec_chunks  <- AVCodec adpcm_ima_amv_decoder.decode = adpcm_decode_frame adpcm.c:894,1268-1292
 *   amv_adpcm_enc_chunks  <- AVCodec adpcm_ima_amv_encoder.encode = adpcm_encode_frame adpcm.c:445,461-496
 *   (amvlib mirror: AmvVideoDecode / AmvAudioDecode, C-AMVDecoder/amvlib/AMVDec.c:259-340 -- see INTEGRATION.md)
 *
 * Contract
 *  - Plain C, plain pointers and sizes.  No CPU fallback: every call either runs the
 *    CUDA kernels or returns a negative AMV_ERR_*; without a usable device amv_create fails.
 *  - `mem` says where ALL buffer arguments of that call live: AMV_MEM_HOST (the library
 *    stages through pinned memory and copies both ways inside the call, which returns
 *    after the results are in the caller's buffers) or AMV_MEM_DEVICE (pointers are device
 *    pointers on the context's device; the call only enqueues work on the context's
 *    stream and returns -- synchronise with amv_sync or the stream you passed).
 *  - A call processes a BATCH of n independent units (frames / chunks).  The reference
 *    callbacks are the n == 1 case.
 *  - Results are bit-exact with the reference C paths for every input inside the
 *    reference's defined domain; outside it (corrupt streams, values that make the
 *    reference index out of its tables) the per-unit status says so and the output for
 *    that unit is unspecified but memory-safe.
 *  - One context per host thread / CUDA stream; contexts are independent.
 */
#ifndef AMVCUDA_H
#define AMVCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AMV_API __attribute__((visibility("default")))
#else
#define AMV_API
#endif

#define AMVCUDA_VERSION 0x000100

typedef struct amv_ctx amv_ctx;

enum amv_mem { AMV_MEM_HOST = 0, AMV_MEM_DEVICE = 1 };

/* return codes (negative = failure of the whole call) */
enum {
    AMV_OK            =  0,
    AMV_ERR_ARG       = -1,   /* bad argument (NULL, n < 0, zero dims, unsupported geometry) */
    AMV_ERR_NODEVICE  = -2,   /* no CUDA device / wrong architecture: there is no CPU path */
    AMV_ERR_CUDA      = -3,   /* a CUDA runtime call failed; see amv_last_error */
    AMV_ERR_NOMEM     = -4,
    AMV_ERR_UNSUPPORTED = -5  /* option outside the contract */
};

/* per-unit status words written to `status[i]` (0 = ok, otherwise an OR of these) */
enum {
    AMV_ST_SHORT     = 1 << 0,  /* packet/chunk shorter than its framing */
    AMV_ST_BADCODE   = 1 << 1,  /* bit pattern that is no Huffman code (mjpegdec.c:362-366) */
    AMV_ST_COEFIDX   = 1 << 2,  /* run past coefficient 63 ("error count", mjpegdec.c:423-424) */
    AMV_ST_MARKER    = 1 << 3,  /* FF xx marker inside the scan data cut it short (mjpegdec.c:1153-1157) */
    AMV_ST_OVERRUN   = 1 << 4,  /* decoder needed more bits than the packet holds */
    AMV_ST_RANGE     = 1 << 5,  /* offset/size outside the supplied buffer, or ADPCM step index > 88 */
    AMV_ST_NOSPACE   = 1 << 6,  /* encoder: packet does not fit the per-frame capacity */
    AMV_ST_HEADER    = 1 << 7   /* plain JPEG: the frame's marker segments differ from the configured header */
};

typedef struct amv_params {
    int      device;          /* CUDA ordinal; -1 = current device */
    void    *stream;          /* cudaStream_t to run on; NULL = the context creates its own */
    uint32_t flags;           /* reserved, 0 */
} amv_params;

/* How the encoder lays packets out in `out`. */
enum amv_layout {
    AMV_LAYOUT_PACKED = 0,    /* back to back in frame order; out_off[i] is written by the call */
    AMV_LAYOUT_SLOTS  = 1     /* frame i at out + i*pkt_cap; out_off[i] = i*pkt_cap is written too */
};

AMV_API int         amv_create(const amv_params *params, amv_ctx **out_ctx);
AMV_API void        amv_destroy(amv_ctx *ctx);
AMV_API int         amv_set_stream(amv_ctx *ctx, void *cuda_stream);
AMV_API int         amv_sync(amv_ctx *ctx);
AMV_API const char *amv_strerror(int err);
AMV_API const char *amv_last_error(const amv_ctx *ctx);
AMV_API int         amv_version(void);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
AMV_API uint64_t    amv_launch_count(const amv_ctx *ctx);
/* pinned host memory for AMV_MEM_HOST callers that want zero staging copies */
AMV_API void       *amv_host_alloc(size_t bytes);
AMV_API void        amv_host_free(void *p);

/* Tuning knobs (never change results):
 *   "decode_log2_lanes"            0..5: decode lanes (subsequences) per frame = 1 << value; -1 = from batch size
 *   "encode_slot_workspace_bytes"  cap of the packed-layout staging workspace (frames are sub-batched to fit)
 *   "profile_events"               1 = bracket each hot kernel launch with CUDA events on the context's stream
 *   "host_chunk_frames"            frames per stage of the AMV_MEM_HOST copy/compute pipeline (0 = choose)
 *   "host_zero_copy_packets"       0 = DMA pinned decoder input into a device copy instead of reading it in place
 *   "encode_rounds"                encoder kernels: 4 (default) k_encode16v2 with the transform regrouped for the two integer
 *                                  pipes (dot-product rows, written-out odd columns), 2 k_encode16v2 with the factorised
 *                                  transform, 1 k_encode16 (each + k_encode for the frames it hands back), 0 the one-kernel
 *                                  encoder; 3 = 2 and 8 = 4 at four instead of five CTAs per SM
 *   "decode_token_pass"            AMV / SP5X token pass: 2 (default) lean pass with 16-bit tokens, 1 lean pass with 32-bit
 *                                  tokens, 0 the flat symbol loop with 32-bit tokens; 1 and 2 also run the lean, checkpointed
 *                                  synchronisation pass when frames are split into lanes, 0 the flat one
 *   "scale_form"                   scaler kernel: 1 (default) tiles, 2 tiles with staged source rows, 0 direct
 *   "resample_form"                audio resampler kernel: 2 (default) phase rows, 1 tiles, 0 direct
 * One option DOES select an algorithm, like the reference's AVCodecContext.trellis does:
 *   "adpcm_trellis"                0 (default) = adpcm_ima_compress_sample (adpcm.c:219-227);
 *                                  1..5 = the -trellis N beam search (adpcm_compress_trellis, adpcm.c:287-443) */
AMV_API int         amv_set_option(amv_ctx *ctx, const char *key, int64_t value);
/* "decode_sync_rounds": rounds the last multi-lane decode needed to self-synchronise (max over warps)
 * "<k>_kernel_launches", then "<k>_kernel_ns" (k = encode|idct|idct_bgr|tokens|unstuff|sync|compact|adpcm_dec|adpcm_enc):
 * launches and summed device time of that kernel since the last "_ns" query (needs profile_events) */
AMV_API int64_t     amv_get_stat(amv_ctx *ctx, const char *key);

/* qscale the reference derives from AVFrame.quality (lambda): update_qscale,
 * mpegvideo_enc.c:143-148 with qmin/qmax = 2/31 (utils.c:497-498). Pure host arithmetic. */
AMV_API int amv_qscale_from_quality(int quality, int qmin, int qmax);

/*
 * Decode n AMV video packets (each `FF D8 | stuffed scan | FF D9`, no tables inside) of
 * w x h pixels into YUVJ420P planes.
 *   pkts/pkts_bytes          one buffer holding all packets
 *   pkt_off[i], pkt_size[i]  location of packet i inside pkts
 *   y,u,v                    plane bases of frame 0; frame i's planes start at
 *                            y + i*fs_y, u + i*fs_c, v + i*fs_c (bytes); rows are ls_y / ls_c apart.
 *                            Y is w x h, Cb/Cr are ceil(w/2) x ceil(h/2); rows are stored top-down
 *                            (the codec's bottom-up order is undone, mjpegdec.c:672-677).
 *   status[i]                per-frame status word (may be NULL)
 */
AMV_API int amv_decode_frames(amv_ctx *ctx,
                              const uint8_t *pkts, uint64_t pkts_bytes,
                              const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                              int w, int h,
                              uint8_t *y, uint8_t *u, uint8_t *v,
                              int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                              int32_t *status, int mem);

/*
 * The sibling codec behind the same reference entry point: SP5X ("Sunplus JPEG", sp5x_decoder,
 * libavcodec/sp5xdec.c:33-188,190-201).  Same tables, Huffman and IDCT as AMV; the packet is a
 * 14-byte header followed by the scan with LITERAL FF bytes to the end of the packet (:78-84),
 * and the picture is stored top-down (no flip: mjpegdec.c:672-677 applies to AMV only).
 * Arguments as amv_decode_frames.
 */
AMV_API int amv_decode_frames_sp5x(amv_ctx *ctx,
                                   const uint8_t *pkts, uint64_t pkts_bytes,
                                   const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                                   int w, int h,
                                   uint8_t *y, uint8_t *u, uint8_t *v,
                                   int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                                   int32_t *status, int mem);

/*
 * Plain baseline MJPEG: the `mjpeg_decoder` of the same source file (libavcodec/mjpegdec.c:1356-1367,
 * ff_mjpeg_decode_frame :1106-1340) -- full JPEG frames that carry their own tables: DQT
 * (ff_mjpeg_decode_dqt :113-145), DHT (ff_mjpeg_decode_dht :148-192), SOF0 (ff_mjpeg_decode_sof :194-345),
 * SOS (ff_mjpeg_decode_sos :738-856).  The scan goes through the same kernels as AMV with those tables;
 * the picture is stored top-down.
 *
 * amv_mjpeg_configure reads the marker segments of ONE sample frame (a HOST pointer; the AVCodec shim
 * hands it the first packet): 8-bit SOF0, three components sampled 4:2:0 (2x2 / 1x1 / 1x1, YUVJ420P), 4:2:2
 * (2x1 / 1x1 / 1x1 or the reference encoder's 2x2 / 1x2 / 1x2, YUVJ422P) or 4:4:4 (YUVJ444P), components
 * 1 and 2 sharing their quantiser and Huffman tables, one interleaved sequential scan, with or without a restart
 * interval (DRI; honoured below 1350 MCUs like mjpegdec.c:726 does).  Anything else
 * is AMV_ERR_UNSUPPORTED.  *w / *h receive the picture size (may be NULL).  The chroma planes of the decode
 * call are ceil(w * hc / hmax) x ceil(h * vc / vmax) -- amv_get_stat "mjpeg_chroma_width" / "mjpeg_chroma_height".
 *
 * amv_decode_frames_mjpeg then decodes frames whose bytes up to the end of the SOS header equal the sample's
 * outside the quantiser values: every frame is dequantised with the tables of its OWN DQT segment (the
 * reference's mjpeg_encoder rewrites them whenever rate control moves the quantiser), while Huffman tables,
 * geometry and segment layout must be the sample's.  Any other frame is left undecoded with AMV_ST_HEADER so
 * the caller can group frames by header.  w, h must be the configured size.  Other arguments as
 * amv_decode_frames.
 */
AMV_API int amv_mjpeg_configure(amv_ctx *ctx, const uint8_t *jpeg, uint32_t size, int *w, int *h);
AMV_API int amv_decode_frames_mjpeg(amv_ctx *ctx,
                                    const uint8_t *pkts, uint64_t pkts_bytes,
                                    const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                                    int w, int h,
                                    uint8_t *y, uint8_t *u, uint8_t *v,
                                    int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                                    int32_t *status, int mem);

/*
 * amvlib flavour of the video decoder: what C-AMVDecoder/amvlib computes for the same packets --
 * replaces AmvVideoDecode (amvlib/AMVDec.c:259-286) -> AmvJpegDecode (amvlib/AmvJpeg.c:1515-1539).
 * amvlib is NOT bit-identical to the ffmpeg fork (own quantiser tables AmvJpeg.c:30-61, a zigzag
 * table with a typo :131-141, Chen-Wang IDCT :1078-1175, DC chain from 0 :1177-1242), so it has its
 * own kernels behind this call.  Output: one bottom-up BGR24 bitmap per frame (picture row y at
 * bgr + i*frame_stride + (h-1-y)*line_bytes, 3 bytes B,G,R per pixel, StoreBuffer :789-840); the
 * reference uses line_bytes = ((w*24+31)/32)*4.  Bytes no pixel covers are left untouched.
 * Defined domain: FF bytes of the scan are followed by 00, IDCT results inside amvlib's clamp
 * table (-512..511; outside it the reference reads foreign memory and this call saturates).
 */
AMV_API int amv_decode_frames_bgr24(amv_ctx *ctx,
                                    const uint8_t *pkts, uint64_t pkts_bytes,
                                    const uint64_t *pkt_off, const uint32_t *pkt_size, int n,
                                    int w, int h,
                                    uint8_t *bgr, int line_bytes, uint64_t frame_stride,
                                    int32_t *status, int mem);

/*
 * The pre/post stage the reference's ffmpeg.c puts next to the codec: range conversion between
 * yuv420p (CCIR 601 range) and the codec's yuvj420p (full range) -- img_convert -> img_apply_table
 * with y/c_ccir_to_jpeg resp. y/c_jpeg_to_ccir (libavcodec/imgconvert.c:1216-1260,2492-2510,
 * colorspace.h:69-84).  dir 0: CCIR -> JPEG (before amv_encode_frames), dir 1: JPEG -> CCIR
 * (after amv_decode_frames).  Planes / strides as in amv_decode_frames; output may alias input.
 */
AMV_API int amv_convert_range(amv_ctx *ctx,
                              const uint8_t *y, const uint8_t *u, const uint8_t *v,
                              int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                              int n, int w, int h, int dir,
                              uint8_t *oy, uint8_t *ou, uint8_t *ov,
                              int ols_y, int ols_c, uint64_t ofs_y, uint64_t ofs_c, int mem);

/*
 * The picture scaler the reference's ffmpeg.c runs in front of the encoder for `-s WxH`: sws_getContext /
 * sws_scale of the fork's libavcodec emulation (libavcodec/imgresample.c:515-690), i.e.
 * img_resample_init(ow, oh, iw, ih) + img_resample per frame (:433-507; component_resample :362-431,
 * h_resample :289-360, v_resample :119-153; banks from av_build_filter, resample2.c:93-141).
 * Input planes / strides as in amv_decode_frames for an iw x ih picture, output planes likewise for
 * ow x oh.  As in the reference the chroma planes are scaled at (iw>>1) x (ih>>1) -> (ow>>1) x (oh>>1)
 * with the luma step sizes; bytes outside that area are neither read nor written.
 */
AMV_API int amv_scale_frames(amv_ctx *ctx,
                             const uint8_t *y, const uint8_t *u, const uint8_t *v,
                             int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                             int n, int iw, int ih,
                             uint8_t *oy, uint8_t *ou, uint8_t *ov,
                             int ols_y, int ols_c, uint64_t ofs_y, uint64_t ofs_c,
                             int ow, int oh, int mem);

/*
 * The same with the pixel-format handling of the fork's sws_scale around it (imgresample.c:599-690): the
 * scaler itself only knows PIX_FMT_YUV420P, so a YUVJ420P source is first taken to YUV420P and a YUVJ420P
 * destination is produced from the scaled YUV420P picture, both through img_convert (= amv_convert_range,
 * dir 1 resp. dir 0).  The AMV encoder takes YUVJ420P, so `ffmpeg -s WxH ... -f amv` always runs the second
 * conversion, and the first one too when the source decoder outputs YUVJ420P (mjpeg, amv).
 * flags: AMV_SCALE_IN_JPEG_RANGE (source is YUVJ420P), AMV_SCALE_OUT_JPEG_RANGE (destination is YUVJ420P).
 * The conversions cover the area the scaler reads / writes.  The source planes are never modified.
 */
#define AMV_SCALE_IN_JPEG_RANGE  1
#define AMV_SCALE_OUT_JPEG_RANGE 2
AMV_API int amv_scale_frames_ex(amv_ctx *ctx,
                                const uint8_t *y, const uint8_t *u, const uint8_t *v,
                                int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                                int n, int iw, int ih,
                                uint8_t *oy, uint8_t *ou, uint8_t *ov,
                                int ols_y, int ols_c, uint64_t ofs_y, uint64_t ofs_c,
                                int ow, int oh, int flags, int mem);

/*
 * The audio resampler of the reference's do_audio_out (ffmpeg.c:501-505) in front of the ADPCM encoder
 * (which takes 22050 Hz mono only, adpcm.c:190-199): audio_resample_init(1, in_channels, out_rate, in_rate)
 * + audio_resample (libavcodec/resample.c:93-235; two channels are averaged, :53-75) -> av_resample
 * (resample2.c:234-323; 16-tap / cutoff 0.8 / 1024-phase Kaiser bank of av_resample_init :185-206).
 *   in        n_in samples per channel, interleaved int16 (4-byte aligned for 2 channels)
 *   out       mono int16, out_cap samples; *n_out receives the count written
 * The result is the concatenation of what the reference returns when the stream is fed to
 * audio_resample in calls of any size >= the filter length (it carries the unconsumed tail from call to
 * call and never flushes it): amv_audio_resample_count(n_in, in_rate, out_rate) samples, all those whose
 * taps lie inside the stream.  Index arithmetic is 64-bit (the reference's int index restarts per call).
 */
AMV_API uint64_t amv_audio_resample_count(uint64_t n_in, int in_rate, int out_rate);
AMV_API int amv_audio_resample(amv_ctx *ctx,
                               const int16_t *in, uint64_t n_in, int in_channels,
                               int in_rate, int out_rate,
                               int16_t *out, uint64_t out_cap, uint64_t *n_out, int mem);
/*
 * The same resampler over a window of a longer stream, for callers that feed it packet by packet the way
 * ffmpeg.c feeds audio_resample (which keeps the unconsumed tail in ReSampleContext.temp and the position in
 * AVResampleContext.index / frac, resample.c:214-216, resample2.c:303-313): `in` holds the stream's samples
 * [in_base, in_base + n_in) and the call writes outputs k_start, k_start + 1, ... up to the last one whose taps
 * end inside the window, i.e. amv_audio_resample_count(in_base + n_in, ...) - k_start samples, or out_cap of them
 * if that is fewer (av_resample's dst_size; the rest comes with a later call).
 * amv_audio_resample_first_tap(k, ...) is the first input sample output k reads (negative for the mirrored head
 * of the stream, which needs in_base == 0): a caller may drop everything in front of first_tap(next k).
 * glue/ffmpeg/amvcuda_resample.c builds its audio_resample replacement on these two; sharding.resample_shard (Python
 * host side) splits one stream across GPUs by output range with them.
 */
AMV_API int amv_audio_resample_from(amv_ctx *ctx,
                                    const int16_t *in, uint64_t in_base, uint64_t n_in, int in_channels,
                                    int in_rate, int out_rate, uint64_t k_start,
                                    int16_t *out, uint64_t out_cap, uint64_t *n_out, int mem);
AMV_API int64_t amv_audio_resample_first_tap(uint64_t k, int in_rate, int out_rate);

/*
 * The filter banks the two stages above run on, as built on the host at call time (no device involved):
 * av_build_filter (resample2.c:93-141) with the arguments of img_resample_init (imgresample.c:476-479:
 * 4 taps x 16 phases, scale 256, cubic; h_incr / v_incr :473-474) resp. av_resample_init
 * (resample2.c:185-206: filter_length taps x 1024 phases, scale 1 << 15, Kaiser beta 9).
 * amv_scale_banks fills 64 coefficients each; amv_audio_resample_bank returns filter_length (>= 1) and, when
 * bank != NULL, fills filter_length * 1024 coefficients (bank_cap in coefficients).
 */
AMV_API int amv_scale_banks(int iw, int ih, int ow, int oh, int16_t *h_bank, int16_t *v_bank,
                            int32_t *h_incr, int32_t *v_incr);
AMV_API int amv_audio_resample_bank(int in_rate, int out_rate, int16_t *bank, uint64_t bank_cap);

/*
 * Encode n YUVJ420P frames into AMV packets, byte-identical to amv_encoder.
 *   y,u,v, ls_*, fs_*        as above (source planes)
 *   qscale                   per-frame quantiser scale 2..31 (NULL = 2, the reference default;
 *                            use amv_qscale_from_quality for AVFrame.quality)
 *   out/out_cap              packet buffer; pkt_cap = capacity reserved per frame
 *   out_off[i], out_size[i]  where packet i was written and its size (0 if status[i] != 0)
 * Requires (h/2)%8 in {0,4} (otherwise the reference reads outside the picture).
 */
AMV_API int amv_encode_frames(amv_ctx *ctx,
                              const uint8_t *y, const uint8_t *u, const uint8_t *v,
                              int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                              int n, int w, int h, const int32_t *qscale,
                              uint8_t *out, uint64_t out_cap, uint32_t pkt_cap, int layout,
                              uint64_t *out_off, uint32_t *out_size,
                              int32_t *status, int mem);

/*
 * Decode n IMA-ADPCM-AMV chunks (le16 predictor, le16 step index, le32 count, nibbles)
 * into mono int16 PCM.  Chunk i yields 2*(chunk_size[i]-8) samples at pcm + pcm_off[i]
 * (offsets in samples).
 */
AMV_API int amv_adpcm_dec_chunks(amv_ctx *ctx,
                                 const uint8_t *chunks, uint64_t chunks_bytes,
                                 const uint64_t *chunk_off, const uint32_t *chunk_size, int n,
                                 int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off,
                                 int32_t *status, int mem);

/*
 * Encode n chunks.  Chunk i takes nsamples[i] (even) samples from pcm + pcm_off[i] and the
 * encoder state step_in[i] (0..88; NULL = 0); writes 8 + nsamples[i]/2 bytes at
 * out + out_off[i] and the state after the chunk to step_out[i] (may be NULL).  A single
 * stream is encoded by chaining step_out[i] -> step_in[i+1] on the host (adpcm.c:466);
 * independent streams/chunks run in parallel.
 */
AMV_API int amv_adpcm_enc_chunks(amv_ctx *ctx,
                                 const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off,
                                 const uint32_t *nsamples, const int16_t *step_in, int16_t *step_out, int n,
                                 uint8_t *out, uint64_t out_bytes, const uint64_t *out_off,
                                 int32_t *status, int mem);

/*
 * Encode `nstreams` continuous streams chunk by chunk with the state carried inside each
 * stream exactly like repeated adpcm_encode_frame calls (adpcm.c:461-496): stream s is
 * chunks [first_chunk[s], first_chunk[s+1]) of the arrays above (first_chunk has
 * nstreams+1 entries); step_in gives the state before each stream's first chunk.
 */
AMV_API int amv_adpcm_enc_streams(amv_ctx *ctx,
                                  const int16_t *pcm, uint64_t pcm_samples, const uint64_t *pcm_off,
                                  const uint32_t *nsamples, const uint32_t *first_chunk, int nstreams, int nchunks,
                                  const int16_t *step_in, int16_t *step_out,
                                  uint8_t *out, uint64_t out_bytes, const uint64_t *out_off,
                                  int32_t *status, int mem);

/* ------------------------------------------------------------------ AMV container (host side) */
typedef struct amv_file_info {
    int      width, height, fps;       /* amvh (amvenc.c:131-178, amvlib/AMVHeader.h:18-40) */
    int      sample_rate, channels;    /* audio strf (riff.c:240-289) */
    uint32_t us_per_frame, nb_frames_header, duration_s;
    uint32_t nvideo, naudio;           /* chunks found in movi */
    uint64_t movi_offset;              /* file offset of the "movi" tag (0x138 in real and reference-muxed files) */
    int      has_end_marker, truncated;
} amv_file_info;

/*
 * Index an AMV file held in memory: per video packet / audio chunk its offset INTO `file` and its
 * size, so `file` itself can be handed to amv_decode_frames / amv_adpcm_dec_chunks as the packet
 * buffer (pinned -> read in place by the kernels).  Replaces, for AMV, avi_read_header /
 * avi_read_packet (libavformat/avidec.c with the amvh hooks :237,283,320,429-434) and amvlib's
 * AmvOpen / AmvReadNextFrame (amvlib/AMVDec.c:15-238).  Arrays hold up to `cap` entries (may be NULL
 * to count); info->nvideo / naudio report what the file holds.
 */
AMV_API int amv_file_index(const uint8_t *file, uint64_t size, amv_file_info *info,
                           uint64_t *v_off, uint32_t *v_size, uint64_t *a_off, uint32_t *a_size, uint32_t cap);

typedef struct amv_mux_params {
    int width, height;
    int tb_num, tb_den;                /* video time base: 1 / fps */
    int sample_rate;                   /* 22050 for the reference encoder */
    int video_bit_rate, audio_bit_rate;/* 0 = the reference's defaults (200000 / 64000); only header fields */
} amv_mux_params;

/*
 * Write an AMV file: n video packets and n audio chunks, alternating.  Byte-identical to the
 * reference's amv_muxer (libavformat/amvenc.c) driven like ffmpeg.c drives it for `-f amv`.
 * Returns the file size, or -(bytes needed) if cap is too small (call with out=NULL, cap=0 to size).
 */
AMV_API int64_t amv_file_mux(const amv_mux_params *mp, int n,
                             const uint8_t *vpk, const uint64_t *v_off, const uint32_t *v_size,
                             const uint8_t *apk, const uint64_t *a_off, const uint32_t *a_size,
                             uint8_t *out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* AMVCUDA_H */

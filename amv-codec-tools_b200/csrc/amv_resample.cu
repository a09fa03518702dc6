// amv_resample.cu -- the two resampling stages ffmpeg.c runs in front of the AMV encoders:
//   * the picture scaler of `-s WxH`: sws_scale of the fork's libavcodec emulation (imgresample.c:515-690) ->
//     img_resample_init / img_resample (:433-507) -> component_resample (:362-431), h_resample (:289-360),
//     v_resample (:119-153): separable 4-tap / 16-phase filter in 16.16 source positions, 8-bit coefficients,
//     the horizontal result clamped to a byte before the vertical pass, edges repeated;
//   * the audio resampler of do_audio_out (ffmpeg.c:501-505): audio_resample (resample.c:131-235; two channels
//     averaged to one, :53-75) -> av_resample (resample2.c:234-323): filter_length taps of a 1024-phase bank,
//     32-bit accumulator, >> 15 with rounding, saturation.
// Both banks come from av_build_filter (resample2.c:93-141), double arithmetic rounded through lrintf: built on
// the host (build_scale_banks / build_resample_bank below, operand order as in the reference so the
// coefficients are bit-identical) and handed to the kernels as data.
// Every output pixel / sample depends only on the input: one thread per 4 output pixels resp. per output sample,
// HBM-bound work (the taps of neighbouring outputs overlap and come from L1 / L2).
#include <cmath>
#include <vector>
#include "amv_common.cuh"
#include "amv_kernels.h"

namespace amv {

// ------------------------------------------------------------------------------------------ banks (host)
namespace {
double bessel_i0(double x) {
    double v = 1, t = 1;
    x = x * x / 4;
    for (int i = 1; i < 50; i++) { t *= x / (i * i); v += t; }
    return v;
}
// type 0: cubic, first derivative -0.5 (the scaler); type >= 2: Kaiser-windowed sinc with beta = type (audio: 9)
void build_bank(int16_t *bank, double factor, int taps, int phases, int scale, int type) {
    const int center = (taps - 1) / 2;
    std::vector<double> tab((size_t)taps);
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
            tab[(size_t)i] = y;
            norm += y;
        }
        for (int i = 0; i < taps; i++) {
            const long c = lrintf((float)(tab[(size_t)i] * scale / norm));
            bank[ph * taps + i] = (int16_t)(c < -32768 ? -32768 : c > 32767 ? 32767 : c);
        }
    }
}
}  // namespace

void build_scale_banks(int iw, int ih, int ow, int oh, ScaleBanks *b) {
    build_bank(b->h, (float)ow / (float)iw, 4, 16, 256, 0);
    build_bank(b->v, (float)oh / (float)ih, 4, 16, 256, 0);
    b->h_incr = (int)(((int64_t)iw * 65536) / ow);
    b->v_incr = (int)(((int64_t)ih * 65536) / oh);
}

int resample_filter_length(int in_rate, int out_rate) {
    double factor = out_rate * 0.8 / in_rate;
    if (factor > 1.0) factor = 1.0;
    const int len = (int)ceil(16 / factor);
    return len < 1 ? 1 : len;
}
void build_resample_bank(int in_rate, int out_rate, int16_t *bank) {
    double factor = out_rate * 0.8 / in_rate;
    if (factor > 1.0) factor = 1.0;
    build_bank(bank, factor, resample_filter_length(in_rate, out_rate), 1024, 1 << 15, 9);
}
// outputs the reference produces from n_in samples: first(k) = (index0 + floor(k * in_rate * 1024 / out_rate)) >> 10; every k
// whose taps start left of sample 0 (they mirror, no end check, resample2.c:256-258), then every k with
// first(k) + len <= n_in (:259-260)
int64_t resample_output_count(int64_t n_in, int in_rate, int out_rate) {
    const int len = resample_filter_length(in_rate, out_rate);
    const int64_t index0 = -1024 * (int64_t)((len - 1) / 2);
    if (n_in <= 0) return 0;
    const __int128 D = (__int128)in_rate * 1024, S = out_rate;
    const int64_t mirrored = (int64_t)(((__int128)(-index0) * S + D - 1) / D);      // k with floor(k*D/S) < -index0
    // largest index with (index >> 10) + len <= n_in is ((n_in - len + 1) << 10) - 1
    const int64_t lim = ((n_in - len + 1) * 1024) - 1 - index0;      // floor(k*D/S) <= lim
    if (lim < 0) return mirrored;
    int64_t k = (int64_t)((((__int128)lim + 1) * S - 1) / D);         // largest k with k*D < (lim+1)*S
    while ((int64_t)(((__int128)(k + 1) * D) / S) <= lim) k++;
    while (k >= 0 && (int64_t)(((__int128)k * D) / S) > lim) k--;
    return k + 1 > mirrored ? k + 1 : mirrored;
}

// first input sample output k reads (negative: the mirrored head)
int64_t resample_first_tap(int64_t k, int in_rate, int out_rate) {
    const int len = resample_filter_length(in_rate, out_rate);
    const int64_t index0 = -1024 * (int64_t)((len - 1) / 2);
    const int64_t index = index0 + (int64_t)(((__int128)k * in_rate * 1024) / out_rate);
    return index >> 10;
}

// ------------------------------------------------------------------------------------------ picture scaler
struct ScaleArgs {
    const uint8_t *src; uint8_t *dst;
    int iw, ih, ow, oh, ils, ols;
    uint64_t ifs, ofs;
    int n, h_incr, v_incr;
};

__device__ __forceinline__ int clip255(int v) { return __vimin_s32_relu(v, 255); }

// Direct form: one thread computes 4 neighbouring output pixels of one row from global memory (VEC: one 32-bit store).
template <bool VEC>
__global__ void __launch_bounds__(256)
k_scale_plane(ScaleArgs a, ScaleBanks banks) {
    __shared__ int2 s_h[16], s_v[16];               // a phase's 4 coefficients as two packed pairs
    if (threadIdx.x < 16) {
        const int16_t *h = banks.h + 4 * threadIdx.x, *v = banks.v + 4 * threadIdx.x;
        s_h[threadIdx.x] = make_int2((uint16_t)h[0] | ((int)h[1] << 16), (uint16_t)h[2] | ((int)h[3] << 16));
        s_v[threadIdx.x] = make_int2((uint16_t)v[0] | ((int)v[1] << 16), (uint16_t)v[2] | ((int)v[3] << 16));
    }
    __syncthreads();
    const int units = (a.ow + 3) >> 2;
    const int64_t total = (int64_t)units * a.oh * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ux = (int)(i % units);
        const int64_t ry = i / units;
        const int y = (int)(ry % a.oh), f = (int)(ry / a.oh);
        const uint8_t *in = a.src + (uint64_t)f * a.ifs;
        const int sy = 2 * 65536 + y * a.v_incr, bottom = sy >> 16;
        const int2 fv = s_v[(sy >> 12) & 15];
        const int fv0 = (int16_t)fv.x, fv1 = fv.x >> 16, fv2 = (int16_t)fv.y, fv3 = fv.y >> 16;
        const uint8_t *r0 = in + (int64_t)a.ils * min(max(bottom - 3, 0), a.ih - 1);
        const uint8_t *r1 = in + (int64_t)a.ils * min(max(bottom - 2, 0), a.ih - 1);
        const uint8_t *r2 = in + (int64_t)a.ils * min(max(bottom - 1, 0), a.ih - 1);
        const uint8_t *r3 = in + (int64_t)a.ils * min(max(bottom, 0), a.ih - 1);
        uint32_t word = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int x = ux * 4 + p;
            const int sx = -65536 + x * a.h_incr, left = sx >> 16;
            const int2 fh = s_h[(sx >> 12) & 15];
            const int h0 = (int16_t)fh.x, h1 = fh.x >> 16, h2 = (int16_t)fh.y, h3 = fh.y >> 16;
            int c0, c1, c2, c3;
            if (left >= 0 && left + 3 < a.iw) { c0 = left; c1 = left + 1; c2 = left + 2; c3 = left + 3; }
            else {
                c0 = min(max(left, 0), a.iw - 1); c1 = min(max(left + 1, 0), a.iw - 1);
                c2 = min(max(left + 2, 0), a.iw - 1); c3 = min(max(left + 3, 0), a.iw - 1);
            }
            const int a0 = clip255((__ldg(r0 + c0) * h0 + __ldg(r0 + c1) * h1 + __ldg(r0 + c2) * h2 + __ldg(r0 + c3) * h3) >> 8);
            const int a1 = clip255((__ldg(r1 + c0) * h0 + __ldg(r1 + c1) * h1 + __ldg(r1 + c2) * h2 + __ldg(r1 + c3) * h3) >> 8);
            const int a2 = clip255((__ldg(r2 + c0) * h0 + __ldg(r2 + c1) * h1 + __ldg(r2 + c2) * h2 + __ldg(r2 + c3) * h3) >> 8);
            const int a3 = clip255((__ldg(r3 + c0) * h0 + __ldg(r3 + c1) * h1 + __ldg(r3 + c2) * h2 + __ldg(r3 + c3) * h3) >> 8);
            const uint32_t px = (uint32_t)clip255((a0 * fv0 + a1 * fv1 + a2 * fv2 + a3 * fv3) >> 8);
            if (VEC) word |= px << (8 * p);
            else if (x < a.ow) a.dst[(uint64_t)f * a.ofs + (int64_t)y * a.ols + x] = (uint8_t)px;
        }
        if (VEC) *reinterpret_cast<uint32_t *>(a.dst + (uint64_t)f * a.ofs + (int64_t)y * a.ols + ux * 4) = word;
    }
}

// Tiled form (the one that normally runs): a block owns a 64 x tile_h tile of the output.  Pass 1 filters every source
// row the tile's vertical taps touch horizontally, once, into shared memory (what the reference keeps in its ring
// of line buffers); pass 2 runs the vertical filter from there, 4 pixels per thread.  Against the direct form this
// drops the horizontal work from 4 rows per output row to (rows touched) / 16 and turns 16 byte loads per pixel
// into (rows touched) / 4 of them plus one 32-bit shared-memory read.
constexpr int kTileW = 64, kTileRowsMax = 320;      // tile height: 256, 128, 64, 32 or 16 output rows, the tallest whose source rows fit

template <bool VEC>
__global__ void __launch_bounds__(256)
k_scale_tile(ScaleArgs a, ScaleBanks banks, int tiles_x, int tiles_y, int tile_h) {
    __shared__ int2 s_h[16], s_v[16];
    __shared__ __align__(16) uint8_t s_line[kTileRowsMax][kTileW];
    if (threadIdx.x < 16) {
        const int16_t *h = banks.h + 4 * threadIdx.x, *v = banks.v + 4 * threadIdx.x;
        s_h[threadIdx.x] = make_int2((uint16_t)h[0] | ((int)h[1] << 16), (uint16_t)h[2] | ((int)h[3] << 16));
        s_v[threadIdx.x] = make_int2((uint16_t)v[0] | ((int)v[1] << 16), (uint16_t)v[2] | ((int)v[3] << 16));
    }
    __syncthreads();
    const int tx = blockIdx.x % tiles_x, tyf = blockIdx.x / tiles_x;
    const int ty = tyf % tiles_y, f = tyf / tiles_y;
    const int x0 = tx * kTileW, y0 = ty * tile_h;
    const int y_last = min(y0 + tile_h, a.oh) - 1;
    const int r_lo = ((2 * 65536 + y0 * a.v_incr) >> 16) - 3;
    const int nrows = ((2 * 65536 + y_last * a.v_incr) >> 16) - r_lo + 1;          // <= kTileRowsMax (checked by the launcher)
    const uint8_t *in = a.src + (uint64_t)f * a.ifs;
    {   // pass 1: column `col` of the tile for rows threadIdx.x / 64, + 4, + 8, ...
        const int col = threadIdx.x & (kTileW - 1);
        const int sx = -65536 + (x0 + col) * a.h_incr, left = sx >> 16;
        const int2 fh = s_h[(sx >> 12) & 15];
        const int h0 = (int16_t)fh.x, h1 = fh.x >> 16, h2 = (int16_t)fh.y, h3 = fh.y >> 16;
        const int c0 = min(max(left, 0), a.iw - 1), c1 = min(max(left + 1, 0), a.iw - 1);
        const int c2 = min(max(left + 2, 0), a.iw - 1), c3 = min(max(left + 3, 0), a.iw - 1);
#pragma unroll 4
        for (int r = threadIdx.x >> 6; r < nrows; r += 4) {
            const uint8_t *row = in + (int64_t)a.ils * min(max(r_lo + r, 0), a.ih - 1);
            s_line[r][col] = (uint8_t)clip255((__ldg(row + c0) * h0 + __ldg(row + c1) * h1 + __ldg(row + c2) * h2 + __ldg(row + c3) * h3) >> 8);
        }
    }
    __syncthreads();
    // pass 2: rows threadIdx.x / 16, + 16, ... of the tile, 4 pixels each
    const int cx = (threadIdx.x & 15) * 4;
    if (x0 + cx >= a.ow) return;
    for (int y = y0 + (threadIdx.x >> 4); y <= y_last; y += 16) {
        const int sy = 2 * 65536 + y * a.v_incr;
        const int2 fv = s_v[(sy >> 12) & 15];
        const int fv0 = (int16_t)fv.x, fv1 = fv.x >> 16, fv2 = (int16_t)fv.y, fv3 = fv.y >> 16;
        const int r = (sy >> 16) - 3 - r_lo;
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(&s_line[r][cx]), w1 = *reinterpret_cast<const uint32_t *>(&s_line[r + 1][cx]);
        const uint32_t w2 = *reinterpret_cast<const uint32_t *>(&s_line[r + 2][cx]), w3 = *reinterpret_cast<const uint32_t *>(&s_line[r + 3][cx]);
        uint32_t word = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int v = (int)((w0 >> (8 * p)) & 0xff) * fv0 + (int)((w1 >> (8 * p)) & 0xff) * fv1 +
                          (int)((w2 >> (8 * p)) & 0xff) * fv2 + (int)((w3 >> (8 * p)) & 0xff) * fv3;
            word |= (uint32_t)clip255(v >> 8) << (8 * p);
        }
        uint8_t *d = a.dst + (uint64_t)f * a.ofs + (int64_t)y * a.ols + x0 + cx;
        if (VEC) *reinterpret_cast<uint32_t *>(d) = word;
        else {
#pragma unroll
            for (int p = 0; p < 4; p++)
                if (x0 + cx + p < a.ow) d[p] = (uint8_t)(word >> (8 * p));
        }
    }
}

// Staged form of the tile kernel (option scale_form 2; measured slower): the stretch of every source row the tile's taps touch is first
// copied into shared memory -- 32-bit loads when the plane's base and pitches allow (WORDS), bytes otherwise -- so
// pass 1 reads its 4 taps per item from shared memory instead of issuing 4 byte loads to L1 each.
// Dynamic shared memory: rows_max x (pitch + 64) bytes.
template <bool VEC, bool WORDS>
__global__ void __launch_bounds__(256)
k_scale_tile_staged(ScaleArgs a, ScaleBanks banks, int tiles_x, int tiles_y, int tile_h, int rows_max, int pitch) {
    AMV_EXTERN_SHARED(uint8_t, s_dyn, 16);
    __shared__ int2 s_h[16], s_v[16];
    uint8_t *s_src = s_dyn;                                   // [rows_max][pitch]
    uint8_t *s_line = s_dyn + (size_t)rows_max * pitch;       // [rows_max][64]
    if (threadIdx.x < 16) {
        const int16_t *h = banks.h + 4 * threadIdx.x, *v = banks.v + 4 * threadIdx.x;
        s_h[threadIdx.x] = make_int2((uint16_t)h[0] | ((int)h[1] << 16), (uint16_t)h[2] | ((int)h[3] << 16));
        s_v[threadIdx.x] = make_int2((uint16_t)v[0] | ((int)v[1] << 16), (uint16_t)v[2] | ((int)v[3] << 16));
    }
    const int tx = blockIdx.x % tiles_x, tyf = blockIdx.x / tiles_x;
    const int ty = tyf % tiles_y, f = tyf / tiles_y;
    const int x0 = tx * kTileW, y0 = ty * tile_h;
    const int y_last = min(y0 + tile_h, a.oh) - 1, x_last = min(x0 + kTileW, a.ow) - 1;
    const int r_lo = ((2 * 65536 + y0 * a.v_incr) >> 16) - 3;
    const int nrows = ((2 * 65536 + y_last * a.v_incr) >> 16) - r_lo + 1;
    // source columns the tile's taps touch (clamped), start rounded down to a word
    const int c_lo = min(max((-65536 + x0 * a.h_incr) >> 16, 0), a.iw - 1) & ~3;
    const int c_hi = min(max(((-65536 + x_last * a.h_incr) >> 16) + 3, 0), a.iw - 1);
    const uint8_t *in = a.src + (uint64_t)f * a.ifs;
    if (WORDS) {
        const int wpr = ((c_hi - c_lo) >> 2) + 1;             // words per row (iw is a multiple of 4: none reaches past the row)
        for (int i = threadIdx.x; i < nrows * wpr; i += 256) {
            const int r = i / wpr, c = i - r * wpr;
            const uint8_t *row = in + (int64_t)a.ils * min(max(r_lo + r, 0), a.ih - 1);
            *reinterpret_cast<uint32_t *>(s_src + (size_t)r * pitch + 4 * c) = __ldg(reinterpret_cast<const uint32_t *>(row + c_lo) + c);
        }
    } else {
        const int bpr = c_hi - c_lo + 1;
        for (int i = threadIdx.x; i < nrows * bpr; i += 256) {
            const int r = i / bpr, c = i - r * bpr;
            const uint8_t *row = in + (int64_t)a.ils * min(max(r_lo + r, 0), a.ih - 1);
            s_src[(size_t)r * pitch + c] = __ldg(row + c_lo + c);
        }
    }
    __syncthreads();
    {   // pass 1 from shared memory: column `col` of the tile for rows threadIdx.x / 64, + 4, + 8, ...
        const int col = threadIdx.x & (kTileW - 1);
        const int sx = -65536 + min(x0 + col, x_last) * a.h_incr, left = sx >> 16;
        const int2 fh = s_h[(sx >> 12) & 15];
        const int h0 = (int16_t)fh.x, h1 = fh.x >> 16, h2 = (int16_t)fh.y, h3 = fh.y >> 16;
        const int c0 = min(max(left, 0), a.iw - 1) - c_lo, c1 = min(max(left + 1, 0), a.iw - 1) - c_lo;
        const int c2 = min(max(left + 2, 0), a.iw - 1) - c_lo, c3 = min(max(left + 3, 0), a.iw - 1) - c_lo;
#pragma unroll 4
        for (int r = threadIdx.x >> 6; r < nrows; r += 4) {
            const uint8_t *row = s_src + (size_t)r * pitch;
            s_line[r * kTileW + col] = (uint8_t)clip255((row[c0] * h0 + row[c1] * h1 + row[c2] * h2 + row[c3] * h3) >> 8);
        }
    }
    __syncthreads();
    // pass 2: rows threadIdx.x / 16, + 16, ... of the tile, 4 pixels each
    const int cx = (threadIdx.x & 15) * 4;
    if (x0 + cx >= a.ow) return;
    for (int y = y0 + (threadIdx.x >> 4); y <= y_last; y += 16) {
        const int sy = 2 * 65536 + y * a.v_incr;
        const int2 fv = s_v[(sy >> 12) & 15];
        const int fv0 = (int16_t)fv.x, fv1 = fv.x >> 16, fv2 = (int16_t)fv.y, fv3 = fv.y >> 16;
        const uint8_t *l = s_line + ((sy >> 16) - 3 - r_lo) * kTileW + cx;
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(l), w1 = *reinterpret_cast<const uint32_t *>(l + kTileW);
        const uint32_t w2 = *reinterpret_cast<const uint32_t *>(l + 2 * kTileW), w3 = *reinterpret_cast<const uint32_t *>(l + 3 * kTileW);
        uint32_t word = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int v = (int)((w0 >> (8 * p)) & 0xff) * fv0 + (int)((w1 >> (8 * p)) & 0xff) * fv1 +
                          (int)((w2 >> (8 * p)) & 0xff) * fv2 + (int)((w3 >> (8 * p)) & 0xff) * fv3;
            word |= (uint32_t)clip255(v >> 8) << (8 * p);
        }
        uint8_t *d = a.dst + (uint64_t)f * a.ofs + (int64_t)y * a.ols + x0 + cx;
        if (VEC) *reinterpret_cast<uint32_t *>(d) = word;
        else {
#pragma unroll
            for (int p = 0; p < 4; p++)
                if (x0 + cx + p < a.ow) d[p] = (uint8_t)(word >> (8 * p));
        }
    }
}

static void launch_scale_plane(const uint8_t *src, uint8_t *dst, int iw, int ih, int ow, int oh, int ils, int ols, uint64_t ifs,
                               uint64_t ofs, int n, const ScaleBanks &b, int form, cudaStream_t s) {
    if (iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0) return;       // a 1-pixel-wide picture has no chroma to scale
    ScaleArgs a{ src, dst, iw, ih, ow, oh, ils, ols, ifs, ofs, n, b.h_incr, b.v_incr };
    const bool vec = (ow & 3) == 0 && ((((uintptr_t)dst | (uintptr_t)ols | ofs) & 3) == 0);
    // the tallest tile whose source rows fit the line buffer: taller tiles share more horizontally filtered rows between
    // their output rows and spread the per-block set-up over more pixels (measured at 2:1: 4.19 / 3.36 / 3.27 ms for
    // 16 / 64 / 128 rows; 1:3 enlargement: 7.2 / 4.15 / 3.56 / 3.29 ms with 256)
    int tile_h = 256;
    const int64_t cols_touched = (((int64_t)(kTileW - 1) * b.h_incr) >> 16) + 5 + 3;      // + word alignment of the first column
    const int64_t pitch = (cols_touched + 3 + 15) & ~(int64_t)15;
    auto rows_of = [&](int th) { return (((int64_t)(th - 1) * b.v_incr) >> 16) + 5; };
    while (tile_h > 16 && (rows_of(tile_h) > kTileRowsMax || (form >= 2 && rows_of(tile_h) * (pitch + kTileW) > 40 * 1024))) tile_h >>= 1;
    const int tiles_x = (ow + kTileW - 1) / kTileW, tiles_y = (oh + tile_h - 1) / tile_h;
    const int64_t tiles = (int64_t)tiles_x * tiles_y * n;
    const int64_t rows_touched = rows_of(tile_h);
    const int64_t smem = rows_touched * (pitch + kTileW);
    // 32-bit staging loads: word-aligned rows of a whole number of words (no load reaches past a row's last pixel)
    const bool words = ((((uintptr_t)src | (uintptr_t)ils | ifs) & 3) == 0) && (iw & 3) == 0;
    if (form >= 2 && rows_touched <= kTileRowsMax && smem <= 40 * 1024 && tiles <= 0x7fffffff) {
        const unsigned g = (unsigned)tiles;
        const int rm = (int)rows_touched, pt = (int)pitch;
        if (vec && words)       AMV_LAUNCH((k_scale_tile_staged<true, true>), g, 256, (size_t)smem, s, a, b, tiles_x, tiles_y, tile_h, rm, pt);
        else if (vec)           AMV_LAUNCH((k_scale_tile_staged<true, false>), g, 256, (size_t)smem, s, a, b, tiles_x, tiles_y, tile_h, rm, pt);
        else if (words)         AMV_LAUNCH((k_scale_tile_staged<false, true>), g, 256, (size_t)smem, s, a, b, tiles_x, tiles_y, tile_h, rm, pt);
        else                    AMV_LAUNCH((k_scale_tile_staged<false, false>), g, 256, (size_t)smem, s, a, b, tiles_x, tiles_y, tile_h, rm, pt);
        return;
    }
    if (form >= 1 && rows_touched <= kTileRowsMax && tiles <= 0x7fffffff) {  // taps straight from global memory (L1)
        if (vec) AMV_LAUNCH(k_scale_tile<true>, (unsigned)tiles, 256, 0, s, a, b, tiles_x, tiles_y, tile_h);
        else     AMV_LAUNCH(k_scale_tile<false>, (unsigned)tiles, 256, 0, s, a, b, tiles_x, tiles_y, tile_h);
        return;
    }
    // extreme reductions (the tile's rows exceed the line buffer): the direct form
    const int64_t total = (int64_t)((ow + 3) >> 2) * oh * n;
    int64_t grid = (total + 255) / 256;
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    if (grid < 1) grid = 1;
    if (vec) AMV_LAUNCH(k_scale_plane<true>, (unsigned)grid, 256, 0, s, a, b);
    else     AMV_LAUNCH(k_scale_plane<false>, (unsigned)grid, 256, 0, s, a, b);
}

int launch_scale_frames(const uint8_t *y, const uint8_t *u, const uint8_t *v, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c,
                        int n, int iw, int ih, uint8_t *oy, uint8_t *ou, uint8_t *ov, int ols_y, int ols_c, uint64_t ofs_y,
                        uint64_t ofs_c, int ow, int oh, const ScaleBanks &b, int form, cudaStream_t s) {
    int launches = 1;
    launch_scale_plane(y, oy, iw, ih, ow, oh, ls_y, ols_y, fs_y, ofs_y, n, b, form, s);
    if ((iw >> 1) > 0 && (ih >> 1) > 0 && (ow >> 1) > 0 && (oh >> 1) > 0) {
        launch_scale_plane(u, ou, iw >> 1, ih >> 1, ow >> 1, oh >> 1, ls_c, ols_c, fs_c, ofs_c, n, b, form, s);
        launch_scale_plane(v, ov, iw >> 1, ih >> 1, ow >> 1, oh >> 1, ls_c, ols_c, fs_c, ofs_c, n, b, form, s);
        launches += 2;
    }
    return launches;
}

// ------------------------------------------------------------------------------------------ audio resampler
template <int CH>
__device__ __forceinline__ int mono_at(const int16_t *in, int64_t i) {
    if (CH == 1) return __ldg(in + i);
    const uint32_t lr = __ldg(reinterpret_cast<const uint32_t *>(in) + i);      // interleaved pair, 4-byte aligned
    return (int)(int16_t)(((int)(int16_t)(lr & 0xffff) + ((int)lr >> 16)) >> 1);
}

// The device bank holds every phase row padded with zero coefficients to a multiple of 8 (len8), so a row starts on a
// 16-byte boundary and is fetched 8 coefficients at a time.
__device__ __forceinline__ int16_t round_sat(uint32_t acc) {
    const int val = ((int)(acc + (1u << 14))) >> 15;
    return (int16_t)max(-32768, min(32767, val));
}

// Direct form: one thread per output sample k straight from global memory; index(k) = index0 + floor(k * in_rate * 1024 / out_rate).
template <int CH>
__device__ __forceinline__ uint32_t taps_direct(const int16_t *__restrict__ in, int64_t n_in, const int16_t *__restrict__ f, int len,
                                                int64_t first, int64_t in_base) {
    uint32_t acc = 0;
    if (first >= 0) {
        for (int i = 0; i < len; i++) acc += (uint32_t)(mono_at<CH>(in, first - in_base + i) * (int)__ldg(f + i));
    } else {
        for (int i = 0; i < len; i++) {                       // left of sample 0 the reference mirrors: src[|i| % src_size]
            int64_t p = first + i;
            if (p < 0) p = -p;
            acc += (uint32_t)(mono_at<CH>(in, p % n_in) * (int)__ldg(f + i));
        }
    }
    return acc;
}
template <int CH>
__global__ void __launch_bounds__(256)
k_audio_resample(const int16_t *__restrict__ in, int64_t n_in, int64_t in_base, const int16_t *__restrict__ bank, int len, int len8,
                 int64_t index0, uint64_t dst_incr /* in_rate * 1024 */, uint32_t src_incr /* out_rate */, int64_t k_base,
                 int16_t *__restrict__ out, int64_t n_out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_out; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t index = index0 + (int64_t)(((uint64_t)(k_base + k) * dst_incr) / src_incr);
        out[k] = round_sat(taps_direct<CH>(in, n_in, bank + (size_t)len8 * (size_t)(index & 1023), len, index >> 10, in_base));
    }
}

// Tiled form: a block owns 256 consecutive outputs.  Their taps cover one contiguous
// stretch of the input, which the block loads once, coalesced and already mixed down to one channel, into shared
// memory as 32-bit samples; each thread then walks its phase row 8 coefficients per 128-bit load.
constexpr int kWinMax = 6144;          // samples of shared memory (24 KB)
constexpr int kSplit = kWinMax / 2 + 16; // k_audio_resample_rows: where the odd-position half of the window starts (16 banks off)

template <int CH>
__global__ void __launch_bounds__(256)
k_audio_resample_tile(const int16_t *__restrict__ in, int64_t n_in, int64_t in_base, const int16_t *__restrict__ bank, int len, int len8,
                      int64_t index0, uint64_t dst_incr, uint32_t src_incr, uint32_t step_q, uint32_t step_r, int64_t k_base,
                      int16_t *__restrict__ out, int64_t n_out) {
    __shared__ int s_win[kWinMax];
    const int64_t k0 = (int64_t)blockIdx.x * 256, k = k0 + threadIdx.x;
    const int t_hi = (int)min((int64_t)255, n_out - 1 - k0);
    // index(k0 + t) = index(k0) + t*q + floor((rem + t*r) / S) with dst_incr = q*S + r: one 64-bit division per block,
    // 32-bit ones per thread (rem < S <= 2^21, t*r < 2^29)
    const uint64_t p0 = (uint64_t)(k_base + k0) * dst_incr, base = p0 / src_incr;
    const uint32_t rem = (uint32_t)(p0 - base * src_incr);
    const int64_t index_lo = index0 + (int64_t)base;
    const int t = min((int)threadIdx.x, t_hi);
    const int64_t index = index_lo + (int64_t)t * step_q + (rem + (uint32_t)t * step_r) / src_incr;
    const int64_t first_lo = index_lo >> 10;
    const int64_t first_hi = (index_lo + (int64_t)t_hi * step_q + (rem + (uint32_t)t_hi * step_r) / src_incr) >> 10;
    const int16_t *f = bank + (size_t)len8 * (size_t)(index & 1023);
    if (first_lo < 0) {                 // the block that holds the mirrored head of the stream
        if (k < n_out) out[k] = round_sat(taps_direct<CH>(in, n_in, f, len, index >> 10, in_base));
        return;
    }
    const int wlen = (int)(first_hi - first_lo) + len8;          // <= kWinMax (checked by the launcher)
    for (int j = threadIdx.x; j < wlen; j += 256)
        s_win[j] = first_lo - in_base + j < n_in ? mono_at<CH>(in, first_lo - in_base + j) : 0;       // only zero coefficients reach past the end
    __syncthreads();
    if (k >= n_out) return;
    const int *w = s_win + (int)((index >> 10) - first_lo);
    uint32_t acc = 0;
    for (int i = 0; i < len8; i += 8) {
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(f + i));
        acc += (uint32_t)(w[i] * (int)(int16_t)c.x) + (uint32_t)(w[i + 1] * ((int)c.x >> 16));
        acc += (uint32_t)(w[i + 2] * (int)(int16_t)c.y) + (uint32_t)(w[i + 3] * ((int)c.y >> 16));
        acc += (uint32_t)(w[i + 4] * (int)(int16_t)c.z) + (uint32_t)(w[i + 5] * ((int)c.z >> 16));
        acc += (uint32_t)(w[i + 6] * (int)(int16_t)c.w) + (uint32_t)(w[i + 7] * ((int)c.w >> 16));
    }
    out[k] = round_sat(acc);
}

// Phase-row form (the one that normally runs): the filter phase of output k repeats with a period P2 that depends on the
// rate pair only (147 outputs for 48 kHz -> 22 050 Hz, 1 for 44.1 kHz, 441 for 8 kHz).  A block of Q threads (Q = the
// multiple of P2 next to 256) owns Q x M consecutive outputs; thread t computes outputs k0 + t + j Q, j < M, which all use
// the same phase row: every 128-bit coefficient load, and the unpacking of its 8 coefficients, serves M outputs, and the
// input stretch (first(k0 + t + j Q) = first(k0 + t) + j * step exactly) is staged in shared memory once as above.
template <int CH, int M, bool SPLIT>
__global__ void __launch_bounds__(1024)
k_audio_resample_rows(const int16_t *__restrict__ in, int64_t n_in, int64_t in_base, const int16_t *__restrict__ bank, int len, int len8,
                      int64_t index0, uint64_t dst_incr, uint32_t src_incr, uint32_t step_q, uint32_t step_r, int q_step /* samples per Q outputs */,
                      int64_t k_base, int16_t *__restrict__ out, int64_t n_out) {
    // SPLIT: the staged stretch is split by sample parity (even positions in the lower half, odd ones kSplit words higher): at the
    // common 2:1 reduction neighbouring threads' windows start two samples apart, which in one linear array makes every
    // 32-bit read a 2-way bank conflict (measured: the kernel ran at the shared-memory wavefront rate); split, the threads of
    // a warp read consecutive words (1.29 -> 1.03 ms).  The launcher asks for it when the windows of neighbouring outputs start
    // an even whole number of samples apart, and for enlargements (0.87 -> 0.79 ms at 8 kHz); at 48 kHz -> 22 050 Hz (2.18
    // samples apart, parities mixed within a warp) the linear array is the faster one (1.54 against 1.81 ms)
    __shared__ int s_win[kWinMax + 64];
    const int Q = blockDim.x, t = threadIdx.x;
    const int64_t k0 = (int64_t)blockIdx.x * Q * M;
    const int64_t k_hi = min(k0 + (int64_t)Q * M, n_out) - 1;
    const uint64_t p0 = (uint64_t)(k_base + k0) * dst_incr, base = p0 / src_incr;
    const uint32_t rem = (uint32_t)(p0 - base * src_incr);
    const int64_t index_lo = index0 + (int64_t)base;
    const int64_t index = index_lo + (int64_t)t * step_q + (rem + (uint32_t)t * step_r) / src_incr;
    const int64_t first_lo = index_lo >> 10;
    const int64_t first_hi = (index0 + (int64_t)(((uint64_t)(k_base + k_hi) * dst_incr) / src_incr)) >> 10;
    const int16_t *f = bank + (size_t)len8 * (size_t)(index & 1023);
    if (first_lo < 0) {                 // the block that holds the mirrored head of the stream
        for (int j = 0; j < M; j++) {
            const int64_t k = k0 + t + (int64_t)j * Q;
            if (k < n_out) out[k] = round_sat(taps_direct<CH>(in, n_in, f, len, (index >> 10) + (int64_t)j * q_step, in_base));
        }
        return;
    }
    const int wlen = (int)(first_hi - first_lo) + len8;          // <= kWinMax (checked by the launcher)
    for (int j = t; j < wlen; j += Q)
        s_win[SPLIT ? (j & 1) * kSplit + (j >> 1) : j] = first_lo - in_base + j < n_in ? mono_at<CH>(in, first_lo - in_base + j) : 0;
    __syncthreads();
    if (k0 + t >= n_out) return;
    const int off0 = (int)((index >> 10) - first_lo);
    uint32_t acc[M];
#pragma unroll
    for (int j = 0; j < M; j++) acc[j] = 0;
    for (int i = 0; i < len8; i += 8) {
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(f + i));
        const int c0 = (int16_t)c.x, c1 = (int)c.x >> 16, c2 = (int16_t)c.y, c3 = (int)c.y >> 16;
        const int c4 = (int16_t)c.z, c5 = (int)c.z >> 16, c6 = (int16_t)c.w, c7 = (int)c.w >> 16;
#pragma unroll
        for (int j = 0; j < M; j++) {
            // taps 0, 2, 4, 6 of this group sit in the half of the window's own parity, taps 1, 3, 5, 7 in the other one
            // (outputs past n_out read staged or stale words; they are not stored)
            const int off = off0 + j * q_step + i;
            if (SPLIT) {
                const int *va = s_win + (off & 1) * kSplit + (off >> 1);
                const int *vb = s_win + ((off + 1) & 1) * kSplit + ((off + 1) >> 1);
                acc[j] += (uint32_t)(va[0] * c0) + (uint32_t)(vb[0] * c1) + (uint32_t)(va[1] * c2) + (uint32_t)(vb[1] * c3) +
                          (uint32_t)(va[2] * c4) + (uint32_t)(vb[2] * c5) + (uint32_t)(va[3] * c6) + (uint32_t)(vb[3] * c7);
            } else {
                const int *v = s_win + off;
                acc[j] += (uint32_t)(v[0] * c0) + (uint32_t)(v[1] * c1) + (uint32_t)(v[2] * c2) + (uint32_t)(v[3] * c3) +
                          (uint32_t)(v[4] * c4) + (uint32_t)(v[5] * c5) + (uint32_t)(v[6] * c6) + (uint32_t)(v[7] * c7);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < M; j++) {
        const int64_t k = k0 + t + (int64_t)j * Q;
        if (k < n_out) out[k] = round_sat(acc[j]);
    }
}

template <int CH, bool SPLIT>
static bool launch_rows_m(int m, unsigned blocks, int Q, const int16_t *in, int64_t n_in, int64_t in_base, const int16_t *bank, int len,
                          int len8, int64_t index0, uint64_t D, uint32_t S, int q_step, int64_t k_base, int16_t *out, int64_t n_out,
                          cudaStream_t s) {
    const uint32_t sq = (uint32_t)(D / S), sr = (uint32_t)(D % S);
    switch (m) {
    case 8: AMV_LAUNCH((k_audio_resample_rows<CH, 8, SPLIT>), blocks, Q, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, sq, sr, q_step, k_base, out, n_out); return true;
    case 4: AMV_LAUNCH((k_audio_resample_rows<CH, 4, SPLIT>), blocks, Q, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, sq, sr, q_step, k_base, out, n_out); return true;
    case 2: AMV_LAUNCH((k_audio_resample_rows<CH, 2, SPLIT>), blocks, Q, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, sq, sr, q_step, k_base, out, n_out); return true;
    default: return false;
    }
}
template <int CH>
static bool launch_rows(int m, unsigned blocks, int Q, const int16_t *in, int64_t n_in, int64_t in_base, const int16_t *bank, int len,
                        int len8, int64_t index0, uint64_t D, uint32_t S, int q_step, int64_t k_base, int16_t *out, int64_t n_out,
                        cudaStream_t s) {
    // parity-split window: neighbouring outputs an even whole number of samples apart, or an enlargement
    const bool split = (D % S == 0 && ((D / S) & 2047) == 0) || D < (uint64_t)S * 1024;
    return split ? launch_rows_m<CH, true>(m, blocks, Q, in, n_in, in_base, bank, len, len8, index0, D, S, q_step, k_base, out, n_out, s)
                 : launch_rows_m<CH, false>(m, blocks, Q, in, n_in, in_base, bank, len, len8, index0, D, S, q_step, k_base, out, n_out, s);
}

static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; }

void launch_audio_resample(const int16_t *in, int64_t n_in, int64_t in_base, int in_ch, const int16_t *bank, int len, int in_rate,
                           int out_rate, int64_t k_base, int16_t *out, int64_t n_out, int audio_form, cudaStream_t s) {
    if (n_out <= 0) return;
    const int len8 = (len + 7) & ~7;
    const int64_t index0 = -1024 * (int64_t)((len - 1) / 2);
    const uint64_t D = (uint64_t)in_rate * 1024;
    const uint32_t S = (uint32_t)out_rate;
    const int64_t span = (int64_t)((255 * D) / S >> 10) + 2 + len8;        // input samples 256 consecutive outputs touch
    const int64_t blocks = (n_out + 255) / 256;
    if (audio_form >= 2) {
        // phase period: after P = S / gcd(D, S) outputs the position has advanced by exactly adv = D / gcd(D, S) (no remainder);
        // the phase (position mod 1024) repeats after P2 = P * 1024 / gcd(adv, 1024) outputs
        const uint64_t g = gcd_u64(D, S), P = S / g, adv = D / g;
        const uint64_t P2 = P * (1024 / gcd_u64(adv, 1024));
        if (P2 <= 1024) {
            const int Q = (int)(P2 * ((256 + P2 - 1) / P2));                 // threads per block: a multiple of P2, >= 256
            const uint64_t q_adv = (uint64_t)(Q / (int)P) * adv;            // position advance per Q outputs: a multiple of 1024
            const int64_t q_step = (int64_t)(q_adv >> 10);
            int m = 8;
            while (m >= 2 && (int64_t)((((uint64_t)Q * m - 1) * D) / S >> 10) + 2 + len8 > kWinMax) m >>= 1;
            const int64_t nb = (n_out + (int64_t)Q * m - 1) / ((int64_t)Q * m);
            if (m >= 2 && q_step < (1 << 20) && nb <= 0x7fffffff) {
                const bool ok = in_ch == 2 ? launch_rows<2>(m, (unsigned)nb, Q, in, n_in, in_base, bank, len, len8, index0, D, S, (int)q_step, k_base, out, n_out, s)
                                           : launch_rows<1>(m, (unsigned)nb, Q, in, n_in, in_base, bank, len, len8, index0, D, S, (int)q_step, k_base, out, n_out, s);
                if (ok) return;
            }
        }
    }
    if (audio_form >= 1 && span <= kWinMax && blocks <= 0x7fffffff) {
        if (in_ch == 2) AMV_LAUNCH(k_audio_resample_tile<2>, (unsigned)blocks, 256, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, (uint32_t)(D / S), (uint32_t)(D % S), k_base, out, n_out);
        else            AMV_LAUNCH(k_audio_resample_tile<1>, (unsigned)blocks, 256, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, (uint32_t)(D / S), (uint32_t)(D % S), k_base, out, n_out);
        return;
    }
    int64_t grid = blocks;
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    if (in_ch == 2) AMV_LAUNCH(k_audio_resample<2>, (unsigned)grid, 256, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, k_base, out, n_out);
    else            AMV_LAUNCH(k_audio_resample<1>, (unsigned)grid, 256, 0, s, in, n_in, in_base, bank, len, len8, index0, D, S, k_base, out, n_out);
}

}  // namespace amv

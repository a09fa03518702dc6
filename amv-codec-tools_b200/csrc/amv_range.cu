// amv_range.cu -- the pre/post stage next to the codec: the range conversion the reference's
// ffmpeg.c inserts through img_convert between yuv420p (CCIR 601 range) and the codec's yuvj420p (full range):
// img_apply_table (imgconvert.c:1236-1260, 2492-2510) with the four 256-entry tables of img_convert_init
// (:1221-1233, colorspace.h:69-84; SCALEBITS 10).  Element-wise, HBM-bound: 16 pixels per thread through
// 128-bit loads and stores, the tables evaluated arithmetically (one multiply-add, shift, clamp per pixel).
#include "amv_common.cuh"
#include "amv_kernels.h"

namespace amv {

template <int DIR, bool CHROMA>
__device__ __forceinline__ uint32_t range_px(uint32_t p) {
    const int x = (int)p;
    if (DIR == 0) {
        const int v = CHROMA ? ((x - 128) * 1161 + (512 + (128 << 10))) >> 10      // C_CCIR_TO_JPEG
                             : (x * 1192 + (512 - 16 * 1192)) >> 10;                // Y_CCIR_TO_JPEG
        return (uint32_t)__vimin_s32_relu(v, 255);                                   // cm[] of the reference
    }
    if (CHROMA) return (uint32_t)max(16, ((x - 128) * 903 + (512 + (128 << 10))) >> 10);   // C_JPEG_TO_CCIR
    return (uint32_t)((x * 879 + (512 + (16 << 10))) >> 10);                                // Y_JPEG_TO_CCIR
}

template <int DIR, bool CHROMA>
__device__ __forceinline__ uint32_t range_word(uint32_t w) {
    return range_px<DIR, CHROMA>(w & 0xff) | (range_px<DIR, CHROMA>((w >> 8) & 0xff) << 8) |
           (range_px<DIR, CHROMA>((w >> 16) & 0xff) << 16) | (range_px<DIR, CHROMA>(w >> 24) << 24);
}

// planes are walked as rows of `width` bytes with a row pitch; rows_total = rows per frame * frames when the
// frames are contiguous in pitch units, otherwise the launcher issues one grid per frame range
template <int DIR, bool CHROMA>
__global__ void __launch_bounds__(256)
k_range(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int width, int rows, int n, int ls_in, int ls_out,
        uint64_t fs_in, uint64_t fs_out, int vec) {
    const int units = vec ? (width + 15) >> 4 : width;          // 16-byte units (aligned planes) or single bytes
    const int64_t total = (int64_t)units * rows * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ux = (int)(i % units);
        const int64_t ry = i / units;
        const int r = (int)(ry % rows), f = (int)(ry / rows);
        const uint8_t *s = src + (uint64_t)f * fs_in + (int64_t)r * ls_in;
        uint8_t *d = dst + (uint64_t)f * fs_out + (int64_t)r * ls_out;
        if (vec) {
            if (ux * 16 + 16 <= width) {
                uint4 q = *reinterpret_cast<const uint4 *>(s + ux * 16);
                q.x = range_word<DIR, CHROMA>(q.x); q.y = range_word<DIR, CHROMA>(q.y);
                q.z = range_word<DIR, CHROMA>(q.z); q.w = range_word<DIR, CHROMA>(q.w);
                *reinterpret_cast<uint4 *>(d + ux * 16) = q;
            } else {
                for (int x = ux * 16; x < width; x++) d[x] = (uint8_t)range_px<DIR, CHROMA>(s[x]);
            }
        } else {
            d[ux] = (uint8_t)range_px<DIR, CHROMA>(s[ux]);
        }
    }
}

// tight planes (pitch = width, frame stride = width * rows, 16-byte aligned ends): the batch is one linear array and the
// index needs no division into (frame, row, column)
template <int DIR, bool CHROMA>
__global__ void __launch_bounds__(256)
k_range_flat(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int64_t units) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 q = src[i];
        q.x = range_word<DIR, CHROMA>(q.x); q.y = range_word<DIR, CHROMA>(q.y);
        q.z = range_word<DIR, CHROMA>(q.z); q.w = range_word<DIR, CHROMA>(q.w);
        dst[i] = q;
    }
}

template <int DIR, bool CHROMA>
static void launch_plane(const uint8_t *src, uint8_t *dst, int width, int rows, int n, int ls_in, int ls_out, uint64_t fs_in,
                         uint64_t fs_out, cudaStream_t s) {
    const uint64_t bytes = (uint64_t)width * rows * n;
    if (ls_in == width && ls_out == width && fs_in == (uint64_t)width * rows && fs_out == fs_in && (bytes & 15) == 0 &&
        ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0)) {
        const int64_t units = (int64_t)(bytes >> 4);
        int64_t grid = (units + 255) / 256;
        if (grid > kNumSMs * 32) grid = kNumSMs * 32;
        if (grid < 1) grid = 1;
        AMV_LAUNCH((k_range_flat<DIR, CHROMA>), (unsigned)grid, 256, 0, s, reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), units);
        return;
    }
    const bool vec = ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)ls_in | (uintptr_t)ls_out | fs_in | fs_out) & 15) == 0);
    const int64_t total = (int64_t)(vec ? (width + 15) >> 4 : width) * rows * n;
    int64_t grid = (total + 255) / 256;
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    if (grid < 1) grid = 1;
    AMV_LAUNCH((k_range<DIR, CHROMA>), (unsigned)grid, 256, 0, s, src, dst, width, rows, n, ls_in, ls_out, fs_in, fs_out, vec ? 1 : 0);
}

// one plane of `width` x `rows` bytes (chroma: the two chroma tables)
void launch_convert_range_plane(const uint8_t *src, uint8_t *dst, int width, int rows, int n, int ls_in, int ls_out, uint64_t fs_in,
                                uint64_t fs_out, int dir, bool chroma, cudaStream_t s) {
    if (width <= 0 || rows <= 0 || n <= 0) return;
    if (dir == 0) { if (chroma) launch_plane<0, true>(src, dst, width, rows, n, ls_in, ls_out, fs_in, fs_out, s);
                    else        launch_plane<0, false>(src, dst, width, rows, n, ls_in, ls_out, fs_in, fs_out, s); }
    else          { if (chroma) launch_plane<1, true>(src, dst, width, rows, n, ls_in, ls_out, fs_in, fs_out, s);
                    else        launch_plane<1, false>(src, dst, width, rows, n, ls_in, ls_out, fs_in, fs_out, s); }
}

void launch_convert_range(const uint8_t *y, const uint8_t *u, const uint8_t *v, uint8_t *oy, uint8_t *ou, uint8_t *ov, int n,
                          int w, int h, int ls_y, int ls_c, uint64_t fs_y, uint64_t fs_c, int ols_y, int ols_c, uint64_t ofs_y,
                          uint64_t ofs_c, int dir, cudaStream_t s) {
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    if (dir == 0) {
        launch_plane<0, false>(y, oy, w, h, n, ls_y, ols_y, fs_y, ofs_y, s);
        launch_plane<0, true>(u, ou, cw, ch, n, ls_c, ols_c, fs_c, ofs_c, s);
        launch_plane<0, true>(v, ov, cw, ch, n, ls_c, ols_c, fs_c, ofs_c, s);
    } else {
        launch_plane<1, false>(y, oy, w, h, n, ls_y, ols_y, fs_y, ofs_y, s);
        launch_plane<1, true>(u, ou, cw, ch, n, ls_c, ols_c, fs_c, ofs_c, s);
        launch_plane<1, true>(v, ov, cw, ch, n, ls_c, ols_c, fs_c, ofs_c, s);
    }
}

}  // namespace amv

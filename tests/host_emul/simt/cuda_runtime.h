// TEST-ONLY stand-in for <cuda_runtime.h>: lets the kernel sources of libamvcuda (csrc/*.cu) compile as plain
// C++ (g++ -DAMV_EMUL) and run on the CPU, one CTA at a time, every CUDA thread a cooperative fiber that yields at
// warp / CTA collectives (simt_rt.cpp).  "Device memory" is host memory, streams and events are no-ops.  The point is
// to run the *whole* kernels -- shuffles, ballots, shared-memory bit packing, the launch geometry -- against the
// oracle in a container without a GPU, before GPU minutes are spent.  This is not a product path: nothing outside
// tests/ builds or loads it, the package never looks for it, and libamvcuda.so has no CPU fallback.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#ifndef AMV_EMUL
#error "tests/host_emul/simt/cuda_runtime.h is the emulator's header; build with -DAMV_EMUL"
#endif

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

// ---------------------------------------------------------------- vector types
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{ x, y }; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }
static inline int2 make_int2(int x, int y) { return int2{ x, y }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }

// ---------------------------------------------------------------- SIMT runtime
namespace simt {
struct Idx { unsigned x, y, z; };
struct ThreadCtx { Idx tid, bid, bdim, gdim; int lane, wid; };
ThreadCtx *cur();
uint8_t *dyn_smem();                      // the CTA's dynamic shared memory (256 KB, 4 KB aligned)
uintptr_t smem_base();                    // origin of the 32-bit "shared window" addresses
void launch_impl(dim3 grid, dim3 block, size_t smem, void (*tramp)(void *), void *closure);
template <class F> void launch(dim3 grid, dim3 block, size_t smem, F &&f) {
    launch_impl(grid, block, smem, [](void *p) { (*static_cast<typename std::remove_reference<F>::type *>(p))(); }, &f);
}
// collectives (all yield until the participating lanes have arrived)
void sync_warp(uint32_t mask);
void sync_block();
int sync_block_or(int pred);
uint64_t shfl(uint32_t mask, uint64_t v, int src_lane);       // value of src_lane (own value if src is out of range)
uint32_t ballot(uint32_t mask, int pred);
uint32_t active_mask();
}  // namespace simt

#define threadIdx (simt::cur()->tid)
#define blockIdx (simt::cur()->bid)
#define blockDim (simt::cur()->bdim)
#define gridDim (simt::cur()->gdim)
#define warpSize 32

#define AMV_EMUL_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })

static inline void __syncthreads() { simt::sync_block(); }
static inline int __syncthreads_or(int p) { return simt::sync_block_or(p); }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { simt::sync_warp(mask); }
static inline uint32_t __activemask() { return simt::active_mask(); }
static inline uint32_t __ballot_sync(uint32_t m, int p) { return simt::ballot(m, p); }
static inline int __any_sync(uint32_t m, int p) { return simt::ballot(m, p) != 0; }
static inline int __all_sync(uint32_t m, int p) { return simt::ballot(m, !p) == 0; }

namespace simt {
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}
template <class T> static inline T __shfl_sync(uint32_t m, T v, int src, int width = 32) {
    const int lane = simt::cur()->lane;
    const int base = lane & ~(width - 1);
    return simt::from_bits<T>(simt::shfl(m, simt::to_bits(v), base + (src & (width - 1))));
}
template <class T> static inline T __shfl_up_sync(uint32_t m, T v, unsigned d, int width = 32) {
    const int lane = simt::cur()->lane;
    const int src = lane - (int)d;
    return simt::from_bits<T>(simt::shfl(m, simt::to_bits(v), src < (lane & ~(width - 1)) ? lane : src));
}
template <class T> static inline T __shfl_down_sync(uint32_t m, T v, unsigned d, int width = 32) {
    const int lane = simt::cur()->lane;
    const int src = lane + (int)d;
    return simt::from_bits<T>(simt::shfl(m, simt::to_bits(v), src > (lane | (width - 1)) ? lane : src));
}
template <class T> static inline T __shfl_xor_sync(uint32_t m, T v, int x, int width = 32) {
    const int lane = simt::cur()->lane;
    (void)width;
    return simt::from_bits<T>(simt::shfl(m, simt::to_bits(v), lane ^ x));
}

// ---------------------------------------------------------------- integer intrinsics
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1); v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4); return __builtin_bswap32(v);
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((int64_t)a * b) >> 32); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {      // ((hi:lo) << (s & 31)) >> 32
    s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {      // ((hi:lo) >> (s & 31)) & 0xffffffff
    s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { return s >= 32 ? lo : __funnelshift_l(lo, hi, s); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { return s >= 32 ? hi : __funnelshift_r(lo, hi, s); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline int __vimin_s32_relu(int a, int b) { const int m = a < b ? a : b; return m < 0 ? 0 : m; }
static inline int __vimax_s32_relu(int a, int b) { const int m = a > b ? a : b; return m < 0 ? 0 : m; }
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;

template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---------------------------------------------------------------- host runtime (device memory = host memory)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct simt_stream_ *cudaStream_t;
typedef struct simt_event_ *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostRegister(void *p, size_t n, unsigned flags);
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaPeekAtLastError();
cudaError_t cudaGetLastError();
cudaError_t cudaDeviceSynchronize();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p);
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(T &sym, const void *src, size_t n, size_t off, cudaMemcpyKind, cudaStream_t) {
    memcpy(reinterpret_cast<char *>(&sym) + off, src, n);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaGetSymbolAddress(void **p, T &sym) { *p = &sym; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

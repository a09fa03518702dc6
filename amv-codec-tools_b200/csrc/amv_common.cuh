// amv_common.cuh -- shared definitions for the libamvcuda kernels (sm_100a).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define AMV_HD __host__ __device__ __forceinline__
#define AMV_D  __device__ __forceinline__
#else
#define AMV_HD inline
#define AMV_D  inline
#endif

// Kernel launch and dynamic shared memory go through two macros so that the test suite's SIMT emulator can compile the
// very same kernel sources as plain C++ and run them on the CPU (-DAMV_EMUL: a test build only, with its own stand-in
// for the CUDA runtime header; the product is nvcc's build, which never defines it and has no CPU path).
#if defined(AMV_EMUL)
#define AMV_LAUNCH(kernel, grid, block, smem, stream, ...) AMV_EMUL_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#define AMV_EXTERN_SHARED(type, name, align) type *name = reinterpret_cast<type *>(simt::dyn_smem())
#else
#define AMV_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define AMV_EXTERN_SHARED(type, name, align) extern __shared__ __align__(align) type name[]
#endif

// per-unit status bits: keep in sync with include/amvcuda.h
#define AMV_ST_SHORT    (1 << 0)
#define AMV_ST_BADCODE  (1 << 1)
#define AMV_ST_COEFIDX  (1 << 2)
#define AMV_ST_MARKER   (1 << 3)
#define AMV_ST_OVERRUN  (1 << 4)
#define AMV_ST_RANGE    (1 << 5)
#define AMV_ST_NOSPACE  (1 << 6)
#define AMV_ST_HEADER   (1 << 7)

namespace amv {

constexpr int kNumSMs = 148;              // B200

// Every frame's un-stuffed scan lives in a 16-byte aligned scratch slot of align16(packet size) +
// kSlotPad bytes (zero padded past the data); its token region holds 4 tokens per slot byte, which
// leaves 4 * kSlotPad = 640 tokens beyond the 1-token-per-2-bits bound: 16 per lane (32 lanes at most) for the
// 16-byte alignment of every lane's first and last token group, plus the room the token kernels keep at the tail.
constexpr uint32_t kSlotPad = 160;

// Picture geometry shared by encoder and decoder kernels.
struct Geom {
    int w, h;          // luma size
    int cw, ch;        // chroma plane size, ceil(w/2) x ceil(h/2)
    int mbw, mbh;      // MCU grid, ceil(w / (8 lh)) x ceil(h / (8 lv)): 16x16 macroblocks for 4:2:0
    int nblk;          // nb * mbw * mbh
    // sampling: blocks per MCU across / down as log2 -- luma (llh, llv), each chroma component (lch, lcv) --
    // and the counts derived from them.  AMV / SP5X / 4:2:0 JPEG: 2x2 luma, 1x1 chroma, 6 blocks per MCU
    int llh, llv, lch, lcv;
    int nl, nc, nb;    // luma blocks, blocks of one chroma component, all blocks of an MCU
    int y0, c0;        // first (bottom-most stored) row of luma / chroma in flipped order
    int flip;          // 1: rows are addressed bottom-up from y0 / c0 (AMV); 0: top-down (SP5X)
};

// Row the codec starts from before walking upwards: vs*(8*mb_h - ((h/2)&7)) - 1
// (mjpegdec.c:672-677 and mjpegenc.c:467-470 use the same expression).
AMV_HD int flip_start_row(int h, int vs) {
    const int mbh = (h + 15) >> 4;
    return vs * (8 * mbh - ((h >> 1) & 7)) - 1;
}

// sampling factors as log2: (1,1,0,0) = 4:2:0, (1,0,0,0) = 4:2:2 as 2x1 / 1x1, (1,1,0,1) = 4:2:2 as 2x2 / 1x2 (what the
// reference's mjpeg_encoder writes), (0,0,0,0) = 4:4:4
AMV_HD Geom make_geom_sampled(int w, int h, int llh, int llv, int lch, int lcv) {
    Geom g;
    g.w = w; g.h = h;
    g.llh = llh; g.llv = llv; g.lch = lch; g.lcv = lcv;
    g.nl = 1 << (llh + llv); g.nc = 1 << (lch + lcv); g.nb = g.nl + 2 * g.nc;
    g.cw = ((w << lch) + (1 << llh) - 1) >> llh; g.ch = ((h << lcv) + (1 << llv) - 1) >> llv;
    g.mbw = (w + (8 << llh) - 1) / (8 << llh); g.mbh = (h + (8 << llv) - 1) / (8 << llv);
    g.nblk = g.nb * g.mbw * g.mbh;
    g.y0 = flip_start_row(h, 2);
    g.c0 = flip_start_row(h, 1);
    g.flip = 1;
    return g;
}
AMV_HD Geom make_geom(int w, int h) { return make_geom_sampled(w, h, 1, 1, 0, 0); }

AMV_HD uint32_t bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}

// (a & mask) | (b & ~mask): one LOP3 (written out because the compiler narrows b's mask first when only a halfword is kept)
AMV_HD uint32_t bitselect(uint32_t a, uint32_t b, uint32_t mask) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
#else
    return (a & mask) | (b & ~mask);
#endif
}

// (a * b >> 32) + c, unsigned: with b = 2^(32 - s) this is (a >> s) + c in one multiply-add
AMV_HD uint32_t mad_hi_u32(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return (uint32_t)(((uint64_t)a * b) >> 32) + c;
#endif
}
// a * b + c mod 2^32.  On the device an opaque multiply-add, so that a chain written as a chain stays one (the compiler
// would otherwise be free to re-associate "2a - (a + b)" back into a second chain of products).
AMV_HD int mad_lo(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return (int)((uint32_t)a * (uint32_t)b + (uint32_t)c);
#endif
}

// the low 12 bits of v, sign-extended: one SGXT (the compiler's shift pair is two instructions)
AMV_HD int sext12(uint32_t v) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("bfe.s32 %0, %1, 0, 12;" : "=r"(d) : "r"(v));
    return d;
#else
    return (int)(v << 20) >> 20;
#endif
}

// bits [15:8] of v: one PRMT
AMV_HD uint32_t byte1(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x4441);
#else
    return (v >> 8) & 0xffu;
#endif
}

AMV_HD int clamp_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// [off, off + len) lies inside [0, total): written so that a hostile offset near 2^64 cannot wrap the sum
AMV_HD bool range_ok(uint64_t off, uint64_t len, uint64_t total) { return off <= total && len <= total - off; }

// index of the most significant set bit (v != 0): one FLO
AMV_HD int msb_index(uint32_t v) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
#else
    return 31 - __builtin_clz(v);
#endif
}

// floor(x / d) and the remainder for any 32-bit x by one high multiply and one correction step: with m = floor(2^32 / d)
// (div_magic; 2^32 - 1 for d = 1) the estimate umulhi(x, m) is the quotient or one below it.  Replaces the ~25-instruction
// emulated division where the divisor is a launch constant.
inline uint32_t div_magic(uint32_t d) { return d >= 2 ? (uint32_t)((1ull << 32) / d) : 0xffffffffu; }
AMV_HD uint32_t div_by_magic(uint32_t x, uint32_t d, uint32_t m, uint32_t &rem) {
#if defined(__CUDA_ARCH__)
    uint32_t q = __umulhi(x, m);
#else
    uint32_t q = (uint32_t)(((uint64_t)x * m) >> 32);
#endif
    uint32_t r = x - q * d;
    if (r >= d) { q++; r -= d; }
    rem = r;
    return q;
}

// count of leading zeros (v != 0): one FLO.SH
AMV_HD int clz_nz(uint32_t v) {
#if defined(__CUDA_ARCH__)
    int r;
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
#else
    return __builtin_clz(v);
#endif
}

#if defined(AMV_EMUL)
// emulator: "shared-window addresses" are offsets from a fixed origin below the emulator's static storage
inline uint32_t smem_addr(const void *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) - simt::smem_base();
    if (a >= (1u << 24)) abort();            // not a __shared__ object (or the emulator's window is misplaced)
    return (uint32_t)a;
}
template <class T> inline T *smem_ptr(uint32_t saddr) { return reinterpret_cast<T *>(simt::smem_base() + saddr); }
inline uint32_t lds32(uint32_t saddr) { return *smem_ptr<uint32_t>(saddr); }
inline int lds_s16(uint32_t saddr) { return *smem_ptr<int16_t>(saddr); }
inline uint32_t lds_u16(uint32_t saddr) { return *smem_ptr<uint16_t>(saddr); }
inline uint4 lds128(uint32_t saddr) { return *smem_ptr<uint4>(saddr); }
template <int OFF> inline uint32_t lds32_at(uint32_t saddr) { return *smem_ptr<uint32_t>(saddr + OFF); }
inline uint2 lds64(uint32_t saddr) { return *smem_ptr<uint2>(saddr); }
inline void sts32(uint32_t saddr, uint32_t v) { *smem_ptr<uint32_t>(saddr) = v; }
inline void sts16(uint32_t saddr, uint32_t v) { *smem_ptr<uint16_t>(saddr) = (uint16_t)v; }
inline void sts8(uint32_t saddr, uint32_t v) { *smem_ptr<uint8_t>(saddr) = (uint8_t)v; }
inline void sts128(uint32_t saddr, const uint4 &v) { *smem_ptr<uint4>(saddr) = v; }
inline void red_or_shared(uint32_t saddr, uint32_t v) { *smem_ptr<uint32_t>(saddr) |= v; }
// 1-D bulk asynchronous copy + mbarrier: the emulator copies at once, so every wait is already satisfied
inline void mbar_init(uint32_t, uint32_t) {}
inline void mbar_arrive_expect_tx(uint32_t, uint32_t) {}
inline void bulk_g2s(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t) { memcpy(smem_ptr<uint8_t>(dst_saddr), src, bytes); }
inline void mbar_wait(uint32_t, uint32_t) {}
inline void fence_proxy_async() {}
inline void cp_async16(uint32_t dst_saddr, const void *src) { memcpy(smem_ptr<uint8_t>(dst_saddr), src, 16); }
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
#elif defined(__CUDACC__)
// Shared memory through explicit 32-bit shared-window addresses: keeps the hot loops free of
// generic-address arithmetic (ptxas otherwise re-derives the shared window base inside them).
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    // opaque to the optimiser: otherwise ptxas re-derives the shared window base (S2R SR_CgaCtaId +
    // LEA, a ~20-cycle special-register read) at every use inside the hot loops
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ int lds_s16(uint32_t saddr) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_shared(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
// ld.shared.u32 [saddr + OFF]: the constant rides in the instruction's immediate field
template <int OFF>
__device__ __forceinline__ uint32_t lds32_at(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(saddr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
// 1-D bulk asynchronous copy global -> shared (the TMA unit's non-tensor form: SASS UBLKCP) completing on an mbarrier.
// dst, src and bytes are multiples of 16.
__device__ __forceinline__ void mbar_init(uint32_t bar_saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t bar_saddr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_saddr, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(bar_saddr), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before its later async-proxy (bulk copy) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// per-thread asynchronous 16-byte copies global -> shared (SASS LDGSTS), grouped and waited for by group count
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
#endif

}  // namespace amv

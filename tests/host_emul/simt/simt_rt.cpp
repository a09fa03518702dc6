// TEST-ONLY: run-time of the CPU SIMT emulator (see cuda_runtime.h in this directory).  One CTA runs at a time; each
// of its threads is a ucontext fiber.  A fiber runs until it reaches a collective (__syncthreads, __syncwarp,
// __shfl*_sync, __ballot/__any/__all_sync), where it parks until every live participant has arrived; the scheduler
// walks the fibers in lane order (or in a seeded random order, SIMT_SHUFFLE=<seed>, to shake out missing
// synchronisation).  A pass in which no fiber can run is a deadlock: the emulator says where and aborts.
#include "cuda_runtime.h"
#include <stdio.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <map>
#include <mutex>
#include <vector>

namespace simt {

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Group {                 // one rendezvous point: a warp (per participation mask) or the CTA
    int arrived = 0;
    unsigned gen = 0;
    uint64_t xchg[2][32];
    uint32_t pred[2] = { 0, 0 };
    int block_or[2] = { 0, 0 };
};

struct Warp {
    Group full;                          // mask == all live lanes
    std::map<uint32_t, Group> partial;   // explicit sub-masks
    uint32_t live = 0;                   // lanes that have not returned
};

struct Fiber {
    ucontext_t uc;
    ThreadCtx tc;
    bool done = true;
    Group *blocked_on = nullptr;
    unsigned blocked_gen = 0;
    Warp *warp = nullptr;
    char *stack = nullptr;
};

struct Block {
    Fiber fib[kMaxThreads];
    Warp warps[kMaxThreads / 32];
    Group cta;
    int nthreads = 0, live = 0;
    ucontext_t sched;
    Fiber *running = nullptr;
    void (*tramp)(void *) = nullptr;
    void *closure = nullptr;
};

Block *g_blk = nullptr;
alignas(4096) uint8_t g_dyn_smem[256 * 1024];
std::mutex g_launch_mutex;               // one CTA at a time, whoever calls

void yield_to_scheduler() { swapcontext(&g_blk->running->uc, &g_blk->sched); }

void release(Group &g) { g.arrived = 0; g.gen++; }

int popcount(uint32_t v) { return __builtin_popcount(v); }

// participants of a warp-level collective: the lanes of `mask` that are still alive
Group &warp_group(Warp &w, uint32_t mask, int *need) {
    const uint32_t m = mask & w.live;
    *need = popcount(m);
    if (m == w.live) return w.full;
    return w.partial[m];
}

void rendezvous(Group &g, int need) {
    Fiber *f = g_blk->running;
    const unsigned gen = g.gen;
    if (++g.arrived >= need) { release(g); return; }
    f->blocked_on = &g; f->blocked_gen = gen;
    while (g.gen == gen) yield_to_scheduler();
    f->blocked_on = nullptr;
}

void fiber_main() {
    Block *b = g_blk;
    Fiber *f = b->running;
    b->tramp(b->closure);
    // thread exit: it no longer takes part in collectives; release whoever was waiting only for it
    f->done = true;
    Warp &w = *f->warp;
    w.live &= ~(1u << f->tc.lane);
    b->live--;
    if (w.live && w.full.arrived >= popcount(w.live)) release(w.full);
    if (b->live && b->cta.arrived >= b->live) release(b->cta);
    swapcontext(&f->uc, &b->sched);
}

}  // namespace

ThreadCtx *cur() { return &g_blk->running->tc; }
uint8_t *dyn_smem() { return g_dyn_smem; }
uintptr_t smem_base() {
    // real shared-window addresses are below 2^18 and the kernels pack them into 24-bit fields: keep every
    // __shared__ object (the dynamic buffer and the kernels' function-local statics, all in this library's .bss)
    // within 16 MB above the origin
    static const uintptr_t base = (reinterpret_cast<uintptr_t>(g_dyn_smem) - (6u << 20)) & ~uintptr_t(0xffff);
    return base;
}

void sync_warp(uint32_t mask) {
    int need;
    Group &g = warp_group(*g_blk->running->warp, mask, &need);
    rendezvous(g, need);
}

void sync_block() { rendezvous(g_blk->cta, g_blk->live); }

int sync_block_or(int pred) {
    Group &g = g_blk->cta;
    const int slot = g.gen & 1;
    if (g.arrived == 0) g.block_or[slot] = 0;
    g.block_or[slot] |= pred != 0;
    rendezvous(g, g_blk->live);
    return g.block_or[slot];
}

uint64_t shfl(uint32_t mask, uint64_t v, int src) {
    int need;
    Fiber *f = g_blk->running;
    Group &g = warp_group(*f->warp, mask, &need);
    const int slot = g.gen & 1;
    g.xchg[slot][f->tc.lane] = v;
    rendezvous(g, need);
    if (src < 0 || src > 31) src = f->tc.lane;
    return g.xchg[slot][src];
}

uint32_t ballot(uint32_t mask, int pred) {
    int need;
    Fiber *f = g_blk->running;
    Group &g = warp_group(*f->warp, mask, &need);
    const int slot = g.gen & 1;
    if (g.arrived == 0) g.pred[slot] = 0;
    if (pred) g.pred[slot] |= 1u << f->tc.lane;
    rendezvous(g, need);
    return g.pred[slot];
}

uint32_t active_mask() { return g_blk->running->warp->live; }

static void on_segv(int sig) {
    char buf[256];
    int n = 0;
    if (g_blk && g_blk->running)
        n = snprintf(buf, sizeof buf, "simt: signal %d in block %u thread %u\n", sig, g_blk->running->tc.bid.x, g_blk->running->tc.tid.x);
    else n = snprintf(buf, sizeof buf, "simt: signal %d outside a kernel\n", sig);
    (void)!write(2, buf, n);
    void *bt[48];
    backtrace_symbols_fd(bt, backtrace(bt, 48), 2);
    _exit(139);
}

void launch_impl(dim3 grid, dim3 block, size_t smem, void (*tramp)(void *), void *closure) {
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    static const bool trace = getenv("SIMT_TRACE") != nullptr;
    if (trace) {
        static bool hooked = false;
        if (!hooked) { hooked = true; static char alt[65536]; stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof alt; ss.ss_flags = 0; sigaltstack(&ss, nullptr);
                       struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_handler = on_segv; sa.sa_flags = SA_ONSTACK; sigaction(SIGSEGV, &sa, nullptr); sigaction(SIGBUS, &sa, nullptr); }
        fprintf(stderr, "simt: launch grid %u x %u x %u, block %u, smem %zu\n", grid.x, grid.y, grid.z, block.x, smem);
    }
    const int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || nt > kMaxThreads || smem > sizeof(g_dyn_smem)) { fprintf(stderr, "simt: bad launch (%d threads, %zu B smem)\n", nt, smem); abort(); }
    static Block *blk = nullptr;
    if (!blk) {
        blk = new Block();
        for (int i = 0; i < kMaxThreads; i++) blk->fib[i].stack = nullptr;
    }
    static const char *shuffle_env = getenv("SIMT_SHUFFLE");
    static uint64_t rng = shuffle_env ? strtoull(shuffle_env, nullptr, 10) * 0x9E3779B97F4A7C15ull + 1 : 0;
    g_blk = blk;
    blk->tramp = tramp; blk->closure = closure; blk->nthreads = nt;
    std::vector<int> order(nt);
    const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
    for (uint64_t bi = 0; bi < nblocks; bi++) {
        blk->live = nt;
        blk->cta = Group();
        for (int w = 0; w < (nt + 31) / 32; w++) { blk->warps[w].full = Group(); blk->warps[w].partial.clear(); blk->warps[w].live = 0; }
        for (int t = 0; t < nt; t++) {
            Fiber &f = blk->fib[t];
            if (!f.stack) {
                f.stack = static_cast<char *>(mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
                if (f.stack == MAP_FAILED) { perror("simt: mmap"); abort(); }
            }
            f.done = false; f.blocked_on = nullptr;
            f.tc.tid = Idx{ (unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y) };
            f.tc.bid = Idx{ (unsigned)(bi % grid.x), (unsigned)((bi / grid.x) % grid.y), (unsigned)(bi / ((uint64_t)grid.x * grid.y)) };
            f.tc.bdim = Idx{ block.x, block.y, block.z };
            f.tc.gdim = Idx{ grid.x, grid.y, grid.z };
            f.tc.lane = t & 31; f.tc.wid = t >> 5;
            f.warp = &blk->warps[t >> 5];
            f.warp->live |= 1u << (t & 31);
            getcontext(&f.uc);
            f.uc.uc_stack.ss_sp = f.stack; f.uc.uc_stack.ss_size = kStackBytes; f.uc.uc_link = nullptr;
            makecontext(&f.uc, fiber_main, 0);
            order[t] = t;
        }
        while (blk->live > 0) {
            bool progressed = false;
            if (rng) for (int i = nt - 1; i > 0; i--) { rng = rng * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[i], order[(rng >> 33) % (i + 1)]); }
            for (int i = 0; i < nt; i++) {
                Fiber &f = blk->fib[order[i]];
                if (f.done) continue;
                if (f.blocked_on && f.blocked_on->gen == f.blocked_gen) continue;
                progressed = true;
                blk->running = &f;
                swapcontext(&blk->sched, &f.uc);
            }
            if (!progressed) {
                fprintf(stderr, "simt: deadlock in block %llu: %d live threads, none runnable (divergent collective?)\n", (unsigned long long)bi, blk->live);
                for (int t = 0; t < nt; t++) if (!blk->fib[t].done) { fprintf(stderr, "  first stuck thread %d (warp %d lane %d), warp live mask %08x, arrived %d\n", t, t >> 5, t & 31, blk->fib[t].warp->live, blk->fib[t].blocked_on ? blk->fib[t].blocked_on->arrived : -1); break; }
                abort();
            }
        }
    }
    g_blk = nullptr;
}

}  // namespace simt

// ---------------------------------------------------------------- host runtime
namespace {
std::mutex g_mem_mutex;
std::map<uintptr_t, std::pair<size_t, int>> g_allocs;      // base -> (bytes, cudaMemoryType)
void *alloc_tracked(size_t n, int type) {
    void *p = nullptr;
    if (posix_memalign(&p, 4096, n ? n : 1) != 0) return nullptr;
    memset(p, 0xA5, n);                                      // device memory is not zeroed
    std::lock_guard<std::mutex> l(g_mem_mutex);
    g_allocs[reinterpret_cast<uintptr_t>(p)] = { n, type };
    return p;
}
void free_tracked(void *p) {
    if (!p) return;
    { std::lock_guard<std::mutex> l(g_mem_mutex); g_allocs.erase(reinterpret_cast<uintptr_t>(p)); }
    free(p);
}
}  // namespace

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "SIMT emulator (tests only)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->totalGlobalMem = 8ull << 30;
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = alloc_tracked(n, cudaMemoryTypeDevice); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free_tracked(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = alloc_tracked(n, cudaMemoryTypeHost); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free_tracked(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *p, size_t n, unsigned) {
    std::lock_guard<std::mutex> l(g_mem_mutex);
    g_allocs[reinterpret_cast<uintptr_t>(p)] = { n, cudaMemoryTypeHost };
    return cudaSuccess;
}
cudaError_t cudaHostUnregister(void *p) { std::lock_guard<std::mutex> l(g_mem_mutex); g_allocs.erase(reinterpret_cast<uintptr_t>(p)); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t i = 0; i < h; i++) memmove(static_cast<char *>(d) + i * dp, static_cast<const char *>(s) + i * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    std::lock_guard<std::mutex> l(g_mem_mutex);
    const uintptr_t x = reinterpret_cast<uintptr_t>(p);
    a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = const_cast<void *>(p);
    auto it = g_allocs.upper_bound(x);
    if (it != g_allocs.begin()) {
        --it;
        if (x < it->first + it->second.first) { a->type = (cudaMemoryType)it->second.second; a->devicePointer = const_cast<void *>(p); }
    }
    return cudaSuccess;
}

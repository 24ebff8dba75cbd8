// cuda_on_cpu/cuda_runtime.h -- TEST INFRASTRUCTURE: just enough of the CUDA programming model to compile this
// repository's kernels with g++ and run them on the host, every CUDA thread a cooperatively scheduled fiber
// (ucontext), warp collectives and __syncthreads() real rendezvous points.  It exists so that kernel SOURCE that
// has never met a GPU (the experimental cuts written after round 1's GPU budget was spent) is executed, thread by
// thread and with its real index arithmetic, against the oracle in the CPU test-suite.
//
// Not modelled: timing, memory spaces (everything is host memory), FMA contraction (g++ runs with
// -ffp-contract=off, nvcc fuses a * b + c in scalar code), inline PTX (the kernels keep their host fall-backs
// behind #ifdef __CUDA_ARCH__).  Results therefore agree with the device's to rounding, not to the bit.
// One CTA runs at a time; a grid is a loop over CTAs.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
typedef void *cudaStream_t;

namespace emu {
struct Fiber {
    ucontext_t ctx;
    bool done = false;
};
constexpr size_t STACK_BYTES = 128 * 1024;  // per CUDA thread; the kernels keep their big arrays in __shared__
inline std::vector<char> &stack_pool() { static std::vector<char> p; return p; }
inline std::vector<unsigned char> &dyn_smem_buf() { static std::vector<unsigned char> b; return b; }
inline unsigned char *dynamic_smem() { return dyn_smem_buf().data(); }  // extern __shared__ of the running CTA
struct Cta {
    unsigned n = 0;
    std::vector<Fiber> f;
    unsigned cur = 0;
    ucontext_t sched;
    std::vector<int> warp_arrived, warp_alive;
    std::vector<unsigned> warp_gen;
    int cta_arrived = 0, cta_alive = 0;
    unsigned cta_gen = 0;
    std::vector<uint64_t> slots;  // one 64-bit exchange slot per thread
    std::function<void()> body;
};
inline Cta *&cta() {
    static Cta *c = nullptr;
    return c;
}
inline dim3 &tid() { static dim3 v; return v; }
inline dim3 &bid() { static dim3 v; return v; }
inline dim3 &bdim() { static dim3 v; return v; }
inline dim3 &gdim() { static dim3 v; return v; }

inline void yield() {
    Cta *c = cta();
    const unsigned me = c->cur;
    swapcontext(&c->f[me].ctx, &c->sched);
    tid().x = me;  // restored by the scheduler as well; kept here for clarity
}
inline void warp_barrier() {
    Cta *c = cta();
    const unsigned w = c->cur >> 5;
    const unsigned g = c->warp_gen[w];
    if (++c->warp_arrived[w] >= c->warp_alive[w]) {
        c->warp_arrived[w] = 0;
        c->warp_gen[w]++;
    } else {
        while (c->warp_gen[w] == g) yield();
    }
}
inline void cta_barrier() {
    Cta *c = cta();
    const unsigned g = c->cta_gen;
    if (++c->cta_arrived >= c->cta_alive) {
        c->cta_arrived = 0;
        c->cta_gen++;
    } else {
        while (c->cta_gen == g) yield();
    }
}
inline void trampoline() {
    Cta *c = cta();
    const unsigned me = c->cur;
    c->body();
    c = cta();
    c->f[me].done = true;
    // a finished thread no longer takes part in rendezvous; release one that was only waiting for it
    const unsigned w = me >> 5;
    if (--c->warp_alive[w] > 0 && c->warp_arrived[w] >= c->warp_alive[w]) {
        c->warp_arrived[w] = 0;
        c->warp_gen[w]++;
    }
    if (--c->cta_alive > 0 && c->cta_arrived >= c->cta_alive) {
        c->cta_arrived = 0;
        c->cta_gen++;
    }
    swapcontext(&c->f[me].ctx, &c->sched);
}
// kernel<<<grid, block, dyn_smem>>>(args...)  ==  emu::launch(grid, block, [&] { kernel(args...); }, dyn_smem);
inline void launch(dim3 grid, unsigned block, std::function<void()> body, size_t dyn_smem = 0) {
    gdim() = grid;
    bdim().x = block;
    if (stack_pool().size() < (size_t)block * STACK_BYTES) stack_pool().resize((size_t)block * STACK_BYTES);
    if (dyn_smem_buf().size() < dyn_smem + 16) dyn_smem_buf().resize(dyn_smem + 16);
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned b = 0; b < grid.x; b++) {
        Cta c;
        c.n = block;
        c.f.resize(block);
        const unsigned nw = (block + 31) / 32;
        c.warp_arrived.assign(nw, 0);
        c.warp_gen.assign(nw, 0);
        c.warp_alive.assign(nw, 0);
        for (unsigned t = 0; t < block; t++) c.warp_alive[t >> 5]++;
        c.cta_alive = (int)block;
        c.slots.assign(block, 0);
        c.body = body;
        cta() = &c;
        bid().x = b;
        bid().y = by;
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = c.f[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stack_pool().data() + (size_t)t * STACK_BYTES;
            f.ctx.uc_stack.ss_size = STACK_BYTES;
            f.ctx.uc_link = &c.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        unsigned long spins = 0;
        for (;;) {
            bool any = false;
            for (unsigned t = 0; t < block; t++) {
                if (c.f[t].done) continue;
                any = true;
                c.cur = t;
                tid().x = t;
                swapcontext(&c.sched, &c.f[t].ctx);
            }
            if (!any) break;
            if (++spins > 200000000ul) { fprintf(stderr, "emu: CTA %u never finishes (lost rendezvous?)\n", b); abort(); }
        }
        cta() = nullptr;
    }
}
inline void launch(unsigned grid, unsigned block, std::function<void()> body, size_t dyn_smem = 0) {
    dim3 g;
    g.x = grid;
    launch(g, block, std::move(body), dyn_smem);
}

template <class T>
inline uint64_t bits(T v) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T unbits(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}
// every lane publishes v, then reads lane `src` of its own warp
template <class T>
inline T exchange(T v, int src) {
    Cta *c = cta();
    const unsigned me = c->cur, base = me & ~31u;
    c->slots[me] = bits(v);
    warp_barrier();
    c = cta();
    const T r = unbits<T>(c->slots[base + (unsigned)(src & 31)]);
    warp_barrier();
    return r;
}
template <class T, class F>
inline T reduce(T v, F op) {
    Cta *c = cta();
    const unsigned me = c->cur, base = me & ~31u;
    c->slots[me] = bits(v);
    warp_barrier();
    c = cta();
    T r = unbits<T>(c->slots[base]);
    const unsigned lanes = std::min(32u, c->n - base);
    for (unsigned l = 1; l < lanes; l++) r = op(r, unbits<T>(c->slots[base + l]));
    warp_barrier();
    return r;
}
}  // namespace emu

namespace emu {
inline dim3 as_dim3(dim3 g) { return g; }
}  // namespace emu

// ---- the runtime API, as far as this repository's host code uses it: one device, everything synchronous ----------
// A stream is a tag; work "enqueued" on it has already run when the call returns (the library enqueues producers
// before consumers, so program order is a valid schedule); events carry no time.
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801,
       cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
typedef struct emu_event_s { int dummy; } *cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) { *free_b = (size_t)3 << 30; *total_b = (size_t)4 << 30; return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = malloc(1); return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(sizeof(emu_event_s)); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *) { memset(h, 0, sizeof(*h)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

#define threadIdx (emu::tid())
#define blockIdx (emu::bid())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

inline void __syncthreads() { emu::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline int emu_lane() { return (int)(emu::cta()->cur & 31u); }

#define EMU_SHFL(T)                                                                                              \
    inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange<T>(v, src); }                            \
    inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::exchange<T>(v, emu_lane() ^ m); }               \
    inline T __shfl_up_sync(unsigned, T v, unsigned d) {                                                         \
        const int l = emu_lane();                                                                                \
        return emu::exchange<T>(v, l >= (int)d ? l - (int)d : l);                                                \
    }                                                                                                            \
    inline T __shfl_down_sync(unsigned, T v, unsigned d) {                                                       \
        const int l = emu_lane();                                                                                \
        return emu::exchange<T>(v, l + (int)d < 32 ? l + (int)d : l);                                            \
    }
EMU_SHFL(float)
EMU_SHFL(double)
EMU_SHFL(int)
EMU_SHFL(unsigned)
EMU_SHFL(long long)
EMU_SHFL(unsigned long long)
#undef EMU_SHFL

inline int __reduce_add_sync(unsigned, int v) { return emu::reduce<int>(v, [](int a, int b) { return a + b; }); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return emu::reduce<unsigned>(v, [](unsigned a, unsigned b) { return a + b; }); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return emu::reduce<unsigned>(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline int __reduce_max_sync(unsigned, int v) { return emu::reduce<int>(v, [](int a, int b) { return a > b ? a : b; }); }
inline int __any_sync(unsigned, int p) { return emu::reduce<int>(p != 0, [](int a, int b) { return a | b; }); }
inline int __all_sync(unsigned, int p) { return emu::reduce<int>(p != 0, [](int a, int b) { return a & b; }); }
inline unsigned __ballot_sync(unsigned, int p) {
    return emu::reduce<unsigned>(p ? 1u << emu_lane() : 0u, [](unsigned a, unsigned b) { return a | b; });
}

template <class T>
inline T __ldg(const T *p) { return *p; }
inline unsigned __float_as_uint(float f) { return emu::unbits<unsigned>(emu::bits(f)); }
inline float __uint_as_float(unsigned u) { return emu::unbits<float>(emu::bits(u)); }
inline int __float_as_int(float f) { return emu::unbits<int>(emu::bits(f)); }
inline float __int_as_float(int u) { return emu::unbits<float>(emu::bits(u)); }
inline long long __double_as_longlong(double d) { return emu::unbits<long long>(emu::bits(d)); }
inline double __longlong_as_double(long long l) { return emu::unbits<double>(emu::bits(l)); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline void __threadfence_block() {}
inline void __threadfence() {}
inline void __threadfence_system() {}
inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }  // one OS thread
inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned atomicExch(unsigned *p, unsigned v) { const unsigned o = *p; *p = v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline float atomicAdd(float *p, float v) { const float o = *p; *p = o + v; return o; }
inline unsigned atomicMin(unsigned *p, unsigned v) { const unsigned o = *p; if (v < o) *p = v; return o; }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicMax(unsigned *p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
// CUDA's global min / max overloads
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline int min(int a, unsigned b) { return (int)((unsigned)a < b ? (unsigned)a : b); }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

// gather.cu -- epoch barrier of the fused feature-row exchange ("fused
// alternative").  finalize_kernel (finalize.cu) has already stored this rank's rows into every
// rank's gather buffer with plain peer stores; what is left of the "all-gather" is publishing
// them: thread t of ONE warp fences, writes this rank's epoch into slot [rank] of rank t's flag
// array (st.release.sys over NVLink) and then spins on slot [t] of its own array until rank t has
// announced the same epoch.  Kernels enqueued behind it on the stream see every rank's rows.
//
// A bounded spin: a peer that never arrives (crashed process) trips the timeout, which is
// recorded in the status slot and reported by bliss_b200_gather_check() instead of hanging the GPU.
#include "common.cuh"
#ifdef BLISS_HOST_EMUL
#include <chrono>
#endif

namespace bliss {

struct PeerFlags {
    unsigned int *flags[MAX_PEERS];  // rank r's flag array: MAX_PEERS epoch slots + 1 status slot
    int world, rank;
};

#ifdef BLISS_HOST_EMUL  // host emulation (tests/cpu_emul): plain loads / stores and the host clock stand in for the PTX
inline unsigned long long global_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline void st_release_sys(unsigned int *p, unsigned int v) { *(volatile unsigned int *)p = v; }
inline unsigned int ld_acquire_sys(const unsigned int *p) { return *(const volatile unsigned int *)p; }
inline void __nanosleep(unsigned) {}
#else
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

__global__ void gather_barrier_kernel(PeerFlags pf, unsigned int epoch, unsigned long long timeout_ns) {
    const int t = threadIdx.x;
    if (t >= pf.world) return;
    __threadfence_system();  // the rows were written by earlier kernels of this stream
    st_release_sys(pf.flags[t] + pf.rank, epoch);
    const unsigned int *mine = pf.flags[pf.rank] + t;
    const unsigned long long t0 = global_ns();
    for (;;) {
        const unsigned int v = ld_acquire_sys(mine);
        if ((int)(v - epoch) >= 0) break;  // wrap-safe: peers may already be one epoch ahead
        if (global_ns() - t0 > timeout_ns) {
            atomicExch(pf.flags[pf.rank] + MAX_PEERS, 1u + (unsigned int)t);
            break;
        }
        __nanosleep(100);
    }
}

int launch_gather_barrier(unsigned int *const *flags, int world, int rank, unsigned int epoch,
                          unsigned long long timeout_ns, cudaStream_t st) {
    PeerFlags pf;
    for (int r = 0; r < MAX_PEERS; r++) pf.flags[r] = r < world ? flags[r] : nullptr;
    pf.world = world;
    pf.rank = rank;
    BLISS_LAUNCH(gather_barrier_kernel, 1, 32, 0, st, pf, epoch, timeout_ns);
    return 1;
}

}  // namespace bliss

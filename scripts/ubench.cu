// ubench.cu -- pipe-rate micro-benchmarks that the FFT kernels' cycle budgets are built on (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench scripts/ubench.cu
// Prints, per test, SM-cycles per warp-instruction per SM (lower = faster) at full occupancy:
//   ffma      scalar FFMA, 8 independent chains per thread
//   ffma2     packed fma.rn.f32x2, 8 independent chains (is FFMA2 one issue slot for two FMAs?)
//   fadd2     packed add.rn.f32x2
//   mix       FFMA2 interleaved with integer ALU ops (does FFMA2 leave issue slots free?)
//   lds64     conflict-free LDS.64 (wavefronts/clk of the shared-memory data pipe)
//   lds128    conflict-free LDS.128
//   shfl      SHFL.BFLY (shares the data pipe with LDS?)
//   lds+shfl  both interleaved
//   f2f/dmul  F2F.F64.F32 + DMUL (geometric_mean's per-bin work)
//   mufu      MUFU.SQRT
//   bulk      cp.async.bulk global->shared 32 KB per CTA iteration (bytes/clk/SM)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float *out, float a, float b) {
    float v[8];
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
    unsigned long long v[8], pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
    for (int i = 0; i < 8; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(f)); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(pa), "l"(pb));
    }
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s ^= v[i];
    ((unsigned long long *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fadd2(float *out, float a) {
    unsigned long long v[8], pa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    for (int i = 0; i < 8; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(f)); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(pa));
    }
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s ^= v[i];
    ((unsigned long long *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 + as many LOP3 (alu pipe): 16 instructions per iteration
__global__ void k_mix(float *out, float a, float b) {
    unsigned long long v[8], pa, pb;
    unsigned int w[8];
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
    for (int i = 0; i < 8; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(f)); w[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(pa), "l"(pb));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(w[(i + 1) & 7]), "r"(it));
        }
    }
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s ^= v[i] + w[i];
    ((unsigned long long *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 + scalar FFMA interleaved (both fma pipe)
__global__ void k_mixf(float *out, float a, float b) {
    unsigned long long v[8], pa, pb;
    float w[8];
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
    for (int i = 0; i < 8; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(f)); w[i] = f; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(w[i]) : "f"(a), "f"(b));
        }
    }
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s ^= v[i] + (unsigned long long)__float_as_uint(w[i]);
    ((unsigned long long *)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>  // 0 lds64, 1 lds128, 2 shfl, 3 lds64+shfl, 4 sts64, 5 lds32
__global__ void k_smem(float *out) {
    __shared__ float4 sm[2048];  // 32 KB
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, i, i, i);
    __syncthreads();
    float acc = 0.f;
    float sv = threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float2 *p2 = (const float2 *)sm + warp * 64 + lane;
    const float4 *p4 = sm + warp * 64 + lane;
    const float *p1 = (const float *)sm + warp * 64 + lane;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 3) {
                float2 t;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"((unsigned)__cvta_generic_to_shared(p2 + 32 * (i & 1))));
                acc += t.x;
            }
            if (MODE == 1) {
                float4 t;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(p4 + 32 * (i & 1))));
                acc += t.x;
            }
            if (MODE == 5) {
                float t;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"((unsigned)__cvta_generic_to_shared(p1 + 32 * (i & 1))));
                acc += t;
            }
            if (MODE == 2 || MODE == 3) {
                sv = __shfl_xor_sync(0xffffffffu, sv, 1 + (i & 15));
            }
            if (MODE == 4) {
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"((unsigned)__cvta_generic_to_shared(p2 + 32 * (i & 1))), "f"(acc), "f"(sv) : "memory");
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sv;
}

__global__ void k_f2f_dmul(double *out, float a) {
    double m[4] = {1.0, 1.0, 1.0, 1.0};
    float f[4];
    for (int i = 0; i < 4; i++) f[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f[i]));
            m[i] *= d;
            m[i] = __longlong_as_double((__double_as_longlong(m[i]) & 0xFFFFFFFFFFFFFll) | 0x3FF0000000000000ll);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = m[0] + m[1] + m[2] + m[3];
}

__global__ void k_mufu(float *out, float a) {
    float v[8];
    for (int i = 0; i < 8; i++) v[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// cp.async.bulk: one elected thread streams CHUNK-byte pieces of global memory into a 2-deep ring, everybody waits
template <int CHUNK>
__global__ void k_bulk(const float *src, size_t src_floats, float *out, int iters) {
    extern __shared__ __align__(128) unsigned char raw[];
    float *buf = (float *)raw;
    __shared__ __align__(8) unsigned long long bar[2];
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const size_t chunk_f = CHUNK / 4;
    size_t pos = ((size_t)blockIdx.x * 977u * chunk_f) % (src_floats - chunk_f);
    pos &= ~(size_t)3;
    auto issue = [&](int stage, size_t p) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + stage * chunk_f);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b0 + 8 * stage), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(src + p), "r"(CHUNK), "r"(b0 + 8 * stage) : "memory");
    };
    if (threadIdx.x == 0) issue(0, pos);
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const int st = it & 1;
        size_t npos = (pos + 131u * chunk_f) % (src_floats - chunk_f);
        npos &= ~(size_t)3;
        if (threadIdx.x == 0 && it + 1 < iters) issue(st ^ 1, npos);
        const unsigned parity = (it >> 1) & 1;
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(b0 + 8 * st), "r"(parity) : "memory");
        }
        acc += buf[st * chunk_f + threadIdx.x];
        __syncthreads();
        pos = npos;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", pr.name, sms, clk_khz);
    const int threads = 256, ctas_per_sm = 4;
    const int grid = sms * ctas_per_sm;
    float *out;
    CK(cudaMalloc(&out, (size_t)grid * threads * 16));
    const double warps_per_sm = threads / 32.0 * ctas_per_sm;
    auto report = [&](const char *name, float ms, double instr_per_thread) {
        const double cyc = ms * 1e-3 * clk_khz * 1e3;
        printf("%-10s %8.3f ms  %.3f SM-cycles per warp-instruction (per SM, all 4 SMSPs together)\n", name, ms,
               cyc / (instr_per_thread * warps_per_sm));
    };
    report("ffma", time_ms([&] { k_ffma<<<grid, threads>>>(out, 1.0001f, 0.5f); }), 8.0 * ITERS);
    report("ffma2", time_ms([&] { k_ffma2<<<grid, threads>>>(out, 1.0001f, 0.5f); }), 8.0 * ITERS);
    report("fadd2", time_ms([&] { k_fadd2<<<grid, threads>>>(out, 0.5f); }), 8.0 * ITERS);
    report("ffma2+lop3", time_ms([&] { k_mix<<<grid, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
    report("ffma2+ffma", time_ms([&] { k_mixf<<<grid, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
    report("lds32", time_ms([&] { k_smem<5><<<grid, threads>>>(out); }), 8.0 * ITERS);
    report("lds64", time_ms([&] { k_smem<0><<<grid, threads>>>(out); }), 8.0 * ITERS);
    report("lds128", time_ms([&] { k_smem<1><<<grid, threads>>>(out); }), 8.0 * ITERS);
    report("sts64", time_ms([&] { k_smem<4><<<grid, threads>>>(out); }), 8.0 * ITERS);
    report("shfl", time_ms([&] { k_smem<2><<<grid, threads>>>(out); }), 8.0 * ITERS);
    report("lds64+shfl", time_ms([&] { k_smem<3><<<grid, threads>>>(out); }), 16.0 * ITERS);
    report("f2f+dmul", time_ms([&] { k_f2f_dmul<<<grid, threads>>>((double *)out, 1.5f); }), 4.0 * ITERS);
    report("mufu.sqrt", time_ms([&] { k_mufu<<<grid, threads>>>(out, 1.5f); }), 8.0 * ITERS);
    // bulk copies: 1 GiB source (larger than L2), 2 and 4 CTAs per SM
    const size_t src_floats = (size_t)256 << 20;
    float *src;
    CK(cudaMalloc(&src, src_floats * 4));
    CK(cudaMemset(src, 0, src_floats * 4));
    for (int cps : {1, 2, 4}) {
        const int iters = 2048;
        CK(cudaFuncSetAttribute(k_bulk<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        float ms = time_ms([&] { k_bulk<32768><<<sms * cps, 128, 65536>>>(src, src_floats, out, iters); });
        CK(cudaGetLastError());
        const double bytes = (double)sms * cps * iters * 32768.0;
        printf("bulk32K x%d CTAs/SM  %8.3f ms  %.1f GB/s  %.1f B/clk/SM\n", cps, ms, bytes / ms / 1e6,
               bytes / (ms * 1e-3 * clk_khz * 1e3) / sms);
        CK(cudaFuncSetAttribute(k_bulk<8832>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        ms = time_ms([&] { k_bulk<8832><<<sms * cps, 128, 32768>>>(src, src_floats, out, iters); });
        CK(cudaGetLastError());
        const double bytes2 = (double)sms * cps * iters * 8832.0;
        printf("bulk8.8K x%d CTAs/SM %8.3f ms  %.1f GB/s  %.1f B/clk/SM\n", cps, ms, bytes2 / ms / 1e6,
               bytes2 / (ms * 1e-3 * clk_khz * 1e3) / sms);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}

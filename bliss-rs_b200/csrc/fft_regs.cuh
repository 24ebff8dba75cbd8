// fft_regs.cuh -- in-register radix-2^k FFT butterflies shared by the 512-point
// phase-vocoder kernel (spectral.cu) and the 8192-point chroma STFT kernel
// (chroma.cu).  Everything is __host__ __device__ so the index logic can be
// unit-tested on a CPU-only box (tests/cpu_emul/).
//
// Convention: forward DFT X[k] = sum_n x[n] exp(-2*pi*i*n*k/N), as rustfft's
// plan_fft_forward (reference call sites src/utils.rs:41-51, src/aubio.rs:235,398).
#pragma once
#include <cuda_runtime.h>

#ifndef BLISS_HD
#define BLISS_HD __host__ __device__ __forceinline__
#endif

namespace bliss {

struct cpx {
    float x, y;
};

#ifdef __CUDA_ARCH__
// MUFU.SQRT (<= 1 ulp): an IEEE-rounded software sqrt costs ~8 instructions per magnitude and the
// magnitudes already carry the f32 FFT's own ~1e-7 rounding noise.
__device__ __forceinline__ float approx_sqrtf(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif

BLISS_HD cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
BLISS_HD cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
BLISS_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// cos(2*pi*k/32), k = 0..8 (f64-rounded-to-f32 literals)
__host__ __device__ constexpr float cos32(int k) {
    // fold k into [0, 8] using symmetries of cosine on a 32-point circle
    k = ((k % 32) + 32) % 32;
    if (k > 16) k = 32 - k;      // cos(2pi - a) = cos(a)
    bool neg = false;
    if (k > 8) { k = 16 - k; neg = true; }  // cos(pi - a) = -cos(a)
    float v = 0.f;
    switch (k) {
        case 0: v = 1.0f; break;
        case 1: v = 0.98078528040323043f; break;
        case 2: v = 0.92387953251128674f; break;
        case 3: v = 0.83146961230254524f; break;
        case 4: v = 0.70710678118654752f; break;
        case 5: v = 0.55557023301960218f; break;
        case 6: v = 0.38268343236508978f; break;
        case 7: v = 0.19509032201612825f; break;
        default: v = 0.0f; break;
    }
    return neg ? -v : v;
}
__host__ __device__ constexpr float sin32(int k) { return cos32(k - 8); }

// multiply by W_N^K = exp(-2*pi*i*K/N), N | 32, K compile-time
template <int K, int N>
BLISS_HD cpx mul_tw(cpx v) {
    constexpr int k32 = (K * (32 / N)) % 32;
    if constexpr (k32 == 0) {
        return v;
    } else if constexpr (k32 == 8) {  // -i
        return cpx{v.y, -v.x};
    } else if constexpr (k32 == 16) {  // -1
        return cpx{-v.x, -v.y};
    } else if constexpr (k32 == 24) {  // +i
        return cpx{-v.y, v.x};
    } else if constexpr (k32 == 4) {  // (1 - i)/sqrt2
        constexpr float h = 0.70710678118654752f;
        return cpx{(v.x + v.y) * h, (v.y - v.x) * h};
    } else if constexpr (k32 == 12) {  // (-1 - i)/sqrt2
        constexpr float h = 0.70710678118654752f;
        return cpx{(v.y - v.x) * h, -(v.x + v.y) * h};
    } else {
        constexpr float c = cos32(k32), s = -sin32(k32);  // W = c + i s
        return cpx{v.x * c - v.y * s, v.x * s + v.y * c};
    }
}

// Decimation-in-frequency radix-2 recursion on v[OFF .. OFF+N), stride 1.
// On return position p (relative) holds X[bitrev_N(p)].
template <int N, int OFF, int TOTAL>
struct DifStage {
    template <int K>
    static BLISS_HD void bfly(cpx (&v)[TOTAL]) {
        if constexpr (K < N / 2) {
            cpx a = v[OFF + K], b = v[OFF + K + N / 2];
            v[OFF + K] = cadd(a, b);
            v[OFF + K + N / 2] = mul_tw<K, N>(csub(a, b));
            bfly<K + 1>(v);
        }
    }
    static BLISS_HD void run(cpx (&v)[TOTAL]) {
        if constexpr (N >= 2) {
            bfly<0>(v);
            DifStage<N / 2, OFF, TOTAL>::run(v);
            DifStage<N / 2, OFF + N / 2, TOTAL>::run(v);
        }
    }
};

template <int N>
BLISS_HD void fft_dif(cpx (&v)[N]) {
    DifStage<N, 0, N>::run(v);
}

__host__ __device__ constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++)
        if (x & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2(int n) {
    int b = 0;
    while ((1 << b) < n) b++;
    return b;
}

}  // namespace bliss

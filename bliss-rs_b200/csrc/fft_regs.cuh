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
// MUFU.SQRT alone: the non-ftz form above costs two scalings and a compare around the MUFU for denormal inputs.
// Callers that keep their data 2^30 above its natural level (exact: the factor is folded into the window and
// taken out again with the final power-of-two scale) can never present a denormal to it.
__device__ __forceinline__ float approx_sqrtf_ftz(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif

// ---- packed FP32 (sm_100a: add/sub/mul/fma .f32x2 -> FADD2 / FMUL2 / FFMA2) -----------------------
// A complex value is a natural f32x2.  One packed instruction does the work of two scalar ones in ONE
// issue slot, and ptxas folds half swaps, per-half negations and scalar broadcasts of the operands into
// instruction modifiers (R.F32x2.LO_HI.NP, R.F32), so a radix-16 butterfly costs 81 FP instructions
// instead of 162.  The FFT kernels are issue-bound (profiles/), not FP-pipe bound, so this is where
// their time goes.  IEEE results per half are those of the scalar instructions.
// Host build (tests/cpu_emul) and -DBLISS_NO_PACKED_FP: plain scalar code with the same meaning.
#if defined(__CUDA_ARCH__) && !defined(BLISS_NO_PACKED_FP)
#define BLISS_PACKED_FP 1
__device__ __forceinline__ unsigned long long pk_(cpx a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ cpx up_(unsigned long long v) {
    cpx r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ cpx padd(cpx a, cpx b) {  // (a.x + b.x, a.y + b.y)
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk_(a)), "l"(pk_(b)));
    return up_(r);
}
__device__ __forceinline__ cpx psub(cpx a, cpx b) {  // (a.x - b.x, a.y - b.y)
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk_(a)), "l"(pk_(b)));
    return up_(r);
}
__device__ __forceinline__ cpx pmul(cpx a, cpx b) {  // (a.x * b.x, a.y * b.y)
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk_(a)), "l"(pk_(b)));
    return up_(r);
}
__device__ __forceinline__ cpx pfma(cpx a, cpx b, cpx c) {  // (a.x * b.x + c.x, a.y * b.y + c.y)
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk_(a)), "l"(pk_(b)), "l"(pk_(c)));
    return up_(r);
}
#else
BLISS_HD cpx padd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
BLISS_HD cpx psub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
BLISS_HD cpx pmul(cpx a, cpx b) { return cpx{a.x * b.x, a.y * b.y}; }
// fused, like the FFMA2 it stands in for (host emulation: tests/cpu_emul; nvcc contracts a * b + c into FFMA anyway)
BLISS_HD cpx pfma(cpx a, cpx b, cpx c) { return cpx{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif

BLISS_HD cpx cadd(cpx a, cpx b) { return padd(a, b); }
BLISS_HD cpx csub(cpx a, cpx b) { return psub(a, b); }
// a * b = b.x (a.x, a.y) + b.y (-a.y, a.x)
BLISS_HD cpx cmul(cpx a, cpx b) { return pfma(a, cpx{b.x, b.x}, pmul(cpx{-a.y, a.x}, cpx{b.y, b.y})); }

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
    } else if constexpr (k32 == 4) {  // (1 - i)/sqrt2:  h (x + y, y - x)
        constexpr float h = 0.70710678118654752f;
        return pmul(padd(v, cpx{v.y, -v.x}), cpx{h, h});
    } else if constexpr (k32 == 12) {  // (-1 - i)/sqrt2:  h (y - x, -x - y)
        constexpr float h = 0.70710678118654752f;
        return pmul(psub(cpx{v.y, -v.x}, v), cpx{h, h});
    } else {
        constexpr float c = cos32(k32), s = -sin32(k32);  // W = c + i s
        return pfma(v, cpx{c, c}, pmul(cpx{-v.y, v.x}, cpx{s, s}));
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

// ---- radix up to 64: twiddles on the 256-point circle ---------------------------------------------------
// cos(2*pi*k/256), k = 0..64 (f64-rounded-to-f32 literals)
__host__ __device__ constexpr float cos256_tab(int k) {
    constexpr float t[65] = {
        1.f, 0.99969881869620425f, 0.99879545620517241f, 0.99729045667869021f,
        0.99518472667219693f, 0.99247953459870997f, 0.98917650996478101f, 0.98527764238894122f,
        0.98078528040323043f, 0.97570213003852857f, 0.97003125319454397f, 0.96377606579543984f,
        0.95694033573220882f, 0.94952818059303667f, 0.94154406518302081f, 0.93299279883473896f,
        0.92387953251128674f, 0.91420975570353069f, 0.90398929312344334f, 0.89322430119551532f,
        0.88192126434835505f, 0.87008699110871146f, 0.85772861000027212f, 0.84485356524970712f,
        0.83146961230254524f, 0.81758481315158371f, 0.80320753148064494f, 0.78834642762660634f,
        0.77301045336273699f, 0.75720884650648457f, 0.74095112535495911f, 0.724247082951467f,
        0.70710678118654757f, 0.68954054473706694f, 0.67155895484701833f, 0.65317284295377676f,
        0.63439328416364549f, 0.61523159058062682f, 0.59569930449243347f, 0.57580819141784534f,
        0.55557023301960229f, 0.53499761988709726f, 0.51410274419322166f, 0.49289819222978409f,
        0.47139673682599781f, 0.4496113296546066f, 0.4275550934302822f, 0.40524131400498986f,
        0.38268343236508984f, 0.35989503653498828f, 0.33688985339222005f, 0.31368174039889157f,
        0.29028467725446233f, 0.26671275747489842f, 0.24298017990326398f, 0.21910124015686977f,
        0.19509032201612833f, 0.17096188876030136f, 0.14673047445536175f, 0.12241067519921628f,
        0.09801714032956077f, 0.073564563599667454f, 0.049067674327418126f, 0.024541228522912264f,
        0.f};
    return t[k];
}
__host__ __device__ constexpr float cos256(int k) {
    k = ((k % 256) + 256) % 256;
    if (k > 128) k = 256 - k;
    bool neg = false;
    if (k > 64) { k = 128 - k; neg = true; }
    const float v = cos256_tab(k);
    return neg ? -v : v;
}
__host__ __device__ constexpr float sin256(int k) { return cos256(k - 64); }

// multiply by W_256^K = exp(-2*pi*i*K/256), K compile-time
template <int K>
BLISS_HD cpx mul_tw256(cpx v) {
    constexpr int k = ((K % 256) + 256) % 256;
    if constexpr (k == 0) {
        return v;
    } else if constexpr (k == 64) {  // -i
        return cpx{v.y, -v.x};
    } else if constexpr (k == 128) {  // -1
        return cpx{-v.x, -v.y};
    } else if constexpr (k == 192) {  // +i
        return cpx{-v.y, v.x};
    } else if constexpr (k == 32) {  // (1 - i)/sqrt2
        constexpr float h = 0.70710678118654752f;
        return pmul(padd(v, cpx{v.y, -v.x}), cpx{h, h});
    } else if constexpr (k == 96) {  // (-1 - i)/sqrt2
        constexpr float h = 0.70710678118654752f;
        return pmul(psub(cpx{v.y, -v.x}, v), cpx{h, h});
    } else {
        constexpr float c = cos256(k), s = -sin256(k);  // W = c + i s
        return pfma(v, cpx{c, c}, pmul(cpx{-v.y, v.x}, cpx{s, s}));
    }
}

// Decimation-in-frequency radix-2 recursion for N | 256 (same structure as DifStage)
template <int N, int OFF, int TOTAL>
struct DifStageG {
    template <int K>
    static BLISS_HD void bfly(cpx (&v)[TOTAL]) {
        if constexpr (K < N / 2) {
            cpx a = v[OFF + K], b = v[OFF + K + N / 2];
            v[OFF + K] = cadd(a, b);
            v[OFF + K + N / 2] = mul_tw256<K * (256 / N)>(csub(a, b));
            bfly<K + 1>(v);
        }
    }
    static BLISS_HD void run(cpx (&v)[TOTAL]) {
        if constexpr (N >= 2) {
            bfly<0>(v);
            DifStageG<N / 2, OFF, TOTAL>::run(v);
            DifStageG<N / 2, OFF + N / 2, TOTAL>::run(v);
        }
    }
};

// 64-point FFT in registers; on return position p holds X[bitrev_6(p)]
BLISS_HD void fft_dif64(cpx (&v)[64]) { DifStageG<64, 0, 64>::run(v); }

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

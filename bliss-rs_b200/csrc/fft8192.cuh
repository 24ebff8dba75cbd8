// fft8192.cuh -- shared-memory 8192-point complex FFT used by the chroma STFT kernel:
// TWO reflect-padded Hann frames per transform (a + i b), 8192 = 16 x 16 x 32,
// three in-register radix passes over one in-place buffer (decimation in frequency).
//
// After pass 3, X[k] with k = k1 + 16*k2 + 256*k3 lives at logical index
// 512*k1 + 32*k2 + k3.  Logical index i is stored at pad(i) so that every pass
// and the natural-order gather of the epilogue are bank-conflict free.
//
// Replaces the per-frame rustfft call of utils::stft (src/utils.rs:41-61).
// __host__ __device__ so tests/cpu_emul can run the passes thread by thread.
#pragma once
#include "fft_regs.cuh"

namespace bliss {
namespace f8k {

constexpr int N = 8192;
constexpr int BUF_CPX = 8192 + 256 + 16;  // pad(8191) + 1 = 8462 -> round up
BLISS_HD int pad(int i) { return i + (i >> 5) + (i >> 9); }
BLISS_HD int xpos(int k) { return 512 * (k & 15) + 32 * ((k >> 4) & 15) + (k >> 8); }

// All three passes address the buffer as  base(thread) + constant * slot, so the index
// arithmetic is a handful of integer ops per butterfly (pad() is folded in by hand):
//   pass 1 store : pad(b + 512 k1)        = b + (b>>5)            + 529 k1      (b < 512)
//   pass 2       : pad(512 blk + j + 32q) = 529 blk + j           + 33 q        (j < 32)
//   pass 3       : pad(32 b + q)          = 33 b + (b>>4)         + q           (b < 256)
//   bin k = tid + 256 m lives at            529 (tid&15) + 33 ((tid>>4)&15) + m (xbase(tid) + m)

// pass 1: butterfly b in [0,512): v[q] = z[b + 512 q] supplied by the caller
BLISS_HD void pass1_store(int b, cpx (&v)[16], const cpx *tw /*[8192]*/, cpx *buf) {
    fft_dif<16>(v);
    cpx *o = buf + b + (b >> 5);
    const cpx *t = tw + 0;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int k1 = bitrev(s, 4);
        cpx r = v[s];
        if (k1 != 0) r = cmul(r, t[b * k1]);
        o[529 * k1] = r;
    }
}

// pass 2: butterfly b in [0,512): block = b>>5 (k1), j = b&31; radix 16 at stride 32
BLISS_HD void pass2(int b, const cpx *tw, cpx *buf) {
    const int blk = b >> 5, j = b & 31;
    cpx *p = buf + 529 * blk + j;
    cpx v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = p[33 * q];
    fft_dif<16>(v);
    const cpx *t = tw + 0;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int k2 = bitrev(s, 4);
        cpx o = v[s];
        if (k2 != 0) o = cmul(o, t[(16 * j) * k2]);
        p[33 * k2] = o;
    }
}

// pass 3: butterfly b in [0,512): 32-block blk = b & 255, parity j = b >> 8; radix 16 at
// stride 2 inside the block, twiddle W32^(j k3).  Leaves u_j[k3] at logical 32 blk + 2 k3 + j;
// the closing radix-2,  X[k1 + 16 k2 + 256 k3 + 4096 k4] = u_0[k3] + (-1)^k4 u_1[k3],  is folded
// into the epilogue (bin_value below).
template <int S>
BLISS_HD void tw32_slots(cpx (&v)[16]) {
    if constexpr (S < 16) {
        v[S] = mul_tw<bitrev(S, 4), 32>(v[S]);
        tw32_slots<S + 1>(v);
    }
}
BLISS_HD void pass3(int b, cpx *buf) {
    const int blk = b & 255, j = b >> 8;
    cpx *p = buf + 33 * blk + (blk >> 4) + j;
    cpx v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = p[2 * q];
    fft_dif<16>(v);
    if (j) tw32_slots<0>(v);
#pragma unroll
    for (int s = 0; s < 16; s++) p[2 * bitrev(s, 4)] = v[s];
}

// padded position of the (u_0, u_1) pair of bin (t + 512 m), t < 512, m < 8:  ebase(t) + 4 m
BLISS_HD int ebase(int t) { return 529 * (t & 15) + 33 * ((t >> 4) & 15) + 2 * (t >> 8); }

// X[k] for k < 4096 (k4 = 0) and X[k + 4096] (k4 = 1) from the pair at p
BLISS_HD cpx pair_sum(const cpx *p) { return cadd(p[0], p[1]); }
BLISS_HD cpx pair_diff(const cpx *p) { return csub(p[0], p[1]); }

// generic (slow) accessor used by the CPU emulation test: X[k], 0 <= k < 8192
BLISS_HD cpx bin_value(const cpx *buf, int k) {
    const int k4 = k >> 12, kk = k & 4095;
    const int pos = pad(512 * (kk & 15) + 32 * ((kk >> 4) & 15) + 2 * (kk >> 8));
    return k4 ? pair_diff(buf + pos) : pair_sum(buf + pos);
}

// magnitudes of the two real frames packed in Z: see pv::untangle_mag
BLISS_HD void untangle_mag(cpx zk, cpx zm, float &magA, float &magB) {
    const float ar = 0.5f * (zk.x + zm.x), ai = 0.5f * (zk.y - zm.y);
    const float br = 0.5f * (zk.y + zm.y), bi = 0.5f * (zm.x - zk.x);
#ifdef __CUDA_ARCH__
    magA = __fsqrt_rn(__fadd_rn(__fmul_rn(ar, ar), __fmul_rn(ai, ai)));
    magB = __fsqrt_rn(__fadd_rn(__fmul_rn(br, br), __fmul_rn(bi, bi)));
#else
    magA = sqrtf(ar * ar + ai * ai);
    magB = sqrtf(br * br + bi * bi);
#endif
}

// reflect-padded sample (utils.rs:11-24): padded index p in [0, n + 8192)
BLISS_HD float padded_sample(const float *x, int n, long long p) {
    long long idx = p - 4096;
    if (idx < 0) idx = -idx;
    else if (idx >= n) idx = 2ll * (n - 1) - idx;
    return x[idx];
}

}  // namespace f8k
}  // namespace bliss

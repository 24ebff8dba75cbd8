// rfft8192.cuh -- 8192-point REAL FFT of one reflect-padded Hann frame, computed as a
// 4096-point complex FFT of z[n] = x[2n] + i x[2n+1] followed by the real-input untangling
//   X[k] = (Z[k] + conj Z[N/2-k]) / 2  -  i W_N^k (Z[k] - conj Z[N/2-k]) / 2,   N = 8192.
// 4096 = 16 x 16 x 16: three in-register radix-16 passes over one in-place shared-memory buffer
// (decimation in frequency), ONE butterfly per thread per pass with 256 threads, 34 KB per frame.
//
// After pass 3, Z[k] with k = k1 + 16 k2 + 256 k3 lives at logical index 256 k1 + 16 k2 + k3;
// logical index i is stored at pad(i) = i + (i>>4) + (i>>8), which makes every pass and the
// natural-order gather of the epilogue bank-conflict free.  All addressing is base(thread) +
// constant * slot (pad() folded in by hand, see the comments at each pass).
//
// Replaces the per-frame rustfft call of utils::stft (src/utils.rs:41-61).
// __host__ __device__ so tests/cpu_emul can run the passes thread by thread.
#pragma once
#include "fft_regs.cuh"

namespace bliss {
namespace r8k {

constexpr int NC = 4096;                       // complex points
constexpr int BUF_CPX = 4096 + 256 + 16 + 16;  // pad(4095) + 1 = 4366 -> round up
BLISS_HD int pad(int i) { return i + (i >> 4) + (i >> 8); }

// pass 1: butterfly b in [0,256): v[q] = z[b + 256 q];  pad(b + 256 k1) = b + (b>>4) + 273 k1.
// tw1 is laid out [k1][b] = W4096^(b k1) so that the 32 lanes of a warp read 256 contiguous bytes.
// LB: pitch of a 16-element column group inside a 256-element row (position = 273 r + LB c + m for the logical
// element 256 r + 16 c + m).  17 = pad() above, the measured layout.  16 (experimental, VARIANT_LAY16) drops the
// per-group padding: passes 1-3 stay conflict-free and the mirror loads of the pair epilogue lose their 2-way
// conflict (thread t reads the block of thread 256 - t: with LB = 17 lanes 0 and 15 of every half-warp meet in one
// bank, 256 instead of 136 wavefronts per frame -- the model is tests/test_host_abi.py).
template <int LB = 17>
BLISS_HD void pass1_store(int b, cpx (&v)[16], const cpx *tw1 /*[16][256]*/, cpx *buf) {
    fft_dif<16>(v);
    cpx *o = LB == 17 ? buf + b + (b >> 4) : buf + (b & 15) + LB * (b >> 4);
    const cpx *t = tw1 + b;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int k1 = bitrev(s, 4);
        cpx r = v[s];
        if (k1 != 0) r = cmul(r, t[256 * k1]);
        o[273 * k1] = r;
    }
}

// pass 1 with 4 twiddle loads instead of 15 (VARIANT_TWPROD): W^b, W^2b, W^4b, W^8b come from the table, the
// other eleven are products of at most three of them (<= 3 extra roundings, ~2e-7 relative).  The kernel is bound
// by the L1 / shared-memory data pipe (profiles/): 11 fewer 64-bit loads per thread are 176 fewer wavefronts per
// frame for 22 more packed FP instructions per thread.  Products are formed right before their use so that only
// the four loaded values stay live.
template <int LB = 17>
BLISS_HD void pass1_store_prod(int b, cpx (&v)[16], const cpx *tw1 /*[16][256]*/, cpx *buf) {
    fft_dif<16>(v);
    cpx *o = LB == 17 ? buf + b + (b >> 4) : buf + (b & 15) + LB * (b >> 4);
    const cpx *t = tw1 + b;
    const cpx t1 = t[256 * 1], t2 = t[256 * 2], t4 = t[256 * 4], t8 = t[256 * 8];
#define BLISS_P1(k1, w) o[273 * (k1)] = cmul(v[bitrev((k1), 4)], (w))
    o[0] = v[0];
    BLISS_P1(8, t8);
    BLISS_P1(4, t4);
    BLISS_P1(12, cmul(t4, t8));
    BLISS_P1(2, t2);
    BLISS_P1(10, cmul(t2, t8));
    const cpx t6 = cmul(t2, t4);
    BLISS_P1(6, t6);
    BLISS_P1(14, cmul(t6, t8));
    BLISS_P1(1, t1);
    BLISS_P1(9, cmul(t1, t8));
    const cpx t5 = cmul(t1, t4);
    BLISS_P1(5, t5);
    BLISS_P1(13, cmul(t5, t8));
    const cpx t3 = cmul(t1, t2);
    BLISS_P1(3, t3);
    BLISS_P1(11, cmul(t3, t8));
    const cpx t7 = cmul(t3, t4);
    BLISS_P1(7, t7);
    BLISS_P1(15, cmul(t7, t8));
#undef BLISS_P1
}

// Periodic Hann pair (w[n], w[n + 1]) for n = 2 (tid + 256 Q) from the thread's own phase (VARIANT_WINSYN):
//   w[n] = 0.5 - 0.5 cos(theta + 2 pi Q / 16),  theta = 2 pi (2 tid [+ 1]) / 8192
//        = 0.5 - 0.5 (cos theta C_Q - sin theta S_Q)
// cw = (cos theta_even, cos theta_odd), sw = (sin theta_even, sin theta_odd) (f64-generated table, one 16-byte load
// per frame); C_Q, S_Q are compile-time constants: two packed FMAs per sample pair instead of one 64-bit load.
// Differs from the f32 `0.5 - 0.5 cosf(...)` table of the reference by ~1e-7 absolute per coefficient.
template <int Q>
BLISS_HD cpx hann_pair(cpx cw, cpx sw) {
    constexpr float a = -0.5f * cos32(2 * Q), b = 0.5f * sin32(2 * Q);
    return pfma(cw, cpx{a, a}, pfma(sw, cpx{b, b}, cpx{0.5f, 0.5f}));
}

// pass 2: butterfly b in [0,256): blk = b>>4 (k1), j = b&15; radix 16 at stride 16 inside the
// 256-block;  pad(256 blk + j + 16 q) = 273 blk + j + 17 q;  twiddle tw2[k2][j] = W256^(j k2)
// (a 2 KB table the kernel keeps in shared memory)
template <int LB = 17>
BLISS_HD void pass2(int b, const cpx *tw2 /*[16][16]*/, cpx *buf) {
    const int blk = b >> 4, j = b & 15;
    cpx *p = buf + 273 * blk + j;
    cpx v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = p[LB * q];
    fft_dif<16>(v);
#pragma unroll
    for (int s = 0; s < 16; s++) {
        const int k2 = bitrev(s, 4);
        cpx o = v[s];
        if (k2 != 0) o = cmul(o, tw2[16 * k2 + j]);
        p[LB * k2] = o;
    }
}

// pass 2 with 4 twiddle loads instead of 15 (VARIANT_TWPROD, like pass1_store_prod): W256^j, ^2j, ^4j, ^8j from
// the shared-memory table, the other eleven by products formed right before their use.
template <int LB = 17>
BLISS_HD void pass2_prod(int b, const cpx *tw2 /*[16][16]*/, cpx *buf) {
    const int blk = b >> 4, j = b & 15;
    cpx *p = buf + 273 * blk + j;
    cpx v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = p[LB * q];
    fft_dif<16>(v);
    const cpx t1 = tw2[16 * 1 + j], t2 = tw2[16 * 2 + j], t4 = tw2[16 * 4 + j], t8 = tw2[16 * 8 + j];
#define BLISS_P2(k2, w) p[LB * (k2)] = cmul(v[bitrev((k2), 4)], (w))
    p[0] = v[0];
    BLISS_P2(8, t8);
    BLISS_P2(4, t4);
    BLISS_P2(12, cmul(t4, t8));
    BLISS_P2(2, t2);
    BLISS_P2(10, cmul(t2, t8));
    const cpx t6 = cmul(t2, t4);
    BLISS_P2(6, t6);
    BLISS_P2(14, cmul(t6, t8));
    BLISS_P2(1, t1);
    BLISS_P2(9, cmul(t1, t8));
    const cpx t5 = cmul(t1, t4);
    BLISS_P2(5, t5);
    BLISS_P2(13, cmul(t5, t8));
    const cpx t3 = cmul(t1, t2);
    BLISS_P2(3, t3);
    BLISS_P2(11, cmul(t3, t8));
    const cpx t7 = cmul(t3, t4);
    BLISS_P2(7, t7);
    BLISS_P2(15, cmul(t7, t8));
#undef BLISS_P2
}

// pass 3: butterfly b in [0,256): 16 consecutive logical elements;  pad(16 b + q) = 17 b + (b>>4) + q
BLISS_HD void pass3(int b, cpx *buf) {
    cpx *p = buf + 17 * b + (b >> 4);
    cpx v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = p[q];
    fft_dif<16>(v);
#pragma unroll
    for (int s = 0; s < 16; s++) p[bitrev(s, 4)] = v[s];
}

// padded position of Z[t + 256 m], t < 256, m < 16:  zbase(t) + m
template <int LB = 17>
BLISS_HD int zbase(int t) { return 273 * (t & 15) + LB * (t >> 4); }

// pass 3 fused with the untangling: thread t transforms the block that ENDS UP holding its own
// natural-order bins (block zbase(t) = logical elements 16 b + q with b = 16 (t & 15) + (t >> 4)),
// so that afterwards v[bitrev(m)] = Z[t + 256 m] is already in its registers.  Only the upper half
// (m = 8..15) goes back to shared memory: those are the mirror values Z[4096 - k] of the bins
// k = t' + 256 m', m' < 8, that thread t' = 256 - t untangles (thread 0 mirrors onto itself and also
// publishes m = 0).  Shared-memory traffic of pass 3 + epilogue: 16 loads + 8 stores + 8 loads per
// thread instead of 16 + 16 + 32.
template <int LB = 17>
BLISS_HD void pass3_regs(int t, cpx (&v)[16], cpx *buf) {
    cpx *p = buf + zbase<LB>(t);
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = p[q];
    fft_dif<16>(v);
#pragma unroll
    for (int m = 8; m < 16; m++) p[m] = v[bitrev(m, 4)];
    if (t == 0) p[0] = v[0];
}

// generic accessor (CPU emulation test)
BLISS_HD cpx z_value(const cpx *buf, int k) {
    return buf[pad(256 * (k & 15) + 16 * ((k >> 4) & 15) + (k >> 8))];
}

// |X[k]| from Z[k], Z[(4096-k)%4096] and w = W8192^k, as `(re*re + im*im).sqrt()` in f32
// (src/utils.rs:57-60)
BLISS_HD float sum_sq_sqrt_half(cpx x) {  // 0.5 * sqrt(x.x^2 + x.y^2), squares and sum un-fused
    const cpx q = pmul(x, x);
#ifdef __CUDA_ARCH__
    return 0.5f * approx_sqrtf(__fadd_rn(q.x, q.y));
#else
    return 0.5f * sqrtf(q.x + q.y);
#endif
}

BLISS_HD float untangle_mag(cpx zk, cpx zm, cpx w) {
    // E = (Zk + conj Zm)/2, O = (Zk - conj Zm)/2;  X = E - i w O
    // (the two factors 1/2 are applied once to the magnitude: exact, power of two)
    const cpx e = padd(zk, cpx{zm.x, -zm.y});
    const cpx o = padd(zk, cpx{-zm.x, zm.y});
    const cpx p = cmul(o, w);
    // -i (pr + i pi) = pi - i pr
    return sum_sq_sqrt_half(padd(e, cpx{p.y, -p.x}));
}

// Both members of a mirror pair from ONE load pair: |X[k]| and |X[4096 - k]| out of Z[k], Z[4096 - k]
// and w = W8192^k.  With E, O, P = w O as in untangle_mag:  X[k] = E - i P,
// X[4096 - k] = conj(E) - i conj(P)  (because W8192^(4096-k) = -conj w, O' = -conj O), whose magnitude
// is that of E + i P.
BLISS_HD void untangle_mag_pair(cpx zk, cpx zm, cpx w, float &mag_k, float &mag_m) {
    const cpx e = padd(zk, cpx{zm.x, -zm.y});
    const cpx o = padd(zk, cpx{-zm.x, zm.y});
    const cpx p = cmul(o, w);
    mag_k = sum_sq_sqrt_half(padd(e, cpx{p.y, -p.x}));  // (er + pi, ei - pr)
    mag_m = sum_sq_sqrt_half(padd(e, cpx{-p.y, p.x}));  // (er - pi, ei + pr)
}

// reflect-padded sample (utils.rs:11-24): idx relative to the song, may be < 0 or >= n
BLISS_HD float reflect_sample(const float *x, int n, long long idx) {
    if (idx < 0) idx = -idx;
    else if (idx >= n) idx = 2ll * (n - 1) - idx;
    return x[idx];
}

}  // namespace r8k
}  // namespace bliss

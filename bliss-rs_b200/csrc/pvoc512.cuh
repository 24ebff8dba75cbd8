// pvoc512.cuh -- one warp transforms TWO consecutive 512-sample phase-vocoder
// frames (timbral frames 2j and 2j+1; the latter is tempo frame j) as ONE complex
// 512-point FFT: z[n] = w[n] * (xA[n] + i xB[n]).
//
//   512 = 16 (in-register, over n1) x 32 (16 in-register + one shuffle radix-2, over n2)
//   n = n2 + 32*n1,   k = k1 + 16*k2
//
// Replaces, per frame: PVoc::do_ (src/aubio.rs:182-264) and PVocTempo::do_
// (:338-425).  aubio's fvec_shift (:219-229) multiplies bin k by (-1)^k, which
// leaves every magnitude -- the only thing either caller keeps -- unchanged, so
// it is dropped here.
//
// The phase functions are __host__ __device__ and take the lane explicitly so
// tests/cpu_emul can run all 32 "lanes" of a warp in a loop on the host.
#pragma once
#include "fft_regs.cuh"

namespace bliss {

namespace pv {

constexpr int ROW = 33;                       // padded row (float2 units) of the 16x32 exchange tile
constexpr int EXCH_CPX = 584;                 // >= max(16*33, zpos(511)+1 = 575)
BLISS_HD int zpos(int k) { return k + (k >> 3); }   // padded natural-order position of bin k
// The same tile with another padding (experimental): SHIFT = 4 keeps the 8-bins-per-lane loads of pvoc512_kernel
// conflict-free and also frees its natural-order STORE (lanes k1 = 0..15 write consecutive slots; with SHIFT = 3
// slots k1 = 0 and k1 = 15 of every store share a bank: 64 instead of 32 wavefronts per frame pair -- the 35
// conflict wavefronts per pair of the ncu capture).  SHIFT = 0: no padding, the conflict-free choice when lanes own
// bins lane + 32 i (stft512_pairs_kernel).  Wavefront model: tests/test_host_abi.py.
template <int SHIFT>
BLISS_HD int zpos_s(int k) { return SHIFT > 0 ? k + (k >> SHIFT) : k; }

// Phase A (lane = n2): r[n1] = z[lane + 32*n1]; radix-16 over n1, twiddle
// W512^(lane*k1), scatter to S[k1][lane].
BLISS_HD void phase_a(int lane, cpx (&r)[16], const cpx *twA, cpx *S) {
    fft_dif<16>(r);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int k1 = bitrev(q, 4);
        cpx v = r[q];
        if (k1 != 0) v = cmul(v, twA[k1 * 32 + lane]);
        S[k1 * ROW + lane] = v;
    }
}

// Phase A with the lane's twiddles W512^(lane k1) formed from four of them (k1 = 1, 2, 4, 8: loop-invariant
// per lane, kept in registers for the whole run) instead of 15 shared-memory loads per frame pair
// (experimental, VARIANT_PV_TWPROD): <= 3 extra roundings per twiddle (~2e-7 relative).
BLISS_HD void phase_a_prod(int lane, cpx (&r)[16], cpx t1, cpx t2, cpx t4, cpx t8, cpx *S) {
    fft_dif<16>(r);
    cpx *o = S + lane;
#define BLISS_PA(k1, w) o[(k1) * ROW] = cmul(r[bitrev((k1), 4)], (w))
    o[0] = r[0];
    BLISS_PA(8, t8);
    BLISS_PA(4, t4);
    BLISS_PA(12, cmul(t4, t8));
    BLISS_PA(2, t2);
    BLISS_PA(10, cmul(t2, t8));
    const cpx t6 = cmul(t2, t4);
    BLISS_PA(6, t6);
    BLISS_PA(14, cmul(t6, t8));
    BLISS_PA(1, t1);
    BLISS_PA(9, cmul(t1, t8));
    const cpx t5 = cmul(t1, t4);
    BLISS_PA(5, t5);
    BLISS_PA(13, cmul(t5, t8));
    const cpx t3 = cmul(t1, t2);
    BLISS_PA(3, t3);
    BLISS_PA(11, cmul(t3, t8));
    const cpx t7 = cmul(t3, t4);
    BLISS_PA(7, t7);
    BLISS_PA(15, cmul(t7, t8));
#undef BLISS_PA
}

// Phase B, lane = p*16 + k1: gather the parity-p half of row k1.
BLISS_HD void phase_b_load(int lane, cpx (&u)[16], const cpx *S) {
    const int k1 = lane & 15, p = lane >> 4;
#pragma unroll
    for (int j = 0; j < 16; j++) u[j] = S[k1 * ROW + 2 * j + p];
}

template <int Q>
BLISS_HD void tw32_all(cpx (&u)[16]) {
    if constexpr (Q < 16) {
        u[Q] = mul_tw<bitrev(Q, 4), 32>(u[Q]);
        tw32_all<Q + 1>(u);
    }
}

// radix-16 over n2' then (odd half only) the W32^k2' twiddle of the final radix-2.
BLISS_HD void phase_b_fft(int lane, cpx (&u)[16]) {
    fft_dif<16>(u);
    if (lane >> 4) tw32_all<0>(u);
}

// final radix-2 between lane (0,k1) and lane (1,k1): `other` is the partner's value
// of the same slot.  Even half keeps F0 + W F1 (k2 = k2'), odd half F0 - W F1 (k2 = k2'+16).
BLISS_HD cpx phase_b_combine(int lane, cpx own, cpx other) {
    const float sg = (lane >> 4) ? -1.f : 1.f;  // other + sg * own: exact, one FFMA2
    return pfma(own, cpx{sg, sg}, other);
}

// bin index held in slot q of lane after the combine
BLISS_HD int bin_of(int lane, int q) {
    const int k1 = lane & 15, p = lane >> 4;
    return k1 + 16 * (bitrev(q, 4) + 16 * p);
}

// Untangle the two real spectra from Z (complex FFT of a + i b):
//   A[k] = (Z[k] + conj Z[N-k]) / 2,   B[k] = (Z[k] - conj Z[N-k]) / (2i)
// and return their magnitudes as `(re*re + im*im).sqrt()` (aubio.rs:248-252).
// SCALE2 = true returns 2|A|, 2|B| (the exact factor 1/2 is folded into a later power-of-two scale).
template <bool SCALE2 = false, bool FTZ = false>
BLISS_HD void untangle_mag(cpx zk, cpx zm, float &magA, float &magB) {
    const float h = SCALE2 ? 1.0f : 0.5f;
    // 2A = (zk.x + zm.x, zk.y - zm.y),  2B = (zk.y + zm.y, zm.x - zk.x)
    cpx a = padd(zk, cpx{zm.x, -zm.y});
    cpx b = padd(cpx{zk.y, -zk.x}, cpx{zm.y, zm.x});
    if (!SCALE2) {
        a = pmul(a, cpx{h, h});
        b = pmul(b, cpx{h, h});
    }
    const cpx qa = pmul(a, a), qb = pmul(b, b);
#ifdef __CUDA_ARCH__
    if (FTZ) {
        magA = approx_sqrtf_ftz(__fadd_rn(qa.x, qa.y));
        magB = approx_sqrtf_ftz(__fadd_rn(qb.x, qb.y));
    } else {
        magA = approx_sqrtf(__fadd_rn(qa.x, qa.y));
        magB = approx_sqrtf(__fadd_rn(qb.x, qb.y));
    }
#else
    magA = sqrtf(qa.x + qa.y);
    magB = sqrtf(qb.x + qb.y);
#endif
}

}  // namespace pv
}  // namespace bliss

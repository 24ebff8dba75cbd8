// spectral.cu -- K1: fused 512-point phase-vocoder kernel (timbral descriptors +
// tempo spectral flux) and K2: time-domain reductions (ZCR, loudness, block energy).
//
// K1 replaces, for every song of the wave:
//   PVoc::do_ / PVocTempo::do_            src/aubio.rs:182-264, 338-425
//   spectral_centroid / spectral_rolloff  src/aubio.rs:16-58  (+ clamp timbral.rs:184-187)
//   geometric_mean + flatness             src/utils.rs:101-117, src/timbral.rs:196-208
//   SpecFlux::do_                         src/aubio.rs:455-467
// driven as Song::analyze_with_options drives them (src/song/mod.rs:433-468).
//
// Data movement: each warp walks a run of consecutive frame pairs of one song and
// keeps a sliding window of 20 samples per lane in registers, so every PCM sample
// is fetched from HBM once per run (plus a one-pair halo) with fully coalesced
// 128-byte warp loads; nothing but 3 floats per frame and 1 float per tempo
// frame is written back.
#include "common.cuh"
#include "pvoc512.cuh"

namespace bliss {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ cpx shfl_xor_cpx(cpx v, int m) {
    return cpx{__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)};
}

// Per-frame descriptors from the 8 consecutive norms a lane holds (bins 8*lane..8*lane+7).
// Returns centroid (Hz), rolloff (Hz), flatness to lane 0 (all lanes compute them).
__device__ __forceinline__ void frame_descriptors(const float (&v)[8], int lane, float &centroid,
                                                  float &rolloff, float &flatness) {
    float s1 = 0.f, sw = 0.f;
    float c[8];
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s1 += v[i];
        sw += (float)(8 * lane + i) * v[i];
        run += v[i] * v[i];
        c[i] = run;
    }
    // geometric_mean's per-group product, utils.rs:104-111 (one 8-group per lane)
    double m = ((double)v[0] * (double)v[1]) * ((double)v[2] * (double)v[3]);
    m *= 3.273390607896142e150;
    m *= ((double)v[4] * (double)v[5]) * ((double)v[6] * (double)v[7]);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
    const int zero = __any_sync(0xffffffffu, m == 0.0);
    int ex = (int)(bits >> 52);
    double mant = __longlong_as_double((long long)((bits & 0xFFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    ex = __reduce_add_sync(0xffffffffu, ex);
    mant = warp_prod(mant);

    s1 = warp_sum(s1);
    sw = warp_sum(sw);
    // inclusive scan of the per-lane energy for the roll-off search (aubio.rs:36-58)
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    const float excl = incl - run;
    const float thr = total * 0.95f;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) cnt += ((excl + c[i]) < thr) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    const float freq = (float)SAMPLE_RATE / 512.f;  // bin_to_freq, aubio.rs:68-71
    float bin = (total == 0.f) ? 0.f : (float)min(cnt + 1, 256);
    rolloff = freq * bin;
    centroid = (s1 == 0.f) ? 0.f : freq * fmaxf(sw / s1, 0.f);
    if (zero) {
        flatness = 0.f;
    } else {
        const float gm = exp2f((log2f((float)mant) + (float)ex) / 256.f - (1023.f + 500.f) / 8.f);
        flatness = gm / (s1 / 256.f);
    }
}

// ---- both frames of a pair at once (experimental, VARIANT_PV_PAIRDESC) --------------------------------------
// Transposed reductions: the first butterfly step sends each half-warp the OTHER frame's partial, so lanes 0..15
// finish frame A's total and lanes 16..31 frame B's with one shuffle per step instead of two.  The tree is the
// one warp_sum / warp_prod walk (a_l + a_(l^16) first, then offsets 8, 4, 2, 1 inside the half), so the totals
// are bit-identical to theirs.
__device__ __forceinline__ float pair_sum(float a, float b, int lane) {
    const bool hi = (lane & 16) != 0;
    float mine = hi ? b : a;
    mine += __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    return mine;
}
__device__ __forceinline__ double pair_prod(double a, double b, int lane) {
    const bool hi = (lane & 16) != 0;
    double mine = hi ? b : a;
    mine *= __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mine *= __shfl_xor_sync(0xffffffffu, mine, o);
    return mine;
}

// per-lane part of frame_descriptors for one frame
struct LanePart {
    float s1, sw, run, c[8];
    double mant;
    int ex;
    bool is_zero;
};
__device__ __forceinline__ LanePart lane_part(const float (&v)[8], int lane) {
    LanePart p;
    p.s1 = 0.f;
    p.sw = 0.f;
    p.run = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p.s1 += v[i];
        p.sw += (float)(8 * lane + i) * v[i];
        p.run += v[i] * v[i];
        p.c[i] = p.run;
    }
    double m = ((double)v[0] * (double)v[1]) * ((double)v[2] * (double)v[3]);
    m *= 3.273390607896142e150;
    m *= ((double)v[4] * (double)v[5]) * ((double)v[6] * (double)v[7]);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
    p.is_zero = (m == 0.0);
    p.ex = (int)(bits >> 52);
    p.mant = __longlong_as_double((long long)((bits & 0xFFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    return p;
}
// roll-off count of one frame (aubio.rs:36-58): warp scan of the per-lane energy, as frame_descriptors
__device__ __forceinline__ void rolloff_count(const LanePart &p, int lane, float &total, int &cnt) {
    float incl = p.run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    const float excl = incl - p.run;
    const float thr = total * 0.95f;
    int c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c += ((excl + p.c[i]) < thr) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, c);
}
// Descriptors of frames A (returned in lanes 0..15) and B (lanes 16..31): same arithmetic as two calls of
// frame_descriptors, with the reductions transposed and the scalar finish (two IEEE divisions, log2f, exp2f:
// 78 instructions per frame) executed once per pair.
__device__ __forceinline__ void pair_descriptors(const float (&va)[8], const float (&vb)[8], int lane,
                                                 float &centroid, float &rolloff, float &flatness) {
    const LanePart a = lane_part(va, lane), b = lane_part(vb, lane);
    const bool hi = (lane & 16) != 0;
    const int zero_a = __any_sync(0xffffffffu, a.is_zero), zero_b = __any_sync(0xffffffffu, b.is_zero);
    const int ex_a = __reduce_add_sync(0xffffffffu, a.ex), ex_b = __reduce_add_sync(0xffffffffu, b.ex);
    const double mant = pair_prod(a.mant, b.mant, lane);
    const float s1 = pair_sum(a.s1, b.s1, lane);
    const float sw = pair_sum(a.sw, b.sw, lane);
    float total_a, total_b;
    int cnt_a, cnt_b;
    rolloff_count(a, lane, total_a, cnt_a);
    rolloff_count(b, lane, total_b, cnt_b);
    const float total = hi ? total_b : total_a;
    const int cnt = hi ? cnt_b : cnt_a;
    const int zero = hi ? zero_b : zero_a;
    const int ex = hi ? ex_b : ex_a;
    const float freq = (float)SAMPLE_RATE / 512.f;
    float bin = (total == 0.f) ? 0.f : (float)min(cnt + 1, 256);
    rolloff = freq * bin;
    centroid = (s1 == 0.f) ? 0.f : freq * fmaxf(sw / s1, 0.f);
    if (zero) {
        flatness = 0.f;
    } else {
        const float gm = exp2f((log2f((float)mant) + (float)ex) / 256.f - (1023.f + 500.f) / 8.f);
        flatness = gm / (s1 / 256.f);
    }
}

// WITH_DESC: timbral descriptors + flux (the analysis path).
// WITH_MAGS: materialise the 257 tempo-frame magnitudes (STFT micro-benchmark,
//            BASELINE.json config 3 = PVocTempo framing: 512 / hop 256).
#ifndef K1_MINBLOCKS
#define K1_MINBLOCKS 2
#endif
// TWPROD (VARIANT_PV_TWPROD) and PAIRDESC (VARIANT_PV_PAIRDESC) are experimental cuts, off by default.
// PAIRDESC: pair_descriptors + the data kept 2^30 above its level so that the magnitudes need MUFU.SQRT only
// (approx_sqrtf_ftz): every result is bit-identical to the default kernel's unless an intermediate of the
// default kernel is denormal.
// ZS: padding shift of the natural-order tile (pvoc512.cuh zpos_s; 3 = the measured layout, 4 = VARIANT_PV_ZPOS4).
template <bool WITH_DESC, bool WITH_MAGS, bool TWPROD = false, bool PAIRDESC = false, int ZS = 3>
__global__ void __launch_bounds__(256, K1_MINBLOCKS)
pvoc512_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
               const unsigned int *__restrict__ item_prefix, int n_songs, unsigned int total_items,
               int pairs_per_item, PvocTables tab, float *__restrict__ centroid,
               float *__restrict__ rolloff, float *__restrict__ flatness, float *__restrict__ flux,
               float *__restrict__ mags_out) {
    __shared__ float s_win[512];
    __shared__ cpx s_twA[16 * 32];
    __shared__ cpx s_ex[8][pv::EXCH_CPX];

    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        s_win[i] = tab.win[i];
        s_twA[i] = tab.twA[i];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * 8u + (unsigned)warp;
    if (item >= total_items) return;
    const int si = find_song(item_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const int j0 = (int)(item - item_prefix[si]) * pairs_per_item;
    const int j1 = min(j0 + pairs_per_item, (int)sd.n_t);
    const float *x = pcm + sd.pcm_off;
    const int n = (int)sd.n;
    cpx *S = s_ex[warp];

    float win_a[16];  // w[lane + 32*n1]
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) win_a[n1] = s_win[lane + 32 * n1] * (PAIRDESC ? 1073741824.f : 1.f);  // 2^30: exact
    cpx tw1 = cpx{1.f, 0.f}, tw2 = tw1, tw4 = tw1, tw8 = tw1;  // W512^(lane k1), k1 = 1, 2, 4, 8 (TWPROD)
    if constexpr (TWPROD) {
        tw1 = s_twA[1 * 32 + lane];
        tw2 = s_twA[2 * 32 + lane];
        tw4 = s_twA[4 * 32 + lane];
        tw8 = s_twA[8 * 32 + lane];
    }

    float old[8];  // previous tempo frame's magnitudes of this lane's bins
#pragma unroll
    for (int i = 0; i < 8; i++) old[i] = 0.f;
    float old256 = 0.f;

    float s[20];  // s[m] = x[256*j - 384 + lane + 32*m]
    const int jstart = (j0 > 0) ? j0 - 1 : j0;  // halo pair: only to seed `old`
    {
        const int base = 256 * jstart - 384 + lane;
#pragma unroll
        for (int m = 0; m < 20; m++) {
            const int idx = base + 32 * m;  // rows 12..19 are >= 0 and < n for every valid pair
            s[m] = (idx >= 0 && idx < n) ? __ldg(x + idx) : 0.f;
        }
    }
    for (int j = jstart; j < j1; j++) {
        // Two real frames ride one complex FFT (A in re, B in im); untangling leaks
        // eps*max(|A|,|B|) of rounding noise into the weaker frame.  B is therefore pre-scaled by a
        // power of two (exact) to A's level and scaled back afterwards, and a frame whose windowed
        // samples are all exactly zero (digital silence) is forced to exact zeros, which is what
        // the reference's separate transforms produce.
        float pka = 0.f, pkb = 0.f;
        cpx r[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            r[n1] = pmul(cpx{s[n1], s[n1 + 4]}, cpx{win_a[n1], win_a[n1]});
            pka = fmaxf(pka, fabsf(r[n1].x));
            pkb = fmaxf(pkb, fabsf(r[n1].y));
        }
        // slide the window by 256 samples and fetch the 8 new rows of the NEXT pair now: the loads are in
        // flight during this pair's FFT instead of stalling the window multiply at the top of the loop
        // (ncu: 10 % of the kernel's stall samples sat on that first multiply)
#pragma unroll
        for (int m = 0; m < 12; m++) s[m] = s[m + 8];
        if (j + 1 < j1) {
            const float *px = x + 256 * (j + 1) - 384 + lane + 32 * 12;  // 256 (j+1) + lane + 32 m < n for j + 1 < n_t
#pragma unroll
            for (int m = 0; m < 8; m++) s[12 + m] = __ldg(px + 32 * m);
        }
        const unsigned int ua = __reduce_max_sync(0xffffffffu, __float_as_uint(pka));
        const unsigned int ub = __reduce_max_sync(0xffffffffu, __float_as_uint(pkb));
        int sh = (int)(ua >> 23) - (int)(ub >> 23);  // exponent difference of the two peaks
        sh = (ua == 0u || ub == 0u) ? 0 : max(-60, min(60, sh));
        const float gscale = __uint_as_float((unsigned)(127 + sh) << 23);
        const float ginv = __uint_as_float((unsigned)(127 - sh) << 23);
        if (sh != 0) {  // warp-uniform; the common case (similar levels) skips the rescale
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) r[n1].y *= gscale;
        }
        if constexpr (TWPROD) pv::phase_a_prod(lane, r, tw1, tw2, tw4, tw8, S);
        else pv::phase_a(lane, r, s_twA, S);
        __syncwarp();
        pv::phase_b_load(lane, r, S);
        __syncwarp();
        pv::phase_b_fft(lane, r);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const cpx o = shfl_xor_cpx(r[q], 16);
            const cpx z = pv::phase_b_combine(lane, r[q], o);
            S[pv::zpos_s<ZS>(pv::bin_of(lane, q))] = z;
        }
        __syncwarp();
        // natural-order epilogue: lane owns bins 8*lane .. 8*lane+7 of both frames
        float ma[8], mb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = 8 * lane + i;
            const cpx zk = S[pv::zpos_s<ZS>(k)];
            const cpx zm = S[pv::zpos_s<ZS>((512 - k) & 511)];
            pv::untangle_mag<true, PAIRDESC>(zk, zm, ma[i], mb[i]);  // 2|A|, 2|B|: halved by ka / kb below
        }
        const cpx zn = S[pv::zpos_s<ZS>(256)];  // Nyquist: A = |Re|, B = |Im|
        float nyq_a = 2.f * fabsf(zn.x), nyq_b = 2.f * fabsf(zn.y);
        if (lane == 0) {  // DC: abs(re), aubio.rs:240 / :403
            const cpx z0 = S[0];
            ma[0] = 2.f * fabsf(z0.x);
            mb[0] = 2.f * fabsf(z0.y);
        }
        __syncwarp();  // S is rewritten by the next pair's phase A
        constexpr float kHalf = PAIRDESC ? 0.5f / 1073741824.f : 0.5f;  // takes the window's 2^30 out again
        const float ka = (ua == 0u) ? 0.f : kHalf;  // powers of two: exact
        const float kb = (ub == 0u) ? 0.f : kHalf * ginv;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ma[i] *= ka;
            mb[i] *= kb;
        }
        nyq_a *= ka;
        nyq_b *= kb;

        const bool emit = (j >= j0);
        if (WITH_MAGS && emit) {
            float *o = mags_out + ((size_t)sd.t_off + (size_t)j) * 257u;
#pragma unroll
            for (int i = 0; i < 8; i++) o[8 * lane + i] = mb[i];
            if (lane == 31) o[256] = nyq_b;
        }
        if (WITH_DESC) {
            // SpecFlux over the 257 correct bins of the tempo frame (aubio.rs:455-467)
            float fl = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                fl += (mb[i] > old[i]) ? (mb[i] - old[i]) : 0.f;
                old[i] = mb[i];
            }
            if (lane == 31) {
                fl += (nyq_b > old256) ? (nyq_b - old256) : 0.f;
                old256 = nyq_b;
            }
            fl = warp_sum(fl);
            if (emit) {
                if (lane == 0) flux[sd.t_off + j] = fl;
                // timbral norms: the 256-bin cvec with Nyquist stored in slot 255 (aubio.rs:255-261)
                if (lane == 31) {
                    ma[7] = nyq_a;
                    mb[7] = nyq_b;
                }
                float c, ro, fa;
                const int fa_idx = 2 * j, fb_idx = 2 * j + 1;
                if constexpr (PAIRDESC) {
                    pair_descriptors(ma, mb, lane, c, ro, fa);
                    const int fidx = 2 * j + (lane >> 4);  // lanes 0..15 hold frame 2j's values, 16..31 frame 2j+1's
                    if ((lane & 15) == 0 && fidx < (int)sd.n_s) {
                        centroid[sd.s_off + fidx] = c;
                        rolloff[sd.s_off + fidx] = ro;
                        flatness[sd.s_off + fidx] = fa;
                    }
                } else {
                if (fa_idx < (int)sd.n_s) {
                    frame_descriptors(ma, lane, c, ro, fa);
                    if (lane == 0) {
                        centroid[sd.s_off + fa_idx] = c;
                        rolloff[sd.s_off + fa_idx] = ro;
                        flatness[sd.s_off + fa_idx] = fa;
                    }
                }
                if (fb_idx < (int)sd.n_s) {
                    frame_descriptors(mb, lane, c, ro, fa);
                    if (lane == 0) {
                        centroid[sd.s_off + fb_idx] = c;
                        rolloff[sd.s_off + fb_idx] = ro;
                        flatness[sd.s_off + fb_idx] = fa;
                    }
                }
                }
            }
        } else if (WITH_MAGS) {
            (void)nyq_a;
        }
    }
}

// ---------------------------------------------------------------------------
// pvoc512v2_kernel (round 2): the same work as pvoc512_kernel -- one warp walks a run of frame pairs, timbral
// frames 2j / 2j+1 ride one complex 512-point FFT, each sample is fetched once through the 20-row sliding register
// window -- re-cut around what the round-1 captures showed the kernel to be bound by (issue slots 70 %, L1 /
// shared-memory data pipe 67 %: 1 111 instructions and ~260 wavefronts per pair with every round-1 cut on):
//   * phase B is a decimation-in-FREQUENCY split of the 32-point DFT over n2: lane (p, k1) forms
//     u[n2] = y[n2] + (-1)^p y[n2 + 16], rotates by W32^(n2 p) and runs ONE radix-16, which leaves it the bins
//     k = k1 + 16 (2 q + p), q = 0..15 -- no radix-2 shuffle stage (32 SHFL + 16 FFMA2 per pair gone);
//   * the mirror bins Z[512 - k] of a lane's eight LOW bins (q < 8) are the eight HIGH bins of ONE other lane
//     (1 - p, 16 - k1), same slot order: 16 SHFL replace the natural-order round trip through shared memory
//     (16 STS.64 + 18 LDS.64, 2-way conflicts), and every lane untangles exactly eight bins of both frames;
//     the bins a lane ends up with are lane + 32 i: a row of magnitudes is written as 128-byte warp stores;
//   * the per-frame descriptors leave the FFT loop: magnitudes of four pairs (eight frames) go to a per-warp tile
//     in shared memory and are reduced FOUR LANES PER FRAME, 64 consecutive bins per lane, serially in registers:
//     2 shuffle steps per quantity instead of 5, one scalar finish (two IEEE divisions, log2f, exp2f) per eight
//     frames instead of per pair; geometric_mean's per-8-bin f64 products (utils.rs:104-111) keep their order.
// Shared memory: 4.1 KB exchange tile + 8.5 KB magnitude tile per warp (101 KB per CTA, two CTAs per SM).
// ---------------------------------------------------------------------------
namespace pv2 {
constexpr int ROW = 33;                      // padded row (cpx) of the 16 x 32 exchange tile: S[k1][n2]
constexpr int EXCH_CPX = 16 * ROW;           // 528 cpx = 4224 B
constexpr int TILE_PAIRS = 4;                // pairs (x 2 frames) per descriptor tile
constexpr int TILE_ROW = 272;                // floats per frame row: bin k at k + 4 (k >> 6), Nyquist at NYQ_POS
constexpr int NYQ_POS = 268;
constexpr int WARP_SMEM_BYTES = EXCH_CPX * 8 + 2 * TILE_PAIRS * TILE_ROW * 4;  // 12 928
static_assert(WARP_SMEM_BYTES % 16 == 0 && (EXCH_CPX * 8) % 16 == 0, "128-bit tile loads");

template <int N2>
BLISS_HD void tw32_nat(cpx (&u)[16]) {  // u[n2] *= W32^n2, natural order
    if constexpr (N2 < 16) {
        u[N2] = mul_tw<N2, 32>(u[N2]);
        tw32_nat<N2 + 1>(u);
    }
}

// Descriptors of the frames of one tile (rows 0 .. 2 n_pairs - 1), four lanes per frame: lane = 4 f + g reduces
// bins 64 g .. 64 g + 63 of row f.  spectral_centroid / spectral_rolloff (aubio.rs:16-58, clamp timbral.rs:184-187),
// flatness = geometric_mean / mean (utils.rs:101-117, timbral.rs:196-208); the 256-bin cvec carries the Nyquist
// magnitude in slot 255 (aubio.rs:255-261).
__device__ __forceinline__ void tile_descriptors(const float *T, int n_pairs, int lane, int frame0, int n_s,
                                                 float *__restrict__ centroid, float *__restrict__ rolloff,
                                                 float *__restrict__ flatness) {
    const int f = lane >> 2, g = lane & 3;
    const float *row = T + f * TILE_ROW;
    const float4 *q = reinterpret_cast<const float4 *>(row + 68 * g);  // tpos(64 g) = 64 g + 4 g
    const float nyq = row[NYQ_POS];
    float s1 = 0.f, sw = 0.f, run = 0.f, c8[8];
    double mant = 1.0;
    int ex = 0;
    bool zero = false;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const float4 a = q[2 * t], b = q[2 * t + 1];
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (t == 7 && g == 3) v[7] = nyq;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            s1 += v[i];
            sw = fmaf((float)(8 * t + i), v[i], sw);
            run = fmaf(v[i], v[i], run);
        }
        c8[t] = run;
        // geometric_mean's per-group product, utils.rs:104-111 (order kept)
        double m = ((double)v[0] * (double)v[1]) * ((double)v[2] * (double)v[3]);
        m *= 3.273390607896142e150;
        m *= ((double)v[4] * (double)v[5]) * ((double)v[6] * (double)v[7]);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
        zero = zero || (m == 0.0);
        ex += (int)(bits >> 52);
        mant *= __longlong_as_double((long long)((bits & 0xFFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    }
    sw = fmaf((float)(64 * g), s1, sw);  // sum (64 g + i) v
    const unsigned int full = 0xffffffffu;
    s1 += __shfl_xor_sync(full, s1, 1);
    s1 += __shfl_xor_sync(full, s1, 2);
    sw += __shfl_xor_sync(full, sw, 1);
    sw += __shfl_xor_sync(full, sw, 2);
    ex += __shfl_xor_sync(full, ex, 1);
    ex += __shfl_xor_sync(full, ex, 2);
    mant *= __shfl_xor_sync(full, mant, 1);
    mant *= __shfl_xor_sync(full, mant, 2);
    const unsigned int zb = __ballot_sync(full, zero);
    const bool any_zero = ((zb >> (lane & ~3)) & 0xFu) != 0u;
    // roll-off: number of bins whose inclusive energy prefix stays below 0.95 of the total (aubio.rs:36-58);
    // the prefix is non-decreasing, so each lane counts inside its own 64 bins and the counts add up
    float incl = run;
    {
        float t1 = __shfl_up_sync(full, incl, 1);
        if (g >= 1) incl += t1;
        t1 = __shfl_up_sync(full, incl, 2);
        if (g >= 2) incl += t1;
    }
    const float total = __shfl_sync(full, incl, lane | 3);
    float excl = __shfl_up_sync(full, incl, 1);
    if (g == 0) excl = 0.f;
    const float thr = total * 0.95f;
    int gb = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) gb += ((excl + c8[t]) < thr) ? 1 : 0;
    const int tg = min(gb, 7);  // the group that holds the crossing (re-scanned with the arithmetic of the first pass)
    float r2 = 0.f;
#pragma unroll
    for (int t = 0; t < 7; t++)
        if (tg > t) r2 = c8[t];
    int within = 0;
    {
        const float4 a = q[2 * tg], b = q[2 * tg + 1];
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (tg == 7 && g == 3) v[7] = nyq;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            r2 = fmaf(v[i], v[i], r2);
            within += ((excl + r2) < thr) ? 1 : 0;
        }
    }
    int cnt = (gb == 8) ? 64 : 8 * gb + within;
    cnt += __shfl_xor_sync(full, cnt, 1);
    cnt += __shfl_xor_sync(full, cnt, 2);
    const float freq = (float)SAMPLE_RATE / 512.f;  // bin_to_freq, aubio.rs:68-71
    const float bin = (total == 0.f) ? 0.f : (float)min(cnt + 1, 256);
    const float ro = freq * bin;
    const float ce = (s1 == 0.f) ? 0.f : freq * fmaxf(sw / s1, 0.f);
    float fl = 0.f;
    if (!any_zero) {
        const float gm = exp2f((log2f((float)mant) + (float)ex) / 256.f - (1023.f + 500.f) / 8.f);
        fl = gm / (s1 / 256.f);
    }
    const int fidx = frame0 + f;
    if (g == 0 && f < 2 * n_pairs && fidx < n_s) {
        centroid[fidx] = ce;
        rolloff[fidx] = ro;
        flatness[fidx] = fl;
    }
}
}  // namespace pv2

__global__ void __launch_bounds__(256, 2)
pvoc512v2_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                 const unsigned int *__restrict__ item_prefix, int n_songs, unsigned int total_items,
                 int pairs_per_item, PvocTables tab, float *__restrict__ centroid, float *__restrict__ rolloff,
                 float *__restrict__ flatness, float *__restrict__ flux, float *__restrict__ /*mags_out: unused*/) {
#ifdef BLISS_HOST_EMUL
    unsigned char *pv2_smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char pv2_smem[];
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * 8u + (unsigned)warp;
    if (item >= total_items) return;
    const int si = find_song(item_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const int j0 = (int)(item - item_prefix[si]) * pairs_per_item;
    const int j1 = min(j0 + pairs_per_item, (int)sd.n_t);
    const float *x = pcm + sd.pcm_off;
    const int n = (int)sd.n;
    cpx *S = reinterpret_cast<cpx *>(pv2_smem + (size_t)warp * pv2::WARP_SMEM_BYTES);
    float *T = reinterpret_cast<float *>(S + pv2::EXCH_CPX);

    float win_a[16];  // 2^30 w[lane + 32 n1]: the magnitudes then need MUFU.SQRT only (exact power of two, taken out below)
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) win_a[n1] = __ldg(tab.win + lane + 32 * n1) * 1073741824.f;
    const float2 *twg = reinterpret_cast<const float2 *>(tab.twA);  // W512^(lane k1), k1 = 1, 2, 4, 8
    const float2 g1 = __ldg(twg + 1 * 32 + lane), g2 = __ldg(twg + 2 * 32 + lane), g4 = __ldg(twg + 4 * 32 + lane),
                 g8 = __ldg(twg + 8 * 32 + lane);
    const cpx tw1 = cpx{g1.x, g1.y}, tw2 = cpx{g2.x, g2.y}, tw4 = cpx{g4.x, g4.y}, tw8 = cpx{g8.x, g8.y};
    const int k1 = lane & 15, p = lane >> 4;
    const float sgn = p ? -1.f : 1.f;
    const int mirror_lane = (k1 != 0) ? 16 * (1 - p) + 16 - k1 : lane;  // holds Z[512 - k] of this lane's low bins

    float old[8];  // previous tempo frame's magnitudes of bins lane + 32 i
#pragma unroll
    for (int i = 0; i < 8; i++) old[i] = 0.f;
    float old256 = 0.f;

    float s[20];  // s[m] = x[256*j - 384 + lane + 32*m]
    const int jstart = (j0 > 0) ? j0 - 1 : j0;  // halo pair: only to seed `old`
    {
        const int base = 256 * jstart - 384 + lane;
#pragma unroll
        for (int m = 0; m < 20; m++) {
            const int idx = base + 32 * m;  // rows 12..19 are >= 0 and < n for every valid pair
            s[m] = (idx >= 0 && idx < n) ? __ldg(x + idx) : 0.f;
        }
    }
    int trow = 0, jt0 = j0;  // pairs in the open tile, its first pair
    for (int j = jstart; j < j1; j++) {
        // frame packing, power-of-two pre-scaling of B and exact zeros for digital silence: as pvoc512_kernel
        float pka = 0.f, pkb = 0.f;
        cpx r[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            r[n1] = pmul(cpx{s[n1], s[n1 + 4]}, cpx{win_a[n1], win_a[n1]});
            pka = fmaxf(pka, fabsf(r[n1].x));
            pkb = fmaxf(pkb, fabsf(r[n1].y));
        }
#pragma unroll
        for (int m = 0; m < 12; m++) s[m] = s[m + 8];
        if (j + 1 < j1) {
            const float *px = x + 256 * (j + 1) - 384 + lane + 32 * 12;  // 256 (j+1) + lane + 32 m < n for j + 1 < n_t
#pragma unroll
            for (int m = 0; m < 8; m++) s[12 + m] = __ldg(px + 32 * m);
        }
        const unsigned int ua = __reduce_max_sync(0xffffffffu, __float_as_uint(pka));
        const unsigned int ub = __reduce_max_sync(0xffffffffu, __float_as_uint(pkb));
        int sh = (int)(ua >> 23) - (int)(ub >> 23);  // exponent difference of the two peaks
        sh = (ua == 0u || ub == 0u) ? 0 : max(-60, min(60, sh));
        const float gscale = __uint_as_float((unsigned)(127 + sh) << 23);
        const float ginv = __uint_as_float((unsigned)(127 - sh) << 23);
        if (sh != 0) {  // warp-uniform; the common case (similar levels) skips the rescale
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) r[n1].y *= gscale;
        }
        pv::phase_a_prod(lane, r, tw1, tw2, tw4, tw8, S);  // S[k1][n2] = W512^(n2 k1) sum_n1 z[n2 + 32 n1] W16^(n1 k1)  (ROW = 33)
        __syncwarp();
        // phase B: X[k1 + 16 (2 q + p)] = sum_{n2 < 16} (y[n2] + (-1)^p y[n2 + 16]) W32^(n2 p) W16^(n2 q)
        {
            const cpx *yrow = S + k1 * pv2::ROW;
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) r[n2] = pfma(yrow[n2 + 16], cpx{sgn, sgn}, yrow[n2]);
        }
        __syncwarp();  // S is rewritten by the next pair's phase A
        if (p) pv2::tw32_nat<1>(r);
        fft_dif<16>(r);  // r[slot] = X[k1 + 16 (2 bitrev(slot) + p)]
        // the eight low bins k = lane + 32 i (q = i < 8) and their mirrors 512 - k = slot q = 15 - i of mirror_lane
        float ma[8], mb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const cpx hi = r[bitrev(15 - i, 4)];
            cpx zm = cpx{__shfl_sync(0xffffffffu, hi.x, mirror_lane), __shfl_sync(0xffffffffu, hi.y, mirror_lane)};
            if (i > 0 && lane == 0) zm = r[bitrev(16 - i, 4)];  // k = 32 i: 512 - k = 16 * 2 (16 - i), this lane's own slot
            pv::untangle_mag<true, true>(r[bitrev(i, 4)], zm, ma[i], mb[i]);  // 2|A|, 2|B|: halved by ka / kb below
        }
        const cpx zn = r[bitrev(8, 4)];  // lane 0: Z[256], the Nyquist bin: A = |Re|, B = |Im| (aubio.rs:255-261, :403-405)
        float nyq_a = 2.f * fabsf(zn.x), nyq_b = 2.f * fabsf(zn.y);
        if (lane == 0) {  // DC: abs(re), aubio.rs:240 / :403
            ma[0] = 2.f * fabsf(r[0].x);
            mb[0] = 2.f * fabsf(r[0].y);
        }
        constexpr float kHalf = 0.5f / 1073741824.f;  // takes the window's 2^30 out again
        const float ka = (ua == 0u) ? 0.f : kHalf;    // powers of two: exact
        const float kb = (ub == 0u) ? 0.f : kHalf * ginv;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ma[i] *= ka;
            mb[i] *= kb;
        }
        nyq_a *= ka;
        nyq_b *= kb;
        // SpecFlux over the 257 correct bins of the tempo frame (aubio.rs:455-467)
        float fl = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            fl += (mb[i] > old[i]) ? (mb[i] - old[i]) : 0.f;
            old[i] = mb[i];
        }
        if (lane == 0) {
            fl += (nyq_b > old256) ? (nyq_b - old256) : 0.f;
            old256 = nyq_b;
        }
        fl = warp_sum(fl);
        if (j >= j0) {  // not the halo pair
            if (lane == 0) flux[sd.t_off + j] = fl;
            float *Ta = T + (2 * trow) * pv2::TILE_ROW + lane, *Tb = Ta + pv2::TILE_ROW;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                Ta[32 * i + 4 * (i >> 1)] = ma[i];  // bin lane + 32 i at tpos = k + 4 (k >> 6)
                Tb[32 * i + 4 * (i >> 1)] = mb[i];
            }
            if (lane == 0) {
                Ta[pv2::NYQ_POS] = nyq_a;
                Tb[pv2::NYQ_POS] = nyq_b;
            }
            trow++;
            if (trow == pv2::TILE_PAIRS || j + 1 == j1) {
                __syncwarp();
                pv2::tile_descriptors(T, trow, lane, 2 * jt0, (int)sd.n_s, centroid + sd.s_off, rolloff + sd.s_off,
                                      flatness + sd.s_off);
                __syncwarp();
                jt0 += trow;
                trow = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// STFT micro-benchmark kernel for the hop-256 framing (BASELINE.json config 3 = PVocTempo::do_,
// src/aubio.rs:338-425), experimental (VARIANT_STFT_PAIRS): tempo frames j and j+1 ride ONE complex
// 512-point FFT and BOTH are wanted, whereas pvoc512_kernel<false, true> transforms the hop-128 pair
// (timbral frames 2j, 2j+1) and throws the even one away -- half the transforms per track.  Each lane
// untangles bins lane + 32 i (the natural-order tile in shared memory allows any assignment), so a frame's
// 257 magnitudes leave as eight 128-byte warp stores instead of eight stores at a 32-byte lane stride.
//   frame j  = x[256 j - 256 .. 256 j + 256),  frame j+1 = x[256 j .. 256 j + 512)   (zeros before the song)
//   window rows s[r] = x[256 j - 256 + lane + 32 r], r = 0..23: A row n1 = s[n1], B row n1 = s[n1 + 8];
//   the next pair starts 16 rows further on.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
stft512_pairs_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                     const unsigned int *__restrict__ item_prefix, int n_songs, unsigned int total_items,
                     int frames_per_item, PvocTables tab, float *__restrict__ mags_out) {
    __shared__ float s_win[512];
    __shared__ cpx s_twA[16 * 32];
    __shared__ cpx s_ex[8][pv::EXCH_CPX];

    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        s_win[i] = tab.win[i];
        s_twA[i] = tab.twA[i];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * 8u + (unsigned)warp;
    if (item >= total_items) return;
    const int si = find_song(item_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const int j0 = (int)(item - item_prefix[si]) * frames_per_item;
    const int j1 = min(j0 + frames_per_item, (int)sd.n_t);
    const float *x = pcm + sd.pcm_off;
    const int n = (int)sd.n;
    cpx *S = s_ex[warp];

    float win_a[16];  // w[lane + 32*n1]
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) win_a[n1] = s_win[lane + 32 * n1];

    float s[24];
    {
        const int base = 256 * j0 - 256 + lane;
#pragma unroll
        for (int m = 0; m < 24; m++) {
            const int idx = base + 32 * m;
            s[m] = (idx >= 0 && idx < n) ? __ldg(x + idx) : 0.f;
        }
    }
    for (int j = j0; j < j1; j += 2) {
        // same frame packing as pvoc512_kernel: B pre-scaled by an exact power of two to A's level, an
        // all-zero windowed frame forced to exact zeros (see there)
        float pka = 0.f, pkb = 0.f;
        cpx r[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            r[n1] = pmul(cpx{s[n1], s[n1 + 8]}, cpx{win_a[n1], win_a[n1]});
            pka = fmaxf(pka, fabsf(r[n1].x));
            pkb = fmaxf(pkb, fabsf(r[n1].y));
        }
        // slide by 512 samples; the 16 new rows of the next pair are requested before this pair's FFT
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = s[m + 16];
        if (j + 2 < j1) {
            const int base = 256 * (j + 2) - 256 + lane + 32 * 8;  // >= 0; frame j+2 is valid: rows 8..15 are inside the song
            const float *px = x + base;
#pragma unroll
            for (int m = 0; m < 8; m++) s[8 + m] = __ldg(px + 32 * m);
            if (j + 3 < (int)sd.n_t) {  // frame j+3 exists: its last 8 rows are inside the song too
#pragma unroll
                for (int m = 8; m < 16; m++) s[8 + m] = __ldg(px + 32 * m);
            } else {
#pragma unroll
                for (int m = 8; m < 16; m++) {
                    const int idx = base + 32 * m;
                    s[8 + m] = (idx < n) ? __ldg(x + idx) : 0.f;
                }
            }
        }
        const unsigned int ua = __reduce_max_sync(0xffffffffu, __float_as_uint(pka));
        const unsigned int ub = __reduce_max_sync(0xffffffffu, __float_as_uint(pkb));
        int sh = (int)(ua >> 23) - (int)(ub >> 23);
        sh = (ua == 0u || ub == 0u) ? 0 : max(-60, min(60, sh));
        const float gscale = __uint_as_float((unsigned)(127 + sh) << 23);
        const float ginv = __uint_as_float((unsigned)(127 - sh) << 23);
        if (sh != 0) {
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) r[n1].y *= gscale;
        }
        pv::phase_a(lane, r, s_twA, S);
        __syncwarp();
        pv::phase_b_load(lane, r, S);
        __syncwarp();
        pv::phase_b_fft(lane, r);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const cpx o = shfl_xor_cpx(r[q], 16);
            const cpx z = pv::phase_b_combine(lane, r[q], o);
            S[pv::zpos_s<0>(pv::bin_of(lane, q))] = z;  // unpadded: conflict-free for bins lane + 32 i
        }
        __syncwarp();
        // lane owns bins lane + 32 i of both frames
        float ma[8], mb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int k = lane + 32 * i;
            const cpx zk = S[pv::zpos_s<0>(k)];
            const cpx zm = S[pv::zpos_s<0>((512 - k) & 511)];
            pv::untangle_mag<true>(zk, zm, ma[i], mb[i]);  // 2|A|, 2|B|
        }
        const cpx zn = S[pv::zpos_s<0>(256)];  // Nyquist: A = |Re|, B = |Im| (aubio.rs:403-405)
        float nyq_a = 2.f * fabsf(zn.x), nyq_b = 2.f * fabsf(zn.y);
        if (lane == 0) {  // DC: abs(re), aubio.rs:403
            const cpx z0 = S[0];
            ma[0] = 2.f * fabsf(z0.x);
            mb[0] = 2.f * fabsf(z0.y);
        }
        __syncwarp();  // S is rewritten by the next pair's phase A
        const float ka = (ua == 0u) ? 0.f : 0.5f;
        const float kb = (ub == 0u) ? 0.f : 0.5f * ginv;
        float *oa = mags_out + ((size_t)sd.t_off + (size_t)j) * 257u;
#pragma unroll
        for (int i = 0; i < 8; i++) oa[lane + 32 * i] = ma[i] * ka;
        if (lane == 0) oa[256] = nyq_a * ka;
        if (j + 1 < j1) {
            float *ob = oa + 257;
#pragma unroll
            for (int i = 0; i < 8; i++) ob[lane + 32 * i] = mb[i] * kb;
            if (lane == 0) ob[256] = nyq_b * kb;
        }
    }
}

// ---------------------------------------------------------------------------
// K2: one warp per 1024-sample loudness chunk.
//   number_crossings        src/utils.rs:81-95      (one call over the whole song, song/mod.rs:470-474)
//   level_lin per chunk     src/misc.rs:12-18       (chunks(1024) incl. short tail, song/mod.rs:478)
//   256-sample block energy -> silence test of Tempo::do_ (aubio.rs:1258-1276, :1431)
// ---------------------------------------------------------------------------
constexpr int TD_CHUNKS_PER_WARP = 8;
// Footprint: 128 threads x <= 56 registers, no shared memory -- what an SM has LEFT beside three CTAs of the chroma STFT
// (3 x 128 x 152 registers, 218 of 227 KB of shared memory).  The idea was that this HBM-bound kernel, enqueued behind
// the compute-bound STFT, trickles through underneath it.  Measured (profiles/knobs_r02.md): it does not, at 56 / 48 /
// 40 / 32 registers, with or without a common carve-out; the step stays the sum of its kernels.  The cut is kept because
// it is the faster one alone (2.43 against 2.50 ms per 1024 tracks, same bits).
constexpr int TD_THREADS = 128;
#ifndef BLISS_TD_MIN_BLOCKS
#define BLISS_TD_MIN_BLOCKS 9  // <= 56 registers
#endif

__global__ void __launch_bounds__(TD_THREADS, BLISS_TD_MIN_BLOCKS)
timedomain_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                  const unsigned int *__restrict__ group_prefix, int n_songs,
                  unsigned int total_groups, float *__restrict__ loud_ms,
                  float *__restrict__ block_energy, unsigned int *__restrict__ zcr_count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * (unsigned)(TD_THREADS / 32) + (unsigned)warp;
    if (item >= total_groups) return;
    const int si = find_song(group_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const unsigned int grp = item - group_prefix[si];
    const float *x = pcm + sd.pcm_off;
    const unsigned int n = sd.n;
    unsigned int crossings = 0;
    const unsigned int ch0 = grp * TD_CHUNKS_PER_WARP;
    float prev_tail = (ch0 > 0) ? __ldg(x + ch0 * 1024u - 1) : 0.f;  // sample just before this run
#pragma unroll 1
    for (unsigned int ch = ch0; ch < min(ch0 + TD_CHUNKS_PER_WARP, sd.n_l); ch++) {
        const unsigned int base = ch * 1024u;
        const unsigned int len = min(1024u, n - base);
        const bool has_prev = base > 0;
        float eb[4] = {0.f, 0.f, 0.f, 0.f};
        const bool fast = len == 1024u && ((sd.pcm_off + base) & 3ull) == 0;
        // two halves of four rows: the four 16-byte loads of a half are issued before anything consumes them
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float v[4][4];
            if (fast) {
                const float4 *p4 = reinterpret_cast<const float4 *>(x + base) + lane + 128 * h;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 t = __ldg(p4 + 32 * k);
                    v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned int p = 4u * lane + 128u * (4 * h + k);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[k][i] = (p + i < len) ? __ldg(x + base + p + i) : 0.f;
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int k = 4 * h + kk;
                const unsigned int p = 4u * lane + 128u * k;
                float e = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) e += v[kk][i] * v[kk][i];
                eb[k >> 1] += e;
                // predecessor of v[k][0]: previous lane's last sample; lane 0 takes the previous row's tail
                float pred = __shfl_up_sync(0xffffffffu, v[kk][3], 1);
                if (lane == 0) pred = prev_tail;
                const bool have_pred = (k > 0) || (lane > 0) || has_prev;
                float before = pred;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (p + i < len) {
                        const bool valid_pair = (i > 0) || have_pred;
                        if (valid_pair && ((before > 0.f) != (v[kk][i] > 0.f))) crossings++;
                    }
                    before = v[kk][i];
                }
                prev_tail = __shfl_sync(0xffffffffu, v[kk][3], 31);
            }
        }
#pragma unroll
        for (int b = 0; b < 4; b++) eb[b] = warp_sum(eb[b]);
        if (lane == 0) loud_ms[sd.l_off + ch] = (eb[0] + eb[1] + eb[2] + eb[3]) / (float)len;
        if (lane < 4) {
            const unsigned int blk = ch * 4u + lane;
            if ((blk + 1u) * 256u <= n) {
                const float e = lane == 0 ? eb[0] : lane == 1 ? eb[1] : lane == 2 ? eb[2] : eb[3];
                block_energy[sd.e_off + blk] = e;
            }
        }
    }
    crossings = __reduce_add_sync(0xffffffffu, crossings);
    if (lane == 0 && crossings) atomicAdd(zcr_count + si, crossings);
}

// ---- launchers ---------------------------------------------------------------
int launch_pvoc512(const float *pcm, const SongDesc *songs, const unsigned int *item_prefix, int n_songs,
                   unsigned int total_items, int pairs_per_item, PvocTables tab, float *centroid,
                   float *rolloff, float *flatness, float *flux, int variant, cudaStream_t st) {
    if (total_items == 0) return 0;
    const unsigned int grid = (total_items + 7u) / 8u;
    auto go = [&](auto kern) {
        BLISS_LAUNCH(kern, grid, 256, 0, st, pcm, songs, item_prefix, n_songs, total_items, pairs_per_item, tab, centroid, rolloff,
                                   flatness, flux, nullptr);
    };
    if ((variant & VARIANT_PVOC_V1) == 0) {  // the round-2 kernel
        constexpr int smem = 8 * pv2::WARP_SMEM_BYTES;
#ifndef BLISS_HOST_EMUL
        // > 48 KB of dynamic shared memory is an opt-in, per device: set on every launch (cheap, and correct for any
        // number of devices and host threads)
        if (cudaFuncSetAttribute(pvoc512v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
#endif
        float *const none = nullptr;
        BLISS_LAUNCH(pvoc512v2_kernel, grid, 256, smem, st, pcm, songs, item_prefix, n_songs, total_items, pairs_per_item, tab,
                     centroid, rolloff, flatness, flux, none);
        return 1;
    }
    const bool tw = (variant & VARIANT_PV_TWPROD) != 0, pd = (variant & VARIANT_PV_PAIRDESC) != 0,
               z4 = (variant & VARIANT_PV_ZPOS4) != 0;
    if (z4) {
        if (tw && pd) go(pvoc512_kernel<true, false, true, true, 4>);
        else if (pd) go(pvoc512_kernel<true, false, false, true, 4>);
        else if (tw) go(pvoc512_kernel<true, false, true, false, 4>);
        else go(pvoc512_kernel<true, false, false, false, 4>);
    } else {
        if (tw && pd) go(pvoc512_kernel<true, false, true, true>);
        else if (pd) go(pvoc512_kernel<true, false, false, true>);
        else if (tw) go(pvoc512_kernel<true, false, true>);
        else go(pvoc512_kernel<true, false>);
    }
    return 1;
}

int launch_stft512_mags(const float *pcm, const SongDesc *songs, const unsigned int *item_prefix,
                        int n_songs, unsigned int total_items, int pairs_per_item, PvocTables tab,
                        float *mags, int variant, cudaStream_t st) {
    if (total_items == 0) return 0;
    const unsigned int grid = (total_items + 7u) / 8u;
    if (variant & VARIANT_STFT_PAIRS) {  // an item's `pairs_per_item` hop-128 pairs are as many hop-256 frames
        BLISS_LAUNCH(stft512_pairs_kernel, grid, 256, 0, st, pcm, songs, item_prefix, n_songs, total_items, pairs_per_item, tab, mags);
        return 1;
    }
    const auto mags_kernel = pvoc512_kernel<false, true>;  // (a template-id with a comma cannot be a macro argument)
    float *const none = nullptr;
    BLISS_LAUNCH(mags_kernel, grid, 256, 0, st, pcm, songs, item_prefix, n_songs, total_items, pairs_per_item, tab, none, none,
                 none, none, mags);
    return 1;
}

// group_prefix counts groups of TD_CHUNKS_PER_WARP (= 8) loudness chunks per song
int launch_timedomain(const float *pcm, const SongDesc *songs, const unsigned int *group_prefix,
                      int n_songs, unsigned int total_groups, float *loud_ms, float *block_energy,
                      unsigned int *zcr_count, cudaStream_t st) {
    if (total_groups == 0) return 0;
    const unsigned int per_cta = TD_THREADS / 32;
    const unsigned int grid = (total_groups + per_cta - 1u) / per_cta;
    BLISS_LAUNCH(timedomain_kernel, grid, TD_THREADS, 0, st, pcm, songs, group_prefix, n_songs, total_groups, loud_ms,
                                            block_energy, zcr_count);
    return 1;
}

// Every kernel of a wave asks for the SAME L1 / shared-memory split (all shared): kernels with different carve-outs
// cannot share an SM, and the latency-bound kernels of one chain are meant to run under the FFT kernels of the other
// (api.cu run_wave).  Called once per device from bliss_b200_init.
#ifndef BLISS_HOST_EMUL
#define BLISS_MAX_SHARED(kern) (void)cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)
#else
#define BLISS_MAX_SHARED(kern) (void)0
#endif
void configure_kernels_spectral() {
    BLISS_MAX_SHARED(pvoc512v2_kernel);
    BLISS_MAX_SHARED(timedomain_kernel);
    BLISS_MAX_SHARED(stft512_pairs_kernel);
}

}  // namespace bliss

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

__global__ void __launch_bounds__(256)
timedomain_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                  const unsigned int *__restrict__ group_prefix, int n_songs,
                  unsigned int total_groups, float *__restrict__ loud_ms,
                  float *__restrict__ block_energy, unsigned int *__restrict__ zcr_count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * 8u + (unsigned)warp;
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
        // all eight 16-byte loads of the lane are issued before anything consumes them
        float v[8][4];
        if (len == 1024u && ((sd.pcm_off + base) & 3ull) == 0) {
            const float4 *p4 = reinterpret_cast<const float4 *>(x + base) + lane;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float4 t = __ldg(p4 + 32 * k);
                v[k][0] = t.x; v[k][1] = t.y; v[k][2] = t.z; v[k][3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned int p = 4u * lane + 128u * k;
#pragma unroll
                for (int i = 0; i < 4; i++) v[k][i] = (p + i < len) ? __ldg(x + base + p + i) : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned int p = 4u * lane + 128u * k;
            float e = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) e += v[k][i] * v[k][i];
            eb[k >> 1] += e;
            // predecessor of v[k][0]: previous lane's last sample; lane 0 takes the previous row's tail
            float pred = __shfl_up_sync(0xffffffffu, v[k][3], 1);
            if (lane == 0) pred = prev_tail;
            const bool have_pred = (k > 0) || (lane > 0) || has_prev;
            float before = pred;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (p + i < len) {
                    const bool valid_pair = (i > 0) || have_pred;
                    if (valid_pair && ((before > 0.f) != (v[k][i] > 0.f))) crossings++;
                }
                before = v[k][i];
            }
            prev_tail = __shfl_sync(0xffffffffu, v[k][3], 31);
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
    const unsigned int grid = (total_groups + 7u) / 8u;
    BLISS_LAUNCH(timedomain_kernel, grid, 256, 0, st, pcm, songs, group_prefix, n_songs, total_groups, loud_ms,
                                            block_energy, zcr_count);
    return 1;
}

}  // namespace bliss

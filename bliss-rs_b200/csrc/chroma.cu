// chroma.cu -- the chroma chain of ChromaDesc::do_ / get_values (src/chroma.rs:73-132):
//   K3  reflect-padded 8192-point Hann STFT (utils.rs:26-64) + pip_track (chroma.rs:269-331)
//   K4  estimate_tuning / pitch_tuning (chroma.rs:334-391): exact median + 100-bin histogram
//   K5  chroma_stft contraction 12 x 4097 . 4097 x frames in f64 (chroma.rs:393-412)
//       fused with normalize_feature_sequence / extract_interval_features (:137-188)
//   plus the table of all 100 possible chroma filterbanks (chroma.rs:197-267).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "rfft8192.cuh"
#include "stft8192_v2.cuh"
#include "stft8192_v3.cuh"

namespace bliss {

// ---------------------------------------------------------------------------
// chroma filterbank for every tuning the estimator can return:
// tuning(idx) = (-50 + 100*0.01*idx)/100, idx = 0..99 (chroma.rs:357-358).
// Layout: table[idx][bin][12] f64 (the 12 chroma weights of one bin contiguous).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double tuning_of(int idx) { return (-50. + (100. * 0.01 * (double)idx)) / 100.; }

__device__ __forceinline__ double freq_bin(int i, double a440) {
    // Array::linspace(0, 22050, 8193)[i] / (a440/16) -> log2 * 12   (chroma.rs:210-214, utils.rs:119-129)
    const double stepf = (double)SAMPLE_RATE / 8192.0;
    double f = 0. + stepf * (double)i;
    f /= a440 / 16.;
    return log2(f) * 12.0;
}

__global__ void chroma_filter_table_kernel(double *__restrict__ table) {
    const int idx = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CH_BINS) return;
    const double tuning = tuning_of(idx);
    const double a440 = 440.0 * pow(2.0, tuning / 12.0);
    const double fb1 = freq_bin(1, a440);
    const double fb = (i == 0) ? fb1 - 1.5 * 12.0 : freq_bin(i, a440);
    const double fbn = freq_bin(i + 1, a440);  // i+1 <= 4097 < 8193 always exists
    double bw = fbn - fb;
    bw = (bw <= 1.) ? 1. : bw;
    double w[12];
    double ss = 0.;
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double d = -(double)r + fb;
        d = fmod(d + 6.0 + 10. * 12.0, 12.0) - 6.0;
        d = d / bw;
        w[r] = exp(-0.5 * (2. * d) * (2. * d));
        ss += w[r] * w[r];
    }
    ss = sqrt(ss);
    if (ss < 2.2250738585072014e-308) ss = 1.;
    double g = (fb / 12.0 - 5.0) / 2.0;
    g = exp(-0.5 * (g * g));
    double *o = table + ((size_t)idx * CH_BINS + i) * 12;
#pragma unroll
    for (int r = 0; r < 12; r++) o[r] = w[(r + 3) % 12] / ss * g;  // np.roll(-3): b[r] = wts[(r+3)%12]
}

__global__ void f64_to_f32_kernel(const double *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// builds the f64 table (chroma.rs:197-267 as written) and the f32 copy chroma_kernel multiplies with
int launch_chroma_filter_table(double *table, float *table32, cudaStream_t st) {
    dim3 grid((CH_BINS + 127) / 128, 100);
    BLISS_LAUNCH(chroma_filter_table_kernel, grid, 128, 0, st, table);
    const size_t n = (size_t)100 * CH_BINS * 12;
    BLISS_LAUNCH(f64_to_f32_kernel, (unsigned)((n + 255) / 256), 256, 0, st, table, table32, n);
    return 2;
}

// ---------------------------------------------------------------------------
// K3: one CTA (256 threads) per chroma frame: 8192-point real FFT as a 4096-point complex FFT
// (rfft8192.cuh), magnitudes to the spill buffer, pip_track peaks to the candidate list.
// ---------------------------------------------------------------------------
constexpr int K3_THREADS = 256;
constexpr int K3_FRAMES_PER_CTA = 4;  // consecutive frames of one song per CTA (amortises the song lookup)

// one pip_track candidate test in f32 (the f64 comparisons of chroma.rs:308 are order
// preserving on f32-exact values); `ref` stays f64 because 0.1*max is not an f32
__device__ __forceinline__ bool pip_is_peak(const float *sm, int c, double ref) {
    const float before = sm[c - 1], elem = sm[c], after = sm[c + 1];
    return after <= elem && before < elem && (double)elem > ref;
}

// magnitudes of bins t + 256 m, m = M..15:  twiddle W8192^(t + 256 m) = W8192^t * W32^m
template <int M>
__device__ __forceinline__ void epilogue_bins(const cpx *pk, const cpx *pm, bool t0, const cpx *buf0, cpx wt,
                                              float (&mag)[17]) {
    if constexpr (M < 16) {
        const cpx zk = pk[M];
        const cpx zm = t0 ? buf0[(16 - M) & 15] : pm[-M];
        mag[M] = r8k::untangle_mag(zk, zm, mul_tw<M, 32>(wt));
        epilogue_bins<M + 1>(pk, pm, t0, buf0, wt, mag);
    }
}

// pair epilogue: bins k = t + 256 M (M = 0..7) from the thread's own registers, their mirrors
// 4096 - k from the block the mirror thread published (rfft8192.cuh, pass3_regs).  Magnitudes go
// straight to the spill row; only bins < 1536 (lo[0..5]) are kept for pip_track.
template <int M>
__device__ __forceinline__ void epilogue_pairs(const cpx (&v)[16], const cpx *pm, bool t0, const cpx *buf0, cpx wt,
                                               float *gm_lo, float *gm_hi, float (&lo)[6], float &mx) {
    if constexpr (M < 8) {
        const cpx zm = t0 ? buf0[(16 - M) & 15] : pm[-M];
        float a, b;
        r8k::untangle_mag_pair(v[bitrev(M, 4)], zm, mul_tw<M, 32>(wt), a, b);
        gm_lo[256 * M] = a;
        gm_hi[-256 * M] = b;
        mx = fmaxf(mx, fmaxf(a, b));
        if constexpr (M < 6) lo[M] = a;
        epilogue_pairs<M + 1>(v, pm, t0, buf0, wt, gm_lo, gm_hi, lo, mx);
    }
}

// VAR (experimental cuts of pass 1, behind bliss_b200_set_variant; 0 = the measured kernel):
constexpr int K3V_TWPROD = 1;  // pass-1 and pass-2 twiddles: 4 loads + 11 products instead of 15 loads each (rfft8192.cuh)
constexpr int K3V_WINSYN = 2;  // Hann pairs from the thread's phase instead of 16 window loads
constexpr int K3V_LAY16 = 4;   // column-group pitch 16 instead of 17 in the FFT buffer: conflict-free mirror loads (rfft8192.cuh)
// Frames that start on an odd sample (hop 2205 is odd: every second one) cannot load (x[2n], x[2n+1]) as one aligned
// 64-bit pair.  K3V_ODDSHIFT transforms the frame rotated by one sample instead, y'[m] = y[(m - 1) mod 8192]: its
// pairs (y[2n-1], y[2n]) ARE aligned, |DFT(y')| = |DFT(y)| (a circular shift is a phase factor), and the only
// element that wraps is the first (y'[0] = y[8191], thread 0).  The window is read from a copy rotated the same way.
constexpr int K3V_ODDSHIFT = 8;
constexpr int K3_HANN_PHASE = CH_WIN;                 // float4[256] behind the window (api.cu build_tables)
constexpr int K3_HANN_SHIFT = CH_WIN + 1024;          // hann[(m + 8191) % 8192]
constexpr int K3_HANN_SHIFT_PHASE = 2 * CH_WIN + 1024;  // float4[256]: phase of samples 2 tid - 1, 2 tid

template <int Q>
__device__ __forceinline__ void window_synth(cpx (&v)[16], cpx cw, cpx sw) {
    if constexpr (Q < 16) {
        v[Q] = pmul(v[Q], r8k::hann_pair<Q>(cw, sw));
        window_synth<Q + 1>(v, cw, sw);
    }
}

template <bool PAIR_EPILOGUE, int VAR = 0>
__global__ void __launch_bounds__(K3_THREADS, 4)
stft8192_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                const unsigned int *__restrict__ frame_prefix, int n_songs,
                const float *__restrict__ hann, const cpx *__restrict__ tw1 /*[16][256] W4096^(b k1)*/,
                const cpx *__restrict__ tw2g /*[16][16] W256^(j k2)*/, const cpx *__restrict__ tw8192,
                float *__restrict__ mags,
                double *__restrict__ cand_mag, double *__restrict__ cand_pitch,
                unsigned int *__restrict__ cand_count) {
    __shared__ __align__(16) cpx buf[r8k::BUF_CPX];
    __shared__ cpx s_tw2[256];
    __shared__ float s_red[K3_THREADS / 32];
    __shared__ unsigned int s_scan[K3_THREADS / 32];
    __shared__ unsigned int s_base;

    constexpr int LB = (PAIR_EPILOGUE && (VAR & K3V_LAY16) != 0) ? 16 : 17;  // the old epilogue gathers by bin: keeps pad()
    const int tid = threadIdx.x;
    s_tw2[tid] = tw2g[tid];  // visible after the barrier that follows pass 1
    const unsigned int item = blockIdx.x;
    const int si = find_song(frame_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const int fbase = (int)(item - frame_prefix[si]) * K3_FRAMES_PER_CTA;
    const float *x = pcm + sd.pcm_off;
    const int n = (int)sd.n;
    const unsigned int mag_row0 = (unsigned int)sd.mag_off;  // spill rows of a wave stay far below 2^32
    const int fend = min(fbase + K3_FRAMES_PER_CTA, (int)sd.n_c_comp);
#pragma unroll 1
    for (int f = fbase; f < fend; f++) {
    // the frame covers samples s0 .. s0+8191 of the reflect-padded song (utils.rs:11-24, :44-47)
    const int s0 = CH_HOP * f - 4096;
    const bool interior = (s0 >= 0) && (s0 + 8191 < n);
    // The next frame shares 5987 of its 8192 samples with this one; the 2205 new ones (70 lines of 128 B)
    // are requested into L2 now, a whole frame time before its loads need them.
    if (tid < 70 && f + 1 < fend) {
        const int nx = s0 + 8192 + 32 * tid;
#ifdef __CUDA_ARCH__
        if (nx < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + nx));
#endif
    }

    // pass 1 straight from global memory: z[nn] = w[2nn] x[2nn] + i w[2nn+1] x[2nn+1], nn = tid + 256 q
    {
        cpx v[16];
        const float2 *ph = reinterpret_cast<const float2 *>(hann) + tid;
        if constexpr ((VAR & K3V_WINSYN) != 0) {
            // samples only; the window comes from the thread's phase entry behind the 8192 window values
            float4 pw = __ldg(reinterpret_cast<const float4 *>(hann + K3_HANN_PHASE) + tid);
            if (interior) {
                const float *pa = x + s0 + 2 * tid;
                if ((reinterpret_cast<size_t>(pa) & 7) == 0) {
                    const float2 *pa2 = reinterpret_cast<const float2 *>(pa);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float2 xx = __ldg(pa2 + 256 * q);
                        v[q] = cpx{xx.x, xx.y};
                    }
                } else if constexpr ((VAR & K3V_ODDSHIFT) != 0) {
                    const float2 *pa2 = reinterpret_cast<const float2 *>(pa - 1);  // pa is 4 mod 8: pa - 1 is aligned
                    pw = __ldg(reinterpret_cast<const float4 *>(hann + K3_HANN_SHIFT_PHASE) + tid);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float2 xx = __ldg(pa2 + 256 * q);
                        v[q] = cpx{xx.x, xx.y};
                    }
                    if (tid == 0) v[0].x = __ldg(x + s0 + 8191);  // y'[0] = y[8191]
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++) v[q] = cpx{__ldg(pa + 512 * q), __ldg(pa + 512 * q + 1)};
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const long long i0 = (long long)s0 + 2 * (tid + 256 * q);
                    v[q] = cpx{r8k::reflect_sample(x, n, i0), r8k::reflect_sample(x, n, i0 + 1)};
                }
            }
            window_synth<0>(v, cpx{pw.x, pw.y}, cpx{pw.z, pw.w});
        } else if (interior) {
            const float *pa = x + s0 + 2 * tid;
            if ((reinterpret_cast<size_t>(pa) & 7) == 0) {  // even frames of an 8-byte aligned song: one 64-bit load per pair
                const float2 *pa2 = reinterpret_cast<const float2 *>(pa);  // (half the L1 wavefronts of two 32-bit loads)
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float2 w = __ldg(ph + 256 * q);
                    const float2 xx = __ldg(pa2 + 256 * q);
                    v[q] = pmul(cpx{xx.x, xx.y}, cpx{w.x, w.y});
                }
            } else if constexpr ((VAR & K3V_ODDSHIFT) != 0) {
                const float2 *pa2 = reinterpret_cast<const float2 *>(pa - 1);  // pa is 4 mod 8: pa - 1 is aligned
                const float2 *phs = reinterpret_cast<const float2 *>(hann + K3_HANN_SHIFT) + tid;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float2 w = __ldg(phs + 256 * q);
                    float2 xx = __ldg(pa2 + 256 * q);
                    if (q == 0 && tid == 0) xx.x = __ldg(x + s0 + 8191);  // y'[0] = y[8191]
                    v[q] = pmul(cpx{xx.x, xx.y}, cpx{w.x, w.y});
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float2 w = __ldg(ph + 256 * q);
                    v[q] = pmul(cpx{__ldg(pa + 512 * q), __ldg(pa + 512 * q + 1)}, cpx{w.x, w.y});
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float2 w = __ldg(ph + 256 * q);
                const long long i0 = (long long)s0 + 2 * (tid + 256 * q);
                v[q] = pmul(cpx{r8k::reflect_sample(x, n, i0), r8k::reflect_sample(x, n, i0 + 1)}, cpx{w.x, w.y});
            }
        }
        if constexpr ((VAR & K3V_TWPROD) != 0) r8k::pass1_store_prod<LB>(tid, v, tw1, buf);
        else r8k::pass1_store<LB>(tid, v, tw1, buf);
    }
    __syncthreads();
    if constexpr ((VAR & K3V_TWPROD) != 0) r8k::pass2_prod<LB>(tid, s_tw2, buf);
    else r8k::pass2<LB>(tid, s_tw2, buf);
    __syncthreads();
    float *sm = reinterpret_cast<float *>(buf);  // magnitudes for pip_track once the FFT data is dead
    float fmx;                                   // frame maximum
    if constexpr (!PAIR_EPILOGUE) {
        r8k::pass3(tid, buf);
        __syncthreads();

        // natural-order magnitudes: thread owns bins tid + 256 m, m = 0..15; thread 0 also bin 4096
        float mag[17];
        mag[16] = 0.f;
        {
            const cpx wt = tw8192[tid];
            const cpx *pk = buf + r8k::zbase(tid);
            const cpx *pm = buf + r8k::zbase((256 - tid) & 255) + 15;
            epilogue_bins<0>(pk, pm, tid == 0, buf, wt, mag);
            if (tid == 0) mag[16] = r8k::untangle_mag(buf[0], buf[0], cpx{-1.f, 0.f});
        }
        float mx = 0.f;
#pragma unroll
        for (int m = 0; m < 17; m++) mx = fmaxf(mx, mag[m]);
        __syncthreads();  // everyone has read buf; reuse it for the magnitudes
        float *gm = mags + (size_t)(mag_row0 + (unsigned int)f) * CH_STRIDE;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            sm[tid + 256 * m] = mag[m];
            gm[tid + 256 * m] = mag[m];
        }
        if (tid == 0) {
            sm[4096] = mag[16];
            gm[4096] = mag[16];
        }
        // frame maximum (pip_track's ref_value = 0.1 * max over all bins, chroma.rs:289-293)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = mx;
        __syncthreads();
        fmx = s_red[0];
#pragma unroll
        for (int w = 1; w < K3_THREADS / 32; w++) fmx = fmaxf(fmx, s_red[w]);

    } else {
        // pass 3 leaves the thread's own bins t + 256 m in registers; only mirror values travel through smem
        float lo[6], mx = 0.f;  // lo[M] = |X[t + 256 M]|: the bins pip_track looks at (56..1484)
        float *gm = mags + (size_t)(mag_row0 + (unsigned int)f) * CH_STRIDE;
        {
            cpx v[16];
            r8k::pass3_regs<LB>(tid, v, buf);
            __syncthreads();
            const cpx wt = tw8192[tid];
            const cpx *pm = buf + r8k::zbase<LB>((256 - tid) & 255) + 15;
            epilogue_pairs<0>(v, pm, tid == 0, buf, wt, gm + tid, gm + 4096 - tid, lo, mx);
            if (tid == 0) {  // the self-mirrored bin 2048: W8192^2048 = -i
                const float mid = r8k::untangle_mag(v[bitrev(8, 4)], v[bitrev(8, 4)], cpx{0.f, -1.f});
                gm[2048] = mid;
                mx = fmaxf(mx, mid);
            }
        }
        __syncthreads();  // every mirror value has been read; reuse buf for the magnitudes pip_track looks at
#pragma unroll
        for (int m = 0; m < 6; m++) sm[tid + 256 * m] = lo[m];
        // frame maximum (pip_track's ref_value = 0.1 * max over all bins, chroma.rs:289-293)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = mx;
        __syncthreads();
        fmx = s_red[0];
#pragma unroll
        for (int w = 1; w < K3_THREADS / 32; w++) fmx = fmaxf(fmx, s_red[w]);
    }

    // pip_track on centre bins 57..1483 (beginning = 56, end = 1486 for n_fft = 8192).
    // Phase 1 counts this thread's peaks, a block scan reserves the output range, phase 2 emits.
    const double ref = 0.1 * (double)fmx;
    unsigned int flags = 0;
#pragma unroll
    for (int m = 0; m < 6; m++) {
        const int c = 57 + tid + K3_THREADS * m;
        if (c <= 1483 && pip_is_peak(sm, c, ref)) flags |= 1u << m;
    }
    const int cnt = __popc(flags);
    unsigned int incl = (unsigned)cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    unsigned int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < K3_THREADS / 32; w++) {
        if (w < (tid >> 5)) woff += s_scan[w];
        tot += s_scan[w];
    }
    if (tid == 0 && tot) s_base = atomicAdd(cand_count + si, tot);
    __syncthreads();
    unsigned long long dst = sd.cand_off + s_base + woff + (incl - (unsigned)cnt);
    while (flags) {
        const int m = __ffs(flags) - 1;
        flags &= flags - 1;
        const int c = 57 + tid + K3_THREADS * m;
        const double before = (double)sm[c - 1], elem = (double)sm[c], after = (double)sm[c + 1];
        const double avg = 0.5 * (after - before);
        double shift = 2. * elem - after - before;
        if (fabs(shift) < 2.2250738585072014e-308) shift += 1.;
        shift = avg / shift;
        // the residue bin of pitch_tuning (log2 / fmod in f64) is left to tuning_kernel, which is
        // latency-bound and overlapped; here only the interpolation of chroma.rs:317-326
        cand_pitch[dst] = ((double)c + shift) * (double)SAMPLE_RATE / 8192.0;
        cand_mag[dst] = elem + 0.5 * avg * shift;
        dst++;
    }
    __syncthreads();  // the magnitudes in `buf` were read by pip_track; the next frame overwrites them
    }  // frames of this CTA
}

// ---------------------------------------------------------------------------
// K4: one CTA per song.  threshold = Midpoint median of the candidate magnitudes
// (exact radix select on the f64 bit patterns: all values are positive), then the
// 100-bin histogram of the residues of candidates with mag >= threshold, argmax =
// first maximum.  Writes the tuning INDEX (tuning = (-50 + idx)/100).
// ---------------------------------------------------------------------------
constexpr int K4_THREADS = 1024;

#ifndef BLISS_HOST_EMUL  // the previous tuning kernel (partial-mask __match_any_sync) is not part of the host emulation
__device__ __forceinline__ void hist_add(unsigned int *hist, unsigned int bucket, bool active) {
    // warp-aggregated shared-memory histogram update
    const unsigned int act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned int peers = __match_any_sync(act, bucket);
    const int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(hist + bucket, (unsigned)__popc(peers));
}

__global__ void __launch_bounds__(K4_THREADS)
tuning_kernel(const double *__restrict__ cand_mag, const double *__restrict__ cand_pitch,
              const unsigned int *__restrict__ cand_count, const SongDesc *__restrict__ songs,
              int *__restrict__ tuning_idx) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix, s_vhi;
    __shared__ unsigned int s_rank, s_cle;
    const SongDesc sd = songs[blockIdx.x];
    const int tid = threadIdx.x;
    const unsigned int n = sd.valid ? cand_count[blockIdx.x] : 0u;
    if (n == 0) {  // estimate_tuning returns 0 when pip_track finds nothing (chroma.rs:377-379)
        if (tid == 0) tuning_idx[blockIdx.x] = 50;
        return;
    }
    const unsigned long long *keys = reinterpret_cast<const unsigned long long *>(cand_mag + sd.cand_off);
    const double *pitches = cand_pitch + sd.cand_off;
    const unsigned int r_lo = (n - 1) / 2, r_hi = n / 2;  // floor / ceil of (n-1)*0.5

    if (tid == 0) { s_prefix = 0ull; s_rank = r_lo; }
    unsigned long long mask = 0ull;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned int n_round = (n + K4_THREADS - 1) / K4_THREADS * K4_THREADS;
        for (unsigned int i = tid; i < n_round; i += K4_THREADS) {
            bool act = false;
            unsigned int bucket = 0;
            if (i < n) {
                const unsigned long long k = keys[i];
                act = (k & mask) == prefix;
                bucket = (unsigned int)((k >> shift) & 255ull);
            }
            hist_add(hist, bucket, act);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int rank = s_rank, cum = 0;
            int b = 0;
            for (; b < 256; b++) {
                if (cum + hist[b] > rank) break;
                cum += hist[b];
            }
            s_rank = rank - cum;
            s_prefix = prefix | ((unsigned long long)b << shift);
        }
        mask |= 255ull << shift;
        __syncthreads();
    }
    const unsigned long long vlo = s_prefix;
    // higher order statistic: equal to vlo when enough duplicates, else the next larger key
    if (tid == 0) { s_cle = 0; s_vhi = ~0ull; }
    __syncthreads();
    if (r_hi != r_lo) {
        unsigned int cle = 0;
        unsigned long long mn = ~0ull;
        for (unsigned int i = tid; i < n; i += K4_THREADS) {
            const unsigned long long k = keys[i];
            cle += (k <= vlo) ? 1u : 0u;
            if (k > vlo && k < mn) mn = k;
        }
        cle = __reduce_add_sync(0xffffffffu, cle);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, mn, o);
            mn = t < mn ? t : mn;
        }
        if ((tid & 31) == 0) {
            atomicAdd(&s_cle, cle);
            atomicMin(&s_vhi, mn);
        }
    }
    __syncthreads();
    unsigned long long vhi = vlo;
    if (r_hi != r_lo && s_cle < r_hi + 1) vhi = s_vhi;
    const double lower = __longlong_as_double((long long)vlo), higher = __longlong_as_double((long long)vhi);
    const double thr = lower + (higher - lower) / 2.;  // ndarray-stats Midpoint

    // histogram of residues with mag >= thr (chroma.rs:385-390 -> pitch_tuning :348-356)
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    {
        const unsigned int n_round = (n + K4_THREADS - 1) / K4_THREADS * K4_THREADS;
        const double *mg = cand_mag + sd.cand_off;
        for (unsigned int i = tid; i < n_round; i += K4_THREADS) {
            bool act = false;
            unsigned int bucket = 0;
            if (i < n) {
                act = mg[i] >= thr;
                if (act) {
                    // pitch_tuning (chroma.rs:342-348): hz_to_octs with tuning 0, 12 bins/octave, residue in
                    // [-0.5, 0.5), 100 bins of 0.01
                    double v = pitches[i] / (440.0 / 16.);
                    v = log2(v);
                    v = fmod(12.0 * v, 1.0);
                    if (v >= 0.5) v -= 1.;
                    int idx = (int)((v - -0.5) / 0.01);
                    bucket = (unsigned int)(idx < 0 ? 0 : (idx > 99 ? 99 : idx));
                }
            }
            hist_add(hist, bucket, act);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int b = 1; b < 100; b++)
            if (hist[b] > hist[best]) best = b;
        tuning_idx[blockIdx.x] = best;
    }
}

#endif  // BLISS_HOST_EMUL
// ---------------------------------------------------------------------------
// K4 (current): same result as tuning_kernel above, four sweeps over the candidates instead of ten.
//   sweep 0  min / max key  -> the bits every key shares are skipped
//   sweep 1+ 12-bit digit histogram below the shared prefix, block scan picks the bucket holding the
//            wanted rank; repeated only while the bucket holds more keys than fit in shared memory
//            (heavy duplicates / the 714-peaks-per-frame worst case)
//   sweep 2  the bucket's keys go to shared memory (+ the smallest key above the bucket)
//            -> both order statistics by rank counting on <= 2048 keys
//   sweep 3  residue histogram of the candidates with mag >= threshold, per-warp private bins
// No warp collectives inside the sweeps, so the loads of consecutive iterations overlap.
// ---------------------------------------------------------------------------
constexpr int K4_BINS = 4096;      // 12-bit digits
constexpr int K4_SURV_CAP = 2048;  // keys of the final bucket kept in shared memory

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    return (unsigned long long)__shfl_xor_sync(0xffffffffu, (long long)v, m);
}

__global__ void __launch_bounds__(K4_THREADS)
tuning_select_kernel(const double *__restrict__ cand_mag, const double *__restrict__ cand_pitch,
                     const unsigned int *__restrict__ cand_count, const SongDesc *__restrict__ songs,
                     int *__restrict__ tuning_idx) {
    __shared__ unsigned int hist[K4_BINS];
    __shared__ unsigned long long surv[K4_SURV_CAP];
    __shared__ unsigned long long s_min[K4_THREADS / 32], s_max[K4_THREADS / 32];
    __shared__ unsigned int s_wsum[K4_THREADS / 32];
    __shared__ unsigned int s_bucket, s_rank, s_count, s_nsurv;
    __shared__ unsigned long long s_vlo, s_vhi, s_next;
    const SongDesc sd = songs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned int n = sd.valid ? cand_count[blockIdx.x] : 0u;
    if (n == 0) {  // estimate_tuning returns 0 when pip_track finds nothing (chroma.rs:377-379)
        if (tid == 0) tuning_idx[blockIdx.x] = 50;
        return;
    }
    const unsigned long long *keys = reinterpret_cast<const unsigned long long *>(cand_mag + sd.cand_off);
    const double *pitches = cand_pitch + sd.cand_off;
    const unsigned int r_lo = (n - 1) / 2, r_hi = n / 2;  // floor / ceil of (n-1)*0.5 (Midpoint quantile)

    // ---- sweep 0: min / max (all keys are positive doubles: bit order == value order) ----
    unsigned long long kmin = ~0ull, kmax = 0ull;
#pragma unroll 4
    for (unsigned int i = tid; i < n; i += K4_THREADS) {
        const unsigned long long k = __ldg(keys + i);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = shfl_xor_u64(kmin, o), b = shfl_xor_u64(kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    if (lane == 0) { s_min[wid] = kmin; s_max[wid] = kmax; }
    if (tid == 0) { s_nsurv = 0; s_next = ~0ull; s_vlo = 0ull; s_vhi = 0ull; }
    __syncthreads();
#pragma unroll 1
    for (int w = 0; w < K4_THREADS / 32; w++) {
        kmin = s_min[w] < kmin ? s_min[w] : kmin;
        kmax = s_max[w] > kmax ? s_max[w] : kmax;
    }

    unsigned long long vlo, vhi;
    if (kmin == kmax) {
        vlo = vhi = kmin;
    } else {
        // bits [63 : d_prev) are shared by every key still in play and equal `prefix`
        int d_prev = 64 - __clzll((long long)(kmin ^ kmax));  // 1..63 (the sign bit is always shared)
        unsigned long long prefix = kmin >> d_prev;
        unsigned int rank = r_lo, count = n;
        int lo_bit;
        while (true) {
            lo_bit = d_prev > 12 ? d_prev - 12 : 0;
            const int width = d_prev - lo_bit;
            const unsigned int dmask = (1u << width) - 1u;
#pragma unroll
            for (int j = 0; j < K4_BINS / K4_THREADS; j++) hist[tid + K4_THREADS * j] = 0u;
            __syncthreads();
#pragma unroll 4
            for (unsigned int i = tid; i < n; i += K4_THREADS) {
                const unsigned long long k = __ldg(keys + i);
                if ((k >> d_prev) == prefix) atomicAdd(&hist[(unsigned int)(k >> lo_bit) & dmask], 1u);
            }
            __syncthreads();
            // bucket holding `rank`: thread owns bins 4 tid .. 4 tid + 3
            unsigned int c[K4_BINS / K4_THREADS], mine = 0;
#pragma unroll
            for (int j = 0; j < K4_BINS / K4_THREADS; j++) { c[j] = hist[(K4_BINS / K4_THREADS) * tid + j]; mine += c[j]; }
            unsigned int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[wid] = incl;
            __syncthreads();
            unsigned int before = incl - mine;
#pragma unroll 1
            for (int w = 0; w < wid; w++) before += s_wsum[w];
            if (before <= rank && rank < before + mine) {  // exactly one thread
                unsigned int e = before;
#pragma unroll
                for (int j = 0; j < K4_BINS / K4_THREADS; j++) {
                    if (rank >= e && rank < e + c[j]) {
                        s_bucket = (unsigned int)((K4_BINS / K4_THREADS) * tid + j);
                        s_rank = rank - e;
                        s_count = c[j];
                    }
                    e += c[j];
                }
            }
            __syncthreads();
            prefix = (prefix << width) | (unsigned long long)s_bucket;
            rank = s_rank;
            count = s_count;
            d_prev = lo_bit;
            if (count <= (unsigned int)K4_SURV_CAP || lo_bit == 0) break;
        }
        // ---- sweep 2: the bucket's keys to shared memory; the smallest key above the bucket ----
        // (lo_bit == 0 with count > cap: every key of the bucket equals `prefix`, nothing to collect)
        const bool collect = count <= (unsigned int)K4_SURV_CAP;
        unsigned long long nxt = ~0ull;
#pragma unroll 4
        for (unsigned int i = tid; i < n; i += K4_THREADS) {
            const unsigned long long k = __ldg(keys + i);
            const unsigned long long top = k >> lo_bit;
            if (top == prefix) {
                if (collect) surv[atomicAdd(&s_nsurv, 1u)] = k;
            } else if (top > prefix) {
                nxt = k < nxt ? k : nxt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = shfl_xor_u64(nxt, o);
            nxt = a < nxt ? a : nxt;
        }
        if (lane == 0 && nxt != ~0ull) atomicMin(&s_next, nxt);
        __syncthreads();
        if (collect) {
            // order statistics `rank` and `rank + 1` of the bucket by rank counting (ties by position)
            for (unsigned int idx = tid; idx < count; idx += K4_THREADS) {
                const unsigned long long key = surv[idx];
                unsigned int r = 0;
                for (unsigned int j = 0; j < count; j++) {
                    const unsigned long long o = surv[j];
                    r += (o < key || (o == key && j < idx)) ? 1u : 0u;
                }
                if (r == rank) s_vlo = key;
                if (r == rank + 1u) s_vhi = key;
            }
            __syncthreads();
            vlo = s_vlo;
            vhi = (rank + 1u < count) ? s_vhi : s_next;
        } else {
            vlo = prefix;
            vhi = (rank + 1u < count) ? prefix : s_next;
        }
        if (r_hi == r_lo) vhi = vlo;
    }
    const double lower = __longlong_as_double((long long)vlo), higher = __longlong_as_double((long long)vhi);
    const double thr = lower + (higher - lower) / 2.;  // ndarray-stats Midpoint

    // ---- sweep 3: histogram of residues with mag >= thr (chroma.rs:385-390 -> pitch_tuning :342-356) ----
    __syncthreads();  // hist / surv are free again
#pragma unroll
    for (int j = 0; j < K4_BINS / K4_THREADS; j++) hist[tid + K4_THREADS * j] = 0u;
    __syncthreads();
    {
        unsigned int *my = hist + 128 * wid;  // per-warp private bins
        const double *mg = cand_mag + sd.cand_off;
#pragma unroll 2
        for (unsigned int i = tid; i < n; i += K4_THREADS) {
            if (__ldg(mg + i) >= thr) {
                // hz_to_octs with tuning 0, 12 bins per octave, residue in [-0.5, 0.5), 100 bins of 0.01
                double v = __ldg(pitches + i) / (440.0 / 16.);
                v = 12.0 * log2(v);
                v = v - trunc(v);  // == fmod(v, 1.0): the fractional part is exact either way
                if (v >= 0.5) v -= 1.;
                const int idx = (int)((v - -0.5) / 0.01);
                atomicAdd(my + (idx < 0 ? 0 : (idx > 99 ? 99 : idx)), 1u);
            }
        }
    }
    __syncthreads();
    if (tid < 128) {
        unsigned int v = 0;
#pragma unroll 1
        for (int w = 0; w < K4_THREADS / 32; w++) v += hist[128 * w + tid];
        hist[tid] = v;  // column `tid` of warp 0's bins: read by this thread only
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int b = 1; b < 100; b++)
            if (hist[b] > hist[best]) best = b;  // argmax = first maximum
        tuning_idx[blockIdx.x] = best;
    }
}

// ---------------------------------------------------------------------------
// K5: chroma contraction + interval features, thread per frame, 128 frames per CTA.
// ---------------------------------------------------------------------------
constexpr int K5_THREADS = 128;            // each thread owns frames tid and tid + 128 of the tile
constexpr int K5_KT = 16;                  // bins staged per step

// Interval / triad templates of chroma.rs:139-152 given as the offsets of their ones:
// dyads {0,d} d=1..6, major {0,4,7}, minor {0,3,7}, diminished {0,3,6}, augmented {0,4,8}.
__host__ __device__ constexpr int tmpl_off(int t, int i) {
    return i == 0 ? 0
         : t < 6  ? (i == 1 ? t + 1 : -1)
         : t == 6 ? (i == 1 ? 4 : 7)
         : t == 7 ? (i == 1 ? 3 : 7)
         : t == 8 ? (i == 1 ? 3 : 6)
                  : (i == 1 ? 4 : 8);
}
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// product over the template rolled right by S (rotate_right, chroma.rs:164-165),
// factors taken in ascending pitch-class order like `x.product()` over the row
template <int T, int S>
__device__ __forceinline__ double tmpl_term(const double (&e)[12]) {
    constexpr int p0 = (tmpl_off(T, 0) + S) % 12, p1 = (tmpl_off(T, 1) + S) % 12;
    if constexpr (tmpl_off(T, 2) < 0) {
        constexpr int a = cmin(p0, p1), b = cmax(p0, p1);
        return (1. * e[a]) * e[b];
    } else {
        constexpr int p2 = (tmpl_off(T, 2) + S) % 12;
        constexpr int a = cmin(p0, cmin(p1, p2)), c = cmax(p0, cmax(p1, p2)), b = p0 + p1 + p2 - a - c;
        return ((1. * e[a]) * e[b]) * e[c];
    }
}
template <int T, int S>
__device__ __forceinline__ void tmpl_sum(const double (&e)[12], double &f) {
    if constexpr (S < 12) {
        f += tmpl_term<T, S>(e);
        tmpl_sum<T, S + 1>(e, f);
    }
}
template <int T>
__device__ __forceinline__ void tmpl_all(const double (&e)[12], double (&feat)[10]) {
    if constexpr (T < 10) {
        double f = 0.;
        tmpl_sum<T, 0>(e, f);
        feat[T] = f;
        tmpl_all<T + 1>(e, feat);
    }
}

// L1 normalisation, exp(15 x), L1 normalisation again, 120 template products (chroma.rs:137-188, :404-410)
__device__ __forceinline__ void frame_interval_features(const double (&acc)[12], double (&feat)[10], double *dbg) {
    double sum = 0.;
#pragma unroll
    for (int c = 0; c < 12; c++) sum += fabs(acc[c]);
    if (sum < 2.2250738585072014e-308) sum = 1.;
    double e[12];
    double esum = 0.;
#pragma unroll
    for (int c = 0; c < 12; c++) {
        const double ch = acc[c] / sum;
        if (dbg) dbg[c] = ch;
        e[c] = exp(ch * 15.);  // chroma_interval_features, chroma.rs:138
        esum += fabs(e[c]);
    }
    if (esum < 0.0001) esum = 1.;  // normalize_feature_sequence, chroma.rs:177-188
#pragma unroll
    for (int c = 0; c < 12; c++) e[c] = e[c] / esum;
    double f[10];
    tmpl_all<0>(e, f);  // extract_interval_features, chroma.rs:157-175
#pragma unroll
    for (int t = 0; t < 10; t++) feat[t] += f[t];
}

__global__ void __launch_bounds__(K5_THREADS)
chroma_kernel(const float *__restrict__ mags, const SongDesc *__restrict__ songs,
              const unsigned int *__restrict__ tile_prefix, int n_songs,
              const float *__restrict__ filt_table, const int *__restrict__ tuning_idx,
              double *__restrict__ tile_partials /*[tiles][10]*/, double *__restrict__ chroma_dbg) {
    __shared__ float s_s[K5_KT][CH_TILE_FRAMES + 1];
    __shared__ __align__(16) float s_w[K5_KT][12];
    __shared__ double s_red[K5_THREADS / 32][10];

    const int tid = threadIdx.x;
    const unsigned int item = blockIdx.x;
    const int si = find_song(tile_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const int tile = (int)(item - tile_prefix[si]);
    const int f0 = tile * CH_TILE_FRAMES;
    const int nf = min(CH_TILE_FRAMES, (int)sd.n_c - f0);  // frames of this tile
    const float *W = filt_table + (size_t)tuning_idx[si] * CH_BINS * 12;
    const float *S = mags + (sd.mag_off + (unsigned long long)f0) * CH_STRIDE;

    // The reference contracts in f64 (chroma.rs:403, ndarray dot on f64).  Here every 16-bin chunk
    // is accumulated with f32 FMAs (products of two f32-exact factors, <= 16 additions) and the chunk
    // sums are added in f64: the result differs from the all-f64 sum by ~1e-7 relative, the level of
    // the f32 FFT's own rounding, at a third of the FP64-pipe time.
    // Thread-owned frames: tid and tid + 128; a bin's 12 weights are read once for both frames.
    double acc0[12], acc1[12];
#pragma unroll
    for (int c = 0; c < 12; c++) { acc0[c] = 0.; acc1[c] = 0.; }

    // Software pipeline: the next k-tile (16 bins x 256 frames of magnitudes + 16 x 12 weights) is
    // fetched into registers while the current one is consumed from shared memory.
    // Thread t stages bin (t & 15) of frames (t >> 4) + 8 i, i = 0..31  -> 64 B per half-warp per row.
    const int st_k = tid & 15, st_f = tid >> 4;
    float pre_s[32];
    float pre_w[2];
    auto fetch = [&](int k0) {
        const int kt = min(K5_KT, CH_BINS - k0);
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int fr = st_f + 8 * i;
            float v = 0.f;
            // rows beyond n_c_comp are zero (utils.rs:27-31)
            if (fr < nf && st_k < kt && (f0 + fr) < (int)sd.n_c_comp) v = __ldg(S + (size_t)fr * CH_STRIDE + k0 + st_k);
            pre_s[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int e = tid + K5_THREADS * i;  // 192 = 16 x 12 weights
            pre_w[i] = (e < kt * 12) ? __ldg(W + (size_t)k0 * 12 + e) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < CH_BINS; k0 += K5_KT) {
        __syncthreads();  // previous tile fully consumed
#pragma unroll
        for (int i = 0; i < 32; i++) s_s[st_k][st_f + 8 * i] = pre_s[i];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int e = tid + K5_THREADS * i;
            if (e < K5_KT * 12) (&s_w[0][0])[e] = pre_w[i];
        }
        __syncthreads();
        if (k0 + K5_KT < CH_BINS) fetch(k0 + K5_KT);
        float p0[12], p1[12];
#pragma unroll
        for (int c = 0; c < 12; c++) { p0[c] = 0.f; p1[c] = 0.f; }
#pragma unroll 4
        for (int kk = 0; kk < K5_KT; kk++) {  // bins past 4096 were staged as zeros
            const float a = s_s[kk][tid], b = s_s[kk][tid + 128];
            const float a2 = a * a, b2 = b * b;  // spectrum.mapv_inplace(|x| x*x), chroma.rs:400
#pragma unroll
            for (int c = 0; c < 12; c++) {
                const float w = s_w[kk][c];
                p0[c] = fmaf(w, a2, p0[c]);
                p1[c] = fmaf(w, b2, p1[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 12; c++) {
            acc0[c] += (double)p0[c];
            acc1[c] += (double)p1[c];
        }
    }
    double feat[10];
#pragma unroll
    for (int t = 0; t < 10; t++) feat[t] = 0.;
    if (tid < nf)
        frame_interval_features(acc0, feat, chroma_dbg ? chroma_dbg + ((size_t)sd.c_tile_off * CH_TILE_FRAMES + f0 + tid) * 12 : nullptr);
    if (tid + 128 < nf)
        frame_interval_features(acc1, feat, chroma_dbg ? chroma_dbg + ((size_t)sd.c_tile_off * CH_TILE_FRAMES + f0 + tid + 128) * 12 : nullptr);
    // sum the tile's frames (mean_axis over frames finishes in the summary kernel)
#pragma unroll
    for (int t = 0; t < 10; t++) {
        double v = feat[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[tid >> 5][t] = v;
    }
    __syncthreads();
    if (tid < 10) {
        double v = 0.;
#pragma unroll
        for (int w = 0; w < K5_THREADS / 32; w++) v += s_red[w][tid];
        tile_partials[((size_t)sd.c_tile_off + tile) * 10 + tid] = v;
    }
}

// ---------------------------------------------------------------------------
// K5 (current): the same contraction and epilogue as chroma_kernel above -- bit-identical results, the
// 16-bin f32 chunks and their f64 sums are kept -- but the magnitude tile never passes through
// registers: cp.async (LDGSTS) copies 64-byte row segments [frame][16 bins] straight into a 3-stage
// shared-memory ring, one __syncthreads per stage, and the consumer reads its two frames and the
// broadcast weights with 128-bit loads (row pitch 80 B = 5 x 16 B: conflict-free for quarter warps).
// Per 16-bin stage and thread: 8 cp.async + 8 LDS.128 (values) + 48 LDS.128 (weights, broadcast)
// for 384 FMA, against 32 LDG + 32 STS + 32 LDS + 192 LDS before.
// ---------------------------------------------------------------------------
#ifndef K5P_STAGES_N
#define K5P_STAGES_N 3
#endif
#ifndef K5P_SWIZZLE
#define K5P_SWIZZLE 1
#endif
constexpr int K5P_STAGES = K5P_STAGES_N;
// Staged row: exactly 64 B, its four 16-byte chunks permuted by (row >> 1) & 3 -- conflict-free for the quarter-warp
// 128-bit loads AND for the LDGSTS writes (K5P_SWIZZLE 0: the first layout, 16 floats + 4 of padding, pitch 80 B,
// whose writes were not).  A stage is 16 instead of 20 KB, and under a 128-register cap FOUR CTAs share an SM.
// Measured on 1024 tracks (scripts/gpu_r02_v.sh, gpu_r02_w.sh): padded rows, 3 CTAs 6.85 ms; swizzled, 3 CTAs 5.93;
// swizzled, 4 CTAs 5.81 (the step: 52.05 -> 51.15 -> 50.61 ms).  A fourth or only two stages change nothing any more
// (5.91 / 5.79 ms): the loads are no longer what the kernel waits for.  Same arithmetic, same bits.
constexpr int K5P_PITCH = K5P_SWIZZLE ? K5_KT : K5_KT + 4;
__host__ __device__ constexpr int k5p_swz(int row) { return K5P_SWIZZLE ? ((row >> 1) & 3) : 0; }

#ifdef BLISS_HOST_EMUL  // host emulation: the copy happens on the spot (a legal completion time for cp.async)
inline void cp_async16(void *smem, const void *gmem, int src_bytes) {
    memset(smem, 0, 16);
    if (src_bytes > 0) memcpy(smem, gmem, (size_t)src_bytes);
}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem);
    // src_bytes < 16: the remainder of the 16 bytes is zero-filled (0: nothing is read)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

#ifndef K5P_MIN_BLOCKS
#define K5P_MIN_BLOCKS (K5P_SWIZZLE ? (K5P_STAGES_N <= 3 ? 4 : 3) : (K5P_STAGES_N == 3 ? 3 : 2))
#endif
__global__ void __launch_bounds__(K5_THREADS, K5P_MIN_BLOCKS)
chroma_pipe_kernel(const float *__restrict__ mags, const SongDesc *__restrict__ songs,
                   const unsigned int *__restrict__ tile_prefix, int n_songs,
                   const float *__restrict__ filt_table, const int *__restrict__ tuning_idx,
                   double *__restrict__ tile_partials /*[tiles][10]*/, double *__restrict__ chroma_dbg) {
#ifdef BLISS_HOST_EMUL
    unsigned char *k5p_smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char k5p_smem[];
#endif
    float *s_s = reinterpret_cast<float *>(k5p_smem);                                   // [STAGES][256][PITCH]
    float *s_w = s_s + K5P_STAGES * CH_TILE_FRAMES * K5P_PITCH;                         // [STAGES][16][12]
    double *s_red = reinterpret_cast<double *>(s_w + K5P_STAGES * K5_KT * 12);          // [4][10]

    const int tid = threadIdx.x;
    const unsigned int item = blockIdx.x;
    const int si = find_song(tile_prefix, n_songs, item);
    const SongDesc sd = songs[si];
    const int tile = (int)(item - tile_prefix[si]);
    const int f0 = tile * CH_TILE_FRAMES;
    const int nf = min(CH_TILE_FRAMES, (int)sd.n_c - f0);  // frames of this tile
    // rows that exist in the spill: frames below n_c_comp (the rest are zero rows, utils.rs:27-31)
    const int nrows = max(0, min(nf, (int)sd.n_c_comp - f0));
    const float *W = filt_table + (size_t)tuning_idx[si] * CH_BINS * 12;
    const float *S = mags + (size_t)((unsigned int)sd.mag_off + (unsigned int)f0) * CH_STRIDE;

    // staging map: thread copies 16-byte chunk (tid & 3) of rows (tid >> 2) + 32 i, i = 0..7
    const int st_c = tid & 3, st_r = tid >> 2;
    auto issue = [&](int stage, int k0) {
        float *dst = s_s + (size_t)stage * CH_TILE_FRAMES * K5P_PITCH + st_r * K5P_PITCH + 4 * (st_c ^ k5p_swz(st_r));  // (rows st_r + 32 i share the permutation)
        const float *src = S + (size_t)st_r * CH_STRIDE + k0 + 4 * st_c;
#pragma unroll
        for (int i = 0; i < CH_TILE_FRAMES / 32; i++)  // rows past nrows: nothing is read, zeros are written
            cp_async16(dst + 32 * i * K5P_PITCH, src + (size_t)32 * i * CH_STRIDE, (st_r + 32 * i) < nrows ? 16 : 0);
        // (the address of a skipped row still lies inside the spill: api.cu pads it by one tile of rows)
        if (tid < K5_KT * 12 / 4)  // 48 chunks of weights
            cp_async16(s_w + stage * K5_KT * 12 + 4 * tid, W + (size_t)k0 * 12 + 4 * tid, 16);
    };

    double acc0[12], acc1[12];  // frames tid and tid + 128
#pragma unroll
    for (int c = 0; c < 12; c++) { acc0[c] = 0.; acc1[c] = 0.; }

    constexpr int N_STAGES_K = (CH_BINS - 1) / K5_KT;  // 256 full stages; bin 4096 is handled after the loop
#pragma unroll
    for (int p = 0; p < K5P_STAGES - 1; p++) {
        issue(p, p * K5_KT);
        cp_async_commit();
    }
#pragma unroll 1
    for (int s = 0; s < N_STAGES_K; s++) {
        cp_async_wait<K5P_STAGES - 2>();  // this thread's copies of stage s have landed
        __syncthreads();                  // ... everyone's; and everyone is done with stage s - 1
        if (s + K5P_STAGES - 1 < N_STAGES_K) issue((s + K5P_STAGES - 1) % K5P_STAGES, (s + K5P_STAGES - 1) * K5_KT);
        cp_async_commit();  // (possibly empty: keeps the group count uniform)
        const int buf = s % K5P_STAGES;
        const float4 *va = reinterpret_cast<const float4 *>(s_s + (size_t)buf * CH_TILE_FRAMES * K5P_PITCH + tid * K5P_PITCH);
        const float4 *vb = reinterpret_cast<const float4 *>(s_s + (size_t)buf * CH_TILE_FRAMES * K5P_PITCH + (tid + 128) * K5P_PITCH);
        const float4 *wq = reinterpret_cast<const float4 *>(s_w + buf * K5_KT * 12);
        const int sw = k5p_swz(tid);  // rows tid and tid + 128 share it
        // (frame a, frame b) ride one FFMA2 per chroma row: .x = frame tid, .y = frame tid + 128; each half is
        // exactly the scalar fmaf(w, s^2, p) of the previous kernel
        cpx p[12];
#pragma unroll
        for (int c = 0; c < 12; c++) p[c] = cpx{0.f, 0.f};
#pragma unroll
        for (int g4 = 0; g4 < K5_KT / 4; g4++) {
            const float4 a4 = va[g4 ^ sw], b4 = vb[g4 ^ sw];
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const cpx sq{av[j] * av[j], bv[j] * bv[j]};  // spectrum.mapv_inplace(|x| x*x), chroma.rs:400
                const float4 w0 = wq[(4 * g4 + j) * 3], w1 = wq[(4 * g4 + j) * 3 + 1], w2 = wq[(4 * g4 + j) * 3 + 2];
                const float wv[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                for (int c = 0; c < 12; c++) p[c] = pfma(cpx{wv[c], wv[c]}, sq, p[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 12; c++) {
            acc0[c] += (double)p[c].x;
            acc1[c] += (double)p[c].y;
        }
    }
    {   // bin 4096: a chunk of its own (the old tiling staged it with 15 zero bins: same sums)
        const float a = (tid < nrows) ? __ldg(S + (size_t)tid * CH_STRIDE + (CH_BINS - 1)) : 0.f;
        const float b = (tid + 128 < nrows) ? __ldg(S + (size_t)(tid + 128) * CH_STRIDE + (CH_BINS - 1)) : 0.f;
        const float a2 = a * a, b2 = b * b;
        const float *wl = W + (size_t)(CH_BINS - 1) * 12;
#pragma unroll
        for (int c = 0; c < 12; c++) {
            const float w = __ldg(wl + c);
            acc0[c] += (double)fmaf(w, a2, 0.f);
            acc1[c] += (double)fmaf(w, b2, 0.f);
        }
    }
    double feat[10];
#pragma unroll
    for (int t = 0; t < 10; t++) feat[t] = 0.;
    if (tid < nf)
        frame_interval_features(acc0, feat, chroma_dbg ? chroma_dbg + ((size_t)sd.c_tile_off * CH_TILE_FRAMES + f0 + tid) * 12 : nullptr);
    if (tid + 128 < nf)
        frame_interval_features(acc1, feat, chroma_dbg ? chroma_dbg + ((size_t)sd.c_tile_off * CH_TILE_FRAMES + f0 + tid + 128) * 12 : nullptr);
    // sum the tile's frames (mean_axis over frames finishes in the summary kernel)
#pragma unroll
    for (int t = 0; t < 10; t++) {
        double v = feat[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[(tid >> 5) * 10 + t] = v;
    }
    __syncthreads();
    if (tid < 10) {
        double v = 0.;
#pragma unroll
        for (int w = 0; w < K5_THREADS / 32; w++) v += s_red[w * 10 + tid];
        tile_partials[((size_t)sd.c_tile_off + tile) * 10 + tid] = v;
    }
}
constexpr size_t K5P_SMEM = (size_t)K5P_STAGES * CH_TILE_FRAMES * K5P_PITCH * 4 + (size_t)K5P_STAGES * K5_KT * 12 * 4 +
                            (K5_THREADS / 32) * 10 * 8;

// ---- launchers ---------------------------------------------------------------
// frame_prefix counts groups of K3_FRAMES_PER_CTA (= 4) frames per song
int launch_stft8192(const float *pcm, const SongDesc *songs, const unsigned int *frame_prefix, int n_songs,
                    unsigned int total_frames, const float *hann, const cpx *tw1, const cpx *tw2,
                    const cpx *tw8192, float *mags, double *cand_mag, double *cand_pitch,
                    unsigned int *cand_count, int variant, cudaStream_t st) {
    if (total_frames == 0) return 0;
    if ((variant & (VARIANT_STFT_V1 | VARIANT_OLD_EPILOGUE)) == 0) {  // the round-2 kernels
        // > 48 KB of dynamic shared memory is an opt-in, per device: set on every launch
        // work items (of four frames) per CTA: as many as still leave ~4 CTAs per resident slot (148 SMs x 3), at most 16
        // (16 against 4 on 1024 tracks: 21.50 against 21.77 ms, profiles/knobs_r02.md); BLISS_B200_STFT_ITEMS overrides
        static const int ipc_env = [] { const char *e = getenv("BLISS_B200_STFT_ITEMS"); return e ? atoi(e) : 0; }();
        const int ipc = ipc_env > 0 ? ipc_env : (int)std::min<unsigned int>(16u, std::max<unsigned int>(1u, total_frames / (148u * 3u * 4u)));
        const unsigned int grid = (total_frames + ipc - 1) / ipc;
        const int fpi = K3_FRAMES_PER_CTA;
        if ((variant & VARIANT_STFT_V3) == 0) {  // 128 threads, two columns per thread (stft8192_v2.cuh): the default
#ifndef BLISS_HOST_EMUL
            if (cudaFuncSetAttribute(stft8192v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2::SMEM_BYTES) != cudaSuccess)
                return -1;
#endif
            BLISS_LAUNCH(stft8192v2_kernel, grid, s2::THREADS, s2::SMEM_BYTES, st, pcm, songs, frame_prefix, n_songs, total_frames, fpi, ipc, hann,
                         tw1, tw2, tw8192, mags, cand_mag, cand_pitch, cand_count);
        } else {  // 256 threads, one column per thread (stft8192_v3.cuh): same speed on B200 (profiles/ncu_r02_stft8192v3_128songs.md)
#ifndef BLISS_HOST_EMUL
            if (cudaFuncSetAttribute(stft8192v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3::SMEM_BYTES) != cudaSuccess)
                return -1;
#endif
            BLISS_LAUNCH(stft8192v3_kernel, grid, s3::THREADS, s3::SMEM_BYTES, st, pcm, songs, frame_prefix, n_songs, total_frames, fpi, ipc, hann,
                         tw1, tw2, tw8192, mags, cand_mag, cand_pitch, cand_count);
        }
        return 1;
    }
    if (variant & VARIANT_OLD_EPILOGUE)
        BLISS_LAUNCH(stft8192_kernel<false>, total_frames, K3_THREADS, 0, st, pcm, songs, frame_prefix, n_songs, hann, tw1, tw2,
                                                                    tw8192, mags, cand_mag, cand_pitch, cand_count);
    else {
        auto go = [&](auto kern) {
            BLISS_LAUNCH(kern, total_frames, K3_THREADS, 0, st, pcm, songs, frame_prefix, n_songs, hann, tw1, tw2, tw8192, mags,
                                                      cand_mag, cand_pitch, cand_count);
        };
        const int var = ((variant & VARIANT_TWPROD) ? K3V_TWPROD : 0) | ((variant & VARIANT_WINSYN) ? K3V_WINSYN : 0) |
                        ((variant & VARIANT_LAY16) ? K3V_LAY16 : 0) | ((variant & VARIANT_ODDSHIFT) ? K3V_ODDSHIFT : 0);
        switch (var) {
            case 1: go(stft8192_kernel<true, 1>); break;
            case 2: go(stft8192_kernel<true, 2>); break;
            case 3: go(stft8192_kernel<true, 3>); break;
            case 4: go(stft8192_kernel<true, 4>); break;
            case 5: go(stft8192_kernel<true, 5>); break;
            case 6: go(stft8192_kernel<true, 6>); break;
            case 7: go(stft8192_kernel<true, 7>); break;
            case 8: go(stft8192_kernel<true, 8>); break;
            case 12: go(stft8192_kernel<true, 12>); break;
            case 15: go(stft8192_kernel<true, 15>); break;
            default:
                if (var == 0) go(stft8192_kernel<true>);
                else if (var & K3V_ODDSHIFT) go(stft8192_kernel<true, 15>);  // other mixes with bit 8192: everything on
                else go(stft8192_kernel<true, 7>);
                break;
        }
    }
    return 1;
}

int launch_tuning(const double *cand_mag, const double *cand_pitch, const unsigned int *cand_count,
                  const SongDesc *songs, int n_songs, int *tuning_idx, int variant, cudaStream_t st) {
    if (n_songs == 0) return 0;
#ifndef BLISS_HOST_EMUL  // (the previous kernel's partial-mask __match_any_sync is not modelled by the host emulation)
    if (variant & VARIANT_OLD_TUNING)
        BLISS_LAUNCH(tuning_kernel, n_songs, K4_THREADS, 0, st, cand_mag, cand_pitch, cand_count, songs, tuning_idx);
    else
#endif
        BLISS_LAUNCH(tuning_select_kernel, n_songs, K4_THREADS, 0, st, cand_mag, cand_pitch, cand_count, songs, tuning_idx);
    return 1;
}

int launch_chroma(const float *mags, const SongDesc *songs, const unsigned int *tile_prefix, int n_songs,
                  unsigned int total_tiles, const float *filt_table, const int *tuning_idx,
                  double *tile_partials, double *chroma_dbg, int variant, cudaStream_t st) {
    if (total_tiles == 0) return 0;
    if (variant & VARIANT_OLD_CHROMA) {
        BLISS_LAUNCH(chroma_kernel, total_tiles, K5_THREADS, 0, st, mags, songs, tile_prefix, n_songs, filt_table,
                                                         tuning_idx, tile_partials, chroma_dbg);
    } else {
        // > 48 KB of dynamic shared memory needs the opt-in, per device: set on every launch
        if (cudaFuncSetAttribute(chroma_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K5P_SMEM) != cudaSuccess)
            return -1;
        BLISS_LAUNCH(chroma_pipe_kernel, total_tiles, K5_THREADS, K5P_SMEM, st, mags, songs, tile_prefix, n_songs, filt_table,
                                                                     tuning_idx, tile_partials, chroma_dbg);
    }
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
void configure_kernels_chroma() {
    BLISS_MAX_SHARED(stft8192v2_kernel);
    BLISS_MAX_SHARED(stft8192v3_kernel);
    BLISS_MAX_SHARED(tuning_select_kernel);
    BLISS_MAX_SHARED(chroma_pipe_kernel);
}

}  // namespace bliss

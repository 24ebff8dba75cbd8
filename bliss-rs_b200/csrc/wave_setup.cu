// wave_setup.cu -- first kernel of every wave: pulls the wave's descriptors (SongDesc array +
// prefix arrays) from pinned host memory into the wave set's device blob with plain loads over
// PCIe, and zeroes the two per-song counters the later kernels accumulate into.
//
// Why a kernel and not cudaMemcpyAsync/cudaMemsetAsync: on the host path the copy engine is busy
// with back-to-back 64..512 MB PCM copies, and a small H2D copy enqueued on another stream is only
// served when that queue runs dry -- measured (BLISS_B200_TRACE_CHUNKS): the kernels of chunks 0..2
// started after chunk 3's copy, 12 ms late, and every later group of chunks likewise.  SM loads from
// mapped host memory do not queue behind the DMA engine.
#include "common.cuh"

namespace bliss {

__global__ void wave_setup_kernel(const uint4 *__restrict__ src_host, uint4 *__restrict__ dst, unsigned int n16,
                                  unsigned int *__restrict__ zcr_count, unsigned int *__restrict__ cand_count,
                                  unsigned int n_songs) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) dst[i] = src_host[i];
    if (i < n_songs) {
        zcr_count[i] = 0u;
        cand_count[i] = 0u;
    }
}

int launch_wave_setup(const void *src_host, void *dst, size_t bytes, unsigned int *zcr_count,
                      unsigned int *cand_count, unsigned int n_songs, cudaStream_t st) {
    const unsigned int n16 = (unsigned int)((bytes + 15) / 16);
    const unsigned int n = n16 > n_songs ? n16 : n_songs;
    if (n == 0) return 0;
    BLISS_LAUNCH(wave_setup_kernel, (n + 255u) / 256u, 256, 0, st, reinterpret_cast<const uint4 *>(src_host),
                                                         reinterpret_cast<uint4 *>(dst), n16, zcr_count, cand_count,
                                                         n_songs);
    return 1;
}

// ---- s16 -> f32 (bliss_b200_analyze_batch_s16) ----------------------------------------------------------
// What the reference's decoder does to signed 16-bit mono 22 050 Hz material before Song::analyze sees it:
// swresample's s16 -> flt conversion, x * (1 / 32768) (src/song/decoder/ffmpeg.rs:36-109; pinned by the decoder
// test of data/s16_mono_22_5kHz.flac, :455-462, which the golden fixture reproduces bit for bit).
// n is a multiple of 4 (chunk layout of api.cu); 8 samples per thread: one 16-byte load, two 16-byte stores.
__global__ void __launch_bounds__(256)
s16_to_f32_kernel(const short *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const float k = 1.0f / 32768.0f;
    if (i + 8 <= n) {
        const int4 v = *reinterpret_cast<const int4 *>(in + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            f[2 * j] = (float)(short)(w[j] & 0xffff) * k;
            f[2 * j + 1] = (float)(short)(w[j] >> 16) * k;
        }
        *reinterpret_cast<float4 *>(out + i) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4 *>(out + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
        for (size_t j = i; j < n; j++) out[j] = (float)in[j] * k;
    }
}

int launch_s16_to_f32(const short *in, float *out, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    const size_t threads = (n + 7) / 8;
    BLISS_LAUNCH(s16_to_f32_kernel, (unsigned int)((threads + 255) / 256), 256, 0, st, in, out, n);
    return 1;
}

// ---- interleaved PCM -> mono f32 (bliss_b200_analyze_batch_pcm, bliss_b200_pcm_to_mono) ------------------
// The sample-format and down-mix steps of the reference's decoders (the sample-rate conversion that follows them for
// sources at another rate is further down):
//   s16 -> flt   x * 2^-15, s32 -> flt   x * 2^-31      swresample's conversions behind
//                                                       src/song/decoder/ffmpeg.rs:36-109 (symphonia's
//                                                       `sample as f32 / 32768.0` rounds identically)
//   stereo       c L + c R, c = (float)sqrt(1/2), products and sum rounded separately: swresample's
//                default FL+FR -> FC matrix for float output, "averaging the channels and multiplying by
//                the square root of 2" in the words of src/song/decoder/symphonia.rs:260-262; pinned by
//                the decoder test of data/s16_stereo_22_5kHz.flac, src/song/decoder/ffmpeg.rs:447-452
//   > 2 channels mean of the channels, summed in channel order: src/song/decoder/symphonia.rs:289-299
// FMT: 1 = s16, 2 = s32, 3 = f32 (include/bliss_b200.h).  Flat over the chunk's frames (the padding frames
// between songs are converted too and never read); n_frames is a multiple of 4: four frames per thread,
// one 16-byte store.
template <int FMT>
__device__ __forceinline__ float pcm_sample(const void *in, size_t i) {
    if (FMT == 1) return __fmul_rn((float)static_cast<const short *>(in)[i], 1.0f / 32768.0f);
    if (FMT == 2) return __fmul_rn((float)static_cast<const int *>(in)[i], 1.0f / 2147483648.0f);
    return static_cast<const float *>(in)[i];
}

template <int FMT>
__global__ void __launch_bounds__(256)
pcm_to_mono_kernel(const void *__restrict__ in, float *__restrict__ out, size_t n_frames, unsigned int channels) {
    const size_t f0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (f0 >= n_frames) return;
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t base = (f0 + k) * channels;
        if (channels == 1) {
            r[k] = pcm_sample<FMT>(in, base);
        } else if (channels == 2) {
            const float c = 0.70710678118654752440f;
            r[k] = __fadd_rn(__fmul_rn(pcm_sample<FMT>(in, base), c), __fmul_rn(pcm_sample<FMT>(in, base + 1), c));
        } else {
            float s = 0.f;
            for (unsigned int ch = 0; ch < channels; ch++) s = __fadd_rn(s, pcm_sample<FMT>(in, base + ch));
            r[k] = __fdiv_rn(s, (float)channels);
        }
    }
    *reinterpret_cast<float4 *>(out + f0) = make_float4(r[0], r[1], r[2], r[3]);
}

int launch_pcm_to_mono(const void *in, float *out, size_t n_frames, int fmt, unsigned int channels, cudaStream_t st) {
    if (n_frames == 0) return 0;
    const unsigned int grid = (unsigned int)(((n_frames + 3) / 4 + 255) / 256);
    if (fmt == 1) BLISS_LAUNCH(pcm_to_mono_kernel<1>, grid, 256, 0, st, in, out, n_frames, channels);
    else if (fmt == 2) BLISS_LAUNCH(pcm_to_mono_kernel<2>, grid, 256, 0, st, in, out, n_frames, channels);
    else BLISS_LAUNCH(pcm_to_mono_kernel<3>, grid, 256, 0, st, in, out, n_frames, channels);
    return 1;
}

// ---- sample-rate conversion to 22 050 Hz (bliss_b200_resample, bliss_b200_analyze_batch_pcm at other rates) ----
// The step of the reference's decoders between the down-mix and Song::analyze: swresample inside the ffmpeg
// decoder (src/song/decoder/ffmpeg.rs:36-109) and rubato's synchronous FFT resampler inside the symphonia one
// (src/song/decoder/symphonia.rs:304-404).  Neither library is part of the reference's tree and the two do not
// agree with each other sample for sample (the reference's own cross-decoder tests compare them through
// tolerances, src/song/decoder/symphonia.rs tests), so there is no bit-level target here: PARITY UNPINNED.
// What runs is the textbook rational resampler, stated so that a public implementation checks it:
// scipy.signal.resample_poly(x, 22050, rate) -- zero-stuff by `up`, Kaiser(beta 5) windowed-sinc low-pass of
// half length 10 max(up, down) at cut-off 1 / max(up, down), keep every `down`-th sample, filter delay removed --
// with the output length of the symphonia decoder, ceil(22050 / rate x n) (:379-380).
// The host (api.cu design_resampler) lays the filter out by phase: row p holds h[p], h[p + up], h[p + 2 up] ...
// (taps4 floats, zero-filled), so that output j reads ONE contiguous row and the inputs i0, i0 - 1, ...:
//   q = (j + pre_remove) down,  i0 = q div up,  p = q mod up,  y[j] = sum_t row_p[t] x[i0 - t]   (f32 FMAs, t rising)
// One thread per output, 1024 outputs per CTA; the CTA finds its song by bisection of the chunk's tile prefix.
struct ResampleJob { unsigned long long in_off, in_len, out_off, out_len; };

__device__ __forceinline__ unsigned int resample_find_job(const unsigned int *__restrict__ tile_prefix, unsigned int n_jobs,
                                                          unsigned int tile) {
    unsigned int lo = 0, hi = n_jobs;
    while (hi - lo > 1) {
        const unsigned int mid = (lo + hi) >> 1;
        if (tile_prefix[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// one output: y[j] = sum_t row[t] x[i0 - t], row in global or shared memory
__device__ __forceinline__ float resample_one(const float *__restrict__ x, unsigned long long in_len, const float4 *row,
                                              unsigned long long i0, unsigned int taps4) {
    float acc = 0.f;
    if (i0 + 1 >= taps4 && i0 < in_len) {  // every tap inside the song
        const float *xp = x + i0;
        for (unsigned int t = 0; t < taps4; t += 4) {
            const float4 c = row[t >> 2];
            acc = fmaf(c.x, xp[-(long long)t], acc);
            acc = fmaf(c.y, xp[-(long long)t - 1], acc);
            acc = fmaf(c.z, xp[-(long long)t - 2], acc);
            acc = fmaf(c.w, xp[-(long long)t - 3], acc);
        }
    } else {
        for (unsigned int t = 0; t < taps4; t += 4) {
            const float4 c = row[t >> 2];
            const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const unsigned long long i = i0 - t - (unsigned int)e;  // wraps past zero: caught by the range test
                const float v = (i0 >= t + (unsigned int)e && i < in_len) ? x[i] : 0.f;
                acc = fmaf(cc[e], v, acc);
            }
        }
    }
    return acc;
}

// General ratio.  TABLE_IN_SMEM: the whole phase table (up x taps4 floats: 26 KB at 48 kHz, 52 KB at 96 kHz) is staged
// in shared memory once per CTA and serves RS_TILES_PER_CTA tiles; each thread then reads its row with 16-byte shared
// loads (consecutive outputs sit in different rows: from global memory that is one L1 line per lane and load).  Rates
// whose table does not fit (up x taps4 x 4 > RS_MAX_SMEM_TABLE) read it through L1 / L2.
constexpr int RS_TILES_PER_CTA = 4;
constexpr size_t RS_MAX_SMEM_TABLE = 160 * 1024;

template <bool TABLE_IN_SMEM>
__global__ void __launch_bounds__(256)
resample_kernel(const float *__restrict__ in, float *__restrict__ out, const ResampleJob *__restrict__ jobs,
                const unsigned int *__restrict__ tile_prefix, unsigned int n_jobs, unsigned int n_tiles,
                const float *__restrict__ tab, unsigned int up, unsigned int down, unsigned int taps4, unsigned int pre_remove) {
#ifdef BLISS_HOST_EMUL
    unsigned char *rs_smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char rs_smem[];
#endif
    const float4 *table = reinterpret_cast<const float4 *>(tab);
    if (TABLE_IN_SMEM) {
        float4 *dst = reinterpret_cast<float4 *>(rs_smem);
        const unsigned int n4 = up * (taps4 >> 2);
        for (unsigned int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(table + i);
        __syncthreads();
        table = dst;
    }
    for (int tt = 0; tt < RS_TILES_PER_CTA; tt++) {
        const unsigned int tile_id = blockIdx.x * RS_TILES_PER_CTA + tt;
        if (tile_id >= n_tiles) break;
        const unsigned int ji = resample_find_job(tile_prefix, n_jobs, tile_id);
        const ResampleJob job = jobs[ji];
        const unsigned long long tile = tile_id - tile_prefix[ji];
        const float *x = in + job.in_off;
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            const unsigned long long j = tile * 1024ull + (unsigned long long)(k * 256) + threadIdx.x;
            if (j >= job.out_len) break;
            const unsigned long long q = (j + pre_remove) * down;
            const unsigned long long i0 = q / up;
            const unsigned int p = (unsigned int)(q - i0 * up);
            out[job.out_off + j] = resample_one(x, job.in_len, table + (size_t)p * (taps4 >> 2), i0, taps4);
        }
    }
}

// Whole-number decimation (up == 1: 44.1 kHz -> D = 2, 88.2 kHz -> D = 4; the filter then has 21 D + 1 taps, taps4 = 22 D).
// With E_r[m] = x[D m + r] the sum splits into D short convolutions at the OUTPUT rate,
//   y[j] = sum_s c[D s] E_0[j' - s] + sum_{r=1}^{D-1} sum_s c[D s + r] E_{D-r}[j' - s - 1],   j' = j + pre_remove,
// so a thread that owns four consecutive outputs slides over 22 + 3 consecutive values of each E_r: the CTA stages its
// stretch of x de-interleaved in shared memory (coalesced loads), every thread pulls each window with seven 16-byte
// loads and runs 4 x 22 FMAs on it from registers.  Loads per FMA: 0.1 against 1 in the general kernel.
template <int D>
__global__ void __launch_bounds__(256)
resample_decimate_kernel(const float *__restrict__ in, float *__restrict__ out, const ResampleJob *__restrict__ jobs,
                         const unsigned int *__restrict__ tile_prefix, unsigned int n_jobs, const float *__restrict__ tab,
                         unsigned int pre_remove) {
    constexpr int S = 22, W = 1024 + 32;
    __shared__ __align__(16) float E[D][W];
    __shared__ __align__(16) float C[D][24];
    const unsigned int ji = resample_find_job(tile_prefix, n_jobs, blockIdx.x);
    const ResampleJob job = jobs[ji];
    const long long j_first = (long long)(blockIdx.x - tile_prefix[ji]) * 1024;
    const long long m_first = j_first + (long long)pre_remove - S - 1;  // E_r[m_first + li] sits at E[r][li]
    const float *x = in + job.in_off;
    if (threadIdx.x < D * 24) {
        const int r = threadIdx.x / 24, sidx = threadIdx.x % 24;
        C[r][sidx] = sidx < S ? tab[D * sidx + r] : 0.f;
    }
    for (int li = threadIdx.x; li < W; li += 256) {
        const long long i = (m_first + li) * D;
        if (i >= 0 && i + D <= (long long)job.in_len) {
            if (D == 2) {
                const float2 v = *reinterpret_cast<const float2 *>(x + i);
                E[0][li] = v.x; E[1 % D][li] = v.y;
            } else {
                const float4 v = *reinterpret_cast<const float4 *>(x + i);
                E[0][li] = v.x; E[1 % D][li] = v.y; E[2 % D][li] = v.z; E[3 % D][li] = v.w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < D; r++) E[r][li] = (i + r >= 0 && i + r < (long long)job.in_len) ? x[i + r] : 0.f;
        }
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < D; r++) {
        const int phase = r == 0 ? 0 : D - r, shift = r == 0 ? 0 : 1;
        float w[28], c[24];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const float4 v = *reinterpret_cast<const float4 *>(&E[phase][4 * threadIdx.x + 4 * k]);
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const float4 v = *reinterpret_cast<const float4 *>(&C[r][4 * k]);
            c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int sidx = 0; sidx < S; sidx++)
#pragma unroll
            for (int rr = 0; rr < 4; rr++) acc[rr] = fmaf(c[sidx], w[rr + S + 1 - sidx - shift], acc[rr]);
    }
    const long long j = j_first + 4 * (long long)threadIdx.x;
    float *o = out + job.out_off + j;
    if (j + 4 <= (long long)job.out_len) {
        *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
        for (int rr = 0; rr < 4; rr++)
            if (j + rr < (long long)job.out_len) o[rr] = acc[rr];
    }
}

// Any ratio whose `down` is a multiple of 4 and whose `up` fits a CTA (48 / 96 / 32 / 24 / 16 / 12 / 8 kHz ...).
// Outputs j = r + up k (k = 0, 1, ...) share the filter row AND the alignment of their input window: their inputs
// sit exactly k down samples apart.  So thread r keeps its row in REGISTERS for the whole tile -- re-indexed once so
// that it lines up with the 16-byte-aligned window that contains x[i0 - T4 + 1 .. i0] (c'[u] = row[T4 - 1 + o - u],
// o = the window's misalignment, constant over k because 4 | down) -- and each output is T4/4 + 1 aligned 16-byte
// shared loads of x (staged once per CTA, coalesced, zero-filled outside the song) and as many FMAs as taps + 4.
// No per-FMA load is left: the first cut issues one load per operand.  nb periods run side by side in a CTA
// (thread = r + up b), RS_PERIODS_PER_THREAD outputs per thread.
constexpr int RS_PERIODS_PER_THREAD = 16;

// c[u] = row[T4 - 1 + O - u] for u = 0 .. T4 + 3 (zero outside the row)
template <int T4, int O>
__device__ __forceinline__ void resample_place_row(const float4 *__restrict__ row4, float (&c)[T4 + 4]) {
#pragma unroll
    for (int u = 0; u < T4 + 4; u++) c[u] = 0.f;
#pragma unroll
    for (int n = 0; n < T4 / 4; n++) {
        const float4 v = __ldg(row4 + n);
        c[T4 - 1 + O - 4 * n] = v.x;      // t = 4 n
        c[T4 - 2 + O - 4 * n] = v.y;
        c[T4 - 3 + O - 4 * n] = v.z;
        c[T4 - 4 + O - 4 * n] = v.w;
    }
}

template <int T4>
__global__ void __launch_bounds__(512)
resample_periodic_kernel(const float *__restrict__ in, float *__restrict__ out, const ResampleJob *__restrict__ jobs,
                         const unsigned int *__restrict__ tile_prefix, unsigned int n_jobs, const float *__restrict__ tab,
                         unsigned int up, unsigned int down, unsigned int pre_remove, unsigned int nb, unsigned int x_floats) {
    constexpr int NW = T4 / 4 + 1;
#ifdef BLISS_HOST_EMUL
    unsigned char *rs_smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(16) unsigned char rs_smem[];
#endif
    float *xs = reinterpret_cast<float *>(rs_smem);
    const unsigned int ji = resample_find_job(tile_prefix, n_jobs, blockIdx.x);
    const ResampleJob job = jobs[ji];
    const unsigned long long tile_out = (unsigned long long)up * nb * RS_PERIODS_PER_THREAD;
    const unsigned long long j_first = (unsigned long long)(blockIdx.x - tile_prefix[ji]) * tile_out;  // a multiple of up
    const unsigned long long base_q = (j_first + pre_remove) * down;
    const long long first0 = (long long)(base_q / up) - (T4 - 1);
    const long long x0 = first0 - (((first0 % 4) + 4) % 4);  // the tile's first input sample (may be negative), 4-aligned
    const float *x = in + job.in_off;
    for (unsigned int li = threadIdx.x * 4; li < x_floats; li += blockDim.x * 4) {
        const long long i = x0 + li;
        float4 v;
        if (i >= 0 && i + 4 <= (long long)job.in_len) {
            v = *reinterpret_cast<const float4 *>(x + i);
        } else {
            v.x = (i >= 0 && i < (long long)job.in_len) ? x[i] : 0.f;
            v.y = (i + 1 >= 0 && i + 1 < (long long)job.in_len) ? x[i + 1] : 0.f;
            v.z = (i + 2 >= 0 && i + 2 < (long long)job.in_len) ? x[i + 2] : 0.f;
            v.w = (i + 3 >= 0 && i + 3 < (long long)job.in_len) ? x[i + 3] : 0.f;
        }
        *reinterpret_cast<float4 *>(xs + li) = v;
    }
    __syncthreads();
    const unsigned int b = threadIdx.x / up, r = threadIdx.x - b * up;
    if (b >= nb) return;
    const unsigned long long qr = base_q + (unsigned long long)r * down;
    const unsigned long long i0 = qr / up;
    const unsigned int p = (unsigned int)(qr - i0 * up);
    const long long first = (long long)i0 - (T4 - 1);
    const int o = (int)(((first % 4) + 4) % 4);
    const unsigned int lo = (unsigned int)((first - o) - x0);
    float c[4 * NW];
    {   // the row comes in as T4/4 16-byte loads; one of four fully unrolled placements puts it where the window wants it
        const float4 *row4 = reinterpret_cast<const float4 *>(tab + (size_t)p * T4);
        switch (o) {
        case 0: resample_place_row<T4, 0>(row4, c); break;
        case 1: resample_place_row<T4, 1>(row4, c); break;
        case 2: resample_place_row<T4, 2>(row4, c); break;
        default: resample_place_row<T4, 3>(row4, c); break;
        }
    }
#pragma unroll 1
    for (int kk = 0; kk < RS_PERIODS_PER_THREAD; kk++) {
        const unsigned int k = b + nb * (unsigned int)kk;
        const unsigned long long j = j_first + r + (unsigned long long)up * k;
        if (j >= job.out_len) break;
        const float4 *w = reinterpret_cast<const float4 *>(xs + lo + (size_t)k * down);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int n = 0; n < NW; n++) {
            const float4 v = w[n];
            a0 = fmaf(c[4 * n], v.x, a0);
            a1 = fmaf(c[4 * n + 1], v.y, a1);
            a2 = fmaf(c[4 * n + 2], v.z, a2);
            a3 = fmaf(c[4 * n + 3], v.w, a3);
        }
        out[job.out_off + j] = (a0 + a1) + (a2 + a3);
    }
}

// Which kernel serves a ratio, the size of its tiles (the host builds the chunk's tile prefix from it) and the row
// length its table is laid out with (a periodic kernel exists for a few row lengths; rows are zero-filled up to it).
static const unsigned int kPeriodicTaps[] = {24, 32, 48, 64, 92};

ResamplePlan resample_plan(unsigned int up, unsigned int down, unsigned int taps4_needed, int variant) {
    ResamplePlan pl;
    pl.kind = RS_KIND_GENERAL;
    pl.tile_out = 1024;
    pl.taps4 = taps4_needed;
    pl.nb = 0;
    if (up == 1 && (down == 2 || down == 4) && taps4_needed == 22 * down) {
        pl.kind = RS_KIND_DECIMATE;
    } else if (up > 1 && up <= 448 && down % 4 == 0) {
        for (unsigned int t : kPeriodicTaps)
            if (t >= taps4_needed) {
                pl.kind = RS_KIND_PERIODIC;
                pl.taps4 = t;
                pl.nb = up >= 224 ? 1 : 224 / up;
                pl.tile_out = up * pl.nb * RS_PERIODS_PER_THREAD;
                break;
            }
    }
    if (pl.kind == RS_KIND_GENERAL && (size_t)up * pl.taps4 * 4 > RS_MAX_SMEM_TABLE) pl.kind = RS_KIND_FIRST_CUT;
    if (variant) {  // the first cut on the same table and tiles
        if (pl.kind == RS_KIND_DECIMATE || pl.kind == RS_KIND_GENERAL) pl.kind = RS_KIND_FIRST_CUT;
        else if (pl.kind == RS_KIND_PERIODIC) pl.kind = RS_KIND_FIRST_CUT_TILED;
    }
    return pl;
}

// One output per thread over tiles of any size (RS_KIND_FIRST_CUT_TILED: the first cut on a periodic kernel's tiles)
__global__ void __launch_bounds__(256)
resample_first_cut_tiled_kernel(const float *__restrict__ in, float *__restrict__ out, const ResampleJob *__restrict__ jobs,
                                const unsigned int *__restrict__ tile_prefix, unsigned int n_jobs, const float *__restrict__ tab,
                                unsigned int up, unsigned int down, unsigned int taps4, unsigned int pre_remove, unsigned int tile_out) {
    const unsigned int ji = resample_find_job(tile_prefix, n_jobs, blockIdx.x);
    const ResampleJob job = jobs[ji];
    const unsigned long long j_first = (unsigned long long)(blockIdx.x - tile_prefix[ji]) * tile_out;
    for (unsigned int jj = threadIdx.x; jj < tile_out; jj += blockDim.x) {
        const unsigned long long j = j_first + jj;
        if (j >= job.out_len) break;
        const unsigned long long q = (j + pre_remove) * down;
        const unsigned long long i0 = q / up;
        const unsigned int p = (unsigned int)(q - i0 * up);
        out[job.out_off + j] = resample_one(in + job.in_off, job.in_len, reinterpret_cast<const float4 *>(tab) + (size_t)p * (taps4 >> 2), i0, taps4);
    }
}

template <int T4>
static int launch_periodic(const float *in, float *out, const ResampleJob *jobs, const unsigned int *tile_prefix, unsigned int n_jobs,
                           unsigned int n_tiles, const float *tab, unsigned int up, unsigned int down, unsigned int pre_remove,
                           unsigned int nb, cudaStream_t st) {
    const unsigned int x_floats = nb * RS_PERIODS_PER_THREAD * down + 4 * (T4 / 4 + 1) + 8;
    const size_t smem = (size_t)x_floats * 4;
    const unsigned int threads = (up * nb + 31u) / 32u * 32u;
#ifndef BLISS_HOST_EMUL
    if (smem > 200 * 1024) return -1;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(resample_periodic_kernel<T4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
#endif
    BLISS_LAUNCH(resample_periodic_kernel<T4>, n_tiles, threads, smem, st, in, out, jobs, tile_prefix, n_jobs, tab, up, down, pre_remove,
                 nb, x_floats);
    return 1;
}

// n_tiles counts tiles of plan.tile_out outputs (the tile prefix the host built from the same plan)
int launch_resample(const float *in, float *out, const void *jobs_v, const unsigned int *tile_prefix, unsigned int n_jobs,
                    unsigned int n_tiles, const float *tab, unsigned int up, unsigned int down, unsigned int pre_remove,
                    const ResamplePlan &pl, cudaStream_t st) {
    if (n_jobs == 0 || n_tiles == 0) return 0;
    const ResampleJob *jobs = static_cast<const ResampleJob *>(jobs_v);
    const unsigned int taps4 = pl.taps4;
    switch (pl.kind) {
    case RS_KIND_DECIMATE:
        if (down == 2) BLISS_LAUNCH(resample_decimate_kernel<2>, n_tiles, 256, 0, st, in, out, jobs, tile_prefix, n_jobs, tab, pre_remove);
        else BLISS_LAUNCH(resample_decimate_kernel<4>, n_tiles, 256, 0, st, in, out, jobs, tile_prefix, n_jobs, tab, pre_remove);
        return 1;
    case RS_KIND_PERIODIC:
        switch (taps4) {
        case 24: return launch_periodic<24>(in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down, pre_remove, pl.nb, st);
        case 32: return launch_periodic<32>(in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down, pre_remove, pl.nb, st);
        case 48: return launch_periodic<48>(in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down, pre_remove, pl.nb, st);
        case 64: return launch_periodic<64>(in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down, pre_remove, pl.nb, st);
        case 92: return launch_periodic<92>(in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down, pre_remove, pl.nb, st);
        default: return -1;
        }
    case RS_KIND_FIRST_CUT_TILED:
        BLISS_LAUNCH(resample_first_cut_tiled_kernel, n_tiles, 256, 0, st, in, out, jobs, tile_prefix, n_jobs, tab, up, down, taps4,
                     pre_remove, pl.tile_out);
        return 1;
    default:
        break;
    }
    const unsigned int grid = (n_tiles + RS_TILES_PER_CTA - 1) / RS_TILES_PER_CTA;
    const size_t table_bytes = (size_t)up * taps4 * 4;
    if (pl.kind == RS_KIND_GENERAL) {
#ifndef BLISS_HOST_EMUL
        if (table_bytes > 48 * 1024 &&
            cudaFuncSetAttribute(resample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes) != cudaSuccess)
            return -1;
#endif
        BLISS_LAUNCH(resample_kernel<true>, grid, 256, table_bytes, st, in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down,
                     taps4, pre_remove);
    } else {
        BLISS_LAUNCH(resample_kernel<false>, grid, 256, 0, st, in, out, jobs, tile_prefix, n_jobs, n_tiles, tab, up, down, taps4,
                     pre_remove);
    }
    return 1;
}

}  // namespace bliss

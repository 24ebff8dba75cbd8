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
// The sample-format and down-mix steps of the reference's decoders for sources that already run at
// 22 050 Hz (a resampler is not part of this library):
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

}  // namespace bliss

// common.cuh -- shared declarations of the B200 analysis pipeline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_regs.cuh"

// Kernel launches go through one macro.  In the product build (nvcc) it IS the launch statement.  Under
// BLISS_HOST_EMUL -- defined only by the CPU test-suite's host emulation (tests/cpu_emul: g++ against a stub
// cuda_runtime.h), never by bliss-rs_b200/_build.py -- it runs the kernel's source on the host, thread by thread,
// at the point of the call.  There is no CPU path in the product: bliss_b200_init fails without a device.
#ifdef BLISS_HOST_EMUL
#define BLISS_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(emu::as_dim3(grid), (unsigned)(block), [=] { kern(__VA_ARGS__); }, (size_t)(smem))
#else
#define BLISS_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

namespace bliss {

constexpr int SAMPLE_RATE = 22050;        // src/lib.rs:143
constexpr int PV_WIN = 512;               // SpectralDesc / BPMDesc WINDOW_SIZE (timbral.rs:40, temporal.rs:40)
constexpr int PV_HOP_TIMBRAL = 128;       // timbral.rs:41
constexpr int PV_HOP_TEMPO = 256;         // temporal.rs:41
constexpr int CH_WIN = 8192;              // chroma.rs:39
constexpr int CH_HOP = 2205;              // chroma.rs:74
constexpr int CH_BINS = 4097;
#ifndef BLISS_CH_STRIDE
#define BLISS_CH_STRIDE 4128  // rows on 128-byte lines (A/B of round 2: chroma contraction 7.51 -> 6.86 ms, stft8192 -0.3 ms against 4104)
#endif
constexpr int CH_STRIDE = BLISS_CH_STRIDE;  // padded row of the magnitude spill (16 B aligned rows)
static_assert(CH_STRIDE >= 4100 && CH_STRIDE % 4 == 0, "a row holds 4097 magnitudes, 16-byte aligned");
constexpr int CH_TILE_FRAMES = 256;      // chroma frames per CTA of chroma_kernel (2 per thread)
constexpr int CH_MAX_PEAKS = 714;         // max local maxima among centre bins 57..1483
constexpr int LOUD_WIN = 1024;            // misc.rs:44
constexpr int MIN_SAMPLES = 8192;         // song/mod.rs:417-430

// One entry per song of the wave currently on the device.  All offsets are in
// elements of the array they index.
struct SongDesc {
    unsigned long long pcm_off;   // into the PCM buffer (samples)
    unsigned long long mag_off;   // into the 8192-STFT magnitude spill (rows of CH_STRIDE floats)
    unsigned long long cand_off;  // into the pip-track candidate arrays
    unsigned int n;               // samples
    unsigned int n_s;             // timbral frames  (n-512)/128+1        song/mod.rs:458-463
    unsigned int n_t;             // tempo frames    (n-512)/256+1        song/mod.rs:435-441
    unsigned int n_c;             // chroma frames   ceil(n/2205) in f32  utils.rs:30
    unsigned int n_c_comp;        // chroma frames actually transformed (zip truncation, utils.rs:44-47)
    unsigned int n_l;             // loudness chunks ceil(n/1024)         song/mod.rs:478
    unsigned int s_off;           // into per-timbral-frame arrays
    unsigned int t_off;           // into per-tempo-frame arrays
    unsigned int l_off;           // into loudness chunk array
    unsigned int e_off;           // into 256-sample block energies
    unsigned int c_tile_off;      // into chroma tile partials
    unsigned int bpm_off;         // into the bpm list
    unsigned int valid;           // 0 => too short, skipped everywhere
    unsigned int pad_;
};

struct PvocTables {   // device pointers
    const float *win;     // [512] hanningz, aubio.rs:150-154
    const cpx *twA;       // [16][32]  W512^(lane*k1)
};

// BLISS_B200_VARIANT (environment, read by bliss_b200_init): bit mask that switches a kernel back to its
// previous implementation, for A/B timing and bisecting on the GPU box.  0 = current kernels.
enum {
    VARIANT_OLD_EPILOGUE = 1,  // stft8192_kernel: pass 3 through shared memory + one bin per untangle
    VARIANT_OLD_TUNING = 2,    // tuning_kernel: 8-pass radix select
    VARIANT_OLD_CHROMA = 4,    // chroma_kernel: thread-per-frame tiles staged through shared memory
    VARIANT_OLD_ACF = 8,       // beattrack_kernel: one autocorrelation lag at a time, scalar loads
    VARIANT_BT512 = 16,        // beattrack_kernel: 512 threads per song (3 songs per SM) instead of 128 (8 per SM)
    // (bit 32 was the radix-64 x radix-64 cut of the chroma STFT: measured slower twice -- 28.5 against 28.1 ms in round 1,
    // 26.5 against 24.3 ms with its load cuts in round 2, profiles/ab_r02.md -- and removed)
    VARIANT_TWPROD = 64,       // stft8192: pass-1 twiddles from 4 loads + 11 products instead of 15 loads (experimental)
    VARIANT_STFT_PAIRS = 256,  // STFT micro-benchmark: hop-256 frames j, j+1 share one FFT, 128-byte magnitude stores (experimental)
    VARIANT_PV_TWPROD = 512,   // pvoc512: phase-A twiddles from 4 per-lane registers + 11 products instead of 15 smem loads per pair (experimental)
    VARIANT_PV_PAIRDESC = 1024, // pvoc512: both frames' descriptors reduced / finished together, MUFU-only magnitudes on 2^30-scaled data (experimental; bit-identical results)
    VARIANT_PV_ZPOS4 = 2048,   // pvoc512: natural-order tile padded k + (k >> 4): conflict-free stores as well as loads (experimental; bit-identical results)
    VARIANT_LAY16 = 4096,      // stft8192: FFT buffer without the per-16 padding: conflict-free mirror loads in the pair epilogue (experimental; bit-identical results)
    VARIANT_ODDSHIFT = 8192,   // stft8192 (both cuts, also with bit 32): frames starting on an odd sample are transformed rotated by one sample, so that their pairs load as aligned 64-bit words (experimental)
    VARIANT_WINSYN = 128,      // stft8192 (both cuts, also with bit 32): Hann window from the thread's phase (2 FFMA2 per pair) instead of 16 (radix-64: 64) loads (experimental)
    // Bits 64 ... 8192 were written blind at the end of round 1 and A/B-timed on one B200 at the start of round 2
    // (profiles/ab_r02.md: all on = 59.7 ms per 1024-track step against 67.4 ms, every bit bit-reproducible): they
    // are now what the library runs.  Internally the bit still means "cut on"; the PUBLIC mask (BLISS_B200_VARIANT,
    // bliss_b200_set_variant) is XORed with VARIANT_PROMOTED, so that 0 = the promoted kernels and a set bit goes
    // back to the kernel measured in round 1, like bits 1..32.
    VARIANT_PVOC_V1 = 16384,   // pvoc512: the round-1 kernel (with its promoted cuts) instead of pvoc512v2_kernel
    VARIANT_STFT_V1 = 32768,   // stft8192: the round-1 kernel (with its promoted cuts) instead of stft8192v2_kernel (bits 1 and 32 imply it)
    VARIANT_OLD_DIST = 65536,  // distance matrix: the round-1 scalar kernel instead of distance_matrix_diag_kernel
    VARIANT_STFT_V3 = 131072,  // stft8192: the 256-thread / one-column kernel (stft8192_v3.cuh) instead of the 128-thread / two-column one; measured at the same 21.8 ms
    VARIANT_ACF_4LAGS = 262144, // beattrack_kernel: round 1's autocorrelation (four consecutive lags per thread) instead of the balanced lag pairs
    VARIANT_RESAMPLE_V1 = 524288, // resampler: the first cut (one output per thread, filter rows from global memory) instead of the decimation / shared-table kernels
    VARIANT_PROMOTED = 64 | 128 | 256 | 512 | 1024 | 2048 | 4096 | 8192,
};

// Sample-rate conversion (wave_setup.cu): which kernel serves a ratio, its tile size and the row length of its table
enum { RS_KIND_FIRST_CUT = 0, RS_KIND_GENERAL = 1, RS_KIND_DECIMATE = 2, RS_KIND_PERIODIC = 3, RS_KIND_FIRST_CUT_TILED = 4 };
struct ResamplePlan {
    int kind;
    unsigned int tile_out, taps4, nb;
};

// Song lookup for flat work lists: largest s with prefix[s] <= item (prefix has n_songs+1
// entries, prefix[n_songs] = total).  Starts from the proportional guess -- exact for equal-length
// songs, where it costs one round trip instead of log2(n_songs) dependent loads -- and falls
// back to bisection when the guess is far off (mixed durations).
__device__ __forceinline__ int find_song(const unsigned int *__restrict__ prefix, int n_songs,
                                         unsigned int item) {
    const unsigned int total = __ldg(prefix + n_songs);
    int s = (int)((float)item * ((float)n_songs / (float)(total ? total : 1u)));  // a guess: float is plenty
    s = min(max(s, 0), n_songs - 1);
#pragma unroll 1
    for (int tries = 0; tries < 4; tries++) {
        const unsigned int lo = __ldg(prefix + s), hi = __ldg(prefix + s + 1);
        if (item < lo) s--;
        else if (item >= hi) s++;
        else return s;
    }
    int lo = 0, hi = n_songs;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- fused feature-row exchange (gather.cu, finalize.cu) -------------------------------------
// Multi-GPU runs shard songs over ranks and every rank needs ALL feature rows before the all-pairs
// distance.  Instead of a collective after the analysis, finalize_kernel stores
// each finished row straight into every rank's gather buffer through peer-mapped pointers.
constexpr int MAX_PEERS = 8;  // one NVSwitch box
struct PeerRows {
    float *base[MAX_PEERS];               // rank r's gather buffer of the open epoch (own rank included)
    int n_peers;                          // 0: no exchange
    unsigned int row_offset, row_stride;  // global row of local song i = row_offset + i * row_stride
};

}  // namespace bliss

/*
 * bliss_b200.h -- C ABI of the B200-native implementation of bliss-rs's per-song
 * analysis hot path (Song::analyze over decoded mono f32 22 050 Hz PCM) and of its
 * feature-vector distances.  This is the drop-in boundary: the entry points below are
 * what a `#[cfg(feature = "b200")]` build of the crate binds through `extern "C"`
 * (see INTEGRATION.md for the Rust side).  Plain pointers and sizes only.
 *
 * bliss-rs has no FFI / plugin interface of its own: the seam is
 * the Rust API.  Each entry point cites the Rust item whose body it replaces, paths
 * relative to the reference checkout (crate bliss-audio 0.13.0).
 *
 * Threading: every function is thread-safe; calls on one process serialise on an
 * internal lock per device context (the reference calls Song::analyze from `cores`
 * worker threads, src/song/decoder.rs:304-329 -- those workers should instead hand
 * their decoded buffers to ONE bliss_b200_analyze_batch call).
 * Ownership: inputs are borrowed for the duration of the call and never written;
 * outputs are caller-allocated.
 */
#ifndef BLISS_B200_H
#define BLISS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes of the calls ------------------------------------------------ */
#define BLISS_B200_OK 0
#define BLISS_B200_E_CUDA (-1)     /* a CUDA runtime call failed; see bliss_b200_last_error() */
#define BLISS_B200_E_ARG (-2)      /* bad argument (null pointer, misaligned device buffer, bad version ...) */
#define BLISS_B200_E_NOT_INIT (-3) /* bliss_b200_init() has not been called */
#define BLISS_B200_E_NOMEM (-4)    /* one song alone does not fit the workspace limit */
#define BLISS_B200_E_NO_DEVICE (-5) /* no usable CUDA device: there is NO CPU fallback */
#define BLISS_B200_E_TIMEOUT (-6)  /* a peer never reached the gather barrier */
#define BLISS_B200_E_UNSUPPORTED (-7) /* input this library does not take (a sample rate other than 22 050 Hz) */

/* ---- per-song status (BlissResult<Analysis>) ------------------------------------ */
#define BLISS_B200_SONG_OK 0
/* BlissError::AnalysisError("empty or too short song."), src/song/mod.rs:417-430 */
#define BLISS_B200_SONG_TOO_SHORT 1
#define BLISS_B200_SONG_INTERNAL 2

/* FeaturesVersion, src/lib.rs:151-186: 2 = Version2 (LATEST, 23 floats), 1 = Version1 (20) */
#define BLISS_B200_FEATURES_V1 1
#define BLISS_B200_FEATURES_V2 2
#define BLISS_B200_SAMPLE_RATE 22050 /* src/lib.rs:143 */

/* Binds this process to CUDA device `device` (>= 0), builds the constant tables
 * (windows, twiddles, the 100 chroma filterbanks of src/chroma.rs:197-267).
 * Idempotent for the same device.  Fails with E_NO_DEVICE when no GPU is present. */
int bliss_b200_init(int device);
/* Every B200 of the box behind ONE process (the reference is one process: worker threads + a channel,
 * src/song/decoder.rs:282-331): contexts on devices 0 .. n_devices-1 (n_devices <= 0: all visible devices).
 * Afterwards one call of bliss_b200_analyze_batch / _s16 / _pcm deals its songs longest-first over all devices (one
 * host thread, copy stream and PCM ring per device) and bliss_b200_distance_matrix splits its rows; results are
 * bit-identical to a single-device call.  The device-pointer entry points keep using device 0.
 * Returns the number of devices in use (> 0) or a negative error code. */
int bliss_b200_init_devices(int n_devices);
/* Number of devices the host-buffer calls use (0 before init). */
int bliss_b200_device_count(void);
void bliss_b200_shutdown(void);
/* Upper bound on device scratch memory a call may hold (default: 40% of the device). */
int bliss_b200_set_workspace_limit(uint64_t bytes);

/* Diagnostic switch (no counterpart in the reference): bit mask that sends a kernel back to its previous
 * implementation -- 1 stft8192 epilogue, 2 tuning select, 4 chroma contraction, 8 autocorrelation,
 * 16 beat-tracker CTA width, 32 radix-64 cut of the 8192-point FFT -- or to an experimental cut that is not
 * the default yet: 64 product twiddles, 128 synthesised Hann window in pass 1 of the 8192-point FFT, 256 hop-256
 * frame pairs in the STFT micro-benchmark, 512 product twiddles in the 512-point FFT, 1024 pair-wise descriptor
 * reductions + MUFU-only magnitudes and 2048 a conflict-free tile padding in the same kernel,
 * 4096 a conflict-free FFT-buffer layout and 8192 aligned loads for odd-start frames in the 8192-point kernel
 * (bliss-rs_b200/csrc/common.cuh).  0 = the current kernels.  Returns the previous mask; BLISS_B200_VARIANT in the environment sets the
 * initial one.  Used by the A/B timings in profiles/ and by the test that all implementations agree. */
int bliss_b200_set_variant(int mask);

const char *bliss_b200_strerror(int code);
const char *bliss_b200_last_error(void);
/* FeaturesVersion::feature_count, src/lib.rs:181-186 (0 for an unknown version) */
uint32_t bliss_b200_feature_count(uint16_t features_version);

/* Song::analyze_with_options(sample_array, options), src/song/mod.rs:413-508.
 * `pcm`: n host floats, mono, 22 050 Hz.  `out`: feature_count floats.
 * Returns a per-song status (>= 0) or a negative call error. */
int bliss_b200_analyze(const float *pcm, uint64_t n_samples, uint16_t features_version, float *out);

/* The batching seam of Decoder::analyze_paths_with_options (src/song/decoder.rs:278-332):
 * n_songs decoded buffers in HOST memory (pinned memory copies fastest), analysed together.
 * out: n_songs x feature_count, row-major.  status: n_songs entries.  One bad song never
 * fails the batch (like the reference, which sends errors as items, :319-325). */
int bliss_b200_analyze_batch(const float *const *pcm, const uint64_t *n_samples, uint32_t n_songs,
                             uint16_t features_version, float *out, int32_t *status);

/* Decode-side feed for 16-bit sources: the same call with the decoder's signed 16-bit mono 22 050 Hz samples
 * as they are BEFORE the reference's resampler turns them into f32 (swresample s16 -> flt: x / 32768,
 * src/song/decoder/ffmpeg.rs:36-109; the decoder test :455-462 pins that conversion bit for bit).  The
 * conversion runs on the device, so half the bytes cross PCIe.  Results are bit-identical to
 * bliss_b200_analyze_batch on the converted samples. */
int bliss_b200_analyze_batch_s16(const int16_t *const *pcm, const uint64_t *n_samples, uint32_t n_songs,
                                 uint16_t features_version, float *out, int32_t *status);

/* Decode-side feed, general form: interleaved frames as the container's codec delivers them -- the sample-format
 * conversion, the down-mix to mono and the conversion to 22 050 Hz of the reference's decoders run on the device
 * behind each chunk's copy:
 *   s16 / s32 -> f32   x * 2^-15 / x * 2^-31 (swresample's conversions, src/song/decoder/ffmpeg.rs:36-109;
 *                      symphonia's `as f32 / 32768.0` rounds identically)
 *   2 channels         c L + c R with c = (float)sqrt(1/2) (swresample's stereo -> mono matrix for float output,
 *                      src/song/decoder/symphonia.rs:260-262; pinned by the decoder test of
 *                      data/s16_stereo_22_5kHz.flac, src/song/decoder/ffmpeg.rs:447-452)
 *   > 2 channels       mean of the channels in channel order (src/song/decoder/symphonia.rs:289-299)
 *   sample_rate        other than 22 050 Hz: the mono signal is converted on the device behind the down-mix
 *                      (bliss_b200_resample below says how, and what that is and is not pinned against)
 * pcm[i]: n_frames[i] frames of `channels` samples; one format and one rate per call.  Results are bit-identical
 * to bliss_b200_analyze_batch on the output of bliss_b200_pcm_to_mono (then bliss_b200_resample). */
#define BLISS_B200_PCM_S16 1
#define BLISS_B200_PCM_S32 2
#define BLISS_B200_PCM_F32 3
#define BLISS_B200_PCM_MAX_CHANNELS 8
#define BLISS_B200_MIN_SAMPLE_RATE 1000u
#define BLISS_B200_MAX_SAMPLE_RATE 768000u
int bliss_b200_analyze_batch_pcm(const void *const *pcm, const uint64_t *n_frames, uint32_t n_songs,
                                 int sample_format, uint32_t channels, uint32_t sample_rate,
                                 uint16_t features_version, float *out, int32_t *status);
/* The conversion alone (host in, host out): what PreAnalyzedSong::sample_array (src/song/decoder.rs:64) holds
 * for such a source.  out: n_frames floats. */
int bliss_b200_pcm_to_mono(const void *pcm, uint64_t n_frames, int sample_format, uint32_t channels, float *out);

/* Sample-rate conversion of a mono f32 signal to 22 050 Hz (host in, host out): the step between the down-mix and
 * Song::analyze in the reference's decoders -- swresample in the ffmpeg decoder (src/song/decoder/ffmpeg.rs:36-109),
 * rubato's synchronous FFT resampler in the symphonia decoder (src/song/decoder/symphonia.rs:304-404).  PARITY
 * UNPINNED: both are third-party libraries outside the reference's tree, they do not agree with each other sample
 * for sample (the reference compares its two decoders through tolerances), and neither can run here.  What this
 * computes is the textbook rational-ratio polyphase resampler in the form scipy.signal.resample_poly(x, 22050, rate)
 * publishes it (Kaiser beta 5 windowed sinc, half length 10 max(up, down), cut-off at the lower Nyquist frequency,
 * delay removed), which the tests check it against to f32 rounding; the output length is the symphonia decoder's,
 * ceil(22050 / rate * n) (:379-380), also returned by bliss_b200_resampled_len.
 * out_capacity: floats `out` can hold; *n_out (may be NULL) receives the length.  22 050 Hz in is a copy. */
uint64_t bliss_b200_resampled_len(uint64_t n_samples, uint32_t sample_rate);
int bliss_b200_resample(const float *pcm, uint64_t n_samples, uint32_t sample_rate, float *out, uint64_t out_capacity,
                        uint64_t *n_out);

/* Same, PCM already resident in device memory: song i is d_pcm[offsets[i] .. offsets[i]+n_samples[i]).
 * d_pcm must be 16-byte aligned; offsets/n_samples/status are HOST arrays; d_out is a DEVICE
 * buffer (n_songs x feature_count).  Work is enqueued on `cuda_stream` (a cudaStream_t, may be
 * NULL for the legacy stream) and the call returns without synchronising.  This is also the
 * form BlissCueFile::get_songs needs (sub-slices of one decoded buffer, src/cue.rs:208-243). */
int bliss_b200_analyze_batch_device(const float *d_pcm, const uint64_t *offsets,
                                    const uint64_t *n_samples, uint32_t n_songs,
                                    uint16_t features_version, float *d_out, int32_t *status,
                                    void *cuda_stream);

/* ---- multi-GPU: fused feature-row exchange ----------------------------------------- *
 * The reference analyses a library on `cores` threads and then ranks ALL songs against a seed
 * (src/song/decoder.rs:291-355 -> src/playlist.rs:256-326).  With songs sharded over the GPUs
 * of one box (one process per GPU) every rank needs all feature rows before its block of the
 * distance matrix.  Instead of a collective after the analysis, the last kernel of the analysis
 * stores each finished row straight into EVERY rank's gather buffer (peer-mapped over
 * NVLink/NVSwitch); what remains of the all-gather is one single-warp epoch barrier.
 *
 *   create (each rank)  -> exchange the BLISS_B200_GATHER_HANDLE_BYTES handles by any means
 *   connect (each rank, all handles in rank order)
 *   per step: scatter (1..n calls; global row of local song i = row_offset + i * row_stride)
 *             commit  (barrier on `cuda_stream`; *d_rows = this rank's complete row buffer)
 *
 * The row buffer is double-buffered by epoch: work that reads *d_rows must be enqueued on the
 * same stream before the NEXT commit.  All ranks must issue the same sequence of commits.
 * Rows no rank wrote keep their previous contents. */
typedef struct bliss_b200_gather bliss_b200_gather;
#define BLISS_B200_GATHER_HANDLE_BYTES 128
int bliss_b200_gather_create(uint32_t world, uint32_t rank, uint64_t max_rows, void *handle_out,
                             bliss_b200_gather **out);
int bliss_b200_gather_connect(bliss_b200_gather *g, const void *all_handles);
/* bliss_b200_analyze_batch_device whose rows also land in every rank's gather buffer;
 * d_out_local (n_songs x feature_count, device) may be NULL. */
int bliss_b200_analyze_batch_device_scatter(bliss_b200_gather *g, const float *d_pcm,
                                            const uint64_t *offsets, const uint64_t *n_samples,
                                            uint32_t n_songs, uint16_t features_version,
                                            uint64_t row_offset, uint64_t row_stride,
                                            float *d_out_local, int32_t *status, void *cuda_stream);
int bliss_b200_gather_commit(bliss_b200_gather *g, void *cuda_stream, const float **d_rows);
/* Synchronises the device; E_TIMEOUT if a barrier gave up on a peer (default 30 s). */
int bliss_b200_gather_check(bliss_b200_gather *g);
int bliss_b200_gather_set_timeout(bliss_b200_gather *g, uint64_t milliseconds);
int bliss_b200_gather_destroy(bliss_b200_gather *g);

/* ---- distances, src/playlist.rs ---------------------------------------------------- */
#define BLISS_B200_METRIC_MAHALANOBIS 0 /* sqrt((a-b)^T M (a-b)), playlist.rs:140-142; M NULL = identity = euclidean_distance :65-71 */
#define BLISS_B200_METRIC_COSINE 2      /* playlist.rs:76-79 */

/* Weight matrix of FeaturesVersion::feature_weights (src/lib.rs:168-173, :209-234): dim x dim */
int bliss_b200_feature_weights(uint16_t features_version, float *m);

/* One pair (Analysis::distance / Song::distance, src/song/mod.rs:364-370, :519-521 use the
 * Mahalanobis metric with feature_weights).  Host pointers.  m: dim x dim row-major or NULL. */
int bliss_b200_distance(const float *a, const float *b, uint32_t dim, int metric, const float *m,
                        float *out);

/* All pairs: out[i*n_cols + j] = d(rows[i], cols[j]).  Host pointers. */
int bliss_b200_distance_matrix(const float *rows, uint32_t n_rows, const float *cols, uint32_t n_cols,
                               uint32_t dim, int metric, const float *m, float *out);
/* Device pointers for rows / cols / out (m stays a HOST pointer); enqueued on cuda_stream. */
int bliss_b200_distance_matrix_device(const float *d_rows, uint32_t n_rows, const float *d_cols,
                                      uint32_t n_cols, uint32_t dim, int metric, const float *m,
                                      float *d_out, void *cuda_stream);

/* closest_to_songs (src/playlist.rs:256-270): order[] = candidate indices sorted (stably) by the
 * sum of distances to the seeds (FunctionDistanceMetric::distance, :56-58); keys (optional,
 * n_cands floats) receives those sums.  Host pointers. */
int bliss_b200_closest_to_songs(const float *seeds, uint32_t n_seeds, const float *cands,
                                uint32_t n_cands, uint32_t dim, int metric, const float *m,
                                uint32_t *order, float *keys);
/* song_to_song (src/playlist.rs:272-326): greedy nearest-neighbour chain from the seeds. */
int bliss_b200_song_to_song(const float *seeds, uint32_t n_seeds, const float *cands, uint32_t n_cands,
                            uint32_t dim, int metric, const float *m, uint32_t *order);

/* ---- STFT micro-benchmark (BASELINE.json config 3) ----------------------------------- */
/* 512-point hanningz phase-vocoder STFT, hop 256 (= PVocTempo, src/aubio.rs:338-425): writes
 * 257 magnitudes per frame; frames of song i start at row frame_offsets_out[i] (host array,
 * n_songs+1 entries, filled by the call).  d_pcm / d_mags are device buffers. */
int bliss_b200_stft512_mag_device(const float *d_pcm, const uint64_t *offsets, const uint64_t *n_samples,
                                  uint32_t n_songs, float *d_mags, uint64_t *frame_offsets_out,
                                  void *cuda_stream);

/* ---- introspection used by tests / bench ------------------------------------------------ */
/* Intermediate results of ONE song (host PCM).  Any pointer may be NULL.  Capacities are the
 * caller's business: centroid/rolloff/flatness n_s = (n-512)/128+1 floats, flux/thresholded
 * n_t = (n-512)/256+1 floats, bpms n_t/16+16 floats, stft8192 n_c x 4097 floats (n_c =
 * ceil(n/2205)), chroma n_c x 12 doubles, interval_features 10 doubles. */
typedef struct {
    float *centroid, *rolloff, *flatness;
    float *flux, *thresholded, *bpms;
    uint32_t *n_bpms;
    float *loudness_chunks; /* ceil(n/1024) */
    uint32_t *zero_crossings;
    float *stft8192;
    uint64_t *n_peaks;   /* pip_track candidate count */
    double *tuning;
    double *chroma;
    double *interval_features;
    double *peak_pitches;   /* pip_track (src/chroma.rs:269-331): interpolated pitches [Hz] and magnitudes of the */
    double *peak_mags;      /* n_peaks candidates, in no particular order; capacity n_c x 714 each */
} bliss_b200_taps;
int bliss_b200_analyze_taps(const float *pcm, uint64_t n_samples, uint16_t features_version, float *out,
                            const bliss_b200_taps *taps);
/* One of the 100 chroma filterbanks the contraction multiplies with (chroma_filter, src/chroma.rs:197-267, for
 * tuning = (-50 + tuning_index) / 100, tuning_index 0..99): out[12][4097] f64, rows = chroma bins as the reference
 * returns them (pinned by data/chroma-filter.npy at 1e-9, src/chroma.rs:705-714). */
int bliss_b200_chroma_filter(int tuning_index, double *out);

/* Per-kernel device time (CUDA events on the launching stream) accumulated while profiling is
 * on.  Kernel ids: 0 pvoc512, 1 timedomain, 2 stft8192, 3 tuning, 4 chroma, 5 peakpick,
 * 6 beattrack, 7 finalize, 8 distance, 9 stft512_mags. */
#define BLISS_B200_N_KERNELS 10
int bliss_b200_set_profiling(int on);
/* Synchronises the device, then fills ms[k] / launches[k] and resets the accumulators. */
int bliss_b200_get_profile(double *ms, uint64_t *launches);
const char *bliss_b200_kernel_name(int kernel_id);
/* Total number of this library's kernels launched since init. */
uint64_t bliss_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif

// api.cu -- host side of the C ABI declared in include/bliss_b200.h: device context,
// wave planning (which songs share one set of scratch buffers), kernel sequencing,
// host<->device staging, profiling counters.  No CPU compute path exists: every
// result is produced by the kernels in spectral.cu / tempo.cu / chroma.cu /
// finalize.cu / distance.cu, and every entry point fails loudly without a GPU.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bliss_b200.h"
#include "common.cuh"

namespace bliss {
// launchers implemented in the other translation units
int launch_pvoc512(const float *, const SongDesc *, const unsigned int *, int, unsigned int, int, PvocTables,
                   float *, float *, float *, float *, int, cudaStream_t);
int launch_stft512_mags(const float *, const SongDesc *, const unsigned int *, int, unsigned int, int,
                        PvocTables, float *, int, cudaStream_t);
void configure_kernels_spectral();
void configure_kernels_chroma();
void configure_kernels_tempo();
void configure_kernels_finalize();
int launch_timedomain(const float *, const SongDesc *, const unsigned int *, int, unsigned int, float *,
                      float *, unsigned int *, cudaStream_t);
int launch_peakpick(const float *, const SongDesc *, const unsigned int *, int, unsigned int, float *,
                    cudaStream_t);
int launch_beattrack(const float *, const float *, const SongDesc *, int, float *, float *, unsigned int *, int,
                     cudaStream_t);
int launch_chroma_filter_table(double *, float *, cudaStream_t);
int launch_stft8192(const float *, const SongDesc *, const unsigned int *, int, unsigned int, const float *,
                    const cpx *, const cpx *, const cpx *, float *, double *, double *, unsigned int *, int,
                    cudaStream_t);
int launch_tuning(const double *, const double *, const unsigned int *, const SongDesc *, int, int *, int,
                  cudaStream_t);
int launch_chroma(const float *, const SongDesc *, const unsigned int *, int, unsigned int, const float *,
                  const int *, double *, double *, int, cudaStream_t);
int launch_finalize(const SongDesc *, int, const float *, const float *, const float *, const float *,
                    const unsigned int *, const float *, const double *, int, float *, unsigned int,
                    const PeerRows &, cudaStream_t);
int launch_wave_setup(const void *, void *, size_t, unsigned int *, unsigned int *, unsigned int, cudaStream_t);
int launch_s16_to_f32(const short *, float *, size_t, cudaStream_t);
int launch_pcm_to_mono(const void *, float *, size_t, int, unsigned int, cudaStream_t);
ResamplePlan resample_plan(unsigned int up, unsigned int down, unsigned int taps4_needed, int variant);
int launch_resample(const float *, float *, const void *, const unsigned int *, unsigned int, unsigned int, const float *,
                    unsigned int, unsigned int, unsigned int, const ResamplePlan &, cudaStream_t);
int launch_gather_barrier(unsigned int *const *, int, int, unsigned int, unsigned long long, cudaStream_t);
int launch_distance_matrix(const float *, unsigned int, const float *, unsigned int, int, int, const float *,
                           float *, cudaStream_t, unsigned int ones_mask, int variant);
int launch_seed_distance(const float *, unsigned int, const float *, unsigned int, int, int, const float *,
                         float *, cudaStream_t);
size_t sort_temp_bytes(unsigned int);
int launch_stable_argsort(const float *, unsigned int, unsigned long long *, unsigned long long *, void *,
                          size_t, unsigned int *, cudaStream_t);
int launch_nearest_alive(const float *, unsigned int, const float *, unsigned int, int, int, const float *,
                         unsigned char *, unsigned int *, unsigned int, float *, cudaStream_t);
}  // namespace bliss

using namespace bliss;

namespace {

thread_local std::string g_last_error;

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return BLISS_B200_E_CUDA;                                                         \
        }                                                                                     \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            cudaError_t e = cudaFree(p);
            p = nullptr;
            cap = 0;
            if (e != cudaSuccess) return e;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

enum { K_PVOC = 0, K_TIME, K_STFT8K, K_TUNING, K_CHROMA, K_PEAK, K_BEAT, K_FINAL, K_DIST, K_STFT512 };
const char *const kKernelNames[BLISS_B200_N_KERNELS] = {
    "pvoc512_kernel", "timedomain_kernel", "stft8192_kernel", "tuning_kernel", "chroma_kernel",
    "peakpick_kernel", "beattrack_kernel", "finalize_kernel", "distance_kernels", "pvoc512_kernel<mags>"};

constexpr int N_STAGE = 2;  // descriptor staging slots per wave set
constexpr int N_SETS = 3;   // waves in flight

// Everything one wave owns: scratch arrays, its two streams (tempo/timbral chain and chroma chain), the
// descriptor staging ring.  N_SETS of them exist so that the latency-bound kernels of one wave
// (tuning, beat tracker: one CTA per song, ~3 ms whatever the wave size) run under the FFT kernels of
// the next waves instead of leaving the SMs idle.
struct WaveSet {
    cudaStream_t main = nullptr, side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
    bool used_in_call = false;
    DevBuf blob;  // SongDesc + prefix arrays (N_STAGE regions)
    DevBuf mags, cand_mag, cand_pitch, cand_count, cent, roll, flat, flux, thr, loud, eb, zcr, tempo, bpm,
        bpm_count, tuning, tiles, chroma_dbg;
    // pinned staging ring for descriptor uploads
    void *stage[N_STAGE] = {nullptr, nullptr};
    size_t stage_cap[N_STAGE] = {0, 0};
    cudaEvent_t stage_ev[N_STAGE] = {nullptr, nullptr};
    bool stage_used[N_STAGE] = {false, false};
    int stage_next = 0;
    void release() {
        DevBuf *all[] = {&blob, &mags, &cand_mag, &cand_pitch, &cand_count, &cent, &roll, &flat, &flux, &thr, &loud,
                         &eb, &zcr, &tempo, &bpm, &bpm_count, &tuning, &tiles, &chroma_dbg};
        for (DevBuf *b : all) b->release();
        for (int i = 0; i < N_STAGE; i++) {
            if (stage[i]) cudaFreeHost(stage[i]);
            stage[i] = nullptr;
            stage_cap[i] = 0;
            if (stage_ev[i]) cudaEventDestroy(stage_ev[i]);
            stage_ev[i] = nullptr;
            stage_used[i] = false;
        }
    }
};

struct Ctx {
    std::mutex mu;
    bool inited = false;
    int device = -1;
    cudaStream_t stream = nullptr, copy_stream = nullptr, conv_stream = nullptr;
    cudaEvent_t ev_begin = nullptr;
    size_t ws_limit = 0;
    int variant = 0;  // BLISS_B200_VARIANT, see common.cuh
    int launch_order = 0;  // BLISS_B200_ORDER: which chain of a wave is enqueued first (run_wave)
    unsigned int metric_ones = 0;  // bit i: diagonal weight i of the metric last prepared is exactly 1 (prepare_metric)
    // constant tables
    DevBuf t_win512, t_twA, t_hann8k, t_tw4k, t_tw2, t_tw8k, t_filt, t_filt32;
    WaveSet ws[N_SETS];
    int next_set = 0;
    // host-API staging
    DevBuf pcm[4], raw16[4], feats, metric, misc[6];
    // sample-rate conversion (design_resampler / enqueue_resample): the filter of the rate last used, the mono input of
    // each ring slot at its own rate, and each slot's job table (pinned host copy + device copy)
    DevBuf rs_tab, rs_in[4], rs_jobs[4];
    void *rs_host[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t rs_host_cap[4] = {0, 0, 0, 0};
    uint32_t rs_rate = 0, rs_up = 0, rs_down = 0, rs_taps4 = 0, rs_pre = 0;
    // profiling
    bool profiling = false;
    struct EvPair { cudaEvent_t a, b; int kid; };
    std::vector<EvPair> ev_pending;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[BLISS_B200_N_KERNELS] = {0};
    unsigned long long prof_launches[BLISS_B200_N_KERNELS] = {0};
    std::atomic<unsigned long long> launches{0};
};

// One context per device.  Slot 0 is the primary context (bliss_b200_init); bliss_b200_init_devices fills slots
// 0 .. n-1 with devices 0 .. n-1, and the host-buffer entry points then shard one call's songs over all of them from
// ONE process (the reference is one process: src/song/decoder.rs:282-331), one worker thread per device.  Every
// function below reaches "its" context through `g`: the calling thread's current context (thread-local; the primary
// one unless a multi-device dispatcher set another).
constexpr int MAX_DEVICES = 16;
Ctx g_ctx[MAX_DEVICES];
std::atomic<int> g_n_ctx{0};          // initialised contexts (slots 0 .. n-1)
std::mutex g_init_mu;                 // serialises init / shutdown of the context table
thread_local Ctx *g_cur = &g_ctx[0];
#define g (*g_cur)

struct ProfScope {
    Ctx::EvPair pr{nullptr, nullptr, -1};
    cudaStream_t st;
    bool on;
    ProfScope(int kid, cudaStream_t s) : st(s), on(g.profiling) {
        if (!on) return;
        auto get = []() {
            cudaEvent_t e;
            if (!g.ev_pool.empty()) { e = g.ev_pool.back(); g.ev_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        pr.a = get();
        pr.b = get();
        pr.kid = kid;
        cudaEventRecord(pr.a, st);
    }
    void done(int n_launched) {
        if (n_launched > 0) g.launches += (unsigned long long)n_launched;
        if (!on) return;
        cudaEventRecord(pr.b, st);
        g.prof_launches[pr.kid] += (unsigned long long)(n_launched > 0 ? n_launched : 0);
        g.ev_pending.push_back(pr);
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- per-song geometry --------------------------------------
struct SongGeom {
    uint32_t n_s, n_t, n_c, n_c_comp, n_l, n_eb, n_pairs8k, n_tiles, bpm_cap;
    size_t scratch_bytes;
};

SongGeom geom_of(uint64_t n) {
    SongGeom q{};
    if (n < (uint64_t)MIN_SAMPLES) return q;
    q.n_s = (uint32_t)((n - 512) / 128 + 1);
    q.n_t = (uint32_t)((n - 512) / 256 + 1);
    q.n_c = (uint32_t)ceilf((float)n / (float)CH_HOP);  // utils.rs:30, computed in f32 like the reference
    const uint64_t windows = n / CH_HOP + 1;            // windows(8192).step_by(2205) over n + 8192 padded samples
    q.n_c_comp = (uint32_t)std::min<uint64_t>(q.n_c, windows);
    q.n_l = (uint32_t)((n + 1023) / 1024);
    q.n_eb = (uint32_t)(n / 256);
    q.n_pairs8k = (q.n_c_comp + 3) / 4;  // stft8192_kernel: one CTA per 4 consecutive chroma frames
    q.n_tiles = (q.n_c + CH_TILE_FRAMES - 1) / CH_TILE_FRAMES;
    q.bpm_cap = q.n_t / 16 + 16;
    const size_t rows = (size_t)q.n_c_comp;
    q.scratch_bytes = rows * CH_STRIDE * 4 + rows * CH_MAX_PEAKS * 16 + (size_t)q.n_s * 12 + (size_t)q.n_t * 8 +
                      (size_t)q.n_l * 4 + (size_t)q.n_eb * 4 + (size_t)q.n_tiles * 80 + (size_t)q.bpm_cap * 4 + 256;
    return q;
}

struct WavePlan {
    std::vector<SongDesc> sd;
    std::vector<uint32_t> k1_prefix, chunk_prefix, t_prefix, pair_prefix, tile_prefix;
    uint32_t pairs_per_item = 64;
    size_t rows = 0, cands = 0, n_s = 0, n_t = 0, n_l = 0, n_eb = 0, tiles = 0, bpms = 0;
};

void plan_wave(const uint64_t *offsets, const uint64_t *n_samples, uint32_t first, uint32_t count,
               bool only_stft512, WavePlan &w) {
    w = WavePlan();
    w.sd.resize(count);
    uint64_t total_pairs = 0;
    for (uint32_t i = 0; i < count; i++) total_pairs += geom_of(n_samples[first + i]).n_t;
    // enough warp-items to fill 148 SMs x 16 warps a few times over, runs as long as possible
    uint64_t r = total_pairs / (148ull * 16ull * 4ull);
    w.pairs_per_item = (uint32_t)std::min<uint64_t>(256, std::max<uint64_t>(8, r));  // (256 against 64: -1 % on 1024 tracks, profiles/knobs_r02.md)
    if (const char *e = getenv("BLISS_B200_PVOC_PAIRS")) w.pairs_per_item = (uint32_t)std::max(4, atoi(e));  // experiments
    w.k1_prefix.assign(count + 1, 0);
    w.chunk_prefix.assign(count + 1, 0);
    w.t_prefix.assign(count + 1, 0);
    w.pair_prefix.assign(count + 1, 0);
    w.tile_prefix.assign(count + 1, 0);
    for (uint32_t i = 0; i < count; i++) {
        const uint64_t n = n_samples[first + i];
        const SongGeom q = geom_of(n);
        SongDesc &d = w.sd[i];
        memset(&d, 0, sizeof(d));
        d.pcm_off = offsets[first + i];
        d.n = (unsigned int)n;
        d.valid = (n >= (uint64_t)MIN_SAMPLES && n < (1ull << 31)) ? 1u : 0u;
        if (d.valid) {
            d.n_s = q.n_s; d.n_t = q.n_t; d.n_c = q.n_c; d.n_c_comp = q.n_c_comp; d.n_l = q.n_l;
        }
        d.mag_off = w.rows;
        d.cand_off = w.cands;
        d.s_off = (unsigned int)w.n_s;
        d.t_off = (unsigned int)w.n_t;
        d.l_off = (unsigned int)w.n_l;
        d.e_off = (unsigned int)w.n_eb;
        d.c_tile_off = (unsigned int)w.tiles;
        d.bpm_off = (unsigned int)w.bpms;
        const uint32_t items = d.valid ? (q.n_t + w.pairs_per_item - 1) / w.pairs_per_item : 0;
        w.k1_prefix[i + 1] = w.k1_prefix[i] + items;
        w.chunk_prefix[i + 1] = w.chunk_prefix[i] + (d.valid ? (q.n_l + 7) / 8 : 0);  // groups of 8 chunks (timedomain_kernel)
        w.t_prefix[i + 1] = w.t_prefix[i] + (d.valid ? q.n_t : 0);
        w.pair_prefix[i + 1] = w.pair_prefix[i] + (d.valid ? q.n_pairs8k : 0);
        w.tile_prefix[i + 1] = w.tile_prefix[i] + (d.valid ? q.n_tiles : 0);
        if (d.valid) {
            w.n_t += q.n_t;
            if (!only_stft512) {
                w.rows += (size_t)q.n_c_comp;
                w.cands += (size_t)q.n_c_comp * CH_MAX_PEAKS;
                w.n_s += q.n_s;
                w.n_l += q.n_l;
                w.n_eb += q.n_eb;
                w.tiles += q.n_tiles;
                w.bpms += q.bpm_cap;
            }
        }
    }
}

// upload descriptors + prefix arrays through the pinned ring; returns device pointers
struct WaveDev {
    const SongDesc *sd;
    const unsigned int *k1_prefix, *chunk_prefix, *t_prefix, *pair_prefix, *tile_prefix;
};

int upload_plan(const WavePlan &w, WaveSet &S, cudaStream_t st, WaveDev &out, unsigned int *zcr_count,
                unsigned int *cand_count) {
    const size_t n = w.sd.size();
    const size_t sd_bytes = align_up(n * sizeof(SongDesc), 256);
    const size_t pf_bytes = align_up((n + 1) * sizeof(uint32_t), 256);
    const size_t total = sd_bytes + 5 * pf_bytes;
    CK(S.blob.ensure(total * N_STAGE));
    const int slot = S.stage_next;
    S.stage_next = (S.stage_next + 1) % N_STAGE;
    if (S.stage_used[slot]) CK(cudaEventSynchronize(S.stage_ev[slot]));
    if (S.stage_cap[slot] < total) {
        if (S.stage[slot]) cudaFreeHost(S.stage[slot]);
        S.stage[slot] = nullptr;
        CK(cudaMallocHost(&S.stage[slot], total + total / 4));
        S.stage_cap[slot] = total + total / 4;
    }
    if (!S.stage_ev[slot]) CK(cudaEventCreateWithFlags(&S.stage_ev[slot], cudaEventDisableTiming));
    char *h = (char *)S.stage[slot];
    memcpy(h, w.sd.data(), n * sizeof(SongDesc));
    const std::vector<uint32_t> *pf[5] = {&w.k1_prefix, &w.chunk_prefix, &w.t_prefix, &w.pair_prefix, &w.tile_prefix};
    for (int i = 0; i < 5; i++) memcpy(h + sd_bytes + i * pf_bytes, pf[i]->data(), (n + 1) * sizeof(uint32_t));
    // each ring slot owns its own region of the device blob, so a wave still executing
    // never sees the next wave's descriptors
    char *d = S.blob.as<char>() + (size_t)slot * (S.blob.cap / N_STAGE / 256 * 256);
    if ((size_t)(S.blob.cap / N_STAGE / 256 * 256) < total) { g_last_error = "descriptor blob too small"; return BLISS_B200_E_CUDA; }
    // fetched by a kernel (wave_setup.cu), not by the copy engine, which may be busy with PCM for a long time
    g.launches += (unsigned long long)launch_wave_setup(h, d, total, zcr_count, cand_count,
                                                        zcr_count ? (unsigned int)n : 0u, st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(S.stage_ev[slot], st));
    S.stage_used[slot] = true;
    out.sd = reinterpret_cast<const SongDesc *>(d);
    const unsigned int *p0 = reinterpret_cast<const unsigned int *>(d + sd_bytes);
    out.k1_prefix = p0;
    out.chunk_prefix = reinterpret_cast<const unsigned int *>(d + sd_bytes + 1 * pf_bytes);
    out.t_prefix = reinterpret_cast<const unsigned int *>(d + sd_bytes + 2 * pf_bytes);
    out.pair_prefix = reinterpret_cast<const unsigned int *>(d + sd_bytes + 3 * pf_bytes);
    out.tile_prefix = reinterpret_cast<const unsigned int *>(d + sd_bytes + 4 * pf_bytes);
    return BLISS_B200_OK;
}

PvocTables pvoc_tables() { return PvocTables{g.t_win512.as<float>(), g.t_twA.as<cpx>()}; }

// one wave of the full analysis: every kernel of the path, enqueued on `st` (tempo/timbral chain) and
// `sb` (chroma chain; may equal st), using the scratch of wave set S
int run_wave(const float *d_pcm, const WavePlan &w, int version, float *d_out, uint32_t out_base, WaveSet &S,
             cudaStream_t st, cudaStream_t sb, bool debug, const PeerRows &peers) {
    const int n = (int)w.sd.size();
    if (n == 0) return BLISS_B200_OK;
    // + one tile of rows: chroma_pipe_kernel forms (never dereferences) addresses of a full tile
    CK(S.mags.ensure((std::max<size_t>(w.rows, 2) + CH_TILE_FRAMES) * CH_STRIDE * sizeof(float)));
    CK(S.cand_mag.ensure(std::max<size_t>(w.cands, 1) * sizeof(double)));
    CK(S.cand_pitch.ensure(std::max<size_t>(w.cands, 1) * sizeof(double)));  // interpolated pitches (f64)
    CK(S.cand_count.ensure((size_t)n * 4));
    CK(S.cent.ensure(std::max<size_t>(w.n_s, 1) * 4));
    CK(S.roll.ensure(std::max<size_t>(w.n_s, 1) * 4));
    CK(S.flat.ensure(std::max<size_t>(w.n_s, 1) * 4));
    CK(S.flux.ensure(std::max<size_t>(w.n_t, 1) * 4));
    CK(S.thr.ensure(std::max<size_t>(w.n_t, 1) * 4));
    CK(S.loud.ensure(std::max<size_t>(w.n_l, 1) * 4));
    CK(S.eb.ensure(std::max<size_t>(w.n_eb, 1) * 4));
    CK(S.zcr.ensure((size_t)n * 4));
    CK(S.tempo.ensure((size_t)n * 4));
    CK(S.bpm.ensure(std::max<size_t>(w.bpms, 1) * 4));
    CK(S.bpm_count.ensure((size_t)n * 4));
    CK(S.tuning.ensure((size_t)n * 4));
    CK(S.tiles.ensure(std::max<size_t>(w.tiles, 1) * 10 * sizeof(double)));
    if (debug) CK(S.chroma_dbg.ensure(std::max<size_t>(w.tiles, 1) * CH_TILE_FRAMES * 12 * sizeof(double)));
    WaveDev dv;
    int rc = upload_plan(w, S, st, dv, S.zcr.as<unsigned int>(), S.cand_count.as<unsigned int>());  // also zeroes both
    if (rc) return rc;

    // Two independent chains per wave: the tempo/timbral chain stays on the caller's stream, the chroma
    // chain runs on a side stream so that its latency-bound kernels (tuning) overlap the other chain's
    // compute-bound ones and vice versa (beat tracker under the chroma STFT).  Joined before finalize.
    // (while per-kernel profiling is on, the caller passes sb == st: both chains serialised so that each
    // kernel's CUDA-event duration is its own and not inflated by the kernel it would overlap with)
    CK(cudaEventRecord(S.ev_fork, st));
    CK(cudaStreamWaitEvent(sb, S.ev_fork, 0));
    // The order in which the two chains are ENQUEUED decides which one the block scheduler serves first: each FFT kernel
    // fills every SM on its own (shared memory / registers), so the chains overlap only where one of them runs a
    // latency-bound kernel.  g.launch_order (BLISS_B200_ORDER): 0 = chroma STFT first, then the tempo / timbral chain
    // interleaved (round 1), 1 = the whole tempo / timbral chain first (the beat tracker then runs under the chroma
    // STFT), 2 = the whole chroma chain first.
    auto k_stft = [&]() -> int {
        ProfScope p(K_STFT8K, sb);
        const int nl = launch_stft8192(d_pcm, dv.sd, dv.pair_prefix, n, w.pair_prefix[n], g.t_hann8k.as<float>(),
                                       g.t_tw4k.as<cpx>(), g.t_tw2.as<cpx>(), g.t_tw8k.as<cpx>(), S.mags.as<float>(), S.cand_mag.as<double>(),
                                       S.cand_pitch.as<double>(), S.cand_count.as<unsigned int>(), g.variant, sb);
        p.done(nl);
        if (nl < 0) {
            g_last_error = "cudaFuncSetAttribute(stft8192v2_kernel, MaxDynamicSharedMemorySize) failed";
            return BLISS_B200_E_CUDA;
        }
        return 0;
    };
    auto k_time_pvoc = [&]() -> int {
        { ProfScope p(K_TIME, st);
          p.done(launch_timedomain(d_pcm, dv.sd, dv.chunk_prefix, n, w.chunk_prefix[n], S.loud.as<float>(),
                                   S.eb.as<float>(), S.zcr.as<unsigned int>(), st)); }
        ProfScope p(K_PVOC, st);
        const int nl = launch_pvoc512(d_pcm, dv.sd, dv.k1_prefix, n, w.k1_prefix[n], (int)w.pairs_per_item, pvoc_tables(),
                                      S.cent.as<float>(), S.roll.as<float>(), S.flat.as<float>(), S.flux.as<float>(), g.variant, st);
        p.done(nl);
        if (nl < 0) {
            g_last_error = "cudaFuncSetAttribute(pvoc512v2_kernel, MaxDynamicSharedMemorySize) failed";
            return BLISS_B200_E_CUDA;
        }
        return 0;
    };
    auto k_tuning = [&]() {
        ProfScope p(K_TUNING, sb);
        p.done(launch_tuning(S.cand_mag.as<double>(), S.cand_pitch.as<double>(),
                             S.cand_count.as<unsigned int>(), dv.sd, n, S.tuning.as<int>(), g.variant, sb));
    };
    auto k_peak = [&]() {
        ProfScope p(K_PEAK, st);
        p.done(launch_peakpick(S.flux.as<float>(), dv.sd, dv.t_prefix, n, w.t_prefix[n], S.thr.as<float>(), st));
    };
    auto k_chroma = [&]() -> int {
        ProfScope p(K_CHROMA, sb);
        const int nl = launch_chroma(S.mags.as<float>(), dv.sd, dv.tile_prefix, n, w.tile_prefix[n], g.t_filt32.as<float>(),
                                     S.tuning.as<int>(), S.tiles.as<double>(), debug ? S.chroma_dbg.as<double>() : nullptr,
                                     g.variant, sb);
        p.done(nl);
        if (nl < 0) {  // the opt-in to > 48 KB of dynamic shared memory was refused
            g_last_error = "cudaFuncSetAttribute(chroma_pipe_kernel, MaxDynamicSharedMemorySize) failed";
            return BLISS_B200_E_CUDA;
        }
        return 0;
    };
    auto k_beat = [&]() {
        ProfScope p(K_BEAT, st);
        p.done(launch_beattrack(S.thr.as<float>(), S.eb.as<float>(), dv.sd, n, S.bpm.as<float>(),
                                S.tempo.as<float>(), S.bpm_count.as<unsigned int>(), g.variant, st));
    };
    if (g.launch_order == 1) {
        if (k_time_pvoc()) return BLISS_B200_E_CUDA;
        k_peak();
        k_beat();
        if (k_stft()) return BLISS_B200_E_CUDA;
        k_tuning();
        if (k_chroma()) return BLISS_B200_E_CUDA;
    } else if (g.launch_order == 2) {
        if (k_stft()) return BLISS_B200_E_CUDA;
        k_tuning();
        if (k_chroma()) return BLISS_B200_E_CUDA;
        if (k_time_pvoc()) return BLISS_B200_E_CUDA;
        k_peak();
        k_beat();
    } else {
        if (k_stft()) return BLISS_B200_E_CUDA;
        if (k_time_pvoc()) return BLISS_B200_E_CUDA;
        k_tuning();
        k_peak();
        if (k_chroma()) return BLISS_B200_E_CUDA;
        k_beat();
    }
    CK(cudaEventRecord(S.ev_join, sb));
    CK(cudaStreamWaitEvent(st, S.ev_join, 0));
    { ProfScope p(K_FINAL, st);
      p.done(launch_finalize(dv.sd, n, S.cent.as<float>(), S.roll.as<float>(), S.flat.as<float>(),
                             S.loud.as<float>(), S.zcr.as<unsigned int>(), S.tempo.as<float>(),
                             S.tiles.as<double>(), version, d_out, out_base, peers, st)); }
    CK(cudaGetLastError());
    return BLISS_B200_OK;
}

// Split [0, n_songs) into waves and keep up to N_SETS of them in flight.
//
// wait_ev: what the waves wait for before touching d_pcm (nullptr: everything enqueued on `st` so far).
// done_ev: nullptr -> `st` is made to wait for every wave before the call returns (the results are
//          ordered on `st` like a plain kernel launch); otherwise done_ev[s] is recorded on wave set s's
//          stream behind this call's last wave there (done_used[s] says whether it was) and `st` is left
//          alone -- the host path uses this to let the next chunk's waves start while these still run.
int analyze_device_locked(const float *d_pcm, const uint64_t *offsets, const uint64_t *n_samples,
                          uint32_t n_songs, int version, float *d_out, int32_t *status, cudaStream_t st,
                          bool debug, const PeerRows *peers = nullptr, cudaEvent_t wait_ev = nullptr,
                          cudaEvent_t *done_ev = nullptr, bool *done_used = nullptr) {
    PeerRows no_peers;
    memset(&no_peers, 0, sizeof(no_peers));
    if (((uintptr_t)d_pcm & 15u) != 0) { g_last_error = "d_pcm must be 16-byte aligned"; return BLISS_B200_E_ARG; }
    // serial mode (per-kernel profiling, debug taps): every wave on set 0, both chains on `st`
    const bool serial = g.profiling || debug;
    // Wave size: as many songs as the workspace share of one set holds.  Splitting a resident batch into
    // more, overlapping waves does NOT pay: measured on 1024 tracks, 1 wave 79.8 ms, 3 waves 80.1, 8 waves
    // 81.5, 16 waves 81.5 -- the FFT kernels keep every SM busy either way and the one-CTA-per-song
    // kernels already overlap the other chain.  The sets earn their keep on the host path, where
    // chunks arrive over PCIe one after the other.  BLISS_B200_WAVE_SONGS caps the wave for experiments.
    const size_t wave_bytes = serial ? g.ws_limit : g.ws_limit / N_SETS;
    uint32_t wave_songs = n_songs;
    if (const char *e = getenv("BLISS_B200_WAVE_SONGS")) wave_songs = std::max(1, atoi(e));
    for (auto &S : g.ws) S.used_in_call = false;
    if (!serial && !wait_ev) {
        CK(cudaEventRecord(g.ev_begin, st));
        wait_ev = g.ev_begin;
    }
    uint32_t first = 0;
    WavePlan w;
    while (first < n_songs) {
        size_t bytes = 0;
        uint32_t count = 0;
        while (first + count < n_songs && count < wave_songs) {
            const size_t b = geom_of(n_samples[first + count]).scratch_bytes + 128;
            if (count > 0 && bytes + b > wave_bytes) break;
            bytes += b;
            count++;
        }
        plan_wave(offsets, n_samples, first, count, false, w);
        int rc;
        if (serial) {
            WaveSet &S = g.ws[0];  // set 0 on the caller's stream: order it behind and ahead of set 0's own waves
            CK(cudaEventRecord(S.ev_done, S.main));
            CK(cudaStreamWaitEvent(st, S.ev_done, 0));
            if (wait_ev) CK(cudaStreamWaitEvent(st, wait_ev, 0));
            rc = run_wave(d_pcm, w, version, d_out, first, S, st, st, debug, peers ? *peers : no_peers);
            if (rc == 0) {
                CK(cudaEventRecord(S.ev_done, st));
                CK(cudaStreamWaitEvent(S.main, S.ev_done, 0));
            }
        } else {
            WaveSet &S = g.ws[g.next_set];
            g.next_set = (g.next_set + 1) % N_SETS;
            CK(cudaStreamWaitEvent(S.main, wait_ev, 0));
            rc = run_wave(d_pcm, w, version, d_out, first, S, S.main, S.side, debug, peers ? *peers : no_peers);
            S.used_in_call = true;
        }
        if (rc) return rc;
        first += count;
    }
    if (!serial) {
        for (int i = 0; i < N_SETS; i++) {
            WaveSet &S = g.ws[i];
            if (done_used) done_used[i] = S.used_in_call;
            if (!S.used_in_call) continue;
            if (done_ev) {
                CK(cudaEventRecord(done_ev[i], S.main));
            } else {
                CK(cudaEventRecord(S.ev_done, S.main));
                CK(cudaStreamWaitEvent(st, S.ev_done, 0));
            }
        }
    } else if (done_ev) {  // serial mode inside the host path: everything is on `st`
        if (done_used) for (int i = 0; i < N_SETS; i++) done_used[i] = (i == 0);
        CK(cudaEventRecord(done_ev[0], st));
    }
    if (status)
        for (uint32_t i = 0; i < n_songs; i++)
            status[i] = n_samples[i] < (uint64_t)MIN_SAMPLES ? BLISS_B200_SONG_TOO_SHORT
                        : n_samples[i] >= (1ull << 31)        ? BLISS_B200_SONG_INTERNAL
                                                              : BLISS_B200_SONG_OK;
    return BLISS_B200_OK;
}

int check_version(uint16_t v) {
    if (v != 1 && v != 2) { g_last_error = "features_version must be 1 or 2"; return BLISS_B200_E_ARG; }
    return 0;
}

int build_tables() {
    // hanningz of PVoc::new / PVocTempo::new (aubio.rs:150-154), f32 math as the reference
    std::vector<float> win(512);
    const float PI_F = 3.14159274101257324f;
    for (int i = 0; i < 512; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * PI_F * (float)i / 512.f));
    std::vector<cpx> twA(16 * 32);
    for (int k1 = 0; k1 < 16; k1++)
        for (int l = 0; l < 32; l++) {
            const double a = -2.0 * M_PI * (double)(k1 * l) / 512.0;
            twA[k1 * 32 + l] = cpx{(float)cos(a), (float)sin(a)};
        }
    // periodic Hann of utils::stft (utils.rs:36-38)
    std::vector<float> hann(2 * (8192 + 4 * 256));
    for (int i = 0; i < 8192; i++) hann[i] = 0.5f - 0.5f * cosf(2.f * (float)i * PI_F / 8192.f);
    // behind the window: per-thread phase (cos e, cos o, sin e, sin o) of samples 2 tid, 2 tid + 1 for the
    // synthesised-window variant of stft8192_kernel (VARIANT_WINSYN, rfft8192.cuh hann_pair)
    for (int t = 0; t < 256; t++) {
        const double te = 2.0 * M_PI * (double)(2 * t) / 8192.0, to = 2.0 * M_PI * (double)(2 * t + 1) / 8192.0;
        hann[8192 + 4 * t + 0] = (float)cos(te);
        hann[8192 + 4 * t + 1] = (float)cos(to);
        hann[8192 + 4 * t + 2] = (float)sin(te);
        hann[8192 + 4 * t + 3] = (float)sin(to);
    }
    // ... and both once more for the frame rotated by one sample (VARIANT_ODDSHIFT, chroma.cu): the window read as
    // hann[(m + 8191) % 8192], the phase of samples 2 tid - 1 and 2 tid
    for (int m = 0; m < 8192; m++) hann[9216 + m] = hann[(m + 8191) % 8192];
    for (int t = 0; t < 256; t++) {
        const double ta = 2.0 * M_PI * (double)(2 * t - 1) / 8192.0, tb = 2.0 * M_PI * (double)(2 * t) / 8192.0;
        hann[17408 + 4 * t + 0] = (float)cos(ta);
        hann[17408 + 4 * t + 1] = (float)cos(tb);
        hann[17408 + 4 * t + 2] = (float)sin(ta);
        hann[17408 + 4 * t + 3] = (float)sin(tb);
    }
    // stft8192v2_kernel (stft8192_v2.cuh): the Hann phase of a frame rotated by r = 0..3 samples, [r][thread < 128]
    // {cos phi_0..3, sin phi_0..3}, phi_j = 2 pi (4 thread + j - r) / 8192
    hann.resize(2 * (8192 + 4 * 256) + 4 * 128 * 8);
    for (int r = 0; r < 4; r++)
        for (int t = 0; t < 128; t++)
            for (int j = 0; j < 4; j++) {
                const double ph = 2.0 * M_PI * (double)(4 * t + j - r) / 8192.0;
                hann[2 * (8192 + 4 * 256) + (size_t)(r * 128 + t) * 8 + j] = (float)cos(ph);
                hann[2 * (8192 + 4 * 256) + (size_t)(r * 128 + t) * 8 + 4 + j] = (float)sin(ph);
            }
    // stft8192v3_kernel (stft8192_v3.cuh): one column per thread, frame rotated by r = 0..1 samples, [r][thread < 256]
    // {cos phi_0, cos phi_1, sin phi_0, sin phi_1}, phi_j = 2 pi (2 thread + j - r) / 8192
    {
        const size_t off3 = hann.size();
        hann.resize(off3 + 2 * 256 * 4);
        for (int r = 0; r < 2; r++)
            for (int t = 0; t < 256; t++)
                for (int j = 0; j < 2; j++) {
                    const double ph = 2.0 * M_PI * (double)(2 * t + j - r) / 8192.0;
                    hann[off3 + (size_t)(r * 256 + t) * 4 + j] = (float)cos(ph);
                    hann[off3 + (size_t)(r * 256 + t) * 4 + 2 + j] = (float)sin(ph);
                }
    }
    // pass-1 twiddles [k1][b] = W4096^(b k1), pass-2 twiddles [k2][j] = W256^(j k2), and W8192^t, t < 256
    // (real-FFT untangling): rfft8192.cuh
    std::vector<cpx> tw4(4096), tw2(256), tw(256);
    for (int k1 = 0; k1 < 16; k1++)
        for (int b = 0; b < 256; b++) {
            const double a = -2.0 * M_PI * (double)(b * k1) / 4096.0;
            tw4[k1 * 256 + b] = cpx{(float)cos(a), (float)sin(a)};
        }
    for (int k2 = 0; k2 < 16; k2++)
        for (int j = 0; j < 16; j++) {
            const double a = -2.0 * M_PI * (double)(j * k2) / 256.0;
            tw2[k2 * 16 + j] = cpx{(float)cos(a), (float)sin(a)};
        }
    for (int m = 0; m < 256; m++) {
        const double a = -2.0 * M_PI * (double)m / 8192.0;
        tw[m] = cpx{(float)cos(a), (float)sin(a)};
    }
    CK(g.t_win512.ensure(win.size() * 4));
    CK(g.t_twA.ensure(twA.size() * sizeof(cpx)));
    CK(g.t_hann8k.ensure(hann.size() * 4));
    CK(g.t_tw8k.ensure(tw.size() * sizeof(cpx)));
    CK(g.t_tw4k.ensure(tw4.size() * sizeof(cpx)));
    CK(cudaMemcpy(g.t_tw4k.p, tw4.data(), tw4.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(g.t_tw2.ensure(tw2.size() * sizeof(cpx)));
    CK(cudaMemcpy(g.t_tw2.p, tw2.data(), tw2.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(g.t_filt.ensure((size_t)100 * CH_BINS * 12 * sizeof(double)));
    CK(cudaMemcpy(g.t_win512.p, win.data(), win.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.t_twA.p, twA.data(), twA.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.t_hann8k.p, hann.data(), hann.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.t_tw8k.p, tw.data(), tw.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    CK(g.t_filt32.ensure((size_t)100 * CH_BINS * 12 * sizeof(float)));
    g.launches += launch_chroma_filter_table(g.t_filt.as<double>(), g.t_filt32.as<float>(), g.stream);
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    return BLISS_B200_OK;
}

// metric upload: diagonal matrices collapse to a weight vector (mode 0), else full (mode 1)
int prepare_metric(int metric, const float *m, uint32_t dim, cudaStream_t st, int &mode, const float *&d_w) {
    d_w = nullptr;
    if (metric == BLISS_B200_METRIC_COSINE) { mode = 2; return 0; }
    if (metric != BLISS_B200_METRIC_MAHALANOBIS) { g_last_error = "unknown metric"; return BLISS_B200_E_ARG; }
    mode = 0;
    g.metric_ones = 0xffffffffu;
    if (!m) return 0;
    bool diag = true;
    for (uint32_t i = 0; i < dim && diag; i++)
        for (uint32_t j = 0; j < dim; j++)
            if (i != j && m[(size_t)i * dim + j] != 0.f) { diag = false; break; }
    CK(g.metric.ensure((size_t)dim * dim * 4 + 256));
    if (diag) {
        std::vector<float> w(dim);
        g.metric_ones = 0;
        for (uint32_t i = 0; i < dim; i++) {
            w[i] = m[(size_t)i * dim + i];
            if (w[i] == 1.0f && i < 32) g.metric_ones |= 1u << i;
        }
        // pageable source: the runtime stages it before returning, so `w` may go out of scope
        CK(cudaMemcpyAsync(g.metric.p, w.data(), dim * 4, cudaMemcpyHostToDevice, st));
    } else {
        mode = 1;
        CK(cudaMemcpyAsync(g.metric.p, m, (size_t)dim * dim * 4, cudaMemcpyHostToDevice, st));
    }
    d_w = g.metric.as<float>();
    return 0;
}

}  // namespace

extern "C" {

const char *bliss_b200_strerror(int code) {
    switch (code) {
        case BLISS_B200_OK: return "ok";
        case BLISS_B200_SONG_TOO_SHORT: return "empty or too short song.";  // src/song/mod.rs:426-430
        case BLISS_B200_SONG_INTERNAL: return "internal analysis error";
        case BLISS_B200_E_CUDA: return "CUDA error";
        case BLISS_B200_E_ARG: return "invalid argument";
        case BLISS_B200_E_NOT_INIT: return "bliss_b200_init() not called";
        case BLISS_B200_E_NOMEM: return "workspace limit too small";
        case BLISS_B200_E_NO_DEVICE: return "no CUDA device (no CPU fallback exists)";
        case BLISS_B200_E_TIMEOUT: return "a peer never reached the gather barrier";
        case BLISS_B200_E_UNSUPPORTED: return "unsupported input (sample rate other than 22050 Hz)";
        default: return "unknown";
    }
}

const char *bliss_b200_last_error(void) { return g_last_error.c_str(); }

uint32_t bliss_b200_feature_count(uint16_t v) { return v == 2 ? 23u : v == 1 ? 20u : 0u; }

// initialise the calling thread's current context `g` on `device` (caller holds g.mu)
static int init_ctx_locked(int device) {
    if (g.inited) {
        if (g.device == device) return BLISS_B200_OK;
        g_last_error = "already initialised on another device";
        return BLISS_B200_E_ARG;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_last_error = "no CUDA device visible";
        return BLISS_B200_E_NO_DEVICE;
    }
    if (device < 0 || device >= count) { g_last_error = "device index out of range"; return BLISS_B200_E_ARG; }
    CK(cudaSetDevice(device));
    g.device = device;
    CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g.conv_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g.ev_begin, cudaEventDisableTiming));
    // BLISS_B200_STREAM_PRIORITY (experiment; measured flat in round 2, profiles/ab_r02.md; default 0 = both chains at
    // the same priority): 1 = the tempo / timbral chain's stream is preferred by the block scheduler, 2 = the chroma chain's
    int prio_main = 0, prio_side = 0;
    if (const char *e = getenv("BLISS_B200_STREAM_PRIORITY")) {
        int lo = 0, hi = 0;  // numerically lower = higher priority
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int mode = atoi(e);
        if (mode == 1) { prio_main = hi; prio_side = lo; }
        if (mode == 2) { prio_main = lo; prio_side = hi; }
    }
    for (auto &S : g.ws) {
        CK(cudaStreamCreateWithPriority(&S.main, cudaStreamNonBlocking, prio_main));
        CK(cudaStreamCreateWithPriority(&S.side, cudaStreamNonBlocking, prio_side));
        CK(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming));
    }
    g.next_set = 0;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    g.ws_limit = (size_t)((double)total_b * 0.40);
    g.variant = VARIANT_PROMOTED;  // user mask 0
    if (const char *e = getenv("BLISS_B200_VARIANT")) g.variant = atoi(e) ^ VARIANT_PROMOTED;
    if (const char *e = getenv("BLISS_B200_ORDER")) g.launch_order = atoi(e);
#ifdef BLISS_HOST_EMUL
    fprintf(stderr, "[bliss_b200] HOST-EMULATED TEST BUILD (tests/cpu_emul): kernels run on the CPU, thread by thread. "
                    "Not a product path -- the product library is built by nvcc and needs a B200.\n");
#endif
    // One carve-out (all shared) for every kernel was tried so that kernels of the two chains could share an SM: no
    // overlap appeared and timedomain_kernel lost its L1 (2.49 -> 3.25 ms), profiles/knobs_r02.md: off unless asked for.
    if (getenv("BLISS_B200_MAX_SHARED_CARVEOUT")) {
        configure_kernels_spectral();
        configure_kernels_chroma();
        configure_kernels_tempo();
        configure_kernels_finalize();
    }
    int rc = build_tables();
    if (rc) return rc;
    g.inited = true;
    return BLISS_B200_OK;
}

int bliss_b200_init(int device) {
    std::lock_guard<std::mutex> lk0(g_init_mu);
    g_cur = &g_ctx[0];
    std::lock_guard<std::mutex> lk(g.mu);
    const int rc = init_ctx_locked(device);
    if (rc == BLISS_B200_OK && g_n_ctx.load() < 1) g_n_ctx = 1;
    return rc;
}

// Multi-device form: contexts on devices 0 .. n_devices-1 (n_devices <= 0: every visible device).
// Afterwards ONE call of bliss_b200_analyze_batch / _s16 / _pcm shards its songs over all of them (longest first onto
// the least loaded device), one host thread and one copy stream per device, and bliss_b200_distance_matrix splits its
// rows the same way.  The device-pointer entry points keep using the primary context (device 0).  Returns the number
// of devices in use (> 0) or a negative error.
int bliss_b200_init_devices(int n_devices) {
    std::lock_guard<std::mutex> lk0(g_init_mu);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_last_error = "no CUDA device visible";
        return BLISS_B200_E_NO_DEVICE;
    }
    if (n_devices <= 0 || n_devices > count) n_devices = count;
    n_devices = std::min(n_devices, MAX_DEVICES);
    if (g_ctx[0].inited && g_ctx[0].device != 0) {
        g_last_error = "the primary context is bound to a device other than 0: shut down first";
        return BLISS_B200_E_ARG;
    }
    int rc = BLISS_B200_OK;
    for (int d = 0; d < n_devices && rc == BLISS_B200_OK; d++) {
        g_cur = &g_ctx[d];
        std::lock_guard<std::mutex> lk(g.mu);
        rc = init_ctx_locked(d);
    }
    g_cur = &g_ctx[0];
    cudaSetDevice(0);
    if (rc != BLISS_B200_OK) return rc;
    if (g_n_ctx.load() < n_devices) g_n_ctx = n_devices;
    return g_n_ctx.load();
}

int bliss_b200_device_count(void) { return g_n_ctx.load(); }

static void shutdown_ctx() {  // the calling thread's current context
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.inited) return;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    DevBuf *all[] = {&g.t_win512, &g.t_twA, &g.t_hann8k, &g.t_tw4k, &g.t_tw2, &g.t_tw8k, &g.t_filt, &g.t_filt32,
                     &g.pcm[0], &g.pcm[1], &g.pcm[2], &g.pcm[3], &g.raw16[0], &g.raw16[1], &g.raw16[2], &g.raw16[3], &g.feats, &g.metric, &g.misc[0], &g.misc[1],
                     &g.misc[2], &g.misc[3], &g.misc[4], &g.misc[5], &g.rs_tab, &g.rs_in[0], &g.rs_in[1], &g.rs_in[2], &g.rs_in[3],
                     &g.rs_jobs[0], &g.rs_jobs[1], &g.rs_jobs[2], &g.rs_jobs[3]};
    for (DevBuf *b : all) b->release();
    for (int i = 0; i < 4; i++) {
        if (g.rs_host[i]) cudaFreeHost(g.rs_host[i]);
        g.rs_host[i] = nullptr;
        g.rs_host_cap[i] = 0;
    }
    g.rs_rate = 0;
    for (auto &S : g.ws) {
        S.release();
        cudaStreamDestroy(S.main);
        cudaStreamDestroy(S.side);
        cudaEventDestroy(S.ev_fork);
        cudaEventDestroy(S.ev_join);
        cudaEventDestroy(S.ev_done);
        S.main = S.side = nullptr;
    }
    for (auto &p : g.ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    g.ev_pending.clear();
    for (auto e : g.ev_pool) cudaEventDestroy(e);
    g.ev_pool.clear();
    cudaStreamDestroy(g.stream);
    cudaStreamDestroy(g.copy_stream);
    cudaStreamDestroy(g.conv_stream);
    cudaEventDestroy(g.ev_begin);
    g.inited = false;
}

void bliss_b200_shutdown(void) {
    std::lock_guard<std::mutex> lk0(g_init_mu);
    for (int d = MAX_DEVICES - 1; d >= 0; d--) {
        g_cur = &g_ctx[d];
        shutdown_ctx();
    }
    g_cur = &g_ctx[0];
    g_n_ctx = 0;
}

// Diagnostic: which kernel implementations run (bit mask of common.cuh VARIANT_*; 0 = current).  Returns the
// previous mask.  The environment variable BLISS_B200_VARIANT sets the initial value.
int bliss_b200_set_variant(int mask) {
    // The cuts promoted in round 2 (profiles/ab_r02.md) are ON for mask 0: the user's bit switches a kernel BACK to
    // its previous implementation, like bits 1..32; internally the promoted bits are stored inverted.
    int prev = 0;
    for (int d = std::max(1, g_n_ctx.load()) - 1; d >= 0; d--) {  // every context; the primary one's mask is returned
        Ctx &c = g_ctx[d];
        std::lock_guard<std::mutex> lk(c.mu);
        prev = c.variant ^ VARIANT_PROMOTED;
        c.variant = mask ^ VARIANT_PROMOTED;
    }
    return prev;
}

int bliss_b200_set_workspace_limit(uint64_t bytes) {
    if (!g_ctx[0].inited) return BLISS_B200_E_NOT_INIT;
    for (int d = 0; d < std::max(1, g_n_ctx.load()); d++) {
        Ctx &c = g_ctx[d];
        std::lock_guard<std::mutex> lk(c.mu);
        c.ws_limit = (size_t)bytes;
    }
    return BLISS_B200_OK;
}

#define REQUIRE_INIT()                                                                  \
    std::lock_guard<std::mutex> lk(g.mu);                                               \
    if (!g.inited) { g_last_error = "bliss_b200_init() not called"; return BLISS_B200_E_NOT_INIT; } \
    CK(cudaSetDevice(g.device));

int bliss_b200_analyze_batch_device(const float *d_pcm, const uint64_t *offsets, const uint64_t *n_samples,
                                    uint32_t n_songs, uint16_t ver, float *d_out, int32_t *status,
                                    void *cuda_stream) {
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (n_songs == 0) return BLISS_B200_OK;
    if (!d_pcm || !offsets || !n_samples || !d_out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    return analyze_device_locked(d_pcm, offsets, n_samples, n_songs, ver, d_out, status,
                                 (cudaStream_t)cuda_stream, false);
}

// ---- fused feature-row exchange across the GPUs of one box -------------------
// One allocation per rank: two row buffers (double-buffered by epoch parity) followed by the flag
// array; exported as a CUDA IPC handle.  Ranks living in the same process (tests) are connected by
// raw pointer instead.
struct GatherWire {  // BLISS_B200_GATHER_HANDLE_BYTES bytes on the wire
    cudaIpcMemHandle_t ipc;
    uint64_t pid, raw_ptr, max_rows;
    int32_t device;
    uint32_t world, rank, magic;
    unsigned char pad[BLISS_B200_GATHER_HANDLE_BYTES - sizeof(cudaIpcMemHandle_t) - 3 * 8 - 4 * 4];
};
static_assert(sizeof(GatherWire) == BLISS_B200_GATHER_HANDLE_BYTES, "wire format");
constexpr uint32_t GATHER_MAGIC = 0xB2006A74u;

struct bliss_b200_gather {
    uint32_t world = 0, rank = 0;
    uint64_t max_rows = 0;
    size_t buf_floats = 0;  // floats per parity buffer
    char *local = nullptr;
    char *peer[MAX_PEERS] = {nullptr};
    bool opened[MAX_PEERS] = {false};
    bool connected = false;
    uint32_t epoch = 1;  // the open epoch; flags start at 0
    uint32_t dim = 0;    // row width of the open epoch (0 = nothing scattered yet)
    unsigned long long timeout_ns = 30ull * 1000000000ull;
    float *rows(int r, uint32_t ep) const { return reinterpret_cast<float *>(peer[r]) + (size_t)(ep & 1u) * buf_floats; }
    unsigned int *flags(int r) const { return reinterpret_cast<unsigned int *>(peer[r] + 2 * buf_floats * 4); }
};

int bliss_b200_gather_create(uint32_t world, uint32_t rank, uint64_t max_rows, void *handle_out,
                             bliss_b200_gather **out) {
    REQUIRE_INIT();
    if (!handle_out || !out || world == 0 || world > (uint32_t)MAX_PEERS || rank >= world || max_rows == 0) {
        g_last_error = "gather_create: need 1 <= world <= 8, rank < world, max_rows > 0";
        return BLISS_B200_E_ARG;
    }
    auto *ga = new bliss_b200_gather();
    ga->world = world;
    ga->rank = rank;
    ga->max_rows = max_rows;
    ga->buf_floats = align_up((size_t)max_rows * 23, 64);
    const size_t bytes = 2 * ga->buf_floats * 4 + 256;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { delete ga; g_last_error = std::string("cudaMalloc: ") + cudaGetErrorString(e); return BLISS_B200_E_CUDA; }
    ga->local = (char *)p;
    e = cudaMemset(p, 0, bytes);
    GatherWire w;
    memset(&w, 0, sizeof(w));
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&w.ipc, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        delete ga;
        g_last_error = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
        return BLISS_B200_E_CUDA;
    }
    w.pid = (uint64_t)getpid();
    w.raw_ptr = (uint64_t)(uintptr_t)p;
    w.max_rows = max_rows;
    w.device = g.device;
    w.world = world;
    w.rank = rank;
    w.magic = GATHER_MAGIC;
    memcpy(handle_out, &w, sizeof(w));
    *out = ga;
    return BLISS_B200_OK;
}

int bliss_b200_gather_connect(bliss_b200_gather *ga, const void *all_handles) {
    REQUIRE_INIT();
    if (!ga || !all_handles) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    if (ga->connected) { g_last_error = "gather already connected"; return BLISS_B200_E_ARG; }
    const GatherWire *w = reinterpret_cast<const GatherWire *>(all_handles);
    for (uint32_t r = 0; r < ga->world; r++) {
        GatherWire h;
        memcpy(&h, &w[r], sizeof(h));
        if (h.magic != GATHER_MAGIC || h.world != ga->world || h.rank != r || h.max_rows != ga->max_rows) {
            g_last_error = "gather_connect: handle " + std::to_string(r) + " does not describe rank " +
                           std::to_string(r) + " of this gather (world / max_rows mismatch?)";
            return BLISS_B200_E_ARG;
        }
        if (r == ga->rank) {
            ga->peer[r] = ga->local;
        } else if (h.pid == (uint64_t)getpid()) {
            // same process (several contexts of one test process): the pointer is directly usable
            if (h.device != g.device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, g.device, h.device));
                if (!can) { g_last_error = "gather_connect: no peer access to device " + std::to_string(h.device); return BLISS_B200_E_CUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                (void)cudaGetLastError();
            }
            ga->peer[r] = (char *)(uintptr_t)h.raw_ptr;
        } else {
            void *p = nullptr;
            CK(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            ga->peer[r] = (char *)p;
            ga->opened[r] = true;
        }
    }
    ga->connected = true;
    return BLISS_B200_OK;
}

int bliss_b200_analyze_batch_device_scatter(bliss_b200_gather *ga, const float *d_pcm, const uint64_t *offsets,
                                            const uint64_t *n_samples, uint32_t n_songs, uint16_t ver,
                                            uint64_t row_offset, uint64_t row_stride, float *d_out_local,
                                            int32_t *status, void *cuda_stream) {
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (!ga || !ga->connected) { g_last_error = "gather not connected"; return BLISS_B200_E_ARG; }
    const uint32_t dim = bliss_b200_feature_count(ver);
    if (ga->dim != 0 && ga->dim != dim) { g_last_error = "one epoch cannot mix feature versions"; return BLISS_B200_E_ARG; }
    if (n_songs == 0) return BLISS_B200_OK;
    if (!d_pcm || !offsets || !n_samples) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    if (row_stride == 0 || row_offset + (uint64_t)(n_songs - 1) * row_stride >= ga->max_rows ||
        row_stride > 0xffffffffull) {
        g_last_error = "scatter rows fall outside the gather buffer";
        return BLISS_B200_E_ARG;
    }
    PeerRows pr;
    memset(&pr, 0, sizeof(pr));
    pr.n_peers = (int)ga->world;
    for (uint32_t r = 0; r < ga->world; r++) pr.base[r] = ga->rows((int)r, ga->epoch);
    pr.row_offset = (unsigned int)row_offset;
    pr.row_stride = (unsigned int)row_stride;
    ga->dim = dim;
    return analyze_device_locked(d_pcm, offsets, n_samples, n_songs, ver, d_out_local, status,
                                 (cudaStream_t)cuda_stream, false, &pr);
}

int bliss_b200_gather_commit(bliss_b200_gather *ga, void *cuda_stream, const float **d_rows) {
    REQUIRE_INIT();
    if (!ga || !ga->connected) { g_last_error = "gather not connected"; return BLISS_B200_E_ARG; }
    unsigned int *fl[MAX_PEERS] = {nullptr};
    for (uint32_t r = 0; r < ga->world; r++) fl[r] = ga->flags((int)r);
    g.launches += (unsigned long long)launch_gather_barrier(fl, (int)ga->world, (int)ga->rank, ga->epoch,
                                                            ga->timeout_ns, (cudaStream_t)cuda_stream);
    CK(cudaGetLastError());
    if (d_rows) *d_rows = ga->rows((int)ga->rank, ga->epoch);
    ga->epoch++;
    ga->dim = 0;
    return BLISS_B200_OK;
}

int bliss_b200_gather_check(bliss_b200_gather *ga) {
    REQUIRE_INIT();
    if (!ga || !ga->local) { g_last_error = "null gather"; return BLISS_B200_E_ARG; }
    unsigned int st = 0;
    CK(cudaMemcpy(&st, ga->flags((int)ga->rank) + MAX_PEERS, 4, cudaMemcpyDeviceToHost));  // synchronises
    if (st != 0) {
        g_last_error = "gather barrier timed out waiting for rank " + std::to_string(st - 1);
        return BLISS_B200_E_TIMEOUT;
    }
    return BLISS_B200_OK;
}

int bliss_b200_gather_set_timeout(bliss_b200_gather *ga, uint64_t milliseconds) {
    if (!ga || milliseconds == 0) return BLISS_B200_E_ARG;
    ga->timeout_ns = (unsigned long long)milliseconds * 1000000ull;
    return BLISS_B200_OK;
}

int bliss_b200_gather_destroy(bliss_b200_gather *ga) {
    if (!ga) return BLISS_B200_OK;
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.inited) {
        cudaSetDevice(g.device);
        cudaDeviceSynchronize();
        for (uint32_t r = 0; r < ga->world; r++)
            if (ga->opened[r]) cudaIpcCloseMemHandle(ga->peer[r]);
        if (ga->local) cudaFree(ga->local);
    }
    delete ga;
    return BLISS_B200_OK;
}

}  // extern "C" (the host path below is internal)

// host buffers: chunks of songs are copied on the copy stream while earlier chunks compute
// HostPcm describes what the host pointers hold.  {F32, 1}: the decoder's output as the reference hands it to
// Song::analyze, copied straight into the PCM buffer.  Anything else (interleaved s16 / s32 / f32 frames of
// `channels` channels at 22 050 Hz) lands in a raw staging buffer and is converted / down-mixed on the device
// behind its copy (wave_setup.cu), so that e.g. 16-bit mono material sends half the bytes over PCIe.
struct HostPcm {
    int fmt;            // BLISS_B200_PCM_*
    uint32_t channels;
    uint32_t rate = (uint32_t)SAMPLE_RATE;
    size_t frame_bytes() const { return (size_t)(fmt == BLISS_B200_PCM_S16 ? 2 : 4) * channels; }
    bool mono_f32() const { return fmt == BLISS_B200_PCM_F32 && channels == 1; }
    bool resample() const { return rate != (uint32_t)SAMPLE_RATE; }
    bool direct() const { return mono_f32() && !resample(); }
};

// ---- sample-rate conversion to 22 050 Hz: the filter and the job table (kernel: wave_setup.cu) ----------------------
// scipy.signal.resample_poly's design, restated (scipy/signal/_signaltools.py resample_poly + _fir_filter_design.py
// firwin; oracle/resample.py is the numpy statement the tests check both against):
//   up / down = 22050 / rate in lowest terms, m = max(up, down), half = 10 m,
//   h[k] = (1/m) sinc((k - half) / m) kaiser_5(k), k = 0 .. 2 half, scaled to sum 1, times up;
//   down - half mod down zeros in front so that the delay is a whole number of outputs (pre_remove of them dropped).
// Laid out by phase for the kernel: row p = h'[p], h'[p + up], ... (taps4 floats, zero-filled).
static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = 0.25 * x * x;
    for (int k = 1; k < 500; k++) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

static uint64_t resample_len(uint64_t n, uint32_t rate) {  // src/song/decoder/symphonia.rs:379-380
    if (rate == (uint32_t)SAMPLE_RATE) return n;
    return (uint64_t)std::ceil((double)SAMPLE_RATE / (double)rate * (double)n);
}

static int design_resampler(uint32_t rate) {  // into the calling context; a no-op for the rate last designed
    if (g.rs_rate == rate) return BLISS_B200_OK;
    uint32_t a = (uint32_t)SAMPLE_RATE, b = rate;
    while (b) { const uint32_t t = a % b; a = b; b = t; }
    const uint32_t up = (uint32_t)SAMPLE_RATE / a, down = rate / a;
    const uint64_t m = std::max(up, down), half = 10 * m, numtaps = 2 * half + 1;
    std::vector<double> h(numtaps);
    const double pi = 3.14159265358979323846, i0b = bessel_i0(5.0);
    double sum = 0.0;
    for (uint64_t k = 0; k < numtaps; k++) {
        const double x = ((double)k - (double)half) / (double)m;
        const double sinc = k == half ? 1.0 : std::sin(pi * x) / (pi * x);
        const double r = ((double)k - (double)half) / (double)half;
        const double w = bessel_i0(5.0 * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
        h[k] = sinc / (double)m * w;
        sum += h[k];
    }
    const uint64_t pre_pad = down - half % down, L = pre_pad + numtaps;
    // the row length the ratio's kernel wants (a few sizes exist for the register-resident kernel; rows are zero-filled)
    const uint32_t taps4 = resample_plan(up, down, (uint32_t)align_up((size_t)((L + up - 1) / up), 4), 0).taps4;
    std::vector<float> tab((size_t)up * taps4, 0.f);
    for (uint32_t p = 0; p < up; p++)
        for (uint32_t t = 0; t < taps4; t++) {
            const uint64_t k = (uint64_t)p + (uint64_t)t * up;
            if (k >= pre_pad && k < L) tab[(size_t)p * taps4 + t] = (float)(h[k - pre_pad] / sum * (double)up);
        }
    CK(cudaDeviceSynchronize());  // a table of another rate may still be in use
    CK(g.rs_tab.ensure(tab.size() * 4));
    CK(cudaMemcpy(g.rs_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    g.rs_rate = rate;
    g.rs_up = up;
    g.rs_down = down;
    g.rs_taps4 = taps4;
    g.rs_pre = (uint32_t)((half + pre_pad) / down);
    return BLISS_B200_OK;
}

// One chunk's conversions: song i reads in[in_off[i] .. +in_len[i]) and writes out[out_off[i] .. +out_len[i]).  The job
// table goes through ring slot `slot`'s pinned host copy and is pulled onto the device by a kernel (wave_setup.cu says
// why); the caller guarantees that the slot's previous conversion has finished.
struct ResampleJobHost { unsigned long long in_off, in_len, out_off, out_len; };
static int enqueue_resample(int slot, const float *in, float *out, const std::vector<ResampleJobHost> &jobs_in, cudaStream_t st) {
    const ResamplePlan plan = resample_plan(g.rs_up, g.rs_down, g.rs_taps4, (g.variant & VARIANT_RESAMPLE_V1) ? 1 : 0);
    std::vector<ResampleJobHost> jobs;
    std::vector<unsigned int> prefix;
    unsigned long long tiles = 0;
    for (const auto &j : jobs_in) {
        if (j.out_len == 0) continue;
        jobs.push_back(j);
        prefix.push_back((unsigned int)tiles);
        tiles += (j.out_len + plan.tile_out - 1) / plan.tile_out;
    }
    if (jobs.empty()) return BLISS_B200_OK;
    if (tiles >= 0x7fffffffull) { g_last_error = "resampler: chunk too large"; return BLISS_B200_E_ARG; }
    prefix.push_back((unsigned int)tiles);
    const size_t jb = jobs.size() * sizeof(ResampleJobHost), bytes = align_up(jb + prefix.size() * 4, 16);
    if (g.rs_host_cap[slot] < bytes) {
        if (g.rs_host[slot]) CK(cudaFreeHost(g.rs_host[slot]));
        g.rs_host[slot] = nullptr;
        g.rs_host_cap[slot] = 0;
        CK(cudaMallocHost(&g.rs_host[slot], bytes * 2));
        g.rs_host_cap[slot] = bytes * 2;
    }
    memcpy(g.rs_host[slot], jobs.data(), jb);
    memcpy(static_cast<char *>(g.rs_host[slot]) + jb, prefix.data(), prefix.size() * 4);
    CK(g.rs_jobs[slot].ensure(bytes));
    g.launches += (unsigned long long)launch_wave_setup(g.rs_host[slot], g.rs_jobs[slot].p, bytes, nullptr, nullptr, 0, st);
    const int launched = launch_resample(in, out, g.rs_jobs[slot].p,
                                         reinterpret_cast<const unsigned int *>(g.rs_jobs[slot].as<char>() + jb),
                                         (unsigned int)jobs.size(), (unsigned int)tiles, g.rs_tab.as<float>(), g.rs_up, g.rs_down,
                                         g.rs_pre, plan, st);
    if (launched < 0) { g_last_error = "resampler: the filter table does not fit the kernel's shared memory"; return BLISS_B200_E_CUDA; }
    g.launches += (unsigned long long)launched;
    CK(cudaGetLastError());
    return BLISS_B200_OK;
}

static int analyze_host_locked(const void *const *pcm_v, const uint64_t *n_samples, uint32_t n_songs,
                               uint16_t ver, float *out, int32_t *status, bool debug,
                               HostPcm hp = HostPcm{BLISS_B200_PCM_F32, 1}) {
    const char *const *pcm = reinterpret_cast<const char *const *>(pcm_v);
    const bool kRaw = !hp.mono_f32();  // frames land in the raw staging buffer and are converted / down-mixed there
    const bool kRs = hp.resample();    // ... and the mono signal is at another rate: converted into the PCM buffer
    const size_t fb = hp.frame_bytes();
    if (kRs)
        if (int rc_d = design_resampler(hp.rate)) return rc_d;
    const uint32_t dim = bliss_b200_feature_count(ver);
    CK(g.feats.ensure((size_t)n_songs * dim * 4));
    // The path is PCIe-bound (15.9 MB per 3-min song).  Songs travel in chunks through a ring of FOUR
    // device PCM buffers; a chunk's waves wait only for that chunk's copy and run on the wave sets'
    // own streams, so the copy engine runs up to three chunks ahead and the ~3.5 ms latency floor of a
    // chunk's kernel chain (tuning, beat tracker) overlaps the next chunks' kernels instead of
    // serialising (measured with BLISS_B200_TRACE before the wave sets existed: 128 MB chunks 37 GB/s,
    // 512 MB chunks 50 GB/s of a 55.6 GB/s link, but 9 ms of exposed first copy).
    // The first chunks are small (64, 128, ... MB) so that compute starts early.
    // BLISS_B200_CHUNK_MB overrides the steady chunk size for experiments.
    size_t chunk_mb = 256;
    if (const char *e = getenv("BLISS_B200_CHUNK_MB")) chunk_mb = std::max(1, atoi(e));
    const size_t chunk_budget = std::min<size_t>(chunk_mb << 20, std::max<size_t>(g.ws_limit / 8, (size_t)64 << 20));
    const bool trace = getenv("BLISS_B200_TRACE") != nullptr;
    cudaEvent_t tr[4] = {nullptr, nullptr, nullptr, nullptr};  // copy begin/end, compute begin/end
    if (trace) {
        for (auto &e : tr) CK(cudaEventCreate(&e));
        CK(cudaEventRecord(tr[0], g.copy_stream));
        CK(cudaEventRecord(tr[2], g.stream));
    }
    std::vector<cudaEvent_t> tr_cb, tr_ce, tr_done;  // per chunk: copy begin / end, compute done (trace only)
    std::vector<size_t> tr_bytes;
    size_t done_bytes = 0;
    constexpr int NBUF = 4;
    cudaEvent_t ev_copy[NBUF], ev_raw[NBUF], ev_done[NBUF][N_SETS];
    bool done_used[NBUF][N_SETS], copy_used[NBUF] = {false, false, false, false};
    // The conversions behind a chunk's copy (sample format, down-mix, sample rate) run on their own stream: on the copy
    // stream they would hold back the next chunk's copy for as long as they take (measured: 16-bit mono sources 6 424 ->
    // 6 614 songs/s, CD-format sources 1 528 -> 1 602 songs/s; BLISS_B200_CONV_ON_COPY_STREAM=1 goes back).
    const bool conv_apart = (kRaw || kRs) && !getenv("BLISS_B200_CONV_ON_COPY_STREAM");
    for (int i = 0; i < NBUF; i++) {
        CK(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_raw[i], cudaEventDisableTiming));
        for (int k = 0; k < N_SETS; k++) {
            CK(cudaEventCreateWithFlags(&ev_done[i][k], cudaEventDisableTiming));
            done_used[i][k] = false;
        }
    }
    // offs / lens: the songs of the chunk at 22 050 Hz in the PCM buffer; offs_in / lens_in: as they arrive (the same
    // numbers unless the call resamples)
    std::vector<uint64_t> offs, lens, offs_in, lens_in;
    std::vector<ResampleJobHost> jobs;
    uint32_t first = 0;
    int c = 0, rc = BLISS_B200_OK;
    while (first < n_songs && rc == BLISS_B200_OK) {
        // chunk = as many songs as fit the PCM budget (at least one)
        const size_t budget = std::min<size_t>(chunk_budget, c < 8 ? ((size_t)64 << 20) << c : chunk_budget);
        size_t samples = 0, samples_in = 0;
        uint32_t count = 0;
        offs.clear();
        lens.clear();
        offs_in.clear();
        lens_in.clear();
        while (first + count < n_songs) {
            const uint64_t n_in = n_samples[first + count], n_22k = resample_len(n_in, hp.rate);
            const size_t len = align_up((size_t)n_22k, 4), len_in = align_up((size_t)n_in, 4);
            // the largest of the buffers involved
            if (count > 0 && std::max((samples + len) * 4, (samples_in + len_in) * std::max<size_t>(4, fb)) > budget) break;
            offs.push_back(samples);
            lens.push_back(n_22k);
            offs_in.push_back(samples_in);
            lens_in.push_back(n_in);
            samples += len;
            samples_in += len_in;
            count++;
        }
        const int b = c % NBUF;
        for (int k = 0; k < N_SETS; k++)  // buffer b free again (also: the host may reallocate it)
            if (done_used[b][k]) CK(cudaEventSynchronize(ev_done[b][k]));
        if (copy_used[b]) CK(cudaEventSynchronize(ev_copy[b]));  // the slot's staging buffers and job table as well
        CK(g.pcm[b].ensure(std::max<size_t>(samples, 4) * 4));
        if (kRaw) CK(g.raw16[b].ensure(std::max<size_t>(samples_in, 4) * fb));
        if (kRs) CK(g.rs_in[b].ensure(std::max<size_t>(samples_in, 4) * 4));
        float *const mono = kRs ? g.rs_in[b].as<float>() : g.pcm[b].as<float>();  // where the mono f32 signal lands
        if (trace) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            CK(cudaEventRecord(e, g.copy_stream));
            tr_cb.push_back(e);
        }
        for (uint32_t i = 0; i < count;) {
            if (lens_in[i] == 0) { i++; continue; }
            if (!pcm[first + i]) { g_last_error = "null pcm pointer"; rc = BLISS_B200_E_ARG; break; }
            // songs that sit back to back in host memory (one big decoded buffer) go as ONE copy
            uint32_t j = i;
            size_t run = (size_t)lens_in[i];
            while (j + 1 < count && lens_in[j + 1] > 0 && (lens_in[j] & 3u) == 0 &&
                   pcm[first + j + 1] == pcm[first + j] + (size_t)lens_in[j] * fb) {
                j++;
                run += (size_t)lens_in[j];
            }
            void *dst = kRaw ? (void *)(g.raw16[b].as<char>() + (size_t)offs_in[i] * fb) : (void *)(mono + offs_in[i]);
            CK(cudaMemcpyAsync(dst, pcm[first + i], run * fb, cudaMemcpyHostToDevice, g.copy_stream));
            i = j + 1;
        }
        if (rc) break;
        cudaStream_t cs = g.copy_stream;  // where the chunk becomes ready
        if (conv_apart) {
            CK(cudaEventRecord(ev_raw[b], g.copy_stream));
            CK(cudaStreamWaitEvent(g.conv_stream, ev_raw[b], 0));
            cs = g.conv_stream;
        }
        if (kRaw) {  // same frame offsets in both buffers; runs behind the chunk's copies
            if (hp.fmt == BLISS_B200_PCM_S16 && hp.channels == 1)
                g.launches += (unsigned long long)launch_s16_to_f32(g.raw16[b].as<short>(), mono, samples_in, cs);
            else
                g.launches += (unsigned long long)launch_pcm_to_mono(g.raw16[b].p, mono, samples_in, hp.fmt, hp.channels, cs);
            CK(cudaGetLastError());
        }
        if (kRs) {  // mono at hp.rate -> the PCM buffer at 22 050 Hz
            jobs.clear();
            for (uint32_t i = 0; i < count; i++) jobs.push_back(ResampleJobHost{offs_in[i], lens_in[i], offs[i], lens[i]});
            if ((rc = enqueue_resample(b, mono, g.pcm[b].as<float>(), jobs, cs)) != BLISS_B200_OK) break;
        }
        CK(cudaEventRecord(ev_copy[b], cs));
        copy_used[b] = true;
        if (trace) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            CK(cudaEventRecord(e, g.copy_stream));
            tr_ce.push_back(e);
            tr_bytes.push_back(samples * 4);
        }
        rc = analyze_device_locked(g.pcm[b].as<float>(), offs.data(), lens.data(), count, ver,
                                   g.feats.as<float>() + (size_t)first * dim, status ? status + first : nullptr,
                                   g.stream, debug, nullptr, ev_copy[b], ev_done[b], done_used[b]);
        if (trace && rc == BLISS_B200_OK) {  // a chunk this small is one wave: its set is the one just used
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            CK(cudaEventRecord(e, g.ws[(g.next_set + N_SETS - 1) % N_SETS].main));
            tr_done.push_back(e);
        }
        first += count;
        done_bytes += samples * 4;
        c++;
    }
    if (rc == BLISS_B200_OK) {
        // the result copy waits for every wave set (and, in serial mode, g.stream ran the waves itself)
        for (auto &S : g.ws) {
            CK(cudaEventRecord(S.ev_done, S.main));
            CK(cudaStreamWaitEvent(g.stream, S.ev_done, 0));
        }
        if (trace) {
            CK(cudaEventRecord(tr[1], g.copy_stream));
            CK(cudaEventRecord(tr[3], g.stream));
        }
        CK(cudaMemcpyAsync(out, g.feats.p, (size_t)n_songs * dim * 4, cudaMemcpyDeviceToHost, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        if (trace) {
            float ms_copy = 0.f, ms_comp = 0.f, ms_all = 0.f;
            cudaEventElapsedTime(&ms_copy, tr[0], tr[1]);
            cudaEventElapsedTime(&ms_comp, tr[2], tr[3]);
            cudaEventElapsedTime(&ms_all, tr[0], tr[3]);
            fprintf(stderr, "[bliss_b200 trace] songs=%u bytes=%.1f MB chunks=%d chunk_mb=%zu copy_stream=%.2f ms (%.1f GB/s) "
                            "first_launch->last_kernel=%.2f ms first_copy->last_kernel=%.2f ms\n",
                    n_songs, done_bytes / 1e6, c, chunk_mb, ms_copy, done_bytes / 1e6 / ms_copy, ms_comp, ms_all);
            // per chunk: when its copy started / ended and when its kernels were done, relative to the first copy
            double busy = 0.;
            for (size_t k = 0; k < tr_ce.size() && k < tr_done.size(); k++) {
                float t_b = 0.f, t_e = 0.f, t_d = 0.f;
                cudaEventElapsedTime(&t_b, tr[0], tr_cb[k]);
                cudaEventElapsedTime(&t_e, tr[0], tr_ce[k]);
                cudaEventElapsedTime(&t_d, tr[0], tr_done[k]);
                busy += t_e - t_b;
                if (getenv("BLISS_B200_TRACE_CHUNKS"))
                    fprintf(stderr, "[bliss_b200 trace]   chunk %2zu %6.1f MB copy %7.2f -> %7.2f ms (%.1f GB/s)  kernels done %7.2f ms (+%.2f)\n",
                            k, tr_bytes[k] / 1e6, t_b, t_e, tr_bytes[k] / 1e6 / (t_e - t_b), t_d, t_d - t_e);
            }
            fprintf(stderr, "[bliss_b200 trace] copy engine busy %.2f ms of %.2f ms (%.1f GB/s while copying)\n", busy, ms_copy,
                    done_bytes / 1e6 / busy);
            for (auto &e : tr) cudaEventDestroy(e);
        }
    } else {
        cudaDeviceSynchronize();
    }
    for (int i = 0; i < NBUF; i++) {
        cudaEventDestroy(ev_copy[i]);
        cudaEventDestroy(ev_raw[i]);
        for (int k = 0; k < N_SETS; k++) cudaEventDestroy(ev_done[i][k]);
    }
    for (auto *v : {&tr_cb, &tr_ce, &tr_done})
        for (auto e : *v) cudaEventDestroy(e);
    return rc;
}

// ---- one call, every device (bliss_b200_init_devices) --------------------------------------------------------
// Songs are independent (Song::analyze is a pure function): the call's songs are dealt longest
// first onto the least loaded device (LPT; equal lengths end up round-robin) and each device's share runs
// analyze_host_locked on its own context from its own host thread -- its own copy stream, wave sets and PCM ring,
// so the H2D copies of all devices are in flight together.  Per-song results do not depend on the device or on the
// batch they travel in: the rows are bit-identical to a single-device call (tests/test_gpu_multidevice.py).
static std::vector<std::vector<uint32_t>> shard_lpt(const uint64_t *n_samples, uint32_t n_songs, int nd) {
    std::vector<uint32_t> order(n_songs);
    for (uint32_t i = 0; i < n_songs; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return n_samples[a] > n_samples[b]; });
    std::vector<std::vector<uint32_t>> mine((size_t)nd);
    std::vector<uint64_t> load((size_t)nd, 0);
    for (uint32_t i : order) {
        int best = 0;
        for (int d = 1; d < nd; d++)
            if (load[d] < load[best]) best = d;
        mine[best].push_back(i);
        load[best] += n_samples[i] + 4096;  // a per-song constant keeps very short songs spread out as well
    }
    for (auto &v : mine) std::sort(v.begin(), v.end());  // original order inside a device: consecutive host buffers stay one copy
    return mine;
}

static int analyze_host_multi(const void *const *pcm, const uint64_t *n_samples, uint32_t n_songs, uint16_t ver,
                              float *out, int32_t *status, HostPcm hp) {
    const int nd = g_n_ctx.load();
    const uint32_t dim = bliss_b200_feature_count(ver);
    const auto mine = shard_lpt(n_samples, n_songs, nd);
    std::vector<int> rcs((size_t)nd, BLISS_B200_OK);
    std::vector<std::string> errs((size_t)nd);
    std::vector<std::thread> workers;
    for (int d = 0; d < nd; d++) {
        if (mine[d].empty()) continue;
        workers.emplace_back([&, d] {
            g_cur = &g_ctx[d];
            std::lock_guard<std::mutex> lk(g.mu);
            const std::vector<uint32_t> &ids = mine[d];
            const uint32_t m = (uint32_t)ids.size();
            auto fail = [&](int rc) { rcs[d] = rc; errs[d] = g_last_error; };
            if (!g.inited) { g_last_error = "device context not initialised"; return fail(BLISS_B200_E_NOT_INIT); }
            if (cudaSetDevice(g.device) != cudaSuccess) { g_last_error = "cudaSetDevice failed"; return fail(BLISS_B200_E_CUDA); }
            std::vector<const void *> ptrs(m);
            std::vector<uint64_t> lens(m);
            for (uint32_t k = 0; k < m; k++) { ptrs[k] = pcm[ids[k]]; lens[k] = n_samples[ids[k]]; }
            std::vector<float> lout((size_t)m * dim, 0.f);
            std::vector<int32_t> lst(m, 0);
            const int rc = analyze_host_locked(ptrs.data(), lens.data(), m, ver, lout.data(), lst.data(), false, hp);
            if (rc) return fail(rc);
            for (uint32_t k = 0; k < m; k++) {
                if (lst[k] == 0) memcpy(out + (size_t)ids[k] * dim, lout.data() + (size_t)k * dim, (size_t)dim * 4);  // a rejected song's row is left alone
                if (status) status[ids[k]] = lst[k];
            }
        });
    }
    for (auto &t : workers) t.join();
    for (int d = 0; d < nd; d++)
        if (rcs[d]) { g_last_error = "device " + std::to_string(d) + ": " + errs[d]; return rcs[d]; }
    return BLISS_B200_OK;
}

// `true` when this call is to be sharded over several devices
static bool use_all_devices(uint32_t n_songs) { return g_n_ctx.load() > 1 && n_songs > 1 && g_cur == &g_ctx[0]; }

extern "C" {

int bliss_b200_analyze_batch(const float *const *pcm, const uint64_t *n_samples, uint32_t n_songs,
                             uint16_t ver, float *out, int32_t *status) {
    if (use_all_devices(n_songs)) {
        if (check_version(ver)) return BLISS_B200_E_ARG;
        if (!pcm || !n_samples || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
        return analyze_host_multi(reinterpret_cast<const void *const *>(pcm), n_samples, n_songs, ver, out, status,
                                  HostPcm{BLISS_B200_PCM_F32, 1});
    }
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (n_songs == 0) return BLISS_B200_OK;
    if (!pcm || !n_samples || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    return analyze_host_locked(reinterpret_cast<const void *const *>(pcm), n_samples, n_songs, ver, out, status, false);
}

int bliss_b200_analyze_batch_s16(const int16_t *const *pcm, const uint64_t *n_samples, uint32_t n_songs,
                                 uint16_t ver, float *out, int32_t *status) {
    if (use_all_devices(n_songs)) {
        if (check_version(ver)) return BLISS_B200_E_ARG;
        if (!pcm || !n_samples || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
        return analyze_host_multi(reinterpret_cast<const void *const *>(pcm), n_samples, n_songs, ver, out, status,
                                  HostPcm{BLISS_B200_PCM_S16, 1});
    }
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (n_songs == 0) return BLISS_B200_OK;
    if (!pcm || !n_samples || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    return analyze_host_locked(reinterpret_cast<const void *const *>(pcm), n_samples, n_songs, ver, out, status, false,
                               HostPcm{BLISS_B200_PCM_S16, 1});
}

static int check_pcm_format(int fmt, uint32_t channels, uint32_t sample_rate) {
    if (fmt != BLISS_B200_PCM_S16 && fmt != BLISS_B200_PCM_S32 && fmt != BLISS_B200_PCM_F32) {
        g_last_error = "unknown sample format " + std::to_string(fmt) + " (BLISS_B200_PCM_S16 / _S32 / _F32)";
        return BLISS_B200_E_ARG;
    }
    if (channels == 0 || channels > BLISS_B200_PCM_MAX_CHANNELS) {
        g_last_error = "channel count " + std::to_string(channels) + " outside 1.." + std::to_string(BLISS_B200_PCM_MAX_CHANNELS);
        return BLISS_B200_E_ARG;
    }
    if (sample_rate < BLISS_B200_MIN_SAMPLE_RATE || sample_rate > BLISS_B200_MAX_SAMPLE_RATE) {
        g_last_error = "sample rate " + std::to_string(sample_rate) + " Hz outside " + std::to_string(BLISS_B200_MIN_SAMPLE_RATE) +
                       ".." + std::to_string(BLISS_B200_MAX_SAMPLE_RATE);
        return BLISS_B200_E_UNSUPPORTED;
    }
    return 0;
}

int bliss_b200_analyze_batch_pcm(const void *const *pcm, const uint64_t *n_frames, uint32_t n_songs, int sample_format,
                                 uint32_t channels, uint32_t sample_rate, uint16_t ver, float *out, int32_t *status) {
    if (use_all_devices(n_songs)) {
        if (check_version(ver)) return BLISS_B200_E_ARG;
        if (int rc = check_pcm_format(sample_format, channels, sample_rate)) return rc;
        if (!pcm || !n_frames || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
        return analyze_host_multi(pcm, n_frames, n_songs, ver, out, status, HostPcm{sample_format, channels, sample_rate});
    }
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (int rc = check_pcm_format(sample_format, channels, sample_rate)) return rc;
    if (n_songs == 0) return BLISS_B200_OK;
    if (!pcm || !n_frames || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    return analyze_host_locked(pcm, n_frames, n_songs, ver, out, status, false, HostPcm{sample_format, channels, sample_rate});
}

uint64_t bliss_b200_resampled_len(uint64_t n_samples, uint32_t sample_rate) {
    if (sample_rate < BLISS_B200_MIN_SAMPLE_RATE || sample_rate > BLISS_B200_MAX_SAMPLE_RATE) return 0;
    return resample_len(n_samples, sample_rate);
}

int bliss_b200_resample(const float *pcm, uint64_t n_samples, uint32_t sample_rate, float *out, uint64_t out_capacity,
                        uint64_t *n_out) {
    REQUIRE_INIT();
    if (int rc = check_pcm_format(BLISS_B200_PCM_F32, 1, sample_rate)) return rc;
    const uint64_t n_22k = resample_len(n_samples, sample_rate);
    if (n_out) *n_out = n_22k;
    if (n_22k == 0) return BLISS_B200_OK;
    if (!pcm || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    if (out_capacity < n_22k) {
        g_last_error = "output buffer holds " + std::to_string(out_capacity) + " samples, " + std::to_string(n_22k) + " needed";
        return BLISS_B200_E_ARG;
    }
    if (sample_rate == (uint32_t)SAMPLE_RATE) {
        memcpy(out, pcm, (size_t)n_samples * 4);
        return BLISS_B200_OK;
    }
    if (int rc = design_resampler(sample_rate)) return rc;
    CK(cudaDeviceSynchronize());  // ring slot 0 belongs to this call alone
    CK(g.rs_in[0].ensure(align_up((size_t)n_samples, 4) * 4));
    CK(g.pcm[0].ensure(align_up((size_t)n_22k, 4) * 4));
    CK(cudaMemcpyAsync(g.rs_in[0].p, pcm, (size_t)n_samples * 4, cudaMemcpyHostToDevice, g.stream));
    const std::vector<ResampleJobHost> jobs{ResampleJobHost{0, n_samples, 0, n_22k}};
    if (int rc = enqueue_resample(0, g.rs_in[0].as<float>(), g.pcm[0].as<float>(), jobs, g.stream)) return rc;
    CK(cudaMemcpyAsync(out, g.pcm[0].p, (size_t)n_22k * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return BLISS_B200_OK;
}

int bliss_b200_pcm_to_mono(const void *pcm, uint64_t n_frames, int sample_format, uint32_t channels, float *out) {
    REQUIRE_INIT();
    if (int rc = check_pcm_format(sample_format, channels, (uint32_t)SAMPLE_RATE)) return rc;
    if (n_frames == 0) return BLISS_B200_OK;
    if (!pcm || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    const HostPcm hp{sample_format, channels};
    const size_t padded = align_up((size_t)n_frames, 4);
    CK(g.raw16[0].ensure(padded * hp.frame_bytes()));
    CK(g.pcm[0].ensure(padded * 4));
    CK(cudaMemcpyAsync(g.raw16[0].p, pcm, (size_t)n_frames * hp.frame_bytes(), cudaMemcpyHostToDevice, g.stream));
    g.launches += (unsigned long long)launch_pcm_to_mono(g.raw16[0].p, g.pcm[0].as<float>(), padded, hp.fmt, hp.channels, g.stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, g.pcm[0].p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return BLISS_B200_OK;
}

int bliss_b200_analyze(const float *pcm, uint64_t n, uint16_t ver, float *out) {
    int32_t status = 0;
    const float *p[1] = {pcm};
    uint64_t len[1] = {n};
    {
        REQUIRE_INIT();
        if (check_version(ver)) return BLISS_B200_E_ARG;
        if (!out || (!pcm && n > 0)) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
        int rc = analyze_host_locked(reinterpret_cast<const void *const *>(p), len, 1, ver, out, &status, false);
        if (rc) return rc;
    }
    return status;
}

int bliss_b200_analyze_taps(const float *pcm, uint64_t n, uint16_t ver, float *out, const bliss_b200_taps *t) {
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (!out || (!pcm && n > 0)) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    int32_t status = 0;
    const float *p[1] = {pcm};
    uint64_t len[1] = {n};
    int rc = analyze_host_locked(reinterpret_cast<const void *const *>(p), len, 1, ver, out, &status, true);
    if (rc) return rc;
    if (status != 0 || !t) return status;
    const SongGeom q = geom_of(n);
    auto dl = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        return dst ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : cudaSuccess;
    };
    CK(dl(t->centroid, g.ws[0].cent.p, (size_t)q.n_s * 4));
    CK(dl(t->rolloff, g.ws[0].roll.p, (size_t)q.n_s * 4));
    CK(dl(t->flatness, g.ws[0].flat.p, (size_t)q.n_s * 4));
    CK(dl(t->flux, g.ws[0].flux.p, (size_t)q.n_t * 4));
    CK(dl(t->thresholded, g.ws[0].thr.p, (size_t)q.n_t * 4));
    uint32_t nb = 0;
    CK(cudaMemcpy(&nb, g.ws[0].bpm_count.p, 4, cudaMemcpyDeviceToHost));
    if (t->n_bpms) *t->n_bpms = nb;
    CK(dl(t->bpms, g.ws[0].bpm.p, (size_t)nb * 4));
    CK(dl(t->loudness_chunks, g.ws[0].loud.p, (size_t)q.n_l * 4));
    CK(dl(t->zero_crossings, g.ws[0].zcr.p, 4));
    if (t->stft8192) {
        CK(cudaMemcpy2D(t->stft8192, (size_t)CH_BINS * 4, g.ws[0].mags.p, (size_t)CH_STRIDE * 4, (size_t)CH_BINS * 4,
                        q.n_c_comp, cudaMemcpyDeviceToHost));
        for (uint32_t f = q.n_c_comp; f < q.n_c; f++) memset(t->stft8192 + (size_t)f * CH_BINS, 0, (size_t)CH_BINS * 4);
    }
    if (t->n_peaks) {
        uint32_t c = 0;
        CK(cudaMemcpy(&c, g.ws[0].cand_count.p, 4, cudaMemcpyDeviceToHost));
        *t->n_peaks = c;
    }
    if (t->peak_pitches || t->peak_mags) {
        unsigned int c = 0;
        CK(cudaMemcpy(&c, g.ws[0].cand_count.p, 4, cudaMemcpyDeviceToHost));
        CK(dl(t->peak_pitches, g.ws[0].cand_pitch.p, (size_t)c * sizeof(double)));
        CK(dl(t->peak_mags, g.ws[0].cand_mag.p, (size_t)c * sizeof(double)));
    }
    if (t->tuning) {
        int idx = 0;
        CK(cudaMemcpy(&idx, g.ws[0].tuning.p, 4, cudaMemcpyDeviceToHost));
        *t->tuning = (-50. + (100. * 0.01 * (double)idx)) / 100.;
    }
    CK(dl(t->chroma, g.ws[0].chroma_dbg.p, (size_t)q.n_c * 12 * sizeof(double)));
    if (t->interval_features) {
        std::vector<double> parts((size_t)q.n_tiles * 10);
        CK(cudaMemcpy(parts.data(), g.ws[0].tiles.p, parts.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 10; k++) {
            double acc = 0.;
            for (uint32_t tl = 0; tl < q.n_tiles; tl++) acc += parts[(size_t)tl * 10 + k];
            t->interval_features[k] = acc / (double)q.n_c;
        }
    }
    return status;
}

int bliss_b200_stft512_mag_device(const float *d_pcm, const uint64_t *offsets, const uint64_t *n_samples,
                                  uint32_t n_songs, float *d_mags, uint64_t *frame_offsets_out, void *cuda_stream) {
    REQUIRE_INIT();
    if (!d_pcm || !offsets || !n_samples || !d_mags) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    WavePlan w;
    plan_wave(offsets, n_samples, 0, n_songs, true, w);
    if (frame_offsets_out) {
        for (uint32_t i = 0; i < n_songs; i++) frame_offsets_out[i] = w.sd[i].t_off;
        frame_offsets_out[n_songs] = w.n_t;
    }
    WaveDev dv;
    WaveSet &S = g.ws[0];  // borrows set 0's descriptor staging: order `st` behind and ahead of its waves
    CK(cudaEventRecord(S.ev_done, S.main));
    CK(cudaStreamWaitEvent(st, S.ev_done, 0));
    int rc = upload_plan(w, S, st, dv, nullptr, nullptr);
    if (rc) return rc;
    {
        ProfScope p(K_STFT512, st);
        p.done(launch_stft512_mags(d_pcm, dv.sd, dv.k1_prefix, (int)n_songs, w.k1_prefix[n_songs],
                                   (int)w.pairs_per_item, pvoc_tables(), d_mags, g.variant, st));
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(S.ev_done, st));
    CK(cudaStreamWaitEvent(S.main, S.ev_done, 0));
    return BLISS_B200_OK;
}

// ---- distances ---------------------------------------------------------------------------
int bliss_b200_feature_weights(uint16_t ver, float *m) {
    if (check_version(ver) || !m) return BLISS_B200_E_ARG;
    const uint32_t dim = bliss_b200_feature_count(ver);
    memset(m, 0, sizeof(float) * dim * dim);
    for (uint32_t i = 0; i < dim; i++) {
        float w = 1.f;
        if (ver == 2) w = (i == 0) ? 0.25f : (i >= 10 ? 3.f / 13.f : 1.f);  // VERSION2_WEIGHTS, lib.rs:209-234
        m[(size_t)i * dim + i] = w;
    }
    return BLISS_B200_OK;
}

int bliss_b200_distance_matrix_device(const float *d_rows, uint32_t n_rows, const float *d_cols, uint32_t n_cols,
                                      uint32_t dim, int metric, const float *m, float *d_out, void *cuda_stream) {
    REQUIRE_INIT();
    if (!d_rows || !d_cols || !d_out || dim == 0 || dim > 64) { g_last_error = "bad argument"; return BLISS_B200_E_ARG; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int mode;
    const float *d_w;
    int rc = prepare_metric(metric, m, dim, st, mode, d_w);
    if (rc) return rc;
    ProfScope p(K_DIST, st);
    const int nl = launch_distance_matrix(d_rows, n_rows, d_cols, n_cols, (int)dim, mode, d_w, d_out, st, g.metric_ones, g.variant);
    p.done(nl);
    if (nl < 0) { g_last_error = "unsupported dim"; return BLISS_B200_E_ARG; }
    CK(cudaGetLastError());
    return BLISS_B200_OK;
}

static int distance_matrix_host_locked(const float *rows, uint32_t n_rows, const float *cols, uint32_t n_cols,
                                       uint32_t dim, int metric, const float *m, float *out) {
    cudaStream_t st = g.stream;
    CK(g.misc[0].ensure((size_t)n_rows * dim * 4 + 16));
    CK(g.misc[1].ensure((size_t)n_cols * dim * 4 + 16));
    CK(g.misc[2].ensure((size_t)n_rows * n_cols * 4 + 16));
    CK(cudaMemcpyAsync(g.misc[0].p, rows, (size_t)n_rows * dim * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.misc[1].p, cols, (size_t)n_cols * dim * 4, cudaMemcpyHostToDevice, st));
    int mode;
    const float *d_w;
    int rc = prepare_metric(metric, m, dim, st, mode, d_w);
    if (rc) return rc;
    ProfScope p(K_DIST, st);
    const int nl = launch_distance_matrix(g.misc[0].as<float>(), n_rows, g.misc[1].as<float>(), n_cols, (int)dim,
                                          mode, d_w, g.misc[2].as<float>(), st, g.metric_ones, g.variant);
    p.done(nl);
    if (nl < 0) { g_last_error = "unsupported dim or too many rows in one call"; return BLISS_B200_E_ARG; }
    CK(cudaGetLastError());  // an invalid launch configuration must not come back as a zero-filled matrix
    CK(cudaMemcpyAsync(out, g.misc[2].p, (size_t)n_rows * n_cols * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BLISS_B200_OK;
}

int bliss_b200_distance_matrix(const float *rows, uint32_t n_rows, const float *cols, uint32_t n_cols,
                               uint32_t dim, int metric, const float *m, float *out) {
    if (!rows || !cols || !out || dim == 0 || dim > 64) { g_last_error = "bad argument"; return BLISS_B200_E_ARG; }
    if (n_rows == 0 || n_cols == 0) return BLISS_B200_OK;
    const int nd = g_n_ctx.load();
    if (nd > 1 && g_cur == &g_ctx[0] && (uint64_t)n_rows * n_cols >= (1ull << 22)) {
        // row blocks over the devices: every entry is computed by the same kernel whatever the
        // split, so the matrix is bit-identical to the single-device one
        std::vector<int> rcs((size_t)nd, BLISS_B200_OK);
        std::vector<std::string> errs((size_t)nd);
        std::vector<std::thread> workers;
        for (int d = 0; d < nd; d++) {
            const uint32_t r0 = (uint32_t)((uint64_t)n_rows * d / nd), r1 = (uint32_t)((uint64_t)n_rows * (d + 1) / nd);
            if (r1 == r0) continue;
            workers.emplace_back([&, d, r0, r1] {
                g_cur = &g_ctx[d];
                std::lock_guard<std::mutex> lk(g.mu);
                if (!g.inited || cudaSetDevice(g.device) != cudaSuccess) { rcs[d] = BLISS_B200_E_CUDA; errs[d] = "device context unavailable"; return; }
                rcs[d] = distance_matrix_host_locked(rows + (size_t)r0 * dim, r1 - r0, cols, n_cols, dim, metric, m,
                                                     out + (size_t)r0 * n_cols);
                if (rcs[d]) errs[d] = g_last_error;
            });
        }
        for (auto &t : workers) t.join();
        for (int d = 0; d < nd; d++)
            if (rcs[d]) { g_last_error = "device " + std::to_string(d) + ": " + errs[d]; return rcs[d]; }
        return BLISS_B200_OK;
    }
    REQUIRE_INIT();
    return distance_matrix_host_locked(rows, n_rows, cols, n_cols, dim, metric, m, out);
}

int bliss_b200_distance(const float *a, const float *b, uint32_t dim, int metric, const float *m, float *out) {
    REQUIRE_INIT();
    if (!a || !b || !out || dim == 0 || dim > 64) { g_last_error = "bad argument"; return BLISS_B200_E_ARG; }
    return distance_matrix_host_locked(a, 1, b, 1, dim, metric, m, out);
}

int bliss_b200_closest_to_songs(const float *seeds, uint32_t n_seeds, const float *cands, uint32_t n_cands,
                                uint32_t dim, int metric, const float *m, uint32_t *order, float *keys) {
    REQUIRE_INIT();
    if (!seeds || !cands || !order || dim == 0 || dim > 64) { g_last_error = "bad argument"; return BLISS_B200_E_ARG; }
    if (n_cands == 0) return BLISS_B200_OK;
    cudaStream_t st = g.stream;
    const size_t tmp_bytes = sort_temp_bytes(n_cands);
    CK(g.misc[0].ensure((size_t)std::max<uint32_t>(n_seeds, 1) * dim * 4));
    CK(g.misc[1].ensure((size_t)n_cands * dim * 4));
    CK(g.misc[2].ensure((size_t)n_cands * 4));
    CK(g.misc[3].ensure((size_t)n_cands * 8 * 2));
    CK(g.misc[4].ensure(tmp_bytes + 16));
    CK(g.misc[5].ensure((size_t)n_cands * 4));
    CK(cudaMemcpyAsync(g.misc[0].p, seeds, (size_t)n_seeds * dim * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.misc[1].p, cands, (size_t)n_cands * dim * 4, cudaMemcpyHostToDevice, st));
    int mode;
    const float *d_w;
    int rc = prepare_metric(metric, m, dim, st, mode, d_w);
    if (rc) return rc;
    ProfScope p(K_DIST, st);
    int nl = launch_seed_distance(g.misc[0].as<float>(), n_seeds, g.misc[1].as<float>(), n_cands, (int)dim, mode,
                                  d_w, g.misc[2].as<float>(), st);
    nl += launch_stable_argsort(g.misc[2].as<float>(), n_cands, g.misc[3].as<unsigned long long>(),
                                g.misc[3].as<unsigned long long>() + n_cands, g.misc[4].p, tmp_bytes,
                                g.misc[5].as<unsigned int>(), st);
    p.done(nl);
    CK(cudaGetLastError());
    std::vector<float> hk;  // the keys come back either way: a NaN distance is an error, not an order
    if (!keys) { hk.resize(n_cands); keys = hk.data(); }
    CK(cudaMemcpyAsync(order, g.misc[5].p, (size_t)n_cands * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(keys, g.misc[2].p, (size_t)n_cands * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n_cands; i++)
        if (keys[i] != keys[i]) {  // the reference's n32(distance) panics on NaN (src/playlist.rs:267)
            g_last_error = "NaN distance for candidate " + std::to_string(i) + " (NaN features or metric)";
            return BLISS_B200_E_ARG;
        }
    return BLISS_B200_OK;
}

int bliss_b200_song_to_song(const float *seeds, uint32_t n_seeds, const float *cands, uint32_t n_cands,
                            uint32_t dim, int metric, const float *m, uint32_t *order) {
    REQUIRE_INIT();
    if (!seeds || !cands || !order || dim == 0 || dim > 64) { g_last_error = "bad argument"; return BLISS_B200_E_ARG; }
    if (n_cands == 0) return BLISS_B200_OK;
    cudaStream_t st = g.stream;
    CK(g.misc[0].ensure((size_t)std::max<uint32_t>(n_seeds, 1) * dim * 4));
    CK(g.misc[1].ensure((size_t)n_cands * dim * 4));
    CK(g.misc[2].ensure((size_t)n_cands));
    CK(g.misc[3].ensure((size_t)n_cands * 4));
    CK(g.misc[4].ensure((size_t)dim * 4 * 2));
    CK(cudaMemcpyAsync(g.misc[0].p, seeds, (size_t)n_seeds * dim * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.misc[1].p, cands, (size_t)n_cands * dim * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(g.misc[2].p, 1, n_cands, st));
    int mode;
    const float *d_w;
    int rc = prepare_metric(metric, m, dim, st, mode, d_w);
    if (rc) return rc;
    ProfScope p(K_DIST, st);
    int nl = 0;
    const float *cur = g.misc[0].as<float>();
    uint32_t n_cur = n_seeds;
    for (uint32_t step = 0; step < n_cands; step++) {
        float *next = g.misc[4].as<float>() + (size_t)(step & 1) * dim;
        nl += launch_nearest_alive(cur, n_cur, g.misc[1].as<float>(), n_cands, (int)dim, mode, d_w,
                                   g.misc[2].as<unsigned char>(), g.misc[3].as<unsigned int>(), step, next, st);
        cur = next;
        n_cur = 1;
    }
    p.done(nl);
    CK(cudaMemcpyAsync(order, g.misc[3].p, (size_t)n_cands * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BLISS_B200_OK;
}

// ---- profiling -----------------------------------------------------------------------------
int bliss_b200_set_profiling(int on) {
    std::lock_guard<std::mutex> lk(g.mu);
    g.profiling = on != 0;
    return BLISS_B200_OK;
}

int bliss_b200_get_profile(double *ms, uint64_t *launches) {
    REQUIRE_INIT();
    CK(cudaDeviceSynchronize());
    for (auto &pr : g.ev_pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pr.a, pr.b) == cudaSuccess) g.prof_ms[pr.kid] += (double)t;
        g.ev_pool.push_back(pr.a);
        g.ev_pool.push_back(pr.b);
    }
    g.ev_pending.clear();
    for (int k = 0; k < BLISS_B200_N_KERNELS; k++) {
        if (ms) ms[k] = g.prof_ms[k];
        if (launches) launches[k] = g.prof_launches[k];
        g.prof_ms[k] = 0.;
        g.prof_launches[k] = 0;
    }
    return BLISS_B200_OK;
}

const char *bliss_b200_kernel_name(int k) { return (k >= 0 && k < BLISS_B200_N_KERNELS) ? kKernelNames[k] : ""; }

int bliss_b200_chroma_filter(int tuning_index, double *out) {
    REQUIRE_INIT();
    if (!out || tuning_index < 0 || tuning_index > 99) { g_last_error = "tuning_index must be 0..99"; return BLISS_B200_E_ARG; }
    // device layout [idx][bin][12] (chroma.cu) -> [12][4097] as chroma_filter returns it
    std::vector<double> t((size_t)CH_BINS * 12);
    CK(cudaMemcpy(t.data(), g.t_filt.as<double>() + (size_t)tuning_index * CH_BINS * 12, t.size() * sizeof(double),
                  cudaMemcpyDeviceToHost));
    for (int b = 0; b < CH_BINS; b++)
        for (int c = 0; c < 12; c++) out[(size_t)c * CH_BINS + b] = t[(size_t)b * 12 + c];
    return BLISS_B200_OK;
}

uint64_t bliss_b200_launch_count(void) {
    uint64_t n = 0;
    for (int d = 0; d < std::max(1, g_n_ctx.load()); d++) n += (uint64_t)g_ctx[d].launches.load();
    return n;
}

}  // extern "C"

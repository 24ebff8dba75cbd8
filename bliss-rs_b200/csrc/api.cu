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
#include <vector>

#include "../../include/bliss_b200.h"
#include "common.cuh"

namespace bliss {
// launchers implemented in the other translation units
int launch_pvoc512(const float *, const SongDesc *, const unsigned int *, int, unsigned int, int, PvocTables,
                   float *, float *, float *, float *, cudaStream_t);
int launch_stft512_mags(const float *, const SongDesc *, const unsigned int *, int, unsigned int, int,
                        PvocTables, float *, cudaStream_t);
int launch_timedomain(const float *, const SongDesc *, const unsigned int *, int, unsigned int, float *,
                      float *, unsigned int *, cudaStream_t);
int launch_peakpick(const float *, const SongDesc *, const unsigned int *, int, unsigned int, float *,
                    cudaStream_t);
int launch_beattrack(const float *, const float *, const SongDesc *, int, float *, float *, unsigned int *,
                     cudaStream_t);
int launch_chroma_filter_table(double *, float *, cudaStream_t);
int launch_stft8192(const float *, const SongDesc *, const unsigned int *, int, unsigned int, const float *,
                    const cpx *, const cpx *, const cpx *, float *, double *, double *, unsigned int *,
                    cudaStream_t);
int launch_tuning(const double *, const double *, const unsigned int *, const SongDesc *, int, int *,
                  cudaStream_t);
int launch_chroma(const float *, const SongDesc *, const unsigned int *, int, unsigned int, const float *,
                  const int *, double *, double *, cudaStream_t);
int launch_finalize(const SongDesc *, int, const float *, const float *, const float *, const float *,
                    const unsigned int *, const float *, const double *, int, float *, unsigned int,
                    const PeerRows &, cudaStream_t);
int launch_gather_barrier(unsigned int *const *, int, int, unsigned int, unsigned long long, cudaStream_t);
int launch_distance_matrix(const float *, unsigned int, const float *, unsigned int, int, int, const float *,
                           float *, cudaStream_t);
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

constexpr int N_STAGE = 4;

struct Ctx {
    std::mutex mu;
    bool inited = false;
    int device = -1;
    cudaStream_t stream = nullptr, copy_stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    size_t ws_limit = 0;
    // constant tables
    DevBuf t_win512, t_twA, t_hann8k, t_tw4k, t_tw2, t_tw8k, t_filt, t_filt32;
    // wave scratch
    DevBuf blob;  // SongDesc + prefix arrays
    DevBuf mags, cand_mag, cand_pitch, cand_count, cent, roll, flat, flux, thr, loud, eb, zcr, tempo, bpm,
        bpm_count, tuning, tiles, chroma_dbg;
    // host-API staging
    DevBuf pcm[4], feats, metric, misc[6];
    // pinned staging ring for descriptor uploads
    void *stage[N_STAGE] = {nullptr, nullptr, nullptr, nullptr};
    size_t stage_cap[N_STAGE] = {0, 0, 0, 0};
    cudaEvent_t stage_ev[N_STAGE] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_used[N_STAGE] = {false, false, false, false};
    int stage_next = 0;
    // profiling
    bool profiling = false;
    struct EvPair { cudaEvent_t a, b; int kid; };
    std::vector<EvPair> ev_pending;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[BLISS_B200_N_KERNELS] = {0};
    unsigned long long prof_launches[BLISS_B200_N_KERNELS] = {0};
    std::atomic<unsigned long long> launches{0};
};

Ctx g;

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
    w.pairs_per_item = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(8, r));
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

int upload_plan(const WavePlan &w, cudaStream_t st, WaveDev &out) {
    const size_t n = w.sd.size();
    const size_t sd_bytes = align_up(n * sizeof(SongDesc), 256);
    const size_t pf_bytes = align_up((n + 1) * sizeof(uint32_t), 256);
    const size_t total = sd_bytes + 5 * pf_bytes;
    CK(g.blob.ensure(total * N_STAGE));
    const int slot = g.stage_next;
    g.stage_next = (g.stage_next + 1) % N_STAGE;
    if (g.stage_used[slot]) CK(cudaEventSynchronize(g.stage_ev[slot]));
    if (g.stage_cap[slot] < total) {
        if (g.stage[slot]) cudaFreeHost(g.stage[slot]);
        g.stage[slot] = nullptr;
        CK(cudaMallocHost(&g.stage[slot], total + total / 4));
        g.stage_cap[slot] = total + total / 4;
    }
    if (!g.stage_ev[slot]) CK(cudaEventCreateWithFlags(&g.stage_ev[slot], cudaEventDisableTiming));
    char *h = (char *)g.stage[slot];
    memcpy(h, w.sd.data(), n * sizeof(SongDesc));
    const std::vector<uint32_t> *pf[5] = {&w.k1_prefix, &w.chunk_prefix, &w.t_prefix, &w.pair_prefix, &w.tile_prefix};
    for (int i = 0; i < 5; i++) memcpy(h + sd_bytes + i * pf_bytes, pf[i]->data(), (n + 1) * sizeof(uint32_t));
    // each ring slot owns its own region of the device blob, so a wave still executing
    // never sees the next wave's descriptors
    char *d = g.blob.as<char>() + (size_t)slot * (g.blob.cap / N_STAGE / 256 * 256);
    if ((size_t)(g.blob.cap / N_STAGE / 256 * 256) < total) { g_last_error = "descriptor blob too small"; return BLISS_B200_E_CUDA; }
    CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(g.stage_ev[slot], st));
    g.stage_used[slot] = true;
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

// one wave of the full analysis: every kernel of the path, enqueued on `st`
int run_wave(const float *d_pcm, const WavePlan &w, int version, float *d_out, uint32_t out_base,
             cudaStream_t st, bool debug, const PeerRows &peers) {
    const int n = (int)w.sd.size();
    if (n == 0) return BLISS_B200_OK;
    CK(g.mags.ensure(std::max<size_t>(w.rows, 2) * CH_STRIDE * sizeof(float)));
    CK(g.cand_mag.ensure(std::max<size_t>(w.cands, 1) * sizeof(double)));
    CK(g.cand_pitch.ensure(std::max<size_t>(w.cands, 1) * sizeof(double)));  // interpolated pitches (f64)
    CK(g.cand_count.ensure((size_t)n * 4));
    CK(g.cent.ensure(std::max<size_t>(w.n_s, 1) * 4));
    CK(g.roll.ensure(std::max<size_t>(w.n_s, 1) * 4));
    CK(g.flat.ensure(std::max<size_t>(w.n_s, 1) * 4));
    CK(g.flux.ensure(std::max<size_t>(w.n_t, 1) * 4));
    CK(g.thr.ensure(std::max<size_t>(w.n_t, 1) * 4));
    CK(g.loud.ensure(std::max<size_t>(w.n_l, 1) * 4));
    CK(g.eb.ensure(std::max<size_t>(w.n_eb, 1) * 4));
    CK(g.zcr.ensure((size_t)n * 4));
    CK(g.tempo.ensure((size_t)n * 4));
    CK(g.bpm.ensure(std::max<size_t>(w.bpms, 1) * 4));
    CK(g.bpm_count.ensure((size_t)n * 4));
    CK(g.tuning.ensure((size_t)n * 4));
    CK(g.tiles.ensure(std::max<size_t>(w.tiles, 1) * 10 * sizeof(double)));
    if (debug) CK(g.chroma_dbg.ensure(std::max<size_t>(w.tiles, 1) * CH_TILE_FRAMES * 12 * sizeof(double)));
    WaveDev dv;
    int rc = upload_plan(w, st, dv);
    if (rc) return rc;
    CK(cudaMemsetAsync(g.zcr.p, 0, (size_t)n * 4, st));
    CK(cudaMemsetAsync(g.cand_count.p, 0, (size_t)n * 4, st));

    // Two independent chains per wave: the tempo/timbral chain stays on the caller's stream, the chroma
    // chain runs on a side stream so that its latency-bound kernels (tuning) overlap the other chain's
    // compute-bound ones and vice versa (beat tracker under the chroma STFT).  Joined before finalize.
    // (while per-kernel profiling is on, both chains are serialised on `st` so that each kernel's
    // CUDA-event duration is its own and not inflated by the kernel it would overlap with)
    cudaStream_t sb = g.profiling ? st : g.side_stream;
    CK(cudaEventRecord(g.ev_fork, st));
    CK(cudaStreamWaitEvent(sb, g.ev_fork, 0));
    { ProfScope p(K_STFT8K, sb);
      p.done(launch_stft8192(d_pcm, dv.sd, dv.pair_prefix, n, w.pair_prefix[n], g.t_hann8k.as<float>(),
                             g.t_tw4k.as<cpx>(), g.t_tw2.as<cpx>(), g.t_tw8k.as<cpx>(), g.mags.as<float>(), g.cand_mag.as<double>(),
                             g.cand_pitch.as<double>(), g.cand_count.as<unsigned int>(), sb)); }
    { ProfScope p(K_TIME, st);
      p.done(launch_timedomain(d_pcm, dv.sd, dv.chunk_prefix, n, w.chunk_prefix[n], g.loud.as<float>(),
                               g.eb.as<float>(), g.zcr.as<unsigned int>(), st)); }
    { ProfScope p(K_PVOC, st);
      p.done(launch_pvoc512(d_pcm, dv.sd, dv.k1_prefix, n, w.k1_prefix[n], (int)w.pairs_per_item, pvoc_tables(),
                            g.cent.as<float>(), g.roll.as<float>(), g.flat.as<float>(), g.flux.as<float>(), st)); }
    { ProfScope p(K_TUNING, sb);
      p.done(launch_tuning(g.cand_mag.as<double>(), g.cand_pitch.as<double>(),
                           g.cand_count.as<unsigned int>(), dv.sd, n, g.tuning.as<int>(), sb)); }
    { ProfScope p(K_PEAK, st);
      p.done(launch_peakpick(g.flux.as<float>(), dv.sd, dv.t_prefix, n, w.t_prefix[n], g.thr.as<float>(), st)); }
    { ProfScope p(K_CHROMA, sb);
      p.done(launch_chroma(g.mags.as<float>(), dv.sd, dv.tile_prefix, n, w.tile_prefix[n], g.t_filt32.as<float>(),
                           g.tuning.as<int>(), g.tiles.as<double>(), debug ? g.chroma_dbg.as<double>() : nullptr, sb)); }
    { ProfScope p(K_BEAT, st);
      p.done(launch_beattrack(g.thr.as<float>(), g.eb.as<float>(), dv.sd, n, g.bpm.as<float>(),
                              g.tempo.as<float>(), g.bpm_count.as<unsigned int>(), st)); }
    CK(cudaEventRecord(g.ev_join, sb));
    CK(cudaStreamWaitEvent(st, g.ev_join, 0));
    { ProfScope p(K_FINAL, st);
      p.done(launch_finalize(dv.sd, n, g.cent.as<float>(), g.roll.as<float>(), g.flat.as<float>(),
                             g.loud.as<float>(), g.zcr.as<unsigned int>(), g.tempo.as<float>(),
                             g.tiles.as<double>(), version, d_out, out_base, peers, st)); }
    CK(cudaGetLastError());
    return BLISS_B200_OK;
}

// split [0, n_songs) into waves that respect the workspace limit, run them in order
int analyze_device_locked(const float *d_pcm, const uint64_t *offsets, const uint64_t *n_samples,
                          uint32_t n_songs, int version, float *d_out, int32_t *status, cudaStream_t st,
                          bool debug, const PeerRows *peers = nullptr) {
    PeerRows no_peers;
    memset(&no_peers, 0, sizeof(no_peers));
    if (((uintptr_t)d_pcm & 15u) != 0) { g_last_error = "d_pcm must be 16-byte aligned"; return BLISS_B200_E_ARG; }
    uint32_t first = 0;
    WavePlan w;
    while (first < n_songs) {
        size_t bytes = 0;
        uint32_t count = 0;
        while (first + count < n_songs) {
            const size_t b = geom_of(n_samples[first + count]).scratch_bytes + 128;
            if (count > 0 && bytes + b > g.ws_limit) break;
            bytes += b;
            count++;
        }
        plan_wave(offsets, n_samples, first, count, false, w);
        int rc = run_wave(d_pcm, w, version, d_out, first, st, debug, peers ? *peers : no_peers);
        if (rc) return rc;
        first += count;
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
    std::vector<float> hann(8192);
    for (int i = 0; i < 8192; i++) hann[i] = 0.5f - 0.5f * cosf(2.f * (float)i * PI_F / 8192.f);
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
    if (!m) return 0;
    bool diag = true;
    for (uint32_t i = 0; i < dim && diag; i++)
        for (uint32_t j = 0; j < dim; j++)
            if (i != j && m[(size_t)i * dim + j] != 0.f) { diag = false; break; }
    CK(g.metric.ensure((size_t)dim * dim * 4 + 256));
    if (diag) {
        std::vector<float> w(dim);
        for (uint32_t i = 0; i < dim; i++) w[i] = m[(size_t)i * dim + i];
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
        default: return "unknown";
    }
}

const char *bliss_b200_last_error(void) { return g_last_error.c_str(); }

uint32_t bliss_b200_feature_count(uint16_t v) { return v == 2 ? 23u : v == 1 ? 20u : 0u; }

int bliss_b200_init(int device) {
    std::lock_guard<std::mutex> lk(g.mu);
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
    CK(cudaStreamCreateWithFlags(&g.side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    g.ws_limit = (size_t)((double)total_b * 0.40);
    int rc = build_tables();
    if (rc) return rc;
    g.inited = true;
    return BLISS_B200_OK;
}

void bliss_b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.inited) return;
    cudaSetDevice(g.device);
    cudaDeviceSynchronize();
    DevBuf *all[] = {&g.t_win512, &g.t_twA, &g.t_hann8k, &g.t_tw4k, &g.t_tw2, &g.t_tw8k, &g.t_filt, &g.t_filt32, &g.blob, &g.mags, &g.cand_mag,
                     &g.cand_pitch, &g.cand_count, &g.cent, &g.roll, &g.flat, &g.flux, &g.thr, &g.loud, &g.eb,
                     &g.zcr, &g.tempo, &g.bpm, &g.bpm_count, &g.tuning, &g.tiles, &g.chroma_dbg, &g.pcm[0],
                     &g.pcm[1], &g.pcm[2], &g.pcm[3], &g.feats, &g.metric, &g.misc[0], &g.misc[1], &g.misc[2], &g.misc[3],
                     &g.misc[4], &g.misc[5]};
    for (DevBuf *b : all) b->release();
    for (int i = 0; i < N_STAGE; i++) {
        if (g.stage[i]) cudaFreeHost(g.stage[i]);
        g.stage[i] = nullptr;
        g.stage_cap[i] = 0;
        if (g.stage_ev[i]) cudaEventDestroy(g.stage_ev[i]);
        g.stage_ev[i] = nullptr;
        g.stage_used[i] = false;
    }
    for (auto &p : g.ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    g.ev_pending.clear();
    for (auto e : g.ev_pool) cudaEventDestroy(e);
    g.ev_pool.clear();
    cudaStreamDestroy(g.stream);
    cudaStreamDestroy(g.copy_stream);
    cudaStreamDestroy(g.side_stream);
    cudaEventDestroy(g.ev_fork);
    cudaEventDestroy(g.ev_join);
    g.inited = false;
}

int bliss_b200_set_workspace_limit(uint64_t bytes) {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.inited) return BLISS_B200_E_NOT_INIT;
    g.ws_limit = (size_t)bytes;
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

// host buffers: chunks of songs are copied on a side stream while the previous chunk computes
static int analyze_host_locked(const float *const *pcm, const uint64_t *n_samples, uint32_t n_songs,
                               uint16_t ver, float *out, int32_t *status, bool debug) {
    const uint32_t dim = bliss_b200_feature_count(ver);
    CK(g.feats.ensure((size_t)n_songs * dim * 4));
    // The path is PCIe-bound (15.9 MB per 3-min song): chunks of songs are copied on a side stream while
    // earlier chunks compute.  A chunk's kernels have ~3.5 ms of latency whatever its size (tuning, beat
    // tracker), more than a small chunk's copy time, so the ring holds FOUR device PCM buffers: the
    // copy engine can run three chunks ahead of the compute stream instead of idling (measured with
    // BLISS_B200_TRACE: two buffers of 128 MB kept the link at 36 of 55 GB/s).
    // BLISS_B200_CHUNK_MB overrides the chunk size for experiments.
    // Measured (BLISS_B200_TRACE, 4 GB batch): 128 MB chunks -> 37 GB/s (compute-latency bound: 32 chunks x
    // 3.5 ms), 256 MB -> 45 GB/s, 512 MB -> 50 GB/s of a 55.6 GB/s link; default = an eighth of the batch.
    size_t total_mb = 0;
    for (uint32_t i = 0; i < n_songs; i++) total_mb += (size_t)n_samples[i] * 4 >> 20;
    size_t chunk_mb = std::min<size_t>(1024, std::max<size_t>(256, total_mb / 8));
    const size_t chunk_budget = std::min<size_t>(chunk_mb << 20, std::max<size_t>(g.ws_limit / 8, (size_t)64 << 20));
    const bool trace = getenv("BLISS_B200_TRACE") != nullptr;
    cudaEvent_t tr[4] = {nullptr, nullptr, nullptr, nullptr};  // copy begin/end, compute begin/end
    if (trace) {
        for (auto &e : tr) CK(cudaEventCreate(&e));
        CK(cudaEventRecord(tr[0], g.copy_stream));
        CK(cudaEventRecord(tr[2], g.stream));
    }
    size_t done_bytes = 0;
    constexpr int NBUF = 4;
    cudaEvent_t ev_copy[NBUF], ev_done[NBUF];
    for (int i = 0; i < NBUF; i++) {
        CK(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    }
    bool used[NBUF] = {false, false, false, false};
    std::vector<uint64_t> offs, lens;
    uint32_t first = 0;
    int c = 0, rc = BLISS_B200_OK;
    while (first < n_songs && rc == BLISS_B200_OK) {
        // chunk = as many songs as fit the PCM budget (at least one)
        size_t samples = 0;
        uint32_t count = 0;
        offs.clear();
        lens.clear();
        while (first + count < n_songs) {
            const size_t len = align_up((size_t)n_samples[first + count], 4);
            if (count > 0 && (samples + len) * 4 > chunk_budget) break;
            offs.push_back(samples);
            lens.push_back(n_samples[first + count]);
            samples += len;
            count++;
        }
        const int b = c % NBUF;
        if (used[b]) CK(cudaEventSynchronize(ev_done[b]));  // buffer b free again (also: host may realloc)
        CK(g.pcm[b].ensure(std::max<size_t>(samples, 4) * 4));
        for (uint32_t i = 0; i < count;) {
            if (lens[i] == 0) { i++; continue; }
            if (!pcm[first + i]) { g_last_error = "null pcm pointer"; rc = BLISS_B200_E_ARG; break; }
            // songs that sit back to back in host memory (one big decoded buffer) go as ONE copy
            uint32_t j = i;
            size_t run = (size_t)lens[i];
            while (j + 1 < count && lens[j + 1] > 0 && (lens[j] & 3u) == 0 &&
                   pcm[first + j + 1] == pcm[first + j] + lens[j]) {
                j++;
                run += (size_t)lens[j];
            }
            CK(cudaMemcpyAsync(g.pcm[b].as<float>() + offs[i], pcm[first + i], run * 4, cudaMemcpyHostToDevice,
                               g.copy_stream));
            i = j + 1;
        }
        if (rc) break;
        CK(cudaEventRecord(ev_copy[b], g.copy_stream));
        CK(cudaStreamWaitEvent(g.stream, ev_copy[b], 0));
        rc = analyze_device_locked(g.pcm[b].as<float>(), offs.data(), lens.data(), count, ver,
                                   g.feats.as<float>() + (size_t)first * dim, status ? status + first : nullptr,
                                   g.stream, debug);
        CK(cudaEventRecord(ev_done[b], g.stream));
        used[b] = true;
        first += count;
        done_bytes += samples * 4;
        c++;
    }
    if (rc == BLISS_B200_OK) {
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
                            "compute_stream=%.2f ms first_copy->last_kernel=%.2f ms\n",
                    n_songs, done_bytes / 1e6, c, chunk_mb, ms_copy, done_bytes / 1e6 / ms_copy, ms_comp, ms_all);
            for (auto &e : tr) cudaEventDestroy(e);
        }
    } else {
        cudaStreamSynchronize(g.stream);
        cudaStreamSynchronize(g.copy_stream);
    }
    for (int i = 0; i < NBUF; i++) { cudaEventDestroy(ev_copy[i]); cudaEventDestroy(ev_done[i]); }
    return rc;
}

int bliss_b200_analyze_batch(const float *const *pcm, const uint64_t *n_samples, uint32_t n_songs,
                             uint16_t ver, float *out, int32_t *status) {
    REQUIRE_INIT();
    if (check_version(ver)) return BLISS_B200_E_ARG;
    if (n_songs == 0) return BLISS_B200_OK;
    if (!pcm || !n_samples || !out) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
    return analyze_host_locked(pcm, n_samples, n_songs, ver, out, status, false);
}

int bliss_b200_analyze(const float *pcm, uint64_t n, uint16_t ver, float *out) {
    int32_t status = 0;
    const float *p[1] = {pcm};
    uint64_t len[1] = {n};
    {
        REQUIRE_INIT();
        if (check_version(ver)) return BLISS_B200_E_ARG;
        if (!out || (!pcm && n > 0)) { g_last_error = "null pointer"; return BLISS_B200_E_ARG; }
        int rc = analyze_host_locked(p, len, 1, ver, out, &status, false);
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
    int rc = analyze_host_locked(p, len, 1, ver, out, &status, true);
    if (rc) return rc;
    if (status != 0 || !t) return status;
    const SongGeom q = geom_of(n);
    auto dl = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        return dst ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : cudaSuccess;
    };
    CK(dl(t->centroid, g.cent.p, (size_t)q.n_s * 4));
    CK(dl(t->rolloff, g.roll.p, (size_t)q.n_s * 4));
    CK(dl(t->flatness, g.flat.p, (size_t)q.n_s * 4));
    CK(dl(t->flux, g.flux.p, (size_t)q.n_t * 4));
    CK(dl(t->thresholded, g.thr.p, (size_t)q.n_t * 4));
    uint32_t nb = 0;
    CK(cudaMemcpy(&nb, g.bpm_count.p, 4, cudaMemcpyDeviceToHost));
    if (t->n_bpms) *t->n_bpms = nb;
    CK(dl(t->bpms, g.bpm.p, (size_t)nb * 4));
    CK(dl(t->loudness_chunks, g.loud.p, (size_t)q.n_l * 4));
    CK(dl(t->zero_crossings, g.zcr.p, 4));
    if (t->stft8192) {
        CK(cudaMemcpy2D(t->stft8192, (size_t)CH_BINS * 4, g.mags.p, (size_t)CH_STRIDE * 4, (size_t)CH_BINS * 4,
                        q.n_c_comp, cudaMemcpyDeviceToHost));
        for (uint32_t f = q.n_c_comp; f < q.n_c; f++) memset(t->stft8192 + (size_t)f * CH_BINS, 0, (size_t)CH_BINS * 4);
    }
    if (t->n_peaks) {
        uint32_t c = 0;
        CK(cudaMemcpy(&c, g.cand_count.p, 4, cudaMemcpyDeviceToHost));
        *t->n_peaks = c;
    }
    if (t->tuning) {
        int idx = 0;
        CK(cudaMemcpy(&idx, g.tuning.p, 4, cudaMemcpyDeviceToHost));
        *t->tuning = (-50. + (100. * 0.01 * (double)idx)) / 100.;
    }
    CK(dl(t->chroma, g.chroma_dbg.p, (size_t)q.n_c * 12 * sizeof(double)));
    if (t->interval_features) {
        std::vector<double> parts((size_t)q.n_tiles * 10);
        CK(cudaMemcpy(parts.data(), g.tiles.p, parts.size() * sizeof(double), cudaMemcpyDeviceToHost));
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
    int rc = upload_plan(w, st, dv);
    if (rc) return rc;
    ProfScope p(K_STFT512, st);
    p.done(launch_stft512_mags(d_pcm, dv.sd, dv.k1_prefix, (int)n_songs, w.k1_prefix[n_songs],
                               (int)w.pairs_per_item, pvoc_tables(), d_mags, st));
    CK(cudaGetLastError());
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
    const int nl = launch_distance_matrix(d_rows, n_rows, d_cols, n_cols, (int)dim, mode, d_w, d_out, st);
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
                                          mode, d_w, g.misc[2].as<float>(), st);
    p.done(nl);
    if (nl < 0) { g_last_error = "unsupported dim"; return BLISS_B200_E_ARG; }
    CK(cudaMemcpyAsync(out, g.misc[2].p, (size_t)n_rows * n_cols * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BLISS_B200_OK;
}

int bliss_b200_distance_matrix(const float *rows, uint32_t n_rows, const float *cols, uint32_t n_cols,
                               uint32_t dim, int metric, const float *m, float *out) {
    REQUIRE_INIT();
    if (!rows || !cols || !out || dim == 0 || dim > 64) { g_last_error = "bad argument"; return BLISS_B200_E_ARG; }
    if (n_rows == 0 || n_cols == 0) return BLISS_B200_OK;
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
    CK(cudaMemcpyAsync(order, g.misc[5].p, (size_t)n_cands * 4, cudaMemcpyDeviceToHost, st));
    if (keys) CK(cudaMemcpyAsync(keys, g.misc[2].p, (size_t)n_cands * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
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

uint64_t bliss_b200_launch_count(void) { return (uint64_t)g.launches.load(); }

}  // extern "C"

// Decoder::analyze_paths of include/bliss_b200.hpp under ThreadSanitizer.  The C ABI is stubbed IN THIS FILE (test
// infrastructure: rows = the first sample of each buffer; a call takes a moment, and a call made while another one is
// in flight is reported), so that the threading of the host mirror itself -- decoding threads, the bounded hand-over,
// the batcher, the error paths -- runs thousands of hand-overs without a device.  Built with -fsanitize=thread and run
// by tests/test_host_abi.py; prints OK.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <set>
#include <thread>

#include "bliss_b200.hpp"

static std::atomic<int> g_in_call{0}, g_overlapping_calls{0}, g_calls{0}, g_max_batch{0};
static std::atomic<bool> g_fail_calls{false};

extern "C" {
int bliss_b200_init(int) { return BLISS_B200_OK; }
int bliss_b200_init_devices(int) { return 1; }
const char *bliss_b200_strerror(int) { return "stub"; }
const char *bliss_b200_last_error(void) { return "stub failure"; }
int bliss_b200_analyze_batch_pcm(const void *const *, const uint64_t *, uint32_t, int, uint32_t, uint32_t, uint16_t, float *, int32_t *) {
    return BLISS_B200_E_ARG;  // not used by this program (no song carries packed frames)
}
int bliss_b200_pcm_to_mono(const void *, uint64_t, int, uint32_t, float *) { return BLISS_B200_E_ARG; }  // likewise
uint64_t bliss_b200_resampled_len(uint64_t n, uint32_t) { return n; }
int bliss_b200_resample(const float *, uint64_t, uint32_t, float *, uint64_t, uint64_t *) { return BLISS_B200_E_ARG; }
uint32_t bliss_b200_feature_count(uint16_t v) { return v == 2 ? 23u : v == 1 ? 20u : 0u; }
int bliss_b200_analyze_batch(const float *const *pcm, const uint64_t *n, uint32_t n_songs, uint16_t ver, float *out, int32_t *status) {
    if (g_in_call.fetch_add(1) != 0) g_overlapping_calls++;
    g_calls++;
    int seen = g_max_batch.load();
    while ((int)n_songs > seen && !g_max_batch.compare_exchange_weak(seen, (int)n_songs)) {}
    std::this_thread::sleep_for(std::chrono::microseconds(300));
    const uint32_t dim = bliss_b200_feature_count(ver);
    for (uint32_t i = 0; i < n_songs; i++) {
        status[i] = n[i] < 8192 ? BLISS_B200_SONG_TOO_SHORT : BLISS_B200_SONG_OK;
        for (uint32_t k = 0; k < dim; k++) out[i * dim + k] = n[i] ? pcm[i][0] : 0.f;
    }
    g_in_call--;
    return g_fail_calls ? BLISS_B200_E_CUDA : BLISS_B200_OK;
}
}

using namespace bliss;

struct NumberDecoder : Decoder {  // "files" are numbers; "bad*" fail to decode, "short*" are too short, "bug" is a bug
    std::atomic<int> now{0}, peak{0};
    PreAnalyzedSong decode(const std::string &path) override {
        const int n = ++now;
        int p = peak.load();
        while (n > p && !peak.compare_exchange_weak(p, n)) {}
        std::this_thread::sleep_for(std::chrono::microseconds(100));
        now--;
        if (path.rfind("bad", 0) == 0) throw BlissError(BlissError::DecodingError, path);
        if (path == "bug") throw std::logic_error("decode() is broken");
        PreAnalyzedSong s;
        s.path = path;
        const bool is_short = path.rfind("short", 0) == 0;
        s.sample_array.assign(is_short ? 100 : 9000, is_short ? 0.f : (float)std::stoi(path));
        return s;
    }
};

int main() {
    NumberDecoder dec;
    std::vector<std::string> paths;
    for (int i = 0; i < 3000; i++) paths.push_back(i % 97 == 5 ? "bad" + std::to_string(i) : i % 89 == 7 ? "short" + std::to_string(i) : std::to_string(i));
    for (unsigned cores : {1u, 3u, 8u}) {
        for (size_t batch : {size_t(1), size_t(7), size_t(64)}) {
            AnalysisOptions o;
            o.number_cores = cores;
            g_max_batch = 0;
            const auto res = dec.analyze_paths(paths, o, batch);
            if (res.size() != paths.size()) return 2;
            std::multiset<std::string> got, want(paths.begin(), paths.end());
            for (const auto &r : res) {
                got.insert(r.first);
                if (const Song *s = std::get_if<Song>(&r.second)) {
                    if (s->path != r.first || s->analysis->as_vec()[22] != (float)std::stoi(r.first)) return 3;  // rows stay with their songs
                } else {
                    const BlissError &e = std::get<BlissError>(r.second);
                    const bool bad = r.first.rfind("bad", 0) == 0, is_short = r.first.rfind("short", 0) == 0;
                    if (!(bad && e.kind == BlissError::DecodingError) && !(is_short && e.kind == BlissError::AnalysisError)) return 4;
                }
            }
            if (got != want) return 5;
            if (g_max_batch > (int)batch) return 6;
        }
    }
    if (g_overlapping_calls != 0) return 7;  // one batcher: GPU calls never overlap
    if (std::thread::hardware_concurrency() >= 2 && dec.peak < 2) return 8;
    {   // a bug in decode() reaches the caller after the workers have stopped
        std::vector<std::string> p2(paths.begin(), paths.begin() + 500);
        p2[250] = "bug";
        AnalysisOptions o;
        o.number_cores = 4;
        try {
            dec.analyze_paths(p2, o, 16);
            return 9;
        } catch (const std::logic_error &) {
        }
        if (dec.now != 0) return 10;
    }
    {   // a failing GPU call (the whole call, not a song) reaches the caller as well
        g_fail_calls = true;
        AnalysisOptions o;
        o.number_cores = 4;
        try {
            dec.analyze_paths(paths, o, 16);
            return 11;
        } catch (const BlissError &e) {
            if (std::string(e.what()).find("stub failure") == std::string::npos) return 12;
        }
        g_fail_calls = false;
        if (dec.now != 0) return 13;
    }
    std::printf("OK %d calls\n", g_calls.load());
    return 0;
}

// bliss_b200.hpp -- C++17 header-only host mirror of the reference crate's API for the
// Song::analyze + distance path, on top of the C ABI in bliss_b200.h.  The reference is compiled
// code (Rust); no Rust toolchain exists in the build image, so this is the compiled-language host
// side a native application links against (the Rust binding itself is in INTEGRATION.md).
//
// Names / argument meaning / error behaviour follow:
//   FeaturesVersion, BlissError            src/lib.rs:151-249
//   Analysis, AnalysisOptions, Song        src/song/mod.rs:45-521
//   PreAnalyzedSong, Decoder               src/song/decoder.rs:34-333
//   distances, closest_to_songs, ...       src/playlist.rs:65-326
#pragma once
#include <algorithm>
#include <cctype>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <variant>
#include <vector>

#include "bliss_b200.h"

namespace bliss {

constexpr uint32_t SAMPLE_RATE = 22050;  // src/lib.rs:143
constexpr uint16_t CHANNELS = 1;         // src/lib.rs:140
constexpr size_t NUMBER_FEATURES = 23;   // AnalysisIndex::COUNT

// src/lib.rs:236-249
struct BlissError : std::runtime_error {
    enum Kind { DecodingError, AnalysisError, ProviderError } kind;
    BlissError(Kind k, const std::string &msg) : std::runtime_error(prefix(k) + msg), kind(k) {}
    static std::string prefix(Kind k) {
        return k == DecodingError   ? "error happened while decoding file - "
               : k == AnalysisError ? "error happened while analyzing file - "
                                    : "error happened with the music library provider - ";
    }
};

enum class FeaturesVersion : uint16_t { Version1 = 1, Version2 = 2 };  // src/lib.rs:151-166
constexpr FeaturesVersion LATEST = FeaturesVersion::Version2;
inline size_t feature_count(FeaturesVersion v) { return bliss_b200_feature_count(static_cast<uint16_t>(v)); }
inline std::vector<float> feature_weights(FeaturesVersion v) {  // src/lib.rs:168-173
    std::vector<float> m(feature_count(v) * feature_count(v));
    bliss_b200_feature_weights(static_cast<uint16_t>(v), m.data());
    return m;
}

namespace detail {
// BLISS_B200_DEVICE=<n> pins the process to one GPU; otherwise every visible B200 is opened and the host-buffer calls
// (analyze_batch, the decoder pipeline) shard one call's songs over all of them inside the library
// (bliss_b200_init_devices): the crate is one process, src/song/decoder.rs:282-331.
inline void ensure_init(int device = -1) {
    static const int rc = [device] {
        int dev = device;
        if (const char *e = std::getenv("BLISS_B200_DEVICE")) dev = std::atoi(e);
        if (dev >= 0) return bliss_b200_init(dev);
        const int n = bliss_b200_init_devices(0);
        return n > 0 ? BLISS_B200_OK : n;
    }();
    if (rc != BLISS_B200_OK)
        throw BlissError(BlissError::AnalysisError, std::string("b200 backend unavailable (no CPU fallback): ") +
                                                        bliss_b200_strerror(rc) + ": " + bliss_b200_last_error());
}
inline void check_call(int rc) {
    if (rc < 0)
        throw BlissError(BlissError::AnalysisError,
                         std::string("b200 backend: ") + bliss_b200_strerror(rc) + ": " + bliss_b200_last_error());
}
inline BlissError status_error(int status) {
    return status == BLISS_B200_SONG_TOO_SHORT
               ? BlissError(BlissError::AnalysisError, "empty or too short song.")  // src/song/mod.rs:426-430
               : BlissError(BlissError::AnalysisError, "b200 backend: internal error");
}
}  // namespace detail

// src/song/mod.rs:103-156
enum class AnalysisIndex : size_t {
    Tempo, Zcr, MeanSpectralCentroid, StdDeviationSpectralCentroid, MeanSpectralRolloff,
    StdDeviationSpectralRolloff, MeanSpectralFlatness, StdDeviationSpectralFlatness, MeanLoudness,
    StdDeviationLoudness, Chroma1, Chroma2, Chroma3, Chroma4, Chroma5, Chroma6, Chroma7, Chroma8, Chroma9,
    Chroma10, Chroma11, Chroma12, Chroma13
};

// src/song/mod.rs:252-269
struct AnalysisOptions {
    FeaturesVersion features_version = LATEST;
    unsigned number_cores = std::max(1u, std::thread::hardware_concurrency());
};

// src/song/mod.rs:240-371
class Analysis {
  public:
    FeaturesVersion features_version;
    Analysis(std::vector<float> analysis, FeaturesVersion v) : features_version(v), v_(std::move(analysis)) {
        if (v_.size() != feature_count(v))
            throw BlissError(BlissError::ProviderError,
                             "Feature count " + std::to_string(v_.size()) +
                                 " does not match the expected version feature count " + std::to_string(feature_count(v)));
    }
    const std::vector<float> &as_vec() const { return v_; }
    float operator[](AnalysisIndex i) const {
        if (features_version != LATEST) throw std::logic_error("Tried to index features with incompatible indexes");
        return v_[static_cast<size_t>(i)];
    }
    // src/song/mod.rs:364-370
    float distance(const Analysis &other) const {
        if (features_version != other.features_version)
            throw std::logic_error("Mismatched features version between two songs or analysis");
        detail::ensure_init();
        const auto m = feature_weights(features_version);
        float d = 0.f;
        detail::check_call(bliss_b200_distance(v_.data(), other.v_.data(), static_cast<uint32_t>(v_.size()),
                                               BLISS_B200_METRIC_MAHALANOBIS, m.data(), &d));
        return d;
    }

  private:
    std::vector<float> v_;
};

// src/song/mod.rs:45-98 (fields the analysis path carries)
struct Song {
    std::string path;
    std::optional<std::string> artist, album_artist, title, album, genre;
    std::optional<int> track_number, disc_number;
    double duration_s = 0.;
    std::optional<Analysis> analysis;
    FeaturesVersion features_version = LATEST;
    struct CueInfo { std::string cue_path, audio_file_path; };  // src/cue.rs: CueInfo
    std::optional<CueInfo> cue_info;                            // set when the song was cut out of a CUE sheet

    // Song::analyze / analyze_with_options, src/song/mod.rs:403-508
    static Analysis analyze(const float *samples, size_t n) { return analyze_with_options(samples, n, AnalysisOptions{}); }
    static Analysis analyze_with_options(const float *samples, size_t n, const AnalysisOptions &o) {
        detail::ensure_init();
        std::vector<float> out(feature_count(o.features_version));
        const int rc = bliss_b200_analyze(samples, n, static_cast<uint16_t>(o.features_version), out.data());
        detail::check_call(rc);
        if (rc != BLISS_B200_SONG_OK) throw detail::status_error(rc);
        return Analysis(std::move(out), o.features_version);
    }
    float distance(const Song &other) const { return analysis->distance(*other.analysis); }  // :519-521
};

using AnalysisResult = std::variant<Analysis, BlissError>;

// The batching seam: many decoded buffers, one GPU call; errors are items (src/song/decoder.rs:319-325).
inline std::vector<AnalysisResult> analyze_batch(const std::vector<const float *> &pcm, const std::vector<uint64_t> &n,
                                                 const AnalysisOptions &o = {}) {
    detail::ensure_init();
    const size_t dim = feature_count(o.features_version);
    std::vector<float> out(dim * pcm.size());
    std::vector<int32_t> status(pcm.size());
    detail::check_call(bliss_b200_analyze_batch(pcm.data(), n.data(), static_cast<uint32_t>(pcm.size()),
                                                static_cast<uint16_t>(o.features_version), out.data(), status.data()));
    std::vector<AnalysisResult> res;
    res.reserve(pcm.size());
    for (size_t i = 0; i < pcm.size(); i++) {
        if (status[i] == BLISS_B200_SONG_OK)
            res.emplace_back(Analysis(std::vector<float>(out.begin() + i * dim, out.begin() + (i + 1) * dim), o.features_version));
        else
            res.emplace_back(detail::status_error(status[i]));
    }
    return res;
}

// Same for 16-bit sources (signed 16-bit mono 22 050 Hz as decoded, before the resampler's s16 -> flt step):
// converted on the device, half the PCIe bytes, results bit-identical to analyze_batch on x / 32768.
inline std::vector<AnalysisResult> analyze_batch_s16(const std::vector<const int16_t *> &pcm,
                                                     const std::vector<uint64_t> &n, const AnalysisOptions &o = {}) {
    detail::ensure_init();
    const size_t dim = feature_count(o.features_version);
    std::vector<float> out(dim * pcm.size());
    std::vector<int32_t> status(pcm.size());
    detail::check_call(bliss_b200_analyze_batch_s16(pcm.data(), n.data(), static_cast<uint32_t>(pcm.size()),
                                                    static_cast<uint16_t>(o.features_version), out.data(), status.data()));
    std::vector<AnalysisResult> res;
    res.reserve(pcm.size());
    for (size_t i = 0; i < pcm.size(); i++) {
        if (status[i] == BLISS_B200_SONG_OK)
            res.emplace_back(Analysis(std::vector<float>(out.begin() + i * dim, out.begin() + (i + 1) * dim), o.features_version));
        else
            res.emplace_back(detail::status_error(status[i]));
    }
    return res;
}

// Decode-side feed, general form (bliss_b200_analyze_batch_pcm): interleaved frames of one sample format and
// channel count at 22 050 Hz; the decoders' sample-format conversion and down-mix (src/song/decoder/ffmpeg.rs:36-109,
// symphonia.rs:260-300) run on the device.  A sample rate other than 22 050 Hz throws: resampling stays on the
// decoder's side of the boundary.
enum class PcmFormat : int { S16 = BLISS_B200_PCM_S16, S32 = BLISS_B200_PCM_S32, F32 = BLISS_B200_PCM_F32 };
inline std::vector<AnalysisResult> analyze_batch_pcm(const std::vector<const void *> &frames, const std::vector<uint64_t> &n_frames,
                                                     PcmFormat format, uint32_t channels, uint32_t sample_rate = BLISS_B200_SAMPLE_RATE,
                                                     const AnalysisOptions &o = {}) {
    detail::ensure_init();
    const size_t dim = feature_count(o.features_version);
    std::vector<float> out(dim * frames.size());
    std::vector<int32_t> status(frames.size());
    detail::check_call(bliss_b200_analyze_batch_pcm(frames.data(), n_frames.data(), static_cast<uint32_t>(frames.size()),
                                                    static_cast<int>(format), channels, sample_rate,
                                                    static_cast<uint16_t>(o.features_version), out.data(), status.data()));
    std::vector<AnalysisResult> res;
    res.reserve(frames.size());
    for (size_t i = 0; i < frames.size(); i++) {
        if (status[i] == BLISS_B200_SONG_OK)
            res.emplace_back(Analysis(std::vector<float>(out.begin() + i * dim, out.begin() + (i + 1) * dim), o.features_version));
        else
            res.emplace_back(detail::status_error(status[i]));
    }
    return res;
}
// the conversion alone: PreAnalyzedSong::sample_array of such a source
inline std::vector<float> pcm_to_mono(const void *frames, uint64_t n_frames, PcmFormat format, uint32_t channels) {
    detail::ensure_init();
    std::vector<float> out(n_frames);
    detail::check_call(bliss_b200_pcm_to_mono(frames, n_frames, static_cast<int>(format), channels, out.data()));
    return out;
}

// mono f32 at sample_rate -> 22 050 Hz on the device (bliss_b200_resample: parity unpinned against the reference's
// swresample / rubato, see include/bliss_b200.h)
inline std::vector<float> resample(const std::vector<float> &mono, uint32_t sample_rate) {
    detail::ensure_init();
    std::vector<float> out(bliss_b200_resampled_len(mono.size(), sample_rate));
    uint64_t n = 0;
    detail::check_call(bliss_b200_resample(mono.data(), mono.size(), sample_rate, out.data(), out.size(), &n));
    out.resize(n);
    return out;
}

// src/song/decoder.rs:34-67
struct PreAnalyzedSong {
    std::string path;
    std::optional<std::string> artist, album_artist, title, album, genre;
    std::optional<int> track_number, disc_number;
    double duration_s = 0.;
    std::vector<float> sample_array;  // mono f32le 22 050 Hz
    // Not in the reference.  A decoder may leave the codec's packed interleaved frames here (at pcm_rate Hz) instead
    // of filling sample_array (pcm_channels > 0 says so): sample-format conversion, down-mix and the conversion to
    // 22 050 Hz then run on the device behind the copy (analyze_batch_pcm), and e.g. 16-bit mono material sends half
    // the bytes over PCIe.
    std::vector<unsigned char> pcm_frames;
    PcmFormat pcm_format = PcmFormat::F32;
    uint32_t pcm_channels = 0;
    uint32_t pcm_rate = SAMPLE_RATE;
    uint64_t n_frames() const { return pcm_channels ? pcm_frames.size() / ((pcm_format == PcmFormat::S16 ? 2 : 4) * pcm_channels) : sample_array.size(); }
    std::vector<float> mono() const {  // sample_array as this backend's decode steps fill it
        if (!pcm_channels) return sample_array;
        std::vector<float> m = pcm_to_mono(pcm_frames.data(), n_frames(), pcm_format, pcm_channels);
        return pcm_rate == SAMPLE_RATE ? m : resample(m, pcm_rate);
    }
};

// One batch of decoded songs -> one entry per song.  Songs that carry sample_array go through analyze_batch; songs
// that carry packed frames are grouped by (sample format, channel count, sample rate) -- one
// bliss_b200_analyze_batch_pcm call takes one of each.
inline std::vector<AnalysisResult> analyze_decoded(const std::vector<PreAnalyzedSong> &songs, const AnalysisOptions &o = {}) {
    std::vector<AnalysisResult> out(songs.size(), AnalysisResult(BlissError(BlissError::AnalysisError, "not analysed")));
    std::vector<char> done(songs.size(), 0);
    for (size_t first = 0; first < songs.size(); first++) {
        if (done[first]) continue;
        const PreAnalyzedSong &lead = songs[first];
        std::vector<size_t> idx;
        for (size_t i = first; i < songs.size(); i++)
            if (!done[i] && songs[i].pcm_channels == lead.pcm_channels && (lead.pcm_channels == 0 || (songs[i].pcm_format == lead.pcm_format && songs[i].pcm_rate == lead.pcm_rate))) {
                idx.push_back(i);
                done[i] = 1;
            }
        std::vector<uint64_t> lens;
        for (size_t i : idx) lens.push_back(songs[i].n_frames());
        std::vector<AnalysisResult> res;
        if (lead.pcm_channels == 0) {
            std::vector<const float *> ptrs;
            for (size_t i : idx) ptrs.push_back(songs[i].sample_array.data());
            res = analyze_batch(ptrs, lens, o);
        } else {
            std::vector<const void *> ptrs;
            for (size_t i : idx) ptrs.push_back(songs[i].pcm_frames.data());
            res = analyze_batch_pcm(ptrs, lens, lead.pcm_format, lead.pcm_channels, lead.pcm_rate, o);
        }
        for (size_t k = 0; k < idx.size(); k++) out[idx[k]] = std::move(res[k]);
    }
    return out;
}

// trait Decoder, src/song/decoder.rs:115-333: implement decode(); the provided functions keep the
// reference's meaning, with analysis batched onto the GPU.
class Decoder {
  public:
    virtual ~Decoder() = default;
    virtual PreAnalyzedSong decode(const std::string &path) = 0;  // throws BlissError(DecodingError)

    Song song_from_path(const std::string &path, const AnalysisOptions &o = {}) {
        PreAnalyzedSong p = decode(path);
        if (p.pcm_channels == 0) return to_song(p, Song::analyze_with_options(p.sample_array.data(), p.sample_array.size(), o), o);
        std::vector<PreAnalyzedSong> one;
        one.push_back(std::move(p));
        AnalysisResult r = std::move(analyze_decoded(one, o)[0]);
        if (auto *e = std::get_if<BlissError>(&r)) throw *e;
        return to_song(one[0], std::get<Analysis>(std::move(r)), o);
    }

    using PathResult = std::pair<std::string, std::variant<Song, BlissError>>;
    // Decoder::analyze_paths_with_options, src/song/decoder.rs:278-332, with the thread layout of the reference up to
    // the analysis: min(available cores, number_cores) threads, each on a contiguous chunk of the paths (:283-304),
    // call decode() -- which therefore has to be safe to call concurrently, like the reference's associated
    // function -- and decoding errors are items (:319-325).  Instead of analysing its own song a worker hands the
    // decoded buffer over (bounded: 2 x batch_songs songs, decoders stall rather than pile PCM up); the calling
    // thread is the batcher: <= batch_songs buffers per bliss_b200_analyze_batch call while the workers keep
    // decoding.  The order of the results is the order of arrival (unspecified in the reference as well).
    // INTEGRATION.md section 3 is the same body in Rust.
    std::vector<PathResult> analyze_paths(const std::vector<std::string> &paths, const AnalysisOptions &o = {},
                                          size_t batch_songs = 64) {
        std::vector<PathResult> out;
        if (paths.empty()) return out;
        batch_songs = std::max<size_t>(batch_songs, 1);
        const size_t cores = std::max<size_t>(1, std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), o.number_cores));
        const size_t chunk_length = std::max<size_t>(paths.size() / cores, 1);
        std::mutex mu;  // guards everything below and `out`
        std::condition_variable not_empty, not_full;
        std::deque<PreAnalyzedSong> decoded;
        size_t workers_left = (paths.size() + chunk_length - 1) / chunk_length;
        std::exception_ptr failure;
        auto worker = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) {
                try {
                    const std::string &pth = paths[i];
                    if (pth.size() >= 4 && cue_extension(pth.substr(pth.size() - 4))) {  // :305-318: every track of the sheet is an item
                        auto items = songs_from_cue(pth, o);
                        std::lock_guard<std::mutex> lk(mu);
                        for (auto &it : items) out.emplace_back(pth, std::move(it));
                        continue;
                    }
                    PreAnalyzedSong p = decode(paths[i]);
                    std::unique_lock<std::mutex> lk(mu);
                    not_full.wait(lk, [&] { return decoded.size() < 2 * batch_songs || failure; });
                    if (failure) break;
                    decoded.push_back(std::move(p));
                    not_empty.notify_one();
                } catch (const BlissError &e) {
                    std::lock_guard<std::mutex> lk(mu);
                    out.emplace_back(paths[i], e);
                } catch (...) {  // a bug in decode(): rethrown on the calling thread
                    std::lock_guard<std::mutex> lk(mu);
                    if (!failure) failure = std::current_exception();
                    break;
                }
            }
            std::lock_guard<std::mutex> lk(mu);
            workers_left--;
            not_empty.notify_one();
        };
        std::vector<std::thread> threads;
        for (size_t lo = 0; lo < paths.size(); lo += chunk_length)
            threads.emplace_back(worker, lo, std::min(paths.size(), lo + chunk_length));
        std::vector<PreAnalyzedSong> batch;
        try {
            for (;;) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    not_empty.wait(lk, [&] { return !decoded.empty() || workers_left == 0 || failure; });
                    if (failure) break;
                    while (!decoded.empty() && batch.size() < batch_songs) {
                        batch.push_back(std::move(decoded.front()));
                        decoded.pop_front();
                    }
                    not_full.notify_all();
                    if (batch.size() < batch_songs && (workers_left > 0 || !decoded.empty())) continue;  // keep filling
                    if (batch.empty()) break;  // every worker is done and everything decoded has been analysed
                }
                auto res = analyze_decoded(batch, o);  // outside the lock: the workers keep decoding
                std::lock_guard<std::mutex> lk(mu);
                for (size_t i = 0; i < batch.size(); i++) {
                    if (auto *a = std::get_if<Analysis>(&res[i])) out.emplace_back(batch[i].path, to_song(batch[i], *a, o));
                    else out.emplace_back(batch[i].path, std::get<BlissError>(res[i]));
                }
                batch.clear();
            }
        } catch (...) {  // e.g. a CUDA failure of the whole call: stop the workers before unwinding past their captures
            std::lock_guard<std::mutex> lk(mu);
            if (!failure) failure = std::current_exception();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            not_full.notify_all();
        }
        for (auto &t : threads) t.join();
        if (failure) std::rethrow_exception(failure);
        return out;
    }

  private:
    static bool cue_extension(std::string e) {
        for (char &c : e) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
        return e == ".cue";
    }
    inline std::vector<std::variant<Song, BlissError>> songs_from_cue(const std::string &path, const AnalysisOptions &o);  // BlissCue, below
    static Song to_song(const PreAnalyzedSong &p, Analysis a, const AnalysisOptions &o) {  // decoder.rs:85-100
        Song s;
        s.path = p.path; s.artist = p.artist; s.album_artist = p.album_artist; s.title = p.title; s.album = p.album;
        s.genre = p.genre; s.track_number = p.track_number; s.disc_number = p.disc_number; s.duration_s = p.duration_s;
        s.analysis = std::move(a);
        s.features_version = o.features_version;
        return s;
    }
};

// A concrete Decoder for the one family of sources this backend takes without the crate's decoders: RIFF/WAVE PCM
// files.  The reference's decoders unpack such a file, convert the sample format to f32, down-mix and resample to
// 22 050 Hz (src/lib.rs:143).  decode() does the first on the host and leaves the packed frames in
// PreAnalyzedSong::pcm_frames (their rate in pcm_rate), the rest runs on the device (the resampler's parity against
// swresample / rubato is unpinned, include/bliss_b200.h).  Widths as ffmpeg's pcm decoders deliver them:
// u8 -> (x - 128) 2^-7 (carried as s16), s16 -> x 2^-15, s24 -> x 2^-23 (carried as s32: x << 8), s32 -> x 2^-31,
// IEEE f32 as it is.  Any other encoding throws DecodingError.
class WavDecoder : public Decoder {
  public:
    PreAnalyzedSong decode(const std::string &path) override {
        auto fail = [&](const std::string &why) { return BlissError(BlissError::DecodingError, "while opening format for file '" + path + "': " + why + "."); };
        std::FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw fail("cannot open");
        std::vector<unsigned char> raw;
        unsigned char buf[1 << 16];
        for (size_t got; (got = std::fread(buf, 1, sizeof buf, f)) > 0;) raw.insert(raw.end(), buf, buf + got);
        std::fclose(f);
        auto u16 = [&](size_t o) { return static_cast<uint32_t>(raw[o] | (raw[o + 1] << 8)); };
        auto u32 = [&](size_t o) { return static_cast<uint32_t>(raw[o]) | (static_cast<uint32_t>(raw[o + 1]) << 8) | (static_cast<uint32_t>(raw[o + 2]) << 16) | (static_cast<uint32_t>(raw[o + 3]) << 24); };
        if (raw.size() < 12 || std::memcmp(raw.data(), "RIFF", 4) != 0 || std::memcmp(raw.data() + 8, "WAVE", 4) != 0) throw fail("not a RIFF/WAVE file");
        uint32_t tag = 0, channels = 0, rate = 0, bits = 0;
        size_t data_off = 0, data_len = 0;
        for (size_t o = 12; o + 8 <= raw.size();) {
            const size_t len = u32(o + 4), body = o + 8;
            if (std::memcmp(raw.data() + o, "fmt ", 4) == 0 && len >= 16 && body + len <= raw.size()) {
                tag = u16(body); channels = u16(body + 2); rate = u32(body + 4); bits = u16(body + 14);
                if (tag == 0xFFFE && len >= 26) tag = u16(body + 24);  // WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
            } else if (std::memcmp(raw.data() + o, "data", 4) == 0) {
                data_off = body;
                data_len = std::min(len, raw.size() - body);  // a truncated file: what is there
                break;
            }
            o = body + len + (len & 1);
        }
        if (!channels || !data_off) throw fail("no fmt / data chunk");
        if (rate < BLISS_B200_MIN_SAMPLE_RATE || rate > BLISS_B200_MAX_SAMPLE_RATE) throw fail("runs at " + std::to_string(rate) + " Hz");
        if (channels > BLISS_B200_PCM_MAX_CHANNELS) throw fail(std::to_string(channels) + " channels");
        const bool pcm = tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32), flt = tag == 3 && bits == 32;
        if (!pcm && !flt) throw fail("encoding " + std::to_string(tag) + " with " + std::to_string(bits) + " bits per sample");
        const size_t width = bits / 8, n = data_len / (width * channels), count = n * channels;
        const unsigned char *d = raw.data() + data_off;
        PreAnalyzedSong p;
        p.path = path;
        p.duration_s = static_cast<double>(n) / rate;
        p.pcm_channels = channels;
        p.pcm_rate = rate;
        p.pcm_format = flt ? PcmFormat::F32 : (bits <= 16 ? PcmFormat::S16 : PcmFormat::S32);
        p.pcm_frames.resize(count * (bits <= 16 ? 2 : 4));
        if (bits == 16 || bits == 32) {
            std::memcpy(p.pcm_frames.data(), d, p.pcm_frames.size());  // little-endian hosts only, like the rest of the library
        } else if (bits == 8) {
            for (size_t i = 0; i < count; i++) {
                const int16_t v = static_cast<int16_t>((static_cast<int>(d[i]) - 128) * 256);
                std::memcpy(p.pcm_frames.data() + 2 * i, &v, 2);
            }
        } else {  // 24
            for (size_t i = 0; i < count; i++) {
                const uint32_t v = (static_cast<uint32_t>(d[3 * i]) << 8) | (static_cast<uint32_t>(d[3 * i + 1]) << 16) | (static_cast<uint32_t>(d[3 * i + 2]) << 24);
                std::memcpy(p.pcm_frames.data() + 4 * i, &v, 4);
            }
        }
        return p;
    }
};

// ---- src/cue.rs -----------------------------------------------------------------------------------
// BlissCue<D>: the songs of a CUE sheet, cut out of ONE decoded buffer per FILE of the sheet (:208-243); here the
// slices of all tracks of all files go to the device in one batched call.  Sheet parsing is the rcue crate's job in
// the reference (0.1.3, not vendored): restated for the commands src/cue.rs reads -- REM comments, PERFORMER, TITLE,
// FILE, TRACK, INDEX mm:ss:ff at 75 frames per second -- in non-strict mode.  Track boundaries as the reference
// computes them: (index.as_secs_f32() * SAMPLE_RATE as f32) as usize, in f32 (:212-213, :231).
namespace cue {
struct Track {
    std::string no;
    std::optional<std::string> title, performer;
    std::vector<std::pair<std::string, std::pair<uint64_t, uint32_t>>> indices;  // (index number, (seconds, nanoseconds))
};
struct CueFile { std::string file; std::vector<Track> tracks; };
struct Cue {
    std::optional<std::string> performer, title;
    std::vector<std::pair<std::string, std::string>> comments;
    std::vector<CueFile> files;
};
namespace detail {
inline std::string trim(const std::string &x) {
    const size_t a = x.find_first_not_of(" \t\r\n"), b = x.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : x.substr(a, b - a + 1);
}
inline std::string unquote(const std::string &x) {
    const std::string t = trim(x);
    return t.size() >= 2 && t.front() == '"' && t.back() == '"' ? t.substr(1, t.size() - 2) : t;
}
inline std::string upper(std::string x) {
    for (char &c : x) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    return x;
}
}  // namespace detail
inline Cue parse(const std::string &text) {
    Cue cue;
    bool in_file = false, in_track = false;
    size_t pos = 0;
    while (pos <= text.size()) {
        const size_t nl = text.find('\n', pos);
        std::string line = detail::trim(text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos));
        pos = nl == std::string::npos ? text.size() + 1 : nl + 1;
        if (line.compare(0, 3, "\xEF\xBB\xBF") == 0) line = line.substr(3);
        if (line.empty()) continue;
        const size_t sp = line.find(' ');
        const std::string cmd = detail::upper(line.substr(0, sp)), rest = sp == std::string::npos ? std::string() : detail::trim(line.substr(sp + 1));
        if (cmd == "REM") {
            const size_t k = rest.find(' ');
            cue.comments.emplace_back(rest.substr(0, k), k == std::string::npos ? std::string() : detail::unquote(rest.substr(k + 1)));
        } else if (cmd == "PERFORMER" || cmd == "TITLE") {
            std::optional<std::string> *dst = nullptr;
            if (in_track) dst = cmd == "TITLE" ? &cue.files.back().tracks.back().title : &cue.files.back().tracks.back().performer;
            else if (!in_file) dst = cmd == "TITLE" ? &cue.title : &cue.performer;
            if (dst) *dst = detail::unquote(rest);
        } else if (cmd == "FILE" && !rest.empty()) {
            std::string name;
            if (rest[0] == '"') {
                const size_t q = rest.find('"', 1);
                if (q == std::string::npos) continue;
                name = rest.substr(1, q - 1);
            } else {
                name = rest.substr(0, rest.find(' '));
            }
            cue.files.push_back(CueFile{name, {}});
            in_file = true;
            in_track = false;
        } else if (cmd == "TRACK" && in_file && !rest.empty()) {
            Track t;
            t.no = rest.substr(0, rest.find(' '));
            cue.files.back().tracks.push_back(t);
            in_track = true;
        } else if (cmd == "INDEX" && in_track) {
            unsigned long long mm = 0, ss = 0, ff = 0;
            char no[16] = {0};
            if (std::sscanf(rest.c_str(), "%15s %llu:%llu:%llu", no, &mm, &ss, &ff) == 4)
                cue.files.back().tracks.back().indices.emplace_back(no, std::make_pair(static_cast<uint64_t>(mm * 60 + ss), static_cast<uint32_t>(ff * 1000000000ull / 75)));
        }
    }
    return cue;
}
inline size_t sample_index(const std::pair<uint64_t, uint32_t> &ts) {  // Duration::as_secs_f32() * SAMPLE_RATE as f32, as usize
    volatile float secs = static_cast<float>(ts.first) + static_cast<float>(ts.second) / 1000000000.0f;
    volatile float idx = secs * static_cast<float>(SAMPLE_RATE);
    return static_cast<size_t>(idx);
}
}  // namespace cue

class BlissCue {
  public:
    explicit BlissCue(Decoder &d) : decoder_(d) {}
    using Item = std::variant<Song, BlissError>;
    // songs_from_path_with_options, :85-106: one entry per track -- a Song with cue_info, or the BlissError the
    // reference would have pushed; a sheet that cannot be read throws DecodingError (:110-116).
    std::vector<Item> songs_from_path(const std::string &path, const AnalysisOptions &o = {}) {
        std::FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw BlissError(BlissError::DecodingError, "when opening CUE file '" + path + "'");
        std::string text;
        char buf[4096];
        for (size_t got; (got = std::fread(buf, 1, sizeof buf, f)) > 0;) text.append(buf, got);
        std::fclose(f);
        const cue::Cue sheet = cue::parse(text);
        std::optional<std::string> genre;
        std::optional<int> disc_number;
        for (const auto &c : sheet.comments) {
            const std::string k = cue::detail::upper(c.first);
            if (k == "GENRE" && !genre) genre = c.second;
            if ((k == "DISCNUMBER" || k == "DISC") && !disc_number) {
                char *end = nullptr;
                const long v = std::strtol(c.second.c_str(), &end, 10);
                if (end != c.second.c_str() && *end == 0) disc_number = static_cast<int>(v);
                else break;
            }
        }
        const size_t slash = path.find_last_of('/');
        const std::string parent = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
        std::vector<Item> out;
        std::vector<PreAnalyzedSong> pieces;  // what goes to the device, in the order of `slots`
        std::vector<size_t> slots;
        for (const cue::CueFile &cf : sheet.files) {
            const std::string audio = parent + cf.file;
            PreAnalyzedSong decoded;
            try {
                decoded = decoder_.decode(audio);
            } catch (const BlissError &e) {
                out.emplace_back(e);
                continue;
            }
            if (decoded.pcm_channels && decoded.pcm_rate != SAMPLE_RATE) {
                // the sheet's indices count 22 050 Hz samples (:212-213): such a file is cut after its conversion
                decoded.sample_array = decoded.mono();
                decoded.pcm_frames.clear();
                decoded.pcm_channels = 0;
            }
            const size_t total = decoded.n_frames();
            if (total == 0) {
                out.emplace_back(BlissError(BlissError::DecodingError, "empty audio file associated to CUE sheet"));
                continue;
            }
            const size_t frame_bytes = decoded.pcm_channels ? (decoded.pcm_format == PcmFormat::S16 ? 2 : 4) * decoded.pcm_channels : 0;
            const size_t n_tracks = cf.tracks.size();
            for (size_t i = 0; i < n_tracks; i++) {
                const cue::Track &t = cf.tracks[i];
                if (t.indices.empty()) continue;
                size_t end = total;  // the last track runs to the end of the file (:229-241)
                if (i + 1 < n_tracks) {
                    if (cf.tracks[i + 1].indices.empty()) continue;
                    end = cue::sample_index(cf.tracks[i + 1].indices.front().second);
                }
                const size_t start = cue::sample_index(t.indices.front().second);
                if (!(start <= end && end <= total)) {  // the reference's slice would panic here
                    out.emplace_back(BlissError(BlissError::DecodingError, "CUE track " + t.no + " of '" + path + "' lies outside its audio file"));
                    continue;
                }
                Song sg;
                char name[32];
                std::snprintf(name, sizeof name, "/CUE_TRACK%03zu", i + 1);
                sg.path = path + name;
                sg.album = sheet.title; sg.artist = t.performer; sg.album_artist = sheet.performer; sg.title = t.title;
                sg.genre = genre; sg.disc_number = disc_number; sg.features_version = o.features_version;
                char *endp = nullptr;
                const long no = std::strtol(t.no.c_str(), &endp, 10);
                if (endp != t.no.c_str() && *endp == 0) sg.track_number = static_cast<int>(no);
                sg.duration_s = static_cast<double>(static_cast<float>(end - start) / static_cast<float>(SAMPLE_RATE));
                sg.cue_info = Song::CueInfo{path, audio};
                PreAnalyzedSong piece;
                if (decoded.pcm_channels) {
                    piece.pcm_channels = decoded.pcm_channels;
                    piece.pcm_format = decoded.pcm_format;
                    piece.pcm_frames.assign(decoded.pcm_frames.begin() + start * frame_bytes, decoded.pcm_frames.begin() + end * frame_bytes);
                } else {
                    piece.sample_array.assign(decoded.sample_array.begin() + start, decoded.sample_array.begin() + end);
                }
                pieces.push_back(std::move(piece));
                slots.push_back(out.size());
                out.emplace_back(std::move(sg));
            }
        }
        if (!pieces.empty()) {
            std::vector<AnalysisResult> res = analyze_decoded(pieces, o);
            for (size_t k = 0; k < res.size(); k++) {
                if (auto *a = std::get_if<Analysis>(&res[k])) std::get<Song>(out[slots[k]]).analysis = std::move(*a);
                else out[slots[k]] = std::get<BlissError>(res[k]);
            }
        }
        return out;
    }

  private:
    Decoder &decoder_;
};

inline std::vector<std::variant<Song, BlissError>> Decoder::songs_from_cue(const std::string &path, const AnalysisOptions &o) {
    return BlissCue(*this).songs_from_path(path, o);
}

// ---- src/playlist.rs ----------------------------------------------------------------------------
namespace playlist {
struct Metric {  // what a DistanceMetricBuilder boils down to on the device
    int metric = BLISS_B200_METRIC_MAHALANOBIS;
    std::vector<float> m;  // dim*dim, empty = identity (euclidean)
    const float *mp() const { return m.empty() ? nullptr : m.data(); }
};
inline Metric euclidean_distance() { return {}; }                                         // :65-71
inline Metric cosine_distance() { return {BLISS_B200_METRIC_COSINE, {}}; }                // :76-79
inline Metric mahalanobis_distance_builder(std::vector<float> m) { return {BLISS_B200_METRIC_MAHALANOBIS, std::move(m)}; }  // :129-131

inline float distance(const std::vector<float> &a, const std::vector<float> &b, const Metric &mt = {}) {
    detail::ensure_init();
    float d = 0.f;
    detail::check_call(bliss_b200_distance(a.data(), b.data(), static_cast<uint32_t>(a.size()), mt.metric, mt.mp(), &d));
    return d;
}
// closest_to_songs, :256-270: returns candidate indices, stable by summed distance to the seeds
inline std::vector<uint32_t> closest_to_songs(const std::vector<float> &seeds, const std::vector<float> &cands,
                                              uint32_t dim, const Metric &mt = {}) {
    detail::ensure_init();
    std::vector<uint32_t> order(cands.size() / dim);
    detail::check_call(bliss_b200_closest_to_songs(seeds.data(), static_cast<uint32_t>(seeds.size() / dim), cands.data(),
                                                   static_cast<uint32_t>(order.size()), dim, mt.metric, mt.mp(),
                                                   order.data(), nullptr));
    return order;
}
// song_to_song, :272-326
inline std::vector<uint32_t> song_to_song(const std::vector<float> &seeds, const std::vector<float> &cands, uint32_t dim,
                                          const Metric &mt = {}) {
    detail::ensure_init();
    std::vector<uint32_t> order(cands.size() / dim);
    detail::check_call(bliss_b200_song_to_song(seeds.data(), static_cast<uint32_t>(seeds.size() / dim), cands.data(),
                                               static_cast<uint32_t>(order.size()), dim, mt.metric, mt.mp(), order.data()));
    return order;
}
// all pairs on the device: out[i * n_cols + j] = d(rows[i], cols[j])
inline std::vector<float> distance_matrix(const std::vector<float> &rows, const std::vector<float> &cols, uint32_t dim,
                                          const Metric &mt = {}) {
    detail::ensure_init();
    const uint32_t nr = static_cast<uint32_t>(rows.size() / dim), nc = static_cast<uint32_t>(cols.size() / dim);
    std::vector<float> out(static_cast<size_t>(nr) * nc);
    if (!out.empty()) detail::check_call(bliss_b200_distance_matrix(rows.data(), nr, cols.data(), nc, dim, mt.metric, mt.mp(), out.data()));
    return out;
}
// variance_based_weight_matrix, :173-221: the diagonal Mahalanobis matrix (dim x dim, row-major) that weights the
// dimensions the seeds agree on -- inverse variance, normalised so that the weights sum to dim.  A few f32 operations
// on n_seeds x dim values, host side as in the reference; feed the result to mahalanobis_distance_builder.
inline std::vector<float> variance_based_weight_matrix(const std::vector<std::vector<float>> &seeds) {
    if (seeds.size() < 2) throw BlissError(BlissError::ProviderError, "seeds must contain more than one element");
    const size_t n = seeds[0].size();
    if (n == 0) throw BlissError(BlissError::ProviderError, "seed feature vectors must not be empty");
    for (const auto &sd : seeds)
        if (sd.size() != n) throw BlissError(BlissError::ProviderError, "all seed feature vectors must have the same length");
    const float n_seeds = static_cast<float>(seeds.size());
    std::vector<float> mean(n, 0.f), variance(n, 0.f), m(n * n, 0.f);
    for (const auto &sd : seeds)
        for (size_t i = 0; i < n; i++) mean[i] += sd[i];
    for (size_t i = 0; i < n; i++) mean[i] /= n_seeds;
    for (const auto &sd : seeds)
        for (size_t i = 0; i < n; i++) {
            const float diff = sd[i] - mean[i];
            variance[i] = variance[i] + diff * diff;
        }
    float sum = 0.f;
    for (size_t i = 0; i < n; i++) {
        variance[i] = 1.0f / (variance[i] / n_seeds + 1e-6f);
        sum += variance[i];
    }
    for (size_t i = 0; i < n; i++) m[i * n + i] = variance[i] * (static_cast<float>(n) / sum);
    return m;
}
// dedup_playlist_custom_distance, :367-402 (dedup_playlist, :343-349, is the euclidean case): a song goes when it is
// closer than the threshold (default 0.05) to the last song kept, or carries the same title and artist tags as it.
// The distances come from one device call.  Returns the indices of the songs that stay.
inline std::vector<size_t> dedup_playlist_custom_distance(const std::vector<Song> &songs, std::optional<float> distance_threshold = {},
                                                          const Metric &mt = {}) {
    std::vector<size_t> keep;
    if (songs.empty()) return keep;
    const float thr = distance_threshold.value_or(0.05f);
    const uint32_t dim = static_cast<uint32_t>(songs[0].analysis->as_vec().size());
    std::vector<float> rows;
    for (const Song &sg : songs) rows.insert(rows.end(), sg.analysis->as_vec().begin(), sg.analysis->as_vec().end());
    const std::vector<float> d = songs.size() > 1 ? distance_matrix(rows, rows, dim, mt) : std::vector<float>();
    const size_t n = songs.size();
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        for (; j < n; j++) {
            const Song &a = songs[i], &b = songs[j];
            const bool same_tags = a.title && b.title && a.artist && b.artist && *a.title == *b.title && *a.artist == *b.artist;
            if (!(d[i * n + j] < thr || same_tags)) break;
        }
        keep.push_back(i);
        i = j;
    }
    return keep;
}
inline std::vector<size_t> dedup_playlist(const std::vector<Song> &songs, std::optional<float> distance_threshold = {}) {
    return dedup_playlist_custom_distance(songs, distance_threshold, euclidean_distance());
}
// closest_album_to_group, :424-485: an "album playlist" -- `group` first, then the albums of `pool` ordered by the
// euclidean distance of their mean analysis to the group's mean analysis (one device call), each album by
// (disc, track) with None < Some(_); songs of the group and songs without an album tag leave the pool.
inline std::vector<Song> closest_album_to_group(const std::vector<Song> &group, const std::vector<Song> &pool_in) {
    if (group.empty()) throw BlissError(BlissError::ProviderError, "Mean of empty slice");
    auto same = [](const Song &a, const Song &b) {  // #[derive(PartialEq)] on Song
        return a.path == b.path && a.artist == b.artist && a.album_artist == b.album_artist && a.title == b.title && a.album == b.album &&
               a.genre == b.genre && a.track_number == b.track_number && a.disc_number == b.disc_number && a.duration_s == b.duration_s &&
               a.features_version == b.features_version && a.analysis.has_value() == b.analysis.has_value() &&
               (!a.analysis || a.analysis->as_vec() == b.analysis->as_vec());
    };
    std::vector<const Song *> pool;
    for (const Song &sg : pool_in) {
        bool in_group = false;
        for (const Song &g : group) in_group = in_group || same(g, sg);
        if (!in_group) pool.push_back(&sg);
    }
    const size_t dim = group[0].analysis->as_vec().size();
    auto mean_of = [&](const std::vector<const Song *> &v) {  // mean_axis(Axis(0)): rows summed in order, then / n, in f32
        std::vector<float> m(dim, 0.f);
        for (const Song *sg : v)
            for (size_t i = 0; i < dim; i++) m[i] += sg->analysis->as_vec()[i];
        for (float &x : m) x /= static_cast<float>(v.size());
        return m;
    };
    std::vector<std::string> names;
    std::vector<std::vector<const Song *>> members;
    for (const Song *sg : pool) {
        if (!sg->album) continue;
        size_t k = 0;
        while (k < names.size() && names[k] != *sg->album) k++;
        if (k == names.size()) { names.push_back(*sg->album); members.emplace_back(); }
        members[k].push_back(sg);
    }
    std::vector<const Song *> group_ptrs;
    for (const Song &g : group) group_ptrs.push_back(&g);
    std::vector<Song> out(group);
    if (names.empty()) return out;
    std::vector<float> means;
    for (const auto &m : members) { const auto v = mean_of(m); means.insert(means.end(), v.begin(), v.end()); }
    const std::vector<float> d = distance_matrix(mean_of(group_ptrs), means, static_cast<uint32_t>(dim));
    std::vector<size_t> order(names.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return d[a] < d[b]; });
    for (size_t k : order) {
        std::vector<const Song *> al = members[k];
        std::stable_sort(al.begin(), al.end(), [](const Song *a, const Song *b) {
            return std::make_pair(a->disc_number, a->track_number) < std::make_pair(b->disc_number, b->track_number);  // nullopt < any value
        });
        for (const Song *sg : al) out.push_back(*sg);
    }
    return out;
}
}  // namespace playlist
}  // namespace bliss

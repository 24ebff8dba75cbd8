// Exercises include/bliss_b200.hpp (the C++17 host mirror of the reference API) against libbliss_b200.so.
// Built and run by tests/test_host_abi.py: without a GPU it must fail loudly ("NO_DEVICE": there is no CPU
// fallback); with one it analyses a synthetic clip through Song::analyze, the batched Decoder seam and the
// 16-bit entry point and prints "OK".
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <set>
#include <unistd.h>

#include "bliss_b200.hpp"

using namespace bliss;

struct ToneDecoder : Decoder {  // a Decoder whose "files" are synthetic tones; "bad" fails to decode
    std::atomic<int> now{0}, peak{0}, calls{0};
    PreAnalyzedSong decode(const std::string &path) override {  // called from analyze_paths' worker threads
        struct Busy {
            ToneDecoder &d;
            explicit Busy(ToneDecoder &dd) : d(dd) {
                d.calls++;
                const int n = ++d.now;
                int p = d.peak.load();
                while (n > p && !d.peak.compare_exchange_weak(p, n)) {}
            }
            ~Busy() { d.now--; }
        } busy(*this);
        const bool is_short = path.rfind("short", 0) == 0;
        if (is_short) std::this_thread::sleep_for(std::chrono::milliseconds(5));
        if (path == "bad") throw BlissError(BlissError::DecodingError, "while opening format for file 'bad'");
        if (path == "bug") throw std::logic_error("decode() is broken");
        PreAnalyzedSong p;
        p.path = path;
        p.title = path;
        const size_t n = is_short ? 4000 : 22050 * 6;
        p.sample_array.resize(n);
        const double f = 220.0 * (1 + (int)path.size());
        for (size_t i = 0; i < n; i++)
            p.sample_array[i] = (float)(0.3 * std::sin(2 * M_PI * f * i / SAMPLE_RATE) + 0.2 * ((i / 5512) % 2 ? 1 : 0) * std::sin(2 * M_PI * 80.0 * i / SAMPLE_RATE));
        p.duration_s = (double)n / SAMPLE_RATE;
        return p;
    }
};

int main() {
    // things that need no device
    if (feature_count(FeaturesVersion::Version2) != 23 || feature_count(FeaturesVersion::Version1) != 20) return 2;
    const auto w = feature_weights(LATEST);
    if (w.size() != 23 * 23 || w[0] != 0.25f || w[24] != 1.0f) return 3;
    try {
        Analysis wrong(std::vector<float>(5, 0.f), LATEST);  // Analysis::new length check, src/song/mod.rs:326-339
        return 4;
    } catch (const BlissError &e) {
        if (e.kind != BlissError::ProviderError) return 5;
    }
    // variance_based_weight_matrix: the reference's own tests, src/playlist.rs:1664-1760 (host side, no device)
    {
        using playlist::variance_based_weight_matrix;
        const auto m = variance_based_weight_matrix({{1.f, 0.f, 1.f}, {1.f, 100.f, 1.f}, {1.f, 200.f, 1.f}});
        if (m.size() != 9 || !(m[0] > m[4]) || !(m[8] > m[4])) return 30;
        if (m[1] != 0.f || m[2] != 0.f || m[3] != 0.f || m[5] != 0.f || m[6] != 0.f || m[7] != 0.f) return 31;
        if (std::fabs(m[0] + m[4] + m[8] - 3.f) >= 1e-4f) return 32;
        const auto id = variance_based_weight_matrix({{1.f, 2.f, 3.f}, {1.f, 2.f, 3.f}, {1.f, 2.f, 3.f}});
        if (std::fabs(id[0] - 1.f) >= 1e-4f || std::fabs(id[4] - 1.f) >= 1e-4f || std::fabs(id[8] - 1.f) >= 1e-4f) return 33;
        const auto two = variance_based_weight_matrix({{0.f, 50.f}, {0.f, 150.f}});
        if (two.size() != 4 || !(two[0] > two[3])) return 34;
        int refused = 0;
        for (const std::vector<std::vector<float>> &bad :
             {std::vector<std::vector<float>>{{1.f, 2.f}}, std::vector<std::vector<float>>{{1.f, 2.f}, {1.f}},
              std::vector<std::vector<float>>{{}, {}}}) {
            try {
                variance_based_weight_matrix(bad);
            } catch (const BlissError &e) {
                refused += e.kind == BlissError::ProviderError;
            }
        }
        if (refused != 3) return 35;
    }
    // src/cue.rs: the reference's test sheet (data/testcue.cue), parsed; 0:11:05 -> sample 244 020 and 0:16:69 -> 373 086,
    // the boundaries behind the durations its test asserts (:311, :356, :402)
    const std::string sheet_text =
        "REM GENRE Random\nREM DATE 2022\nREM DISCNUMBER 1\nPERFORMER \"Polochon_street\"\nTITLE \"Album for CUE test\"\n"
        "FILE \"%s\" WAVE\n  TRACK 01 AUDIO\n    TITLE \"Renaissance\"\n    PERFORMER \"David TMX\"\n    INDEX 01 0:00:00\n"
        "  TRACK 02 AUDIO\n    TITLE \"Piano\"\n    PERFORMER \"Polochon_street\"\n    INDEX 01 0:11:05\n"
        "  TRACK 03 AUDIO\n    TITLE \"Tone\"\n    PERFORMER \"Polochon_street\"\n    INDEX 01 0:16:69\n\n"
        "FILE \"not-existing.wav\" WAVE\n  TRACK 01 AUDIO\n    TITLE \"Nope\"\n    PERFORMER \"Charlie\"\n    INDEX 01 0:00:00\n";
    {
        const cue::Cue c = cue::parse(sheet_text);
        if (c.files.size() != 2 || c.files[0].tracks.size() != 3 || c.files[0].file != "%s" || c.files[1].file != "not-existing.wav") return 60;
        if (*c.performer != "Polochon_street" || *c.title != "Album for CUE test" || c.comments.size() != 3 || c.comments[2].second != "1") return 61;
        const cue::Track &t2 = c.files[0].tracks[1];
        if (t2.no != "02" || *t2.title != "Piano" || *t2.performer != "Polochon_street" || t2.indices.size() != 1) return 62;
        if (cue::sample_index(t2.indices[0].second) != 244020 || cue::sample_index(c.files[0].tracks[2].indices[0].second) != 373086) return 63;
        if ((float)244020 / 22050.f != 11.066666603f || (float)(373086 - 244020) / 22050.f != 5.853333473f) return 64;
    }
    ToneDecoder dec;
    try {
        const Song s = dec.song_from_path("a");
        if (s.analysis->as_vec().size() != 23) return 6;
        const Song t = dec.song_from_path("bbb");
        const float d = s.distance(t), d0 = s.distance(s);
        if (!(d > 0.f) || d0 != 0.f) return 7;
        // the batching seam: errors are items, the batch survives them (src/song/decoder.rs:319-325)
        auto res = dec.analyze_paths({"a", "bad", "short", "bbb"}, {}, 2);
        if (res.size() != 4) return 8;
        int ok = 0, err = 0;
        for (auto &r : res) {
            if (auto *song = std::get_if<Song>(&r.second)) {
                ok++;
                const Song &ref = r.first == "a" ? s : t;
                if (std::memcmp(song->analysis->as_vec().data(), ref.analysis->as_vec().data(), 23 * sizeof(float)) != 0) return 9;
            } else {
                err++;
                const auto &e = std::get<BlissError>(r.second);
                if (r.first == "short" && std::string(e.what()).find("empty or too short song.") == std::string::npos) return 10;
                if (r.first == "bad" && e.kind != BlissError::DecodingError) return 11;
            }
        }
        if (ok != 2 || err != 2) return 12;
        // the same seam with decoding threads: 3 cores over 14 paths -> chunks of 4 -> 4 worker threads (the reference's
        // paths.chunks(len / cores), src/song/decoder.rs:300-304), every path answered exactly once, rows stay with
        // their songs, the calling thread batches <= 3 songs per GPU call while the workers decode
        {
            std::vector<std::string> paths = {"short0", "short1", "a", "short2", "short3", "bad", "short4", "short5",
                                              "short6", "bbb", "short7", "short8", "short9", "shortA"};
            AnalysisOptions o3;
            o3.number_cores = 3;
            dec.peak = 0;
            dec.calls = 0;
            auto many = dec.analyze_paths(paths, o3, 3);
            if (many.size() != paths.size() || dec.calls != (int)paths.size()) return 21;
            std::multiset<std::string> want(paths.begin(), paths.end()), got;
            for (auto &r : many) {
                got.insert(r.first);
                if (auto *song = std::get_if<Song>(&r.second)) {
                    if (r.first != "a" && r.first != "bbb") return 22;
                    const Song &ref = r.first == "a" ? s : t;
                    if (song->path != r.first || std::memcmp(song->analysis->as_vec().data(), ref.analysis->as_vec().data(), 23 * sizeof(float)) != 0) return 23;
                } else if ((r.first == "bad") != (std::get<BlissError>(r.second).kind == BlissError::DecodingError)) {
                    return 24;
                }
            }
            if (got != want) return 25;
            if (std::thread::hardware_concurrency() >= 2 && dec.peak < 2) return 26;  // decoders did run side by side
            try {  // a failure of decode() that is not a BlissError reaches the caller, after the workers have stopped
                dec.analyze_paths({"short0", "bug", "short1", "short2"}, o3, 2);
                return 27;
            } catch (const std::logic_error &) {
            }
            if (dec.now != 0) return 28;
            if (!dec.analyze_paths({}, o3).empty()) return 29;
        }
        // 16-bit entry point == f32 entry point on x / 32768
        PreAnalyzedSong p = dec.decode("a");
        std::vector<int16_t> q(p.sample_array.size());
        std::vector<float> back(q.size());
        for (size_t i = 0; i < q.size(); i++) {
            q[i] = (int16_t)std::lrint(p.sample_array[i] * 32767.0);
            back[i] = (float)q[i] / 32768.0f;
        }
        auto r16 = analyze_batch_s16({q.data()}, {q.size()});
        auto r32 = analyze_batch({back.data()}, {back.size()});
        if (std::get<Analysis>(r16[0]).as_vec() != std::get<Analysis>(r32[0]).as_vec()) return 13;
        // interleaved stereo 16-bit frames == f32 entry point on the down-mixed samples (c L + c R, c = sqrt(1/2))
        std::vector<int16_t> st(2 * q.size());
        for (size_t i = 0; i < q.size(); i++) {
            st[2 * i] = q[i];
            st[2 * i + 1] = (int16_t)(q[i] / 2);
        }
        const std::vector<float> mono = pcm_to_mono(st.data(), q.size(), PcmFormat::S16, 2);
        const float c = 0.70710678118654752440f;
        for (size_t i = 0; i < q.size(); i += 997) {
            const volatile float l = ((float)st[2 * i] / 32768.0f) * c, r = ((float)st[2 * i + 1] / 32768.0f) * c;
            if (mono[i] != l + r) return 15;
        }
        auto rst = analyze_batch_pcm({st.data()}, {q.size()}, PcmFormat::S16, 2);
        auto rmo = analyze_batch({mono.data()}, {mono.size()});
        if (std::get<Analysis>(rst[0]).as_vec() != std::get<Analysis>(rmo[0]).as_vec()) return 16;
        try {
            analyze_batch_pcm({st.data()}, {q.size()}, PcmFormat::S16, 2, 999);
            return 17;  // outside BLISS_B200_MIN_SAMPLE_RATE .. _MAX_SAMPLE_RATE: must refuse
        } catch (const BlissError &) {
        }
        // the same frames taken as 44.1 kHz material: down-mix + resampler + analysis in one call = the three steps apart
        const std::vector<float> mono_cd = resample(mono, 44100);
        if (mono_cd.size() != (mono.size() + 1) / 2 || bliss_b200_resampled_len(mono.size(), 44100) != mono_cd.size()) return 18;
        auto rcd = analyze_batch_pcm({st.data()}, {q.size()}, PcmFormat::S16, 2, 44100);
        auto rcd2 = analyze_batch({mono_cd.data()}, {mono_cd.size()});
        if (std::get<Analysis>(rcd[0]).as_vec() != std::get<Analysis>(rcd2[0]).as_vec()) return 19;
        // WavDecoder: 22 050 Hz WAV files through the decoder pipeline, the file's own frames converted on the device
        {
            const char *tmp = std::getenv("TMPDIR");
            const std::string dir = std::string(tmp ? tmp : "/tmp") + "/bliss_b200_host_mirror_" + std::to_string((long)::getpid());
            auto write_wav = [&](const std::string &name, uint32_t tag, uint32_t channels, uint32_t rate, uint32_t bits, const void *data, size_t bytes) {
                const std::string path = dir + "_" + name;
                std::FILE *f = std::fopen(path.c_str(), "wb");
                auto w32 = [&](uint32_t v) { std::fwrite(&v, 4, 1, f); };
                auto w16 = [&](uint16_t v) { std::fwrite(&v, 2, 1, f); };
                std::fwrite("RIFF", 1, 4, f); w32((uint32_t)(36 + bytes)); std::fwrite("WAVEfmt ", 1, 8, f); w32(16);
                w16((uint16_t)tag); w16((uint16_t)channels); w32(rate); w32(rate * channels * bits / 8); w16((uint16_t)(channels * bits / 8)); w16((uint16_t)bits);
                std::fwrite("data", 1, 4, f); w32((uint32_t)bytes);
                std::fwrite(data, 1, bytes, f);
                std::fclose(f);
                return path;
            };
            std::vector<unsigned char> s24(3 * q.size());
            std::vector<int32_t> s24_as_s32(q.size());
            for (size_t i = 0; i < q.size(); i++) {
                const int32_t v = (int32_t)q[i] * 256 + 5;  // a 24-bit value
                s24[3 * i] = (unsigned char)(v & 255); s24[3 * i + 1] = (unsigned char)((v >> 8) & 255); s24[3 * i + 2] = (unsigned char)((v >> 16) & 255);
                s24_as_s32[i] = (int32_t)((uint32_t)v << 8);
            }
            const std::vector<std::string> wavs = {
                write_wav("mono16.wav", 1, 1, 22050, 16, q.data(), 2 * q.size()), write_wav("stereo16.wav", 1, 2, 22050, 16, st.data(), 2 * st.size()),
                write_wav("mono24.wav", 1, 1, 22050, 24, s24.data(), s24.size()), write_wav("float.wav", 3, 1, 22050, 32, back.data(), 4 * back.size()),
                write_wav("cd.wav", 1, 2, 44100, 16, st.data(), 2 * st.size()), dir + "_missing.wav"};
            WavDecoder wd;
            AnalysisOptions o2;
            o2.number_cores = 2;
            auto got = wd.analyze_paths(wavs, o2, 3);
            auto r24 = analyze_batch_pcm({s24_as_s32.data()}, {s24_as_s32.size()}, PcmFormat::S32, 1);
            if (got.size() != wavs.size()) return 50;
            int songs_ok = 0, refused = 0;
            for (auto &r : got) {
                const std::string name = r.first.substr(dir.size() + 1);
                if (auto *song = std::get_if<Song>(&r.second)) {
                    songs_ok++;
                    const std::vector<float> &v = song->analysis->as_vec();
                    if (name == "mono16.wav" && v != std::get<Analysis>(r16[0]).as_vec()) return 51;
                    if (name == "stereo16.wav" && v != std::get<Analysis>(rst[0]).as_vec()) return 52;
                    if (name == "mono24.wav" && v != std::get<Analysis>(r24[0]).as_vec()) return 53;
                    if (name == "float.wav" && v != std::get<Analysis>(r32[0]).as_vec()) return 54;
                    if (name == "cd.wav" && v != std::get<Analysis>(rcd[0]).as_vec()) return 59;
                    if (std::fabs(song->duration_s - (name == "cd.wav" ? 3.0 : 6.0)) > 1e-9) return 55;
                } else {
                    refused += std::get<BlissError>(r.second).kind == BlissError::DecodingError && (name == "missing.wav");
                }
            }
            if (songs_ok != 5 || refused != 1) return 56;
            if (wd.song_from_path(wavs[1]).analysis->as_vec() != std::get<Analysis>(rst[0]).as_vec()) return 57;
            if (wd.decode(wavs[1]).mono() != mono) return 58;
            // a CUE sheet over the stereo file: both tracks are slices of one decoded buffer, analysed in one call, and
            // equal the analysis of the same slices of the down-mixed samples; .cue paths inside analyze_paths
            {
                std::string text = sheet_text.substr(0, sheet_text.find("  TRACK 03"));
                const std::string wav_name = wavs[1].substr(wavs[1].find_last_of('/') + 1);
                text.replace(text.find("%s"), 2, wav_name);
                text.replace(text.find("0:11:05"), 7, "0:02:37");
                const std::string cue_path = dir + "_album.cue";
                std::FILE *cf = std::fopen(cue_path.c_str(), "wb");
                std::fwrite(text.data(), 1, text.size(), cf);
                std::fclose(cf);
                const size_t cut = cue::sample_index({2, 37ull * 1000000000ull / 75});
                auto tracks = BlissCue(wd).songs_from_path(cue_path);
                auto direct = analyze_batch({mono.data(), mono.data() + cut}, {cut, mono.size() - cut});
                if (tracks.size() != 2 || cut < 54000 || cut > 55000) return 70;
                for (int k = 0; k < 2; k++) {
                    const Song *sg = std::get_if<Song>(&tracks[k]);
                    if (!sg || sg->analysis->as_vec() != std::get<Analysis>(direct[k]).as_vec()) return 71;
                    if (sg->path != cue_path + (k ? "/CUE_TRACK002" : "/CUE_TRACK001") || *sg->title != (k ? "Piano" : "Renaissance")) return 72;
                    if (*sg->album != "Album for CUE test" || *sg->album_artist != "Polochon_street" || *sg->genre != "Random" || *sg->disc_number != 1 || *sg->track_number != k + 1) return 73;
                    if (sg->cue_info->cue_path != cue_path || sg->cue_info->audio_file_path != wavs[1]) return 74;
                }
                auto items = wd.analyze_paths({wavs[0], cue_path}, o2, 3);
                int under_sheet = 0;
                for (auto &r : items) under_sheet += r.first == cue_path && std::holds_alternative<Song>(r.second);
                if (items.size() != 3 || under_sheet != 2) return 75;
                std::remove(cue_path.c_str());
            }
            for (const auto &wav : wavs) std::remove(wav.c_str());
        }
        // playlist: closest_to_songs keeps the seed first
        std::vector<float> cands;
        for (const Song *x : {&t, &s}) cands.insert(cands.end(), x->analysis->as_vec().begin(), x->analysis->as_vec().end());
        const auto order = playlist::closest_to_songs(s.analysis->as_vec(), cands, 23, playlist::mahalanobis_distance_builder(w));
        if (order.size() != 2 || order[0] != 1) return 14;
        // dedup_playlist[_custom_distance]: the reference's own test, src/playlist.rs:507-640
        {
            auto mk = [](const char *path, float filler, float v16, const char *title, const char *artist) {
                Song x;
                x.path = path;
                std::vector<float> a(23, 1.f);
                for (int i = 0; i < 16; i++) a[i] = filler;
                if (filler != 1.f) a[16] = v16;
                x.analysis = Analysis(a, LATEST);
                if (title) x.title = title;
                if (artist) x.artist = artist;
                return x;
            };
            const std::vector<Song> pl = {mk("path-to-first", 1.f, 1.f, nullptr, nullptr), mk("path-to-dupe", 1.f, 1.f, nullptr, nullptr),
                                          mk("path-to-second", 2.f, 1.9f, "dupe-title", "dupe-artist"),
                                          mk("path-to-third", 2.f, 2.5f, "dupe-title", "dupe-artist"),
                                          mk("path-to-fourth", 2.f, 0.f, "dupe-title", "no-dupe-artist"),
                                          mk("path-to-fourth", 2.f, 0.001f, nullptr, nullptr)};
            const std::vector<size_t> kept = {0, 2, 4}, first_only = {0};
            if (playlist::dedup_playlist_custom_distance(pl, {}, playlist::euclidean_distance()) != kept) return 40;
            if (playlist::dedup_playlist_custom_distance(pl, 20.f, playlist::euclidean_distance()) != first_only) return 41;
            if (playlist::dedup_playlist(pl, 20.f) != first_only || playlist::dedup_playlist(pl) != kept) return 42;
            if (!playlist::dedup_playlist({}).empty()) return 43;
            const auto dm = playlist::distance_matrix(pl[0].analysis->as_vec(), pl[2].analysis->as_vec(), 23);
            if (dm.size() != 1 || dm[0] != playlist::distance(pl[0].analysis->as_vec(), pl[2].analysis->as_vec())) return 44;
        }
        // closest_album_to_group: the reference's own test, src/playlist.rs:1113-1262
        {
            auto mk = [](const char *path, float val, const char *album, const char *artist, int track, int disc) {
                Song x;
                x.path = path;
                x.analysis = Analysis(std::vector<float>(23, val), LATEST);
                if (album) x.album = album;
                if (artist) x.artist = artist;
                if (track) x.track_number = track;
                if (disc) x.disc_number = disc;
                return x;
            };
            const Song first = mk("path-to-first", 0.f, "Album", "Artist", 1, 1), second = mk("path-to-third", 10.f, "Album", "Another Artist", 2, 1);
            const Song o11 = mk("path-to-second-2", 0.15f, "Another Album", "Artist", 1, 1), o12 = mk("path-to-second", 0.1f, "Another Album", "Artist", 2, 1);
            const Song o21 = mk("path-to-fourth", 20.f, "Another Album", "Another Artist", 1, 2), o24 = mk("path-to-fourth", 20.f, "Another Album", "Another Artist", 4, 2);
            const Song none = mk("path-to-fifth", 40.f, nullptr, "Third Artist", 0, 0);
            const auto got = playlist::closest_album_to_group({first, second}, {first, o12, o24, second, o21, o11, none});
            const std::vector<std::pair<std::string, int>> want = {{"path-to-first", 1}, {"path-to-third", 2}, {"path-to-second-2", 1},
                                                                   {"path-to-second", 2}, {"path-to-fourth", 1}, {"path-to-fourth", 4}};
            if (got.size() != want.size()) return 80;
            for (size_t i = 0; i < want.size(); i++)
                if (got[i].path != want[i].first || *got[i].track_number != want[i].second) return 81;
            std::vector<Song> pool2;
            for (int i : {2, 1}) pool2.push_back(mk(("far-" + std::to_string(i)).c_str(), 30.f + i, "Far", nullptr, i, 0));
            for (int i : {2, 1}) pool2.push_back(mk(("near-" + std::to_string(i)).c_str(), 6.f + i, "Near", nullptr, i, 0));
            pool2.push_back(first);
            const auto two = playlist::closest_album_to_group({first, second}, pool2);
            const std::vector<std::string> names = {"path-to-first", "path-to-third", "near-1", "near-2", "far-1", "far-2"};
            if (two.size() != names.size()) return 82;
            for (size_t i = 0; i < names.size(); i++)
                if (two[i].path != names[i]) return 83;
            try {
                playlist::closest_album_to_group({}, pool2);
                return 84;
            } catch (const BlissError &e) {
                if (e.kind != BlissError::ProviderError) return 85;
            }
        }
        std::puts("OK");
        return 0;
    } catch (const BlissError &e) {
        if (std::string(e.what()).find("no CPU fallback") != std::string::npos) {
            std::puts("NO_DEVICE");
            return 0;
        }
        std::fprintf(stderr, "unexpected BlissError: %s\n", e.what());
        return 20;
    }
}

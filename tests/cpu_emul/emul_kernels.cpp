// emul_kernels.cpp -- runs the SOURCE of this repository's CUDA kernels on the host (cuda_on_cpu/cuda_runtime.h:
// every CUDA thread a fiber, warp collectives and __syncthreads() real rendezvous points) on one short song and
// writes what they produce as raw little-endian arrays; tests/test_host_abi.py compares those with the oracle and
// the experimental kernel cuts (BLISS_B200_VARIANT bits) with the measured kernels.
//
//   g++ -std=c++17 -O1 -ffp-contract=off -DBLISS_HOST_EMUL -I tests/cpu_emul/cuda_on_cpu emul_kernels.cpp
//   ./emul_kernels song.f32 out_dir [full | batch n1 n2 ...]
//
// TEST INFRASTRUCTURE.  The kernels are compiled unmodified: the .cu files are #included, their launchers
// (<<< >>>) and the kernels outside this emulation are compiled out by BLISS_HOST_EMUL, inline PTX has host
// fall-backs behind #ifdef __CUDA_ARCH__.
#include <cuda_runtime.h>

#include <string>

#include "../../bliss-rs_b200/csrc/spectral.cu"
#include "../../bliss-rs_b200/csrc/wave_setup.cu"
#include "../../bliss-rs_b200/csrc/chroma.cu"
#include "../../bliss-rs_b200/csrc/tempo.cu"
#include "../../bliss-rs_b200/csrc/finalize.cu"

using namespace bliss;

static std::string g_out;

template <class T>
static void dump(const char *name, const std::vector<T> &v) {
    const std::string p = g_out + "/" + name;
    FILE *f = fopen(p.c_str(), "wb");
    if (!f) { perror(p.c_str()); exit(2); }
    fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s song.f32 out_dir [full]\n", argv[0]); return 2; }
    g_out = argv[2];
    const bool batch_mode = argc > 3 && std::string(argv[3]) == "batch";
    const bool stages = !(argc > 3 && (std::string(argv[3]) == "full" || batch_mode));  // "full": only the whole-path runs
    std::vector<float> x;
    {
        FILE *f = fopen(argv[1], "rb");
        if (!f) { perror(argv[1]); return 2; }
        fseek(f, 0, SEEK_END);
        const long bytes = ftell(f);
        fseek(f, 0, SEEK_SET);
        x.resize(bytes / 4);
        if (fread(x.data(), 4, x.size(), f) != x.size()) return 2;
        fclose(f);
    }
    const unsigned n = (unsigned)x.size();
    x.resize(n + 8, 0.f);  // the kernels never read past n; the slack only keeps a violation from crashing first
    // geometry of one song (api.cu geom_of): src/song/mod.rs:435-478, src/utils.rs:30
    SongDesc sd;
    memset(&sd, 0, sizeof(sd));
    sd.n = n;
    sd.n_s = (n - 512) / 128 + 1;
    sd.n_t = (n - 512) / 256 + 1;
    sd.n_c = (unsigned)ceilf((float)n / 2205.f);
    sd.n_c_comp = std::min(sd.n_c, n / 2205 + 1);  // windows(8192).step_by(2205) over the n + 8192 padded samples
    sd.n_l = (n + 1023) / 1024;
    sd.valid = 1;
    const std::vector<SongDesc> songs(1, sd);

    // ---- tables (api.cu build_tables) --------------------------------------------------------------------
    const float PI_F = 3.14159274101257324f;
    std::vector<float> win(512);
    for (int i = 0; i < 512; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * PI_F * (float)i / 512.f));
    std::vector<cpx> twA(16 * 32);
    for (int k1 = 0; k1 < 16; k1++)
        for (int l = 0; l < 32; l++) {
            const double a = -2.0 * M_PI * (double)(k1 * l) / 512.0;
            twA[k1 * 32 + l] = cpx{(float)cos(a), (float)sin(a)};
        }
    PvocTables tab{win.data(), twA.data()};
    std::vector<float> hann(2 * (8192 + 4 * 256));
    for (int i = 0; i < 8192; i++) hann[i] = 0.5f - 0.5f * cosf(2.f * (float)i * PI_F / 8192.f);
    for (int t = 0; t < 256; t++) {
        const double te = 2.0 * M_PI * (double)(2 * t) / 8192.0, to = 2.0 * M_PI * (double)(2 * t + 1) / 8192.0;
        hann[8192 + 4 * t + 0] = (float)cos(te);
        hann[8192 + 4 * t + 1] = (float)cos(to);
        hann[8192 + 4 * t + 2] = (float)sin(te);
        hann[8192 + 4 * t + 3] = (float)sin(to);
    }
    for (int m = 0; m < 8192; m++) hann[9216 + m] = hann[(m + 8191) % 8192];
    for (int t = 0; t < 256; t++) {
        const double ta = 2.0 * M_PI * (double)(2 * t - 1) / 8192.0, tb = 2.0 * M_PI * (double)(2 * t) / 8192.0;
        hann[17408 + 4 * t + 0] = (float)cos(ta);
        hann[17408 + 4 * t + 1] = (float)cos(tb);
        hann[17408 + 4 * t + 2] = (float)sin(ta);
        hann[17408 + 4 * t + 3] = (float)sin(tb);
    }
    hann.resize(2 * (8192 + 4 * 256) + 4 * 128 * 8);  // stft8192v2_kernel's rotated phases (api.cu build_tables)
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
    std::vector<cpx> tw4(4096), tw2(256), tw8(256);
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
        tw8[m] = cpx{(float)cos(a), (float)sin(a)};
    }

    // ---- pvoc512_kernel: the measured build and the experimental cuts --------------------------------------
    const int ppi = 8;  // pairs per work item: several items, so that halo pairs and item boundaries are exercised
    const unsigned items = (sd.n_t + ppi - 1) / ppi;
    const std::vector<unsigned> item_prefix = {0u, items};
    auto run_pvoc = [&](const char *tag, auto kern) {
        if (!stages) return;
        std::vector<float> cen(sd.n_s, -7.f), rol(sd.n_s, -7.f), fla(sd.n_s, -7.f), flux(sd.n_t, -7.f);
        emu::launch((items + 7) / 8, 256, [&] {
            kern(x.data(), songs.data(), item_prefix.data(), 1, items, ppi, tab, cen.data(), rol.data(), fla.data(),
                 flux.data(), nullptr);
        }, 8 * pv2::WARP_SMEM_BYTES);
        dump((std::string("centroid_") + tag).c_str(), cen);
        dump((std::string("rolloff_") + tag).c_str(), rol);
        dump((std::string("flatness_") + tag).c_str(), fla);
        dump((std::string("flux_") + tag).c_str(), flux);
    };
    run_pvoc("default", pvoc512_kernel<true, false>);
    run_pvoc("v512", pvoc512_kernel<true, false, true>);
    run_pvoc("v1024", pvoc512_kernel<true, false, false, true>);
    run_pvoc("v2048", pvoc512_kernel<true, false, false, false, 4>);
    run_pvoc("v3584", pvoc512_kernel<true, false, true, true, 4>);
    run_pvoc("v2", pvoc512v2_kernel);

    // ---- STFT micro-benchmark kernels ----------------------------------------------------------------------
    if (stages) {
        std::vector<float> mags((size_t)sd.n_t * 257, -7.f);
        emu::launch((items + 7) / 8, 256, [&] {
            pvoc512_kernel<false, true>(x.data(), songs.data(), item_prefix.data(), 1, items, ppi, tab, nullptr, nullptr,
                                        nullptr, nullptr, mags.data());
        });
        dump("stft512_default", mags);
        const int fpi = 13;  // an odd item length: pairs straddle nothing, single frames end items
        const unsigned items2 = (sd.n_t + fpi - 1) / fpi;
        const std::vector<unsigned> prefix2 = {0u, items2};
        std::fill(mags.begin(), mags.end(), -7.f);
        emu::launch((items2 + 7) / 8, 256, [&] {
            stft512_pairs_kernel(x.data(), songs.data(), prefix2.data(), 1, items2, fpi, tab, mags.data());
        });
        dump("stft512_v256", mags);
    }

    // ---- timedomain_kernel ---------------------------------------------------------------------------------
    if (stages) {
        const unsigned groups = (sd.n_l + 7) / 8;
        const std::vector<unsigned> gp = {0u, groups};
        std::vector<float> loud(sd.n_l, -7.f), eb(n / 256 + 1, -7.f);
        std::vector<unsigned> zcr(1, 0u);
        emu::launch((groups + 3) / 4, TD_THREADS, [&] {
            timedomain_kernel(x.data(), songs.data(), gp.data(), 1, groups, loud.data(), eb.data(), zcr.data());
        });
        dump("loudness_chunks", loud);
        dump("zcr_count", zcr);
    }

    // ---- pcm_to_mono_kernel: the song as 16-bit stereo (L = x, R = x / 2), 32-bit mono, 3-channel float -----
    if (stages) {
        const size_t frames = (n + 3) / 4 * 4;
        std::vector<short> st(2 * frames, 0);
        std::vector<int> s32(frames, 0);
        std::vector<float> f3(3 * frames, 0.f);
        for (unsigned i = 0; i < n; i++) {
            const short s = (short)lrintf(x[i] * 32767.f);
            st[2 * i] = s;
            st[2 * i + 1] = (short)(s / 2);
            s32[i] = (int)s * 65536 + (int)(i % 251);
            f3[3 * i] = x[i];
            f3[3 * i + 1] = -0.5f * x[i];
            f3[3 * i + 2] = 0.25f;
        }
        std::vector<float> out(frames, -7.f);
        const unsigned grid = (unsigned)((frames / 4 + 255) / 256);
        emu::launch(grid, 256, [&] { pcm_to_mono_kernel<1>(st.data(), out.data(), frames, 2u); });
        dump("mono_from_s16_stereo", out);
        dump("in_s16_stereo", st);
        emu::launch(grid, 256, [&] { pcm_to_mono_kernel<2>(s32.data(), out.data(), frames, 1u); });
        dump("mono_from_s32", out);
        dump("in_s32", s32);
        emu::launch(grid, 256, [&] { pcm_to_mono_kernel<3>(f3.data(), out.data(), frames, 3u); });
        dump("mono_from_f32x3", out);
        dump("in_f32x3", f3);
    }

    // ---- stft8192_kernel: the measured build and the experimental cuts ---------------------------------------
    if (stages) {
        const unsigned ctas = (sd.n_c_comp + K3_FRAMES_PER_CTA - 1) / K3_FRAMES_PER_CTA;
        const std::vector<unsigned> fp = {0u, ctas};
        auto run_stft = [&](const char *tag, auto kern) {
            std::vector<float> mags((size_t)sd.n_c_comp * CH_STRIDE, -7.f);
            std::vector<double> cm((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.), cp((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.);
            std::vector<unsigned> cc(1, 0u);
            emu::launch(ctas, K3_THREADS, [&] {
                kern(x.data(), songs.data(), fp.data(), 1, hann.data(), tw4.data(), tw2.data(), tw8.data(), mags.data(),
                     cm.data(), cp.data(), cc.data());
            });
            std::vector<float> dense((size_t)sd.n_c_comp * CH_BINS);
            for (unsigned f = 0; f < sd.n_c_comp; f++)
                memcpy(&dense[(size_t)f * CH_BINS], &mags[(size_t)f * CH_STRIDE], CH_BINS * 4);
            dump((std::string("stft8192_") + tag).c_str(), dense);
            dump((std::string("peaks_") + tag).c_str(), cc);
            cm.resize(cc[0]);
            cp.resize(cc[0]);
            std::sort(cm.begin(), cm.end());
            std::sort(cp.begin(), cp.end());
            dump((std::string("peak_mags_") + tag).c_str(), cm);
            dump((std::string("peak_pitches_") + tag).c_str(), cp);
        };
        {   // stft8192v2_kernel (round 2): 128 threads, bulk-copy staging, four work items per CTA
            std::vector<float> mags((size_t)sd.n_c_comp * CH_STRIDE, -7.f);
            std::vector<double> cm((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.), cp((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.);
            std::vector<unsigned> cc(1, 0u);
            emu::launch((ctas + s2::ITEMS_PER_CTA - 1) / s2::ITEMS_PER_CTA, s2::THREADS, [&] {
                stft8192v2_kernel(x.data(), songs.data(), fp.data(), 1, ctas, K3_FRAMES_PER_CTA, s2::ITEMS_PER_CTA, hann.data(), tw4.data(), tw2.data(),
                                  tw8.data(), mags.data(), cm.data(), cp.data(), cc.data());
            }, s2::SMEM_BYTES);
            std::vector<float> dense((size_t)sd.n_c_comp * CH_BINS);
            for (unsigned f = 0; f < sd.n_c_comp; f++)
                memcpy(&dense[(size_t)f * CH_BINS], &mags[(size_t)f * CH_STRIDE], CH_BINS * 4);
            dump("stft8192_v2", dense);
            dump("peaks_v2", cc);
            cm.resize(cc[0]);
            cp.resize(cc[0]);
            std::sort(cm.begin(), cm.end());
            std::sort(cp.begin(), cp.end());
            dump("peak_mags_v2", cm);
            dump("peak_pitches_v2", cp);
        }
        {   // stft8192v3_kernel (round 2): 256 threads, one column per thread, mirror halves exchanged through the buffer
            std::vector<float> mags((size_t)sd.n_c_comp * CH_STRIDE, -7.f);
            std::vector<double> cm((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.), cp((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.);
            std::vector<unsigned> cc(1, 0u);
            emu::launch((ctas + s3::ITEMS_PER_CTA - 1) / s3::ITEMS_PER_CTA, s3::THREADS, [&] {
                stft8192v3_kernel(x.data(), songs.data(), fp.data(), 1, ctas, K3_FRAMES_PER_CTA, s3::ITEMS_PER_CTA, hann.data(), tw4.data(), tw2.data(),
                                  tw8.data(), mags.data(), cm.data(), cp.data(), cc.data());
            }, s3::SMEM_BYTES);
            std::vector<float> dense((size_t)sd.n_c_comp * CH_BINS);
            for (unsigned f = 0; f < sd.n_c_comp; f++)
                memcpy(&dense[(size_t)f * CH_BINS], &mags[(size_t)f * CH_STRIDE], CH_BINS * 4);
            dump("stft8192_v3", dense);
            dump("peaks_v3", cc);
            cm.resize(cc[0]);
            cp.resize(cc[0]);
            std::sort(cm.begin(), cm.end());
            std::sort(cp.begin(), cp.end());
            dump("peak_mags_v3", cm);
            dump("peak_pitches_v3", cp);
        }
        run_stft("default", stft8192_kernel<true>);
        run_stft("v64", stft8192_kernel<true, K3V_TWPROD>);
        run_stft("v128", stft8192_kernel<true, K3V_WINSYN>);
        run_stft("v4096", stft8192_kernel<true, K3V_LAY16>);
        run_stft("v4288", stft8192_kernel<true, K3V_TWPROD | K3V_WINSYN | K3V_LAY16>);
        run_stft("v8192", stft8192_kernel<true, K3V_ODDSHIFT>);
        run_stft("v12480", stft8192_kernel<true, K3V_TWPROD | K3V_WINSYN | K3V_LAY16 | K3V_ODDSHIFT>);
        run_stft("old_epilogue", stft8192_kernel<false>);
    }
    // ---- the whole path, kernel after kernel as run_wave (api.cu) enqueues them: 23 features ------------------
    {
        std::vector<double> table((size_t)100 * CH_BINS * 12, 0.);  // only the row of the estimated tuning is filled
        std::vector<float> table32(table.size(), 0.f);
        auto full = [&](const char *tag, auto pvoc_kern, auto stft_kern) {
            // tempo / timbral chain
            const unsigned groups = (sd.n_l + 7) / 8;
            const std::vector<unsigned> gp = {0u, groups}, tp = {0u, sd.n_t};
            std::vector<float> loud(sd.n_l, 0.f), eb(n / 256 + 1, 0.f), cen(sd.n_s), rol(sd.n_s), fla(sd.n_s), flux(sd.n_t),
                thr(sd.n_t), bpm(sd.n_t / 16 + 16, 0.f), tempo(1, 0.f);
            std::vector<unsigned> zcr(1, 0u), nbpm(1, 0u);
            emu::launch((groups + 3) / 4, TD_THREADS, [&] {
                timedomain_kernel(x.data(), songs.data(), gp.data(), 1, groups, loud.data(), eb.data(), zcr.data());
            });
            emu::launch((items + 7) / 8, 256, [&] {
                pvoc_kern(x.data(), songs.data(), item_prefix.data(), 1, items, ppi, tab, cen.data(), rol.data(), fla.data(),
                          flux.data(), nullptr);
            }, 8 * pv2::WARP_SMEM_BYTES);
            emu::launch((sd.n_t + 255) / 256, 256, [&] { peakpick_kernel(flux.data(), songs.data(), tp.data(), 1, sd.n_t, thr.data()); });
            emu::launch(1, 128, [&] {
                beattrack_kernel<128>(thr.data(), eb.data(), songs.data(), bpm.data(), tempo.data(), nbpm.data(), 0);
            });
            {   // the three autocorrelation cuts of beattrack_kernel (0 balanced lag pairs, 1 one lag at a time, 2 four
                // consecutive lags per thread) keep every lag's sum in the reference's order: the same bits
                std::vector<float> all;
                for (int mode = 0; mode < 3; mode++) {
                    std::vector<float> b2(bpm.size(), 0.f), t2(1, 0.f);
                    std::vector<unsigned> n2(1, 0u);
                    emu::launch(1, 128, [&] {
                        beattrack_kernel<128>(thr.data(), eb.data(), songs.data(), b2.data(), t2.data(), n2.data(), mode);
                    });
                    all.push_back(t2[0]);
                    all.push_back((float)n2[0]);
                    for (unsigned i = 0; i < n2[0] && i < 64; i++) all.push_back(b2[i]);
                    all.push_back(-12345.f);
                }
                dump((std::string("acf_modes_") + tag).c_str(), all);
            }
            // chroma chain
            const unsigned ctas = (sd.n_c_comp + K3_FRAMES_PER_CTA - 1) / K3_FRAMES_PER_CTA;
            const std::vector<unsigned> fp = {0u, ctas};
            std::vector<float> mags(((size_t)sd.n_c_comp + CH_TILE_FRAMES) * CH_STRIDE, 0.f);
            std::vector<double> cm((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.), cp((size_t)sd.n_c_comp * CH_MAX_PEAKS, 0.);
            std::vector<unsigned> cc(1, 0u);
            emu::launch(ctas, K3_THREADS, [&] {
                stft_kern(x.data(), songs.data(), fp.data(), 1, hann.data(), tw4.data(), tw2.data(), tw8.data(), mags.data(),
                          cm.data(), cp.data(), cc.data());
            });
            std::vector<int> tuning(1, -1);
            emu::launch(1, K4_THREADS, [&] { tuning_select_kernel(cm.data(), cp.data(), cc.data(), songs.data(), tuning.data()); });
            if (tuning[0] < 0 || tuning[0] > 99) { printf("tuning index %d\n", tuning[0]); exit(3); }
            {   // the filterbank of that tuning (chroma_filter_table_kernel computes row blockIdx.y of the table)
                dim3 grid;
                grid.x = (CH_BINS + 127) / 128;
                emu::bid().y = 0;
                for (unsigned bx = 0; bx < grid.x; bx++) {  // one row only: blockIdx.y is set by hand
                    emu::launch(1, 128, [&] {
                        emu::bid().x = bx;
                        emu::bid().y = (unsigned)tuning[0];
                        chroma_filter_table_kernel(table.data());
                    });
                }
                emu::bid().y = 0;
                const size_t lo = (size_t)tuning[0] * CH_BINS * 12;
                for (size_t i = lo; i < lo + (size_t)CH_BINS * 12; i++) table32[i] = (float)table[i];  // f64_to_f32_kernel
            }
            const unsigned tiles = (sd.n_c + CH_TILE_FRAMES - 1) / CH_TILE_FRAMES;
            const std::vector<unsigned> tlp = {0u, tiles};
            std::vector<double> partials((size_t)tiles * 10, 0.);
            emu::launch(tiles, K5_THREADS, [&] {
                chroma_pipe_kernel(mags.data(), songs.data(), tlp.data(), 1, table32.data(), tuning.data(), partials.data(), nullptr);
            }, K5P_SMEM);
            // summary
            PeerRows peers;
            memset(&peers, 0, sizeof(peers));
            std::vector<float> out(24, 0.f);
            emu::launch(1, K9_THREADS, [&] {
                finalize_kernel(songs.data(), cen.data(), rol.data(), fla.data(), loud.data(), zcr.data(), tempo.data(),
                                partials.data(), 2, out.data(), 0u, peers);
            });
            out.resize(23);
            dump((std::string("features_") + tag).c_str(), out);
            std::vector<float> misc = {(float)tuning[0], (float)nbpm[0], tempo[0]};
            dump((std::string("misc_") + tag).c_str(), misc);
        };
        if (!batch_mode) full("default", pvoc512_kernel<true, false>, stft8192_kernel<true>);
        if (!batch_mode) full("all_cuts", pvoc512_kernel<true, false, true, true, 4>, stft8192_kernel<true, K3V_TWPROD | K3V_WINSYN | K3V_LAY16 | K3V_ODDSHIFT>);
    }
    // ---- a ragged batch: "batch n1 n2 ..." cuts the file into consecutive songs (4-sample aligned starts, a too short
    //      one allowed) and runs the whole path over all of them at once, descriptors and prefix arrays laid out as
    //      plan_wave (api.cu) lays them out: song lookup, item boundaries and per-song offsets of every kernel ----------
    if (argc > 4 && std::string(argv[3]) == "batch") {
        std::vector<SongDesc> bs;
        std::vector<unsigned> k1p(1, 0u), chp(1, 0u), tpp(1, 0u), prp(1, 0u), tlp(1, 0u);
        size_t off = 0, rows = 0, cands = 0, ns = 0, nt = 0, nl = 0, neb = 0, tiles = 0, bpms = 0;
        const int bppi = 16;
        for (int a = 4; a < argc; a++) {
            const unsigned len = (unsigned)atoi(argv[a]);
            SongDesc d;
            memset(&d, 0, sizeof(d));
            d.pcm_off = off;
            d.n = len;
            d.valid = len >= (unsigned)MIN_SAMPLES ? 1u : 0u;
            unsigned q_ns = 0, q_nt = 0, q_nc = 0, q_ncc = 0, q_nl = 0, q_tiles = 0;
            if (d.valid) {
                q_ns = (len - 512) / 128 + 1;
                q_nt = (len - 512) / 256 + 1;
                q_nc = (unsigned)ceilf((float)len / 2205.f);
                q_ncc = std::min(q_nc, len / 2205 + 1);
                q_nl = (len + 1023) / 1024;
                q_tiles = (q_nc + CH_TILE_FRAMES - 1) / CH_TILE_FRAMES;
                d.n_s = q_ns; d.n_t = q_nt; d.n_c = q_nc; d.n_c_comp = q_ncc; d.n_l = q_nl;
            }
            d.mag_off = rows; d.cand_off = cands;
            d.s_off = (unsigned)ns; d.t_off = (unsigned)nt; d.l_off = (unsigned)nl; d.e_off = (unsigned)neb;
            d.c_tile_off = (unsigned)tiles; d.bpm_off = (unsigned)bpms;
            k1p.push_back(k1p.back() + (d.valid ? (q_nt + bppi - 1) / bppi : 0));
            chp.push_back(chp.back() + (d.valid ? (q_nl + 7) / 8 : 0));
            tpp.push_back(tpp.back() + q_nt);
            prp.push_back(prp.back() + (d.valid ? (q_ncc + 3) / 4 : 0));
            tlp.push_back(tlp.back() + q_tiles);
            if (d.valid) {
                rows += q_ncc; cands += (size_t)q_ncc * CH_MAX_PEAKS; ns += q_ns; nt += q_nt; nl += q_nl; neb += len / 256;
                tiles += q_tiles; bpms += q_nt / 16 + 16;
            }
            bs.push_back(d);
            off += (len + 3) / 4 * 4;
        }
        if (off > n + 8) { fprintf(stderr, "batch longer than the file\n"); return 2; }
        const int k = (int)bs.size();
        std::vector<double> table((size_t)100 * CH_BINS * 12, 0.);
        std::vector<float> table32(table.size(), 0.f);
        auto batch = [&](const char *tag, auto pvoc_kern, auto stft_kern) {
            std::vector<float> loud(nl + 1, 0.f), eb(neb + 1, 0.f), cen(ns + 1), rol(ns + 1), fla(ns + 1), flux(nt + 1), thr(nt + 1),
                bpm(bpms + 1, 0.f), tempo(k, 0.f), mags((rows + CH_TILE_FRAMES) * CH_STRIDE, 0.f), out((size_t)k * 23, 0.f);
            std::vector<unsigned> zcr(k, 0u), nbpm(k, 0u), cc(k, 0u);
            std::vector<double> cm(cands + 1, 0.), cp(cands + 1, 0.), partials(tiles * 10 + 10, 0.);
            std::vector<int> tuning(k, 0);
            emu::launch((chp[k] + 3) / 4, TD_THREADS, [&] { timedomain_kernel(x.data(), bs.data(), chp.data(), k, chp[k], loud.data(), eb.data(), zcr.data()); });
            emu::launch((k1p[k] + 7) / 8, 256, [&] {
                pvoc_kern(x.data(), bs.data(), k1p.data(), k, k1p[k], bppi, tab, cen.data(), rol.data(), fla.data(), flux.data(), nullptr);
            }, 8 * pv2::WARP_SMEM_BYTES);
            emu::launch((tpp[k] + 255) / 256, 256, [&] { peakpick_kernel(flux.data(), bs.data(), tpp.data(), k, tpp[k], thr.data()); });
            emu::launch((unsigned)k, 128, [&] { beattrack_kernel<128>(thr.data(), eb.data(), bs.data(), bpm.data(), tempo.data(), nbpm.data(), 0); });
            emu::launch(prp[k], K3_THREADS, [&] {
                stft_kern(x.data(), bs.data(), prp.data(), k, hann.data(), tw4.data(), tw2.data(), tw8.data(), mags.data(), cm.data(),
                          cp.data(), cc.data());
            });
            emu::launch((unsigned)k, K4_THREADS, [&] { tuning_select_kernel(cm.data(), cp.data(), cc.data(), bs.data(), tuning.data()); });
            for (int i = 0; i < k; i++) {
                if (!bs[i].valid) continue;
                for (unsigned bx = 0; bx < (CH_BINS + 127) / 128; bx++)
                    emu::launch(1, 128, [&] {
                        emu::bid().x = bx;
                        emu::bid().y = (unsigned)tuning[i];
                        chroma_filter_table_kernel(table.data());
                    });
                emu::bid().y = 0;
                const size_t lo = (size_t)tuning[i] * CH_BINS * 12;
                for (size_t e = lo; e < lo + (size_t)CH_BINS * 12; e++) table32[e] = (float)table[e];
            }
            emu::launch(tlp[k], K5_THREADS, [&] {
                chroma_pipe_kernel(mags.data(), bs.data(), tlp.data(), k, table32.data(), tuning.data(), partials.data(), nullptr);
            }, K5P_SMEM);
            PeerRows peers;
            memset(&peers, 0, sizeof(peers));
            emu::launch((unsigned)k, K9_THREADS, [&] {
                finalize_kernel(bs.data(), cen.data(), rol.data(), fla.data(), loud.data(), zcr.data(), tempo.data(), partials.data(), 2,
                                out.data(), 0u, peers);
            });
            dump((std::string("batch_features_") + tag).c_str(), out);
        };
        batch("default", pvoc512_kernel<true, false>, stft8192_kernel<true>);
        batch("all_cuts", pvoc512_kernel<true, false, true, true, 4>, stft8192_kernel<true, K3V_TWPROD | K3V_WINSYN | K3V_LAY16 | K3V_ODDSHIFT>);
        printf("batch of %d songs\nOK\n", k);
        return 0;
    }
    printf("n %u n_s %u n_t %u n_c %u n_c_comp %u n_l %u\nOK\n", n, sd.n_s, sd.n_t, sd.n_c, sd.n_c_comp, sd.n_l);
    return 0;
}

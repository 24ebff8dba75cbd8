// Host TRANSCRIPTION of stft512_pairs_kernel's loop (bliss-rs_b200/csrc/spectral.cu; one warp = 32 loop
// iterations per phase, the FFT phases are the kernel's own __host__ __device__ functions of pvoc512.cuh).
// Checks the design of the hop-256 pairing: which window rows feed which frame, the slide by 16 rows, that
// every unguarded load of the kernel stays inside the song for any item length, that every bin of every
// frame is written exactly by the pair that owns it, digital silence -> exact zeros, and the magnitudes
// against an f64 DFT of the PVocTempo framing (src/aubio.rs:338-425).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "../../bliss-rs_b200/csrc/pvoc512.cuh"
using namespace bliss;
int main() {
    const int n = 256 * 37 + 123 + 512;  // arbitrary length
    const int n_t = (n - 512) / 256 + 1;
    std::vector<float> x(n);
    srand(3);
    for (auto &v : x) v = (float)(rand() / (double)RAND_MAX - 0.5);
    for (int i = 3000; i < 3600; i++) x[i] = 0.f;  // a stretch of digital silence
    const float PI_F = 3.14159274101257324f;
    std::vector<float> win(512);
    for (int i = 0; i < 512; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * PI_F * (float)i / 512.f));
    std::vector<cpx> twA(16 * 32);
    for (int k1 = 0; k1 < 16; k1++)
        for (int l = 0; l < 32; l++) { double a = -2.0 * M_PI * (double)(k1 * l) / 512.0; twA[k1 * 32 + l] = cpx{(float)cos(a), (float)sin(a)}; }
    std::vector<float> out((size_t)n_t * 257, -1.f);
    double worst = 0;
    for (int frames_per_item : {8, 13, 64}) {
        std::fill(out.begin(), out.end(), -1.f);
        for (int j0 = 0; j0 < n_t; j0 += frames_per_item) {
            const int j1 = std::min(j0 + frames_per_item, n_t);
            float s[32][24];
            for (int lane = 0; lane < 32; lane++) {
                const int base = 256 * j0 - 256 + lane;
                for (int m = 0; m < 24; m++) { const int idx = base + 32 * m; s[lane][m] = (idx >= 0 && idx < n) ? x[idx] : 0.f; }
            }
            for (int j = j0; j < j1; j += 2) {
                cpx r[32][16];
                float pka = 0, pkb = 0;
                for (int lane = 0; lane < 32; lane++)
                    for (int n1 = 0; n1 < 16; n1++) {
                        const float w = win[lane + 32 * n1];
                        r[lane][n1] = cpx{s[lane][n1] * w, s[lane][n1 + 8] * w};
                        pka = fmaxf(pka, fabsf(r[lane][n1].x)); pkb = fmaxf(pkb, fabsf(r[lane][n1].y));
                    }
                for (int lane = 0; lane < 32; lane++) {
                    for (int m = 0; m < 8; m++) s[lane][m] = s[lane][m + 16];
                    if (j + 2 < j1) {
                        const int base = 256 * (j + 2) - 256 + lane + 32 * 8;
                        for (int m = 0; m < 8; m++) { if (base + 32 * m < 0 || base + 32 * m >= n) { printf("OOB unguarded load\n"); return 2; } s[lane][8 + m] = x[base + 32 * m]; }
                        if (j + 3 < n_t) {
                            for (int m = 8; m < 16; m++) { if (base + 32 * m >= n) { printf("OOB unguarded load (B)\n"); return 2; } s[lane][8 + m] = x[base + 32 * m]; }
                        } else {
                            for (int m = 8; m < 16; m++) { const int idx = base + 32 * m; s[lane][8 + m] = (idx < n) ? x[idx] : 0.f; }
                        }
                    }
                }
                unsigned ua, ub; memcpy(&ua, &pka, 4); memcpy(&ub, &pkb, 4);
                int sh = (int)(ua >> 23) - (int)(ub >> 23);
                sh = (ua == 0u || ub == 0u) ? 0 : std::max(-60, std::min(60, sh));
                unsigned gs = (unsigned)(127 + sh) << 23, gi = (unsigned)(127 - sh) << 23;
                float gscale, ginv; memcpy(&gscale, &gs, 4); memcpy(&ginv, &gi, 4);
                if (sh != 0) for (int lane = 0; lane < 32; lane++) for (int n1 = 0; n1 < 16; n1++) r[lane][n1].y *= gscale;
                std::vector<cpx> S(pv::EXCH_CPX);
                for (int lane = 0; lane < 32; lane++) pv::phase_a(lane, r[lane], twA.data(), S.data());
                for (int lane = 0; lane < 32; lane++) { pv::phase_b_load(lane, r[lane], S.data()); pv::phase_b_fft(lane, r[lane]); }
                std::vector<cpx> Z(pv::EXCH_CPX);
                for (int lane = 0; lane < 32; lane++)
                    for (int q = 0; q < 16; q++) Z[pv::zpos_s<0>(pv::bin_of(lane, q))] = pv::phase_b_combine(lane, r[lane][q], r[lane ^ 16][q]);
                const float ka = (ua == 0u) ? 0.f : 0.5f, kb = (ub == 0u) ? 0.f : 0.5f * ginv;
                for (int lane = 0; lane < 32; lane++) {
                    float ma[8], mb[8];
                    for (int i = 0; i < 8; i++) { const int k = lane + 32 * i; pv::untangle_mag<true>(Z[pv::zpos_s<0>(k)], Z[pv::zpos_s<0>((512 - k) & 511)], ma[i], mb[i]); }
                    const cpx zn = Z[pv::zpos_s<0>(256)];
                    float nyq_a = 2.f * fabsf(zn.x), nyq_b = 2.f * fabsf(zn.y);
                    if (lane == 0) { ma[0] = 2.f * fabsf(Z[0].x); mb[0] = 2.f * fabsf(Z[0].y); }
                    float *oa = out.data() + (size_t)j * 257;
                    for (int i = 0; i < 8; i++) oa[lane + 32 * i] = ma[i] * ka;
                    if (lane == 0) oa[256] = nyq_a * ka;
                    if (j + 1 < j1) { float *ob = oa + 257; for (int i = 0; i < 8; i++) ob[lane + 32 * i] = mb[i] * kb; if (lane == 0) ob[256] = nyq_b * kb; }
                }
            }
        }
        // reference: f64 DFT of each windowed frame (PVocTempo framing, zeros before the song)
        double err = 0;
        for (int m = 0; m < n_t; m++) {
            double fr[512]; double scale = 1e-30;
            for (int i = 0; i < 512; i++) { const int idx = 256 * (m + 1) - 512 + i; fr[i] = (idx >= 0 ? (double)x[idx] : 0.0) * (double)win[i]; }
            std::vector<double> mag(257);
            for (int k = 0; k <= 256; k++) {
                double re = 0, im = 0;
                for (int i = 0; i < 512; i++) { double a = -2.0 * M_PI * (double)((i * k) % 512) / 512.0; re += fr[i] * cos(a); im += fr[i] * sin(a); }
                mag[k] = sqrt(re * re + im * im); scale = fmax(scale, mag[k]);
            }
            bool silent = true; for (int i = 0; i < 512; i++) if (fr[i] != 0.0) silent = false;
            for (int k = 0; k <= 256; k++) {
                const float g = out[(size_t)m * 257 + k];
                if (g < 0) { printf("frame %d bin %d never written\n", m, k); return 3; }
                if (silent && g != 0.f) { printf("silent frame %d not exactly zero\n", m); return 4; }
                err = fmax(err, fabs(g - mag[k]) / scale);
            }
        }
        printf("frames_per_item %d: max rel err %.3e over %d frames\n", frames_per_item, err, n_t);
        worst = fmax(worst, err);
    }
    if (worst > 2e-6) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}

// Host-only emulation of the warp / CTA FFT index logic of the CUDA kernels.
// Built with nvcc but launches nothing: every "lane" / "thread" is a loop iteration.
// Prints max relative error vs an f64 DFT for (a) the 512-point two-frame warp FFT
// incl. untangling, (b) the 8192-point three-pass FFT incl. untangling.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../bliss-rs_b200/csrc/rfft8192.cuh"
#include "../../bliss-rs_b200/csrc/pvoc512.cuh"

using namespace bliss;

static void dft_real(const std::vector<double> &x, std::vector<double> &mag, int n) {
    mag.assign(n / 2 + 1, 0.0);
    for (int k = 0; k <= n / 2; k++) {
        double re = 0, im = 0;
        for (int i = 0; i < n; i++) {
            double a = -2.0 * M_PI * (double)((long long)i * k % n) / n;
            re += x[i] * cos(a);
            im += x[i] * sin(a);
        }
        mag[k] = sqrt(re * re + im * im);
    }
}

int main() {
    srand(1);
    double worst = 0;
    // ---------------- 512-point, two frames per warp ----------------
    {
        std::vector<cpx> twA(16 * 32);
        for (int k1 = 0; k1 < 16; k1++)
            for (int l = 0; l < 32; l++) {
                double a = -2.0 * M_PI * (double)(k1 * l) / 512.0;
                twA[k1 * 32 + l] = cpx{(float)cos(a), (float)sin(a)};
            }
        std::vector<double> a(512), b(512);
        for (int i = 0; i < 512; i++) {
            a[i] = (rand() / (double)RAND_MAX - 0.5);
            b[i] = (rand() / (double)RAND_MAX - 0.5) * 0.3;
        }
        std::vector<cpx> S(pv::EXCH_CPX), Z(pv::EXCH_CPX);
        cpx regs[32][16];
        for (int lane = 0; lane < 32; lane++) {
            cpx r[16];
            for (int n1 = 0; n1 < 16; n1++) r[n1] = cpx{(float)a[lane + 32 * n1], (float)b[lane + 32 * n1]};
            pv::phase_a(lane, r, twA.data(), S.data());
        }
        for (int lane = 0; lane < 32; lane++) {
            pv::phase_b_load(lane, regs[lane], S.data());
            pv::phase_b_fft(lane, regs[lane]);
        }
        for (int lane = 0; lane < 32; lane++)
            for (int q = 0; q < 16; q++) {
                cpx z = pv::phase_b_combine(lane, regs[lane][q], regs[lane ^ 16][q]);
                Z[pv::zpos(pv::bin_of(lane, q))] = z;
            }
        std::vector<double> ma, mb;
        dft_real(a, ma, 512);
        dft_real(b, mb, 512);
        double scale = 0;
        for (int k = 0; k <= 256; k++) scale = fmax(scale, fmax(ma[k], mb[k]));
        double err = 0;
        for (int k = 0; k <= 256; k++) {
            float fa, fb;
            pv::untangle_mag(Z[pv::zpos(k)], Z[pv::zpos((512 - k) & 511)], fa, fb);
            err = fmax(err, fabs(fa - ma[k]) / scale);
            err = fmax(err, fabs(fb - mb[k]) / scale);
        }
        printf("fft512 pair: max rel err %.3e\n", err);
        worst = fmax(worst, err);
        // the other paddings of the natural-order tile (zpos_s<4>: VARIANT_PV_ZPOS4, zpos_s<0>: stft512_pairs_kernel)
        // are injective, fit the tile, and carry the same untangling
        {
            std::vector<int> seen4(pv::EXCH_CPX, 0), seen0(pv::EXCH_CPX, 0);
            for (int k = 0; k < 512; k++) {
                const int p4 = pv::zpos_s<4>(k), p0 = pv::zpos_s<0>(k);
                if (p4 >= pv::EXCH_CPX || p0 >= pv::EXCH_CPX || seen4[p4]++ || seen0[p0]++) { printf("zpos_s collision at %d\n", k); return 13; }
            }
            std::vector<cpx> Z4(pv::EXCH_CPX);
            for (int k = 0; k < 512; k++) Z4[pv::zpos_s<4>(k)] = Z[pv::zpos(k)];
            for (int k = 0; k <= 256; k++) {
                float fa, fb, ga, gb;
                pv::untangle_mag(Z[pv::zpos(k)], Z[pv::zpos((512 - k) & 511)], fa, fb);
                pv::untangle_mag(Z4[pv::zpos_s<4>(k)], Z4[pv::zpos_s<4>((512 - k) & 511)], ga, gb);
                if (fa != ga || fb != gb) { printf("zpos_s<4> changes bin %d\n", k); return 14; }
            }
        }
        // VARIANT_PV_TWPROD: phase A with product twiddles stores (nearly) what the table form stores
        {
            std::vector<cpx> S2(pv::EXCH_CPX);
            for (int lane = 0; lane < 32; lane++) {
                cpx r[16], u[16];
                for (int n1 = 0; n1 < 16; n1++) r[n1] = u[n1] = cpx{(float)a[lane + 32 * n1], (float)b[lane + 32 * n1]};
                pv::phase_a(lane, r, twA.data(), S.data());
                pv::phase_a_prod(lane, u, twA[1 * 32 + lane], twA[2 * 32 + lane], twA[4 * 32 + lane], twA[8 * 32 + lane], S2.data());
            }
            double dmax = 0, vmax = 0;
            for (int k1 = 0; k1 < 16; k1++)
                for (int lane = 0; lane < 32; lane++) {
                    const cpx p = S[k1 * pv::ROW + lane], q = S2[k1 * pv::ROW + lane];
                    dmax = fmax(dmax, fmax(fabs(p.x - q.x), fabs(p.y - q.y)));
                    vmax = fmax(vmax, fmax(fabs(p.x), fabs(p.y)));
                }
            printf("fft512 phase-A product twiddles: max diff %.3e of %.3e\n", dmax, vmax);
            if (!(dmax <= 1e-6 * vmax)) { printf("phase-A product twiddles disagree\n"); return 11; }
        }
    }
    // ---------------- 8192-point real FFT as 4096 complex, one frame per CTA ----------------
    {
        std::vector<cpx> tw(4096), tw2(256), tw8(256);  // tw: [k1][b] = W4096^(b k1); tw2: [k2][j] = W256^(j k2)
        for (int k1 = 0; k1 < 16; k1++)
            for (int b = 0; b < 256; b++) {
                double a = -2.0 * M_PI * (double)(b * k1) / 4096.0;
                tw[k1 * 256 + b] = cpx{(float)cos(a), (float)sin(a)};
            }
        for (int k2 = 0; k2 < 16; k2++)
            for (int j = 0; j < 16; j++) {
                double a = -2.0 * M_PI * (double)(j * k2) / 256.0;
                tw2[k2 * 16 + j] = cpx{(float)cos(a), (float)sin(a)};
            }
        for (int t = 0; t < 256; t++) {
            double a = -2.0 * M_PI * (double)t / 8192.0;
            tw8[t] = cpx{(float)cos(a), (float)sin(a)};
        }
        std::vector<double> a(8192);
        for (int i = 0; i < 8192; i++) a[i] = (rand() / (double)RAND_MAX - 0.5);
        std::vector<cpx> buf(r8k::BUF_CPX);
        for (int bb = 0; bb < 256; bb++) {
            cpx v[16];
            for (int q = 0; q < 16; q++) {
                const int nn = bb + 256 * q;
                v[q] = cpx{(float)a[2 * nn], (float)a[2 * nn + 1]};
            }
            r8k::pass1_store(bb, v, tw.data(), buf.data());
        }
        for (int bb = 0; bb < 256; bb++) r8k::pass2(bb, tw2.data(), buf.data());
        for (int bb = 0; bb < 256; bb++) r8k::pass3(bb, buf.data());
        std::vector<double> ma;
        dft_real(a, ma, 8192);
        double scale = 0;
        for (int k = 0; k <= 4096; k++) scale = fmax(scale, ma[k]);
        double err = 0;
        // epilogue exactly as the kernel addresses it: thread t owns bins t + 256 m
        for (int t = 0; t < 256; t++) {
            const cpx *pk = buf.data() + r8k::zbase(t);
            for (int m = 0; m < 16; m++) {
                const int k = t + 256 * m;
                const cpx zk = pk[m];
                cpx zm;
                if (t == 0) zm = buf[r8k::zbase(0) + ((16 - m) & 15)];
                else zm = buf[r8k::zbase(256 - t) + 15 - m];
                const cpx chk = r8k::z_value(buf.data(), (4096 - k) & 4095);
                if (zm.x != chk.x || zm.y != chk.y) { printf("mirror addressing mismatch k=%d\n", k); return 3; }
                const double ang = -2.0 * M_PI * (double)k / 8192.0;
                const cpx w = cmul(tw8[t], cpx{(float)cos(-2.0 * M_PI * m / 32.0), (float)sin(-2.0 * M_PI * m / 32.0)});
                (void)ang;
                const float mg = r8k::untangle_mag(zk, zm, w);
                err = fmax(err, fabs(mg - ma[k]) / scale);
            }
        }
        {
            const cpx z0 = buf[r8k::zbase(0)];
            const float mg = r8k::untangle_mag(z0, z0, cpx{-1.f, 0.f});
            err = fmax(err, fabs(mg - ma[4096]) / scale);
        }
        printf("rfft8192 (4096 complex): max rel err %.3e\n", err);
        worst = fmax(worst, err);
        // ---- fused pass 3 + pair epilogue (pass3_regs / untangle_mag_pair), thread by thread ----
        {
            std::vector<cpx> buf2(r8k::BUF_CPX);
            for (int bb = 0; bb < 256; bb++) {
                cpx v[16];
                for (int q = 0; q < 16; q++) {
                    const int nn = bb + 256 * q;
                    v[q] = cpx{(float)a[2 * nn], (float)a[2 * nn + 1]};
                }
                r8k::pass1_store(bb, v, tw.data(), buf2.data());
            }
            for (int bb = 0; bb < 256; bb++) r8k::pass2(bb, tw2.data(), buf2.data());
            static cpx regs[256][16];
            // poison what pass3_regs does not publish, so that a read of an unpublished slot shows up
            for (int t = 0; t < 256; t++) r8k::pass3_regs(t, regs[t], buf2.data());
            for (int t = 0; t < 256; t++)
                for (int m = (t == 0 ? 1 : 0); m < 8; m++) buf2[r8k::zbase(t) + m] = cpx{NAN, NAN};
            std::vector<int> hit(4097, 0);
            double err2 = 0;
            for (int t = 0; t < 256; t++) {
                const cpx *pm = buf2.data() + r8k::zbase((256 - t) & 255) + 15;
                for (int M = 0; M < 8; M++) {
                    const int k = t + 256 * M;
                    const cpx zk = regs[t][bitrev(M, 4)];
                    const cpx zm = (t == 0) ? buf2[(16 - M) & 15] : pm[-M];
                    const cpx w = cmul(tw8[t], cpx{(float)cos(-2.0 * M_PI * M / 32.0), (float)sin(-2.0 * M_PI * M / 32.0)});
                    float mk, mm;
                    r8k::untangle_mag_pair(zk, zm, w, mk, mm);
                    hit[k]++;
                    hit[4096 - k]++;
                    err2 = fmax(err2, fabs(mk - ma[k]) / scale);
                    err2 = fmax(err2, fabs(mm - ma[4096 - k]) / scale);
                }
            }
            {
                const cpx z = regs[0][bitrev(8, 4)];
                const float mg = r8k::untangle_mag(z, z, cpx{0.f, -1.f});
                hit[2048]++;
                err2 = fmax(err2, fabs(mg - ma[2048]) / scale);
            }
            for (int k = 0; k <= 4096; k++)
                if (hit[k] != 1) { printf("pair epilogue: bin %d produced %d times\n", k, hit[k]); return 4; }
            if (!(err2 == err2)) { printf("pair epilogue read an unpublished slot\n"); return 5; }
            printf("rfft8192 fused pass3 + pair epilogue: max rel err %.3e\n", err2);
            worst = fmax(worst, err2);
        }
        // ---- VARIANT_LAY16: the same three passes + pair epilogue on the buffer without the per-16 padding ----
        {
            std::vector<cpx> buf16(r8k::BUF_CPX, cpx{NAN, NAN});
            for (int bb = 0; bb < 256; bb++) {
                cpx v[16];
                for (int q = 0; q < 16; q++) {
                    const int nn = bb + 256 * q;
                    v[q] = cpx{(float)a[2 * nn], (float)a[2 * nn + 1]};
                }
                r8k::pass1_store<16>(bb, v, tw.data(), buf16.data());
            }
            for (int bb = 0; bb < 256; bb++) r8k::pass2<16>(bb, tw2.data(), buf16.data());
            static cpx regs16[256][16];
            for (int t = 0; t < 256; t++) r8k::pass3_regs<16>(t, regs16[t], buf16.data());
            for (int t = 0; t < 256; t++)
                for (int m = (t == 0 ? 1 : 0); m < 8; m++) buf16[r8k::zbase<16>(t) + m] = cpx{NAN, NAN};
            std::vector<int> hit(4097, 0);
            double err16 = 0;
            bool same = true;
            for (int t = 0; t < 256; t++) {
                const cpx *pm = buf16.data() + r8k::zbase<16>((256 - t) & 255) + 15;
                for (int M = 0; M < 8; M++) {
                    const int k = t + 256 * M;
                    const cpx zk = regs16[t][bitrev(M, 4)];
                    const cpx zm = (t == 0) ? buf16[(16 - M) & 15] : pm[-M];
                    const cpx chk = r8k::z_value(buf.data(), (4096 - k) & 4095);  // the padded layout's value of the same bin
                    if (zm.x != chk.x || zm.y != chk.y) same = false;
                    const cpx w = cmul(tw8[t], cpx{(float)cos(-2.0 * M_PI * M / 32.0), (float)sin(-2.0 * M_PI * M / 32.0)});
                    float mk, mm;
                    r8k::untangle_mag_pair(zk, zm, w, mk, mm);
                    hit[k]++;
                    hit[4096 - k]++;
                    err16 = fmax(err16, fabs(mk - ma[k]) / scale);
                    err16 = fmax(err16, fabs(mm - ma[4096 - k]) / scale);
                }
            }
            hit[2048]++;
            for (int k = 0; k <= 4096; k++)
                if (hit[k] != 1) { printf("LAY16 epilogue: bin %d produced %d times\n", k, hit[k]); return 15; }
            if (!same) { printf("LAY16 mirror values differ from the padded layout's\n"); return 16; }
            if (!(err16 == err16)) { printf("LAY16 epilogue read an unpublished slot\n"); return 17; }
            printf("rfft8192 without the per-16 padding (LAY16): max rel err %.3e, mirror values identical\n", err16);
            worst = fmax(worst, err16);
        }
        // ---- VARIANT_TWPROD: pass-1 twiddles from 4 loads + 11 products store (nearly) what 15 loads store ----
        {
            std::vector<cpx> bufA(r8k::BUF_CPX), bufB(r8k::BUF_CPX);
            for (int bb = 0; bb < 256; bb++) {
                cpx v[16], u[16];
                for (int q = 0; q < 16; q++) {
                    const int nn = bb + 256 * q;
                    v[q] = u[q] = cpx{(float)a[2 * nn], (float)a[2 * nn + 1]};
                }
                r8k::pass1_store(bb, v, tw.data(), bufA.data());
                r8k::pass1_store_prod(bb, u, tw.data(), bufB.data());
            }
            double dmax = 0, vmax = 0;
            for (int i = 0; i < 4096; i++) {
                const cpx p = bufA[r8k::pad(i)], q = bufB[r8k::pad(i)];
                dmax = fmax(dmax, fmax(fabs(p.x - q.x), fabs(p.y - q.y)));
                vmax = fmax(vmax, fmax(fabs(p.x), fabs(p.y)));
            }
            printf("pass-1 product twiddles: max diff %.3e of %.3e\n", dmax, vmax);
            if (!(dmax <= 1e-6 * vmax)) { printf("product twiddles disagree\n"); return 9; }
            bufB = bufA;  // same input to both forms of pass 2
            for (int bb = 0; bb < 256; bb++) {
                r8k::pass2(bb, tw2.data(), bufA.data());
                r8k::pass2_prod(bb, tw2.data(), bufB.data());
            }
            dmax = vmax = 0;
            for (int i = 0; i < 4096; i++) {
                const cpx p = bufA[r8k::pad(i)], q = bufB[r8k::pad(i)];
                dmax = fmax(dmax, fmax(fabs(p.x - q.x), fabs(p.y - q.y)));
                vmax = fmax(vmax, fmax(fabs(p.x), fabs(p.y)));
            }
            printf("pass-2 product twiddles: max diff %.3e of %.3e\n", dmax, vmax);
            if (!(dmax <= 1e-6 * vmax)) { printf("pass-2 product twiddles disagree\n"); return 12; }
        }
        // ---- VARIANT_WINSYN: Hann pairs from the thread's phase against the reference's f32 window ----
        {
            const float PI_F = 3.14159274101257324f;
            double wmax = 0;
            cpx w[256][16];
            for (int t = 0; t < 256; t++) {
                const double te = 2.0 * M_PI * (double)(2 * t) / 8192.0, to = 2.0 * M_PI * (double)(2 * t + 1) / 8192.0;
                const cpx cw{(float)cos(te), (float)cos(to)}, sw{(float)sin(te), (float)sin(to)};
#define BLISS_W(Q) w[t][Q] = r8k::hann_pair<Q>(cw, sw);
                BLISS_W(0) BLISS_W(1) BLISS_W(2) BLISS_W(3) BLISS_W(4) BLISS_W(5) BLISS_W(6) BLISS_W(7)
                BLISS_W(8) BLISS_W(9) BLISS_W(10) BLISS_W(11) BLISS_W(12) BLISS_W(13) BLISS_W(14) BLISS_W(15)
#undef BLISS_W
                for (int q = 0; q < 16; q++) {
                    const int n0 = 2 * (t + 256 * q);
                    const float r0 = 0.5f - 0.5f * cosf(2.f * (float)n0 * PI_F / 8192.f);        // utils.rs:36-38
                    const float r1 = 0.5f - 0.5f * cosf(2.f * (float)(n0 + 1) * PI_F / 8192.f);
                    wmax = fmax(wmax, fmax(fabs(w[t][q].x - r0), fabs(w[t][q].y - r1)));
                }
            }
            printf("synthesised Hann window: max abs diff to the f32 table %.3e\n", wmax);
            if (!(wmax <= 4e-7)) { printf("synthesised window disagrees\n"); return 10; }
        }
        std::vector<int> seen(r8k::BUF_CPX, 0);
        for (int i = 0; i < 4096; i++) {
            int p = r8k::pad(i);
            if (p >= r8k::BUF_CPX || seen[p]++) { printf("pad collision at %d\n", i); return 2; }
        }
    }
    if (worst > 5e-6) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}

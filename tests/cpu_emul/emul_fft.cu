// Host-only emulation of the warp / CTA FFT index logic of the CUDA kernels.
// Built with nvcc but launches nothing: every "lane" / "thread" is a loop iteration.
// Prints max relative error vs an f64 DFT for (a) the 512-point two-frame warp FFT
// incl. untangling, (b) the 8192-point three-pass FFT incl. untangling.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../bliss-rs_b200/csrc/fft8192.cuh"
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
    }
    // ---------------- 8192-point, two frames per CTA ----------------
    {
        std::vector<cpx> tw(8192);
        for (int m = 0; m < 8192; m++) {
            double a = -2.0 * M_PI * (double)m / 8192.0;
            tw[m] = cpx{(float)cos(a), (float)sin(a)};
        }
        std::vector<double> a(8192), b(8192);
        for (int i = 0; i < 8192; i++) {
            a[i] = (rand() / (double)RAND_MAX - 0.5);
            b[i] = (rand() / (double)RAND_MAX - 0.5) * 2.0;
        }
        std::vector<cpx> buf(f8k::BUF_CPX);
        for (int bb = 0; bb < 512; bb++) {
            cpx v[16];
            for (int q = 0; q < 16; q++) v[q] = cpx{(float)a[bb + 512 * q], (float)b[bb + 512 * q]};
            f8k::pass1_store(bb, v, tw.data(), buf.data());
        }
        for (int bb = 0; bb < 512; bb++) f8k::pass2(bb, tw.data(), buf.data());
        for (int bb = 0; bb < 512; bb++) f8k::pass3(bb, buf.data());
        // spot-check 200 bins (full f64 DFT of 8192 x 4097 is slow-ish but fine: do all)
        std::vector<double> ma, mb;
        dft_real(a, ma, 8192);
        dft_real(b, mb, 8192);
        double scale = 0;
        for (int k = 0; k <= 4096; k++) scale = fmax(scale, fmax(ma[k], mb[k]));
        double err = 0;
        for (int k = 0; k <= 4096; k++) {
            float fa, fb;
            f8k::untangle_mag(f8k::bin_value(buf.data(), k), f8k::bin_value(buf.data(), (8192 - k) & 8191), fa, fb);
            err = fmax(err, fabs(fa - ma[k]) / scale);
            err = fmax(err, fabs(fb - mb[k]) / scale);
        }
        printf("fft8192 pair: max rel err %.3e\n", err);
        worst = fmax(worst, err);
        // hand-folded epilogue addressing (as the kernel does it) must agree with bin_value()
        for (int t = 0; t < 512; t++)
            for (int m = 0; m < 8; m++) {
                const int k = t + 512 * m;
                const cpx zk = f8k::pair_sum(buf.data() + f8k::ebase(t) + 4 * m);
                cpx zm;
                if (k == 0) zm = zk;
                else if (t == 0) zm = f8k::pair_diff(buf.data() + f8k::ebase(0) + 4 * (8 - m));
                else zm = f8k::pair_diff(buf.data() + f8k::ebase(512 - t) + 4 * (7 - m));
                const cpx rk = f8k::bin_value(buf.data(), k), rm = f8k::bin_value(buf.data(), (8192 - k) & 8191);
                if (zk.x != rk.x || zk.y != rk.y || zm.x != rm.x || zm.y != rm.y) { printf("epilogue addressing mismatch k=%d\n", k); return 3; }
            }
        // padding must be injective
        std::vector<int> seen(f8k::BUF_CPX, 0);
        for (int i = 0; i < 8192; i++) {
            int p = f8k::pad(i);
            if (p >= f8k::BUF_CPX || seen[p]++) { printf("pad collision at %d\n", i); return 2; }
        }
    }
    if (worst > 5e-6) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}

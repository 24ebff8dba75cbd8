// stft8192_v3.cuh -- stft8192v3_kernel: stft8192v2_kernel's data path (bulk-copy staging one frame ahead, rotated
// frames, synthesised window, register twiddles, one 33 KB FFT buffer laid out 258 A + 16 B + c) with ONE radix-16
// column per thread and 256 threads per frame.
//
// Why: the 128-thread kernel keeps two columns (64 data registers) per thread so that mirror pairs meet in registers;
// at 148 registers and 73 KB of shared memory that is 3 CTAs = 12 warps per SM, and its ncu capture shows what that
// costs -- issue slots 44 % busy, FP32 pipe ~55 %, a quarter of all stall samples at the three barriers
// (profiles/ncu_r02_stft8192v2_b_128songs.md): nothing is saturated, the SM is waiting.  With one column per thread
// the same three CTAs hold 24 warps.  The price is the mirror exchange the two-column form avoided: after pass 3 a
// thread writes the upper half of its column (8 values, into its OWN block of the buffer, so no barrier before the
// write) and reads the upper half of thread 256 - t's column: 4 + 4 conflict-free 128-bit accesses and one more
// barrier per frame; and the pass-2 twiddles come from four registers by products (+22 packed FP instructions per
// thread) because fifteen of them no longer fit.
#pragma once
#include "stft8192_v2.cuh"

namespace bliss {
namespace s3 {

constexpr int THREADS = 256;
constexpr int ITEMS_PER_CTA = 4;
constexpr size_t SMEM_BYTES = s2::SMEM_BYTES;
constexpr int PHASE_OFF = s2::PHASE_OFF + 4 * 128 * 8;  // float offset of the [2][256][4] phase table (api.cu build_tables)

template <int A>
BLISS_HD void window_rows1(cpx (&v)[16], cpx cw, cpx sw) {
    if constexpr (A < 8) {
        const cpx g = s2::hann_row<A>(cw, sw);
        v[A] = pmul(v[A], g);
        v[A + 8] = pmul(v[A + 8], psub(cpx{s2::KSCALE, s2::KSCALE}, g));
        window_rows1<A + 1>(v, cw, sw);
    }
}

// fifteen twiddles w^1 .. w^15 from w^1, w^2, w^4, w^8, applied to slot bitrev(K) and stored at o[STRIDE * K]
#define BLISS_S3_TW(K, W) o[STRIDE * (K)] = cmul(v[bitrev((K), 4)], (W))
template <int STRIDE>
BLISS_HD void twiddle_store(const cpx (&v)[16], cpx t1, cpx t2, cpx t4, cpx t8, cpx *o) {
    o[0] = v[0];
    BLISS_S3_TW(8, t8);
    BLISS_S3_TW(4, t4);
    BLISS_S3_TW(12, cmul(t4, t8));
    BLISS_S3_TW(2, t2);
    BLISS_S3_TW(10, cmul(t2, t8));
    const cpx t6 = cmul(t2, t4);
    BLISS_S3_TW(6, t6);
    BLISS_S3_TW(14, cmul(t6, t8));
    BLISS_S3_TW(1, t1);
    BLISS_S3_TW(9, cmul(t1, t8));
    const cpx t5 = cmul(t1, t4);
    BLISS_S3_TW(5, t5);
    BLISS_S3_TW(13, cmul(t5, t8));
    const cpx t3 = cmul(t1, t2);
    BLISS_S3_TW(3, t3);
    BLISS_S3_TW(11, cmul(t3, t8));
    const cpx t7 = cmul(t3, t4);
    BLISS_S3_TW(7, t7);
    BLISS_S3_TW(15, cmul(t7, t8));
}
#undef BLISS_S3_TW

// bins k = t + 256 C, C < 8, from the thread's own low half and the mirror thread's high half (zm[C] = Z[4096 - k])
template <int C>
BLISS_HD void epilogue(const cpx (&v)[16], const cpx (&zm)[8], cpx wt, float *glo, float *ghi, float *plo, float &mx) {
    if constexpr (C < 8) {
        float a, b;
        s2::untangle_pair(v[bitrev(C, 4)], zm[C], mul_tw<C, 32>(wt), a, b);
        glo[256 * C] = a;
        ghi[-256 * C] = b;
        if constexpr (C < 6) plo[256 * C] = a;
        mx = fmaxf(mx, fmaxf(a, b));
        epilogue<C + 1>(v, zm, wt, glo, ghi, plo, mx);
    }
}

}  // namespace s3

__global__ void __launch_bounds__(s3::THREADS, 3)
stft8192v3_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                  const unsigned int *__restrict__ frame_prefix, int n_songs, unsigned int total_items,
                  int frames_per_item, int items_per_cta, const float *__restrict__ hann /* + s3::PHASE_OFF: [2][256][4] */,
                  const cpx *__restrict__ tw1 /*[16][256] W4096^(b k1)*/, const cpx *__restrict__ tw2g /*[16][16] W256^(j k2)*/,
                  const cpx *__restrict__ tw8192, float *__restrict__ mags, double *__restrict__ cand_mag,
                  double *__restrict__ cand_pitch, unsigned int *__restrict__ cand_count) {
#ifdef BLISS_HOST_EMUL
    unsigned char *s3_smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(128) unsigned char s3_smem[];
#endif
    float *X = reinterpret_cast<float *>(s3_smem);
    cpx *Y = reinterpret_cast<cpx *>(s3_smem + s2::X_FLOATS * 4);
    float *P = reinterpret_cast<float *>(s3_smem + s2::X_FLOATS * 4 + s2::Y_CPX * 8);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(P + s2::PIP_FLOATS + 64);
    __shared__ s2::FrameDesc s_fd[2];
    __shared__ s2::Cursor cur;
    __shared__ float s_red[s3::THREADS / 32];

    const int u = threadIdx.x, lane = u & 31;
    // ---- per-thread constants of the whole CTA -------------------------------------------------------------
    const float2 *t1g = reinterpret_cast<const float2 *>(tw1) + u;  // W4096^(u k1) at [k1][u]
    const float2 f1 = __ldg(t1g + 256 * 1), f2 = __ldg(t1g + 256 * 2), f4 = __ldg(t1g + 256 * 4), f8 = __ldg(t1g + 256 * 8);
    const cpx t1 = cpx{f1.x, f1.y}, t2 = cpx{f2.x, f2.y}, t4 = cpx{f4.x, f4.y}, t8 = cpx{f8.x, f8.y};
    const float2 *t2g = reinterpret_cast<const float2 *>(tw2g) + (u & 15);  // W256^(c B) at [B][c]
    const float2 h1 = __ldg(t2g + 16 * 1), h2 = __ldg(t2g + 16 * 2), h4 = __ldg(t2g + 16 * 4), h8 = __ldg(t2g + 16 * 8);
    const cpx q1 = cpx{h1.x, h1.y}, q2 = cpx{h2.x, h2.y}, q4 = cpx{h4.x, h4.y}, q8 = cpx{h8.x, h8.y};
    const float2 wtg = __ldg(reinterpret_cast<const float2 *>(tw8192) + u);
    const cpx wt = cpx{wtg.x, wtg.y};  // W8192^u
    // pass 3: column t = u at 258 A + 16 B; its mirror column 256 - u (thread 0 and thread 128 mirror onto themselves)
    const int pA = u & 15, pB = u >> 4;
    const int base1 = 258 * pA + 16 * pB;
    const int um = (256 - u) & 255;
    const int base2 = 258 * (um & 15) + 16 * (um >> 4);

    if (u == 0) {
        s2::mbar_init(bar, 1);
        cur.item = blockIdx.x * (unsigned)items_per_cta;
        cur.item_end = min(cur.item + (unsigned)items_per_cta, total_items);
        cur.song_item0 = cur.song_item1 = 0u;
        s_fd[0].valid = 0;
        s_fd[1].valid = 0;
        if (s2::cursor_load_item(cur, songs, frame_prefix, n_songs, frames_per_item)) {
            s2::cursor_describe(cur, pcm, mags, s_fd[0]);
            if (!s_fd[0].edge)
                s2::bulk_load(X, s_fd[0].x + s_fd[0].s0 - s_fd[0].r, (8192u + (s_fd[0].r ? 4u : 0u)) * 4u, bar);
        }
    }
    __syncthreads();
    if (s_fd[0].valid && s_fd[0].edge) {  // reflect-padded frame: filled by hand (four frames of a song)
        const float *x = s_fd[0].x;
        const int n = s_fd[0].n, s0 = s_fd[0].s0;
        for (int m = u; m < 8192; m += s3::THREADS) X[m] = r8k::reflect_sample(x, n, (long long)s0 + m);
        __syncthreads();
        if (u == 0) s2::mbar_arrive(bar);
    }

    for (unsigned int it = 0;; it++) {
        const int ph = (int)(it & 1u);
        const s2::FrameDesc &fd = s_fd[ph];
        if (!fd.valid) break;
        // The copy started r = (start & 3) samples early.  64-bit loads need an even offset only: the frame is
        // transformed rotated by r & 1 samples, y'[m] = y[(m - (r & 1)) mod 8192] = X[m + (r & 2)] (m >= r & 1)
        const int rot = fd.r & 1, xoff = fd.r & 2;
        s2::mbar_wait(bar, (unsigned)ph);
        if (u == 0 && rot) X[xoff] = X[8191 + fd.r];  // y'[0] = y[8191]
        // ---- pass 1: column u (z[nn] = y'[2 nn] + i y'[2 nn + 1], nn = 256 a + u) -------------------------------------
        {
            cpx v[16];
            const float2 *xin = reinterpret_cast<const float2 *>(X + xoff) + u;
#pragma unroll
            for (int a = 0; a < 16; a++) {
                const float2 t = xin[256 * a];
                v[a] = cpx{t.x, t.y};
            }
            {   // Hann window of the rotated frame: phases of samples 2u - rot, 2u + 1 - rot
                const float4 pw = __ldg(reinterpret_cast<const float4 *>(hann + s3::PHASE_OFF) + rot * s3::THREADS + u);
                s3::window_rows1<0>(v, cpx{pw.x, pw.y}, cpx{pw.z, pw.w});
            }
            fft_dif<16>(v);
            s3::twiddle_store<258>(v, t1, t2, t4, t8, Y + u);  // Y[258 A + u]
        }
        __syncthreads();  // B1: Y complete, X consumed
        if (u == 0) {     // next frame: descriptor + bulk copy, a whole frame ahead
            s2::FrameDesc &nd = s_fd[ph ^ 1];
            nd.valid = 0;
            cur.f++;
            bool more = cur.f < cur.fend;
            if (!more) {
                cur.item++;
                more = s2::cursor_load_item(cur, songs, frame_prefix, n_songs, frames_per_item);
            }
            if (more) {
                s2::cursor_describe(cur, pcm, mags, nd);
                if (!nd.edge) s2::bulk_load(X, nd.x + nd.s0 - nd.r, (8192u + (nd.r ? 4u : 0u)) * 4u, bar);
            }
        }
        // ---- pass 2: column (A, c) = (u >> 4, u & 15): radix 16 over b at stride 16, twiddle W256^(c B), in place ----
        {
            cpx *p = Y + 258 * (u >> 4) + (u & 15);
            cpx v[16];
#pragma unroll
            for (int b = 0; b < 16; b++) v[b] = p[16 * b];
            fft_dif<16>(v);
            s3::twiddle_store<16>(v, q1, q2, q4, q8, p);
        }
        __syncthreads();  // B2
        {   // a reflect-padded next frame is filled by hand (X is idle: no copy was issued for it)
            const s2::FrameDesc &nd = s_fd[ph ^ 1];
            if (nd.valid && nd.edge) {
                const float *x = nd.x;
                const int n = nd.n, s0 = nd.s0;
                for (int m = u; m < 8192; m += s3::THREADS) X[m] = r8k::reflect_sample(x, n, (long long)s0 + m);
                if (u == 0) s2::mbar_arrive(bar);
            }
        }
        // ---- pass 3 on column t = u; the upper half goes back into the thread's own block for its mirror thread --------
        float mx = 0.f;
        {
            cpx v[16], zm[8];
            float4 *q1p = reinterpret_cast<float4 *>(Y + base1);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 a = q1p[c];
                v[2 * c] = cpx{a.x, a.y};
                v[2 * c + 1] = cpx{a.z, a.w};
            }
            fft_dif<16>(v);  // v[bitrev(C)] = Z[u + 256 C]
#pragma unroll
            for (int c = 0; c < 4; c++) {  // own block, slots 0..7 <- Z[u + 256 (8 + slot)]
                const cpx e = v[bitrev(8 + 2 * c, 4)], o = v[bitrev(9 + 2 * c, 4)];
                q1p[c] = make_float4(e.x, e.y, o.x, o.y);
            }
            if (u == 0) {
                // thread 0 mirrors onto itself one index further on (4096 - 256 C = 256 (16 - C)) and pairs Z[0] with
                // itself (X[0] / X[4096]): it publishes Z[256 * 9 .. 256 * 15], Z[0] so that the common read below fits
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const cpx e = v[bitrev(9 + 2 * c, 4)], o = (c < 3) ? v[bitrev(10 + 2 * c, 4)] : v[0];
                    q1p[c] = make_float4(e.x, e.y, o.x, o.y);
                }
            }
            __syncthreads();  // B3: every upper half is published (each thread wrote only the block it had read itself)
            {
                // Z[4096 - k], k = u + 256 C: column 256 - u, index 15 - C = slot 7 - C of the mirror thread's block
                const float4 *q2p = reinterpret_cast<const float4 *>(Y + base2);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float4 a = q2p[c];
                    zm[7 - 2 * c] = cpx{a.x, a.y};
                    zm[6 - 2 * c] = cpx{a.z, a.w};
                }
            }
            float *grow = fd.row;
            s3::epilogue<0>(v, zm, wt, grow + u, grow + 4096 - u, P + u, mx);
            if (u == 0) {  // the self-mirrored bin 2048: W8192^2048 = -i
                float mid, dummy;
                s2::untangle_pair(v[bitrev(8, 4)], v[bitrev(8, 4)], cpx{0.f, -1.f}, mid, dummy);
                grow[2048] = mid;
                mx = fmaxf(mx, mid);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[u >> 5] = mx;
        __syncthreads();  // B4: bins < 1536 and the warp maxima are in shared memory; every mirror read of Y is done
        // ---- pip_track on centre bins 57..1483 (beginning = 56, end = 1486 for n_fft = 8192): 6 centres per thread ----
        {
            float fmx = s_red[0];
#pragma unroll
            for (int w = 1; w < s3::THREADS / 32; w++) fmx = fmaxf(fmx, s_red[w]);
            const double ref = 0.1 * (double)fmx;
            float thr = (float)ref;  // (double)elem > ref  <=>  elem > thr, thr = the largest f32 that is <= ref
            if ((double)thr > ref) thr = __uint_as_float(__float_as_uint(thr) - 1u);
            unsigned int flags = 0;
            const int b0 = 56 + 6 * u;  // centres c = b0 + 1 + i, i < 6
            if (u < 238) {
                float m[8];
                const float2 *q = reinterpret_cast<const float2 *>(P + b0);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 t = q[i];
                    m[2 * i] = t.x;
                    m[2 * i + 1] = t.y;
                }
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const float before = m[i], elem = m[i + 1], after = m[i + 2];
                    if (after <= elem && before < elem && elem > thr && fmx > 0.f) flags |= 1u << i;
                }
                if (u == 237) flags &= 0x1fu;  // centre 1484 is past the last one (1483)
            }
            const int cnt = __popc(flags);
            unsigned int incl = (unsigned)cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned int wtot = __shfl_sync(0xffffffffu, incl, 31);
            unsigned int wbase = 0;
            if (lane == 0 && wtot) wbase = atomicAdd(cand_count + fd.si, wtot);  // one reservation per warp: order is free (chroma.rs:361-391)
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            unsigned long long dst = fd.cand_off + wbase + (incl - (unsigned)cnt);
            while (flags) {
                const int i = __ffs(flags) - 1;
                flags &= flags - 1;
                const int c = b0 + 1 + i;
                const double before = (double)P[c - 1], elem = (double)P[c], after = (double)P[c + 1];
                const double avg = 0.5 * (after - before);
                double shift = 2. * elem - after - before;
                if (fabs(shift) < 2.2250738585072014e-308) shift += 1.;
                shift = avg / shift;
                cand_pitch[dst] = ((double)c + shift) * (double)SAMPLE_RATE / 8192.0;
                cand_mag[dst] = elem + 0.5 * avg * shift;
                dst++;
            }
        }
        // no barrier here: the next writes to Y / P / s_red / s_fd[ph] sit behind the next frame's B1 .. B3
    }
}

}  // namespace bliss

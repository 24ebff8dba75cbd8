// stft8192_v2.cuh -- stft8192v2_kernel (round 2): the reflect-padded 8192-point Hann STFT of utils::stft
// (src/utils.rs:26-64) + pip_track (src/chroma.rs:269-331), re-cut around the resource the round-1 kernel was bound
// by -- the L1 / shared-memory data pipe (84 % busy: ~2 200 shared-memory + ~1 200 global wavefronts per frame).
//
//   * TMA staging.  One elected thread streams the frame's 8192 samples global -> shared with ONE 1-D bulk copy
//     (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP), issued a whole frame ahead: no sample or window load
//     ever touches the LSU, no register is held across the load latency.  (The first cut also sent the finished
//     magnitude row out as a shared -> global bulk store; measured on B200 the CTA then sat at a barrier waiting
//     for the store's reads of the row -- 12 % of all stall samples, profiles/ncu_r02_stft8192v2_first_128songs.md
//     -- so the magnitudes now leave the registers as 128-byte warp stores and only bins < 1536 are staged for pip_track.)
//   * Alignment by rotation.  A frame starts at any sample (hop 2205, arbitrary song offsets); the bulk copy
//     starts at the 16-byte boundary below it and the frame is transformed ROTATED by r = start & 3 samples,
//     y'[m] = y[(m - r) mod 8192]: |DFT| is unchanged, every load is an aligned 128-bit one, and the only samples
//     that wrap are the first r (patched by one thread).  The window is synthesised with the same rotation.
//   * 128 threads per frame, TWO radix-16 columns per thread and pass, 4096 = 16 x 16 x 16 in place in one
//     33 KB buffer laid out P(A, B, c) = 258 A + 16 B + c: pass 1 stores and pass 3 loads are 128-bit, every
//     access of every pass is bank-conflict free (the arithmetic is next to each pass).
//   * Mirror pairs in registers.  In pass 3 thread u transforms column t = u AND its mirror column 256 - u, so
//     Z[k] and Z[4096 - k] of all its bins meet in ITS registers: the real-input untangling needs no exchange
//     through shared memory at all (round 1: 8 stores + 8 conflicting loads per thread).
//   * Twiddles: pass 1 from four per-thread registers (W^m, W^2m, W^4m, W^8m) by products, the second column by
//     the constant rotation W4096^A; pass 2 all fifteen W256^(cB) live in registers for the whole CTA; the
//     untangling uses W8192^t W32^C with W32^C a constant rotation.  No twiddle is loaded inside the frame loop.
//   * Hann window from the thread's phase (2 FFMA2 per sample pair, the upper half of the column by
//     w(a + 8) = 1 - w(a)), carried 2^30 high so that every magnitude is ONE MUFU.SQRT (ftz) and one multiply.
// Shared-memory traffic per frame: 256 (sample loads) + 1024 (two exchanges) + 128 (magnitude row) + ~20
// (pip_track) wavefronts, against ~3 400 in round 1.  The kernel is then bound by the FP32 pipe (~120 k packed
// FP instructions per frame; FFMA2 issues at one per two cycles and SMSP -- profiles/ab_r02.md).
#pragma once
#include "common.cuh"
#include "fft_regs.cuh"
#include "rfft8192.cuh"

namespace bliss {
namespace s2 {

constexpr int THREADS = 128;
constexpr int ITEMS_PER_CTA = 4;           // work items (of K3_FRAMES_PER_CTA = 4 frames each) one CTA walks
constexpr int Y_CPX = 258 * 16;            // 4128 complex = 33 024 B
constexpr int X_FLOATS = 8256;             // 8192 + 4 staged samples, rounded to the size of Y
constexpr int PIP_FLOATS = 1536;           // magnitudes of bins 0..1535 staged for pip_track (centre bins 57..1483)
constexpr float KSCALE = 1073741824.f;     // 2^30 carried by the window
constexpr float KINV = 0.5f / 1073741824.f;  // takes it out again, with the 1/2 of the real-input untangling
constexpr size_t SMEM_BYTES = (size_t)X_FLOATS * 4 + (size_t)Y_CPX * 8 + (size_t)PIP_FLOATS * 4 + 512;
constexpr int PHASE_OFF = 2 * (8192 + 4 * 256);  // float offset of the [4][128][8] rotated-phase table behind the Hann tables

// exp(-2 pi i A / 4096), A = 1..15 (f64-rounded-to-f32): the rotation between a thread's two pass-1 columns
__host__ __device__ constexpr double tsin_(double x) {  // |x| < 0.03: the series converges in a few terms
    double t = x, s = x;
    for (int i = 1; i < 8; i++) {
        t *= -x * x / ((2 * i) * (2 * i + 1));
        s += t;
    }
    return s;
}
__host__ __device__ constexpr double tcos_(double x) {
    double t = 1.0, s = 1.0;
    for (int i = 1; i < 8; i++) {
        t *= -x * x / ((2 * i - 1) * (2 * i));
        s += t;
    }
    return s;
}
template <int A>
BLISS_HD cpx w4096() {
    constexpr double ang = 6.283185307179586476925286766559 * (double)A / 4096.0;
    constexpr float c = (float)tcos_(ang), s = (float)-tsin_(ang);
    return cpx{c, s};
}

// ---- TMA / mbarrier primitives (sm_90+ PTX; the host emulation copies on the spot) -------------------------
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#else
inline void mbar_init(unsigned long long *, int) {}
inline void mbar_arrive(unsigned long long *) {}
inline void mbar_wait(unsigned long long *, unsigned) {}
inline void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
#endif

struct FrameDesc {           // what the CTA needs to know about one frame (written by thread 0, one frame ahead)
    const float *x;          // the song
    float *row;              // its magnitude row in the spill
    unsigned long long cand_off;
    int n, s0, r, edge, si, valid;
};

// thread 0's cursor over the CTA's frames: items [item, item_end) of K3_FRAMES_PER_CTA consecutive frames each
struct Cursor {   // lives in shared memory: only thread 0 touches it
    unsigned int item, item_end;
    unsigned int song_item0, song_item1;  // the items [song_item0, song_item1) belong to song si
    int f, fend, si;
    SongDesc sd;
};

__device__ __forceinline__ bool cursor_load_item(Cursor &c, const SongDesc *songs, const unsigned int *frame_prefix,
                                                 int n_songs, int frames_per_item) {
    while (c.item < c.item_end) {
        if (c.item < c.song_item0 || c.item >= c.song_item1) {  // another song: the only global loads of the cursor
            c.si = find_song(frame_prefix, n_songs, c.item);
            c.sd = songs[c.si];
            c.song_item0 = frame_prefix[c.si];
            c.song_item1 = frame_prefix[c.si + 1];
        }
        c.f = (int)(c.item - c.song_item0) * frames_per_item;
        c.fend = min(c.f + frames_per_item, (int)c.sd.n_c_comp);
        if (c.f < c.fend) return true;
        c.item++;
    }
    return false;
}

__device__ __forceinline__ void cursor_describe(const Cursor &c, const float *pcm, float *mags, FrameDesc &d) {
    d.x = pcm + c.sd.pcm_off;
    d.row = mags + (size_t)((unsigned int)c.sd.mag_off + (unsigned int)c.f) * CH_STRIDE;
    d.cand_off = c.sd.cand_off;
    d.n = (int)c.sd.n;
    d.s0 = CH_HOP * c.f - 4096;  // the frame covers samples s0 .. s0 + 8191 of the reflect-padded song (utils.rs:11-24, :44-47)
    d.si = c.si;
    const bool interior = d.s0 >= 0 && d.s0 + 8191 < d.n;
    const int r = (int)((c.sd.pcm_off + (unsigned long long)(interior ? d.s0 : 0)) & 3ull);
    // the copy starts r samples early and, if r != 0, ends 4 - r samples late: both must stay inside the song
    // (the samples before it belong to the same buffer: pcm_off + s0 >= r by construction)
    const bool tma = interior && (r == 0 || d.s0 - r + 8196 <= d.n);
    d.r = tma ? r : 0;
    d.edge = tma ? 0 : 1;
    d.valid = 1;
}

// Hann pair of row a (a < 8) from the thread's phase: 2^30 (0.5 - 0.5 cos(phi + 2 pi a / 16))
template <int A>
BLISS_HD cpx hann_row(cpx cw, cpx sw) {
    constexpr float ca = -0.5f * KSCALE * cos32(2 * A), sa = 0.5f * KSCALE * sin32(2 * A);
    return pfma(cw, cpx{ca, ca}, pfma(sw, cpx{sa, sa}, cpx{0.5f * KSCALE, 0.5f * KSCALE}));
}
template <int A>
BLISS_HD void window_rows(cpx (&v)[16], cpx cw, cpx sw) {
    if constexpr (A < 8) {
        const cpx g = hann_row<A>(cw, sw);
        v[A] = pmul(v[A], g);
        v[A + 8] = pmul(v[A + 8], psub(cpx{KSCALE, KSCALE}, g));  // cos(phi + pi + x) = -cos(phi + x)
        window_rows<A + 1>(v, cw, sw);
    }
}

// pass-1 twiddles and stores: slot s of both columns holds index A = bitrev(s); column 2u by W4096^(2u A)
// (products of t1, t2, t4, t8), column 2u + 1 by that times W4096^A; both to Y[258 A + 2u] as one 128-bit store
template <int A>
BLISS_HD void p1_store(const cpx (&v1)[16], const cpx (&v2)[16], cpx w, float4 *y4) {
    const cpx a = cmul(v1[bitrev(A, 4)], w);
    const cpx b = cmul(v2[bitrev(A, 4)], cmul(w, w4096<A>()));
    y4[129 * A] = make_float4(a.x, a.y, b.x, b.y);  // 258 A complex = 129 A float4
}

// |X[k]|, |X[4096 - k]| from Z[k], Z[4096 - k], w = W8192^k on data carried 2^30 high (rfft8192.cuh untangle_mag_pair)
BLISS_HD float mag_of(cpx x) {
    const cpx q = pmul(x, x);
#ifdef __CUDA_ARCH__
    return KINV * approx_sqrtf_ftz(__fadd_rn(q.x, q.y));
#else
    return KINV * sqrtf(q.x + q.y);
#endif
}
BLISS_HD void untangle_pair(cpx zk, cpx zm, cpx w, float &mag_k, float &mag_m) {
    const cpx e = padd(zk, cpx{zm.x, -zm.y});
    const cpx o = padd(zk, cpx{-zm.x, zm.y});
    const cpx p = cmul(o, w);
    mag_k = mag_of(padd(e, cpx{p.y, -p.x}));
    mag_m = mag_of(padd(e, cpx{-p.y, p.x}));
}

// bins k = u + 256 C (from the thread's column) and 4096 - k (from its mirror column): straight to the spill row
// (consecutive threads -> consecutive floats: 128-byte warp stores); bins < 1536 also to the pip_track buffer
template <int C>
BLISS_HD void epilogue_generic(const cpx (&v1)[16], const cpx (&v2)[16], cpx wt, float *glo, float *ghi, float *plo,
                               float *phi, float &mx) {
    if constexpr (C < 16) {
        float a, b;
        untangle_pair(v1[bitrev(C, 4)], v2[bitrev(15 - C, 4)], mul_tw<C, 32>(wt), a, b);
        glo[256 * C] = a;    // k = u + 256 C
        ghi[-256 * C] = b;   // 4096 - k
        if constexpr (C < 6) plo[256 * C] = a;
        if constexpr (C >= 10) phi[-256 * C] = b;
        mx = fmaxf(mx, fmaxf(a, b));
        epilogue_generic<C + 1>(v1, v2, wt, glo, ghi, plo, phi, mx);
    }
}

}  // namespace s2

__global__ void __launch_bounds__(s2::THREADS, 3)
stft8192v2_kernel(const float *__restrict__ pcm, const SongDesc *__restrict__ songs,
                  const unsigned int *__restrict__ frame_prefix, int n_songs, unsigned int total_items,
                  int frames_per_item, int items_per_cta, const float *__restrict__ hann /* + s2::PHASE_OFF: [4][128][8] */,
                  const cpx *__restrict__ tw1 /*[16][256] W4096^(b k1)*/, const cpx *__restrict__ tw2g /*[16][16] W256^(j k2)*/,
                  const cpx *__restrict__ tw8192, float *__restrict__ mags, double *__restrict__ cand_mag,
                  double *__restrict__ cand_pitch, unsigned int *__restrict__ cand_count) {
#ifdef BLISS_HOST_EMUL
    unsigned char *s2_smem = emu::dynamic_smem();
#else
    extern __shared__ __align__(128) unsigned char s2_smem[];
#endif
    float *X = reinterpret_cast<float *>(s2_smem);                    // staged samples (bulk-copy destination)
    cpx *Y = reinterpret_cast<cpx *>(s2_smem + s2::X_FLOATS * 4);    // FFT buffer (three passes in place)
    float *P = reinterpret_cast<float *>(s2_smem + s2::X_FLOATS * 4 + s2::Y_CPX * 8);  // magnitudes of bins < 1536 for pip_track
    cpx *S0 = reinterpret_cast<cpx *>(P + s2::PIP_FLOATS);                             // thread 0's two self-mirrored columns (32 cpx)
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(S0 + 32);
    __shared__ s2::FrameDesc s_fd[2];
    __shared__ s2::Cursor cur;  // thread 0's cursor over the CTA's frames
    __shared__ float s_red[s2::THREADS / 32];

    const int u = threadIdx.x, lane = u & 31;
    // ---- per-thread constants of the whole CTA -------------------------------------------------------------
    const float2 *t1g = reinterpret_cast<const float2 *>(tw1) + 2 * u;  // W4096^(2u k1) at [k1][2u]
    const float2 f1 = __ldg(t1g + 256 * 1), f2 = __ldg(t1g + 256 * 2), f4 = __ldg(t1g + 256 * 4), f8 = __ldg(t1g + 256 * 8);
    const cpx t1 = cpx{f1.x, f1.y}, t2 = cpx{f2.x, f2.y}, t4 = cpx{f4.x, f4.y}, t8 = cpx{f8.x, f8.y};
    cpx w2[16];  // W256^(c B), c = u & 15 (pass 2)
#pragma unroll
    for (int B = 1; B < 16; B++) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(tw2g) + 16 * B + (u & 15));
        w2[B] = cpx{t.x, t.y};
    }
    w2[0] = cpx{1.f, 0.f};
    const float2 wtg = __ldg(reinterpret_cast<const float2 *>(tw8192) + (u == 0 ? 128 : u));
    const cpx wt = cpx{wtg.x, wtg.y};  // W8192^u (thread 0: W8192^128, its second column)
    // pass 3: column t = u at 258 A + 16 B and its mirror column 256 - u (thread 0: columns 0 and 128)
    const int pA = u & 15, pB = u >> 4;
    const int base1 = 258 * pA + 16 * pB;
    const int base2 = (u == 0) ? 128 : 258 * ((16 - pA) & 15) + 16 * (pA ? 15 - pB : 16 - pB);

    // ---- thread 0: first frame of this CTA, its copy ---------------------------------------------------------
    // The CTA's bookkeeping thread (frame cursor, bulk copies).  (Rotating the role over the warps by CTA -- warp w of
    // every CTA lives on sub-partition w -- was measured: no difference, profiles/knobs_r02.md.)
    const bool leader = u == 0;
    if (leader) {
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
        for (int m = u; m < 8192; m += s2::THREADS) X[m] = r8k::reflect_sample(x, n, (long long)s0 + m);
        __syncthreads();
        if (leader) s2::mbar_arrive(bar);
    }

    for (unsigned int it = 0;; it++) {
        const int ph = (int)(it & 1u);
        const s2::FrameDesc &fd = s_fd[ph];
        if (!fd.valid) break;
        const int rot = fd.r;
        s2::mbar_wait(bar, (unsigned)ph);
        if (u == 0) {  // the rotated frame's first `rot` samples are the last ones of the frame (X[8192 + m])
            for (int m = 0; m < rot; m++) X[m] = X[8192 + m];
        }
        // ---- pass 1: columns 2u, 2u + 1 (z[nn] = y'[2 nn] + i y'[2 nn + 1], nn = 256 a + column) -----------------
        {
            cpx v1[16], v2[16];
            const float4 *xin = reinterpret_cast<const float4 *>(X) + u;
#pragma unroll
            for (int a = 0; a < 16; a++) {
                const float4 t = xin[128 * a];
                v1[a] = cpx{t.x, t.y};
                v2[a] = cpx{t.z, t.w};
            }
            {   // Hann window of the rotated frame: phases of samples 4u - rot + j, j = 0..3
                const float4 *pp = reinterpret_cast<const float4 *>(hann + s2::PHASE_OFF) + (rot * s2::THREADS + u) * 2;
                const float4 pc = __ldg(pp), ps = __ldg(pp + 1);
                s2::window_rows<0>(v1, cpx{pc.x, pc.y}, cpx{ps.x, ps.y});
                s2::window_rows<0>(v2, cpx{pc.z, pc.w}, cpx{ps.z, ps.w});
            }
            fft_dif<16>(v1);
            fft_dif<16>(v2);
            float4 *y4 = reinterpret_cast<float4 *>(Y) + u;  // Y[2u]
            y4[0] = make_float4(v1[0].x, v1[0].y, v2[0].x, v2[0].y);
            s2::p1_store<8>(v1, v2, t8, y4);
            s2::p1_store<4>(v1, v2, t4, y4);
            s2::p1_store<12>(v1, v2, cmul(t4, t8), y4);
            s2::p1_store<2>(v1, v2, t2, y4);
            s2::p1_store<10>(v1, v2, cmul(t2, t8), y4);
            const cpx t6 = cmul(t2, t4);
            s2::p1_store<6>(v1, v2, t6, y4);
            s2::p1_store<14>(v1, v2, cmul(t6, t8), y4);
            s2::p1_store<1>(v1, v2, t1, y4);
            s2::p1_store<9>(v1, v2, cmul(t1, t8), y4);
            const cpx t5 = cmul(t1, t4);
            s2::p1_store<5>(v1, v2, t5, y4);
            s2::p1_store<13>(v1, v2, cmul(t5, t8), y4);
            const cpx t3 = cmul(t1, t2);
            s2::p1_store<3>(v1, v2, t3, y4);
            s2::p1_store<11>(v1, v2, cmul(t3, t8), y4);
            const cpx t7 = cmul(t3, t4);
            s2::p1_store<7>(v1, v2, t7, y4);
            s2::p1_store<15>(v1, v2, cmul(t7, t8), y4);
        }
        __syncthreads();  // B1: Y complete, X consumed
        if (leader) {     // next frame: descriptor + bulk copy, a whole frame ahead
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
        // ---- pass 2: columns (A, c), (A + 8, c): radix 16 over b at stride 16, twiddle W256^(c B), in place ------
        {
            cpx *p = Y + 258 * (u >> 4) + (u & 15);
            cpx va[16], vb[16];
#pragma unroll
            for (int b = 0; b < 16; b++) {
                va[b] = p[16 * b];
                vb[b] = p[16 * b + 258 * 8];
            }
            fft_dif<16>(va);
            fft_dif<16>(vb);
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const int B = bitrev(s, 4);
                p[16 * B] = B ? cmul(va[s], w2[B]) : va[s];
                p[16 * B + 258 * 8] = B ? cmul(vb[s], w2[B]) : vb[s];
            }
        }
        __syncthreads();  // B2
        {   // a reflect-padded next frame is filled by hand (X is idle: no copy was issued for it)
            const s2::FrameDesc &nd = s_fd[ph ^ 1];
            if (nd.valid && nd.edge) {
                const float *x = nd.x;
                const int n = nd.n, s0 = nd.s0;
                for (int m = u; m < 8192; m += s2::THREADS) X[m] = r8k::reflect_sample(x, n, (long long)s0 + m);
                if (leader) s2::mbar_arrive(bar);  // (the barriers below order the fill before the next frame's loads)
            }
        }
        // ---- pass 3 on the thread's column and its mirror column; untangling in registers --------------------------
        float mx = 0.f;
        {
            cpx v1[16], v2[16];
            const float4 *q1 = reinterpret_cast<const float4 *>(Y + base1), *q2 = reinterpret_cast<const float4 *>(Y + base2);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 a = q1[c], b = q2[c];
                v1[2 * c] = cpx{a.x, a.y};
                v1[2 * c + 1] = cpx{a.z, a.w};
                v2[2 * c] = cpx{b.x, b.y};
                v2[2 * c + 1] = cpx{b.z, b.w};
            }
            fft_dif<16>(v1);
            fft_dif<16>(v2);
            float *grow = fd.row;
            if (u != 0) {
                s2::epilogue_generic<0>(v1, v2, wt, grow + u, grow + 4096 - u, P + u, P + 4096 - u, mx);
            } else {
                // thread 0's columns 0 and 128 mirror onto themselves: 17 pairs, handed to lanes 0..16 of this warp
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    S0[c] = v1[bitrev(c, 4)];        // Z[256 c]
                    S0[16 + c] = v2[bitrev(c, 4)];   // Z[128 + 256 c]
                }
            }
        }
        if (u < 32) {  // warp 0
            __syncwarp();
            if (lane < 17) {
                // lane L <= 8: k = 256 L <-> 4096 - k (L = 0: X[0] and X[4096], L = 8: the self-mirrored bin 2048);
                // L >= 9:  k = 128 + 256 (L - 9) <-> 4096 - k.  W8192^k = W64^j, j = k / 128 = tw1[k1 = j][b = 64]
                const bool c0 = lane <= 8;
                const int C = c0 ? lane : lane - 9;
                const cpx zk = c0 ? S0[C] : S0[16 + C];
                const cpx zm = c0 ? S0[(16 - C) & 15] : S0[16 + 15 - C];
                const int jw = c0 ? 2 * C : 2 * C + 1;
                cpx w = cpx{0.f, -1.f};  // W64^16
                if (jw < 16) {
                    const float2 t = __ldg(reinterpret_cast<const float2 *>(tw1) + 256 * jw + 64);
                    w = cpx{t.x, t.y};
                }
                float a, b;
                s2::untangle_pair(zk, zm, w, a, b);
                const int k = c0 ? 256 * C : 128 + 256 * C;
                float *grow = fd.row;
                grow[k] = a;
                grow[4096 - k] = b;
                if (k < s2::PIP_FLOATS) P[k] = a;
                if (4096 - k < s2::PIP_FLOATS) P[4096 - k] = b;
                mx = fmaxf(mx, fmaxf(a, b));
            }
            __syncwarp();
        }
        // frame maximum (pip_track's ref_value = 0.1 * max over all bins, chroma.rs:289-293)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[u >> 5] = mx;
        __syncthreads();  // B3: bins < 1536 and the warp maxima are in shared memory; every pass-3 load of Y is done
        // ---- pip_track on centre bins 57..1483 (beginning = 56, end = 1486 for n_fft = 8192): 12 centres per thread -
        {
            const float fmx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
            const double ref = 0.1 * (double)fmx;
            // (double)elem > ref for an f32 elem  <=>  elem > thr, thr = the largest f32 that is <= ref
            float thr = (float)ref;
            if ((double)thr > ref) thr = __uint_as_float(__float_as_uint(thr) - 1u);  // ref >= 0: one ulp down (0 stays 0: fmx = 0 has no peaks)
            unsigned int flags = 0;
            const int b0 = 56 + 12 * u;  // this thread's centres c = b0 + 1 + i, i < 12
            // (testing `elem > thr` first and the neighbours only for the survivors was measured: the divergent loop costs
            // more than the 24 compares it saves, 22.5 against 21.5 ms, profiles/knobs_r02.md)
            if (u < 119) {
                float m[16];
                const float4 *q = reinterpret_cast<const float4 *>(P + b0);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 t = q[i];
                    m[4 * i] = t.x; m[4 * i + 1] = t.y; m[4 * i + 2] = t.z; m[4 * i + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const float before = m[i], elem = m[i + 1], after = m[i + 2];
                    if (after <= elem && before < elem && elem > thr && (fmx > 0.f)) flags |= 1u << i;
                }
                if (u == 118) flags &= 0x7ffu;  // centre 1484 is past the last one (1483)
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
                // the residue bin of pitch_tuning (log2 / fmod in f64) is left to tuning_kernel; here only the
                // interpolation of chroma.rs:317-326
                cand_pitch[dst] = ((double)c + shift) * (double)SAMPLE_RATE / 8192.0;
                cand_mag[dst] = elem + 0.5 * avg * shift;
                dst++;
            }
        }
        // no barrier here: the next writes to Y / P / s_red / s_fd[ph] sit behind the next frame's B1 or B2, which no
        // warp passes before every warp has left this frame
    }
}

}  // namespace bliss

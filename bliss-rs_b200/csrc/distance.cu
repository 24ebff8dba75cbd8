// distance.cu -- K10: feature-vector distances of src/playlist.rs on the device.
//   mahalanobis_distance  sqrt((a-b)^T M (a-b))   playlist.rs:140-142  (default metric lib.rs:168-178)
//   euclidean_distance    M = I                   playlist.rs:65-71
//   cosine_distance       1 - a.b/(|a||b|)        playlist.rs:76-79
//   closest_to_songs keys: sum over seeds         playlist.rs:56-58, 256-270
// f32 throughout, accumulated in the order ndarray uses (unrolled_dot: 8 partial
// sums) so that the reference's exact-equality tests (lib.rs:273-291,
// playlist.rs:1009-1108) hold bit for bit.  The direct (a-b)^2 form is kept on
// purpose: the GEMM form |a|^2+|b|^2-2ab cancels for the near-duplicates that
// dedup_playlist thresholds at 0.05 (playlist.rs:381-382).
#ifndef BLISS_HOST_EMUL  // tests/cpu_emul/emul_distance.cpp runs the kernels below on the host (no CUB, no <<< >>>)
#include <cub/device/device_radix_sort.cuh>
#endif

#include "common.cuh"

namespace bliss {

constexpr int MAX_DIM = 64;

// ndarray numeric_util::unrolled_dot order
template <typename FA, typename FB>
__device__ __forceinline__ float unrolled_dot(FA xa, FB xb, int len) {
    float p[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 8 <= len; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) p[k] = __fadd_rn(p[k], __fmul_rn(xa(i + k), xb(i + k)));
    }
    float sum = 0.f;
    sum = __fadd_rn(sum, __fadd_rn(p[0], p[4]));
    sum = __fadd_rn(sum, __fadd_rn(p[1], p[5]));
    sum = __fadd_rn(sum, __fadd_rn(p[2], p[6]));
    sum = __fadd_rn(sum, __fadd_rn(p[3], p[7]));
    for (; i < len; i++) sum = __fadd_rn(sum, __fmul_rn(xa(i), xb(i)));
    return sum;
}

// mode: 0 = diagonal weights w[dim] (identity when w == nullptr), 1 = full matrix m[dim*dim], 2 = cosine
// DIM > 0: compile-time dimension (registers); DIM == 0: runtime dim <= MAX_DIM.
template <int DIM>
__device__ __forceinline__ float pair_distance(const float *a, const float *b, int dim_rt, int mode,
                                               const float *w_or_m) {
    const int dim = DIM > 0 ? DIM : dim_rt;
    if (mode == 2) {
        const float ab = unrolled_dot([&](int i) { return a[i]; }, [&](int i) { return b[i]; }, dim);
        const float aa = unrolled_dot([&](int i) { return a[i]; }, [&](int i) { return a[i]; }, dim);
        const float bb = unrolled_dot([&](int i) { return b[i]; }, [&](int i) { return b[i]; }, dim);
        return __fsub_rn(1.f, __fdiv_rn(ab, __fmul_rn(__fsqrt_rn(aa), __fsqrt_rn(bb))));
    }
    float d[DIM > 0 ? DIM : MAX_DIM], t[DIM > 0 ? DIM : MAX_DIM];
#pragma unroll
    for (int i = 0; i < dim; i++) d[i] = __fsub_rn(a[i], b[i]);
    if (mode == 0) {
#pragma unroll
        for (int i = 0; i < dim; i++) t[i] = w_or_m ? __fmul_rn(d[i], w_or_m[i]) : d[i];
    } else {
        for (int j = 0; j < dim; j++) {
            float s = 0.f;
            for (int i = 0; i < dim; i++) s = __fadd_rn(s, __fmul_rn(d[i], w_or_m[i * dim + j]));
            t[j] = s;
        }
    }
    const float s = unrolled_dot([&](int i) { return t[i]; }, [&](int i) { return d[i]; }, dim);
    return __fsqrt_rn(s);
}

// All-pairs block: out[i][j] = dist(rows[i], cols[j]).  CTA (128 threads) tile = 32 rows x 256 columns; thread t keeps
// its two column vectors (t and t + 128) in registers and walks the 32 rows of the tile, whose
// vectors are broadcast from shared memory; a warp's 32 results of one row are 32 consecutive floats.
template <int DIM>
__global__ void __launch_bounds__(128)
distance_matrix_kernel(const float *__restrict__ rows, unsigned int n_rows,
                       const float *__restrict__ cols, unsigned int n_cols, int mode,
                       const float *__restrict__ w_or_m, float *__restrict__ out) {
    constexpr int TR = 32, TC = 256, NT = 128;  // 2 columns per thread
    __shared__ float s_a[TR][DIM];
    __shared__ float s_w[(DIM <= 32) ? DIM * DIM : 1];
    const unsigned int r0 = blockIdx.y * TR, c0 = blockIdx.x * TC;
    for (int e = threadIdx.x; e < TR * DIM; e += NT) {
        const unsigned int r = r0 + e / DIM;
        s_a[e / DIM][e % DIM] = r < n_rows ? rows[(size_t)r * DIM + e % DIM] : 0.f;
    }
    const float *wm = nullptr;
    if (w_or_m) {
        const int nw = (mode == 1) ? DIM * DIM : DIM;
        if (DIM <= 32) {
            for (int e = threadIdx.x; e < nw; e += NT) s_w[e] = w_or_m[e];
            wm = s_w;
        } else {
            wm = w_or_m;
        }
    }
    // this thread's two columns, kept in registers for the whole tile
    const unsigned int ca = c0 + threadIdx.x, cb = c0 + (unsigned)NT + threadIdx.x;
    float va[DIM], vb[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) {
        va[i] = ca < n_cols ? __ldg(cols + (size_t)ca * DIM + i) : 0.f;
        vb[i] = cb < n_cols ? __ldg(cols + (size_t)cb * DIM + i) : 0.f;
    }
    __syncthreads();
    const unsigned int nr = min((unsigned)TR, n_rows - r0);
#pragma unroll 1
    for (unsigned int lr = 0; lr < nr; lr++) {
        float *o = out + (size_t)(r0 + lr) * n_cols;
        const float da = pair_distance<DIM>(s_a[lr], va, DIM, mode, wm);
        const float db = pair_distance<DIM>(s_a[lr], vb, DIM, mode, wm);
        if (ca < n_cols) o[ca] = da;
        if (cb < n_cols) o[cb] = db;
    }
}

// ---- diagonal-metric fast path (round 2) -----------------------------------------------------------------------
// The metrics the crate ships are diagonal (src/lib.rs:168-178, 209-234): sqrt(sum_i w_i d_i^2), evaluated by ndarray
// as unrolled_dot(d * w, d).  Same arithmetic as pair_distance (every product and every sum rounded on its own, in
// ndarray's order), two columns per thread carried as ONE packed f32x2 lane pair:
//   d = a - b                FADD2 (the row value is a scalar broadcast operand)
//   t = d * w_i              FMUL2, skipped where w_i == 1 (x * 1 is exact: 9 of the 23 v2 weights, all of v1's)
//   q = t * d                FMUL2
//   p_k = p_k * 1 + q        FFMA2 with the constant 1: one rounding of the exact p_k + q, i.e. the reference's
//                            separate add -- written as an FMA because ptxas contracts mul.f32x2 + add.f32x2 into one
//                            FFMA2 (a fused product, other bits), which it cannot do to an FMA
// The first round of the eight partial sums is p_k = q (0 + q is exact).  82 packed FP instructions per two pairs
// instead of ~190 scalar ones: the kernel moves from issue-bound to the FP32 pipe (83 roundings per pair cannot fuse).
#ifndef BLISS_HOST_EMUL  // (the host emulation keeps running the scalar kernel, whose bits this one must reproduce)
// ONES: compile-time mask of the features whose weight is exactly 1 (0 = none known: every product is made)
template <int DIM, unsigned int ONES>
__global__ void __launch_bounds__(128)
distance_matrix_diag_kernel(const float *__restrict__ rows, unsigned int n_rows, const float *__restrict__ cols,
                            unsigned int n_cols, const float *__restrict__ w, float *__restrict__ out, float one_rt) {
    constexpr int TR = 64, NT = 128, PITCH = (DIM + 3) / 4 * 4;
    // 1.0f as a RUN-TIME value: with the literal, ptxas simplifies fma(p, 1, q) to p + q and then contracts the
    // product behind q into it (FFMA2: a fused product, not the reference's bits) -- seen in the SASS
    const cpx one = cpx{one_rt, one_rt};
    __shared__ __align__(16) float s_a[TR][PITCH];
    const unsigned int r0 = blockIdx.y * TR, c0 = blockIdx.x * (2 * NT);
    for (int e = threadIdx.x; e < TR * PITCH; e += NT) {
        const unsigned int r = r0 + e / PITCH;
        const int i = e % PITCH;
        s_a[e / PITCH][i] = (r < n_rows && i < DIM) ? rows[(size_t)r * DIM + i] : 0.f;
    }
    const unsigned int ca = c0 + threadIdx.x, cb = c0 + (unsigned)NT + threadIdx.x;
    cpx v[DIM], wv[DIM];  // (column ca, column cb) per feature; weights as (w_i, w_i)
#pragma unroll
    for (int i = 0; i < DIM; i++) {
        v[i].x = ca < n_cols ? __ldg(cols + (size_t)ca * DIM + i) : 0.f;
        v[i].y = cb < n_cols ? __ldg(cols + (size_t)cb * DIM + i) : 0.f;
        const float wi = w ? __ldg(w + i) : 1.f;
        wv[i] = cpx{wi, wi};
    }
    __syncthreads();
    const unsigned int nr = min((unsigned)TR, n_rows - r0);
#pragma unroll 1
    for (unsigned int lr = 0; lr < nr; lr++) {
        float a[PITCH];
        const float4 *ra = reinterpret_cast<const float4 *>(s_a[lr]);
#pragma unroll
        for (int k = 0; k < PITCH / 4; k++) {
            const float4 t4 = ra[k];
            a[4 * k] = t4.x; a[4 * k + 1] = t4.y; a[4 * k + 2] = t4.z; a[4 * k + 3] = t4.w;
        }
        cpx p[8], sum = cpx{0.f, 0.f};
        constexpr int BODY = DIM / 8 * 8;
#pragma unroll
        for (int i = 0; i < DIM; i++) {
            const cpx d = psub(cpx{a[i], a[i]}, v[i]);
            const cpx t = ((ONES >> i) & 1u) ? d : pmul(d, wv[i]);
            const cpx q = pmul(t, d);
            if (i < 8) {
                p[i] = q;  // 0 + q
            } else if (i < BODY) {
                p[i & 7] = pfma(p[i & 7], one, q);
            } else {
                if (i == BODY) {  // ndarray: sum = ((p0 + p4) + (p1 + p5)) + (p2 + p6)) + (p3 + p7), then the tail in order
                    sum = padd(p[0], p[4]);
                    sum = padd(sum, padd(p[1], p[5]));
                    sum = padd(sum, padd(p[2], p[6]));
                    sum = padd(sum, padd(p[3], p[7]));
                }
                sum = pfma(sum, one, q);
            }
        }
        float *o = out + (size_t)(r0 + lr) * n_cols;
        if (ca < n_cols) __stcs(o + ca, __fsqrt_rn(sum.x));
        if (cb < n_cols) __stcs(o + cb, __fsqrt_rn(sum.y));
    }
}
#endif  // BLISS_HOST_EMUL

// any dim <= MAX_DIM: one thread per output (used for custom metrics on other vector sizes)
__global__ void __launch_bounds__(256)
distance_matrix_generic_kernel(const float *__restrict__ rows, unsigned int n_rows,
                               const float *__restrict__ cols, unsigned int n_cols, int dim, int mode,
                               const float *__restrict__ w_or_m, float *__restrict__ out) {
    const unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    for (unsigned int r = blockIdx.y; r < n_rows; r += gridDim.y)  // grid.y is capped at 65535: rows are strided over it
        out[(size_t)r * n_cols + c] = pair_distance<0>(rows + (size_t)r * dim, cols + (size_t)c * dim, dim, mode, w_or_m);
}

// keys[j] = sum_i dist(seed_i, cand_j), seeds visited in order (Iterator::sum, playlist.rs:56-58)
__global__ void __launch_bounds__(256)
seed_distance_kernel(const float *__restrict__ seeds, unsigned int n_seeds,
                     const float *__restrict__ cands, unsigned int n_cands, int dim, int mode,
                     const float *__restrict__ w_or_m, float *__restrict__ keys) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cands) return;
    float b[MAX_DIM];
    for (int i = 0; i < dim; i++) b[i] = cands[(size_t)j * dim + i];
    float s = 0.f;
    for (unsigned int i = 0; i < n_seeds; i++)
        s = __fadd_rn(s, pair_distance<0>(seeds + (size_t)i * dim, b, dim, mode, w_or_m));
    keys[j] = s;
}

// composite 64-bit key = (ordered float bits << 32) | index  -> any sort is stable
__global__ void make_sort_keys_kernel(const float *__restrict__ keys, unsigned int n,
                                      unsigned long long *__restrict__ out) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned int u = __float_as_uint(keys[j]);
    if (u == 0x80000000u) u = 0u;                     // n32: -0.0 == +0.0
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // total order of finite floats
    out[j] = ((unsigned long long)u << 32) | (unsigned long long)j;
}

__global__ void unpack_order_kernel(const unsigned long long *__restrict__ sorted, unsigned int n,
                                    unsigned int *__restrict__ order) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) order[j] = (unsigned int)(sorted[j] & 0xffffffffull);
}

// argmin (first minimum) over `alive` candidates of dist(cur, cand): one CTA, used by song_to_song
__global__ void __launch_bounds__(1024)
nearest_alive_kernel(const float *__restrict__ cur, unsigned int n_cur, const float *__restrict__ cands,
                     unsigned int n_cands, int dim, int mode, const float *__restrict__ w_or_m,
                     unsigned char *__restrict__ alive, unsigned int *__restrict__ order,
                     unsigned int step, float *__restrict__ next_cur) {
    __shared__ float s_v[32];
    __shared__ unsigned int s_i[32];
    float best = INFINITY;
    unsigned int bi = 0xffffffffu;
    for (unsigned int j = threadIdx.x; j < n_cands; j += blockDim.x) {
        if (!alive[j]) continue;
        float s = 0.f;
        for (unsigned int i = 0; i < n_cur; i++)
            s = __fadd_rn(s, pair_distance<0>(cur + (size_t)i * dim, cands + (size_t)j * dim, dim, mode, w_or_m));
        if (s < best || bi == 0xffffffffu) { best = s; bi = j; }  // first minimum: j ascending per thread
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const unsigned int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0xffffffffu && (bi == 0xffffffffu || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned int w = 1; w < blockDim.x / 32; w++) {
            const float ov = s_v[w];
            const unsigned int oi = s_i[w];
            if (oi != 0xffffffffu && (bi == 0xffffffffu || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
        }
        order[step] = bi;
        alive[bi] = 0;
        for (int i = 0; i < dim; i++) next_cur[i] = cands[(size_t)bi * dim + i];
    }
}

// ---- launchers ---------------------------------------------------------------
// ones_mask: bit i set where the diagonal weight w_i is exactly 1 (or no weights at all); variant bit
// VARIANT_OLD_DIST (65536) keeps the round-1 kernel for A/B and for the bit-exactness test of the packed one
int launch_distance_matrix(const float *rows, unsigned int n_rows, const float *cols, unsigned int n_cols,
                           int dim, int mode, const float *w_or_m, float *out, cudaStream_t st, unsigned int ones_mask,
                           int variant) {
    const bool DIST_DIAG_FAST_PATH = (variant & VARIANT_OLD_DIST) == 0;
    (void)ones_mask;
    (void)DIST_DIAG_FAST_PATH;
    if (n_rows == 0 || n_cols == 0) return 0;
#ifndef BLISS_HOST_EMUL
    if (mode == 0 && (dim == 23 || dim == 20) && (DIST_DIAG_FAST_PATH)) {  // diagonal metric: the packed kernel
        dim3 g2((n_cols + 255u) / 256u, (n_rows + 63u) / 64u);
        constexpr unsigned int V2_ONES = 0x3FEu;  // VERSION2_WEIGHTS (src/lib.rs:209-234): w_1 .. w_9 = 1
        const unsigned int all = (1u << dim) - 1u;
        if (!w_or_m) ones_mask = all;
        if (dim == 23) {
            if ((ones_mask & all) == all) BLISS_LAUNCH((distance_matrix_diag_kernel<23, 0x7FFFFFu>), g2, 128, 0, st, rows, n_rows, cols, n_cols, w_or_m, out, 1.0f);
            else if ((ones_mask & V2_ONES) == V2_ONES) BLISS_LAUNCH((distance_matrix_diag_kernel<23, V2_ONES>), g2, 128, 0, st, rows, n_rows, cols, n_cols, w_or_m, out, 1.0f);
            else BLISS_LAUNCH((distance_matrix_diag_kernel<23, 0u>), g2, 128, 0, st, rows, n_rows, cols, n_cols, w_or_m, out, 1.0f);
        } else {
            if ((ones_mask & all) == all) BLISS_LAUNCH((distance_matrix_diag_kernel<20, 0xFFFFFu>), g2, 128, 0, st, rows, n_rows, cols, n_cols, w_or_m, out, 1.0f);
            else BLISS_LAUNCH((distance_matrix_diag_kernel<20, 0u>), g2, 128, 0, st, rows, n_rows, cols, n_cols, w_or_m, out, 1.0f);
        }
        return 1;
    }
#endif
    if (n_rows > 32u * 65535u) return -1;  // (2 M rows per call: the callers block their rows long before that)
    dim3 grid((n_cols + 255u) / 256u, (n_rows + 31u) / 32u);
    if (dim == 23) BLISS_LAUNCH(distance_matrix_kernel<23>, grid, 128, 0, st, rows, n_rows, cols, n_cols, mode, w_or_m, out);
    else if (dim == 20) BLISS_LAUNCH(distance_matrix_kernel<20>, grid, 128, 0, st, rows, n_rows, cols, n_cols, mode, w_or_m, out);
    else if (dim <= MAX_DIM) BLISS_LAUNCH(distance_matrix_generic_kernel, dim3((n_cols + 255u) / 256u, n_rows < 65535u ? n_rows : 65535u), 256, 0, st, rows, n_rows, cols, n_cols, dim, mode, w_or_m, out);
    else return -1;
    return 1;
}

int launch_seed_distance(const float *seeds, unsigned int n_seeds, const float *cands, unsigned int n_cands,
                         int dim, int mode, const float *w_or_m, float *keys, cudaStream_t st) {
    if (n_cands == 0) return 0;
    BLISS_LAUNCH(seed_distance_kernel, (n_cands + 255u) / 256u, 256, 0, st, seeds, n_seeds, cands, n_cands, dim, mode,
                                                                  w_or_m, keys);
    return 1;
}

// stable ascending order of keys -> order[]; tmp buffers supplied by the caller
size_t sort_temp_bytes(unsigned int n) {
    size_t bytes = 0;
#ifdef BLISS_HOST_EMUL
    bytes = 16 + (size_t)n;  // host emulation: std::sort stands in for the device radix sort, no scratch needed
#else
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const unsigned long long *)nullptr,
                                   (unsigned long long *)nullptr, (int)n);
#endif
    return bytes;
}

int launch_stable_argsort(const float *keys, unsigned int n, unsigned long long *k_in,
                          unsigned long long *k_out, void *tmp, size_t tmp_bytes, unsigned int *order,
                          cudaStream_t st) {
    if (n == 0) return 0;
    BLISS_LAUNCH(make_sort_keys_kernel, (n + 255u) / 256u, 256, 0, st, keys, n, k_in);
#ifdef BLISS_HOST_EMUL
    (void)tmp; (void)tmp_bytes;
    std::copy(k_in, k_in + n, k_out);
    std::sort(k_out, k_out + n);  // the keys carry the index in their low half: any correct sort gives the stable order
#else
    cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, k_in, k_out, (int)n, 0, 64, st);
#endif
    BLISS_LAUNCH(unpack_order_kernel, (n + 255u) / 256u, 256, 0, st, k_out, n, order);
    return 3;
}

int launch_nearest_alive(const float *cur, unsigned int n_cur, const float *cands, unsigned int n_cands,
                         int dim, int mode, const float *w_or_m, unsigned char *alive, unsigned int *order,
                         unsigned int step, float *next_cur, cudaStream_t st) {
    BLISS_LAUNCH(nearest_alive_kernel, 1, 1024, 0, st, cur, n_cur, cands, n_cands, dim, mode, w_or_m, alive, order, step,
                                             next_cur);
    return 1;
}

}  // namespace bliss

// tempo.cu -- K7 (onset peak-picker thresholding) and K8 (beat tracker + BPM median).
// Compiled with -fmad=false: the reference (Rust, via its aubio transcription) never
// fuses multiply-adds in this chain and the chain is full of discrete decisions.
//
//   K7  PeakPicker::do_ thresholded value   src/aubio.rs:733-768 (+ filtfilt :661-685,
//       quickselect median :482-554)        -- each hop depends only on of[t-6..t]
//   K8  Tempo::do_ glue                     src/aubio.rs:1378-1443
//       BeatTracking::do_ / checkstate      src/aubio.rs:966-1227
//       BPMDesc::do_ / get_value            src/temporal.rs:50-77
//
// K8 is one CTA per song: the state machine is sequential over ~n_t/128 cycles,
// every cycle's autocorrelation / comb filterbank / phase histogram is spread over
// the CTA, keeping each element's f32 accumulation order identical to the reference's loops.
#include "common.cuh"

namespace bliss {

// ------------------------------- K7 ----------------------------------------
__global__ void __launch_bounds__(256)
peakpick_kernel(const float *__restrict__ flux, const SongDesc *__restrict__ songs,
                const unsigned int *__restrict__ t_prefix, int n_songs, unsigned int total,
                float *__restrict__ thr_out) {
    const unsigned int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int lo = find_song(t_prefix, n_songs, gid);  // t_prefix = exclusive prefix of n_t
    const SongDesc sd = songs[lo];
    const int t = (int)(gid - t_prefix[lo]);
    const float *of = flux + sd.t_off;
    float p[7], tmp[7];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const int u = t - 6 + i;
        p[i] = (u >= 0) ? of[u] : 0.f;  // onset_keep starts as zeros (aubio.rs:719)
    }
    const float b0 = 0.1599879f, b1 = 0.31997577f, b2 = 0.1599879f, a1 = 0.23484048f, a2 = 0.0f;
    // forward pass (state reset before and after, aubio.rs:661-685)
    {
        float x1 = 0.f, x2 = 0.f, y1 = 0.f, y2 = 0.f;
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const float x0 = p[i];
            const float y0 = b0 * x0 + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
            x2 = x1; x1 = x0; y2 = y1; y1 = y0;
            p[i] = y0;
        }
    }
#pragma unroll
    for (int i = 0; i < 7; i++) tmp[6 - i] = p[i];
    {
        float x1 = 0.f, x2 = 0.f, y1 = 0.f, y2 = 0.f;
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const float x0 = tmp[i];
            const float y0 = b0 * x0 + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
            x2 = x1; x1 = x0; y2 = y1; y1 = y0;
            tmp[i] = y0;
        }
    }
#pragma unroll
    for (int i = 0; i < 7; i++) p[i] = tmp[6 - i];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 7; i++) sum += p[i];
    const float mean = sum / 7.f;
    // median of 7 = element of rank 3 (what aubio's quickselect returns)
    float median = p[0];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int less = 0, eq_before = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            less += (p[k] < p[i]) ? 1 : 0;
            eq_before += (k < i && p[k] == p[i]) ? 1 : 0;
        }
        if (less + eq_before == 3) median = p[i];
    }
    thr_out[sd.t_off + t] = p[5] - median - mean * 0.3f;  // threshold 0.3: aubio.rs:1347
}

// ------------------------------- K8 ----------------------------------------
// BT threads per song.  Every phase is written as a loop over its 512 / 256 / 128 elements with stride BT,
// so the same code runs with one element per thread (BT = 512, the previous layout, kept as
// VARIANT_BT512) or with 128 threads and up to four elements each (current): the kernel is a chain of
// ~17 barrier-separated, latency-bound phases per cycle, most of them 128 wide or single-threaded, so what
// counts is how many songs are resident per SM (8 CTAs of 128 threads instead of 3 of 512) and how
// cheap a barrier is, not how wide one song runs.
struct __align__(16) BtShared {
    float df[512], dfrev[512], acf[512], phout[512];
    float dfwv[512];
    float acfout[128], rwv[128], gwv[128], out[128];
    float phwv[256];
    float red_v[16];
    int red_i[16];
    // scalar state (aubio.rs:834-862)
    unsigned int timesig;
    float lastbeat;
    int counter;
    unsigned int flagstep;
    float gp, bp, rp, rp1, rp2;
    int mode;       // 0 flagconst, 1 context, 2 initial
    int phase_gauss;
    int maxidx;
    unsigned int n_bpm;
};

// aubio.rs:787-799 vec_max_elem: last index of the maximum, 0.0 is the floor
// (returns 0 when every element is negative).  Each thread brings the best (value, index) of its own
// elements (ties: larger index); all threads must call.
template <int BT>
__device__ int block_max_elem(BtShared &sh, float v, int idx, bool active) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float bv = active ? v : -INFINITY;
    int bi = active ? idx : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sh.red_v[warp] = bv; sh.red_i[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mv = sh.red_v[0];
        int mi = sh.red_i[0];
        for (int w = 1; w < BT / 32; w++) {
            if (sh.red_v[w] > mv || (sh.red_v[w] == mv && sh.red_i[w] > mi)) { mv = sh.red_v[w]; mi = sh.red_i[w]; }
        }
        sh.maxidx = (mv >= 0.f && mi >= 0) ? mi : 0;
    }
    __syncthreads();
    return sh.maxidx;
}

// aubio.rs:576-604
__device__ float quadratic_peak_pos(const float *x, int len, int pos) {
    if (pos == 0 || pos >= len - 1) return (float)pos;
    const float s0 = x[pos - 1], s1 = x[pos], s2 = x[pos + 1];
    return (float)pos + 0.5f * (s0 - s2) / (s0 - 2.0f * s1 + s2);
}

// aubio.rs:864-907
__device__ unsigned int get_timesig(const float *acf, int acflen, int gp) {
    if (gp < 2) return 4;
    float three = 0.f, four = 0.f;
    if (acflen > 6 * gp + 2) {
        for (int k = -2; k < 2; k++) {
            three += acf[3 * gp + k];
            four += acf[4 * gp + k];
        }
    } else {
        for (int k = -2; k < 2; k++) {
            const int i3 = 3 * gp + k, i6 = 6 * gp + k, i4 = 4 * gp + k, i2 = 2 * gp + k;
            if (i3 < acflen && i6 < acflen) three += acf[i3] + acf[i6];
            else if (i3 < acflen) three += acf[i3];
            if (i4 < acflen && i2 < acflen) four += acf[i4] + acf[i2];
            else if (i4 < acflen) four += acf[i4];
        }
    }
    return three > four ? 3u : 4u;
}

template <int BT>
__global__ void __launch_bounds__(BT, BT == 512 ? 3 : 8)
beattrack_kernel(const float *__restrict__ thr, const float *__restrict__ block_energy,
                 const SongDesc *__restrict__ songs, float *__restrict__ bpm_list,
                 float *__restrict__ tempo_feature, unsigned int *__restrict__ bpm_count, int scalar_acf) {
    static_assert(BT >= 128 && BT <= 512 && BT % 32 == 0, "128-wide phases run as `if (tid < 128)`");
    __shared__ BtShared sh;
    const SongDesc sd = songs[blockIdx.x];
    const int tid = threadIdx.x;
    if (!sd.valid) {
        if (tid == 0) { tempo_feature[blockIdx.x] = -1.f; bpm_count[blockIdx.x] = 0; }
        return;
    }
    const float *D = thr + sd.t_off;          // thresholded onset function, D[t]
    const float *eb = block_energy + sd.e_off;
    float *bpms = bpm_list + sd.bpm_off;
    const int n_t = (int)sd.n_t;
    const int step = 128, winlen = 512, laglen = 128;
    const float g_var = 3.901f;

    // BeatTracking::new, aubio.rs:911-962
    const float rayparam_f = 60.0f * (float)SAMPLE_RATE / 120.0f / 256.0f;
    const unsigned int rayparam_u = (unsigned int)rayparam_f;
    {
        const float dfwvnorm = expf((logf(2.0f) / rayparam_f) * (float)(winlen + 2));
        for (int e = tid; e < winlen; e += BT) sh.dfwv[e] = expf((logf(2.0f) / rayparam_f) * (float)(e + 1)) / dfwvnorm;
        if (tid < laglen) {
            const float i_f = (float)(tid + 1);
            const float r2 = rayparam_f * rayparam_f;
            sh.rwv[tid] = (i_f / r2) * expf(-(i_f * i_f) / (2.0f * r2));
            sh.gwv[tid] = 0.f;
            sh.out[tid] = 0.f;
        }
        for (int e = tid; e < 2 * laglen; e += BT) sh.phwv[e] = 1.0f;
        if (tid == 0) {
            sh.timesig = 0; sh.lastbeat = 0.f; sh.counter = 0; sh.flagstep = 0;
            sh.gp = 0.f; sh.bp = 0.f; sh.rp = 1.0f; sh.rp1 = 0.f; sh.rp2 = 0.f;
            sh.n_bpm = 0;
        }
    }
    __syncthreads();

    const int cycles = n_t / 128;  // cycle c runs at hop 128c-1 (aubio.rs:1390-1406)
    for (int c = 1; c <= cycles; c++) {
        // dfframe[j] = DD[128c - 512 + j], DD[u] = thr[u-1] (u >= 1), 0 otherwise:
        // the very first thresholded value lands in dfframe[385] (blockpos pre-increment)
        for (int e = tid; e < winlen; e += BT) {
            const int u = 128 * c - 512 + e;
            sh.df[e] = (u >= 1) ? D[u - 1] : 0.f;
        }
        __syncthreads();
        const unsigned int timesig0 = sh.timesig;
        const int numelem = timesig0 == 0 ? 4 : (int)timesig0;
        // gp of the previous cycle.  Read HERE, behind the barrier above: thread 0 rewrites sh.gp in
        // checkstate below, and when gp <= 0 there is no barrier between that write and a read placed
        // there -- a late warp would see the new gp, take the other side of a branch that contains
        // barriers and desynchronise the CTA (seen as run-to-run tempo flips on 2-3 % of 1024 songs).
        const float gp_in = sh.gp;
        // dfrev = reverse(df * dfwv)
        for (int e = tid; e < winlen; e += BT) sh.dfrev[winlen - 1 - e] = sh.df[e] * sh.dfwv[e];
        // vec_autocorr, aubio.rs:819-828:  acf[i] = sum_{j < 512-i} df[j] df[j+i] / (512 - i), each lag's sum in
        // the reference's order (j ascending, un-fused multiply and add).
        if (scalar_acf == 1) {  // VARIANT_OLD_ACF: one lag at a time, two scalar loads per multiply-add
            for (int lag = tid; lag < winlen; lag += BT) {
                float tmp = 0.f;
                const float *a = sh.df, *b = sh.df + lag;
                const int cnt = winlen - lag;
                int j = 0;
                for (; j + 8 <= cnt; j += 8) {  // products first (independent), then the ordered adds
                    float pr[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) pr[u] = a[j + u] * b[j + u];
#pragma unroll
                    for (int u = 0; u < 8; u++) tmp += pr[u];
                }
                for (; j < cnt; j++) tmp += a[j] * b[j];
                sh.acf[lag] = tmp / (float)cnt;
            }
        } else if (scalar_acf == 0) {
            // Balanced cut (round 2).  With four consecutive lags per thread the work per thread falls linearly with
            // the lag (128 .. 1 blocks): warp 0 of EVERY resident CTA carries 40 % of its song's multiply-adds, and warp w
            // of every CTA lives on sub-partition w -- one of the SM's four schedulers did 4 x the work of another
            // (the kernel's 3.0 ms were that scheduler's queue).  Here thread t owns the lag pair (2t, 2t + 1) AND its
            // mirror pair (510 - 2t, 511 - 2t): 1026 multiply-adds for every thread, all four warps alike.  Each lag's
            // sum keeps the reference's order (j ascending, un-fused multiply and add): bit-identical results.
            for (int t = tid; t < 128; t += BT) {
                const float4 *df4 = reinterpret_cast<const float4 *>(sh.df);
                const float2 *df2 = reinterpret_cast<const float2 *>(sh.df);
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    const int i0 = half == 0 ? 2 * t : 510 - 2 * t;  // even: lags i0, i0 + 1
                    float acc0 = 0.f, acc1 = 0.f;
                    // blocks of four j with every term of both lags inside the frame: 4 jq + 3 + (i0 + 1) <= 511
                    const int nblk = (i0 <= 507) ? (507 - i0) / 4 + 1 : 0;
                    float2 w01 = df2[i0 / 2];                                    // df[i0 + j .. i0 + j + 3], j = 0
                    float2 w23 = (i0 / 2 + 1 < 256) ? df2[i0 / 2 + 1] : make_float2(0.f, 0.f);
                    for (int jq = 0; jq < nblk; jq++) {
                        const float4 a = df4[jq];
                        const float2 w45 = df2[i0 / 2 + 2 * jq + 2];             // df[i0 + j + 4], df[i0 + j + 5]
                        const float2 w67 = (i0 / 2 + 2 * jq + 3 < 256) ? df2[i0 / 2 + 2 * jq + 3] : make_float2(0.f, 0.f);
                        const float p00 = a.x * w01.x, p01 = a.x * w01.y;
                        const float p10 = a.y * w01.y, p11 = a.y * w23.x;
                        const float p20 = a.z * w23.x, p21 = a.z * w23.y;
                        const float p30 = a.w * w23.y, p31 = a.w * w45.x;
                        acc0 += p00; acc1 += p01;
                        acc0 += p10; acc1 += p11;
                        acc0 += p20; acc1 += p21;
                        acc0 += p30; acc1 += p31;
                        w01 = w45;
                        w23 = w67;
                    }
                    // the last terms, in order: j = 4 nblk .. 511 - i0 (lag i0) / 510 - i0 (lag i0 + 1)
                    for (int j = 4 * nblk; j + i0 <= 511; j++) {
                        acc0 += sh.df[j] * sh.df[j + i0];
                        if (j + i0 + 1 <= 511) acc1 += sh.df[j] * sh.df[j + i0 + 1];
                    }
                    sh.acf[i0] = acc0 / (float)(winlen - i0);
                    sh.acf[i0 + 1] = acc1 / (float)(winlen - i0 - 1);
                }
            }
        } else {
            // Thread t < 128 owns the four consecutive lags 4t..4t+3 and walks j in blocks of four: one
            // broadcast float4 of df[j..j+3] and one float4 of df[4t+j+4..+7] per 16 multiply-adds (the
            // sliding window df[4t+j..+6] stays in registers) instead of two scalar loads per multiply-add.
            for (int t = tid; t < 128; t += BT) {
                const int i0 = 4 * t;
                const float4 *df4 = reinterpret_cast<const float4 *>(sh.df);
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const int nblk = 128 - t;  // (512 - i0) / 4 blocks; the last one is partial for lags > i0
                float4 bc = df4[t];        // df[i0 + j .. i0 + j + 3], j = 0
                for (int jq = 0; jq < nblk - 1; jq++) {
                    const float4 a = df4[jq];
                    const float4 bn = df4[t + jq + 1];
                    const float av[4] = {a.x, a.y, a.z, a.w};
                    const float w[7] = {bc.x, bc.y, bc.z, bc.w, bn.x, bn.y, bn.z};
                    float pr[4][4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int m = 0; m < 4; m++) pr[u][m] = av[u] * w[u + m];
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int m = 0; m < 4; m++) acc[m] += pr[u][m];
                    bc = bn;
                }
                {   // last block: j = 508 - i0; lag i0 + m still has terms u = 0 .. 3 - m
                    const float4 a = df4[nblk - 1];
                    const float av[4] = {a.x, a.y, a.z, a.w};
                    const float w[4] = {bc.x, bc.y, bc.z, bc.w};
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int m = 0; m < 4; m++)
                            if (u + m < 4) acc[m] += av[u] * w[u + m];
                }
#pragma unroll
                for (int m = 0; m < 4; m++) sh.acf[i0 + m] = acc[m] / (float)(winlen - i0 - m);
            }
        }
        __syncthreads();  // acf complete before the comb filterbank reads arbitrary lags
        // shift-invariant comb filterbank, general model (aubio.rs:992-1003)
        float myv = 0.f;
        if (tid < laglen) {
            float acc = 0.f;
            if (tid >= 1 && tid < laglen - 1) {
                for (int a = 1; a <= numelem; a++)
                    for (int b = 1; b < 2 * a; b++) {
                        const int idx = tid * a + b - 1;
                        if (idx < winlen) acc += sh.acf[idx] / (2.0f * (float)a - 1.0f);
                    }
            }
            acc *= sh.rwv[tid];
            sh.acfout[tid] = acc;
            myv = acc;
        }
        __syncthreads();
        int maxindex = block_max_elem<BT>(sh, myv, tid, tid < laglen);
        if (tid == 0) {
            if (maxindex > 0 && maxindex < laglen - 1) sh.rp = quadratic_peak_pos(sh.acfout, laglen, maxindex);
            else sh.rp = (float)rayparam_u;
        }
        __syncthreads();
        // ---- checkstate, aubio.rs:1096-1227 ----
        if (gp_in > 0.f) {  // context-dependent comb (no 1/(2a-1)), Gaussian weighting
            float acc = 0.f;
            if (tid < laglen) {
                if (tid >= 1 && tid < laglen - 1) {
                    for (unsigned int a = 1; a <= timesig0; a++)
                        for (unsigned int b = 1; b < 2 * a; b++) {
                            const int idx = tid * (int)a + (int)b - 1;
                            if (idx < winlen) acc += sh.acf[idx];
                        }
                }
                acc *= sh.gwv[tid];
            }
            __syncthreads();
            if (tid < laglen) sh.acfout[tid] = acc;
            __syncthreads();
            maxindex = block_max_elem<BT>(sh, acc, tid, tid < laglen);
        }
        if (tid == 0) {
            int counter = sh.counter;
            unsigned int flagstep = sh.flagstep;
            float gp = gp_in;
            const float rp = sh.rp;
            float rp1 = sh.rp1, rp2 = sh.rp2;
            bool flagconst = false;
            if (gp > 0.f) gp = quadratic_peak_pos(sh.acfout, laglen, maxindex);
            else gp = 0.f;
            if (counter == 0) {
                if (fabsf(gp - rp) > 2.0f * g_var) { flagstep = 1; counter = 3; }
                else flagstep = 0;
            }
            if (counter == 1 && flagstep == 1) {
                if (fabsf(2.0f * rp - rp1 - rp2) < g_var) { flagconst = true; counter = 0; }
                else { flagconst = false; counter = 2; }
            } else if (counter > 0) {
                counter -= 1;
            }
            rp2 = rp1;
            rp1 = rp;
            float bp;
            int mode;
            int phase_gauss = 0;
            if (flagconst) {
                gp = rp;
                sh.timesig = get_timesig(sh.acf, winlen, (int)gp);
                bp = gp;
                mode = 0;
            } else if (sh.timesig > 0) {
                bp = gp;
                mode = 1;
                phase_gauss = ((float)step > sh.lastbeat) ? 1 : 0;
            } else {
                bp = rp;
                mode = 2;
            }
            sh.mode = mode;
            sh.phase_gauss = phase_gauss;
            sh.counter = counter; sh.flagstep = flagstep; sh.gp = gp;
            sh.rp1 = rp1; sh.rp2 = rp2;
            sh.bp = bp;  // pre-doubling value: the phase weighting below uses it (aubio.rs:1197)
        }
        __syncthreads();
        {
            const int mode = sh.mode;
            const float bp_pre = sh.bp, gp = sh.gp, lastbeat = sh.lastbeat;
            if (mode == 0 && tid < laglen) {
                const float diff = (float)(tid + 1) - gp;
                sh.gwv[tid] = expf(-0.5f * diff * diff / (g_var * g_var));
            }
            const bool gauss = (mode == 1 && sh.phase_gauss);
            for (int e = tid; e < 2 * laglen; e += BT) {
                if (gauss) {
                    const float diff = 1.0f + (float)e - (float)step + lastbeat;
                    sh.phwv[e] = expf(-0.5f * diff * diff / (bp_pre / 8.0f));
                } else {
                    sh.phwv[e] = 1.0f;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            float bp = sh.bp;
            while (bp > 0.f && bp < 25.f) bp *= 2.0f;  // aubio.rs:1216-1218
            sh.bp = bp;
        }
        __syncthreads();
        const float bp = sh.bp;
        if (bp == 0.f) {
            if (tid < step) sh.out[tid] = 0.f;
            __syncthreads();
        } else {
            // beat phase, aubio.rs:1026-1057
            const int kmax = (int)floorf((float)winlen / bp);
            float best_v = -INFINITY;
            int best_i = -1;
            for (int e = tid; e < winlen; e += BT) {
                float ph = 0.f;
                if ((float)e < bp) {
                    for (int k = 0; k < kmax; k++) {
                        const int idx = e + (int)floorf(bp * (float)k + 0.5f);
                        if (idx < winlen) ph += sh.dfrev[idx];
                    }
                }
                if (e < 2 * laglen) ph *= sh.phwv[e];
                sh.phout[e] = ph;
                if (ph >= best_v) { best_v = ph; best_i = e; }  // e ascends: ties keep the larger index
            }
            __syncthreads();
            maxindex = block_max_elem<BT>(sh, best_v, best_i, true);
            if (tid < step) sh.out[tid] = 0.f;
            __syncthreads();
            if (tid == 0) {
                float phase;
                if (maxindex >= winlen - 1) phase = (float)step - sh.lastbeat;
                else phase = quadratic_peak_pos(sh.phout, winlen, maxindex);
                phase += 1.0f;
                int i = 1;
                float beat = bp - phase;
                if (((float)step - sh.lastbeat - phase) < -0.40f * bp) beat += bp;
                while (beat + bp < 0.f) beat += bp;
                if (beat >= 0.f && i < step) { sh.out[i] = beat; i++; }
                while (beat + bp <= (float)step && i < step) {
                    beat += bp;
                    sh.out[i] = beat;
                    i++;
                }
                sh.lastbeat = beat;
                sh.out[0] = (float)i;
            }
            __syncthreads();
        }
        // Tempo::do_ steps 6 + BPMDesc::do_: hops t = 128c-1+p carry blockpos p
        if (tid < step) {
            const int t = 128 * c - 1 + tid;
            if (t < n_t) {
                const int num_beats = (int)sh.out[0];
                float tempo_out = 0.f;
                for (int i = 1; i < num_beats; i++) {
                    const float beat_pos = sh.out[i];
                    if (tid == (int)floorf(beat_pos)) {
                        tempo_out = beat_pos - floorf(beat_pos);
                        // is_silence over the 512-sample window handed to do_ (song/mod.rs:435-441)
                        const float lvl = (eb[t] + eb[t + 1]) / 512.f;
                        if (10.0f * log10f(lvl) < -90.0f) tempo_out = 0.f;
                    }
                }
                if (tempo_out > 0.f) {
                    // get_bpm, aubio.rs:1231-1239
                    const float period_samples = 256.f * bp;
                    const float period_s = period_samples / (float)SAMPLE_RATE;
                    const float bpm = (bp != 0.f) ? 60.0f / period_s : 0.f;
                    const unsigned int slot = atomicAdd(&sh.n_bpm, 1u);
                    bpms[slot] = bpm;
                }
            }
        }
        __syncthreads();
    }
    // BPMDesc::get_value, temporal.rs:66-77: Midpoint median, normalise by 206
    __threadfence_block();
    __syncthreads();
    const int nb = (int)sh.n_bpm;
    if (nb == 0) {
        if (tid == 0) { tempo_feature[blockIdx.x] = -1.f; bpm_count[blockIdx.x] = 0; }
        return;
    }
    const int r_lo = (nb - 1) / 2, r_hi = nb / 2;  // floor/ceil((nb-1)/2)
    if (tid == 0) { sh.red_v[0] = 0.f; sh.red_v[1] = 0.f; }
    __syncthreads();
    for (int i = tid; i < nb; i += BT) {
        const float v = bpms[i];
        int rank = 0;
        for (int k = 0; k < nb; k++) {
            const float w = bpms[k];
            rank += (w < v || (w == v && k < i)) ? 1 : 0;
        }
        if (rank == r_lo) sh.red_v[0] = v;
        if (rank == r_hi) sh.red_v[1] = v;
    }
    __syncthreads();
    if (tid == 0) {
        const float lower = sh.red_v[0], higher = sh.red_v[1];
        const float median = lower + (higher - lower) / 2.f;
        tempo_feature[blockIdx.x] = 2.f * (median - 0.f) / (206.f - 0.f) - 1.f;
        bpm_count[blockIdx.x] = (unsigned int)nb;
    }
}

int launch_peakpick(const float *flux, const SongDesc *songs, const unsigned int *t_prefix, int n_songs,
                    unsigned int total, float *thr, cudaStream_t st) {
    if (total == 0) return 0;
    BLISS_LAUNCH(peakpick_kernel, (total + 255u) / 256u, 256, 0, st, flux, songs, t_prefix, n_songs, total, thr);
    return 1;
}

int launch_beattrack(const float *thr, const float *block_energy, const SongDesc *songs, int n_songs,
                     float *bpm_list, float *tempo_feature, unsigned int *bpm_count, int variant,
                     cudaStream_t st) {
    if (n_songs == 0) return 0;
    // 0 = balanced lag pairs (round 2), 1 = one lag at a time (VARIANT_OLD_ACF), 2 = four consecutive lags per thread (round 1)
    const int scalar_acf = (variant & VARIANT_OLD_ACF) ? 1 : (variant & VARIANT_ACF_4LAGS) ? 2 : 0;
    if (variant & VARIANT_BT512)
        BLISS_LAUNCH(beattrack_kernel<512>, n_songs, 512, 0, st, thr, block_energy, songs, bpm_list, tempo_feature, bpm_count,
                                                       scalar_acf);
    else
        BLISS_LAUNCH(beattrack_kernel<128>, n_songs, 128, 0, st, thr, block_energy, songs, bpm_list, tempo_feature, bpm_count,
                                                       scalar_acf);
    return 1;
}

// Every kernel of a wave asks for the SAME L1 / shared-memory split (all shared): kernels with different carve-outs
// cannot share an SM, and the latency-bound kernels of one chain are meant to run under the FFT kernels of the other
// (api.cu run_wave).  Called once per device from bliss_b200_init.
#ifndef BLISS_HOST_EMUL
#define BLISS_MAX_SHARED(kern) (void)cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)
#else
#define BLISS_MAX_SHARED(kern) (void)0
#endif
void configure_kernels_tempo() {
    BLISS_MAX_SHARED(peakpick_kernel);
    BLISS_MAX_SHARED(beattrack_kernel<128>);
}

}  // namespace bliss

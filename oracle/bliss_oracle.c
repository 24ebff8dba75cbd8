/*
 * bliss_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See bliss_oracle.h.  Every function cites the reference file:line it restates
 * (paths relative to /root/reference).  Build with -ffp-contract=off: Rust never
 * fuses a*b+c unless mul_add is written, and the places where the reference
 * does write mul_add use fmaf() explicitly here.
 */
#define _GNU_SOURCE
#include "bliss_oracle.h"

#include <malloc.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define PI_F 3.14159274101257324f /* std::f32::consts::PI */

/* ======================================================================== */
/* complex f32 FFT (stands in for rustfft 6.4.1; Cargo.lock:1321)           */
/* radix-2 decimation in time, f64-generated twiddles rounded to f32.        */
/* ======================================================================== */
typedef struct {
    uint32_t n;
    uint32_t *rev;
    float *twr, *twi; /* per-stage contiguous twiddles: stage with half-size h starts at offset h-1 */
} fft_plan;

static fft_plan g_plans[8];
static int g_nplans = 0;
static pthread_mutex_t g_plan_mu = PTHREAD_MUTEX_INITIALIZER;

/* Plans are append-only: the hot path scans the published ones without taking the lock (a global
 * mutex per FFT call serialised the ~50 000 transforms per song of every worker thread). */
static const fft_plan *get_plan(uint32_t n) {
    int np = __atomic_load_n(&g_nplans, __ATOMIC_ACQUIRE);
    for (int i = 0; i < np; i++)
        if (g_plans[i].n == n) return &g_plans[i];
    pthread_mutex_lock(&g_plan_mu);
    for (int i = 0; i < g_nplans; i++)
        if (g_plans[i].n == n) {
            pthread_mutex_unlock(&g_plan_mu);
            return &g_plans[i];
        }
    if (g_nplans == 8) abort();
    fft_plan *p = &g_plans[g_nplans];
    p->n = n;
    p->rev = (uint32_t *)malloc(sizeof(uint32_t) * n);
    p->twr = (float *)malloc(sizeof(float) * n);
    p->twi = (float *)malloc(sizeof(float) * n);
    uint32_t bits = 0;
    while ((1u << bits) < n) bits++;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t r = 0;
        for (uint32_t b = 0; b < bits; b++)
            if (i & (1u << b)) r |= 1u << (bits - 1 - b);
        p->rev[i] = r;
    }
    for (uint32_t half = 1; half < n; half <<= 1)
        for (uint32_t k = 0; k < half; k++) {
            double a = -2.0 * M_PI * (double)k / (double)(2 * half);
            p->twr[half - 1 + k] = (float)cos(a);
            p->twi[half - 1 + k] = (float)sin(a);
        }
    __atomic_store_n(&g_nplans, g_nplans + 1, __ATOMIC_RELEASE);
    pthread_mutex_unlock(&g_plan_mu);
    return p;
}

/* radix-2 decimation in time; every stage streams unit-stride through data and twiddles so the
 * compiler can vectorise the butterflies (-O3 -march=native) */
void bo_fft(float *restrict d, uint32_t n) {
    const fft_plan *p = get_plan(n);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t r = p->rev[i];
        if (r > i) {
            float tr = d[2 * i], ti = d[2 * i + 1];
            d[2 * i] = d[2 * r];
            d[2 * i + 1] = d[2 * r + 1];
            d[2 * r] = tr;
            d[2 * r + 1] = ti;
        }
    }
    for (uint32_t half = 1; half < n; half <<= 1) {
        const float *restrict wr = p->twr + (half - 1), *restrict wi = p->twi + (half - 1);
        for (uint32_t base = 0; base < n; base += 2 * half) {
            float *restrict a = d + 2 * base;
            float *restrict b = a + 2 * half;
            for (uint32_t k = 0; k < half; k++) {
                float br = b[2 * k], bi = b[2 * k + 1];
                float tr = br * wr[k] - bi * wi[k];
                float ti = br * wi[k] + bi * wr[k];
                float ar = a[2 * k], ai = a[2 * k + 1];
                a[2 * k] = ar + tr;
                a[2 * k + 1] = ai + ti;
                b[2 * k] = ar - tr;
                b[2 * k + 1] = ai - ti;
            }
        }
    }
}

/* ======================================================================== */
/* utils.rs                                                                  */
/* ======================================================================== */

/* utils.rs:66-68: plain f32 sum / n */
static float mean_f32(const float *v, uint32_t n) {
    float s = 0.f;
    for (uint32_t i = 0; i < n; i++) s += v[i];
    return s / (float)n;
}

/* ndarray 0.17 var_axis(ddof=0) -> std: one-pass Welford in f32 with mul_add
 * (called at timbral.rs:61-63, 85-87, 115-117 and misc.rs:52). */
static float std_pop_f32(const float *v, uint32_t n) {
    float mean = 0.f, sum_sq = 0.f;
    for (uint32_t i = 0; i < n; i++) {
        float count = (float)(i + 1);
        float x = v[i];
        float delta = x - mean;
        mean = mean + delta / count;
        sum_sq = fmaf(x - mean, delta, sum_sq);
    }
    return sqrtf(sum_sq / (float)n);
}

/* utils.rs:70-77 */
static float normalize(float value, float min_v, float max_v) {
    return 2.f * (value - min_v) / (max_v - min_v) - 1.f;
}

/* utils.rs:11-24: reflect, edge sample excluded */
void bo_reflect_pad(const float *x, uint64_t n, uint32_t pad, float *out) {
    for (uint32_t i = 0; i < pad; i++) out[i] = x[pad - i];
    memcpy(out + pad, x, sizeof(float) * n);
    for (uint32_t i = 0; i < pad; i++) out[pad + n + i] = x[n - 2 - i];
}

/* utils.rs:30: (len as f32 / hop as f32).ceil() as usize */
uint32_t bo_stft_num_frames(uint64_t n, uint32_t hop) {
    return (uint32_t)ceilf((float)n / (float)hop);
}

/* utils.rs:26-64.  Output frame-major (the reference permutes axes at :63). */
uint32_t bo_stft(const float *x, uint64_t n, uint32_t win, uint32_t hop, double *out) {
    uint32_t frames = bo_stft_num_frames(n, hop);
    uint32_t bins = win / 2 + 1;
    memset(out, 0, sizeof(double) * (size_t)frames * bins);
    uint64_t np = n + 2 * (uint64_t)(win / 2);
    float *padded = (float *)malloc(sizeof(float) * np);
    bo_reflect_pad(x, n, win / 2, padded);
    float *hann = (float *)malloc(sizeof(float) * win);
    for (uint32_t i = 0; i < win; i++)
        hann[i] = 0.5f - 0.5f * cosf(2.f * (float)i * PI_F / (float)win);
    float *buf = (float *)malloc(sizeof(float) * 2 * win);
    uint64_t n_windows = (np - win) / hop + 1; /* .windows(win).step_by(hop) */
    for (uint32_t f = 0; f < frames && f < n_windows; f++) { /* zip truncates */
        const float *w = padded + (uint64_t)f * hop;
        for (uint32_t i = 0; i < win; i++) {
            buf[2 * i] = w[i] * hann[i];
            buf[2 * i + 1] = 0.f;
        }
        bo_fft(buf, win);
        double *row = out + (size_t)f * bins;
        for (uint32_t k = 0; k < bins; k++) {
            float re = buf[2 * k], im = buf[2 * k + 1];
            row[k] = (double)sqrtf(re * re + im * im);
        }
    }
    free(buf);
    free(hann);
    free(padded);
    return frames;
}

/* utils.rs:81-95 */
uint32_t bo_number_crossings(const float *x, uint64_t n) {
    uint32_t crossings = 0;
    int was_positive = x[0] > 0.f;
    for (uint64_t i = 0; i < n; i++) {
        int is_positive = x[i] > 0.f;
        if (was_positive != is_positive) {
            crossings++;
            was_positive = is_positive;
        }
    }
    return crossings;
}

/* utils.rs:101-117 */
float bo_geometric_mean(const float *in, uint32_t n) {
    int32_t exponents = 0;
    double mantissas = 1.;
    for (uint32_t c = 0; c + 8 <= n; c += 8) {
        const float *ch = in + c;
        double m = ((double)ch[0] * (double)ch[1]) * ((double)ch[2] * (double)ch[3]);
        m *= 3.273390607896142e150; /* 2^500 */
        m *= ((double)ch[4] * (double)ch[5]) * ((double)ch[6] * (double)ch[7]);
        if (m == 0.) return 0.f;
        uint64_t bits;
        memcpy(&bits, &m, 8);
        exponents += (int32_t)(bits >> 52);
        uint64_t mb = (bits & 0xFFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
        double mm;
        memcpy(&mm, &mb, 8);
        mantissas *= mm;
    }
    return exp2f((log2f((float)mantissas) + (float)exponents) / (float)n - (1023.f + 500.f) / 8.f);
}

/* ======================================================================== */
/* aubio.rs, timbral part + timbral.rs                                       */
/* ======================================================================== */

/* aubio.rs:150-154 / :308-311: hanningz */
static void hanningz(float *w, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) w[i] = 0.5f * (1.0f - cosf(2.0f * PI_F * (float)i / (float)n));
}

/* Shared by PVoc::do_ (aubio.rs:182-264) and PVocTempo::do_ (:338-425):
 * frame k's 512-sample buffer is the last 512 samples ending at hop*(k+1)
 * (zeros before the start of the song), windowed, half-swapped, FFT'd. */
static void pvoc_fft(const float *x, uint32_t k, uint32_t hop, const float *win, float *buf /*1024*/) {
    float data[512];
    int64_t start = (int64_t)hop * ((int64_t)k + 1) - 512;
    for (int i = 0; i < 512; i++) {
        int64_t idx = start + i;
        data[i] = (idx < 0 ? 0.f : x[idx]) * win[i];
    }
    for (int j = 0; j < 256; j++) { /* fvec_shift, aubio.rs:219-229 */
        float t = data[j];
        data[j] = data[j + 256];
        data[j + 256] = t;
    }
    for (int i = 0; i < 512; i++) {
        buf[2 * i] = data[i];
        buf[2 * i + 1] = 0.f;
    }
    bo_fft(buf, 512);
}

/* aubio.rs:16-29 */
static float spectral_centroid(const float *norm, uint32_t n) {
    float sum = 0.f;
    for (uint32_t j = 0; j < n; j++) sum += norm[j];
    if (sum == 0.f) return 0.f;
    float sc = 0.f;
    for (uint32_t j = 0; j < n; j++) sc += (float)j * norm[j];
    return sc / sum;
}

/* aubio.rs:36-58 */
static float spectral_rolloff(const float *norm, uint32_t n) {
    float cumsum = 0.f, rollsum = 0.f;
    for (uint32_t j = 0; j < n; j++) cumsum += norm[j] * norm[j];
    if (cumsum == 0.f) return 0.f;
    cumsum *= 0.95f;
    uint32_t j = 0;
    while (rollsum < cumsum && j < n) {
        rollsum += norm[j] * norm[j];
        j++;
    }
    return (float)j;
}

/* aubio.rs:68-71 */
static float bin_to_freq(float bin, float sample_rate, float fft_size) {
    float freq = sample_rate / fft_size;
    return freq * (bin > 0.f ? bin : 0.f);
}

void bo_timbral_frames(const float *x, uint64_t n, uint32_t n_frames, float *centroid,
                       float *rolloff, float *flatness, float *norms_out) {
    (void)n;
    float win[512], buf[1024], norm[256];
    hanningz(win, 512);
    for (uint32_t k = 0; k < n_frames; k++) {
        pvoc_fft(x, k, 128, win, buf);
        /* aubio.rs:237-261: the "buggy" 256-bin cvec */
        norm[0] = fabsf(buf[0]);
        for (int i = 1; i < 255; i++) {
            float re = buf[2 * i], im = buf[2 * i + 1];
            norm[i] = sqrtf(re * re + im * im);
        }
        norm[255] = fabsf(buf[2 * 256]);
        if (norms_out) memcpy(norms_out + (size_t)k * 256, norm, sizeof(norm));
        /* timbral.rs:154-209 */
        centroid[k] = bin_to_freq(spectral_centroid(norm, 256), (float)BO_SAMPLE_RATE, 512.f);
        float bin = spectral_rolloff(norm, 256);
        if (bin > 256.f) bin = 256.f;
        rolloff[k] = bin_to_freq(bin, (float)BO_SAMPLE_RATE, 512.f);
        float gm = bo_geometric_mean(norm, 256);
        flatness[k] = (gm == 0.f) ? 0.f : gm / mean_f32(norm, 256);
    }
}

void bo_summarise(const float *v, uint32_t n, int kind, float out[2]) {
    float m = mean_f32(v, n), s = std_pop_f32(v, n);
    if (kind == 0) { /* timbral.rs:212-215 */
        out[0] = normalize(m, 0.f, (float)BO_SAMPLE_RATE / 2.f);
        out[1] = normalize(s, 0.f, (float)BO_SAMPLE_RATE / 2.f);
    } else { /* timbral.rs:104-122 */
        out[0] = 2.f * (m - 0.f) / (1.f - 0.f) - 1.f;
        out[1] = 2.f * (s - 0.f) / (1.f - 0.f) - 1.f;
    }
}

/* timbral.rs:231-258, one do_ call on the whole song (song/mod.rs:470-474) */
float bo_zcr(const float *x, uint64_t n) {
    uint32_t c = bo_number_crossings(x, n);
    return normalize((float)c / (float)n, 0.f, 1.f);
}

/* misc.rs:12-18 */
static float level_lin(const float *d, uint64_t n) {
    float e = 0.f;
    for (uint64_t i = 0; i < n; i++) e += d[i] * d[i];
    return e / (float)n;
}

/* misc.rs:39-71; chunks(1024) incl. short tail in Song::analyze
 * (song/mod.rs:478), chunks_exact in the reference's own unit test. */
void bo_loudness(const float *x, uint64_t n, int chunks_exact, float out[2]) {
    uint64_t nch = chunks_exact ? n / 1024 : (n + 1023) / 1024;
    float *v = (float *)malloc(sizeof(float) * (nch ? nch : 1));
    for (uint64_t c = 0; c < nch; c++) {
        uint64_t s = c * 1024, len = (s + 1024 <= n) ? 1024 : n - s;
        v[c] = level_lin(x + s, len);
    }
    float std_v = std_pop_f32(v, (uint32_t)nch);
    float mean_v = mean_f32(v, (uint32_t)nch);
    if (mean_v < 1e-9f) mean_v = 1e-9f;
    if (std_v < 1e-9f) std_v = 1e-9f;
    out[0] = normalize(10.0f * log10f(mean_v), -90.f, 0.f);
    out[1] = normalize(10.0f * log10f(std_v), -90.f, 0.f);
    free(v);
}

/* ======================================================================== */
/* aubio.rs, tempo part                                                      */
/* ======================================================================== */

static float vec_mean(const float *d, int n) { /* aubio.rs:471-478 */
    if (n == 0) return 0.f;
    float s = 0.f;
    for (int i = 0; i < n; i++) s += d[i];
    return s / (float)n;
}

static void swapf(float *a, float *b) {
    float t = *a;
    *a = *b;
    *b = t;
}

/* aubio.rs:482-554 quickselect median (lower median) */
static float vec_median(float *data, int n) {
    if (n == 0) return 0.f;
    int low = 0, high = n - 1;
    int median = (low + high) / 2;
    for (;;) {
        if (high <= low) return data[median];
        if (high == low + 1) {
            if (data[low] > data[high]) swapf(&data[low], &data[high]);
            return data[median];
        }
        int middle = (low + high) / 2;
        if (data[middle] > data[high]) swapf(&data[middle], &data[high]);
        if (data[low] > data[high]) swapf(&data[low], &data[high]);
        if (data[middle] > data[low]) swapf(&data[middle], &data[low]);
        swapf(&data[middle], &data[low + 1]);
        int ll = low + 1, hh = high;
        for (;;) {
            ll++;
            while (data[low] > data[ll]) ll++;
            hh--;
            while (data[hh] > data[low]) hh--;
            if (hh < ll) break;
            swapf(&data[ll], &data[hh]);
        }
        swapf(&data[low], &data[hh]);
        if (hh <= median) low = ll;
        if (hh >= median) high = hh - 1;
    }
}

/* aubio.rs:576-604 */
static float vec_quadratic_peak_pos(const float *x, int len, int pos) {
    if (pos == 0 || pos >= len - 1) return (float)pos;
    float s0 = x[pos - 1], s1 = x[pos], s2 = x[pos + 1];
    return (float)pos + 0.5f * (s0 - s2) / (s0 - 2.0f * s1 + s2);
}

/* aubio.rs:787-799: last index of the maximum, 0.0 is the floor */
static int vec_max_elem(const float *d, int n) {
    int pos = 0;
    float tmp = 0.f;
    for (int j = 0; j < n; j++)
        if (tmp <= d[j]) {
            pos = j;
            tmp = d[j];
        }
    return pos;
}

typedef struct {
    float b0, b1, b2, a1, a2, x1, x2, y1, y2;
} biquad;

static float biquad_step(biquad *f, float x0) { /* aubio.rs:636-649 */
    float y0 = f->b0 * x0 + f->b1 * f->x1 + f->b2 * f->x2 - f->a1 * f->y1 - f->a2 * f->y2;
    f->x2 = f->x1;
    f->x1 = x0;
    f->y2 = f->y1;
    f->y1 = y0;
    return y0;
}
static void biquad_reset(biquad *f) { f->x1 = f->x2 = f->y1 = f->y2 = 0.f; }

/* aubio.rs:661-685 */
static void do_filtfilt(biquad *f, float *data, float *tmp, int len) {
    for (int i = 0; i < len; i++) data[i] = biquad_step(f, data[i]);
    biquad_reset(f);
    for (int i = 0; i < len; i++) tmp[len - i - 1] = data[i];
    for (int i = 0; i < len; i++) tmp[i] = biquad_step(f, tmp[i]);
    biquad_reset(f);
    for (int i = 0; i < len; i++) data[i] = tmp[len - i - 1];
}

typedef struct { /* aubio.rs:692-779 */
    float threshold;
    biquad bq;
    float onset_keep[7], onset_proc[7], scratch[7], onset_peek[3];
    float thresholded;
} peakpicker;

static void pp_init(peakpicker *p) {
    memset(p, 0, sizeof(*p));
    p->threshold = 0.1f;
    p->bq.b0 = 0.1599879f;
    p->bq.b1 = 0.31997577f;
    p->bq.b2 = 0.1599879f;
    p->bq.a1 = 0.23484048f;
    p->bq.a2 = 0.0f;
}

static void pp_do(peakpicker *p, float onset) { /* aubio.rs:733-768; return value unused by Tempo */
    for (int i = 0; i < 6; i++) p->onset_keep[i] = p->onset_keep[i + 1];
    p->onset_keep[6] = onset;
    memcpy(p->onset_proc, p->onset_keep, sizeof(p->onset_keep));
    do_filtfilt(&p->bq, p->onset_proc, p->scratch, 7);
    float mean = vec_mean(p->onset_proc, 7);
    memcpy(p->scratch, p->onset_proc, sizeof(p->onset_proc));
    float median = vec_median(p->scratch, 7);
    p->onset_peek[0] = p->onset_peek[1];
    p->onset_peek[1] = p->onset_peek[2];
    p->thresholded = p->onset_proc[5] - median - mean * p->threshold;
    p->onset_peek[2] = p->thresholded;
}

#define BT_WINLEN 512
#define BT_LAGLEN 128
typedef struct { /* aubio.rs:834-862 */
    uint32_t hop_size, samplerate;
    float rwv[BT_LAGLEN], gwv[BT_LAGLEN], dfwv[BT_WINLEN], dfrev[BT_WINLEN], acf[BT_WINLEN],
        acfout[BT_LAGLEN], phwv[2 * BT_LAGLEN], phout[BT_WINLEN];
    uint32_t timesig, step, rayparam;
    float lastbeat;
    int32_t counter;
    uint32_t flagstep;
    float g_var, gp, bp, rp, rp1, rp2;
} beattracking;

/* aubio.rs:864-907 */
static uint32_t bt_get_timesig(const float *acf, int acflen, int gp) {
    if (gp < 2) return 4;
    float three = 0.f, four = 0.f;
    if (acflen > 6 * gp + 2) {
        for (int k = -2; k < 2; k++) {
            three += acf[3 * gp + k];
            four += acf[4 * gp + k];
        }
    } else {
        for (int k = -2; k < 2; k++) {
            int i3 = 3 * gp + k, i6 = 6 * gp + k, i4 = 4 * gp + k, i2 = 2 * gp + k;
            if (i3 < acflen && i6 < acflen) three += acf[i3] + acf[i6];
            else if (i3 < acflen) three += acf[i3];
            if (i4 < acflen && i2 < acflen) four += acf[i4] + acf[i2];
            else if (i4 < acflen) four += acf[i4];
        }
    }
    return three > four ? 3 : 4;
}

/* aubio.rs:911-962 */
static void bt_init(beattracking *b, uint32_t hop_size, uint32_t samplerate) {
    memset(b, 0, sizeof(*b));
    const int winlen = BT_WINLEN, laglen = BT_LAGLEN;
    float rayparam_float = 60.0f * (float)samplerate / 120.0f / (float)hop_size;
    b->rayparam = (uint32_t)rayparam_float;
    float dfwvnorm = expf((logf(2.0f) / rayparam_float) * (float)(winlen + 2));
    b->hop_size = hop_size;
    b->samplerate = samplerate;
    b->step = winlen / 4;
    float r2 = rayparam_float * rayparam_float;
    for (int i = 0; i < laglen; i++) {
        float i_f = (float)(i + 1);
        b->rwv[i] = (i_f / r2) * expf(-(i_f * i_f) / (2.0f * r2));
    }
    for (int i = 0; i < winlen; i++)
        b->dfwv[i] = expf((logf(2.0f) / rayparam_float) * (float)(i + 1)) / dfwvnorm;
    for (int i = 0; i < 2 * laglen; i++) b->phwv[i] = 1.0f;
    b->g_var = 3.901f;
    b->rp = 1.0f;
}

/* aubio.rs:1096-1227 */
static void bt_checkstate(beattracking *b) {
    const int laglen = BT_LAGLEN, acflen = BT_WINLEN;
    const uint32_t step = b->step;
    int32_t counter = b->counter;
    uint32_t flagstep = b->flagstep;
    float gp = b->gp, rp = b->rp, rp1 = b->rp1, rp2 = b->rp2, bp;
    int flagconst = 0;
    if (gp > 0.f) {
        for (int i = 0; i < laglen; i++) b->acfout[i] = 0.f;
        for (int i = 1; i < laglen - 1; i++)
            for (uint32_t a = 1; a <= b->timesig; a++)
                for (uint32_t bb = 1; bb < 2 * a; bb++) {
                    int idx = i * (int)a + (int)bb - 1;
                    if (idx < acflen) b->acfout[i] += b->acf[idx];
                }
        for (int i = 0; i < laglen; i++) b->acfout[i] *= b->gwv[i];
        int maxindex = vec_max_elem(b->acfout, laglen);
        gp = vec_quadratic_peak_pos(b->acfout, laglen, maxindex);
    } else {
        gp = 0.f;
    }
    if (counter == 0) {
        if (fabsf(gp - rp) > 2.0f * b->g_var) {
            flagstep = 1;
            counter = 3;
        } else {
            flagstep = 0;
        }
    }
    if (counter == 1 && flagstep == 1) {
        if (fabsf(2.0f * rp - rp1 - rp2) < b->g_var) {
            flagconst = 1;
            counter = 0;
        } else {
            flagconst = 0;
            counter = 2;
        }
    } else if (counter > 0) {
        counter -= 1;
    }
    rp2 = rp1;
    rp1 = rp;
    if (flagconst) {
        gp = rp;
        b->timesig = bt_get_timesig(b->acf, acflen, (int)gp);
        for (int j = 0; j < laglen; j++) {
            float diff = (float)(j + 1) - gp;
            b->gwv[j] = expf(-0.5f * diff * diff / (b->g_var * b->g_var));
        }
        bp = gp;
        for (int j = 0; j < 2 * laglen; j++) b->phwv[j] = 1.0f;
    } else if (b->timesig > 0) {
        bp = gp;
        if ((float)step > b->lastbeat) {
            for (int j = 0; j < 2 * laglen; j++) {
                float diff = 1.0f + (float)j - (float)step + b->lastbeat;
                b->phwv[j] = expf(-0.5f * diff * diff / (bp / 8.0f));
            }
        } else {
            for (int j = 0; j < 2 * laglen; j++) b->phwv[j] = 1.0f;
        }
    } else {
        bp = rp;
        for (int j = 0; j < 2 * laglen; j++) b->phwv[j] = 1.0f;
    }
    while (bp > 0.f && bp < 25.f) bp *= 2.0f;
    b->counter = counter;
    b->flagstep = flagstep;
    b->gp = gp;
    b->bp = bp;
    b->rp1 = rp1;
    b->rp2 = rp2;
}

/* aubio.rs:966-1092 */
static void bt_do(beattracking *b, const float *dfframe, float *output /*step*/) {
    const int step = (int)b->step, laglen = BT_LAGLEN, winlen = BT_WINLEN;
    int numelem = b->timesig == 0 ? 4 : (int)b->timesig;
    for (int i = 0; i < winlen; i++) b->dfrev[i] = dfframe[i] * b->dfwv[i];
    for (int j = 0; j < winlen / 2; j++) swapf(&b->dfrev[j], &b->dfrev[winlen - 1 - j]);
    for (int i = 0; i < winlen; i++) { /* vec_autocorr, aubio.rs:819-828 */
        float tmp = 0.f;
        for (int j = i; j < winlen; j++) tmp += dfframe[j - i] * dfframe[j];
        b->acf[i] = tmp / (float)(winlen - i);
    }
    for (int i = 0; i < laglen; i++) b->acfout[i] = 0.f;
    for (int i = 1; i < laglen - 1; i++)
        for (int a = 1; a <= numelem; a++)
            for (int bb = 1; bb < 2 * a; bb++)
                if (i * a + bb - 1 < winlen)
                    b->acfout[i] += b->acf[i * a + bb - 1] / (2.0f * (float)a - 1.0f);
    for (int i = 0; i < laglen; i++) b->acfout[i] *= b->rwv[i];
    int maxindex = vec_max_elem(b->acfout, laglen);
    if (maxindex > 0 && maxindex < laglen - 1)
        b->rp = vec_quadratic_peak_pos(b->acfout, laglen, maxindex);
    else
        b->rp = (float)b->rayparam;
    bt_checkstate(b);
    float bp = b->bp;
    if (bp == 0.f) {
        for (int i = 0; i < step; i++) output[i] = 0.f;
        return;
    }
    int kmax = (int)floorf((float)winlen / bp);
    for (int i = 0; i < winlen; i++) b->phout[i] = 0.f;
    for (int i = 0; (float)i < bp && i < winlen; i++)
        for (int k = 0; k < kmax; k++) {
            int idx = i + (int)floorf(bp * (float)k + 0.5f);
            if (idx < winlen) b->phout[i] += b->dfrev[idx];
        }
    for (int i = 0; i < 2 * laglen; i++) b->phout[i] *= b->phwv[i]; /* min(512,256) */
    maxindex = vec_max_elem(b->phout, winlen);
    float phase;
    if (maxindex >= winlen - 1) phase = (float)step - b->lastbeat;
    else phase = vec_quadratic_peak_pos(b->phout, winlen, maxindex);
    phase += 1.0f;
    for (int i = 0; i < step; i++) output[i] = 0.f;
    int i = 1;
    float beat = bp - phase;
    if (((float)step - b->lastbeat - phase) < -0.40f * bp) beat += bp;
    while (beat + bp < 0.f) beat += bp;
    if (beat >= 0.f && i < step) {
        output[i] = beat;
        i++;
    }
    while (beat + bp <= (float)step && i < step) {
        beat += bp;
        output[i] = beat;
        i++;
    }
    b->lastbeat = beat;
    output[0] = (float)i;
}

static float bt_get_bpm(const beattracking *b) { /* aubio.rs:1231-1239 */
    if (b->bp != 0.f) {
        float period_samples = (float)b->hop_size * b->bp;
        float period_s = period_samples / (float)b->samplerate;
        return 60.0f / period_s;
    }
    return 0.f;
}

static int cmp_f32(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

void bo_tempo_norms(const float *x, uint64_t n, uint32_t n_frames, float *norms) {
    (void)n;
    float win[512], buf[1024];
    hanningz(win, 512);
    for (uint32_t t = 0; t < n_frames; t++) {
        pvoc_fft(x, t, 256, win, buf);
        float *g = norms + (size_t)t * 257;
        g[0] = fabsf(buf[0]);
        for (int i = 1; i < 256; i++) {
            float re = buf[2 * i], im = buf[2 * i + 1];
            g[i] = sqrtf(re * re + im * im);
        }
        g[256] = fabsf(buf[2 * 256]);
    }
}

/* temporal.rs:32-85 driving aubio.rs:1284-1450 (Tempo::new / do_). */
float bo_tempo(const float *x, uint64_t n, uint32_t n_frames, uint32_t silence_len, float *flux_out,
               float *thr_out, float *bpms_out, uint32_t *n_bpms_out) {
    (void)n;
    const uint32_t hop = 256;
    /* Tempo::new: winlen = next_pow2((5.8*sr/hop) as usize) = 512, step = 128 */
    const int winlen = BT_WINLEN, step = BT_WINLEN / 4;
    float win[512], buf[1024], grain[257], oldmag[257];
    hanningz(win, 512);
    memset(oldmag, 0, sizeof(oldmag));
    peakpicker pp;
    pp_init(&pp);
    pp.threshold = 0.3f; /* aubio.rs:1347 */
    beattracking *bt = (beattracking *)malloc(sizeof(beattracking));
    bt_init(bt, hop, BO_SAMPLE_RATE);
    float dfframe[BT_WINLEN], out[BT_WINLEN / 4];
    memset(dfframe, 0, sizeof(dfframe));
    memset(out, 0, sizeof(out));
    int blockpos = 0;
    float *bpms = (float *)malloc(sizeof(float) * (n_frames ? n_frames : 1));
    uint32_t nb = 0;
    for (uint32_t t = 0; t < n_frames; t++) {
        /* 1. PVocTempo::do_ (aubio.rs:338-425): 257 correct bins */
        pvoc_fft(x, t, hop, win, buf);
        grain[0] = fabsf(buf[0]);
        for (int i = 1; i < 256; i++) {
            float re = buf[2 * i], im = buf[2 * i + 1];
            grain[i] = sqrtf(re * re + im * im);
        }
        grain[256] = fabsf(buf[2 * 256]);
        /* 2. SpecFlux::do_ (aubio.rs:455-467) */
        float of = 0.f;
        for (int j = 0; j < 257; j++) {
            if (grain[j] > oldmag[j]) of += grain[j] - oldmag[j];
            oldmag[j] = grain[j];
        }
        if (flux_out) flux_out[t] = of;
        /* 3. aubio.rs:1390-1406 */
        if (blockpos == step - 1) {
            bt_do(bt, dfframe, out);
            for (int i = 0; i < winlen - step; i++) dfframe[i] = dfframe[i + step];
            for (int i = winlen - step; i < winlen; i++) dfframe[i] = 0.f;
            blockpos = -1;
        }
        blockpos += 1;
        /* 4-5. aubio.rs:1410-1416 */
        pp_do(&pp, of);
        float thresholded = pp.thresholded;
        if (thr_out) thr_out[t] = thresholded;
        dfframe[winlen - step + blockpos] = thresholded;
        /* 6. aubio.rs:1419-1438 */
        float tempo_out = 0.f;
        int num_beats = (int)out[0];
        for (int i = 1; i < num_beats; i++) {
            float beat_pos = out[i];
            if (blockpos == (int)floorf(beat_pos)) {
                tempo_out = beat_pos - floorf(beat_pos);
                /* is_silence(input, -90): aubio.rs:1258-1276 */
                float lvl = level_lin(x + (uint64_t)t * hop, silence_len);
                if (10.0f * log10f(lvl) < -90.0f) tempo_out = 0.f;
            }
        }
        /* temporal.rs:50-57 */
        if (tempo_out > 0.f) bpms[nb++] = bt_get_bpm(bt);
    }
    if (n_bpms_out) *n_bpms_out = nb;
    if (bpms_out) memcpy(bpms_out, bpms, sizeof(float) * nb);
    float result;
    if (nb == 0) {
        result = -1.f; /* temporal.rs:66-70 */
    } else {
        /* ndarray-stats quantile_mut(0.5, Midpoint) on n32 */
        qsort(bpms, nb, sizeof(float), cmp_f32);
        double pos = (double)(nb - 1) * 0.5;
        uint32_t lo = (uint32_t)floor(pos), hi = (uint32_t)ceil(pos);
        float lower = bpms[lo], higher = bpms[hi];
        float median = lower + (higher - lower) / 2.f;
        result = normalize(median, 0.f, 206.f);
    }
    free(bpms);
    free(bt);
    return result;
}

/* ======================================================================== */
/* chroma.rs                                                                 */
/* ======================================================================== */

static int cmp_f64(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* chroma.rs:269-331 */
uint64_t bo_pip_track(const double *S, uint32_t frames, uint32_t n_fft, double *pitches,
                      double *mags) {
    const double sr = (double)BO_SAMPLE_RATE;
    const double fmin = 150.0, fmax = 4000.0 < sr / 2.0 ? 4000.0 : sr / 2.0;
    const uint32_t bins = 1 + n_fft / 2;
    /* Array::linspace(0, sr/2, bins) */
    const double stepf = (sr / 2.) / (double)(bins - 1);
    int beginning = -1, end = -1;
    for (uint32_t i = 0; i < bins; i++) {
        double f = 0. + stepf * (double)i;
        if (fmin <= f && f < fmax) {
            if (beginning < 0) beginning = (int)i;
            end = (int)i;
        }
    }
    if (beginning < 0) return 0;
    double *ref = (double *)malloc(sizeof(double) * frames);
    for (uint32_t j = 0; j < frames; j++) {
        const double *col = S + (size_t)j * bins;
        double mx = col[0];
        for (uint32_t k = 0; k < bins; k++) mx = (mx > col[k]) ? mx : col[k];
        ref[j] = 0.1 * mx;
    }
    uint64_t cnt = 0;
    int rows = end - 3 - beginning;
    /* The reference's Zip walks its [bins x frames] view, which is the transpose of a frame-major
     * buffer (utils.rs:63), along memory order; visiting frame by frame here does the same.  The
     * order of the emitted peaks is irrelevant downstream (median + histogram). */
    for (uint32_t j = 0; j < frames; j++)
        for (int i = 0; i < rows; i++) {
            const double *col = S + (size_t)j * bins;
            double before = col[beginning + i], elem = col[beginning + 1 + i],
                   after = col[beginning + 2 + i];
            if (elem > ref[j] && after <= elem && before < elem) {
                double avg = 0.5 * (after - before);
                double shift = 2. * elem - after - before;
                if (fabs(shift) < 2.2250738585072014e-308) shift += 1.;
                shift = avg / shift;
                pitches[cnt] = ((double)(i + beginning + 1) + shift) * sr / (double)n_fft;
                mags[cnt] = elem + 0.5 * avg * shift;
                cnt++;
            }
        }
    free(ref);
    return cnt;
}

/* utils.rs:119-129 hz_to_octs_inplace */
void bo_hz_to_octs(double *f, uint64_t n, double tuning, uint32_t bins_per_octave) {
    const double a440 = 440.0 * pow(2.0, tuning / (double)bins_per_octave);
    for (uint64_t i = 0; i < n; i++) f[i] = log2(f[i] / (a440 / 16.));
}

/* chroma.rs:334-359 (hz_to_octs_inplace with tuning 0, 12 bins) */
double bo_pitch_tuning(double *f, uint64_t n, double resolution) {
    if (n == 0) return 0.0;
    uint32_t nb = (uint32_t)((0.5 - -0.5) / resolution);
    uint64_t *counts = (uint64_t *)calloc(nb, sizeof(uint64_t));
    bo_hz_to_octs(f, n, 0.0, 12);
    for (uint64_t i = 0; i < n; i++) {
        double v = f[i];
        v = fmod(12.0 * v, 1.0);
        if (v >= 0.5) v -= 1.;
        f[i] = v;
        uint64_t idx = (uint64_t)((v - -0.5) / resolution);
        if (idx >= nb) idx = nb - 1; /* the reference would panic here */
        counts[idx]++;
    }
    uint32_t best = 0;
    for (uint32_t i = 1; i < nb; i++)
        if (counts[i] > counts[best]) best = i; /* argmax: first maximum */
    free(counts);
    return (-50. + (100. * resolution * (double)best)) / 100.;
}

/* chroma.rs:361-391 */
double bo_estimate_tuning(const double *S, uint32_t frames, uint32_t n_fft) {
    size_t cap = (size_t)frames * (n_fft / 2 + 1);
    double *pitch = (double *)malloc(sizeof(double) * cap);
    double *mag = (double *)malloc(sizeof(double) * cap);
    uint64_t cnt = bo_pip_track(S, frames, n_fft, pitch, mag);
    double tuning = 0.;
    if (cnt > 0) {
        /* filter p > 0 */
        uint64_t m = 0;
        for (uint64_t i = 0; i < cnt; i++)
            if (pitch[i] > 0.) {
                pitch[m] = pitch[i];
                mag[m] = mag[i];
                m++;
            }
        double *sorted = (double *)malloc(sizeof(double) * (m ? m : 1));
        memcpy(sorted, mag, sizeof(double) * m);
        qsort(sorted, m, sizeof(double), cmp_f64);
        double pos = (double)(m - 1) * 0.5;
        uint64_t lo = (uint64_t)floor(pos), hi = (uint64_t)ceil(pos);
        double thr = sorted[lo] + (sorted[hi] - sorted[lo]) / 2.; /* Midpoint */
        free(sorted);
        uint64_t k = 0;
        for (uint64_t i = 0; i < m; i++)
            if (mag[i] >= thr) pitch[k++] = pitch[i];
        tuning = bo_pitch_tuning(pitch, k, 0.01);
    }
    free(pitch);
    free(mag);
    return tuning;
}

/* chroma.rs:197-267 */
void bo_chroma_filter(uint32_t n_fft, double tuning, double *out) {
    const double ctroct = 5.0, octwidth = 2., nc = 12.0, nc2 = 6.0; /* round(12/2) */
    const uint32_t len = n_fft + 1, keep = 1 + n_fft / 2;
    double *fb = (double *)malloc(sizeof(double) * len);
    double *bw = (double *)malloc(sizeof(double) * len);
    double *wts = (double *)malloc(sizeof(double) * 12 * len);
    const double stepf = (double)BO_SAMPLE_RATE / (double)(len - 1);
    const double a440 = 440.0 * pow(2.0, tuning / 12.0);
    for (uint32_t i = 0; i < len; i++) {
        double f = 0. + stepf * (double)i;
        f /= a440 / 16.;
        fb[i] = log2(f) * nc;
    }
    fb[0] = fb[1] - 1.5 * nc;
    for (uint32_t i = 0; i + 1 < len; i++) {
        double d = fb[i + 1] - fb[i];
        bw[i] = d <= 1. ? 1. : d;
    }
    bw[len - 1] = 1.;
    for (uint32_t r = 0; r < 12; r++)
        for (uint32_t i = 0; i < len; i++) {
            double d = -(double)r + fb[i];
            d = fmod(d + nc2 + 10. * nc, nc) - nc2;
            d = d / bw[i];
            wts[r * len + i] = exp(-0.5 * (2. * d) * (2. * d));
        }
    for (uint32_t i = 0; i < len; i++) {
        double s = 0.;
        for (uint32_t r = 0; r < 12; r++) s += wts[r * len + i] * wts[r * len + i];
        s = sqrt(s);
        if (s < 2.2250738585072014e-308) s = 1.;
        double g = (fb[i] / nc - ctroct) / octwidth;
        g = exp(-0.5 * (g * g));
        for (uint32_t r = 0; r < 12; r++) wts[r * len + i] = wts[r * len + i] / s * g;
    }
    /* np.roll(-3, axis 0): b[r] = wts[(r+3)%12]; keep first 1+n_fft/2 columns */
    for (uint32_t r = 0; r < 12; r++)
        memcpy(out + (size_t)r * keep, wts + (size_t)((r + 3) % 12) * len, sizeof(double) * keep);
    free(fb);
    free(bw);
    free(wts);
}

/* chroma.rs:393-412 */
void bo_chroma_stft(double *S, uint32_t frames, uint32_t n_fft, double tuning, double *chroma) {
    const uint32_t bins = 1 + n_fft / 2;
    for (size_t i = 0; i < (size_t)frames * bins; i++) S[i] = S[i] * S[i];
    double *W = (double *)malloc(sizeof(double) * 12 * bins);
    bo_chroma_filter(n_fft, tuning, W);
    for (uint32_t j = 0; j < frames; j++) {
        const double *col = S + (size_t)j * bins;
        double raw[12], sum = 0.;
        for (uint32_t r = 0; r < 12; r++) {
            const double *w = W + (size_t)r * bins;
            double acc = 0.;
            for (uint32_t k = 0; k < bins; k++) acc += w[k] * col[k];
            raw[r] = acc;
            sum += fabs(acc);
        }
        if (sum < 2.2250738585072014e-308) sum = 1.;
        for (uint32_t r = 0; r < 12; r++) chroma[(size_t)r * frames + j] = raw[r] / sum;
    }
    free(W);
}

/* chroma.rs:177-188 (column-wise L1) */
void bo_normalize_feature_sequence(const double *in, uint32_t rows, uint32_t cols, double *out) {
    for (uint32_t j = 0; j < cols; j++) {
        double sum = 0.;
        for (uint32_t r = 0; r < rows; r++) sum += fabs(in[(size_t)r * cols + j]);
        if (sum < 0.0001) sum = 1.;
        for (uint32_t r = 0; r < rows; r++) out[(size_t)r * cols + j] = in[(size_t)r * cols + j] / sum;
    }
}

/* chroma.rs:139-152, [12 offsets][10 templates] */
static const int TEMPLATES[12][10] = {
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, {1, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 1, 0, 0, 0, 0, 1, 1, 0}, {0, 0, 0, 1, 0, 0, 1, 0, 0, 1}, {0, 0, 0, 0, 1, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 1, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 0, 1, 1, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 1},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
};

/* chroma.rs:157-175 */
void bo_extract_interval_features(const double *chroma, uint32_t frames, double *out) {
    for (uint32_t t = 0; t < 10; t++)
        for (uint32_t j = 0; j < frames; j++) out[(size_t)t * frames + j] = 0.;
    for (uint32_t t = 0; t < 10; t++)
        for (uint32_t shift = 0; shift < 12; shift++) {
            int rolled[12];
            for (uint32_t p = 0; p < 12; p++) rolled[(p + shift) % 12] = TEMPLATES[p][t]; /* rotate_right */
            for (uint32_t j = 0; j < frames; j++) {
                double prod = 1.;
                for (uint32_t p = 0; p < 12; p++)
                    prod *= rolled[p] ? chroma[(size_t)p * frames + j] : 1.; /* f.powi(s) */
                out[(size_t)t * frames + j] += prod;
            }
        }
}

/* chroma.rs:137-155 */
int bo_chroma_interval_features(const double *chroma, uint32_t frames, double out[10]) {
    if (frames == 0) return 1;
    double *e = (double *)malloc(sizeof(double) * 12 * frames);
    double *nrm = (double *)malloc(sizeof(double) * 12 * frames);
    double *feat = (double *)malloc(sizeof(double) * 10 * frames);
    for (size_t i = 0; i < (size_t)12 * frames; i++) e[i] = exp(chroma[i] * 15.);
    bo_normalize_feature_sequence(e, 12, frames, nrm);
    bo_extract_interval_features(nrm, frames, feat);
    for (uint32_t t = 0; t < 10; t++) {
        double s = 0.;
        for (uint32_t j = 0; j < frames; j++) s += feat[(size_t)t * frames + j];
        out[t] = s / (double)frames;
    }
    free(e);
    free(nrm);
    free(feat);
    return 0;
}

/* chroma.rs:97-126 (v2) and :128-132 (v1) */
void bo_chroma_values(const double f_in[10], int version, float *out) {
    if (version == 1) {
        for (int i = 0; i < 10; i++) out[i] = 2.f * ((float)f_in[i] - 0.f) / (0.12f - 0.f) - 1.f;
        return;
    }
    double f[10];
    memcpy(f, f_in, sizeof(f));
    double n1 = 0., n2 = 0.;
    for (int i = 0; i < 6; i++) n1 += f[i] * f[i];
    for (int i = 6; i < 10; i++) n2 += f[i] * f[i];
    n1 = sqrt(n1);
    n2 = sqrt(n2);
    if (n1 > 0.)
        for (int i = 0; i < 6; i++) f[i] /= n1;
    if (n2 > 0.)
        for (int i = 6; i < 10; i++) f[i] /= n2;
    for (int i = 0; i < 10; i++) out[i] = normalize((float)f[i], 0.f, 1.f);
    float a = 2.f * ((float)n1 - 0.f) / (0.25f - 0.f) - 1.f;
    out[10] = a < 1.f ? a : 1.f;
    float b = 2.f * ((float)n2 - 0.f) / (0.025f - 0.f) - 1.f;
    out[11] = b < 1.f ? b : 1.f;
    double angle = atan2(20. * n2, n1 + 1e-12);
    out[12] = 2.f * ((float)angle - 0.f) / (1.57079637050628662f - 0.f) - 1.f;
}

/* ChromaDesc::do_ (chroma.rs:73-86) + get_values on one whole-song call */
int bo_chroma(const float *x, uint64_t n, int version, float *out, double *tuning_out,
              double *chroma_out) {
    const uint32_t win = 8192, hop = 2205, bins = 4097;
    uint32_t frames = bo_stft_num_frames(n, hop);
    double *S = (double *)malloc(sizeof(double) * (size_t)frames * bins);
    bo_stft(x, n, win, hop, S);
    double tuning = bo_estimate_tuning(S, frames, win);
    if (tuning_out) *tuning_out = tuning;
    double *chroma = (double *)malloc(sizeof(double) * 12 * (size_t)frames);
    bo_chroma_stft(S, frames, win, tuning, chroma);
    if (chroma_out) memcpy(chroma_out, chroma, sizeof(double) * 12 * (size_t)frames);
    double f[10];
    int rc = bo_chroma_interval_features(chroma, frames, f);
    if (rc == 0) bo_chroma_values(f, version, out);
    free(chroma);
    free(S);
    return rc;
}

/* ======================================================================== */
/* song/mod.rs: Song::analyze_with_options                                   */
/* ======================================================================== */
int bo_analyze(const float *pcm, uint64_t n, int version, float *out) {
    if (n < 8192) return BO_TOO_SHORT; /* song/mod.rs:417-430 */
    /* windows(512).step_by(hop): (n-512)/hop + 1 */
    uint32_t n_t = (uint32_t)((n - 512) / 256 + 1);
    uint32_t n_s = (uint32_t)((n - 512) / 128 + 1);
    float tempo = bo_tempo(pcm, n, n_t, 512, NULL, NULL, NULL, NULL);
    float *c = (float *)malloc(sizeof(float) * 3 * (size_t)n_s);
    bo_timbral_frames(pcm, n, n_s, c, c + n_s, c + 2 * (size_t)n_s, NULL);
    float cent[2], roll[2], flat[2], loud[2];
    bo_summarise(c, n_s, 0, cent);
    bo_summarise(c + n_s, n_s, 0, roll);
    bo_summarise(c + 2 * (size_t)n_s, n_s, 1, flat);
    free(c);
    float zcr = bo_zcr(pcm, n);
    bo_loudness(pcm, n, 0, loud);
    /* song/mod.rs:493-498 */
    out[0] = tempo;
    out[1] = zcr;
    out[2] = cent[0];
    out[3] = cent[1];
    out[4] = roll[0];
    out[5] = roll[1];
    out[6] = flat[0];
    out[7] = flat[1];
    out[8] = loud[0];
    out[9] = loud[1];
    return bo_chroma(pcm, n, version, out + 10, NULL, NULL) ? 2 : BO_OK;
}

typedef struct {
    const float *const *pcm;
    const uint64_t *n;
    uint32_t lo, hi;
    int version;
    float *out;
    int32_t *status;
} worker_arg;

static void *worker(void *p) {
    worker_arg *a = (worker_arg *)p;
    int dim = a->version == 1 ? 20 : 23;
    for (uint32_t i = a->lo; i < a->hi; i++) {
        int rc = bo_analyze(a->pcm[i], a->n[i], a->version, a->out + (size_t)i * dim);
        if (a->status) a->status[i] = rc;
    }
    return NULL;
}

/* song/decoder.rs:278-332: paths.chunks(len/cores) -> one thread per chunk */
int bo_analyze_batch(const float *const *pcm, const uint64_t *n, uint32_t n_songs, int version,
                     float *out, int32_t *status, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > n_songs) n_threads = (int)(n_songs ? n_songs : 1);
    /* keep the per-song 59 MB spectra inside the per-thread arenas instead of mmap/munmap-ing them
     * for every song: with ~100 worker threads the page-fault / mmap-lock storm otherwise dominates */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    get_plan(512);
    get_plan(8192);
    uint32_t chunk = (n_songs + n_threads - 1) / n_threads;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    worker_arg *args = (worker_arg *)malloc(sizeof(worker_arg) * n_threads);
    int started = 0;
    for (int t = 0; t < n_threads; t++) {
        uint32_t lo = t * chunk, hi = lo + chunk > n_songs ? n_songs : lo + chunk;
        if (lo >= hi) break;
        args[t] = (worker_arg){pcm, n, lo, hi, version, out, status};
        pthread_create(&th[t], NULL, worker, &args[t]);
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    free(args);
    return 0;
}

/* ======================================================================== */
/* playlist.rs / lib.rs                                                      */
/* ======================================================================== */

/* ndarray 0.17 numeric_util::unrolled_dot: what Array1::dot runs for
 * contiguous f32 (called at playlist.rs:70, :77, :141). */
static float unrolled_dot(const float *xs, const float *ys, uint32_t len) {
    float sum = 0.f, p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t i = 0;
    for (; i + 8 <= len; i += 8)
        for (int k = 0; k < 8; k++) p[k] = p[k] + xs[i + k] * ys[i + k];
    sum = sum + (p[0] + p[4]);
    sum = sum + (p[1] + p[5]);
    sum = sum + (p[2] + p[6]);
    sum = sum + (p[3] + p[7]);
    for (; i < len; i++) sum = sum + xs[i] * ys[i];
    return sum;
}

/* playlist.rs:140-142: (a-b).dot(m).dot(&(a-b)).sqrt() in f32 */
float bo_mahalanobis_distance(const float *a, const float *b, const float *m, uint32_t dim) {
    float d[64], t[64];
    for (uint32_t i = 0; i < dim; i++) d[i] = a[i] - b[i];
    for (uint32_t j = 0; j < dim; j++) {
        float s = 0.f;
        for (uint32_t i = 0; i < dim; i++) s += d[i] * m[(size_t)i * dim + j];
        t[j] = s;
    }
    return sqrtf(unrolled_dot(t, d, dim));
}

float bo_euclidean_distance(const float *a, const float *b, uint32_t dim) { /* playlist.rs:65-71 */
    float d[64];
    for (uint32_t i = 0; i < dim; i++) d[i] = a[i] - b[i];
    return sqrtf(unrolled_dot(d, d, dim)); /* m = eye: (a-b).dot(I) == a-b exactly */
}

float bo_cosine_distance(const float *a, const float *b, uint32_t dim) { /* playlist.rs:76-79 */
    float ab = unrolled_dot(a, b, dim), aa = unrolled_dot(a, a, dim), bb = unrolled_dot(b, b, dim);
    return 1.f - ab / (sqrtf(aa) * sqrtf(bb));
}

/* lib.rs:168-173, :209-234 */
void bo_feature_weights(int version, float *m) {
    uint32_t dim = version == 1 ? 20 : 23;
    memset(m, 0, sizeof(float) * dim * dim);
    for (uint32_t i = 0; i < dim; i++) {
        float w = 1.f;
        if (version != 1) {
            if (i == 0) w = 0.25f;
            else if (i >= 10) w = 3.f / 13.f;
        }
        m[(size_t)i * dim + i] = w;
    }
}

float bo_default_distance(const float *a, const float *b, int version) {
    float m[23 * 23];
    bo_feature_weights(version, m);
    return bo_mahalanobis_distance(a, b, m, version == 1 ? 20 : 23);
}

static float dist_m(const float *a, const float *b, uint32_t dim, const float *m) {
    return m ? bo_mahalanobis_distance(a, b, m, dim) : bo_euclidean_distance(a, b, dim);
}

/* playlist.rs:256-270 with FunctionDistanceMetric::distance (:56-58) */
void bo_closest_to_songs(const float *seeds, uint32_t n_seeds, const float *cands, uint32_t n_cands,
                         uint32_t dim, const float *m, uint32_t *order, float *keys_out) {
    float *keys = (float *)malloc(sizeof(float) * (n_cands ? n_cands : 1));
    for (uint32_t j = 0; j < n_cands; j++) {
        float s = 0.f;
        for (uint32_t i = 0; i < n_seeds; i++)
            s += dist_m(seeds + (size_t)i * dim, cands + (size_t)j * dim, dim, m);
        keys[j] = s;
        order[j] = j;
    }
    /* stable insertion-free merge: simple stable sort by key */
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * (n_cands ? n_cands : 1));
    for (uint32_t w = 1; w < n_cands; w *= 2) {
        for (uint32_t lo = 0; lo < n_cands; lo += 2 * w) {
            uint32_t mid = lo + w < n_cands ? lo + w : n_cands;
            uint32_t hi = lo + 2 * w < n_cands ? lo + 2 * w : n_cands;
            uint32_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (keys[order[j]] < keys[order[i]]) ? order[j++] : order[i++];
            while (i < mid) tmp[k++] = order[i++];
            while (j < hi) tmp[k++] = order[j++];
        }
        memcpy(order, tmp, sizeof(uint32_t) * n_cands);
    }
    if (keys_out) memcpy(keys_out, keys, sizeof(float) * n_cands);
    free(tmp);
    free(keys);
}

/* playlist.rs:272-326: greedy nearest-neighbour chain; argmin = first minimum */
void bo_song_to_song(const float *seeds, uint32_t n_seeds, const float *cands, uint32_t n_cands,
                     uint32_t dim, const float *m, uint32_t *order) {
    uint32_t *pool = (uint32_t *)malloc(sizeof(uint32_t) * (n_cands ? n_cands : 1));
    for (uint32_t j = 0; j < n_cands; j++) pool[j] = j;
    uint32_t pool_n = n_cands;
    const float *vec = seeds;
    uint32_t nvec = n_seeds;
    for (uint32_t out = 0; out < n_cands; out++) {
        uint32_t best = 0;
        float bestd = 0.f;
        for (uint32_t j = 0; j < pool_n; j++) {
            float s = 0.f;
            for (uint32_t i = 0; i < nvec; i++)
                s += dist_m(vec + (size_t)i * dim, cands + (size_t)pool[j] * dim, dim, m);
            if (j == 0 || s < bestd) {
                bestd = s;
                best = j;
            }
        }
        uint32_t chosen = pool[best];
        order[out] = chosen;
        memmove(pool + best, pool + best + 1, sizeof(uint32_t) * (pool_n - best - 1));
        pool_n--;
        vec = cands + (size_t)chosen * dim;
        nvec = 1;
    }
    free(pool);
}

/* ---- decode-side feed: sample format + down-mix for sources already at 22 050 Hz ----------
 * What the reference's decoders do between the codec's output and PreAnalyzedSong::sample_array
 * (src/song/decoder.rs:64) when no rate change is needed:
 *   s16 / s32 -> f32: swresample's x * (1.0f / (1 << 15)) and x * (1.0f / (1U << 31)) behind
 *     src/song/decoder/ffmpeg.rs:36-109 (symphonia's sample conversion divides by the same powers
 *     of two and rounds identically);
 *   stereo: swresample's FL+FR -> FC matrix for float output, c*L + c*R with c = (float)M_SQRT1_2 --
 *     "averaging the channels and multiplying by the square root of 2 ... recovers the exact
 *     behavior of ffmpeg", src/song/decoder/symphonia.rs:260-262, :280-287;
 *   more than two channels: `chunk.iter().sum::<f32>() / num_channels as f32`, :289-299.
 * Pinned by the adler32 the reference asserts for data/s16_stereo_22_5kHz.flac
 * (src/song/decoder/ffmpeg.rs:447-452); both channels of that file are identical, so the hash
 * pins the two constants but not the rounding order of the two products.
 * fmt: 1 = s16, 2 = s32, 3 = f32. */
static inline float bo_pcm_sample(const void *in, int fmt, size_t i) {
    if (fmt == 1) return (float)((const int16_t *)in)[i] * (1.0f / 32768.0f);
    if (fmt == 2) return (float)((const int32_t *)in)[i] * (1.0f / 2147483648.0f);
    return ((const float *)in)[i];
}

int bo_pcm_to_mono(const void *pcm, uint64_t n_frames, int fmt, uint32_t channels, float *out) {
    if (fmt < 1 || fmt > 3 || channels == 0) return -1;
    const float c = (float)0.70710678118654752440; /* M_SQRT1_2 */
    for (uint64_t f = 0; f < n_frames; f++) {
        const size_t base = (size_t)f * channels;
        if (channels == 1) {
            out[f] = bo_pcm_sample(pcm, fmt, base);
        } else if (channels == 2) {
            const float l = bo_pcm_sample(pcm, fmt, base) * c;
            const float r = bo_pcm_sample(pcm, fmt, base + 1) * c;
            out[f] = l + r;
        } else {
            float s = 0.f;
            for (uint32_t ch = 0; ch < channels; ch++) s += bo_pcm_sample(pcm, fmt, base + ch);
            out[f] = s / (float)channels;
        }
    }
    return 0;
}

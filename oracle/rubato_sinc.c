/* oracle/rubato_sinc.c — TEST INFRASTRUCTURE ONLY (see sk_oracle.h).
 *
 * Restatement of rubato 0.14.1 `SincFixedIn<f32>` as soundkit calls it
 * (soundkit/src/audio_pipeline.rs:474-491, soundkit-decoder/src/lib.rs:1939-1954):
 *   SincInterpolationParameters{sinc_len 256, f_cutoff 0.95, Linear, oversampling 256, BlackmanHarris2}
 * rubato is a crates.io dependency ("0.14.1", soundkit/Cargo.toml:15) that is NOT under
 * /root/reference; this file restates its published algorithm (sinc.rs make_sincs, windows.rs
 * blackman_harris, asynchro_sinc.rs SincFixedIn::process_into_buffer + get_nearest_times_2,
 * interpolator_avx.rs get_sinc_interpolated, interpolation.rs interp_lin).
 * SAMPLE VALUES: PARITY UNPINNED (only the output-length relation of
 * soundkit-decoder/src/lib.rs:5187-5238 is pinned by the reference).
 */
#define _GNU_SOURCE
#include "sk_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define SINC_LEN 256
#define FACTOR 256

struct sko_resampler {
    size_t channels, chunk;
    double ratio;        /* out/in */
    double last_index;   /* -(sinc_len/2) initially */
    float* sincs;        /* [FACTOR][SINC_LEN] */
    float** buffer;      /* [channels][chunk + 2*SINC_LEN] */
    /* scratch for the index walk */
    int32_t* w_i0; uint8_t* w_s0; float* w_frac; size_t w_cap;
};

/* rubato sinc.rs: make_sincs::<f32>(npoints=256, factor=256, f_cutoff, BlackmanHarris2),
 * windows.rs: blackman_harris (denominator N, not N-1), squared for the "2" variant.
 * All arithmetic in f32; sin/cos are libm sinf/cosf (what Rust's f32::sin/cos call). */
static void make_sincs(float f_cutoff, float* sincs) {
    const size_t tot = (size_t)SINC_LEN * FACTOR;
    float* y = (float*)malloc(tot * sizeof(float));
    const float PI = 3.14159265358979323846264338327950288f;
    const float pi2 = 2.0f * PI, pi4 = 4.0f * PI, pi6 = 6.0f * PI;
    const float np = (float)tot;
    const float a = 0.35875f, b = 0.48829f, c = 0.14128f, d = 0.01168f;
    float sum = 0.0f;
    for (size_t x = 0; x < tot; x++) {
        float xf = (float)x;
        float w = a - b * cosf(pi2 * xf / np) + c * cosf(pi4 * xf / np) - d * cosf(pi6 * xf / np);
        w = w * w;
        float v = (xf - (float)(tot / 2)) * f_cutoff / (float)FACTOR;
        float s = (v == 0.0f) ? 1.0f : sinf(v * PI) / (v * PI);
        float val = w * s;
        sum += val;
        y[x] = val;
    }
    sum /= (float)FACTOR;
    for (size_t p = 0; p < SINC_LEN; p++)
        for (size_t n = 0; n < FACTOR; n++)
            sincs[(FACTOR - n - 1) * SINC_LEN + p] = y[FACTOR * p + n] / sum;
    free(y);
}

sko_resampler* sko_resampler_new(uint32_t in_rate, uint32_t out_rate, size_t chunk, size_t channels) {
    sko_resampler* r = (sko_resampler*)calloc(1, sizeof(*r));
    r->channels = channels; r->chunk = chunk;
    r->ratio = (double)out_rate / (double)in_rate;
    /* make_interpolator: sinc_len already a multiple of 8; cutoff scaled when downsampling */
    float fc = r->ratio >= 1.0 ? 0.95f : 0.95f * (float)r->ratio;
    r->sincs = (float*)aligned_alloc(64, (size_t)FACTOR * SINC_LEN * sizeof(float));
    make_sincs(fc, r->sincs);
    r->buffer = (float**)calloc(channels ? channels : 1, sizeof(float*));
    for (size_t c = 0; c < channels; c++) r->buffer[c] = (float*)calloc(chunk + 2 * SINC_LEN + 8, sizeof(float));
    r->last_index = -(double)(SINC_LEN / 2);
    return r;
}

void sko_resampler_free(sko_resampler* r) {
    if (!r) return;
    for (size_t c = 0; c < r->channels; c++) free(r->buffer[c]);
    free(r->buffer); free(r->sincs); free(r->w_i0); free(r->w_s0); free(r->w_frac); free(r);
}

size_t sko_resampler_out_max(const sko_resampler* r) { return (size_t)((double)r->chunk * r->ratio * 2.0 + 10.0); }
const float* sko_resampler_table(const sko_resampler* r) { return r->sincs; }
double sko_resampler_last_index(const sko_resampler* r) { return r->last_index; }

/* the f64 index walk of process_into_buffer (Linear arm) + get_nearest_times_2.
 * t_ratio_increment is exactly 0.0 for a fixed ratio, so `t_ratio += 0.0` is a no-op. */
static size_t walk(double last_index, double ratio, size_t chunk, int32_t* i0, uint8_t* s0, float* frac,
                   size_t cap, double* idx_out) {
    double t_ratio = 1.0 / ratio;
    long end_idx = (long)chunk - (SINC_LEN + 1) - (long)ceil(t_ratio);
    double idx = last_index;
    size_t n = 0;
    while (idx < (double)end_idx) {
        idx += t_ratio;
        if (n < cap) {
            double fl = floor(idx);
            if (i0) i0[n] = (int32_t)fl;
            if (s0) s0[n] = (uint8_t)(long)floor((idx - fl) * (double)FACTOR);
            if (frac) { double f = idx * (double)FACTOR; frac[n] = (float)(f - floor(f)); }
        }
        n++;
    }
    if (idx_out) *idx_out = idx;
    return n;
}

size_t sko_resampler_walk(const sko_resampler* r, int32_t* i0, uint8_t* s0, float* frac, size_t cap) {
    return walk(r->last_index, r->ratio, r->chunk, i0, s0, frac, cap, NULL);
}

/* interpolator_avx.rs get_sinc_interpolated: 8 lane accumulators, lane l takes taps l, l+8, ...
 * with FUSED multiply-add in tap order; then (hi128 + lo128), hadd, hadd =
 * ((a0+a4)+(a1+a5)) + ((a2+a6)+(a3+a7)). */
#if defined(__x86_64__)
__attribute__((target("avx2,fma")))
static float dot_avx(const float* wave, const float* sinc) {
    __m256 acc = _mm256_setzero_ps();
    for (int k = 0; k < SINC_LEN; k += 8)
        acc = _mm256_fmadd_ps(_mm256_loadu_ps(wave + k), _mm256_load_ps(sinc + k), acc);
    __m128 hi = _mm256_extractf128_ps(acc, 1);
    __m128 lo = _mm_add_ps(hi, _mm256_castps256_ps128(acc));
    __m128 t2 = _mm_hadd_ps(lo, lo);
    __m128 t1 = _mm_hadd_ps(t2, t2);
    return _mm_cvtss_f32(t1);
}
#endif
static float dot_scalar(const float* wave, const float* sinc) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < SINC_LEN; k += 8)
        for (int l = 0; l < 8; l++) a[l] = fmaf(wave[k + l], sinc[k + l], a[l]);
    return ((a[0] + a[4]) + (a[1] + a[5])) + ((a[2] + a[6]) + (a[3] + a[7]));
}
static int g_use_avx = -1;
int sko_resampler_uses_avx(void) {
    if (g_use_avx < 0) {
#if defined(__x86_64__)
        g_use_avx = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
#else
        g_use_avx = 0;
#endif
    }
    return g_use_avx;
}
static inline float dot(const float* wave, const float* sinc) {
#if defined(__x86_64__)
    if (g_use_avx) return dot_avx(wave, sinc);
#endif
    return dot_scalar(wave, sinc);
}

typedef struct {
    const sko_resampler* r; float* const* out; size_t n0, n1;
} job_t;

static void* run_job(void* p) {
    job_t* j = (job_t*)p;
    const sko_resampler* r = j->r;
    for (size_t n = j->n0; n < j->n1; n++) {
        long i0 = r->w_i0[n]; int s0 = r->w_s0[n];
        long i1 = i0; int s1 = s0 + 1;
        if (s1 >= FACTOR) { s1 -= FACTOR; i1 += 1; }           /* get_nearest_times_2 */
        float frac = r->w_frac[n];
        for (size_t c = 0; c < r->channels; c++) {
            const float* buf = r->buffer[c];
            float p0 = dot(buf + (i0 + 2 * SINC_LEN), r->sincs + (size_t)s0 * SINC_LEN);
            float p1 = dot(buf + (i1 + 2 * SINC_LEN), r->sincs + (size_t)s1 * SINC_LEN);
            j->out[c][n] = p0 + frac * (p1 - p0);              /* interp_lin */
        }
    }
    return NULL;
}

size_t sko_resampler_process(sko_resampler* r, const float* const* in, float* const* out, int threads) {
    sko_resampler_uses_avx();
    const size_t chunk = r->chunk;
    /* history shift: buffer.copy_within(chunk .. chunk + 2*sinc_len, 0); then copy input */
    for (size_t c = 0; c < r->channels; c++) {
        memmove(r->buffer[c], r->buffer[c] + chunk, 2 * SINC_LEN * sizeof(float));
        memcpy(r->buffer[c] + 2 * SINC_LEN, in[c], chunk * sizeof(float));
    }
    size_t cap = sko_resampler_out_max(r);
    if (r->w_cap < cap) {
        free(r->w_i0); free(r->w_s0); free(r->w_frac);
        r->w_i0 = (int32_t*)malloc(cap * sizeof(int32_t));
        r->w_s0 = (uint8_t*)malloc(cap);
        r->w_frac = (float*)malloc(cap * sizeof(float));
        r->w_cap = cap;
    }
    double idx;
    size_t n = walk(r->last_index, r->ratio, chunk, r->w_i0, r->w_s0, r->w_frac, cap, &idx);
    if (threads <= 1 || n < 4096) {
        job_t j = {r, out, 0, n};
        run_job(&j);
    } else {
        if (threads > 64) threads = 64;
        pthread_t th[64]; job_t jobs[64];
        for (int t = 0; t < threads; t++) {
            jobs[t] = (job_t){r, out, n * (size_t)t / (size_t)threads, n * (size_t)(t + 1) / (size_t)threads};
            pthread_create(&th[t], NULL, run_job, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    r->last_index = idx - (double)chunk;
    return n;
}

size_t sko_resampler_process_partial(sko_resampler* r, const float* const* in_or_null, size_t frames,
                                     float* const* out, int threads) {
    float** padded = (float**)malloc(r->channels * sizeof(float*));
    if (frames > r->chunk) frames = r->chunk;
    for (size_t c = 0; c < r->channels; c++) {
        padded[c] = (float*)calloc(r->chunk ? r->chunk : 1, sizeof(float));
        if (in_or_null && frames) memcpy(padded[c], in_or_null[c], frames * sizeof(float));
    }
    size_t n = sko_resampler_process(r, (const float* const*)padded, out, threads);
    for (size_t c = 0; c < r->channels; c++) free(padded[c]);
    free(padded);
    return n;
}

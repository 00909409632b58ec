/* oracle/pipeline.c — TEST INFRASTRUCTURE ONLY (see sk_oracle.h).
 * Restates soundkit/src/audio_pipeline.rs:438-493 (downsample_audio) and
 * soundkit-decoder/src/lib.rs:79, 1917-2060 (StreamingResampler), 3261-3290, 3324-3456
 * (the rate-changing arm of apply_output_options with a mono target). */
#include "sk_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define RESAMPLE_CHUNK_SIZE 4096 /* soundkit-decoder/src/lib.rs:79 */

static int rate_ok(uint32_t r) { /* COMMON_SAMPLE_RATES, audio_pipeline.rs:12-13 */
    static const uint32_t ok[9] = {8000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000};
    for (int i = 0; i < 9; i++) if (ok[i] == r) return 1;
    return 0;
}

size_t sko_downsample_audio(const uint8_t* in, size_t bytes, int bits, int fmt, size_t channels,
                            uint32_t in_rate, uint32_t out_rate, float* out, size_t out_cap, int threads) {
    if (channels == 0) return (size_t)-1;                              /* :440-443 */
    if (bits != 16 && bits != 24 && bits != 32) return (size_t)-1;      /* :445-450 */
    if (in_rate == 0 || out_rate == 0) return (size_t)-1;               /* :456-458 */
    if (!rate_ok(in_rate) || !rate_ok(out_rate)) return (size_t)-1;     /* :460-466 */
    size_t bps = (size_t)bits / 8;
    size_t frames_cap = bytes / (bps * channels) + 1;
    float* planar = (float*)malloc(frames_cap * channels * sizeof(float));
    size_t frames = sko_pcm_to_planar_f32(in, bytes, bits, fmt, 0, channels, SKO_MODE_A, planar); /* :468 */
    if (frames == (size_t)-1) { free(planar); return (size_t)-1; }
    /* SincFixedIn::new(out/in, 2.0, params, data[0].len(), data.len()) + one process()  :482-491 */
    sko_resampler* r = sko_resampler_new(in_rate, out_rate, frames, channels);
    size_t omax = sko_resampler_out_max(r);
    float** ip = (float**)malloc(channels * sizeof(float*));
    float** op = (float**)malloc(channels * sizeof(float*));
    float* tmp = (float*)malloc((omax ? omax : 1) * channels * sizeof(float));
    for (size_t c = 0; c < channels; c++) { ip[c] = planar + c * frames; op[c] = tmp + c * omax; }
    size_t n = sko_resampler_process(r, (const float* const*)ip, op, threads);
    if (n > out_cap) n = out_cap;
    for (size_t c = 0; c < channels; c++) memcpy(out + c * out_cap, op[c], n * sizeof(float));
    free(tmp); free(ip); free(op); free(planar);
    sko_resampler_free(r);
    return n;
}

/* ---- StreamingResampler ------------------------------------------------ */
struct sko_stream {
    sko_resampler* r;
    size_t channels;
    uint32_t in_rate, out_rate;
    float** accum;        /* per channel pending input */
    size_t accum_len, accum_cap;
    float** obuf;         /* per channel scratch of out_max */
    size_t omax;
};

sko_stream* sko_stream_new(uint32_t in_rate, uint32_t out_rate, size_t channels) {
    sko_stream* s = (sko_stream*)calloc(1, sizeof(*s));
    s->r = sko_resampler_new(in_rate, out_rate, RESAMPLE_CHUNK_SIZE, channels);
    s->channels = channels; s->in_rate = in_rate; s->out_rate = out_rate;
    s->omax = sko_resampler_out_max(s->r);
    s->accum = (float**)calloc(channels, sizeof(float*));
    s->obuf = (float**)calloc(channels, sizeof(float*));
    for (size_t c = 0; c < channels; c++) s->obuf[c] = (float*)malloc(s->omax * sizeof(float));
    return s;
}
void sko_stream_free(sko_stream* s) {
    if (!s) return;
    for (size_t c = 0; c < s->channels; c++) { free(s->accum[c]); free(s->obuf[c]); }
    free(s->accum); free(s->obuf); sko_resampler_free(s->r); free(s);
}

static int emit(sko_stream* s, size_t n, float* out, size_t out_cap, size_t* out_len) {
    if (*out_len + n > out_cap) return -1;
    for (size_t c = 0; c < s->channels; c++) memcpy(out + c * out_cap + *out_len, s->obuf[c], n * sizeof(float));
    *out_len += n;
    return 0;
}

/* StreamingResampler::process, :1970-2015 (the accum compaction is a memory detail) */
int sko_stream_process(sko_stream* s, const float* planar, size_t frames, float* out, size_t out_cap,
                       size_t* out_len) {
    if (s->accum_len + frames > s->accum_cap) {
        s->accum_cap = (s->accum_len + frames) * 2 + RESAMPLE_CHUNK_SIZE;
        for (size_t c = 0; c < s->channels; c++) s->accum[c] = (float*)realloc(s->accum[c], s->accum_cap * sizeof(float));
    }
    for (size_t c = 0; c < s->channels; c++) memcpy(s->accum[c] + s->accum_len, planar + c * frames, frames * sizeof(float));
    s->accum_len += frames;
    size_t start = 0;
    const float* ip[256];
    while (s->accum_len - start >= RESAMPLE_CHUNK_SIZE) {
        for (size_t c = 0; c < s->channels; c++) ip[c] = s->accum[c] + start;
        size_t n = sko_resampler_process(s->r, ip, s->obuf, 1);
        start += RESAMPLE_CHUNK_SIZE;
        if (n && emit(s, n, out, out_cap, out_len)) return -1;
    }
    if (start) {
        for (size_t c = 0; c < s->channels; c++) memmove(s->accum[c], s->accum[c] + start, (s->accum_len - start) * sizeof(float));
        s->accum_len -= start;
    }
    return 0;
}

/* StreamingResampler::flush, :2017-2059 */
int sko_stream_flush(sko_stream* s, float* out, size_t out_cap, size_t* out_len) {
    size_t remaining = s->accum_len;
    const float* ip[256];
    if (remaining > 0) {
        size_t padded = RESAMPLE_CHUNK_SIZE - remaining;
        for (size_t c = 0; c < s->channels; c++) ip[c] = s->accum[c];
        size_t n = sko_resampler_process_partial(s->r, ip, remaining, s->obuf, 1);
        if (padded > 0) {
            size_t trim = (size_t)round(((double)padded * (double)s->out_rate) / (double)s->in_rate);
            n = n > trim ? n - trim : 0;
        }
        s->accum_len = 0;
        if (n && emit(s, n, out, out_cap, out_len)) return -1;
    } else {
        size_t n = sko_resampler_process_partial(s->r, NULL, 0, s->obuf, 1);
        if (n && emit(s, n, out, out_cap, out_len)) return -1;
    }
    return 0;
}

/* apply_output_options rate-change arm with output_channels = 1 (lib.rs:3385-3431):
 * audio_data_to_f32_channels (mode B) -> StreamingResampler::process + flush -> per emitted
 * chunk downmix_channels(.., 1).  Downmix is per-sample so chunking does not matter. */
size_t sko_normalize_chain_mono(const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian,
                                size_t channels, uint32_t in_rate, uint32_t out_rate, float* out,
                                size_t out_cap) {
    size_t bps = (size_t)((bits + 7) / 8);
    size_t frames_cap = bytes / (bps * (channels ? channels : 1)) + 1;
    float* planar = (float*)malloc(frames_cap * (channels ? channels : 1) * sizeof(float));
    size_t frames = sko_pcm_to_planar_f32(in, bytes, bits, fmt, big_endian, channels, SKO_MODE_B, planar);
    if (frames == (size_t)-1) { free(planar); return (size_t)-1; }
    sko_stream* s = sko_stream_new(in_rate, out_rate, channels);
    size_t cap = (size_t)((double)(frames + 2 * RESAMPLE_CHUNK_SIZE) * (double)out_rate / (double)in_rate) + 64;
    float* res = (float*)malloc(cap * channels * sizeof(float));
    size_t len = 0;
    int rc = sko_stream_process(s, planar, frames, res, cap, &len);
    if (!rc) rc = sko_stream_flush(s, res, cap, &len);
    size_t n = (size_t)-1;
    if (!rc) {
        /* compact [channels][cap] -> [channels][len] then downmix */
        float* packed = (float*)malloc((len ? len : 1) * channels * sizeof(float));
        for (size_t c = 0; c < channels; c++) memcpy(packed + c * len, res + c * cap, len * sizeof(float));
        n = len < out_cap ? len : out_cap;
        float* mono = (float*)malloc((len ? len : 1) * sizeof(float));
        if (channels > 1) sko_downmix_channels(packed, channels, len, 1, mono);
        else memcpy(mono, packed, len * sizeof(float));
        memcpy(out, mono, n * sizeof(float));
        free(mono); free(packed);
    }
    free(res); free(planar); sko_stream_free(s);
    return n;
}

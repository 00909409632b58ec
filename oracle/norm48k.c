/*
 * norm48k.c — CPU restatement of soundkit's StreamingStereo48kNormalizer (TEST INFRASTRUCTURE ONLY, see sk_oracle.h).
 *
 * Follows soundkit/src/audio_pipeline.rs:133-413: decoder-convention sample conversion (pcm_sample_to_f32, :176-227),
 * mono duplicated / first two channels (:155-166), finite_pcm clamp (:396-402), 2-tap linear interpolation over the
 * whole-stream timeline with `source_position = output_frames as f64 * rate / 48000.0` (:326-347), and the exact
 * rounded-duration tail that repeats the final sample (:366-394).  Pinned by the reference's own unit tests
 * (:715-755, :797-823): see tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "sk_oracle.h"

struct sko_norm48k {
    uint32_t rate; uint8_t channels;
    uint64_t source_frames, output_frames;
    float prev_l, prev_r;
    int has_previous, finished;
};

static float finite_pcm(float s) { /* :396-402 */
    if (!isfinite(s)) return 0.0f;
    return s < -1.0f ? -1.0f : (s > 1.0f ? 1.0f : s);
}

/* pcm_sample_to_f32, :176-227; returns 0 on an unsupported format */
static int sample_to_f32(const uint8_t* b, int bits, int fmt, int be, float* out) {
    if (bits == 16) {
        int16_t v = be ? (int16_t)((b[0] << 8) | b[1]) : (int16_t)((b[1] << 8) | b[0]);
        *out = (float)v / 32768.0f; return 1;
    }
    if (bits == 24) {
        uint32_t u = be ? ((uint32_t)b[0] << 16) | ((uint32_t)b[1] << 8) | b[2] : ((uint32_t)b[2] << 16) | ((uint32_t)b[1] << 8) | b[0];
        int32_t v = (u & 0x800000u) ? (int32_t)(u | 0xff000000u) : (int32_t)u;
        *out = (float)v / 8388608.0f; return 1;
    }
    if (bits == 32) {
        uint32_t u = be ? ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]
                        : ((uint32_t)b[3] << 24) | ((uint32_t)b[2] << 16) | ((uint32_t)b[1] << 8) | b[0];
        if (fmt == SKO_FMT_FLOAT) { memcpy(out, &u, 4); return 1; }
        *out = (float)(int32_t)u / 2147483648.0f; return 1;
    }
    return 0;
}

sko_norm48k* sko_norm48k_new(void) { return (sko_norm48k*)calloc(1, sizeof(sko_norm48k)); }
void sko_norm48k_free(sko_norm48k* n) { free(n); }
uint64_t sko_norm48k_source_frames(const sko_norm48k* n) { return n->source_frames; }
uint64_t sko_norm48k_output_frames(const sko_norm48k* n) { return n->output_frames; }

static float stream_sample(const float* s, size_t len, float previous, uint64_t start, uint64_t frame) { /* :404-413 */
    if (frame < start) return previous;
    uint64_t i = frame - start;
    return i < len ? finite_pcm(s[i]) : previous;
}

/* push, :288-363.  Returns 0 = Ok (n_out frames written, possibly 0 = Ok(None)), negative = the Err arm:
 * -1 finished, -2 invalid geometry, -3 geometry changed, -4 unsupported depth, -5 incomplete frame / no channels */
int sko_norm48k_push(sko_norm48k* n, const uint8_t* data, size_t bytes, int bits, int fmt, int be, int channels, uint32_t rate,
                     float* left, float* right, size_t cap, size_t* n_out) {
    *n_out = 0;
    if (n->finished) return -1;
    if (rate == 0 || channels == 0) return -2;
    if (n->rate == 0) { n->rate = rate; n->channels = (uint8_t)channels; }
    else if (n->rate != rate || n->channels != channels) return -3;
    if (bits != 16 && bits != 24 && bits != 32) return -4;
    const size_t sb = (size_t)bits / 8, fb = sb * (size_t)channels;
    if (bytes == 0 || bytes % fb) return -5;
    const size_t frames = bytes / fb;
    float* l = (float*)malloc(frames * sizeof(float));
    float* r = (float*)malloc(frames * sizeof(float));
    for (size_t f = 0; f < frames; f++) { /* audio_to_stereo_f32, :133-174 */
        float a, b;
        sample_to_f32(data + f * fb, bits, fmt, be, &a);
        if (channels > 1) sample_to_f32(data + f * fb + sb, bits, fmt, be, &b); else b = a;
        l[f] = finite_pcm(a); r[f] = finite_pcm(b);
    }
    const uint64_t start = n->source_frames, end = start + frames;
    const float final_l = l[frames - 1], final_r = r[frames - 1];
    size_t w = 0;
    int rc = 0;
    if (rate == 48000) {
        if (frames > cap) rc = -6;
        else { memcpy(left, l, frames * 4); memcpy(right, r, frames * 4); w = frames; n->output_frames += frames; }
    } else {
        for (;;) {
            const double pos = (double)n->output_frames * (double)rate / 48000.0;
            const uint64_t lower = (uint64_t)floor(pos), upper = lower + 1;
            if (upper >= end) break;
            if (w >= cap) { rc = -6; break; }
            const float fraction = (float)(pos - (double)lower);
            const float ll = stream_sample(l, frames, n->prev_l, start, lower), rl = stream_sample(r, frames, n->prev_r, start, lower);
            const float lu = stream_sample(l, frames, n->prev_l, start, upper), ru = stream_sample(r, frames, n->prev_r, start, upper);
            left[w] = finite_pcm(ll + ((lu - ll) * fraction));
            right[w] = finite_pcm(rl + ((ru - rl) * fraction));
            w++;
            n->output_frames++;
        }
    }
    free(l); free(r);
    if (rc) return rc;
    n->source_frames = end; n->prev_l = final_l; n->prev_r = final_r; n->has_previous = 1;
    *n_out = w;
    return 0;
}

/* finish, :366-394.  0 = Ok, -1 already finished, -7 "decoded source contained no PCM" */
int sko_norm48k_finish(sko_norm48k* n, float* left, float* right, size_t cap, size_t* n_out) {
    *n_out = 0;
    if (n->finished) return -1;
    n->finished = 1;
    if (!n->has_previous || n->source_frames == 0 || n->rate == 0) return -7;
    if (n->rate == 48000) return 0;
    double t = round(((double)n->source_frames * 48000.0) / (double)n->rate);
    if (t < 1.0) t = 1.0;
    const uint64_t target = (uint64_t)t;
    const uint64_t remaining = target > n->output_frames ? target - n->output_frames : 0;
    n->output_frames = target;
    if (remaining > cap) return -6;
    for (uint64_t i = 0; i < remaining; i++) { left[i] = n->prev_l; right[i] = n->prev_r; }
    *n_out = (size_t)remaining;
    return 0;
}

/* oracle/pcm.c — TEST INFRASTRUCTURE ONLY (see sk_oracle.h).
 * Restates soundkit/src/audio_bytes.rs:3-373, soundkit/src/audio_pipeline.rs:17-131 and the
 * conversion helpers of soundkit-decoder/src/lib.rs:1815-1827, 3458-3701.
 * Built with -ffp-contract=off: plain Rust arithmetic never fuses multiply-add. */
#include "sk_oracle.h"
#include <math.h>
#include <string.h>

/* ---- Rust scalar semantics ------------------------------------------- */
/* `x as i32` for f32/f64: round toward zero, saturate, NaN -> 0 */
static int32_t as_i32_f(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
static int32_t as_i32_d(double x) {
    if (x != x) return 0;
    if (x >= 2147483648.0) return INT32_MAX;
    if (x <= -2147483648.0) return INT32_MIN;
    return (int32_t)x;
}
static int16_t as_i16_f(float x) {
    if (x != x) return 0;
    if (x >= 32767.0f) return INT16_MAX;
    if (x <= -32768.0f) return INT16_MIN;
    return (int16_t)x;
}
/* f32::clamp(-1.0, 1.0): NaN stays NaN */
static float clamp1(float x) {
    if (x < -1.0f) return -1.0f;
    if (x > 1.0f) return 1.0f;
    return x;
}
static int is_finite_f(float x) { return (x - x) == 0.0f; }

static int16_t ld_i16le(const uint8_t* p) { return (int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8)); }
static int16_t ld_i16be(const uint8_t* p) { return (int16_t)((uint16_t)p[1] | ((uint16_t)p[0] << 8)); }
static int32_t ld_i32le(const uint8_t* p) {
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}
static int32_t ld_i32be(const uint8_t* p) {
    return (int32_t)((uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24));
}
/* sign-extended 24-bit loads (audio_bytes.rs:40-45, :312-315) */
static int32_t ld_s24le(const uint8_t* p) {
    uint32_t u = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    return (u & 0x800000u) ? (int32_t)(u | 0xFF000000u) : (int32_t)u;
}
static int32_t ld_s24be(const uint8_t* p) {
    uint32_t u = (uint32_t)p[2] | ((uint32_t)p[1] << 8) | ((uint32_t)p[0] << 16);
    return (u & 0x800000u) ? (int32_t)(u | 0xFF000000u) : (int32_t)u;
}
static float bits_f32(int32_t b) { float f; memcpy(&f, &b, 4); return f; }

/* ---- audio_bytes.rs flat conversions --------------------------------- */
size_t sko_i16le_to_f32(const uint8_t* in, size_t bytes, float* out) {            /* :3-15 */
    size_t n = bytes / 2;
    for (size_t i = 0; i < n; i++) out[i] = (float)ld_i16le(in + 2 * i) / 32768.0f;
    return n;
}
void sko_i16_to_i16le(const int16_t* in, size_t n, uint8_t* out) {                /* :17-23 */
    for (size_t i = 0; i < n; i++) { out[2 * i] = (uint8_t)in[i]; out[2 * i + 1] = (uint8_t)((uint16_t)in[i] >> 8); }
}
size_t sko_s16le_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :25-34, :231-238 */
    size_t n = bytes / 2;
    for (size_t i = 0; i < n; i++) out[i] = ld_i16le(in + 2 * i);
    return n;
}
size_t sko_s16be_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :222-229 */
    size_t n = bytes / 2;
    for (size_t i = 0; i < n; i++) out[i] = ld_i16be(in + 2 * i);
    return n;
}
size_t sko_s16le_to_i32(const uint8_t* in, size_t bytes, int32_t* out) {          /* :240-248 */
    size_t n = bytes / 2;
    for (size_t i = 0; i < n; i++) out[i] = (int32_t)ld_i16le(in + 2 * i);
    return n;
}
size_t sko_s24le_to_i32(const uint8_t* in, size_t bytes, int32_t* out) {          /* :36-49 */
    size_t n = bytes / 3;
    for (size_t i = 0; i < n; i++) out[i] = ld_s24le(in + 3 * i);
    return n;
}
size_t sko_s24le_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :51-64 */
    size_t n = bytes / 3;
    for (size_t i = 0; i < n; i++) out[i] = (int16_t)(ld_s24le(in + 3 * i) >> 8);
    return n;
}
size_t sko_s24be_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :66-79 */
    size_t n = bytes / 3;
    for (size_t i = 0; i < n; i++) out[i] = (int16_t)(ld_s24be(in + 3 * i) >> 8);
    return n;
}
size_t sko_s32le_to_i32(const uint8_t* in, size_t bytes, int32_t* out) {          /* :81-89 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = ld_i32le(in + 4 * i);
    return n;
}
size_t sko_s32be_to_i32(const uint8_t* in, size_t bytes, int32_t* out) {          /* :91-99 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = ld_i32be(in + 4 * i);
    return n;
}
size_t sko_s32le_to_s24(const uint8_t* in, size_t bytes, int32_t* out) {          /* :101-110 keeps LOW 24 bits */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = ld_i32le(in + 4 * i) & 0x00FFFFFF;
    return n;
}
size_t sko_s32be_to_s24(const uint8_t* in, size_t bytes, int32_t* out) {          /* :112-121 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = ld_i32be(in + 4 * i) & 0x00FFFFFF;
    return n;
}
/* :123-143 — divisor (2.0f32.powi(31) - 1.0) rounds to 2^31 in f32 */
size_t sko_s32le_to_f32(const uint8_t* in, size_t bytes, float* out) {
    size_t n = bytes / 4;
    volatile float div = 2147483648.0f - 1.0f;
    for (size_t i = 0; i < n; i++) out[i] = (float)ld_i32le(in + 4 * i) / div;
    return n;
}
size_t sko_s32be_to_f32(const uint8_t* in, size_t bytes, float* out) {
    size_t n = bytes / 4;
    volatile float div = 2147483648.0f - 1.0f;
    for (size_t i = 0; i < n; i++) out[i] = (float)ld_i32be(in + 4 * i) / div;
    return n;
}
size_t sko_s32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :145-154 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = (int16_t)(ld_i32le(in + 4 * i) >> 16);
    return n;
}
size_t sko_s32be_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :156-165 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = (int16_t)(ld_i32be(in + 4 * i) >> 16);
    return n;
}
size_t sko_f32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :167-176 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = as_i16_f(clamp1(bits_f32(ld_i32le(in + 4 * i))) * 32767.0f);
    return n;
}
size_t sko_f32be_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {          /* :178-187 */
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = as_i16_f(clamp1(bits_f32(ld_i32be(in + 4 * i))) * 32767.0f);
    return n;
}
/* :189-203 — both multipliers are 2^31 in f32; `as` saturates */
static int32_t f32_to_i32_sample(float x) {
    float c = clamp1(x);
    if (c >= 0.0f) return as_i32_f(c * 2147483648.0f /* i32::MAX as f32 */);
    return as_i32_f(c * 2147483648.0f /* -(i32::MIN as f32) */);
}
/* :205-220 and soundkit-decoder/src/lib.rs:3653-3659 */
static int32_t f32_to_s24_sample(float x) {
    float c = clamp1(x);
    if (c >= 0.0f) return as_i32_f(c * 8388607.0f);
    return as_i32_f(c * 8388608.0f);
}
size_t sko_f32le_to_i32(const uint8_t* in, size_t bytes, int32_t* out) {
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = f32_to_i32_sample(bits_f32(ld_i32le(in + 4 * i)));
    return n;
}
size_t sko_f32le_to_s24(const uint8_t* in, size_t bytes, int32_t* out) {
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) out[i] = f32_to_s24_sample(bits_f32(ld_i32le(in + 4 * i)));
    return n;
}

/* ---- interleave / deinterleave --------------------------------------- */
void sko_interleave_vecs_i16(const int16_t* planar, size_t channels, size_t frames, uint8_t* out) { /* :250-262 */
    for (size_t i = 0; i < frames; i++)
        for (size_t c = 0; c < channels; c++) {
            int16_t v = planar[c * frames + i];
            *out++ = (uint8_t)v;
            *out++ = (uint8_t)((uint16_t)v >> 8);
        }
}
size_t sko_deinterleave_vecs_i16(const uint8_t* in, size_t bytes, size_t channels, int16_t* planar) { /* :264-278 */
    size_t frames = bytes / (channels * 2);
    for (size_t i = 0; i < frames; i++)
        for (size_t c = 0; c < channels; c++) planar[c * frames + i] = ld_i16le(in + (i * channels + c) * 2);
    return frames;
}
size_t sko_deinterleave_vecs_s24(const uint8_t* in, size_t bytes, size_t channels, int32_t* planar) { /* :280-294 */
    size_t frames = bytes / (channels * 3);
    for (size_t i = 0; i < frames; i++)
        for (size_t c = 0; c < channels; c++) planar[c * frames + i] = ld_s24le(in + (i * channels + c) * 3);
    return frames;
}
size_t sko_deinterleave_vecs_f32(const uint8_t* in, size_t bytes, size_t channels, float* planar) { /* :296-310 bit copy */
    size_t frames = bytes / (channels * 4);
    for (size_t i = 0; i < frames; i++)
        for (size_t c = 0; c < channels; c++) memcpy(&planar[c * frames + i], in + (i * channels + c) * 4, 4);
    return frames;
}
void sko_stereo_to_mono_take_left(const int16_t* in, size_t n, int16_t* out) {    /* :317-342 */
    for (size_t i = 0; i < n / 2; i++) out[i] = in[2 * i];
}
void sko_stereo_to_mono_avg(const int16_t* in, size_t n, int16_t* out) {          /* :344-373; `/` truncates toward 0 */
    for (size_t i = 0; i < n / 2; i++) out[i] = (int16_t)(((int32_t)in[2 * i] + (int32_t)in[2 * i + 1]) / 2);
}

/* ---- audio_pipeline.rs scalers --------------------------------------- */
void sko_vec_f32_to_i16(const float* in, size_t n, int16_t* out) {                /* :17-27 */
    for (size_t i = 0; i < n; i++) out[i] = as_i16_f(clamp1(in[i]) * 32767.0f);
}
void sko_vec_i16_to_f32(const int16_t* in, size_t n, float* out) {                /* :29-38 */
    for (size_t i = 0; i < n; i++) out[i] = (float)in[i] / 32768.0f;
}
void sko_vec_i32_to_f32(const int32_t* in, size_t n, float* out) {                /* :40-50; 2147483647.0f32 == 2^31 */
    for (size_t i = 0; i < n; i++) out[i] = (float)in[i] / 2147483648.0f;
}

/* ---- to planar f32: mode A (audio_pipeline.rs:74-98) / mode B (decoder lib.rs:3563-3617) ---- */
size_t sko_pcm_to_planar_f32(const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian,
                             size_t channels, int mode, float* out) {
    if (channels == 0) return (size_t)-1;
    if (mode == SKO_MODE_A) {
        /* always little-endian; 16 -> /2^15; 24 -> sign-extended /2^31 (sic); 32 int -> /2^31;
         * 32 float -> bit copy; trailing partial frame dropped (chunks_exact). */
        if (bits == 32 && fmt != SKO_FMT_FLOAT) {
            /* :80-88 — idx % channels split of ALL whole samples (a trailing partial frame's
             * samples still land in the leading channels).  We require whole frames here. */
            size_t n = bytes / 4, frames = n / channels;
            if (n % channels) return (size_t)-1;
            for (size_t i = 0; i < n; i++)
                out[(i % channels) * frames + i / channels] = (float)ld_i32le(in + 4 * i) / 2147483648.0f;
            return frames;
        }
        if (bits == 16) {
            size_t frames = bytes / (channels * 2);
            for (size_t i = 0; i < frames; i++)
                for (size_t c = 0; c < channels; c++)
                    out[c * frames + i] = (float)ld_i16le(in + (i * channels + c) * 2) / 32768.0f;
            return frames;
        }
        if (bits == 24) {
            size_t frames = bytes / (channels * 3);
            for (size_t i = 0; i < frames; i++)
                for (size_t c = 0; c < channels; c++)
                    out[c * frames + i] = (float)ld_s24le(in + (i * channels + c) * 3) / 2147483648.0f;
            return frames;
        }
        if (bits == 32) return sko_deinterleave_vecs_f32(in, bytes, channels, out);
        return (size_t)-1; /* "deserialize_audio failed: unsuporrted type" */
    }
    /* mode B */
    size_t bps = (size_t)((bits + 7) / 8);
    if (!(bits == 16 || bits == 24 || bits == 32) || bytes % bps != 0 || (bytes / bps) % channels != 0)
        return (size_t)-1; /* "PCM data is unsupported or contains a partial frame" */
    if (fmt == SKO_FMT_FLOAT && bits != 32) return (size_t)-1; /* "floating-point PCM must contain 32-bit samples" */
    size_t n = bytes / bps, frames = n / channels;
    for (size_t i = 0; i < n; i++) {
        const uint8_t* p = in + i * bps;
        float s;
        if (fmt == SKO_FMT_FLOAT) s = bits_f32(big_endian ? ld_i32be(p) : ld_i32le(p));
        else if (bits == 16) s = (float)(big_endian ? ld_i16be(p) : ld_i16le(p)) / 32768.0f;
        else if (bits == 24) s = (float)(big_endian ? ld_s24be(p) : ld_s24le(p)) / 8388608.0f;
        else s = (float)(big_endian ? ld_i32be(p) : ld_i32le(p)) / 2147483648.0f;
        out[(i % channels) * frames + i / channels] = is_finite_f(s) ? s : 0.0f; /* :3614 */
    }
    return frames;
}

/* mixdown_to_mono_f32, audio_pipeline.rs:105-131: accumulate from 0.0 in channel order, then * (1/n) */
void sko_mixdown_to_mono_f32(const float* planar, size_t channels, size_t frames, float* out) {
    if (channels == 0) return;
    if (channels == 1) { memcpy(out, planar, frames * sizeof(float)); return; }
    float scale = 1.0f / (float)channels;
    for (size_t i = 0; i < frames; i++) {
        float m = 0.0f;
        for (size_t c = 0; c < channels; c++) m += planar[c * frames + i];
        out[i] = m * scale;
    }
}

/* downmix_channels, soundkit-decoder/src/lib.rs:3492-3561 */
size_t sko_downmix_channels(const float* planar, size_t channels, size_t frames, int target, float* out) {
    if (channels == 0 || target == 0) return 0;
    if (target == 1) {
        float scale = 1.0f / (float)channels;
        for (size_t i = 0; i < frames; i++) {
            float m = 0.0f;
            for (size_t c = 0; c < channels; c++) m += planar[c * frames + i] * scale;
            out[i] = m;
        }
        return 1;
    }
    if (target == 2 && channels > 2) {
        const float k = 0.707f;
        float* l = out; float* r = out + frames;
        float maxv = 0.0f;
        for (size_t i = 0; i < frames; i++) {
            l[i] = planar[i];
            r[i] = planar[frames + i];
            l[i] += k * planar[2 * frames + i];
            r[i] += k * planar[2 * frames + i];
            if (channels > 4) {
                l[i] += k * planar[4 * frames + i];
                if (channels > 5) r[i] += k * planar[5 * frames + i];
            }
        }
        /* .map(abs).fold(0.0, f32::max): f32::max ignores NaN operands */
        for (size_t i = 0; i < 2 * frames; i++) { float a = fabsf(out[i]); if (a > maxv) maxv = a; }
        if (maxv > 1.0f) {
            float scale = 1.0f / maxv;
            for (size_t i = 0; i < 2 * frames; i++) out[i] *= scale;
        }
        return 2;
    }
    memcpy(out, planar, (size_t)target * frames * sizeof(float)); /* first N channels */
    return (size_t)target;
}

/* float_sample_to_i16, soundkit-decoder/src/lib.rs:1815-1827 */
int16_t sko_float_sample_to_i16(float s) {
    float f = is_finite_f(s) ? clamp1(s) : 0.0f;
    double scaled = f < 0.0f ? (double)f * 32768.0 : (double)f * 32767.0;
    int32_t r = as_i32_d(round(scaled));
    if (r < INT16_MIN) r = INT16_MIN;
    if (r > INT16_MAX) r = INT16_MAX;
    return (int16_t)r;
}

/* f32_channels_to_bytes, soundkit-decoder/src/lib.rs:3619-3701 */
size_t sko_f32_channels_to_bytes(const float* planar, size_t channels, size_t frames, int bits, int fmt,
                                 uint8_t* out) {
    if (channels == 0) return 0;
    if (fmt == SKO_FMT_FLOAT) {
        if (bits != 32) return (size_t)-1;
        for (size_t i = 0; i < frames; i++)
            for (size_t c = 0; c < channels; c++) { memcpy(out, &planar[c * frames + i], 4); out += 4; }
        return frames * channels * 4;
    }
    if (bits != 16 && bits != 24 && bits != 32) return (size_t)-1;
    for (size_t i = 0; i < frames; i++)
        for (size_t c = 0; c < channels; c++) {
            float x = planar[c * frames + i];
            if (bits == 16) {
                int16_t v = sko_float_sample_to_i16(x);
                *out++ = (uint8_t)v; *out++ = (uint8_t)((uint16_t)v >> 8);
            } else if (bits == 24) {
                uint32_t v = (uint32_t)f32_to_s24_sample(x);
                *out++ = (uint8_t)v; *out++ = (uint8_t)(v >> 8); *out++ = (uint8_t)(v >> 16);
            } else {
                uint32_t v = (uint32_t)f32_to_i32_sample(x);
                *out++ = (uint8_t)v; *out++ = (uint8_t)(v >> 8); *out++ = (uint8_t)(v >> 16); *out++ = (uint8_t)(v >> 24);
            }
        }
    return frames * channels * (size_t)(bits / 8);
}

/* exact_signed_pcm_to_i16, soundkit-decoder/src/lib.rs:3458-3489 */
size_t sko_exact_signed_pcm_to_i16(const uint8_t* in, size_t bytes, int bits, int big_endian, uint8_t* out) {
    if (bits != 24 && bits != 32) return (size_t)-1;
    size_t bps = (size_t)bits / 8;
    if (bytes % bps) return (size_t)-1; /* "integer PCM contains a partial sample" */
    int shift = bits - 16;
    size_t n = bytes / bps;
    for (size_t i = 0; i < n; i++) {
        const uint8_t* p = in + i * bps;
        int32_t s = bits == 24 ? (big_endian ? ld_s24be(p) : ld_s24le(p)) : (big_endian ? ld_i32be(p) : ld_i32le(p));
        int16_t v = (int16_t)(s >> shift);
        out[2 * i] = (uint8_t)v; out[2 * i + 1] = (uint8_t)((uint16_t)v >> 8);
    }
    return n * 2;
}


/* ---- packet scalers (soundkit/src/audio_packet.rs) and container-PCM unpackers ---------- */
/* audio_packet.rs:176-180, 243-247: f32::from(i16) / f32::from(i16::MAX) */
size_t sko_pkt_s16le_to_f32(const uint8_t* in, size_t bytes, float* out) {
    size_t n = bytes / 2;
    for (size_t i = 0; i < n; i++) out[i] = (float)(int16_t)(in[2 * i] | (in[2 * i + 1] << 8)) / 32767.0f;
    return n;
}
/* audio_packet.rs:182-186, 250-254: LE::read_i24 as f32 / (1 << 23) as f32 */
size_t sko_pkt_s24le_to_f32(const uint8_t* in, size_t bytes, float* out) {
    size_t n = bytes / 3;
    for (size_t i = 0; i < n; i++) {
        uint32_t u = in[3 * i] | ((uint32_t)in[3 * i + 1] << 8) | ((uint32_t)in[3 * i + 2] << 16);
        int32_t v = (u & 0x800000u) ? (int32_t)(u | 0xff000000u) : (int32_t)u;
        out[i] = (float)v / 8388608.0f;
    }
    return n;
}
/* audio_packet.rs:105-124 (PCMSigned arm): (s32 as i64 * i16::MAX as i64 / i32::MAX as i64) as i32, clamped to i16 */
size_t sko_pkt_s32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) {
        int32_t s = (int32_t)(in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24));
        int32_t q = (int32_t)(((int64_t)s * 32767) / 2147483647);
        out[i] = (int16_t)(q < -32768 ? -32768 : (q > 32767 ? 32767 : q));
    }
    return n;
}
/* audio_packet.rs:113-124 (float arm): (x * 32767.0) as i32 (Rust `as`: truncates, saturates, NaN -> 0), clamped to i16 */
size_t sko_pkt_f32le_to_i16(const uint8_t* in, size_t bytes, int16_t* out) {
    size_t n = bytes / 4;
    for (size_t i = 0; i < n; i++) {
        float x; memcpy(&x, in + 4 * i, 4);
        float y = x * 32767.0f;
        int32_t q = (y != y) ? 0 : (y >= 2147483648.0f ? INT32_MAX : (y <= -2147483648.0f ? INT32_MIN : (int32_t)y));
        out[i] = (int16_t)(q < -32768 ? -32768 : (q > 32767 ? 32767 : q));
    }
    return n;
}
/* soundkit-decoder/src/lib.rs:362-367 (AVI 8-bit PCM): (i16::from(sample) - 128) << 8 */
size_t sko_u8_to_s16(const uint8_t* in, size_t bytes, int16_t* out) {
    for (size_t i = 0; i < bytes; i++) out[i] = (int16_t)(((int)in[i] - 128) * 256);
    return bytes;
}
/* unpack_dvd_lpcm, soundkit-decoder/src/lib.rs:750-804.  Returns bytes written, or (size_t)-1 for the Err arms
 * (partial block / 20-bit). */
size_t sko_unpack_dvd_lpcm(const uint8_t* packed, size_t bytes, int bits, int channels, uint8_t* out) {
    if (bits == 16) {
        if (bytes % ((size_t)channels * 2)) return (size_t)-1;
        for (size_t i = 0; i + 1 < bytes; i += 2) { out[i] = packed[i + 1]; out[i + 1] = packed[i]; }
        return bytes;
    }
    if (bits == 20) return (size_t)-1;
    size_t groups_per_block = (channels == 1 || channels == 2 || channels == 4) ? 1 : (channels == 8 ? 2 : (size_t)channels);
    size_t block_size = (channels == 1 || channels == 2 || channels == 4) ? 12 : (channels == 8 ? 24 : (size_t)channels * 12);
    if (bytes % block_size) return (size_t)-1;
    size_t w = 0;
    for (size_t b = 0; b < bytes; b += block_size) {
        size_t cursor = 0;
        for (size_t g = 0; g < groups_per_block; g++) {
            const uint8_t* high = packed + b + cursor;
            const uint8_t* low = packed + b + cursor + 8;
            for (int s = 0; s < 4; s++) { out[w++] = low[s]; out[w++] = high[2 * s + 1]; out[w++] = high[2 * s]; }
            cursor += 12;
        }
    }
    return w;
}

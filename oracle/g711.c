/* oracle/g711.c — TEST INFRASTRUCTURE ONLY (see sk_oracle.h).
 * Restates soundkit-g711/src/lib.rs:3-322 (mu-law / A-law, table-free bit arithmetic). */
#include "sk_oracle.h"

#define ULAW_BIAS 0x84   /* soundkit-g711/src/lib.rs:3 */
#define ULAW_CLIP 32635  /* :4 */
static const int32_t SEGMENT_ENDS[8] = {0xff, 0x1ff, 0x3ff, 0x7ff, 0xfff, 0x1fff, 0x3fff, 0x7fff}; /* :5-14 */

/* segment_for, :317-322 */
static int segment_for(int32_t v) {
    for (int i = 0; i < 8; i++)
        if (v <= SEGMENT_ENDS[i]) return i;
    return 8;
}

/* encode_mulaw_sample, :235-254 */
static uint8_t enc_mu(int16_t s) {
    int32_t pcm = s;
    uint8_t mask;
    if (pcm < 0) { pcm = ULAW_BIAS - pcm; mask = 0x7f; }
    else         { pcm += ULAW_BIAS;      mask = 0xff; }
    if (pcm > ULAW_CLIP) pcm = ULAW_CLIP;
    int seg = segment_for(pcm);
    uint8_t enc = seg >= 8 ? 0x7f : (uint8_t)((seg << 4) | ((pcm >> (seg + 3)) & 0x0f));
    return enc ^ mask;
}

/* decode_mulaw_sample, :256-269 */
static int16_t dec_mu(uint8_t b) {
    uint8_t u = (uint8_t)~b;
    int32_t m = ((int32_t)(u & 0x0f)) << 3;
    m += ULAW_BIAS;
    m <<= (u & 0x70) >> 4;
    int32_t d = (u & 0x80) ? ULAW_BIAS - m : m - ULAW_BIAS;
    return (int16_t)d;
}

/* encode_alaw_sample, :271-294 */
static uint8_t enc_a(int16_t s) {
    int32_t pcm = s;
    uint8_t mask;
    if (pcm >= 0) mask = 0xd5;
    else { pcm = -pcm - 1; mask = 0x55; }
    int seg = segment_for(pcm);
    uint8_t enc;
    if (seg >= 8) enc = 0x7f;
    else {
        enc = (uint8_t)(seg << 4);
        if (seg < 2) enc |= (uint8_t)((pcm >> 4) & 0x0f);
        else         enc |= (uint8_t)((pcm >> (seg + 3)) & 0x0f);
    }
    return enc ^ mask;
}

/* decode_alaw_sample, :296-315 */
static int16_t dec_a(uint8_t b) {
    uint8_t a = b ^ 0x55;
    int seg = (a & 0x70) >> 4;
    int32_t m = ((int32_t)(a & 0x0f)) << 4;
    if (seg == 0) m += 8;
    else if (seg == 1) m += 0x108;
    else { m += 0x108; m <<= (seg - 1); }
    return (a & 0x80) ? (int16_t)m : (int16_t)(-(int16_t)m);
}

uint8_t sko_g711_encode_sample(int law, int16_t s) { return law == SKO_LAW_MULAW ? enc_mu(s) : enc_a(s); }  /* :22-27 */
int16_t sko_g711_decode_sample(int law, uint8_t b) { return law == SKO_LAW_MULAW ? dec_mu(b) : dec_a(b); }  /* :29-34 */

/* encode_i16, :36-50 */
void sko_g711_encode_i16(int law, const int16_t* in, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) out[i] = sko_g711_encode_sample(law, in[i]);
}
/* G711Encoder::encode_i32, :124-138 — (sample >> 16) as i16 */
void sko_g711_encode_i32(int law, const int32_t* in, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; i++) out[i] = sko_g711_encode_sample(law, (int16_t)(in[i] >> 16));
}
/* decode_i16, :52-66 */
void sko_g711_decode_i16(int law, const uint8_t* in, size_t n, int16_t* out) {
    for (size_t i = 0; i < n; i++) out[i] = sko_g711_decode_sample(law, in[i]);
}
/* G711Decoder::decode_i32, :192-211 — i32::from(x) << 16 */
void sko_g711_decode_i32(int law, const uint8_t* in, size_t n, int32_t* out) {
    for (size_t i = 0; i < n; i++)
        out[i] = (int32_t)((uint32_t)(int32_t)sko_g711_decode_sample(law, in[i]) << 16);
}
/* G711Decoder::decode_f32, :213-232 — f32::from(x) / 32768.0 */
void sko_g711_decode_f32(int law, const uint8_t* in, size_t n, float* out) {
    for (size_t i = 0; i < n; i++) out[i] = (float)sko_g711_decode_sample(law, in[i]) / 32768.0f;
}

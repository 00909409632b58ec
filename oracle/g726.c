/*
 * g726.c — CPU restatement of soundkit-g726 (TEST INFRASTRUCTURE ONLY, see sk_oracle.h).
 *
 * Follows soundkit-g726/src/lib.rs:215-548 (G726Core: encode_sample :249, decode_code :279, predictor_zero :305,
 * predictor_pole :313, step_size :317, update :334, quan :465, fmult :472, quantize :499, reconstruct :517,
 * pack_code_group :538, unpack_code_group :568) — itself a port of the Sun Microsystems G.72x reference code.
 * Pinned by the reference's fixtures: testdata/g726/<name>_{16,24,32,40}.g726 decode to exactly the payloads of
 * golden/g726/<name>_{16,24,32,40}.decoded.wav (tests/test_oracle_golden.py).
 */
#include <stdlib.h>
#include <string.h>

#include "sk_oracle.h"

static const int Q16[1] = {261};
static const int DQLN16[4] = {116, 365, 365, 116};
static const int WI16[4] = {-22, 439, 439, -22};
static const int FI16[4] = {0, 0xE00, 0xE00, 0};
static const int Q24[3] = {8, 218, 331};
static const int DQLN24[8] = {-2048, 135, 273, 373, 373, 273, 135, -2048};
static const int WI24[8] = {-4, 30, 137, 582, 582, 137, 30, -4};
static const int FI24[8] = {0, 0x200, 0x400, 0xE00, 0xE00, 0x400, 0x200, 0};
static const int Q32[7] = {-124, 80, 178, 246, 300, 349, 400};
static const int DQLN32[16] = {-2048, 4, 135, 213, 273, 323, 373, 425, 425, 373, 323, 273, 213, 135, 4, -2048};
static const int WI32[16] = {-12, 18, 41, 64, 112, 198, 355, 1122, 1122, 355, 198, 112, 64, 41, 18, -12};
static const int FI32[16] = {0, 0, 0, 0x200, 0x200, 0x200, 0x600, 0xE00, 0xE00, 0x600, 0x200, 0x200, 0x200, 0, 0, 0};
static const int Q40[15] = {-122, -16, 68, 139, 198, 250, 298, 339, 378, 413, 445, 475, 502, 528, 553};
static const int DQLN40[32] = {-2048, -66, 28, 104, 169, 224, 274, 318, 358, 395, 429, 459, 488, 514, 539, 566,
                               566, 539, 514, 488, 459, 429, 395, 358, 318, 274, 224, 169, 104, 28, -66, -2048};
static const int WI40[32] = {14, 14, 24, 39, 40, 41, 58, 100, 141, 179, 219, 280, 358, 440, 529, 696,
                             696, 529, 440, 358, 280, 219, 179, 141, 100, 58, 41, 40, 39, 24, 14, 14};
static const int FI40[32] = {0, 0, 0, 0, 0, 0x200, 0x200, 0x200, 0x200, 0x200, 0x400, 0x600, 0x800, 0xA00, 0xC00, 0xC00,
                             0xC00, 0xC00, 0xA00, 0x800, 0x600, 0x400, 0x200, 0x200, 0x200, 0x200, 0x200, 0, 0, 0, 0, 0};

typedef struct { int bits, group_samples, group_bytes, nq; const int *q, *dqln, *wi, *fi; } rate_t;
static int rate_for(int bits, rate_t* r) { /* G726Rate, :62-189 */
    switch (bits) {
        case 2: *r = (rate_t){2, 4, 1, 1, Q16, DQLN16, WI16, FI16}; return 1;
        case 3: *r = (rate_t){3, 8, 3, 3, Q24, DQLN24, WI24, FI24}; return 1;
        case 4: *r = (rate_t){4, 2, 1, 7, Q32, DQLN32, WI32, FI32}; return 1;
        case 5: *r = (rate_t){5, 8, 5, 15, Q40, DQLN40, WI40, FI40}; return 1;
    }
    return 0;
}

/* G726Core as 24 ints: yl yu dms dml ap a[2] b[6] pk[2] dq[6] sr[2] td (:198-230) */
enum { YL = 0, YU, DMS, DML, AP, A0, A1, B0, PK0 = B0 + 6, PK1, DQ0, SR0 = DQ0 + 6, SR1, TD, NSTATE };

void sko_g726_state_init(int32_t* s) {
    memset(s, 0, SKO_G726_STATE_INTS * sizeof(int32_t));
    s[YL] = 34816; s[YU] = 544;
    for (int i = 0; i < 6; i++) s[DQ0 + i] = 32;
    s[SR0] = 32; s[SR1] = 32;
}

static int quan_pow2(int v) { /* quan(value, &POWER2): first i with v < 2^i, i < 15, else 15 (:465-470) */
    for (int i = 0; i < 15; i++) if (v < (1 << i)) return i;
    return 15;
}
static int quan_tab(int v, const int* t, int n) { for (int i = 0; i < n; i++) if (v < t[i]) return i; return n; }

static int fmult(int an, int srn) { /* :472-497 */
    int anmag = an > 0 ? an : ((-an) & 0x1FFF);
    int anexp = quan_pow2(anmag) - 6;
    int anmant = anmag == 0 ? 32 : (anexp >= 0 ? anmag >> anexp : anmag << -anexp);
    int wanexp = anexp + ((srn >> 6) & 0x0F) - 13;
    int wanmant = (anmant * (srn & 0x3F) + 0x30) >> 4;
    int retval = wanexp >= 0 ? ((wanmant << wanexp) & 0x7FFF) : (wanmant >> -wanexp);
    return ((an ^ srn) < 0) ? -retval : retval;
}
static int predictor_zero(const int32_t* s) { int t = 0; for (int i = 0; i < 6; i++) t += fmult(s[B0 + i] >> 2, s[DQ0 + i]); return t; }
static int predictor_pole(const int32_t* s) { return fmult(s[A1] >> 2, s[SR1]) + fmult(s[A0] >> 2, s[SR0]); }
static int step_size(const int32_t* s) { /* :317-332 */
    if (s[AP] >= 256) return s[YU];
    int y = s[YL] >> 6, dif = s[YU] - y, al = s[AP] >> 2;
    if (dif > 0) return y + ((dif * al) >> 6);
    if (dif < 0) return y + ((dif * al + 0x3F) >> 6);
    return y;
}
static int quantize(int d, int y, const rate_t* r) { /* :499-515 */
    int dqm = d < 0 ? -d : d;
    int exp = quan_pow2(dqm >> 1);
    int mant = ((dqm << 7) >> exp) & 0x7F;
    int dln = (exp << 7) + mant - (y >> 2);
    int i = quan_tab(dln, r->q, r->nq);
    int mask = (1 << r->bits) - 1;
    if (d < 0) return mask - i;
    return i == 0 ? mask : i;
}
static int reconstruct(int sign, int dqln, int y) { /* :517-536 */
    int dql = dqln + (y >> 2);
    if (dql < 0) return sign ? -0x8000 : 0;
    int dex = (dql >> 7) & 15, dqt = 128 + (dql & 127);
    int dq = (dqt << 7) >> (14 - dex);
    return sign ? dq - 0x8000 : dq;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void update(int32_t* s, int y, int wi, int fi, int dq, int sr, int dqsez, int is40) { /* :334-462 */
    int pk0 = dqsez < 0, mag = dq & 0x7FFF;
    int ylint = s[YL] >> 15, ylfrac = (s[YL] >> 10) & 0x1F;
    int thr1 = (32 + ylfrac) << ylint;
    int thr2 = ylint > 9 ? (31 << 10) : thr1;
    int dqthr = (thr2 + (thr2 >> 1)) >> 1;
    int tr = (s[TD] != 0 && mag > dqthr);
    s[YU] = clampi(y + ((wi - y) >> 5), 544, 5120);
    s[YL] += s[YU] + ((-s[YL]) >> 6);
    int a2p;
    if (tr) {
        s[A0] = s[A1] = 0;
        for (int i = 0; i < 6; i++) s[B0 + i] = 0;
        a2p = 0;
    } else {
        int pks1 = pk0 ^ s[PK0];
        a2p = s[A1] - (s[A1] >> 7);
        if (dqsez != 0) {
            int fa1 = pks1 ? s[A0] : -s[A0];
            if (fa1 < -8191) a2p -= 0x100; else if (fa1 > 8191) a2p += 0xFF; else a2p += fa1 >> 5;
            if (pk0 ^ s[PK1]) {
                if (a2p <= -12160) a2p = -12288; else if (a2p >= 12416) a2p = 12288; else a2p -= 0x80;
            } else if (a2p <= -12416) a2p = -12288;
            else if (a2p >= 12160) a2p = 12288;
            else a2p += 0x80;
        }
        s[A1] = a2p;
        s[A0] -= s[A0] >> 8;
        if (dqsez != 0) s[A0] += pks1 == 0 ? 192 : -192;
        int a1ul = 15360 - a2p;
        s[A0] = clampi(s[A0], -a1ul, a1ul);
        for (int i = 0; i < 6; i++) {
            s[B0 + i] -= s[B0 + i] >> (is40 ? 9 : 8);
            if (dq & 0x7FFF) s[B0 + i] += ((dq ^ s[DQ0 + i]) >= 0) ? 128 : -128;
        }
    }
    for (int i = 5; i >= 1; i--) s[DQ0 + i] = s[DQ0 + i - 1];
    if (mag == 0) s[DQ0] = dq >= 0 ? 0x20 : -0x3E0;
    else {
        int exp = quan_pow2(mag);
        int value = (exp << 6) + ((mag << 6) >> exp);
        s[DQ0] = dq >= 0 ? value : value - 0x400;
    }
    s[SR1] = s[SR0];
    if (sr == 0) s[SR0] = 0x20;
    else if (sr > 0) { int exp = quan_pow2(sr); s[SR0] = (exp << 6) + ((sr << 6) >> exp); }
    else if (sr > -32768) { int m = -sr, exp = quan_pow2(m); s[SR0] = (exp << 6) + ((m << 6) >> exp) - 0x400; }
    else s[SR0] = -0x3E0;
    s[PK1] = s[PK0]; s[PK0] = pk0;
    s[TD] = tr ? 0 : (a2p < -11776);
    s[DMS] += (fi - s[DMS]) >> 5;
    s[DML] += ((fi << 2) - s[DML]) >> 7;
    int t = (s[DMS] << 2) - s[DML];
    if (t < 0) t = -t;
    if (tr) s[AP] = 256;
    else if (y < 1536 || s[TD] != 0 || t >= (s[DML] >> 3)) s[AP] += (0x200 - s[AP]) >> 4;
    else s[AP] += (-s[AP]) >> 4;
}

static int encode_sample(int32_t* s, int16_t sample, const rate_t* r) { /* :249-277 */
    int sl = (int)sample >> 2;
    int sezi = predictor_zero(s), sez = sezi >> 1;
    int se = (sezi + predictor_pole(s)) >> 1;
    int d = sl - se, y = step_size(s);
    int i = quantize(d, y, r);
    int dq = reconstruct((i & (1 << (r->bits - 1))) != 0, r->dqln[i], y);
    int dq_mask = r->bits == 5 ? 0x7FFF : 0x3FFF;
    int sr = dq < 0 ? se - (dq & dq_mask) : se + dq;
    int dqsez = sr + sez - se;
    update(s, y, r->wi[i] << 5, r->fi[i], dq, sr, dqsez, r->bits == 5);
    return i & ((1 << r->bits) - 1);
}
static int16_t decode_code(int32_t* s, int code, const rate_t* r) { /* :279-303 */
    int i = code & ((1 << r->bits) - 1);
    int sezi = predictor_zero(s), sez = sezi >> 1;
    int se = (sezi + predictor_pole(s)) >> 1;
    int y = step_size(s);
    int dq = reconstruct((i & (1 << (r->bits - 1))) != 0, r->dqln[i], y);
    int dq_mask = r->bits == 5 ? 0x7FFF : 0x3FFF;
    int sr = dq < 0 ? se - (dq & dq_mask) : se + dq;
    int dqsez = sr - se + sez;
    update(s, y, r->wi[i] << 5, r->fi[i], dq, sr, dqsez, r->bits == 5);
    return (int16_t)clampi(sr << 2, -32768, 32767);
}

/* pack_code_group / unpack_code_group, :538-596.  packing 0 = Left (MSB first), 1 = Right (LSB first) */
static void pack_group(const uint8_t* codes, int n, int bits, int packing, uint8_t* out, int nbytes) {
    memset(out, 0, (size_t)nbytes);
    int bi = 0;
    for (int c = 0; c < n; c++) {
        if (packing == 0) { for (int b = bits - 1; b >= 0; b--, bi++) if ((codes[c] >> b) & 1) out[bi / 8] |= (uint8_t)(1 << (7 - (bi % 8))); }
        else { for (int b = 0; b < bits; b++, bi++) if ((codes[c] >> b) & 1) out[bi / 8] |= (uint8_t)(1 << (bi % 8)); }
    }
}
static void unpack_group(const uint8_t* in, int n, int bits, int packing, uint8_t* codes) {
    int bi = 0;
    for (int c = 0; c < n; c++) {
        codes[c] = 0;
        if (packing == 0) { for (int b = 0; b < bits; b++, bi++) codes[c] = (uint8_t)((codes[c] << 1) | ((in[bi / 8] >> (7 - (bi % 8))) & 1)); }
        else { for (int b = 0; b < bits; b++, bi++) codes[c] |= (uint8_t)(((in[bi / 8] >> (bi % 8)) & 1) << b); }
    }
}

/* whole groups only (the Rust wrappers keep the partial group in pending_bytes / pending_samples, :631-742, :821-860);
 * `state` carries G726Core between calls.  Return samples / bytes written, (size_t)-1 on a bad rate or partial group. */
size_t sko_g726_decode(int bits, int packing, int32_t* state, const uint8_t* in, size_t bytes, int16_t* out) {
    rate_t r;
    if (!rate_for(bits, &r) || bytes % (size_t)r.group_bytes) return (size_t)-1;
    uint8_t codes[8];
    size_t w = 0;
    for (size_t g = 0; g < bytes; g += (size_t)r.group_bytes) {
        unpack_group(in + g, r.group_samples, r.bits, packing, codes);
        for (int c = 0; c < r.group_samples; c++) out[w++] = decode_code(state, codes[c], &r);
    }
    return w;
}
size_t sko_g726_encode(int bits, int packing, int32_t* state, const int16_t* in, size_t n, uint8_t* out) {
    rate_t r;
    if (!rate_for(bits, &r) || n % (size_t)r.group_samples) return (size_t)-1;
    uint8_t codes[8];
    size_t w = 0;
    for (size_t g = 0; g < n; g += (size_t)r.group_samples) {
        for (int c = 0; c < r.group_samples; c++) codes[c] = (uint8_t)encode_sample(state, in[g + c], &r);
        pack_group(codes, r.group_samples, r.bits, packing, out + w, r.group_bytes);
        w += (size_t)r.group_bytes;
    }
    return w;
}

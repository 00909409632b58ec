// sk_g726.cu — SURVEY.md 8(f) row 4: G.726 ADPCM (16/24/32/40 kbit/s) for batches of independent telephony streams.
//
// Replaces soundkit-g726/src/lib.rs:215-596 (G726Core::encode_sample :249, decode_code :279, predictor_zero :305,
// predictor_pole :313, step_size :317, update :334, fmult :472, quantize :499, reconstruct :517 and the
// left/right-justified code packing :538-596).  Integer arithmetic only, bit-exact.
//
// ADPCM is sample-serial inside a stream (every sample updates the predictor), so the parallel axis is the STREAM:
// one thread per stream, the 24-word predictor state in registers, the per-rate tables in shared memory (indexed by
// a per-lane code), the quantiser's threshold search unrolled against compile-time constants, and
// `quan(x, POWER2)` (a 15-step linear search in the reference) as one FLO.  The code bitstream is read / written
// through a 64-bit bit reservoir (MSB-first for Left packing, LSB-first for Right), 32 bits at a time.  This path is
// instruction-latency bound (~300 dependent integer ops per sample), not HBM bound: 65 536 streams give 443 resident
// threads per SM, and the 8 fmult() of one sample are independent of each other (ILP 8).
#include <new>

#include "sk_common.cuh"

#define G726_NSTATE SK_G726_STATE_INTS
enum { G_YL = 0, G_YU, G_DMS, G_DML, G_AP, G_A0, G_A1, G_B0, G_PK0 = G_B0 + 6, G_PK1, G_DQ0, G_SR0 = G_DQ0 + 6, G_SR1, G_TD };

template <int BITS> struct G726Tab;
template <> struct G726Tab<2> {
    static constexpr int NQ = 1, GS = 4, GB = 1;
    static constexpr int q[1] = {261};
    static constexpr int dqln[4] = {116, 365, 365, 116};
    static constexpr int wi[4] = {-22, 439, 439, -22};
    static constexpr int fi[4] = {0, 0xE00, 0xE00, 0};
};
template <> struct G726Tab<3> {
    static constexpr int NQ = 3, GS = 8, GB = 3;
    static constexpr int q[3] = {8, 218, 331};
    static constexpr int dqln[8] = {-2048, 135, 273, 373, 373, 273, 135, -2048};
    static constexpr int wi[8] = {-4, 30, 137, 582, 582, 137, 30, -4};
    static constexpr int fi[8] = {0, 0x200, 0x400, 0xE00, 0xE00, 0x400, 0x200, 0};
};
template <> struct G726Tab<4> {
    static constexpr int NQ = 7, GS = 2, GB = 1;
    static constexpr int q[7] = {-124, 80, 178, 246, 300, 349, 400};
    static constexpr int dqln[16] = {-2048, 4, 135, 213, 273, 323, 373, 425, 425, 373, 323, 273, 213, 135, 4, -2048};
    static constexpr int wi[16] = {-12, 18, 41, 64, 112, 198, 355, 1122, 1122, 355, 198, 112, 64, 41, 18, -12};
    static constexpr int fi[16] = {0, 0, 0, 0x200, 0x200, 0x200, 0x600, 0xE00, 0xE00, 0x600, 0x200, 0x200, 0x200, 0, 0, 0};
};
template <> struct G726Tab<5> {
    static constexpr int NQ = 15, GS = 8, GB = 5;
    static constexpr int q[15] = {-122, -16, 68, 139, 198, 250, 298, 339, 378, 413, 445, 475, 502, 528, 553};
    static constexpr int dqln[32] = {-2048, -66, 28, 104, 169, 224, 274, 318, 358, 395, 429, 459, 488, 514, 539, 566,
                                     566, 539, 514, 488, 459, 429, 395, 358, 318, 274, 224, 169, 104, 28, -66, -2048};
    static constexpr int wi[32] = {14, 14, 24, 39, 40, 41, 58, 100, 141, 179, 219, 280, 358, 440, 529, 696,
                                   696, 529, 440, 358, 280, 219, 179, 141, 100, 58, 41, 40, 39, 24, 14, 14};
    static constexpr int fi[32] = {0, 0, 0, 0, 0, 0x200, 0x200, 0x200, 0x200, 0x200, 0x400, 0x600, 0x800, 0xA00, 0xC00, 0xC00,
                                   0xC00, 0xC00, 0xA00, 0x800, 0x600, 0x400, 0x200, 0x200, 0x200, 0x200, 0x200, 0, 0, 0, 0, 0};
};

// quan(value, &POWER2) (:465-470): first i < 15 with value < 2^i, else 15 == min(bit_length(value), 15), 0 for value <= 0
__device__ __forceinline__ int g726_quan2(int v) { return v <= 0 ? 0 : min(32 - __clz(v), 15); }

__device__ __forceinline__ int g726_fmult(int an, int srn) {  // :472-497
    const int anmag = an > 0 ? an : ((-an) & 0x1FFF);
    const int anexp = g726_quan2(anmag) - 6;
    const int anmant = anmag == 0 ? 32 : (anexp >= 0 ? anmag >> anexp : anmag << -anexp);
    const int wanexp = anexp + ((srn >> 6) & 0x0F) - 13;
    const int wanmant = (anmant * (srn & 0x3F) + 0x30) >> 4;
    const int retval = wanexp >= 0 ? ((wanmant << wanexp) & 0x7FFF) : (wanmant >> -wanexp);
    return ((an ^ srn) < 0) ? -retval : retval;
}

struct G726Core { int s[G726_NSTATE]; };

__device__ __forceinline__ int g726_step_size(const G726Core& c) {  // :317-332
    if (c.s[G_AP] >= 256) return c.s[G_YU];
    const int y = c.s[G_YL] >> 6, dif = c.s[G_YU] - y, al = c.s[G_AP] >> 2;
    if (dif > 0) return y + ((dif * al) >> 6);
    if (dif < 0) return y + ((dif * al + 0x3F) >> 6);
    return y;
}
__device__ __forceinline__ int g726_reconstruct(bool sign, int dqln, int y) {  // :517-536
    const int dql = dqln + (y >> 2);
    if (dql < 0) return sign ? -0x8000 : 0;
    const int dex = (dql >> 7) & 15, dqt = 128 + (dql & 127);
    const int dq = (dqt << 7) >> (14 - dex);
    return sign ? dq - 0x8000 : dq;
}
// log-format of a magnitude (the `exp << 6 + mantissa` form stored in dq[] / sr[], :425-452)
__device__ __forceinline__ int g726_logfmt(int mag) { const int e = g726_quan2(mag); return (e << 6) + ((mag << 6) >> e); }

template <bool IS40>
__device__ __forceinline__ void g726_update(G726Core& c, int y, int wi, int fi, int dq, int sr, int dqsez) {  // :334-462
    int* s = c.s;
    const int pk0 = dqsez < 0, mag = dq & 0x7FFF;
    const int ylint = s[G_YL] >> 15, ylfrac = (s[G_YL] >> 10) & 0x1F;
    const int thr1 = (32 + ylfrac) << ylint;
    const int thr2 = ylint > 9 ? (31 << 10) : thr1;
    const int dqthr = (thr2 + (thr2 >> 1)) >> 1;
    const bool tr = s[G_TD] != 0 && mag > dqthr;
    s[G_YU] = max(544, min(5120, y + ((wi - y) >> 5)));
    s[G_YL] += s[G_YU] + ((-s[G_YL]) >> 6);
    int a2p = 0;
    if (tr) {
        s[G_A0] = 0; s[G_A1] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) s[G_B0 + i] = 0;
    } else {
        const int pks1 = pk0 ^ s[G_PK0];
        a2p = s[G_A1] - (s[G_A1] >> 7);
        if (dqsez != 0) {
            const int fa1 = pks1 ? s[G_A0] : -s[G_A0];
            a2p += fa1 < -8191 ? -0x100 : (fa1 > 8191 ? 0xFF : (fa1 >> 5));
            if (pk0 ^ s[G_PK1]) a2p = a2p <= -12160 ? -12288 : (a2p >= 12416 ? 12288 : a2p - 0x80);
            else a2p = a2p <= -12416 ? -12288 : (a2p >= 12160 ? 12288 : a2p + 0x80);
        }
        s[G_A1] = a2p;
        s[G_A0] -= s[G_A0] >> 8;
        if (dqsez != 0) s[G_A0] += pks1 == 0 ? 192 : -192;
        const int a1ul = 15360 - a2p;
        s[G_A0] = max(-a1ul, min(a1ul, s[G_A0]));
#pragma unroll
        for (int i = 0; i < 6; i++) {
            s[G_B0 + i] -= s[G_B0 + i] >> (IS40 ? 9 : 8);
            if (mag != 0) s[G_B0 + i] += ((dq ^ s[G_DQ0 + i]) >= 0) ? 128 : -128;
        }
    }
#pragma unroll
    for (int i = 5; i >= 1; i--) s[G_DQ0 + i] = s[G_DQ0 + i - 1];
    s[G_DQ0] = mag == 0 ? (dq >= 0 ? 0x20 : -0x3E0) : (dq >= 0 ? g726_logfmt(mag) : g726_logfmt(mag) - 0x400);
    s[G_SR1] = s[G_SR0];
    s[G_SR0] = sr == 0 ? 0x20 : (sr > 0 ? g726_logfmt(sr) : (sr > -32768 ? g726_logfmt(-sr) - 0x400 : -0x3E0));
    s[G_PK1] = s[G_PK0];
    s[G_PK0] = pk0;
    s[G_TD] = tr ? 0 : (a2p < -11776);
    s[G_DMS] += (fi - s[G_DMS]) >> 5;
    s[G_DML] += ((fi << 2) - s[G_DML]) >> 7;
    if (tr) s[G_AP] = 256;
    else if (y < 1536 || s[G_TD] != 0 || abs((s[G_DMS] << 2) - s[G_DML]) >= (s[G_DML] >> 3)) s[G_AP] += (0x200 - s[G_AP]) >> 4;
    else s[G_AP] += (-s[G_AP]) >> 4;
}

// signal estimate shared by both directions: sezi (zeros) and se (zeros + poles); the 8 fmult() are independent
__device__ __forceinline__ void g726_predict(const G726Core& c, int& sez, int& se) {
    int sezi = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) sezi += g726_fmult(c.s[G_B0 + i] >> 2, c.s[G_DQ0 + i]);
    const int pole = g726_fmult(c.s[G_A1] >> 2, c.s[G_SR1]) + g726_fmult(c.s[G_A0] >> 2, c.s[G_SR0]);
    sez = sezi >> 1;
    se = (sezi + pole) >> 1;
}

// quan(dln, q_table) (:505): the tables are sorted, so "first index with dln < q[i]" is the number of thresholds <= dln;
// written against literals (G726Tab<BITS>::q) so the search is NQ compares and adds with no memory access
template <int BITS> __device__ __forceinline__ int g726_qcount(int dln);
template <> __device__ __forceinline__ int g726_qcount<2>(int d) { return d >= 261; }
template <> __device__ __forceinline__ int g726_qcount<3>(int d) { return (d >= 8) + (d >= 218) + (d >= 331); }
template <> __device__ __forceinline__ int g726_qcount<4>(int d) {
    return (d >= -124) + (d >= 80) + (d >= 178) + (d >= 246) + (d >= 300) + (d >= 349) + (d >= 400);
}
template <> __device__ __forceinline__ int g726_qcount<5>(int d) {
    return (d >= -122) + (d >= -16) + (d >= 68) + (d >= 139) + (d >= 198) + (d >= 250) + (d >= 298) + (d >= 339) + (d >= 378) + (d >= 413) +
           (d >= 445) + (d >= 475) + (d >= 502) + (d >= 528) + (d >= 553);
}

// tab: shared memory [3][32] = dqln | wi << 5 | fi
template <int BITS>
__device__ __forceinline__ int g726_decode_code(G726Core& c, int i, const int* tab) {  // :279-303
    int sez, se;
    g726_predict(c, sez, se);
    const int y = g726_step_size(c);
    const int dq = g726_reconstruct((i & (1 << (BITS - 1))) != 0, tab[i], y);
    const int sr = dq < 0 ? se - (dq & (BITS == 5 ? 0x7FFF : 0x3FFF)) : se + dq;
    g726_update<BITS == 5>(c, y, tab[32 + i], tab[64 + i], dq, sr, sr - se + sez);
    return max(-32768, min(32767, sr << 2));
}
template <int BITS>
__device__ __forceinline__ int g726_encode_sample(G726Core& c, int sample, const int* tab) {  // :249-277
    int sez, se;
    g726_predict(c, sez, se);
    const int d = (sample >> 2) - se;
    const int y = g726_step_size(c);
    // quantize, :499-515: the sorted threshold search is a count of thresholds <= dln, against immediates
    const int dqm = abs(d);
    const int e = g726_quan2(dqm >> 1);
    const int dln = (e << 7) + (((dqm << 7) >> e) & 0x7F) - (y >> 2);
    const int q = g726_qcount<BITS>(dln);
    constexpr int MASK = (1 << BITS) - 1;
    const int i = d < 0 ? MASK - q : (q == 0 ? MASK : q);
    const int dq = g726_reconstruct((i & (1 << (BITS - 1))) != 0, tab[i], y);
    const int sr = dq < 0 ? se - (dq & (BITS == 5 ? 0x7FFF : 0x3FFF)) : se + dq;
    g726_update<BITS == 5>(c, y, tab[32 + i], tab[64 + i], dq, sr, sr + sez - se);
    return i & MASK;
}

// [rate][3][32] = dqln | wi | fi in constant memory (filled from G726Tab on first use); each CTA copies its rate's rows to
// shared memory because the hot-loop index is a per-lane code (constant memory would serialise divergent indices)
__constant__ int g726_c_tab[4][96];
template <int BITS>
__device__ __forceinline__ void g726_load_tables(int* tab) {
    for (int i = threadIdx.x; i < 96; i += blockDim.x) tab[i] = i >= 32 && i < 64 ? g726_c_tab[BITS - 2][i] << 5 : g726_c_tab[BITS - 2][i];
    __syncthreads();
}
__device__ __forceinline__ void g726_core_default(G726Core& c) {  // Default for G726Core, :214-230
#pragma unroll
    for (int i = 0; i < G726_NSTATE; i++) c.s[i] = 0;
    c.s[G_YL] = 34816; c.s[G_YU] = 544;
#pragma unroll
    for (int i = 0; i < 6; i++) c.s[G_DQ0 + i] = 32;
    c.s[G_SR0] = 32; c.s[G_SR1] = 32;
}
__device__ __forceinline__ void g726_core_load(G726Core& c, const int32_t* state, size_t stream) {
    if (!state) { g726_core_default(c); return; }
#pragma unroll
    for (int i = 0; i < G726_NSTATE; i++) c.s[i] = state[stream * G726_NSTATE + i];
}
__device__ __forceinline__ void g726_core_store(const G726Core& c, int32_t* state, size_t stream) {
    if (!state) return;
#pragma unroll
    for (int i = 0; i < G726_NSTATE; i++) state[stream * G726_NSTATE + i] = c.s[i];
}

// 32 bits of the code stream starting at byte `off` (< bytes); bytes past the end read as 0.  `wordwise`: the stream
// base is 4-byte aligned, so whole words come from one 32-bit load.
__device__ __forceinline__ uint32_t g726_fetch32(const uint8_t* p, size_t off, size_t bytes, bool wordwise) {
    if (wordwise && off + 4 <= bytes) return __ldg(reinterpret_cast<const uint32_t*>(p + off));
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (off + b < bytes) w |= (uint32_t)p[off + b] << (8 * b);
    return w;
}

constexpr int kG726Threads = 128;

// KIND: output element (sk_kind).  One thread per stream; `bytes` per stream is a whole number of code groups.
template <int BITS, int PACKING, int KIND>
__global__ void __launch_bounds__(kG726Threads) g726_decode_kernel(const uint8_t* __restrict__ in, size_t in_stride, size_t n_streams, size_t bytes,
                                                                   void* __restrict__ outv, size_t out_stride, int32_t* __restrict__ state) {
    __shared__ int tab[96];
    g726_load_tables<BITS>(tab);
    const size_t stream = (size_t)blockIdx.x * kG726Threads + threadIdx.x;
    if (stream >= n_streams) return;
    G726Core c;
    g726_core_load(c, state, stream);
    const uint8_t* p = in + stream * in_stride;
    const bool wordwise = ((uintptr_t)p & 3) == 0;
    const size_t n = bytes * 8 / BITS;
    uint64_t res = 0;
    int navail = 0;
    size_t off = 0;
    constexpr int OB = KIND == SK_KIND_I16 ? 2 : 4;
    uint8_t* o = reinterpret_cast<uint8_t*>(outv) + stream * out_stride * OB;
    const bool vec = ((uintptr_t)o & 15) == 0;
    constexpr int PER = 16 / OB;  // samples per 128-bit store
    uint32_t buf[4];
    for (size_t k = 0; k < n; k++) {
        if (navail < BITS) {
            const uint32_t w = g726_fetch32(p, off, bytes, wordwise);
            off += 4;
            if (PACKING == 0) res |= (uint64_t)bswap32(w) << (32 - navail);  // MSB first: reservoir is left aligned
            else res |= (uint64_t)w << navail;                                // LSB first: right aligned
            navail += 32;
        }
        int code;
        if (PACKING == 0) { code = (int)(res >> (64 - BITS)); res <<= BITS; }
        else { code = (int)(res & ((1u << BITS) - 1)); res >>= BITS; }
        navail -= BITS;
        const int v = g726_decode_code<BITS>(c, code, tab);
        // G726Decoder::decode_{i16,i32,f32}, :821-951
        uint32_t e;
        if (KIND == SK_KIND_I16) e = (uint32_t)v & 0xffffu;
        else if (KIND == SK_KIND_I32) e = (uint32_t)v << 16;
        else e = __float_as_uint(__fmul_rn((float)v, 1.0f / 32768.0f));
        const int slot = (int)(k % PER);
        if (OB == 2) { if (slot & 1) buf[slot >> 1] |= e << 16; else buf[slot >> 1] = e; }
        else buf[slot] = e;
        if (slot == PER - 1 || k + 1 == n) {
            const size_t first = k - slot;
            if (vec && slot == PER - 1) st_stream_v4(o + first * OB, make_uint4(buf[0], buf[1], buf[2], buf[3]));
            else {
                for (int j = 0; j <= slot; j++) {
                    if (OB == 2) reinterpret_cast<uint16_t*>(o)[first + j] = (uint16_t)(buf[j >> 1] >> (16 * (j & 1)));
                    else reinterpret_cast<uint32_t*>(o)[first + j] = buf[j];
                }
            }
        }
    }
    g726_core_store(c, state, stream);
}

// KIND: input element, i16 or i32 ((s >> 16) as i16, :735-738).  `n` samples per stream, a whole number of groups.
template <int BITS, int PACKING, int KIND>
__global__ void __launch_bounds__(kG726Threads) g726_encode_kernel(const void* __restrict__ inv, size_t in_stride, size_t n_streams, size_t n,
                                                                   uint8_t* __restrict__ out, size_t out_stride, int32_t* __restrict__ state) {
    __shared__ int tab[96];
    g726_load_tables<BITS>(tab);
    const size_t stream = (size_t)blockIdx.x * kG726Threads + threadIdx.x;
    if (stream >= n_streams) return;
    G726Core c;
    g726_core_load(c, state, stream);
    constexpr int IB = KIND == SK_KIND_I16 ? 2 : 4;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(inv) + stream * in_stride * IB;
    uint8_t* o = out + stream * out_stride;
    const bool wordwise = ((uintptr_t)o & 3) == 0;
    const size_t bytes = n * BITS / 8;
    uint64_t res = 0;
    int nbits = 0;
    size_t off = 0;
    for (size_t k = 0; k < n; k++) {
        const int sample = KIND == SK_KIND_I16 ? (int)__ldg(reinterpret_cast<const int16_t*>(p) + k)
                                               : (int)(int16_t)(__ldg(reinterpret_cast<const int32_t*>(p) + k) >> 16);
        const uint64_t code = (uint64_t)g726_encode_sample<BITS>(c, sample, tab);
        if (PACKING == 0) res |= code << (64 - BITS - nbits); else res |= code << nbits;
        nbits += BITS;
        while (nbits >= 32 || (k + 1 == n && nbits > 0)) {  // whole words; at the end of the stream also the partial one
            const uint32_t w = PACKING == 0 ? bswap32((uint32_t)(res >> 32)) : (uint32_t)res;
            if (wordwise && off + 4 <= bytes) *reinterpret_cast<uint32_t*>(o + off) = w;
            else
                for (int b = 0; b < 4; b++)
                    if (off + b < bytes) o[off + b] = (uint8_t)(w >> (8 * b));
            off += 4;
            if (PACKING == 0) res <<= 32; else res >>= 32;
            nbits -= 32;  // goes negative only after the final partial word
        }
    }
    g726_core_store(c, state, stream);
}

// ---- host side -----------------------------------------------------------------------------------
extern "C" void sk_g726_state_init(int32_t* state) {  // Default for G726Core
    if (!state) return;
    memset(state, 0, G726_NSTATE * sizeof(int32_t));
    state[G_YL] = 34816; state[G_YU] = 544;
    for (int i = 0; i < 6; i++) state[G_DQ0 + i] = 32;
    state[G_SR0] = 32; state[G_SR1] = 32;
}

template <int BITS>
static void g726_fill_rows(int (*h)[96]) {
    using T = G726Tab<BITS>;
    for (int i = 0; i < (1 << BITS); i++) { h[BITS - 2][i] = T::dqln[i]; h[BITS - 2][32 + i] = T::wi[i]; h[BITS - 2][64 + i] = T::fi[i]; }
}
static int g726_upload_tables(sk_ctx* ctx) {  // per device, once per context
    if (ctx->g726_tables_ready) return SK_OK;
    int h[4][96];
    memset(h, 0, sizeof(h));
    g726_fill_rows<2>(h); g726_fill_rows<3>(h); g726_fill_rows<4>(h); g726_fill_rows<5>(h);
    SK_CUDA(ctx, cudaMemcpyToSymbolAsync(g726_c_tab, h, sizeof(h), 0, cudaMemcpyHostToDevice, ctx->stream));
    SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `h` is on the stack
    ctx->g726_tables_ready = true;
    return SK_OK;
}

static bool g726_group(int bits, size_t* gs, size_t* gb) {  // samples_per_byte_group / bytes_per_group, :114-130
    switch (bits) {
        case 2: *gs = 4; *gb = 1; return true;
        case 3: *gs = 8; *gb = 3; return true;
        case 4: *gs = 2; *gb = 1; return true;
        case 5: *gs = 8; *gb = 5; return true;
    }
    return false;
}

template <int BITS, int PACKING>
static void g726_launch_decode(sk_ctx* ctx, const uint8_t* in, size_t in_stride, size_t n_streams, size_t bytes, int kind, void* out,
                               size_t out_stride, int32_t* state) {
    const unsigned grid = (unsigned)((n_streams + kG726Threads - 1) / kG726Threads);
    if (kind == SK_KIND_I16) g726_decode_kernel<BITS, PACKING, SK_KIND_I16><<<grid, kG726Threads, 0, ctx->stream>>>(in, in_stride, n_streams, bytes, out, out_stride, state);
    else if (kind == SK_KIND_I32) g726_decode_kernel<BITS, PACKING, SK_KIND_I32><<<grid, kG726Threads, 0, ctx->stream>>>(in, in_stride, n_streams, bytes, out, out_stride, state);
    else g726_decode_kernel<BITS, PACKING, SK_KIND_F32><<<grid, kG726Threads, 0, ctx->stream>>>(in, in_stride, n_streams, bytes, out, out_stride, state);
}
template <int BITS, int PACKING>
static void g726_launch_encode(sk_ctx* ctx, const void* in, size_t in_stride, size_t n_streams, size_t n, int kind, uint8_t* out, size_t out_stride,
                               int32_t* state) {
    const unsigned grid = (unsigned)((n_streams + kG726Threads - 1) / kG726Threads);
    if (kind == SK_KIND_I16) g726_encode_kernel<BITS, PACKING, SK_KIND_I16><<<grid, kG726Threads, 0, ctx->stream>>>(in, in_stride, n_streams, n, out, out_stride, state);
    else g726_encode_kernel<BITS, PACKING, SK_KIND_I32><<<grid, kG726Threads, 0, ctx->stream>>>(in, in_stride, n_streams, n, out, out_stride, state);
}

#define G726_DISPATCH(FN, ...)                                                         \
    do {                                                                               \
        if (packing == 0) {                                                            \
            if (bits == 2) FN<2, 0>(__VA_ARGS__); else if (bits == 3) FN<3, 0>(__VA_ARGS__); \
            else if (bits == 4) FN<4, 0>(__VA_ARGS__); else FN<5, 0>(__VA_ARGS__);     \
        } else {                                                                       \
            if (bits == 2) FN<2, 1>(__VA_ARGS__); else if (bits == 3) FN<3, 1>(__VA_ARGS__); \
            else if (bits == 4) FN<4, 1>(__VA_ARGS__); else FN<5, 1>(__VA_ARGS__);     \
        }                                                                              \
    } while (0)

extern "C" int sk_g726_decode_batch_dev(sk_ctx* ctx, int bits, int packing, const uint8_t* in, size_t in_stride, size_t n_streams,
                                        size_t bytes, int out_kind, void* out, size_t out_stride, int32_t* state) {
    if (!ctx) return SK_E_ARG;
    size_t gs, gb;
    if (!g726_group(bits, &gs, &gb)) SK_FAIL(ctx, SK_E_ARG, "G.726 code size must be 2, 3, 4 or 5 bits, got %d", bits);
    if (packing != 0 && packing != 1) SK_FAIL(ctx, SK_E_ARG, "unknown G.726 packing %d", packing);
    if (out_kind < SK_KIND_I16 || out_kind > SK_KIND_F32) SK_FAIL(ctx, SK_E_ARG, "unknown output kind %d", out_kind);
    if (bytes % gb) SK_FAIL(ctx, SK_E_ARG, "G.726 input is not a whole number of %zu-byte code groups", gb);
    if (!bytes || !n_streams) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    const size_t ob = out_kind == SK_KIND_I16 ? 2 : 4;
    if (((uintptr_t)out | (out_stride * ob)) & (ob - 1)) SK_FAIL(ctx, SK_E_ARG, "output not aligned to its element size");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc = g726_upload_tables(ctx); if (rc) return rc; }
    G726_DISPATCH(g726_launch_decode, ctx, in, in_stride, n_streams, bytes, out_kind, out, out_stride, state);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

extern "C" int sk_g726_encode_batch_dev(sk_ctx* ctx, int bits, int packing, int in_kind, const void* in, size_t in_stride, size_t n_streams,
                                        size_t n, uint8_t* out, size_t out_stride, int32_t* state) {
    if (!ctx) return SK_E_ARG;
    size_t gs, gb;
    if (!g726_group(bits, &gs, &gb)) SK_FAIL(ctx, SK_E_ARG, "G.726 code size must be 2, 3, 4 or 5 bits, got %d", bits);
    if (packing != 0 && packing != 1) SK_FAIL(ctx, SK_E_ARG, "unknown G.726 packing %d", packing);
    if (in_kind != SK_KIND_I16 && in_kind != SK_KIND_I32) SK_FAIL(ctx, SK_E_ARG, "G.726 encode takes i16 or i32 input");
    if (n % gs) SK_FAIL(ctx, SK_E_ARG, "G.726 input is not a whole number of %zu-sample code groups", gs);
    if (!n || !n_streams) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    const size_t ib = in_kind == SK_KIND_I16 ? 2 : 4;
    if (((uintptr_t)in | (in_stride * ib)) & (ib - 1)) SK_FAIL(ctx, SK_E_ARG, "input not aligned to its element size");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc = g726_upload_tables(ctx); if (rc) return rc; }
    G726_DISPATCH(g726_launch_encode, ctx, in, in_stride, n_streams, n, in_kind, out, out_stride, state);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

// host-pointer, one stream, explicit predictor state (the Rust G726Decoder / G726Encoder own `core` and the pending
// partial group; they hand whole groups to these calls)
extern "C" int sk_g726_decode(sk_ctx* ctx, int bits, int packing, int32_t* state, const uint8_t* in, size_t bytes, int out_kind, void* out) {
    if (!ctx) return SK_E_ARG;
    if (!state) SK_FAIL(ctx, SK_E_ARG, "null state");
    size_t gs, gb;
    if (!g726_group(bits, &gs, &gb)) SK_FAIL(ctx, SK_E_ARG, "G.726 code size must be 2, 3, 4 or 5 bits, got %d", bits);
    if (bytes % gb) SK_FAIL(ctx, SK_E_ARG, "G.726 input is not a whole number of %zu-byte code groups", gb);
    if (!bytes) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = bytes / gb * gs, ob = out_kind == SK_KIND_I16 ? 2 : 4;
    const size_t in_off = 256;  // state lives in front of the codes in the device input arena
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_off + bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, n * ob))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, state, G726_NSTATE * sizeof(int32_t)))) return rc;
    if ((rc = sk_h2d(ctx, (char*)ctx->d_in + in_off, in, bytes))) return rc;
    if ((rc = sk_g726_decode_batch_dev(ctx, bits, packing, (const uint8_t*)ctx->d_in + in_off, 0, 1, bytes, out_kind, ctx->d_out, 0, (int32_t*)ctx->d_in)))
        return rc;
    if ((rc = sk_d2h(ctx, state, ctx->d_in, G726_NSTATE * sizeof(int32_t)))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, n * ob);
}

extern "C" int sk_g726_encode(sk_ctx* ctx, int bits, int packing, int32_t* state, int in_kind, const void* in, size_t n, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (!state) SK_FAIL(ctx, SK_E_ARG, "null state");
    size_t gs, gb;
    if (!g726_group(bits, &gs, &gb)) SK_FAIL(ctx, SK_E_ARG, "G.726 code size must be 2, 3, 4 or 5 bits, got %d", bits);
    if (in_kind != SK_KIND_I16 && in_kind != SK_KIND_I32) SK_FAIL(ctx, SK_E_ARG, "G.726 encode takes i16 or i32 input");
    if (n % gs) SK_FAIL(ctx, SK_E_ARG, "G.726 input is not a whole number of %zu-sample code groups", gs);
    if (!n) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ib = in_kind == SK_KIND_I16 ? 2 : 4, bytes = n / gs * gb;
    const size_t in_off = 256;
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_off + n * ib))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, bytes + 4))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, state, G726_NSTATE * sizeof(int32_t)))) return rc;
    if ((rc = sk_h2d(ctx, (char*)ctx->d_in + in_off, in, n * ib))) return rc;
    if ((rc = sk_g726_encode_batch_dev(ctx, bits, packing, in_kind, (const char*)ctx->d_in + in_off, 0, 1, n, (uint8_t*)ctx->d_out, 0, (int32_t*)ctx->d_in)))
        return rc;
    if ((rc = sk_d2h(ctx, state, ctx->d_in, G726_NSTATE * sizeof(int32_t)))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, bytes);
}

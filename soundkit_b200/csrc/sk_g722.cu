// sk_g722.cu — SURVEY.md 8(f) row 4: G.722 64 kbit/s sub-band ADPCM (mode 1) for batches of independent wideband telephony streams.
//
// Replaces what soundkit-g722/src/lib.rs:1-230 reaches through the crates.io dependency `ezk-g722 = "0.1.1"`
// (libg722::{decoder::Decoder, encoder::Encoder}, Bitrate::Mode1_64000, eight_k = false, packed = false): ITU-T G.722 blocks 1-6
// per sub-band (quantiser, inverse quantisers, log scale factor adaptation, the 2-pole / 6-zero predictor of block 4) and the
// 24-tap quadrature mirror filter between the two 8 kHz bands and the 16 kHz signal.  Integer arithmetic only, bit-exact.
// The decoder is pinned by the reference's fixture (testdata/g722/*.g722 -> golden/g722/*.decoded.wav); that fixture also pins the
// one place where the crate departs from the Recommendation's reset values: both bands start with a ZERO scale factor (DETL 32 /
// DETH 8 in the Recommendation decode the first 2 570 samples of the fixture differently).
//
// Like G.726 the codec is sample-serial inside a stream, so the parallel axis is the STREAM: one thread per stream, both band
// predictors (22 words each) and the QMF delay line (24 words) in registers, tables in shared / constant memory, one code byte in
// and two samples out per step.  Instruction-latency bound, not HBM bound.
#include <new>

#include "sk_common.cuh"

#define G722_NSTATE SK_G722_STATE_INTS  // per band: s, sz, r1, r2, a1, a2, p1, p2, d1..d6, b1..b6, nb, det (22 words) x 2, then x[24]
constexpr int kG722Threads = 128;
constexpr int kBandWords = 22;

struct G722Band { int s, sz, r1, r2, a1, a2, p1, p2, d[6], b[6], nb, det; };
struct G722State { G722Band lo, hi; int x[24]; };

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

__constant__ int c_g722_ilb[32] = {2048, 2093, 2139, 2186, 2233, 2282, 2332, 2383, 2435, 2489, 2543, 2599, 2656, 2714, 2774, 2834,
                                   2896, 2960, 3025, 3091, 3158, 3228, 3298, 3371, 3444, 3520, 3597, 3676, 3756, 3838, 3922, 4008};
__constant__ int c_g722_qm6[64] = {-136, -136, -136, -136, -24808, -21904, -19008, -16704, -14984, -13512, -12280, -11192, -10232, -9360, -8576, -7856,
                                   -7192, -6576, -6000, -5456, -4944, -4464, -4008, -3576, -3168, -2776, -2400, -2032, -1688, -1360, -1040, -728,
                                   24808, 21904, 19008, 16704, 14984, 13512, 12280, 11192, 10232, 9360, 8576, 7856, 7192, 6576, 6000, 5456,
                                   4944, 4464, 4008, 3576, 3168, 2776, 2400, 2032, 1688, 1360, 1040, 728, 432, 136, -432, -136};
__constant__ int c_g722_qm4[16] = {0, -20456, -12896, -8968, -6288, -4240, -2584, -1200, 20456, 12896, 8968, 6288, 4240, 2584, 1200, 0};
__constant__ int c_g722_q6[32] = {0, 35, 72, 110, 150, 190, 233, 276, 323, 370, 422, 473, 530, 587, 650, 714,
                                  786, 858, 940, 1023, 1121, 1219, 1339, 1458, 1612, 1765, 1980, 2195, 2557, 2919, 0, 0};
__constant__ int c_g722_iln[32] = {0, 63, 62, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 0};
__constant__ int c_g722_ilp[32] = {0, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43, 42, 41, 40, 39, 38, 37, 36, 35, 34, 33, 32, 0};

// WL[RL42[ril]] folded: index = ril (4 bits).  rl42 = {0,7,6,5,4,3,2,1,7,6,5,4,3,2,1,0}; wl = {-60,-30,58,172,334,538,1198,3042}
__constant__ int c_g722_wl[16] = {-60, 3042, 1198, 538, 334, 172, 58, -30, 3042, 1198, 538, 334, 172, 58, -30, -60};

// Every table the hot loop indexes with a per-lane value lives in SHARED memory: every lane of a warp is a different stream, so the
// indices diverge, and the constant cache serves one address per cycle (ncu on the first version, all tables in constant memory:
// 46 % of the decoder's stall samples on short_scoreboard).  One 16-byte row per 6-bit low-band code gathers what a step needs.
struct G722Tab {
    int4 low[64];  // [code]: {qm4[code >> 2], wl[code >> 2], qm6[code], 0}; the encoder reads (x, y) of row ril << 2
    int ilb[32];
    int q6[32];    // QUANTL decision levels (encoder)
    int il[32];    // iln | ilp << 8 (encoder)
};
__device__ __forceinline__ void g722_tab_fill(G722Tab& T) {  // whole CTA; ends with a barrier
    for (int i = threadIdx.x; i < 64; i += blockDim.x) T.low[i] = make_int4(c_g722_qm4[i >> 2], c_g722_wl[i >> 2], c_g722_qm6[i], 0);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) { T.ilb[i] = c_g722_ilb[i]; T.q6[i] = c_g722_q6[i]; T.il[i] = c_g722_iln[i] | (c_g722_ilp[i] << 8); }
    __syncthreads();
}
__device__ __forceinline__ int g722_wh(int ihigh) {  // WH[RH2[ihigh]]: rh2 = {2,1,2,1}, wh = {0,-214,798}
    return (ihigh & 1) ? -214 : 798;
}
__device__ __forceinline__ int g722_qm2(int ihigh) {  // {-7408, -1616, 7408, 1616}
    const int mag = (ihigh & 1) ? 1616 : 7408;
    return (ihigh & 2) ? mag : -mag;
}

// blocks 3L / 3H: SCALEL (base 8) / SCALEH (base 10)
__device__ __forceinline__ int g722_scale(const G722Tab& T, int nb, int base) {
    const int wd1 = (nb >> 6) & 31, wd2 = base - (nb >> 11);
    const int v = T.ilb[wd1];
    const int wd3 = wd2 < 0 ? (v << -wd2) : (v >> wd2);
    return wd3 << 2;
}

// block 4: RECONS, PARREC, UPPOL2, UPPOL1, UPZERO, DELAYA, FILTEP, FILTEZ, PREDIC
__device__ __forceinline__ void g722_block4(G722Band& b, int d) {
    const int r0 = sat16(b.s + d), p0 = sat16(b.sz + d);
    const int sg0 = p0 >> 15, sg1 = b.p1 >> 15, sg2 = b.p2 >> 15;
    int wd1 = sat16(b.a1 * 4);
    int wd2 = sg0 == sg1 ? -wd1 : wd1;
    if (wd2 > 32767) wd2 = 32767;
    int wd3 = (sg0 == sg2 ? 128 : -128) + (wd2 >> 7) + ((b.a2 * 32512) >> 15);
    const int ap2 = max(-12288, min(12288, wd3));
    wd1 = sg0 == sg1 ? 192 : -192;
    wd2 = (b.a1 * 32640) >> 15;
    int ap1 = sat16(wd1 + wd2);
    wd3 = sat16(15360 - ap2);
    ap1 = ap1 > wd3 ? wd3 : (ap1 < -wd3 ? -wd3 : ap1);
    // UPZERO on the old delay line, then DELAYA
    wd1 = d == 0 ? 0 : 128;
    const int sgd = d >> 15;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        wd2 = (b.d[i] >> 15) == sgd ? wd1 : -wd1;
        wd3 = (b.b[i] * 32640) >> 15;
        b.b[i] = sat16(wd2 + wd3);
    }
#pragma unroll
    for (int i = 5; i > 0; i--) b.d[i] = b.d[i - 1];
    b.d[0] = d;
    b.r2 = b.r1; b.r1 = r0; b.p2 = b.p1; b.p1 = p0; b.a1 = ap1; b.a2 = ap2;
    // FILTEP
    wd1 = sat16(b.r1 + b.r1);
    wd1 = (b.a1 * wd1) >> 15;
    wd2 = sat16(b.r2 + b.r2);
    wd2 = (b.a2 * wd2) >> 15;
    const int sp = sat16(wd1 + wd2);
    // FILTEZ (summed from the oldest tap, as the reference loop does; integer addition, order-free)
    int sz = 0;
#pragma unroll
    for (int i = 5; i >= 0; i--) sz += (b.b[i] * sat16(b.d[i] + b.d[i])) >> 15;
    b.sz = sat16(sz);
    b.s = sat16(sp + b.sz);
}

__device__ __forceinline__ void g722_load(G722State& s, const int32_t* state, size_t stream) {
    if (!state) {
        int* w = reinterpret_cast<int*>(&s);
#pragma unroll
        for (int i = 0; i < G722_NSTATE; i++) w[i] = 0;
        return;
    }
    const int32_t* p = state + stream * G722_NSTATE;
    int* w = reinterpret_cast<int*>(&s);
#pragma unroll
    for (int i = 0; i < G722_NSTATE; i++) w[i] = p[i];
}
__device__ __forceinline__ void g722_store(const G722State& s, int32_t* state, size_t stream) {
    if (!state) return;
    int32_t* p = state + stream * G722_NSTATE;
    const int* w = reinterpret_cast<const int*>(&s);
#pragma unroll
    for (int i = 0; i < G722_NSTATE; i++) p[i] = w[i];
}

// 24-tap QMF halves: x holds the newest pair in x[22], x[23]
__device__ __forceinline__ void g722_qmf(const int (&x)[24], int& sum_even_taps, int& sum_odd_taps) {
    constexpr int c[12] = {3, -11, 12, 32, -210, 951, 3876, -805, 362, -156, 53, -11};
    int e = 0, o = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) { e += x[2 * i] * c[i]; o += x[2 * i + 1] * c[11 - i]; }
    sum_even_taps = e; sum_odd_taps = o;
}

// one code byte -> two samples (libg722 Decoder::decode, 64 kbit/s, 16 kHz output)
__device__ __forceinline__ void g722_decode_byte(const G722Tab& T, G722State& s, int code, int& y0, int& y1) {
    const int4 row = T.low[code & 0x3f];
    const int ihigh = (code >> 6) & 3;
    const int wd2 = (s.lo.det * row.z) >> 15;                       // 5L INVQBL
    int rlow = s.lo.s + wd2;                                        // 5L RECONS
    rlow = max(-16384, min(16383, rlow));                           // 6L LIMIT
    const int dlowt = (s.lo.det * row.x) >> 15;                     // 2L INVQAL
    int nb = ((s.lo.nb * 127) >> 7) + row.y;                        // 3L LOGSCL
    s.lo.nb = max(0, min(18432, nb));
    s.lo.det = g722_scale(T, s.lo.nb, 8);
    g722_block4(s.lo, dlowt);

    const int dhigh = (s.hi.det * g722_qm2(ihigh)) >> 15;           // 2H INVQAH
    int rhigh = dhigh + s.hi.s;                                     // 5H RECONS
    rhigh = max(-16384, min(16383, rhigh));                         // 6H LIMIT
    nb = ((s.hi.nb * 127) >> 7) + g722_wh(ihigh);                   // 3H LOGSCH
    s.hi.nb = max(0, min(22528, nb));
    s.hi.det = g722_scale(T, s.hi.nb, 10);
    g722_block4(s.hi, dhigh);

#pragma unroll
    for (int i = 0; i < 22; i++) s.x[i] = s.x[i + 2];               // receive QMF
    s.x[22] = rlow + rhigh;
    s.x[23] = rlow - rhigh;
    int xout2, xout1;
    g722_qmf(s.x, xout2, xout1);
    y0 = sat16(xout1 >> 11);
    y1 = sat16(xout2 >> 11);
}

// two samples -> one code byte (libg722 Encoder::encode)
__device__ __forceinline__ int g722_encode_pair(const G722Tab& T, G722State& s, int a0, int a1) {
#pragma unroll
    for (int i = 0; i < 22; i++) s.x[i] = s.x[i + 2];               // transmit QMF
    s.x[22] = a0;
    s.x[23] = a1;
    int sumodd, sumeven;
    g722_qmf(s.x, sumodd, sumeven);
    const int xlow = (sumeven + sumodd) >> 14, xhigh = (sumeven - sumodd) >> 14;

    const int el = sat16(xlow - s.lo.s);                            // 1L SUBTRA
    int wd = el >= 0 ? el : -(el + 1);                              // 1L QUANTL
    // the reference walks i = 1 .. 29 until wd < (q6[i] * det) >> 12.  The levels rise with i (det >= 0), so that index is 1 + the
    // number of levels <= wd: a five-step descent over the sorted levels, no data-dependent branch (the walk made every warp pay
    // for its slowest lane)
    int pos = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int k = pos + step;
        const bool ge = wd >= ((T.q6[k & 31] * s.lo.det) >> 12);
        pos = (k <= 29 && ge) ? k : pos;
    }
    const int i = pos + 1;
    const int ilw = T.il[i];
    const int ilow = el < 0 ? (ilw & 0xff) : (ilw >> 8);
    const int ril = ilow >> 2;                                      // 2L INVQAL
    const int2 rr = *reinterpret_cast<const int2*>(&T.low[ril << 2]);
    const int dlow = (s.lo.det * rr.x) >> 15;
    int nb = ((s.lo.nb * 127) >> 7) + rr.y;                         // 3L LOGSCL
    s.lo.nb = max(0, min(18432, nb));
    s.lo.det = g722_scale(T, s.lo.nb, 8);
    g722_block4(s.lo, dlow);

    const int eh = sat16(xhigh - s.hi.s);                           // 1H SUBTRA
    wd = eh >= 0 ? eh : -(eh + 1);                                  // 1H QUANTH
    const int mih = wd >= ((564 * s.hi.det) >> 12) ? 2 : 1;
    // ihn = {0,1,0}, ihp = {0,3,2}
    const int ihigh = eh < 0 ? (mih == 1 ? 1 : 0) : (mih == 1 ? 3 : 2);
    const int dhigh = (s.hi.det * g722_qm2(ihigh)) >> 15;           // 2H INVQAH
    nb = ((s.hi.nb * 127) >> 7) + g722_wh(ihigh);                   // 3H LOGSCH
    s.hi.nb = max(0, min(22528, nb));
    s.hi.det = g722_scale(T, s.hi.nb, 10);
    g722_block4(s.hi, dhigh);
    return (ihigh << 6) | ilow;
}

// KIND: output element (sk_kind).  One thread per stream.
template <int KIND>
__global__ void __launch_bounds__(kG722Threads) g722_decode_kernel(const uint8_t* __restrict__ in, size_t in_stride, size_t n_streams, size_t bytes,
                                                                   void* __restrict__ outv, size_t out_stride, int32_t* __restrict__ state) {
    __shared__ G722Tab T;
    g722_tab_fill(T);
    const size_t stream = (size_t)blockIdx.x * kG722Threads + threadIdx.x;
    if (stream >= n_streams) return;
    G722State s;
    g722_load(s, state, stream);
    const uint8_t* p = in + stream * in_stride;
    constexpr int OB = KIND == SK_KIND_I16 ? 2 : 4;
    uint8_t* o = reinterpret_cast<uint8_t*>(outv) + stream * out_stride * OB;
    const bool in_words = ((uintptr_t)p & 3) == 0, out_vec = ((uintptr_t)o & 15) == 0;
    // G722Decoder::decode_{i16,i32,f32}, soundkit-g722/src/lib.rs:161-232
    auto conv = [](int v) -> uint32_t {
        if (KIND == SK_KIND_I16) return (uint32_t)v & 0xffffu;
        if (KIND == SK_KIND_I32) return (uint32_t)v << 16;
        return __float_as_uint(__fmul_rn((float)v, 1.0f / 32768.0f));
    };
    size_t j = 0;
    // 4 code bytes per step: one 32-bit load in, 8 samples = 16 (i16) or 32 (i32 / f32) bytes out
    for (; j + 4 <= bytes; j += 4) {
        uint32_t w;
        if (in_words) w = __ldg(reinterpret_cast<const uint32_t*>(p + j));
        else w = (uint32_t)p[j] | ((uint32_t)p[j + 1] << 8) | ((uint32_t)p[j + 2] << 16) | ((uint32_t)p[j + 3] << 24);
        uint32_t e[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int y0, y1;
            g722_decode_byte(T, s, (int)((w >> (8 * k)) & 0xffu), y0, y1);
            e[2 * k] = conv(y0); e[2 * k + 1] = conv(y1);
        }
        uint8_t* q = o + 2 * j * OB;
        if (OB == 2) {
            const uint4 v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
            if (out_vec) st_stream_v4(q, v);
            else { uint16_t* h = reinterpret_cast<uint16_t*>(q); for (int k = 0; k < 8; k++) h[k] = (uint16_t)e[k]; }
        } else {
            if (out_vec) { st_stream_v4(q, make_uint4(e[0], e[1], e[2], e[3])); st_stream_v4(q + 16, make_uint4(e[4], e[5], e[6], e[7])); }
            else { uint32_t* h = reinterpret_cast<uint32_t*>(q); for (int k = 0; k < 8; k++) h[k] = e[k]; }
        }
    }
    for (; j < bytes; j++) {
        int y0, y1;
        g722_decode_byte(T, s, (int)p[j], y0, y1);
        if (OB == 2) { uint16_t* h = reinterpret_cast<uint16_t*>(o) + 2 * j; h[0] = (uint16_t)conv(y0); h[1] = (uint16_t)conv(y1); }
        else { uint32_t* h = reinterpret_cast<uint32_t*>(o) + 2 * j; h[0] = conv(y0); h[1] = conv(y1); }
    }
    g722_store(s, state, stream);
}

// KIND: input element, i16 or i32 ((s >> 16) as i16, soundkit-g722/src/lib.rs:119-122).  `n` (even) samples per stream.
template <int KIND>
__global__ void __launch_bounds__(kG722Threads) g722_encode_kernel(const void* __restrict__ inv, size_t in_stride, size_t n_streams, size_t n,
                                                                   uint8_t* __restrict__ out, size_t out_stride, int32_t* __restrict__ state) {
    __shared__ G722Tab T;
    g722_tab_fill(T);
    const size_t stream = (size_t)blockIdx.x * kG722Threads + threadIdx.x;
    if (stream >= n_streams) return;
    G722State s;
    g722_load(s, state, stream);
    constexpr int IB = KIND == SK_KIND_I16 ? 2 : 4;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(inv) + stream * in_stride * IB;
    uint8_t* o = out + stream * out_stride;
    const bool out_words = ((uintptr_t)o & 3) == 0;
    auto sample = [&](size_t k) -> int {
        return KIND == SK_KIND_I16 ? (int)__ldg(reinterpret_cast<const int16_t*>(p) + k) : (int)(int16_t)(__ldg(reinterpret_cast<const int32_t*>(p) + k) >> 16);
    };
    const size_t codes = n / 2;
    size_t c = 0;
    for (; c + 4 <= codes; c += 4) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) w |= (uint32_t)g722_encode_pair(T, s, sample(2 * (c + k)), sample(2 * (c + k) + 1)) << (8 * k);
        if (out_words) *reinterpret_cast<uint32_t*>(o + c) = w;
        else for (int k = 0; k < 4; k++) o[c + k] = (uint8_t)(w >> (8 * k));
    }
    for (; c < codes; c++) o[c] = (uint8_t)g722_encode_pair(T, s, sample(2 * c), sample(2 * c + 1));
    g722_store(s, state, stream);
}

// ---- host side -----------------------------------------------------------------------------------
extern "C" void sk_g722_state_init(int32_t* state) {  // the crate's decoder / encoder start from an all-zero band state (see the header)
    if (state) memset(state, 0, G722_NSTATE * sizeof(int32_t));
}

extern "C" int sk_g722_decode_batch_dev(sk_ctx* ctx, const uint8_t* in, size_t in_stride, size_t n_streams, size_t bytes, int out_kind, void* out,
                                        size_t out_stride, int32_t* state) {
    if (!ctx) return SK_E_ARG;
    if (out_kind < SK_KIND_I16 || out_kind > SK_KIND_F32) SK_FAIL(ctx, SK_E_ARG, "unknown output kind %d", out_kind);
    if (!bytes || !n_streams) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    const size_t ob = out_kind == SK_KIND_I16 ? 2 : 4;
    if (((uintptr_t)out | (out_stride * ob)) & (ob - 1)) SK_FAIL(ctx, SK_E_ARG, "output not aligned to its element size");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((n_streams + kG722Threads - 1) / kG722Threads);
    if (out_kind == SK_KIND_I16) g722_decode_kernel<SK_KIND_I16><<<grid, kG722Threads, 0, ctx->stream>>>(in, in_stride, n_streams, bytes, out, out_stride, state);
    else if (out_kind == SK_KIND_I32) g722_decode_kernel<SK_KIND_I32><<<grid, kG722Threads, 0, ctx->stream>>>(in, in_stride, n_streams, bytes, out, out_stride, state);
    else g722_decode_kernel<SK_KIND_F32><<<grid, kG722Threads, 0, ctx->stream>>>(in, in_stride, n_streams, bytes, out, out_stride, state);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

extern "C" int sk_g722_encode_batch_dev(sk_ctx* ctx, int in_kind, const void* in, size_t in_stride, size_t n_streams, size_t n, uint8_t* out,
                                        size_t out_stride, int32_t* state) {
    if (!ctx) return SK_E_ARG;
    if (in_kind != SK_KIND_I16 && in_kind != SK_KIND_I32) SK_FAIL(ctx, SK_E_ARG, "G.722 encode takes i16 or i32 input");
    if (n % 2) SK_FAIL(ctx, SK_E_ARG, "G.722 input is not a whole number of sample pairs");
    if (!n || !n_streams) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    const size_t ib = in_kind == SK_KIND_I16 ? 2 : 4;
    if (((uintptr_t)in | (in_stride * ib)) & (ib - 1)) SK_FAIL(ctx, SK_E_ARG, "input not aligned to its element size");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((n_streams + kG722Threads - 1) / kG722Threads);
    if (in_kind == SK_KIND_I16) g722_encode_kernel<SK_KIND_I16><<<grid, kG722Threads, 0, ctx->stream>>>(in, in_stride, n_streams, n, out, out_stride, state);
    else g722_encode_kernel<SK_KIND_I32><<<grid, kG722Threads, 0, ctx->stream>>>(in, in_stride, n_streams, n, out, out_stride, state);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

// host-pointer, one stream, explicit codec state (the Rust G722Decoder / G722Encoder own it, and the encoder's pending odd sample)
extern "C" int sk_g722_decode(sk_ctx* ctx, int32_t* state, const uint8_t* in, size_t bytes, int out_kind, void* out) {
    if (!ctx) return SK_E_ARG;
    if (!state) SK_FAIL(ctx, SK_E_ARG, "null state");
    if (out_kind < SK_KIND_I16 || out_kind > SK_KIND_F32) SK_FAIL(ctx, SK_E_ARG, "unknown output kind %d", out_kind);
    if (!bytes) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = bytes * 2, ob = out_kind == SK_KIND_I16 ? 2 : 4;
    const size_t in_off = 512;  // the state lives in front of the codes in the device input arena
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_off + bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, n * ob))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, state, G722_NSTATE * sizeof(int32_t)))) return rc;
    if ((rc = sk_h2d(ctx, (char*)ctx->d_in + in_off, in, bytes))) return rc;
    if ((rc = sk_g722_decode_batch_dev(ctx, (const uint8_t*)ctx->d_in + in_off, 0, 1, bytes, out_kind, ctx->d_out, 0, (int32_t*)ctx->d_in))) return rc;
    if ((rc = sk_d2h(ctx, state, ctx->d_in, G722_NSTATE * sizeof(int32_t)))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, n * ob);
}

extern "C" int sk_g722_encode(sk_ctx* ctx, int32_t* state, int in_kind, const void* in, size_t n, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (!state) SK_FAIL(ctx, SK_E_ARG, "null state");
    if (in_kind != SK_KIND_I16 && in_kind != SK_KIND_I32) SK_FAIL(ctx, SK_E_ARG, "G.722 encode takes i16 or i32 input");
    if (n % 2) SK_FAIL(ctx, SK_E_ARG, "G.722 input is not a whole number of sample pairs");
    if (!n) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ib = in_kind == SK_KIND_I16 ? 2 : 4, bytes = n / 2;
    const size_t in_off = 512;
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_off + n * ib))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, bytes + 4))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, state, G722_NSTATE * sizeof(int32_t)))) return rc;
    if ((rc = sk_h2d(ctx, (char*)ctx->d_in + in_off, in, n * ib))) return rc;
    if ((rc = sk_g722_encode_batch_dev(ctx, in_kind, (const char*)ctx->d_in + in_off, 0, 1, n, (uint8_t*)ctx->d_out, 0, (int32_t*)ctx->d_in))) return rc;
    if ((rc = sk_d2h(ctx, state, ctx->d_in, G722_NSTATE * sizeof(int32_t)))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, bytes);
}

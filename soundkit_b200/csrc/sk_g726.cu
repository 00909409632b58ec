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
// PTX shifts clamp their amount to the register width (a C++ shift by 32 is undefined): x << 32 == 0, x >> 32 == 0
__device__ __forceinline__ uint32_t g726_shl(uint32_t x, int n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }
__device__ __forceinline__ uint32_t g726_shr(uint32_t x, int n) { uint32_t r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }

// One delay-line entry (dq[i] / sr[i], the `exp << 6 + mantissa` log format of :425-452 with the sign as −0x400) together with
// the fields every later sample reads from it, decoded once when the entry is made instead of eight times per sample:
// m8 = mantissa << 8, kk = 30 − exponent, sg = ±1.
struct G726Tap { int m8, kk, sg; };
__device__ __forceinline__ G726Tap g726_tap(int raw) {
    G726Tap t;
    t.m8 = (raw & 0x3F) << 8; t.kk = 30 - ((raw >> 6) & 0xF); t.sg = (raw >> 31) | 1;
    return t;
}
// the word the reference keeps for an entry (exact for every entry the algorithm itself makes out of a 15-bit magnitude)
__device__ __forceinline__ int g726_tap_raw(const G726Tap& t) { return ((30 - t.kk) << 6) + (t.m8 >> 8) - (t.sg < 0 ? 0x400 : 0); }
// log format of a magnitude < 2^15 (0 -> 0x20, :441 / :450) with the sign `neg`
__device__ __forceinline__ G726Tap g726_tap_of(int mag, bool neg) {
    const int c = __clz(mag);                                             // 32 for 0
    const int m = max((int)(g726_shl((uint32_t)mag, c) >> 26), 32);       // the 6 leading bits; 32 for mag == 0
    G726Tap t;
    t.m8 = m << 8; t.kk = c - 2; t.sg = neg ? -1 : 1;
    return t;
}

// fmult (:472-497) against a decoded entry.  anexp = bit_length(anmag) − 6 and anmant = the 6 leading bits of anmag (32 for 0);
// the reference's two-sided shift `wanexp >= 0 ? (wanmant << wanexp) & 0x7FFF : wanmant >> -wanexp` is ONE right shift of
// wanmant << 12 by 12 − wanexp = kk − floor(log2 anmag) ∈ [2, 31] (no low bit is lost while wanexp >= 0, and the mask is a no-op
// for the right shifts: wanmant < 2^8); `(anmant·m + 0x30) >> 4 << 12` is `(anmant·(m << 8) + 0x3000) & ~0xFFF`.
__device__ __forceinline__ int g726_fmult(int an, const G726Tap& t) {
    const int anmag = an > 0 ? an : (abs(an) & 0x1FFF);
    const int c = __clz(anmag);
    const int anmant = max((int)(g726_shl((uint32_t)anmag, c) >> 26), 32);
    const uint32_t p = (uint32_t)(anmant * t.m8 + 0x3000) & ~0xFFFu;
    const int r = (int)g726_shr(p, t.kk + c - 31) & 0x7FFF;
    const int s = (an ^ t.sg) >> 31;
    return (r ^ s) - s;
}

struct G726Core {
    int yl, yu, dms, dml, ap, a0, a1, b[6], pk0, pk1, td;
    G726Tap dq[6], sr[2];
    int sr_raw[2];  // sr can leave the 16-bit range, its log-format word is then not canonical: kept as the reference holds it
};

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

// One sample of either direction: encode_sample (:249-277: `in` = the 16-bit sample, returns the code) or decode_code
// (:279-303: `in` = the code, returns the sample), including update (:334-462).  Written without a data-dependent branch:
// every lane of a warp is a different stream, so a branch taken by one lane is paid by all 32 (the first version's loop
// executed 569 instructions per sample, most of them on partially active warps).  tab[code] = {dqln, wi << 5, fi, 0}.
template <int BITS, bool ENC>
__device__ __forceinline__ int g726_step(G726Core& s, int in, const int4* tab) {
    constexpr int MASK = (1 << BITS) - 1;
    // predictor_zero (:305) / predictor_pole (:313): the 8 products are independent of each other
    int sezi = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) sezi += g726_fmult(s.b[i] >> 2, s.dq[i]);
    const int pole = g726_fmult(s.a1 >> 2, s.sr[1]) + g726_fmult(s.a0 >> 2, s.sr[0]);
    const int sez = sezi >> 1, se = (sezi + pole) >> 1;
    // step_size (:317-332); the rounding term 0x3F applies to negative differences only and dif == 0 adds nothing
    const int yq = s.yl >> 6, dif = s.yu - yq;
    const int ymix = yq + ((dif * (s.ap >> 2) + ((dif >> 31) & 0x3F)) >> 6);
    const int y = s.ap >= 256 ? s.yu : ymix;
    int code = in;
    if (ENC) {  // quantize (:499-515)
        const int d = (in >> 2) - se;
        const int dqm = abs(d);
        const int e = g726_quan2(dqm >> 1);
        const int dln = (e << 7) + (((dqm << 7) >> e) & 0x7F) - (y >> 2);
        const int q = g726_qcount<BITS>(dln);
        code = d < 0 ? MASK - q : (q == 0 ? MASK : q);
    }
    const int4 row = tab[code];
    // reconstruct (:517-536): dq = ±mag in sign-magnitude (sign ? mag − 0x8000 : mag)
    const bool neg = (code & (1 << (BITS - 1))) != 0;
    const int dql = row.x + (y >> 2);
    const int mag = dql < 0 ? 0 : (int)g726_shr((uint32_t)(128 + (dql & 127)) << 7, 14 - ((dql >> 7) & 15));
    const int sr = neg ? se - (mag & (BITS == 5 ? 0x7FFF : 0x3FFF)) : se + mag;
    const int dqsez = sr - se + sez;
    // update (:334-462)
    const int pk0 = dqsez < 0 ? 1 : 0;
    const int ylint = s.yl >> 15, ylfrac = (s.yl >> 10) & 0x1F;
    const int thr2 = ylint > 9 ? (31 << 10) : ((32 + ylfrac) << ylint);
    const bool tr = s.td != 0 && mag > ((thr2 + (thr2 >> 1)) >> 1);
    const int yu = max(544, min(5120, y + ((row.y - y) >> 5)));
    s.yl += yu + ((-s.yl) >> 6);
    s.yu = yu;
    // pole coefficients: both saturating selects of the reference are clamps (fa1 >> 5 into [-256, 255]; a2p ± 0x80 into ±12288)
    const bool nz = dqsez != 0;
    const int pks1 = pk0 ^ s.pk0;
    int a2p = s.a1 - (s.a1 >> 7);
    {
        const int fa1 = pks1 ? s.a0 : -s.a0;
        const int t = a2p + max(-256, min(255, fa1 >> 5)) + ((pk0 ^ s.pk1) ? -0x80 : 0x80);
        a2p = nz ? max(-12288, min(12288, t)) : a2p;
    }
    const int a1ul = 15360 - a2p;
    const int a0 = max(-a1ul, min(a1ul, s.a0 - (s.a0 >> 8) + (nz ? (pks1 ? -192 : 192) : 0)));
    s.a1 = tr ? 0 : a2p;
    s.a0 = tr ? 0 : a0;
    // zero coefficients: ±128 by the sign agreement of dq with each delayed dq (none for a zero magnitude)
    const int gs = mag == 0 ? 0 : (neg ? -128 : 128);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const int b = s.b[i] - (s.b[i] >> (BITS == 5 ? 9 : 8)) + gs * s.dq[i].sg;
        s.b[i] = tr ? 0 : b;
    }
#pragma unroll
    for (int i = 5; i >= 1; i--) s.dq[i] = s.dq[i - 1];
    s.dq[0] = g726_tap_of(mag, neg);
    s.sr[1] = s.sr[0];
    s.sr_raw[1] = s.sr_raw[0];
    {   // sr in log format (:443-452), literally: sr outside ±32767 makes a mantissa that carries into the exponent field
        const int x = sr > 0 ? sr : -sr;
        const int e = g726_quan2(x);
        const int lf = (e << 6) + ((x << 6) >> e);
        s.sr_raw[0] = sr == 0 ? 0x20 : (sr > 0 ? lf : (sr > -32768 ? lf - 0x400 : -0x3E0));
        s.sr[0] = g726_tap(s.sr_raw[0]);
    }
    s.pk1 = s.pk0;
    s.pk0 = pk0;
    s.td = (!tr && a2p < -11776) ? 1 : 0;
    s.dms += (row.z - s.dms) >> 5;
    s.dml += ((row.z << 2) - s.dml) >> 7;
    const bool fast = y < 1536 || s.td != 0 || abs((s.dms << 2) - s.dml) >= (s.dml >> 3);
    s.ap = tr ? 256 : s.ap + (((fast ? 0x200 : 0) - s.ap) >> 4);
    return ENC ? code : max(-32768, min(32767, sr << 2));
}

// [rate][3][32] = dqln | wi | fi in constant memory (filled from G726Tab on first use); each CTA copies its rate's rows to
// shared memory as one 16-byte row per code {dqln, wi << 5, fi, 0} (one LDS.128 per sample; the index is a per-lane code,
// constant memory would serialise divergent indices)
__constant__ int g726_c_tab[4][96];
template <int BITS>
__device__ __forceinline__ void g726_load_tables(int4* tab) {
    for (int i = threadIdx.x; i < 32; i += blockDim.x)
        tab[i] = make_int4(g726_c_tab[BITS - 2][i], g726_c_tab[BITS - 2][32 + i] << 5, g726_c_tab[BITS - 2][64 + i], 0);
    __syncthreads();
}
__device__ __forceinline__ void g726_core_load(G726Core& c, const int32_t* state, size_t stream) {
    int r[G726_NSTATE];
    if (state) {
#pragma unroll
        for (int i = 0; i < G726_NSTATE; i++) r[i] = state[stream * G726_NSTATE + i];
    } else {  // Default for G726Core, :214-230
#pragma unroll
        for (int i = 0; i < G726_NSTATE; i++) r[i] = 0;
        r[G_YL] = 34816; r[G_YU] = 544;
#pragma unroll
        for (int i = 0; i < 6; i++) r[G_DQ0 + i] = 32;
        r[G_SR0] = 32; r[G_SR1] = 32;
    }
    c.yl = r[G_YL]; c.yu = r[G_YU]; c.dms = r[G_DMS]; c.dml = r[G_DML]; c.ap = r[G_AP]; c.a0 = r[G_A0]; c.a1 = r[G_A1];
    c.pk0 = r[G_PK0]; c.pk1 = r[G_PK1]; c.td = r[G_TD];
#pragma unroll
    for (int i = 0; i < 6; i++) { c.b[i] = r[G_B0 + i]; c.dq[i] = g726_tap(r[G_DQ0 + i]); }
    c.sr[0] = g726_tap(r[G_SR0]); c.sr[1] = g726_tap(r[G_SR1]);
    c.sr_raw[0] = r[G_SR0]; c.sr_raw[1] = r[G_SR1];
}
__device__ __forceinline__ void g726_core_store(const G726Core& c, int32_t* state, size_t stream) {
    if (!state) return;
    int32_t* r = state + stream * G726_NSTATE;
    r[G_YL] = c.yl; r[G_YU] = c.yu; r[G_DMS] = c.dms; r[G_DML] = c.dml; r[G_AP] = c.ap; r[G_A0] = c.a0; r[G_A1] = c.a1;
    r[G_PK0] = c.pk0; r[G_PK1] = c.pk1; r[G_TD] = c.td;
#pragma unroll
    for (int i = 0; i < 6; i++) { r[G_B0 + i] = c.b[i]; r[G_DQ0 + i] = g726_tap_raw(c.dq[i]); }
    r[G_SR0] = c.sr_raw[0]; r[G_SR1] = c.sr_raw[1];
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

// A block of 32 samples is exactly BITS 32-bit words of codes.  The words are held as one multi-word integer `w` (word 0 =
// the earliest codes; MSB-first packing byte-swaps on load so that the earliest code is the TOP of word 0): the 8 codes of a
// sub-block are its first 8·BITS bits, then the whole integer moves on by 8·BITS bits with compile-time funnel shifts.
template <int BITS, int PACKING, int U = 8>
__device__ __forceinline__ void g726_words_advance(uint32_t (&w)[BITS]) {
    constexpr int S = U * BITS, WS = S / 32, BS = S % 32;
#pragma unroll
    for (int i = 0; i < BITS; i++) {
        const uint32_t a = i + WS < BITS ? w[i + WS] : 0u, b = i + WS + 1 < BITS ? w[i + WS + 1] : 0u;
        if (BS == 0) w[i] = a;
        else w[i] = PACKING == 0 ? __funnelshift_l(b, a, BS) : __funnelshift_r(a, b, BS);
    }
}
template <int BITS, int PACKING>
__device__ __forceinline__ int g726_code_at(uint64_t cw, int j) {  // cw: first 64 bits of the words (word 0 first)
    return PACKING == 0 ? (int)(cw >> (64 - (j + 1) * BITS)) & ((1 << BITS) - 1) : (int)(cw >> (j * BITS)) & ((1 << BITS) - 1);
}

template <int KIND> __device__ __forceinline__ uint32_t g726_out_elem(int v) {  // G726Decoder::decode_{i16,i32,f32}, :821-951
    if (KIND == SK_KIND_I16) return (uint32_t)v & 0xffffu;
    if (KIND == SK_KIND_I32) return (uint32_t)v << 16;
    return __float_as_uint(__fmul_rn((float)v, 1.0f / 32768.0f));
}

// KIND: output element (sk_kind).  One thread per stream; `bytes` per stream is a whole number of code groups.
template <int BITS, int PACKING, int KIND>
__global__ void __launch_bounds__(kG726Threads) g726_decode_kernel(const uint8_t* __restrict__ in, size_t in_stride, size_t n_streams, size_t bytes,
                                                                   void* __restrict__ outv, size_t out_stride, int32_t* __restrict__ state) {
    __shared__ int4 tab[32];
    g726_load_tables<BITS>(tab);
    const size_t stream = (size_t)blockIdx.x * kG726Threads + threadIdx.x;
    if (stream >= n_streams) return;
    G726Core c;
    g726_core_load(c, state, stream);
    const uint8_t* p = in + stream * in_stride;
    const bool wordwise = ((uintptr_t)p & 3) == 0;
    const size_t n = bytes * 8 / BITS;
    constexpr int OB = KIND == SK_KIND_I16 ? 2 : 4;
    uint8_t* o = reinterpret_cast<uint8_t*>(outv) + stream * out_stride * OB;
    const bool vec = ((uintptr_t)o & 15) == 0;
#ifndef G726_DEC_SUB
#define G726_DEC_SUB 4
#endif
    // samples decoded per trip of the resident loop.  4, not 8: the 8-sample body (2 500 instructions, 40 KB) ran with 19 % of its stall
    // samples on no_instructions (instruction-cache misses); the 4-sample body decodes 69.5 instead of 59.5 G samples/s
    constexpr int U = G726_DEC_SUB;
    // whole 32-sample blocks: the next block's words are requested before the current block is decoded
    const size_t nblk = wordwise ? n / 32 : 0;
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(p);
    uint32_t nx[BITS];
    if (nblk) {
#pragma unroll
        for (int i = 0; i < BITS; i++) nx[i] = __ldg(pw + i);
    }
    for (size_t blk = 0; blk < nblk; blk++) {
        uint32_t w[BITS];
#pragma unroll
        for (int i = 0; i < BITS; i++) w[i] = PACKING == 0 ? bswap32(nx[i]) : nx[i];
        if (blk + 1 < nblk) {
#pragma unroll
            for (int i = 0; i < BITS; i++) nx[i] = __ldg(pw + (blk + 1) * BITS + i);
        }
        uint8_t* ob = o + blk * 32 * OB;
#pragma unroll 1
        for (int sb = 0; sb < 32 / U; sb++) {
            const uint64_t cw = PACKING == 0 ? ((uint64_t)w[0] << 32 | w[1]) : ((uint64_t)w[1] << 32 | w[0]);
            uint32_t e[U];
#pragma unroll
            for (int j = 0; j < U; j++) e[j] = g726_out_elem<KIND>(g726_step<BITS, false>(c, g726_code_at<BITS, PACKING>(cw, j), tab));
            if (OB == 2) {
                uint32_t q[U / 2];
#pragma unroll
                for (int j = 0; j < U / 2; j++) q[j] = e[2 * j] | e[2 * j + 1] << 16;
                if (vec) {
                    if (U == 8) st_stream_v4(ob + sb * 16, make_uint4(q[0], q[1], q[U / 2 - 2], q[U / 2 - 1]));
                    else st_stream_v2(ob + sb * 8, make_uint2(q[0], q[1]));
                } else {
#pragma unroll
                    for (int j = 0; j < U; j++) reinterpret_cast<uint16_t*>(ob)[sb * U + j] = (uint16_t)(q[j >> 1] >> (16 * (j & 1)));
                }
            } else if (vec) {
#pragma unroll
                for (int j = 0; j < U; j += 4) st_stream_v4(ob + sb * U * 4 + j * 4, make_uint4(e[j], e[j + 1], e[j + 2], e[j + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < U; j++) reinterpret_cast<uint32_t*>(ob)[sb * U + j] = e[j];
            }
            g726_words_advance<BITS, PACKING, U>(w);
        }
    }
    // the rest (fewer than 32 samples, or a stream that does not start on a word): a 64-bit bit reservoir, one sample at a time
    uint64_t res = 0;
    int navail = 0;
    size_t off = nblk * 4 * BITS;
#pragma unroll 1
    for (size_t k = nblk * 32; k < n; k++) {
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
        const uint32_t e = g726_out_elem<KIND>(g726_step<BITS, false>(c, code, tab));
        if (OB == 2) reinterpret_cast<uint16_t*>(o)[k] = (uint16_t)e;
        else reinterpret_cast<uint32_t*>(o)[k] = e;
    }
    g726_core_store(c, state, stream);
}

template <int KIND> __device__ __forceinline__ int g726_in_sample(const uint8_t* p, size_t k) {  // i32: (s >> 16) as i16, :735-738
    return KIND == SK_KIND_I16 ? (int)__ldg(reinterpret_cast<const int16_t*>(p) + k) : (int)(int16_t)(__ldg(reinterpret_cast<const int32_t*>(p) + k) >> 16);
}

// KIND: input element, i16 or i32.  `n` samples per stream, a whole number of groups.
template <int BITS, int PACKING, int KIND>
__global__ void __launch_bounds__(kG726Threads) g726_encode_kernel(const void* __restrict__ inv, size_t in_stride, size_t n_streams, size_t n,
                                                                   uint8_t* __restrict__ out, size_t out_stride, int32_t* __restrict__ state) {
    __shared__ int4 tab[32];
    g726_load_tables<BITS>(tab);
    const size_t stream = (size_t)blockIdx.x * kG726Threads + threadIdx.x;
    if (stream >= n_streams) return;
    G726Core c;
    g726_core_load(c, state, stream);
    constexpr int IB = KIND == SK_KIND_I16 ? 2 : 4;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(inv) + stream * in_stride * IB;
    uint8_t* o = out + stream * out_stride;
    const bool wordwise = ((uintptr_t)o & 3) == 0;
    const bool vec = ((uintptr_t)p & 15) == 0;
    const size_t bytes = n * BITS / 8;
    // whole 32-sample blocks = BITS output words, built as one multi-word integer (see g726_words_advance): each sub-block's
    // codes enter at the far end and the integer moves on, so after four sub-blocks the first codes sit in word 0
    const size_t nblk = wordwise ? n / 32 : 0;
#ifndef G726_ENC_SUB
#define G726_ENC_SUB 4
#endif
    constexpr int U = G726_ENC_SUB;      // samples encoded per trip of the resident loop
    constexpr int NW = U * IB / 4;       // input words per sub-block
    uint32_t nx[NW];
    auto fetch = [&](size_t k) {
        if (vec) {
            if constexpr (NW % 4 == 0) {
#pragma unroll
                for (int i = 0; i < NW / 4; i++) {
                    const uint4 v = ld_stream_v4(p + k * IB + i * 16);
                    nx[4 * i] = v.x; nx[4 * i + 1] = v.y; nx[4 * i + 2] = v.z; nx[4 * i + 3] = v.w;
                }
            } else {
                const uint2 v = ld_stream_v2(p + k * IB);
                nx[0] = v.x; nx[1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NW; j++) {
                if (IB == 2) nx[j] = (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(p) + k + 2 * j) | (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(p) + k + 2 * j + 1) << 16;
                else nx[j] = __ldg(reinterpret_cast<const uint32_t*>(p) + k + j);
            }
        }
    };
    if (nblk) fetch(0);
    for (size_t blk = 0; blk < nblk; blk++) {
        uint32_t w[BITS];
#pragma unroll
        for (int i = 0; i < BITS; i++) w[i] = 0;
#pragma unroll 1
        for (int sb = 0; sb < 32 / U; sb++) {
            uint32_t q[NW];
#pragma unroll
            for (int i = 0; i < NW; i++) q[i] = nx[i];
            const size_t knext = blk * 32 + (sb + 1) * U;
            if (knext + U <= nblk * 32) fetch(knext);
            uint64_t cw = 0;
#pragma unroll
            for (int j = 0; j < U; j++) {
                const int sample = IB == 2 ? (int)(int16_t)(q[j >> 1] >> (16 * (j & 1))) : (int)(int16_t)((int32_t)q[j] >> 16);
                const uint64_t code = (uint64_t)g726_step<BITS, true>(c, sample, tab);
                cw |= PACKING == 0 ? code << (U * BITS - (j + 1) * BITS) : code << (j * BITS);
            }
            g726_words_advance<BITS, PACKING, U>(w);
            constexpr int S = U * BITS, T = 32 * BITS - S;  // the sub-block occupies bits [T, T + S) counted from the first code
            if (PACKING == 0) {  // MSB first: the last S bits of the integer = the low end of the last words
                w[BITS - 1] |= (uint32_t)cw;
                if constexpr (S > 32) w[BITS - 2] |= (uint32_t)(cw >> 32);
            } else {
                w[T / 32] |= (uint32_t)(cw << (T % 32));
                if constexpr (T % 32 + S > 32) w[T / 32 + 1] |= (uint32_t)(cw >> (32 - T % 32));
            }
        }
#pragma unroll
        for (int i = 0; i < BITS; i++) reinterpret_cast<uint32_t*>(o)[blk * BITS + i] = PACKING == 0 ? bswap32(w[i]) : w[i];
    }
    uint64_t res = 0;
    int nbits = 0;
    size_t off = nblk * 4 * BITS;
#pragma unroll 1
    for (size_t k = nblk * 32; k < n; k++) {
        const uint64_t code = (uint64_t)g726_step<BITS, true>(c, g726_in_sample<KIND>(p, k), tab);
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

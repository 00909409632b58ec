// sk_resample_dev.cuh — device code and launchers of the integer-ratio resampler kernels, shared by sk_resample.cu and the
// per-step translation units sk_resample_int<STEP>.cu (the fully unrolled FIR is expensive to compile: one file per
// step lets `make -j` build them in parallel).
#pragma once
#include <math.h>

#include <algorithm>

#include "sk_common.cuh"

#define SINC_LEN 256
#define OVERSAMPLING 256
#define STREAM_CHUNK 4096  // RESAMPLE_CHUNK_SIZE, soundkit-decoder/src/lib.rs:79

static inline uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

struct sk_plan {
    size_t n_out = 0;
    bool integer_step = false;
    int step = 0;
    long long i0_first = 0;  // integer path: i0[n] = i0_first + step * n
    // general path
    int32_t* d_i0 = nullptr; uint8_t* d_s0 = nullptr; float* d_frac = nullptr;
    int period = 8;          // outputs per phase period used for the phase-major tile mapping
    int max_window = 0;      // largest input window (frames) any tile of 32*period outputs needs
    int period2 = 0;         // resample_gen2_kernel: outputs per period (a multiple of L), 0 = kernel not applicable
    int max_window2 = 0;     // largest input window of a 64-period tile
    int max_step3 = -1;      // general plans: largest i0[n + 3] - i0[n] (window advance inside a group of 4 consecutive outputs)
    int max_step7 = -1;      // ... and i0[n + 7] - i0[n] (groups of 8 outputs: the tolerance-mode kernel)
    uint64_t last_use = 0;   // plan cache clock value of the last lookup (least recently used plans are evicted first)
};

// ================================================================================================
// device: window staging (optional fused PCM conversion)
// ================================================================================================
struct SrcDesc {
    const uint8_t* base;      // stream 0, lane 0
    long long stream_stride;  // bytes between streams
    long long lane_stride;    // planar: bytes between channel planes; interleaved: unused
    int width;                // 0 = planar f32, 2/3/4 = interleaved sample bytes
    int channels;             // channels per stream
    int big_endian, is_float, scrub;
    int pair_ok;              // base and stride are 8-byte aligned: a stereo frame may be fetched with one 4 / 8-byte load
    int byte_loads;           // base or stride not aligned to the sample width: assemble 16 / 32-bit samples from bytes
    float scale;              // exact power of two for integer sources
    long long origin;         // frame index i lives at element i + origin
    long long valid_lo, valid_hi;  // frames outside [lo, hi) read as 0.0 (initial history / flush padding)
};

__device__ __forceinline__ float decode_sample_bits(uint32_t u, const SrcDesc& s) {
    // u holds the sample bytes assembled little-endian (after the optional byte swap)
    if (s.width == 2) return __fmul_rn((float)(int16_t)u, s.scale);
    if (s.width == 3) return __fmul_rn((float)(((int32_t)(u << 8)) >> 8), s.scale);
    if (s.is_float) {
        float f = __uint_as_float(u);
        if (s.scrub && (u & 0x7f800000u) == 0x7f800000u) f = 0.0f;
        return f;
    }
    return __fmul_rn(__int2float_rn((int32_t)u), s.scale);
}

// sample of channel `ch` of frame `f` (absolute frame index, already range checked) of stream `sbase`
__device__ __forceinline__ uint32_t load_sample_bits(const uint8_t* sbase, long long f, int ch, const SrcDesc& s) {
    const uint8_t* p = sbase + ((f + s.origin) * s.channels + ch) * (long long)s.width;
    uint32_t u;
    if (s.byte_loads) {  // a caller's slice may start at any byte
        u = 0;
        for (int b = 0; b < s.width; b++) u |= (uint32_t)p[b] << (8 * (s.big_endian ? s.width - 1 - b : b));
        return u;
    }
    if (s.width == 2) { u = *reinterpret_cast<const uint16_t*>(p); if (s.big_endian) u = ((u & 0xff) << 8) | (u >> 8); }
    else if (s.width == 3) {
        uint32_t b0 = p[0], b1 = p[1], b2 = p[2];
        u = s.big_endian ? (b2 | (b1 << 8) | (b0 << 16)) : (b0 | (b1 << 8) | (b2 << 16));
    } else { u = *reinterpret_cast<const uint32_t*>(p); if (s.big_endian) u = bswap32(u); }
    return u;
}

// Fill win0/win1[0..count) with frames [w0, w0+count) of lanes (lane0, lane0+1) of one stream.
// Thread-per-frame: consecutive threads touch consecutive frames, so global reads coalesce into
// whole sectors and the shared-memory stores are conflict free.  Returns (block-wide OR'ed later)
// whether a value with |x| >= 2^113 (or non-finite) was staged — only possible for float sources.
template <int NTHREADS>
__device__ __forceinline__ bool stage_window(const SrcDesc& s, long long stream, int lane0, bool lane1_valid, long long w0, int count,
                                             float* win0, float* win1) {
    bool big = false;
    if (s.width == 0) {
        // planar f32 (the stateful resamplers' device rows): 4-byte cp.async per sample — any alignment, every copy of the window in
        // flight at once, no register round trip — then the fetching thread checks its own elements for huge / non-finite values
        const float* p0 = reinterpret_cast<const float*>(s.base + stream * s.stream_stride + (long long)lane0 * s.lane_stride);
        const float* p1 = reinterpret_cast<const float*>(s.base + stream * s.stream_stride + (long long)(lane0 + 1) * s.lane_stride);
        for (int i = threadIdx.x; i < count; i += NTHREADS) {
            const long long f = w0 + i;
            const bool ok = f >= s.valid_lo && f < s.valid_hi;
            if (ok) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(win0 + i)), "l"(p0 + f + s.origin) : "memory");
            else win0[i] = 0.0f;
            if (win1) {
                if (ok && lane1_valid) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(win1 + i)), "l"(p1 + f + s.origin) : "memory");
                else win1[i] = 0.0f;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        for (int i = threadIdx.x; i < count; i += NTHREADS) {
            big |= (__float_as_uint(win0[i]) & 0x7f800000u) >= 0x78000000u;
            if (win1) big |= (__float_as_uint(win1[i]) & 0x7f800000u) >= 0x78000000u;
        }
    } else {
        const uint8_t* sbase = s.base + stream * s.stream_stride;
        const bool stereo_pair = (s.channels == 2) && lane1_valid && lane0 == 0 && s.pair_ok;
        for (int i = threadIdx.x; i < count; i += NTHREADS) {
            const long long f = w0 + i;
            float a = 0.0f, b = 0.0f;
            if (f >= s.valid_lo && f < s.valid_hi) {
                if (stereo_pair && s.width == 2) {
                    uint32_t u = *reinterpret_cast<const uint32_t*>(sbase + (f + s.origin) * 4);
                    uint32_t l = u & 0xffffu, r = u >> 16;
                    if (s.big_endian) { l = ((l & 0xff) << 8) | (l >> 8); r = ((r & 0xff) << 8) | (r >> 8); }
                    a = decode_sample_bits(l, s); b = decode_sample_bits(r, s);
                } else if (stereo_pair && s.width == 3) {
                    const uint16_t* p = reinterpret_cast<const uint16_t*>(sbase + (f + s.origin) * 6);
                    uint32_t h0 = p[0], h1 = p[1], h2 = p[2];
                    uint32_t l, r;
                    if (!s.big_endian) { l = h0 | ((h1 & 0xff) << 16); r = (h1 >> 8) | (h2 << 8); }
                    else { l = ((h0 & 0xff) << 16) | (h0 & 0xff00) | (h1 & 0xff); r = ((h1 >> 8) << 16) | ((h2 & 0xff) << 8) | (h2 >> 8); }
                    a = decode_sample_bits(l, s); b = decode_sample_bits(r, s);
                } else if (stereo_pair && s.width == 4) {
                    uint2 u = *reinterpret_cast<const uint2*>(sbase + (f + s.origin) * 8);
                    if (s.big_endian) { u.x = bswap32(u.x); u.y = bswap32(u.y); }
                    a = decode_sample_bits(u.x, s); b = decode_sample_bits(u.y, s);
                } else {
                    a = decode_sample_bits(load_sample_bits(sbase, f, lane0, s), s);
                    if (lane1_valid) b = decode_sample_bits(load_sample_bits(sbase, f, lane0 + 1, s), s);
                }
                if (s.is_float) big |= ((__float_as_uint(a) & 0x7f800000u) >= 0x78000000u) | ((__float_as_uint(b) & 0x7f800000u) >= 0x78000000u);
            }
            win0[i] = a; if (win1) win1[i] = b;
        }
    }
    return big;
}

// rubato interpolator_avx.rs reduction: (hi128 + lo128) then two hadds
__device__ __forceinline__ float reduce8(const float* a) {
    return __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[4]), __fadd_rn(a[1], a[5])), __fadd_rn(__fadd_rn(a[2], a[6]), __fadd_rn(a[3], a[7])));
}

// generic 256-tap dot product in rubato's lane order; x in shared memory (any alignment), h 16-byte aligned global/shared
__device__ __forceinline__ float dot256(const float* x, const float* __restrict__ h) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < SINC_LEN; k += 8) {
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(h + k));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(h + k + 4));
        a[0] = __fmaf_rn(x[k + 0], h0.x, a[0]); a[1] = __fmaf_rn(x[k + 1], h0.y, a[1]);
        a[2] = __fmaf_rn(x[k + 2], h0.z, a[2]); a[3] = __fmaf_rn(x[k + 3], h0.w, a[3]);
        a[4] = __fmaf_rn(x[k + 4], h1.x, a[4]); a[5] = __fmaf_rn(x[k + 5], h1.y, a[5]);
        a[6] = __fmaf_rn(x[k + 6], h1.z, a[6]); a[7] = __fmaf_rn(x[k + 7], h1.w, a[7]);
    }
    return reduce8(a);
}

struct OutDesc {
    float* base;
    long long stream_stride;  // floats between streams
    long long lane_stride;    // floats between channel planes (mix none)
    int mix;                  // sk_mix_mode; != NONE only with exactly two lanes per stream
    long long n_out;
    // byte epilogue (sk_normalize_batch_pcm_dev): the mono result leaves as f32_channels_to_bytes would write it
    int pcm_bits = 0;         // 0 = planar f32 (the fields above), else 16 / 24 / 32
    int pcm_float = 0;        // PCMFloat (32-bit only)
    uint8_t* pcm_base = nullptr;
    long long pcm_stream_stride = 0;  // bytes between streams
};

// one mono sample as f32_channels_to_bytes writes it (soundkit-decoder/src/lib.rs:3619-3683): little-endian bits in the low bytes
__device__ __forceinline__ uint32_t pcm_quantize(float x, int bits, int is_float) {
    if (is_float) return __float_as_uint(x);
    if (bits == 16) return (uint32_t)float_sample_to_i16_dev(x) & 0xffffu;
    if (bits == 24) return (uint32_t)f32_to_s24_dev(x) & 0xffffffu;
    return (uint32_t)f32_to_i32_dev(x);
}

// ================================================================================================
// integer-step kernel
// ================================================================================================
// One table row in two layouts (kernel parameter -> constant bank).  The packed FMA consumes taps in
// adjacent pairs (j, j+1): `he` serves pairs that start at an even tap; `ho[i] = h[i-1]` (zero outside)
// serves pairs that start at an odd tap, including the edge pairs (-1, 0) and (255, 256) whose outer
// element is a zero coefficient.
struct __align__(16) Taps { float he[SINC_LEN]; float ho[SINC_LEN + 8]; };

// d += a * b on two packed f32 lanes (sm_100 FFMA2): both lanes are IEEE fused multiply-adds
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int STEP> struct IntCfg;
// R outputs per thread and the bank skew between the two lanes' windows, chosen so that the 8
// threads of an LDS.128 quarter-warp (4 thread-blocks x 2 lanes) hit 8 distinct 4-bank groups:
// thread stride S = STEP*R words; need {S*b mod 32, b<4} U {S*b + SKEW mod 32} to be 8 distinct
// multiples of 4.  CTAS = resident CTAs per SM of the persistent TMA kernel (register budget).
// Step 3 (48 -> 16 kHz, the headline) runs 12 outputs per thread at 2 CTAs/SM: one LDCU.128 then feeds 12 FFMA2 instead of 8
// and the input span per output shrinks, which is worth more than the third CTA (57.8 -> 60.1 G frames/s on B200).
template <> struct IntCfg<3>  { static constexpr int R = 12, SKEW = 16, CTAS = 2; };  // S = 36: {0,4,8,12} + 16
template <> struct IntCfg<2>  { static constexpr int R = 6, SKEW = 16, CTAS = 3; };   // S = 12: {0,12,24,4} + 16 (R = 12 / 2 CTAs measured: 58.3 vs 59.2 TFLOP/s)
template <> struct IntCfg<4>  { static constexpr int R = 7, SKEW = 16, CTAS = 3; };   // S = 28: {0,28,24,20} + 16
template <> struct IntCfg<6>  { static constexpr int R = 6, SKEW = 16, CTAS = 3; };   // S = 36: {0,4,8,12} + 16
template <> struct IntCfg<12> { static constexpr int R = 7, SKEW = 16, CTAS = 3; };  // S = 84: {0,20,8,28} + 16

constexpr int kIntThreads = 256;                 // 128 thread-blocks x 2 lanes
constexpr int kIntBlocksPerLane = kIntThreads / 2;
// integer steps whose kernels are also built with the byte epilogue (48 -> 24 / 16 / 8 kHz and friends); the others reach
// sk_normalize_batch_pcm_dev through the f32 scratch
template <int STEP> __host__ __device__ constexpr bool int_step_has_pcm() { return STEP == 2 || STEP == 3 || STEP == 6; }

template <int STEP> __host__ __device__ constexpr int int_tile_outputs() { return kIntBlocksPerLane * IntCfg<STEP>::R; }
template <int STEP> __host__ __device__ constexpr int int_window() { return STEP * (int_tile_outputs<STEP>() - 1) + SINC_LEN; }
template <int STEP> __host__ __device__ constexpr int int_lane_stride() { return ((int_window<STEP>() + 8 + 31) / 32) * 32 + IntCfg<STEP>::SKEW; }
template <int STEP> __host__ __device__ constexpr int int_smem_bytes() { return 2 * int_lane_stride<STEP>() * (int)sizeof(float); }

// The FIR core: walk the thread's input span once.  Every LDS.128 yields two packed pairs (x[m], x[m+1]),
// m even, and each pair feeds one FFMA2 per output r: taps (j, j+1) with j = m - LEAD - STEP*r go to the
// accumulator pair of lanes (j & 7, (j+1) & 7).  Lane l of output r therefore still receives exactly its
// taps l, l+8, ... in increasing order with fused multiply-adds — rubato's AVX accumulation — while the
// instruction count is half of the scalar form (R*128 FFMA2 instead of R*256 FFMA) and the coefficient
// pair is a uniform-register operand fed from the constant bank (kernel parameter).
// Outputs whose j is odd for even m use the shifted table `ho` and pair lanes (1,2)(3,4)(5,6)(7,0); their
// first / last pair carries one zero coefficient (x * 0 = +-0 leaves a non-negative-zero accumulator
// unchanged for finite x; tiles that stage non-finite or huge floats take the exact slow path instead).
template <int STEP, int R, int LEAD>
__device__ __forceinline__ void fir_core2(const float* __restrict__ w, const Taps& taps, float (&lanes)[R][8]) {
    constexpr int NIN = ((LEAD + STEP * (R - 1) + SINC_LEN + 1 + 3) / 4) * 4;
    unsigned long long acc[R][4];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int p = 0; p < 4; p++) acc[r][p] = 0ull;
#pragma unroll
    for (int v = 0; v < NIN / 4; v++) {
        unsigned long long x[2];
        asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x[0]), "=l"(x[1]) : "r"((unsigned)__cvta_generic_to_shared(w + 4 * v)));
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int m = 4 * v + 2 * e;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int j = m - LEAD - STEP * r;  // tap of the pair's low element
                const bool odd = ((LEAD + STEP * r) & 1) != 0;
                if (!odd) {
                    if (j >= 0 && j + 1 < SINC_LEN) ffma2(acc[r][(j & 7) >> 1], x[e], *reinterpret_cast<const unsigned long long*>(&taps.he[j]));
                } else {
                    if (j >= -1 && j <= SINC_LEN - 1) ffma2(acc[r][(j & 7) >> 1], x[e], *reinterpret_cast<const unsigned long long*>(&taps.ho[j + 1]));
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const bool odd = ((LEAD + STEP * r) & 1) != 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const float lo = __uint_as_float((unsigned)(acc[r][p] & 0xffffffffull)), hi = __uint_as_float((unsigned)(acc[r][p] >> 32));
            if (!odd) { lanes[r][2 * p] = lo; lanes[r][2 * p + 1] = hi; }
            else { lanes[r][2 * p + 1] = lo; lanes[r][(2 * p + 2) & 7] = hi; }
        }
    }
}

// store R outputs of one thread: optional (L, R) -> mono mix through a lane-pair shuffle, 128-bit stores
template <int R, bool PCM>  // PCM: the mono result leaves as bytes (compile-time, so the f32 path carries none of the quantiser code)
__device__ __forceinline__ void int_epilogue(const float (&y)[R], const OutDesc& dst, long long stream, long long n0, int b, int lane, int lane0,
                                             bool lane1_valid) {
    const long long nb = n0 + (long long)R * b;  // first output of this thread
    if (dst.mix != SK_MIX_NONE) {
        // lanes are (L, R) of one stream in adjacent threads: exchange, then each thread stores half
        // lane 0 (L) finishes outputs [0, H), lane 1 (R) outputs [H, R): each sends the other half across
        constexpr int H = (R + 1) / 2;
        float o[H];
#pragma unroll
        for (int q = 0; q < H; q++) {
            const int mine = q < R - H ? q + H : R - 1;  // index lane 0 gives away (clamped when R is odd)
            const float give = lane == 0 ? y[mine] : y[q];
            const float got = __shfl_xor_sync(0xffffffffu, give, 1);
            const float keep = lane == 0 ? y[q] : y[mine];
            const float l = lane == 0 ? keep : got, rr = lane == 0 ? got : keep;
            if (dst.mix == SK_MIX_SCALE_THEN_SUM) o[q] = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(l, 0.5f)), __fmul_rn(rr, 0.5f));
            else o[q] = __fmul_rn(__fadd_rn(__fadd_rn(0.0f, l), rr), 0.5f);
        }
        const long long no = nb + (long long)H * lane;  // first output this thread stores
        const int cnt = lane == 0 ? H : R - H;
        if (PCM) {  // bytes out: quantise the mono samples in registers, 8 / 12 / 16 bytes per thread when aligned
            uint8_t* bp = dst.pcm_base + stream * dst.pcm_stream_stride;
            const int ob = dst.pcm_bits >> 3;
            if (R == 8 && nb + 8 <= dst.n_out && ((dst.pcm_stream_stride & 15) == 0) && (((uintptr_t)dst.pcm_base) & 15) == 0) {
                uint32_t q4[4];
#pragma unroll
                for (int q = 0; q < 4; q++) q4[q] = pcm_quantize(o[q], dst.pcm_bits, dst.pcm_float);
                if (ob == 2) *reinterpret_cast<uint2*>(bp + no * 2) = make_uint2(q4[0] | (q4[1] << 16), q4[2] | (q4[3] << 16));
                else if (ob == 3) {
                    uint32_t* w = reinterpret_cast<uint32_t*>(bp + no * 3);
                    w[0] = q4[0] | (q4[1] << 24); w[1] = (q4[1] >> 8) | (q4[2] << 16); w[2] = (q4[2] >> 16) | (q4[3] << 8);
                } else *reinterpret_cast<uint4*>(bp + no * 4) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
            } else if (R == 12 && nb + 12 <= dst.n_out && ((dst.pcm_stream_stride & 7) == 0) && (((uintptr_t)dst.pcm_base) & 7) == 0) {
                uint32_t q6[6];  // 6 samples per thread: 12 bytes (4-byte aligned), 18 bytes (2-byte aligned) or 24 bytes (8-byte aligned)
#pragma unroll
                for (int q = 0; q < 6; q++) q6[q] = pcm_quantize(o[q % H], dst.pcm_bits, dst.pcm_float);
                if (ob == 2) {
                    uint32_t* w = reinterpret_cast<uint32_t*>(bp + no * 2);
#pragma unroll
                    for (int v = 0; v < 3; v++) w[v] = q6[2 * v] | (q6[2 * v + 1] << 16);
                } else if (ob == 3) {
                    uint16_t* h = reinterpret_cast<uint16_t*>(bp + no * 3);
#pragma unroll
                    for (int v = 0; v < 3; v++) {
                        const uint32_t a = q6[2 * v], c = q6[2 * v + 1];
                        h[3 * v] = (uint16_t)a; h[3 * v + 1] = (uint16_t)((a >> 16) | (c << 8)); h[3 * v + 2] = (uint16_t)(c >> 8);
                    }
                } else {
                    uint2* w = reinterpret_cast<uint2*>(bp + no * 4);
#pragma unroll
                    for (int v = 0; v < 3; v++) w[v] = make_uint2(q6[2 * v], q6[2 * v + 1]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < H; q++)
                    if (q < cnt && no + q < dst.n_out) {
                        const uint32_t v = pcm_quantize(o[q], dst.pcm_bits, dst.pcm_float);
                        for (int b = 0; b < ob; b++) bp[(no + q) * ob + b] = (uint8_t)(v >> (8 * b));
                    }
            }
            return;
        }
        float* op = dst.base + stream * dst.stream_stride;
        if (R == 8 && nb + 8 <= dst.n_out && ((dst.stream_stride & 3) == 0)) {
            *reinterpret_cast<float4*>(op + no) = make_float4(o[0], o[1], o[2], o[3]);
        } else if (R == 12 && nb + 12 <= dst.n_out && ((dst.stream_stride & 1) == 0) && (((uintptr_t)dst.base) & 7) == 0) {
#pragma unroll
            for (int v = 0; v < 3; v++) *reinterpret_cast<float2*>(op + no + 2 * v) = make_float2(o[(2 * v) % H], o[(2 * v + 1) % H]);
        } else {
#pragma unroll
            for (int q = 0; q < H; q++)
                if (q < cnt && no + q < dst.n_out) op[no + q] = o[q];
        }
    } else {
        if (lane == 1 && !lane1_valid) return;
        float* op = dst.base + stream * dst.stream_stride + (long long)(lane0 + lane) * dst.lane_stride;
        if (R % 4 == 0 && nb + R <= dst.n_out && ((dst.stream_stride | dst.lane_stride) & 3) == 0) {
#pragma unroll
            for (int v = 0; v < R / 4; v++) *reinterpret_cast<float4*>(op + nb + 4 * v) = make_float4(y[4 * v], y[4 * v + 1], y[4 * v + 2], y[4 * v + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (nb + r < dst.n_out) op[nb + r] = y[r];
        }
    }
}

template <int STEP, bool PCM>
__global__ void __launch_bounds__(kIntThreads, 2)
resample_int_kernel(const __grid_constant__ Taps taps, const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst,
                    long long i0_first, int lanes_per_stream, const float* __restrict__ row1) {
    constexpr int R = IntCfg<STEP>::R;
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int W = int_window<STEP>();
    constexpr int WS = int_lane_stride<STEP>();
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_big;

    // blockIdx.y enumerates lane pairs: pair p of stream s covers lanes 2p, 2p+1
    const int pairs_per_stream = (lanes_per_stream + 1) >> 1;
    const long long stream = blockIdx.y / pairs_per_stream;
    const int lane0 = (blockIdx.y % pairs_per_stream) * 2;
    const bool lane1_valid = lane0 + 1 < lanes_per_stream;
    const long long n0 = (long long)blockIdx.x * T;
    const long long w0 = i0_first + (long long)STEP * n0;

    if (threadIdx.x == 0) s_big = 0;
    __syncthreads();
    // W + 4: when the tile's last output has an odd window offset its final FFMA2 pairs tap 255 with a zero coefficient, and
    // that zero multiplies the word one past the window — it must be staged data (finite), not whatever shared memory held
    const bool big = stage_window<kIntThreads>(src, stream, lane0, lane1_valid, w0, W + 4, smem, smem + WS);
    if (big) s_big = 1;
    __syncthreads();

    const int lane = threadIdx.x & 1, b = threadIdx.x >> 1;
    const float* w = smem + lane * WS + STEP * R * b;
    float acc[R][8];
    fir_core2<STEP, R, 0>(w, taps, acc);

    float y[R];
#pragma unroll
    for (int r = 0; r < R; r++) y[r] = reduce8(acc[r]);
    if (s_big) {
        // some |x| >= 2^113 (or non-finite) in this window: evaluate the reference's full expression
        // p0 + 0.0 * (p1 - p0) with plain dot products (subindex 0 and 1, same input index)
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float p0 = dot256(w + STEP * r, row1 - SINC_LEN), p1 = dot256(w + STEP * r, row1);
            y[r] = __fadd_rn(p0, __fmul_rn(0.0f, __fsub_rn(p1, p0)));
        }
    }
    int_epilogue<R, PCM>(y, dst, stream, n0, b, lane, lane0, lane1_valid);
}

// ================================================================================================
// integer-step kernel v2: persistent CTAs, TMA-staged raw windows, interleaved stereo sources
// ================================================================================================
// One CTA per SM slot walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The raw interleaved bytes
// of tile k+2 are fetched by the TMA engine (cp.async.bulk -> mbarrier) while tile k is converted and
// convolved, so HBM latency never sits in front of the FMA pipe, and staging costs ~1 instruction per
// 16 bytes instead of a scalar load / convert / store chain per frame.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

// window of tile geometry for the TMA kernel: the shared window starts LEAD frames before the first
// tap so that (a) its first raw byte is 16-byte aligned in the stream and (b) every thread's LDS.128
// base stays 16-byte aligned; the taps are shifted by LEAD at compile time instead.
template <int STEP> __host__ __device__ constexpr int tma_lead() { return STEP % 8; }
template <int STEP> __host__ __device__ constexpr int tma_frames() { return ((tma_lead<STEP>() + int_window<STEP>() + 7) / 8) * 8; }
template <int STEP> __host__ __device__ constexpr int tma_lane_stride() { return ((tma_frames<STEP>() + 8 + 31) / 32) * 32 + IntCfg<STEP>::SKEW; }
template <int STEP, int SRCW> __host__ __device__ constexpr int tma_raw_bytes() { return tma_frames<STEP>() * 2 * SRCW; }
// Raw-byte buffers of the TMA staging.  Step 3 (the headline shape: 2 CTAs per SM by registers) requests tile k + 2 while tile k
// is convolved, into a second buffer.  The other steps are limited by shared memory instead (their windows are STEP / 3 times
// longer per output): ONE buffer, refilled for tile k + 1 as soon as tile k has been converted — the request still has the whole
// FIR of tile k to arrive — brings a third CTA per SM back.  The same holds for 32-bit sources at step 3: two 39 KB raw buffers beside
// the 39 KB of windows left ONE CTA per SM (s32 / f32 stereo ran at 48.9 instead of 59 TFLOP/s).  SK_INT_NRAW overrides for A/B builds.
#ifdef SK_INT_NRAW
template <int STEP, int SRCW> __host__ __device__ constexpr int tma_raw_buffers() { return SK_INT_NRAW; }
#else
template <int STEP, int SRCW> __host__ __device__ constexpr int tma_raw_buffers() { return STEP == 3 && SRCW <= 3 ? 2 : 1; }
#endif
template <int STEP, int SRCW> __host__ __device__ constexpr int tma_smem_bytes() {
    return 2 * tma_lane_stride<STEP>() * (int)sizeof(float) + tma_raw_buffers<STEP, SRCW>() * tma_raw_bytes<STEP, SRCW>() + 64;
}

// 4 interleaved stereo frames (FB words in memory order) -> 4 L and 4 R f32 samples.  Everything about
// the format is a template parameter: no per-sample branches.  Integer samples are staged UNSCALED (as the integer
// value): the exact power-of-two scale is folded into the taps on the host instead
// (fma(x * 2^-k, h, a) == fma(x, h * 2^-k, a) bit for bit), which removes one FMUL per sample from the staging pass.
template <int SRCW, bool BE, bool ISFLOAT, bool SCRUB>
__device__ __forceinline__ void convert4(const uint32_t* w, float* l, float* r) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float f;
            if (SRCW == 2) {
                uint32_t h = c == 0 ? (w[k] & 0xffffu) : (w[k] >> 16);
                if (BE) h = __byte_perm(h, 0, 0x3201) & 0xffffu;
                f = (float)(int16_t)h;
            } else if (SRCW == 3) {
                const int o = 3 * (2 * k + c), wi = o >> 2, sh = o & 3;
                const uint32_t lo = w[wi], hi = (sh >= 2) ? w[wi + 1] : 0u;
                // one PRMT per sample: result bytes 0..2 are the sample's bytes, least significant first, and byte 3 replicates the
                // sign of its top byte (selector nibble | 8), i.e. the sign-extended 24-bit integer
                const uint32_t top = !BE ? (uint32_t)(sh + 2) : (uint32_t)sh, low = !BE ? (uint32_t)sh : (uint32_t)(sh + 2);
                const uint32_t sel = ((top | 8u) << 12) | (top << 8) | ((uint32_t)(sh + 1) << 4) | low;
                uint32_t v;  // prmt.b32 directly: the __byte_perm intrinsic documents only 3 selector bits per nibble
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(lo), "r"(hi), "r"(sel));
                f = __int2float_rn((int32_t)v);
            } else {
                uint32_t u = w[2 * k + c];
                if (BE) u = bswap32(u);
                if (ISFLOAT) f = (SCRUB && (u & 0x7f800000u) == 0x7f800000u) ? 0.0f : __uint_as_float(u);
                else f = __int2float_rn((int32_t)u);
            }
            if (c == 0) l[k] = f; else r[k] = f;
        }
    }
}

// raw bytes of one tile (shared memory, filled by the TMA) -> two f32 windows.  Groups [g_lo, g_hi) of 4
// frames are completely inside the copied range; the few others (stream start / end) go frame by frame
// through global memory with range checks.  Returns whether a huge / non-finite float was staged.
template <int SRCW, int WF, int WS, bool BE, bool ISFLOAT, bool SCRUB>
__device__ __forceinline__ bool convert_window(const uint8_t* __restrict__ rb, float* __restrict__ win, const SrcDesc& src, long long stream,
                                               long long wa, int g_lo, int g_hi) {
    constexpr int FB = 2 * SRCW;
    bool big = false;
    for (int g = threadIdx.x; g < WF / 4; g += kIntThreads) {
        float l[4], r[4];
        if (g >= g_lo && g < g_hi) {
            uint32_t w[FB];
            if (SRCW == 3) {
#pragma unroll
                for (int v = 0; v < 3; v++) {
                    const uint2 q = *reinterpret_cast<const uint2*>(rb + g * 24 + 8 * v);
                    w[2 * v] = q.x; w[2 * v + 1] = q.y;
                }
            } else {
#pragma unroll
                for (int v = 0; v < FB / 4; v++) {
                    const uint4 q = *reinterpret_cast<const uint4*>(rb + g * (4 * FB) + 16 * v);
                    w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
                }
            }
            convert4<SRCW, BE, ISFLOAT, SCRUB>(w, l, r);
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const long long ff = wa + 4 * g + q;
                l[q] = 0.0f; r[q] = 0.0f;
                if (ff >= src.valid_lo && ff < src.valid_hi) {
                    const uint8_t* sbase = src.base + stream * src.stream_stride;
                    l[q] = decode_sample_bits(load_sample_bits(sbase, ff, 0, src), src);
                    r[q] = decode_sample_bits(load_sample_bits(sbase, ff, 1, src), src);
                }
            }
        }
        if (ISFLOAT) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                big |= ((__float_as_uint(l[q]) & 0x7f800000u) >= 0x78000000u) | ((__float_as_uint(r[q]) & 0x7f800000u) >= 0x78000000u);
        }
        *reinterpret_cast<float4*>(win + 4 * g) = make_float4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<float4*>(win + WS + 4 * g) = make_float4(r[0], r[1], r[2], r[3]);
    }
    return big;
}

#ifdef SK_TRACE_CTAS
__device__ unsigned long long sk_cta_trace[3 * 1024];
#endif
// per-tile geometry, computed once by the TMA-issuing thread and handed to the CTA through shared memory
struct TileGeom { int stream; int n0; int wa; int g_lo; int g_hi; int pad[3]; };


template <int STEP, int SRCW, bool PCM>
__global__ void __launch_bounds__(kIntThreads, IntCfg<STEP>::CTAS)
resample_int_tma_kernel(const __grid_constant__ Taps taps, const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, long long i0_first,
                        unsigned tiles_per_stream, unsigned total_tiles, const float* __restrict__ row1, unsigned* __restrict__ ticket) {
    constexpr int R = IntCfg<STEP>::R;
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int LEAD = tma_lead<STEP>();
    constexpr int WF = tma_frames<STEP>();
    constexpr int WS = tma_lane_stride<STEP>();
    constexpr int FB = 2 * SRCW;  // bytes per stereo frame
    constexpr int RAWB = tma_raw_bytes<STEP, SRCW>();
    constexpr int NB = tma_raw_buffers<STEP, SRCW>();  // raw buffers == tiles requested ahead
    // the word one past the tile's window is only read (against a zero coefficient) when the last output's offset is odd; the
    // converted region must then reach it
    static_assert(((LEAD + STEP * (R - 1)) & 1) == 0 || WF > LEAD + int_window<STEP>(), "stage one more frame for this STEP / R");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* win = reinterpret_cast<float*>(smem_raw);
    uint8_t* raw = smem_raw + 2 * WS * sizeof(float);
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + NB * RAWB);
    __shared__ TileGeom geoms[4];

    // Tiles are handed out by ticket (one atomic per tile), not by a static stride: the three co-resident CTAs of an SM do not
    // progress at the same rate (measured with %globaltimer: with equal static shares the first CTA finished after 57 % of the
    // kernel and its SM ran under-occupied from then on), so every CTA keeps drawing tiles until the batch is exhausted.

#ifdef SK_TRACE_CTAS  // development aid (tools/cta_trace.py): start time, SM and end time of every CTA
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned long long t; unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        sk_cta_trace[3 * blockIdx.x] = t; sk_cta_trace[3 * blockIdx.x + 1] = sm;
    }
#endif
    // one elected thread: ticket and geometry of the CTA's k-th tile -> geoms[k & 3] (stream = -1: no tile left), then the TMA
    // request for its raw bytes
    auto issue = [&](unsigned k) {
        const unsigned t = atomicAdd(ticket, 1u);
        if (t >= total_tiles) { geoms[k & 3].stream = -1; return; }
        const unsigned stream = t / tiles_per_stream;
        const long long n0 = (long long)(t - stream * tiles_per_stream) * T;
        const long long wa = i0_first + (long long)STEP * n0 - LEAD;                  // first frame of the shared window
        const long long stream_bytes16 = ((src.valid_hi * FB) / 16) * 16;             // bytes of a stream the TMA may touch
        long long b_lo = wa * FB; if (b_lo < 0) b_lo = 0;
        long long b_hi = (wa + WF) * FB; if (b_hi > stream_bytes16) b_hi = stream_bytes16;
        if (b_hi < b_lo) b_hi = b_lo;
        TileGeom g;
        g.stream = (int)stream; g.n0 = (int)n0; g.wa = (int)wa;
        g.g_lo = (int)((b_lo - wa * FB + 4 * FB - 1) / (4 * FB));  // groups of 4 frames fully inside the copied bytes
        g.g_hi = (int)((b_hi - wa * FB) / (4 * FB));
        geoms[k & 3] = g;
        const int buf = (int)(k % NB);
        const uint32_t bar = smem_u32(&bars[buf]);
        const uint32_t bytes = (uint32_t)(b_hi - b_lo);
        mbar_expect_tx(bar, bytes);
        if (bytes) tma_load_1d(smem_u32(raw + buf * RAWB + (b_lo - wa * FB)), src.base + (long long)stream * src.stream_stride + b_lo, bytes, bar);
    };

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_fence_init();
        issue(0);
        if (NB == 2) issue(1);
    }
    __syncthreads();

    const int lane = threadIdx.x & 1, b = threadIdx.x >> 1;
    for (unsigned k = 0;; k++) {
        const TileGeom g = geoms[k & 3];
        if (g.stream < 0) break;  // tickets are monotonic: nothing was requested for this or any later k
        const int buf = (int)(k % NB);
        mbar_wait(smem_u32(&bars[buf]), (k / NB) & 1);

        const uint8_t* rb = raw + buf * RAWB;
        bool big;
        if (SRCW == 4 && src.is_float) {
            if (src.scrub) big = src.big_endian ? convert_window<SRCW, WF, WS, true, true, true>(rb, win, src, g.stream, g.wa, g.g_lo, g.g_hi)
                                                : convert_window<SRCW, WF, WS, false, true, true>(rb, win, src, g.stream, g.wa, g.g_lo, g.g_hi);
            else big = src.big_endian ? convert_window<SRCW, WF, WS, true, true, false>(rb, win, src, g.stream, g.wa, g.g_lo, g.g_hi)
                                      : convert_window<SRCW, WF, WS, false, true, false>(rb, win, src, g.stream, g.wa, g.g_lo, g.g_hi);
        } else {
            big = src.big_endian ? convert_window<SRCW, WF, WS, true, false, false>(rb, win, src, g.stream, g.wa, g.g_lo, g.g_hi)
                                 : convert_window<SRCW, WF, WS, false, false, false>(rb, win, src, g.stream, g.wa, g.g_lo, g.g_hi);
        }
        const int any_big = __syncthreads_or((int)big);  // windows ready, raw[buf] free
        if (threadIdx.x == 0) issue(k + NB);

        // ---- FIR
        const float* w = win + lane * WS + STEP * R * b;
        float acc[R][8];
        fir_core2<STEP, R, LEAD>(w, taps, acc);
        float y[R];
#pragma unroll
        for (int r = 0; r < R; r++) y[r] = reduce8(acc[r]);
        if (any_big) {  // exact slow path, see resample_int_kernel
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float p0 = dot256(w + LEAD + STEP * r, row1 - SINC_LEN), p1 = dot256(w + LEAD + STEP * r, row1);
                y[r] = __fadd_rn(p0, __fmul_rn(0.0f, __fsub_rn(p1, p0)));
            }
        }
        int_epilogue<R, PCM>(y, dst, g.stream, g.n0, b, lane, 0, true);
        __syncthreads();  // windows free for the next tile
    }
#ifdef SK_TRACE_CTAS
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sk_cta_trace[3 * blockIdx.x + 2] = t;
    }
#endif
}

// ================================================================================================
// integer-step kernel v3: persistent CTAs, PLANAR f32 rows fetched by the TMA straight into double-buffered windows
// ================================================================================================
// The stateful resamplers (sk_resampler, sk_resampler_batch) keep every (stream, channel) lane as a device row of f32
// [history | pending input].  No conversion pass is needed for such a source, so the TMA writes the two lanes' windows of tile
// k + 1 directly into the second window buffer while tile k is convolved; tiles (lane pair x 128 R outputs) are drawn by ticket
// like the interleaved kernel's.  Requires 16-byte aligned rows and (origin + i0_first - LEAD) % 4 == 0, which the batched
// handle arranges by choosing where in the row it keeps its history; anything else takes resample_int_kernel.
template <int STEP> __host__ __device__ constexpr int planar_smem_bytes() { return 4 * tma_lane_stride<STEP>() * (int)sizeof(float) + 64; }
struct PlanarGeom { int stream; int lane0; int n0; int wa; int c_lo; int c_hi; int lane1; int pad; };  // [c_lo, c_hi): window frames the TMA copies

template <int STEP>
__global__ void __launch_bounds__(kIntThreads, IntCfg<STEP>::CTAS)
resample_int_planar_kernel(const __grid_constant__ Taps taps, const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, long long i0_first,
                           int lanes_per_stream, unsigned tiles_per_lane, unsigned total_tiles, const float* __restrict__ row1, unsigned* __restrict__ ticket) {
    constexpr int R = IntCfg<STEP>::R;
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int LEAD = tma_lead<STEP>();
    constexpr int WF = tma_frames<STEP>();
    constexpr int WS = tma_lane_stride<STEP>();
    static_assert(((LEAD + STEP * (R - 1)) & 1) == 0 || WF > LEAD + int_window<STEP>(), "stage one more frame for this STEP / R");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* wins = reinterpret_cast<float*>(smem_raw);  // [2 buffers][2 lanes][WS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 4 * WS * sizeof(float));
    __shared__ PlanarGeom geoms[4];
    const unsigned pps = (unsigned)(lanes_per_stream + 1) >> 1;  // lane pairs per stream

    // one elected thread: ticket and geometry of the CTA's k-th tile -> geoms[k & 3] (stream = -1: no tile left), then the two
    // bulk copies of the part of the window that lies inside the row's valid frames, rounded inward to whole 16-byte units
    auto issue = [&](unsigned k) {
        const unsigned t = atomicAdd(ticket, 1u);
        if (t >= total_tiles) { geoms[k & 3].stream = -1; return; }
        const unsigned pair = t / tiles_per_lane, stream = pair / pps;
        PlanarGeom g;
        g.stream = (int)stream; g.lane0 = 2 * (int)(pair - stream * pps); g.lane1 = g.lane0 + 1 < lanes_per_stream;
        g.n0 = (int)((t - pair * tiles_per_lane) * (unsigned)T);
        const long long wa = i0_first + (long long)STEP * g.n0 - LEAD;  // first frame of the shared window
        g.wa = (int)wa;
        const long long lo = wa > src.valid_lo ? wa : src.valid_lo, hi = wa + WF < src.valid_hi ? wa + WF : src.valid_hi;
        g.c_lo = (int)((lo - wa + 3) & ~3ll);
        g.c_hi = hi > wa ? (int)((hi - wa) & ~3ll) : 0;
        if (g.c_hi < g.c_lo) g.c_hi = g.c_lo;
        geoms[k & 3] = g;
        const int buf = (int)(k & 1);
        const uint32_t bar = smem_u32(&bars[buf]);
        const uint32_t bytes = (uint32_t)(g.c_hi - g.c_lo) * 4u;
        // the buffer was read by the FIR and patched at its edges through the generic proxy: order that before the async writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, g.lane1 ? 2 * bytes : bytes);
        if (bytes) {
            const uint8_t* p0 = src.base + (long long)stream * src.stream_stride + (long long)g.lane0 * src.lane_stride + (wa + g.c_lo + src.origin) * 4;
            float* w = wins + buf * 2 * WS + g.c_lo;
            tma_load_1d(smem_u32(w), p0, bytes, bar);
            if (g.lane1) tma_load_1d(smem_u32(w + WS), p0 + src.lane_stride, bytes, bar);
        }
    };

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_fence_init();
        issue(0);
        issue(1);
    }
    __syncthreads();

    const int lane = threadIdx.x & 1, b = threadIdx.x >> 1;
    for (unsigned k = 0;; k++) {
        const PlanarGeom g = geoms[k & 3];
        if (g.stream < 0) break;  // tickets are monotonic: nothing was requested for this or any later k
        const int buf = (int)(k & 1);
        float* win = wins + buf * 2 * WS;
        // ---- the frames the bulk copies do not cover (up to 3 at either end of the valid range, everything outside it, all of a
        // missing second lane): frame by frame, zero outside the row's valid frames
        {
            const float* p0 = reinterpret_cast<const float*>(src.base + (long long)g.stream * src.stream_stride + (long long)g.lane0 * src.lane_stride);
            const int edge = g.c_lo + (WF - g.c_hi), edge1 = g.lane1 ? edge : WF;
            for (int i = threadIdx.x; i < edge + edge1; i += kIntThreads) {
                const int l = i >= edge, j = l ? i - edge : i;
                const int f = (l && !g.lane1) ? j : (j < g.c_lo ? j : g.c_hi + (j - g.c_lo));  // window frame
                const long long ff = (long long)g.wa + f;
                float v = 0.0f;
                if ((!l || g.lane1) && ff >= src.valid_lo && ff < src.valid_hi) v = __ldg(p0 + (l ? src.lane_stride / 4 : 0) + ff + src.origin);
                win[l * WS + f] = v;
            }
        }
        mbar_wait(smem_u32(&bars[buf]), (k >> 1) & 1);
        __syncthreads();
        // ---- huge / non-finite inputs take the exact path
        bool big = false;
        for (int i = threadIdx.x; i < 2 * (WF / 4); i += kIntThreads) {
            const int l = i >= WF / 4, j = l ? i - WF / 4 : i;
            const uint4 u = *reinterpret_cast<const uint4*>(win + l * WS + 4 * j);
            big |= ((u.x & 0x7f800000u) >= 0x78000000u) | ((u.y & 0x7f800000u) >= 0x78000000u) | ((u.z & 0x7f800000u) >= 0x78000000u) |
                   ((u.w & 0x7f800000u) >= 0x78000000u);
        }
        const int any_big = __syncthreads_or((int)big);

        // ---- FIR
        const float* w = win + lane * WS + STEP * R * b;
        float acc[R][8];
        fir_core2<STEP, R, LEAD>(w, taps, acc);
        float y[R];
#pragma unroll
        for (int r = 0; r < R; r++) y[r] = reduce8(acc[r]);
        if (any_big) {  // exact slow path, see resample_int_kernel
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float p0 = dot256(w + LEAD + STEP * r, row1 - SINC_LEN), p1 = dot256(w + LEAD + STEP * r, row1);
                y[r] = __fadd_rn(p0, __fmul_rn(0.0f, __fsub_rn(p1, p0)));
            }
        }
        int_epilogue<R, false>(y, dst, g.stream, g.n0, b, lane, g.lane0, g.lane1 != 0);
        __syncthreads();  // this window buffer is free: request the tile after next into it
        if (threadIdx.x == 0) issue(k + 2);
    }
}

// ================================================================================================
// host: launchers of the integer-step kernels
// ================================================================================================
template <int STEP>
static int launch_int_planar(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams,
                             int lanes_per_stream) {
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int smem = planar_smem_bytes<STEP>();
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_planar_kernel<STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long tiles_per_lane = (long long)((plan->n_out + T - 1) / T);
    const long long total = tiles_per_lane * (long long)((lanes_per_stream + 1) / 2) * (long long)n_streams;
    const int per_sm = std::max(1, std::min(IntCfg<STEP>::CTAS, (220 * 1024) / smem));
    const long long grid = std::min<long long>(total, (long long)ctx->num_sms * per_sm);
    constexpr unsigned kTicketRing = 256;
    if (!ctx->d_tickets) SK_CUDA(ctx, cudaMalloc((void**)&ctx->d_tickets, kTicketRing * sizeof(unsigned)));
    unsigned* ticket = ctx->d_tickets + (ctx->ticket_next++ % kTicketRing);
    SK_CUDA(ctx, cudaMemsetAsync(ticket, 0, sizeof(unsigned), ctx->stream));
    resample_int_planar_kernel<STEP><<<(unsigned)grid, kIntThreads, smem, ctx->stream>>>(taps, src, dst, plan->i0_first, lanes_per_stream, (unsigned)tiles_per_lane,
                                                                                       (unsigned)total, d_table + SINC_LEN, ticket);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}
template <int STEP, int SRCW>
static int launch_int_tma(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams) {
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int smem = tma_smem_bytes<STEP, SRCW>();
    constexpr bool kPcm = int_step_has_pcm<STEP>();  // byte epilogue instantiated for the common steps only (compile time)
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_tma_kernel<STEP, SRCW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (kPcm) SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_tma_kernel<STEP, SRCW, kPcm>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long tiles_per_stream = (long long)((plan->n_out + T - 1) / T);
    const long long total = tiles_per_stream * (long long)n_streams;
    const int per_sm = std::max(1, std::min(IntCfg<STEP>::CTAS, (220 * 1024) / smem));
    const long long grid = std::min<long long>(total, (long long)ctx->num_sms * per_sm);
    // integer sources are staged unscaled (see convert4): fold the exact power-of-two scale into the taps and let the
    // per-frame edge path (decode_sample_bits) stage the same unscaled values
    SrcDesc s2 = src;
    Taps t2 = taps;
    if (!src.is_float) {
        for (int i = 0; i < SINC_LEN; i++) t2.he[i] *= src.scale;
        for (int i = 0; i < SINC_LEN + 8; i++) t2.ho[i] *= src.scale;
        s2.scale = 1.0f;
    }
    if (dst.pcm_bits && !kPcm) SK_FAIL(ctx, SK_E_STATE, "internal: no byte epilogue for integer step %d", STEP);
    constexpr unsigned kTicketRing = 256;
    if (!ctx->d_tickets) SK_CUDA(ctx, cudaMalloc((void**)&ctx->d_tickets, kTicketRing * sizeof(unsigned)));
    unsigned* ticket = ctx->d_tickets + (ctx->ticket_next++ % kTicketRing);
    SK_CUDA(ctx, cudaMemsetAsync(ticket, 0, sizeof(unsigned), ctx->stream));
    if (dst.pcm_bits)
        resample_int_tma_kernel<STEP, SRCW, kPcm><<<(unsigned)grid, kIntThreads, smem, ctx->stream>>>(t2, s2, dst, plan->i0_first, (unsigned)tiles_per_stream,
                                                                                                    (unsigned)total, d_table + SINC_LEN, ticket);
    else
        resample_int_tma_kernel<STEP, SRCW, false><<<(unsigned)grid, kIntThreads, smem, ctx->stream>>>(t2, s2, dst, plan->i0_first, (unsigned)tiles_per_stream,
                                                                                                     (unsigned)total, d_table + SINC_LEN, ticket);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

int sk_launch_int3_tma_w2(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams);
int sk_launch_int3_tma_w3(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams);
int sk_launch_int3_tma_w4(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams);

template <int STEP>
static int launch_int(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* h_table, const float* d_table,
                      size_t n_streams, int lanes_per_stream) {
    Taps taps;
    memset(&taps, 0, sizeof(taps));
    memcpy(taps.he, h_table, SINC_LEN * sizeof(float));      // subindex 0
    memcpy(taps.ho + 1, h_table, SINC_LEN * sizeof(float));  // ho[i] = h[i-1]
    // fast path: interleaved stereo, 16-byte aligned streams, canonical first index (batch plans)
    if (src.width != 0 && src.channels == 2 && lanes_per_stream == 2 && src.origin == 0 && src.valid_lo == 0 &&
        plan->i0_first == -(long long)(SINC_LEN / 2) + STEP && ((uintptr_t)src.base % 16) == 0 && (src.stream_stride % 16) == 0 &&
        (long long)n_streams * (long long)((plan->n_out + int_tile_outputs<STEP>() - 1) / int_tile_outputs<STEP>()) < (1ll << 31)) {
        if constexpr (STEP == 3) {  // its six TMA kernels are the longest compiles of the build: one translation unit per sample width
            if (src.width == 2) return sk_launch_int3_tma_w2(ctx, taps, src, dst, plan, d_table, n_streams);
            if (src.width == 3) return sk_launch_int3_tma_w3(ctx, taps, src, dst, plan, d_table, n_streams);
            return sk_launch_int3_tma_w4(ctx, taps, src, dst, plan, d_table, n_streams);
        } else {
            if (src.width == 2) return launch_int_tma<STEP, 2>(ctx, taps, src, dst, plan, d_table, n_streams);
            if (src.width == 3) return launch_int_tma<STEP, 3>(ctx, taps, src, dst, plan, d_table, n_streams);
            return launch_int_tma<STEP, 4>(ctx, taps, src, dst, plan, d_table, n_streams);
        }
    }
    // planar f32 rows (the stateful resamplers), aligned so that every tile window starts on a 16-byte boundary of its row
    if (src.width == 0 && !dst.pcm_bits && ((uintptr_t)src.base % 16) == 0 && (src.stream_stride % 16) == 0 && (src.lane_stride % 16) == 0 &&
        ((src.origin + plan->i0_first - tma_lead<STEP>()) & 3) == 0 && plan->i0_first > -(1ll << 30) && plan->i0_first < (1ll << 30) && plan->n_out < (1u << 30) &&
        (long long)n_streams * ((lanes_per_stream + 1) / 2) * (long long)((plan->n_out + int_tile_outputs<STEP>() - 1) / int_tile_outputs<STEP>()) < (1ll << 31))
        return launch_int_planar<STEP>(ctx, taps, src, dst, plan, d_table, n_streams, lanes_per_stream);
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int smem = int_smem_bytes<STEP>();
    constexpr bool kPcm = int_step_has_pcm<STEP>();
    if (dst.pcm_bits && !kPcm) SK_FAIL(ctx, SK_E_STATE, "internal: no byte epilogue for integer step %d", STEP);
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_kernel<STEP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (kPcm) SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_kernel<STEP, kPcm>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const size_t tiles = (plan->n_out + T - 1) / T;
    const size_t pairs = (size_t)((lanes_per_stream + 1) / 2) * n_streams;
    const size_t pps = (size_t)((lanes_per_stream + 1) / 2);
    const size_t slab = (65535 / pps) * pps;  // gridDim.y is limited to 65535: whole streams per slab
    for (size_t p0 = 0; p0 < pairs; p0 += slab) {
        const size_t np = std::min<size_t>(slab, pairs - p0);
        SrcDesc s2 = src; OutDesc d2 = dst;
        const size_t stream0 = p0 / pps;
        s2.base += (long long)stream0 * src.stream_stride;
        d2.base += (long long)stream0 * dst.stream_stride;
        dim3 grid((unsigned)tiles, (unsigned)np);
        d2.pcm_base += (long long)stream0 * dst.pcm_stream_stride;
        if (dst.pcm_bits) resample_int_kernel<STEP, kPcm><<<grid, kIntThreads, smem, ctx->stream>>>(taps, s2, d2, plan->i0_first, lanes_per_stream, d_table + SINC_LEN);
        else resample_int_kernel<STEP, false><<<grid, kIntThreads, smem, ctx->stream>>>(taps, s2, d2, plan->i0_first, lanes_per_stream, d_table + SINC_LEN);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}


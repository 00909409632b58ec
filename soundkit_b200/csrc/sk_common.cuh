// sk_common.cuh — shared host/device helpers of libsoundkit_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/soundkit_cuda.h"

#define SK_NUM_SMS_FALLBACK 148

struct sk_plan;  // resampler index plan (sk_resample.cu)

struct sk_ctx {
    int device = 0;
    int num_sms = SK_NUM_SMS_FALLBACK;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream kernels are launched on (own or caller's)
    uint64_t launches = 0;
    std::string err;
    // pinned host staging + device arena for the host-pointer entry points (grow-only)
    void* h_stage = nullptr; size_t h_cap = 0;
    void* d_in = nullptr;    size_t d_in_cap = 0;
    void* d_out = nullptr;   size_t d_out_cap = 0;
    void* d_tmp = nullptr;   size_t d_tmp_cap = 0;     // scratch of sk_normalize_batch_pcm_dev (mono f32 before quantisation)
    // tile tickets of the persistent integer-step kernel: one counter per launch, taken round-robin from a ring so that
    // launches in flight on different streams of one context never share a counter (sk_resample_dev.cuh)
    unsigned int* d_tickets = nullptr; unsigned int ticket_next = 0;
    void* batch_pipe = nullptr;                       // sk_resample.cu: slab pipeline of sk_normalize_batch (streams, events, device slabs)
    bool g726_tables_ready = false;                   // sk_g726.cu: __constant__ tables uploaded on this device
    // sinc tables per ratio, built lazily
    std::map<uint64_t, float*> sinc_tables;          // key = (in_rate << 32) | out_rate -> device [256][256]
    std::map<uint64_t, std::vector<float>> sinc_host; // same key -> host copy
    std::map<uint64_t, float*> sinc_tables4;         // same key -> four word-shifted zero-padded copies (sk_resample_gen3.cu)
    std::map<std::string, sk_plan*> plans;            // batch plans keyed by rate/frames/semantics (bounded LRU, sk_resample.cu)
    uint64_t plan_clock = 0;
    void* d_res = nullptr;   size_t d_res_cap = 0;     // result planes of sk_normalize_stream (grow-only; no per-call cudaMalloc)
    void* d_rows = nullptr;  size_t d_rows_cap = 0;    // planar f32 rows of the two-pass chain for mono / multi-channel sources (sk_resample.cu)
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};     // pinned-stage slot events of sk_h2d / sk_d2h (created once)
    size_t batch_slab_bytes = (size_t)256 << 20;      // input bytes per slab of the sk_normalize_batch pipeline (sk_ctx_set_batch_slab_bytes)
    bool resample_fast = false;                       // sk_ctx_set_resample_mode: opt-in tolerance mode of the general-ratio resampler
    bool disable_gen3 = false;                        // diagnostic (SK_DISABLE_GEN3=1): general ratios use the older 2 x 2 tile kernel
    unsigned long long* d_stats = nullptr;            // [0] outputs of the general-ratio resampler that took the per-output exact path
};

#define SK_FAIL(ctx, code, ...)                                 \
    do {                                                        \
        char _b[512];                                           \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                  \
        (ctx)->err = _b;                                        \
        return (code);                                          \
    } while (0)

#define SK_CUDA(ctx, expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) SK_FAIL(ctx, SK_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define SK_LAUNCH_CHECK(ctx)                                                                         \
    do {                                                                                             \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess) SK_FAIL(ctx, SK_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(_e)); \
        (ctx)->launches++;                                                                           \
    } while (0)

// grow-only staging buffers ----------------------------------------------------------------
int sk_ensure_host(sk_ctx* ctx, size_t bytes);
int sk_ensure_dev_in(sk_ctx* ctx, size_t bytes);
int sk_ensure_dev_out(sk_ctx* ctx, size_t bytes);
int sk_ensure_stats(sk_ctx* ctx);  // device counters (zeroed on first use)
// host -> device through the pinned stage (chunked), device -> host likewise; both synchronous wrt host
int sk_h2d(sk_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
// Small host-pointer calls: the kernel reads and writes the context's pinned staging buffer directly over PCIe (pinned
// allocations are device-addressable at the same address), so a call is memcpy + one launch + one synchronize + memcpy
// instead of two DMA copies, one launch and two synchronizes.  sk_small_begin returns false when the call is too large for
// this path (the caller then stages through device memory).
constexpr size_t kSkSmallCallBytes = 64u << 10;
bool sk_small_begin(sk_ctx* ctx, const void* in_host, size_t in_bytes, size_t out_bytes, void** in_dev, void** out_dev);
// the two halves of the pinned staging buffer (kSkStageHalf bytes each: inputs / outputs of a zero-copy call); null on allocation failure
constexpr size_t kSkStageHalf = 8u << 20;
char* sk_stage_half(sk_ctx* ctx, int half);
int sk_small_end(sk_ctx* ctx, void* out_host, const void* out_dev, size_t out_bytes);
int sk_d2h(sk_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

static inline size_t sk_round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __CUDACC__
// 128-bit streaming loads/stores: data touched once, keep it out of L1
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// non-volatile forms (read-only data): the compiler may hoist several of them ahead of their uses
__device__ __forceinline__ uint4 ld_stream_v4_nv(const void* p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_v2(const void* p) {
    uint2 r;
    asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
    uint32_t r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_v2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// ---- float -> integer sample quantisers shared by sk_pcm.cu (f32_channels_to_bytes) and the fused chain's byte epilogue
// Rust f32::clamp(-1.0, 1.0): NaN stays NaN
__device__ __forceinline__ float clamp1(float x) { return x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); }
__device__ __forceinline__ bool finite_f(float x) { return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u; }

// float_sample_to_i16, soundkit-decoder/src/lib.rs:1815-1827: non-finite -> 0, clamp, x < 0 ? x * 32768 : x * 32767 IN F64,
// f64::round (half away from zero), clamp to i16.  Evaluated with ONE f32 operation that rounds once: round-half-away of
// |v| is floor(|v| + 0.5), and fma.rm(|x|, c, 2^22 + 0.5) forms |x|*c + 0.5 exactly before its single rounding toward -inf
// to a multiple of 0.5 (every float in [2^22, 2^23) is one) — flooring a value that was floored to halves is flooring the
// value itself, so the mantissa >> 1 IS the rounded magnitude (its low 16 bits; the exponent sits above bit 22).  |x| <= 1
// keeps the magnitude <= 32767 (>= 0) / 32768 (< 0): the reference's final clamp never acts.  No f64, no conversion
// instruction (history: the FP64 form made the s16 byte writer compute bound; a Fast2Sum f32 form with FRND + F2I still
// kept the ALU pipe 80 % busy at 0.88 of the HBM rate).  Callers keep the low 16 bits.
__device__ __forceinline__ int32_t float_sample_to_i16_dev(float s) {
    const bool fin = fabsf(s) <= 3.402823466e+38f;       // false for NaN and +-Inf
    const float m = fin ? fminf(fabsf(s), 1.0f) : 0.0f;
    const bool neg = s < 0.0f;
    const float q = __fmaf_rd(m, neg ? 32768.0f : 32767.0f, 4194304.5f);
    const int r = (int)((__float_as_uint(q) >> 1) & 0xffffu);
    return neg ? -r : r;
}
// audio_bytes.rs:205-220 / decoder lib.rs:3653-3659
__device__ __forceinline__ int32_t f32_to_s24_dev(float x) {
    float c = clamp1(x);
    return c >= 0.0f ? __float2int_rz(__fmul_rn(c, 8388607.0f)) : __float2int_rz(__fmul_rn(c, 8388608.0f));
}
// audio_bytes.rs:189-203 / decoder lib.rs:3667-3673: both multipliers are 2^31 in f32, `as` saturates
__device__ __forceinline__ int32_t f32_to_i32_dev(float x) { return __float2int_rz(__fmul_rn(clamp1(x), 2147483648.0f)); }
__device__ __forceinline__ int32_t f32_to_i16_trunc_dev(float x) { return __float2int_rz(__fmul_rn(clamp1(x), 32767.0f)); }
#endif

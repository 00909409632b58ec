// sk_pcm.cu — kernel 1: sample-width / endian conversion and (de)interleave.
//
// Replaces the scalar loops of soundkit/src/audio_bytes.rs:3-373, soundkit/src/audio_pipeline.rs:17-131
// and soundkit-decoder/src/lib.rs:1815-1827, 3458-3701.  All of it is HBM-bound byte/integer work
// (6-8 B per sample), so the design is about coalescing, not arithmetic:
//   * every thread owns a group of G frames chosen so that its input span (G*CH*B bytes) and each
//     of its per-channel output spans are whole 16-byte vectors -> only 128-bit loads and stores,
//     also for packed 24-bit (48-byte groups);
//   * samples are cut out of the register-resident words with PRMT at compile-time offsets;
//   * integer->float scaling is by exact powers of two, float->int uses saturating round-toward-
//     zero converts (Rust `as`), and nothing is ever fused (plain Rust arithmetic has no FMA):
//     the file is built with -fmad=false and uses explicit _rn intrinsics where order matters.
// A scalar kernel (one thread per sample, byte loads) covers channel counts > 2, unaligned caller
// pointers and the < G-frame tail; it is correctness-only.
#include <algorithm>

#include "sk_common.cuh"

// ------------------------------------------------------------------------------------------------
// sample extraction from a register array holding the thread's bytes in memory order
// ------------------------------------------------------------------------------------------------
template <int B, bool BE>
__device__ __forceinline__ int32_t get_sample(const uint32_t* w, int k) {  // k is compile-time after unrolling
    if (B == 1) {
        return (int32_t)((w[k >> 2] >> (8 * (k & 3))) & 0xffu);  // unsigned byte
    } else if (B == 2) {
        uint32_t word = w[k >> 1];
        uint32_t h = (k & 1) ? (word >> 16) : (word & 0xffffu);
        if (BE) h = ((h & 0xff) << 8) | (h >> 8);
        return (int32_t)(int16_t)h;
    } else if (B == 3) {
        const int o = 3 * k, wi = o >> 2, sh = o & 3;
        const uint32_t lo = w[wi];
        const uint32_t hi = (sh >= 2) ? w[wi + 1] : 0u;  // bytes sh..sh+2 spill into the next word when sh >= 2
        // place the 3 bytes in the upper 24 bits (MSB of the sample in byte 3), then arithmetic >> 8
        uint32_t sel;
        if (!BE) sel = ((uint32_t)(sh + 2) << 12) | ((uint32_t)(sh + 1) << 8) | ((uint32_t)sh << 4);
        else     sel = ((uint32_t)sh << 12) | ((uint32_t)(sh + 1) << 8) | ((uint32_t)(sh + 2) << 4);
        uint32_t v = __byte_perm(lo, hi, sel);
        return ((int32_t)v) >> 8;
    } else {
        uint32_t word = w[k];
        if (BE) word = bswap32(word);
        return (int32_t)word;
    }
}

// ------------------------------------------------------------------------------------------------
// flat ops
// ------------------------------------------------------------------------------------------------
struct FlatOpInfo { int in_b; bool be; int out_b; };
__host__ __device__ constexpr FlatOpInfo flat_info(int op) {
    switch (op) {
        case SK_OP_I16LE_TO_F32: return {2, false, 4};
        case SK_OP_S16LE_TO_I16: return {2, false, 2};
        case SK_OP_S16BE_TO_I16: return {2, true, 2};
        case SK_OP_S16LE_TO_I32: return {2, false, 4};
        case SK_OP_S24LE_TO_I32: return {3, false, 4};
        case SK_OP_S24LE_TO_I16: return {3, false, 2};
        case SK_OP_S24BE_TO_I16: return {3, true, 2};
        case SK_OP_S32LE_TO_I32: return {4, false, 4};
        case SK_OP_S32BE_TO_I32: return {4, true, 4};
        case SK_OP_S32LE_TO_S24: return {4, false, 4};
        case SK_OP_S32BE_TO_S24: return {4, true, 4};
        case SK_OP_S32LE_TO_F32: return {4, false, 4};
        case SK_OP_S32BE_TO_F32: return {4, true, 4};
        case SK_OP_S32LE_TO_I16: return {4, false, 2};
        case SK_OP_S32BE_TO_I16: return {4, true, 2};
        case SK_OP_F32LE_TO_I16: return {4, false, 2};
        case SK_OP_F32BE_TO_I16: return {4, true, 2};
        case SK_OP_F32LE_TO_I32: return {4, false, 4};
        case SK_OP_F32LE_TO_S24: return {4, false, 4};
        case SK_OP_PKT_S16LE_TO_F32: return {2, false, 4};
        case SK_OP_PKT_S24LE_TO_F32: return {3, false, 4};
        case SK_OP_PKT_S32LE_TO_I16: return {4, false, 2};
        case SK_OP_PKT_F32LE_TO_I16: return {4, false, 2};
        case SK_OP_U8_TO_S16: return {1, false, 2};
        default: return {0, false, 0};
    }
}

// raw = sign-extended little-endian-corrected sample bits; returns the output element's bits
template <int OP>
__device__ __forceinline__ uint32_t flat_apply(int32_t raw) {
    switch (OP) {
        case SK_OP_I16LE_TO_F32: return __float_as_uint(__fmul_rn((float)raw, 1.0f / 32768.0f));
        case SK_OP_S16LE_TO_I16: case SK_OP_S16BE_TO_I16: return (uint32_t)raw & 0xffffu;
        case SK_OP_S16LE_TO_I32: case SK_OP_S24LE_TO_I32: case SK_OP_S32LE_TO_I32: case SK_OP_S32BE_TO_I32: return (uint32_t)raw;
        case SK_OP_S24LE_TO_I16: case SK_OP_S24BE_TO_I16: return (uint32_t)(raw >> 8) & 0xffffu;
        case SK_OP_S32LE_TO_S24: case SK_OP_S32BE_TO_S24: return (uint32_t)raw & 0x00ffffffu;
        case SK_OP_S32LE_TO_F32: case SK_OP_S32BE_TO_F32: return __float_as_uint(__fmul_rn(__int2float_rn(raw), 4.656612873077393e-10f));
        case SK_OP_S32LE_TO_I16: case SK_OP_S32BE_TO_I16: return (uint32_t)(raw >> 16) & 0xffffu;
        case SK_OP_F32LE_TO_I16: case SK_OP_F32BE_TO_I16: return (uint32_t)f32_to_i16_trunc_dev(__int_as_float(raw)) & 0xffffu;
        case SK_OP_F32LE_TO_I32: return (uint32_t)f32_to_i32_dev(__int_as_float(raw));
        case SK_OP_F32LE_TO_S24: return (uint32_t)f32_to_s24_dev(__int_as_float(raw));
        case SK_OP_PKT_S16LE_TO_F32: return __float_as_uint(__fdiv_rn((float)raw, 32767.0f));
        case SK_OP_PKT_S24LE_TO_F32: return __float_as_uint(__fmul_rn((float)raw, 1.0f / 8388608.0f));
        case SK_OP_PKT_S32LE_TO_I16: {  // i64 product, truncating division, clamp (the clamp never bites: |q| <= 32767)
            long long q = ((long long)raw * 32767ll) / 2147483647ll;
            return (uint32_t)max(-32768ll, min(32767ll, q)) & 0xffffu;
        }
        case SK_OP_PKT_F32LE_TO_I16: return (uint32_t)max(-32768, min(32767, __float2int_rz(__fmul_rn(__int_as_float(raw), 32767.0f)))) & 0xffffu;
        case SK_OP_U8_TO_S16: return (uint32_t)((raw - 128) << 8) & 0xffffu;
        default: return 0;
    }
}

constexpr int kPcmThreads = 256;
constexpr int kPcmWarps = kPcmThreads / 32;

// Warp-tiled global I/O.  A lane's natural unit is a contiguous span of NV 16-byte vectors (its group of samples); issued
// directly, each warp-wide access would touch every NV-th vector of a 512*NV-byte run (partially used sectors, 0.67-0.86
// of HBM).  Instead the warp moves the run with NV fully contiguous 512-byte accesses and the vectors change lanes through
// NV*512 bytes of warp-private shared memory.  NV == 1 needs no exchange.
template <int NV>
__device__ __forceinline__ void warp_load_spans(uint4* xch, const uint4* __restrict__ gsrc, uint32_t* w, int lane) {
    if constexpr (NV <= 2) {  // measured: up to two vectors per lane, direct L1-allocating loads of the lane's own span are as fast or faster
#pragma unroll
        for (int v = 0; v < NV; v++) { const uint4 q = __ldg(gsrc + lane * NV + v); w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w; }
    } else {
        uint4 r[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) r[v] = ld_stream_v4_nv(gsrc + v * 32 + lane);
        __syncwarp();
#pragma unroll
        for (int v = 0; v < NV; v++) xch[v * 32 + lane] = r[v];
        __syncwarp();
#pragma unroll
        for (int v = 0; v < NV; v++) { const uint4 q = xch[lane * NV + v]; w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w; }
    }
}
template <int NV>
__device__ __forceinline__ void warp_store_spans(uint4* xch, uint4* __restrict__ gdst, const uint32_t* o, int lane) {
    if constexpr (NV == 1) {
        st_stream_v4(gdst + lane, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
        __syncwarp();
#pragma unroll
        for (int v = 0; v < NV; v++) xch[lane * NV + v] = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        __syncwarp();
#pragma unroll
        for (int v = 0; v < NV; v++) st_stream_v4(gdst + v * 32 + lane, xch[v * 32 + lane]);
    }
}

// G samples per lane: input G*B bytes = NV_IN 16-byte vectors, output G*OB bytes = NV_OUT vectors; a warp owns 32 groups
template <int OP>
__global__ void __launch_bounds__(kPcmThreads) flat_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t ntiles) {
    constexpr FlatOpInfo I = flat_info(OP);
    constexpr int B = I.in_b, OB = I.out_b;
    constexpr int G = B == 1 ? 16 : (B == 2 ? 8 : (B == 3 ? 16 : (OB == 2 ? 8 : 4)));
    constexpr int NV_IN = G * B / 16, NW_OUT = G * OB / 4, NV_OUT = NW_OUT / 4;
    constexpr int NVX = NV_IN > NV_OUT ? NV_IN : NV_OUT;
    __shared__ uint4 xch_all[kPcmWarps][NVX * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* xch = xch_all[warp];
    for (size_t t = (size_t)blockIdx.x * kPcmWarps + warp; t < ntiles; t += (size_t)gridDim.x * kPcmWarps) {
        uint32_t w[NV_IN * 4];
        warp_load_spans<NV_IN>(xch, reinterpret_cast<const uint4*>(in + t * (size_t)(32 * G * B)), w, lane);
        uint32_t o[NW_OUT];
#pragma unroll
        for (int k = 0; k < G; k++) {
            uint32_t r = flat_apply<OP>(I.be ? get_sample<B, true>(w, k) : get_sample<B, false>(w, k));
            if (OB == 4) o[k] = r;
            else if (k & 1) o[k >> 1] |= r << 16;
            else o[k >> 1] = r;
        }
        warp_store_spans<NV_OUT>(xch, reinterpret_cast<uint4*>(out + t * (size_t)(32 * G * OB)), o, lane);
    }
}

template <int OP>
__global__ void flat_scalar_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t first, size_t n) {
    constexpr FlatOpInfo I = flat_info(OP);
    for (size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t* p = in + i * I.in_b;
        uint32_t u = 0;
        if (!I.be) { for (int b = 0; b < I.in_b; b++) u |= (uint32_t)p[b] << (8 * b); }
        else       { for (int b = 0; b < I.in_b; b++) u = (u << 8) | p[b]; }
        int32_t raw = I.in_b == 1 ? (int32_t)u : (I.in_b == 2 ? (int32_t)(int16_t)u : (I.in_b == 3 ? ((int32_t)(u << 8)) >> 8 : (int32_t)u));
        uint32_t r = flat_apply<OP>(raw);
        if (I.out_b == 2) reinterpret_cast<uint16_t*>(out)[i] = (uint16_t)r;
        else reinterpret_cast<uint32_t*>(out)[i] = r;
    }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
// One work item per thread, as many CTAs as that takes (the kernels keep a grid-stride loop for > 2^22 CTAs).  Measured on
// B200: a small persistent grid whose CTAs stride through memory in lockstep reaches 0.85-0.89 of the HBM copy rate on these
// streaming kernels, the same kernel launched as a large grid of short-lived CTAs 1.0-1.08 (G.711 encode: 5.8 -> 7.0 TB/s).
static inline int grid_for(sk_ctx* ctx, size_t work_items, int threads) {
    (void)ctx;
    size_t need = (work_items + threads - 1) / threads;
    const size_t cap = (size_t)1 << 22;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

template <int OP>
static int launch_flat(sk_ctx* ctx, const uint8_t* in, size_t n, uint8_t* out) {
    constexpr FlatOpInfo I = flat_info(OP);
    constexpr int G = I.in_b == 1 ? 16 : (I.in_b == 2 ? 8 : (I.in_b == 3 ? 16 : (I.out_b == 2 ? 8 : 4)));
    const size_t ntiles = (aligned16(in) && aligned16(out)) ? n / (32 * (size_t)G) : 0;  // warp tiles of 32 groups
    if (ntiles) {
        flat_kernel<OP><<<grid_for(ctx, ntiles * 32, kPcmThreads), kPcmThreads, 0, ctx->stream>>>(in, out, ntiles);
        SK_LAUNCH_CHECK(ctx);
    }
    if (ntiles * 32 * G < n) {
        flat_scalar_kernel<OP><<<grid_for(ctx, n - ntiles * 32 * G, 256), 256, 0, ctx->stream>>>(in, out, ntiles * 32 * G, n);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

extern "C" int sk_flat_op_widths(int op, int* in_width, int* out_width) {
    FlatOpInfo I = flat_info(op);
    if (!I.in_b) return SK_E_ARG;
    if (in_width) *in_width = I.in_b;
    if (out_width) *out_width = I.out_b;
    return SK_OK;
}

extern "C" int sk_pcm_flat_convert_dev(sk_ctx* ctx, int op, const uint8_t* in, size_t bytes, void* outv) {
    if (!ctx) return SK_E_ARG;
    FlatOpInfo I = flat_info(op);
    if (!I.in_b) SK_FAIL(ctx, SK_E_ARG, "unknown flat op %d", op);
    size_t n = bytes / I.in_b;
    if (!n) return SK_OK;
    if (!in || !outv) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t* out = (uint8_t*)outv;
    switch (op) {
#define SK_FLAT_CASE(OP) case OP: return launch_flat<OP>(ctx, in, n, out);
        SK_FLAT_CASE(SK_OP_I16LE_TO_F32) SK_FLAT_CASE(SK_OP_S16LE_TO_I16) SK_FLAT_CASE(SK_OP_S16BE_TO_I16)
        SK_FLAT_CASE(SK_OP_S16LE_TO_I32) SK_FLAT_CASE(SK_OP_S24LE_TO_I32) SK_FLAT_CASE(SK_OP_S24LE_TO_I16)
        SK_FLAT_CASE(SK_OP_S24BE_TO_I16) SK_FLAT_CASE(SK_OP_S32LE_TO_I32) SK_FLAT_CASE(SK_OP_S32BE_TO_I32)
        SK_FLAT_CASE(SK_OP_S32LE_TO_S24) SK_FLAT_CASE(SK_OP_S32BE_TO_S24) SK_FLAT_CASE(SK_OP_S32LE_TO_F32)
        SK_FLAT_CASE(SK_OP_S32BE_TO_F32) SK_FLAT_CASE(SK_OP_S32LE_TO_I16) SK_FLAT_CASE(SK_OP_S32BE_TO_I16)
        SK_FLAT_CASE(SK_OP_F32LE_TO_I16) SK_FLAT_CASE(SK_OP_F32BE_TO_I16) SK_FLAT_CASE(SK_OP_F32LE_TO_I32)
        SK_FLAT_CASE(SK_OP_F32LE_TO_S24) SK_FLAT_CASE(SK_OP_PKT_S16LE_TO_F32) SK_FLAT_CASE(SK_OP_PKT_S24LE_TO_F32)
        SK_FLAT_CASE(SK_OP_PKT_S32LE_TO_I16) SK_FLAT_CASE(SK_OP_PKT_F32LE_TO_I16) SK_FLAT_CASE(SK_OP_U8_TO_S16)
#undef SK_FLAT_CASE
    }
    return SK_E_ARG;
}

extern "C" int sk_pcm_flat_convert(sk_ctx* ctx, int op, const uint8_t* in, size_t bytes, void* out, size_t* n_out) {
    if (!ctx) return SK_E_ARG;
    FlatOpInfo I = flat_info(op);
    if (!I.in_b) SK_FAIL(ctx, SK_E_ARG, "unknown flat op %d", op);
    size_t n = bytes / I.in_b;
    if (n_out) *n_out = n;
    if (!n) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    void *s_in, *s_out;
    if (sk_small_begin(ctx, in, n * I.in_b, n * I.out_b, &s_in, &s_out)) {
        if ((rc = sk_pcm_flat_convert_dev(ctx, op, (const uint8_t*)s_in, n * I.in_b, s_out))) return rc;
        return sk_small_end(ctx, out, s_out, n * I.out_b);
    }
    if ((rc = sk_ensure_dev_in(ctx, n * I.in_b))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, n * I.out_b))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, n * I.in_b))) return rc;
    if ((rc = sk_pcm_flat_convert_dev(ctx, op, (const uint8_t*)ctx->d_in, n * I.in_b, ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, n * I.out_b);
}

// ------------------------------------------------------------------------------------------------
// unpack_dvd_lpcm, soundkit-decoder/src/lib.rs:750-804
// ------------------------------------------------------------------------------------------------
// 24-bit: a 12-byte group holds the 16 high bits of 4 samples (big-endian) followed by their 4 low bytes; the output is
// the same 4 samples as packed little-endian 24-bit: out[3s..3s+3] = {in[8+s], in[2s+1], in[2s]}.  Pure byte permutation:
// 3 words in, 4 PRMT, 3 words out; a thread owns 4 groups (48 B) so every access is a 128-bit vector.
__device__ __forceinline__ void dvd24_group(const uint32_t* w, uint32_t* o) {
    o[0] = __byte_perm(w[0], w[2], 0x5014);                                   // in8 in1 in0 in9
    o[1] = __byte_perm(__byte_perm(w[0], w[2], 0x0623), w[1], 0x5210);         // in3 in2 in10 in5
    o[2] = __byte_perm(w[1], w[2], 0x2370);                                   // in4 in11 in7 in6
}
__global__ void __launch_bounds__(kPcmThreads) dvd_lpcm24_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t nquads,
                                                                  size_t ngroups, int vec) {
    const size_t stride = (size_t)gridDim.x * kPcmThreads;
    if (vec) {
        for (size_t q = (size_t)blockIdx.x * kPcmThreads + threadIdx.x; q < nquads; q += stride) {
            uint32_t w[12], o[12];
            const uint4* p = reinterpret_cast<const uint4*>(in + q * 48);
#pragma unroll
            for (int v = 0; v < 3; v++) { uint4 t = __ldg(p + v); w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w; }
#pragma unroll
            for (int g = 0; g < 4; g++) dvd24_group(w + 3 * g, o + 3 * g);
            uint4* d = reinterpret_cast<uint4*>(out + q * 48);
#pragma unroll
            for (int v = 0; v < 3; v++) st_stream_v4(d + v, make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]));
        }
    }
    // groups not covered by whole quads (or everything, for unaligned buffers): one thread per group, byte accesses
    const size_t first = vec ? nquads * 4 : 0;
    for (size_t g = first + (size_t)blockIdx.x * kPcmThreads + threadIdx.x; g < ngroups; g += stride) {
        const uint8_t* p = in + g * 12;
        uint8_t* d = out + g * 12;
#pragma unroll
        for (int s = 0; s < 4; s++) { d[3 * s] = p[8 + s]; d[3 * s + 1] = p[2 * s + 1]; d[3 * s + 2] = p[2 * s]; }
    }
}

extern "C" int sk_dvd_lpcm_unpack_dev(sk_ctx* ctx, const uint8_t* packed, size_t bytes, int bits, int channels, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 8) SK_FAIL(ctx, SK_E_ARG, "DVD LPCM channel count %d out of range", channels);
    if (bits == 20) SK_FAIL(ctx, SK_E_ARG, "20-bit DVD LPCM decoding is not implemented");
    if (bits != 16 && bits != 24) SK_FAIL(ctx, SK_E_ARG, "unsupported DVD LPCM sample size %d", bits);
    if (bits == 16) {
        if (bytes % ((size_t)channels * 2)) SK_FAIL(ctx, SK_E_ARG, "DVD LPCM ends with a partial sample block");
        return sk_pcm_flat_convert_dev(ctx, SK_OP_S16BE_TO_I16, packed, bytes, out);  // [sample[1], sample[0]]
    }
    const size_t block = (channels == 1 || channels == 2 || channels == 4) ? 12 : (channels == 8 ? 24 : (size_t)channels * 12);
    if (bytes % block) SK_FAIL(ctx, SK_E_ARG, "DVD LPCM ends with a partial 24-bit sample block");
    if (!bytes) return SK_OK;
    if (!packed || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ngroups = bytes / 12;
    const int vec = aligned16(packed) && aligned16(out);
    const size_t nquads = vec ? ngroups / 4 : 0;
    dvd_lpcm24_kernel<<<grid_for(ctx, vec ? nquads + 1 : ngroups, kPcmThreads), kPcmThreads, 0, ctx->stream>>>(packed, out, nquads, ngroups, vec);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

extern "C" int sk_dvd_lpcm_unpack(sk_ctx* ctx, const uint8_t* packed, size_t bytes, int bits, int channels, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (bytes && (!packed || !out)) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, bytes + 16))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, bytes + 16))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, packed, bytes))) return rc;
    if ((rc = sk_dvd_lpcm_unpack_dev(ctx, (const uint8_t*)ctx->d_in, bytes, bits, channels, (uint8_t*)ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, bytes);
}

// ------------------------------------------------------------------------------------------------
// interleaved bytes -> planar (typed): the deinterleave family and both to-f32 conventions
// ------------------------------------------------------------------------------------------------
enum { CONV_F32_SCALED = 0, CONV_F32_BITS = 1, CONV_F32_BITS_SCRUB = 2, CONV_RAW_I32 = 3, CONV_RAW_I16 = 4 };

template <int CONV>
__device__ __forceinline__ uint32_t planar_apply(int32_t raw, float scale) {
    if (CONV == CONV_F32_SCALED) return __float_as_uint(__fmul_rn(__int2float_rn(raw), scale));
    if (CONV == CONV_F32_BITS) return (uint32_t)raw;
    if (CONV == CONV_F32_BITS_SCRUB) return finite_f(__int_as_float(raw)) ? (uint32_t)raw : 0u;  // lib.rs:3614
    if (CONV == CONV_RAW_I32) return (uint32_t)raw;
    return (uint32_t)raw & 0xffffu;
}

// grid.y = stream.  Thread = G frames of all CH channels.
template <int B, int CH, bool BE, int CONV>
__global__ void __launch_bounds__(kPcmThreads) deint_kernel(const uint8_t* __restrict__ in, size_t in_stride, uint8_t* __restrict__ out,
                                                             size_t out_stride_bytes, size_t frames, size_t ngroups, float scale) {
    constexpr int OB = CONV == CONV_RAW_I16 ? 2 : 4;
    // G frames: input G*CH*B bytes must be whole 16-byte vectors, per-channel output G*OB bytes too
    constexpr int G = (OB == 2) ? 8 : ((B == 3) ? 8 * (CH == 1 ? 2 : 1) : (B == 2 && CH == 1 ? 8 : 4));
    constexpr int NV_IN = G * CH * B / 16;
    static_assert(G * CH * B % 16 == 0 && G * OB % 16 == 0, "vector group");
    const uint8_t* sin = in + (size_t)blockIdx.y * in_stride;
    uint8_t* sout = out + (size_t)blockIdx.y * out_stride_bytes;
    const size_t stride = (size_t)gridDim.x * kPcmThreads;
    for (size_t g = (size_t)blockIdx.x * kPcmThreads + threadIdx.x; g < ngroups; g += stride) {
        uint32_t w[NV_IN * 4];
        const uint4* p = reinterpret_cast<const uint4*>(sin + g * (size_t)(G * CH * B));
#pragma unroll
        for (int v = 0; v < NV_IN; v++) {
            uint4 q = __ldg(p + v);
            w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            uint32_t o[G * OB / 4];
#pragma unroll
            for (int k = 0; k < G; k++) {
                uint32_t r = planar_apply<CONV>(get_sample<B, BE>(w, k * CH + c), scale);
                if (OB == 4) o[k] = r;
                else if (k & 1) o[k >> 1] |= r << 16;
                else o[k >> 1] = r;
            }
            uint4* q = reinterpret_cast<uint4*>(sout + ((size_t)c * frames + g * G) * OB);
#pragma unroll
            for (int v = 0; v < G * OB / 16; v++) st_stream_v4(q + v, make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]));
        }
    }
}

// Packed 24-bit stereo (BASELINE configs[1], s24le): a thread's natural unit (8 frames = 48 B in, 2 x 32 B out) makes
// every 128-bit access a strided, partially used sector.  Here a WARP owns 256 consecutive frames: three coalesced
// 512-byte loads land in warp-private shared memory, each lane then reads the 24 bytes of 4 frames (LDS.64 x 3, stride
// 24 B: conflict-free) and writes one 128-bit vector per channel, so every global access is a whole contiguous line.
template <bool BE, int CONV>
__global__ void __launch_bounds__(kPcmThreads) deint_s24_stereo_kernel(const uint8_t* __restrict__ in, size_t in_stride, uint8_t* __restrict__ out,
                                                                        size_t out_stride_bytes, size_t frames, size_t ntiles, float scale) {
    __shared__ __align__(16) uint8_t stage[kPcmThreads / 32][1536];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t* sin = in + (size_t)blockIdx.y * in_stride;
    uint8_t* sout = out + (size_t)blockIdx.y * out_stride_bytes;
    const size_t wstride = (size_t)gridDim.x * (kPcmThreads / 32);
    for (size_t t = (size_t)blockIdx.x * (kPcmThreads / 32) + warp; t < ntiles; t += wstride) {
        const uint4* p = reinterpret_cast<const uint4*>(sin + t * 1536);
        uint4 v0 = ld_stream_v4_nv(p + lane), v1 = ld_stream_v4_nv(p + 32 + lane), v2 = ld_stream_v4_nv(p + 64 + lane);
        __syncwarp();  // previous tile's reads of the staging buffer are done
        uint4* st = reinterpret_cast<uint4*>(stage[warp]);
        st[lane] = v0; st[32 + lane] = v1; st[64 + lane] = v2;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint2* q = reinterpret_cast<const uint2*>(stage[warp] + (i * 32 + lane) * 24);
            const uint2 a = q[0], b = q[1], c = q[2];
            const uint32_t w[7] = {a.x, a.y, b.x, b.y, c.x, c.y, 0u};
            uint32_t l[4], r[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                l[k] = planar_apply<CONV>(get_sample<3, BE>(w, 2 * k), scale);
                r[k] = planar_apply<CONV>(get_sample<3, BE>(w, 2 * k + 1), scale);
            }
            const size_t f0 = t * 256 + (size_t)(i * 32 + lane) * 4;
            st_stream_v4(sout + f0 * 4, make_uint4(l[0], l[1], l[2], l[3]));
            st_stream_v4(sout + (frames + f0) * 4, make_uint4(r[0], r[1], r[2], r[3]));
        }
    }
}

// One channel (16- / 24-bit in, 4-byte samples out): a lane's group is 2 or 4 output vectors, so the direct form above stores every
// 2nd / 4th vector of a run per instruction (half-written sectors: 0.75-0.78 of the HBM rate).  Same cure as the flat kernels: a
// WARP owns 32 consecutive groups and the vectors change lanes through warp-private shared memory (warp_load_spans /
// warp_store_spans), so every global access is a whole contiguous 512-byte run.  grid.y = stream.
template <int B, bool BE, int CONV>
__global__ void __launch_bounds__(kPcmThreads) deint_mono_kernel(const uint8_t* __restrict__ in, size_t in_stride, uint8_t* __restrict__ out,
                                                                  size_t out_stride_bytes, size_t ntiles, float scale) {
    constexpr int G = B == 3 ? 16 : 8;
    constexpr int NV_IN = G * B / 16, NV_OUT = G * 4 / 16;
    constexpr int NVX = NV_IN > NV_OUT ? NV_IN : NV_OUT;
    __shared__ uint4 xch_all[kPcmWarps][NVX * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* xch = xch_all[warp];
    const uint8_t* sin = in + (size_t)blockIdx.y * in_stride;
    uint8_t* sout = out + (size_t)blockIdx.y * out_stride_bytes;
    for (size_t t = (size_t)blockIdx.x * kPcmWarps + warp; t < ntiles; t += (size_t)gridDim.x * kPcmWarps) {
        uint32_t w[NV_IN * 4 + 1];
        warp_load_spans<NV_IN>(xch, reinterpret_cast<const uint4*>(sin + t * (size_t)(32 * G * B)), w, lane);
        w[NV_IN * 4] = 0u;
        uint32_t o[G];
#pragma unroll
        for (int k = 0; k < G; k++) o[k] = planar_apply<CONV>(get_sample<B, BE>(w, k), scale);
        warp_store_spans<NV_OUT>(xch, reinterpret_cast<uint4*>(sout + t * (size_t)(32 * G * 4)), o, lane);
    }
}

// 3 .. 16 channels (5.1 / 7.1 beds): a CTA copies a tile of 256 frames into shared memory with coalesced 128-bit loads, then thread f
// walks the channels of frame f, so that every warp store is 128 contiguous bytes of one channel plane.  Reading sample (f, c)
// from 32 lanes is a stride of CH * B bytes; where that is a multiple of 16 bytes (all 32 reads in 8 or fewer banks) the tile is
// stored XOR-swizzled — word w at w ^ ((w >> 5) & swz), a permutation inside each 32-word row — which spreads the lanes over
// all banks again.  (The one-thread-per-sample kernel below ran these shapes at 0.22-0.27 of the HBM rate.)
constexpr int kTileFrames = 256;
constexpr int kTileMaxCh = 16;
template <int B, bool BE, int CONV>
__global__ void __launch_bounds__(kPcmThreads) deint_tile_kernel(const uint8_t* __restrict__ in, size_t in_stride, uint8_t* __restrict__ out,
                                                                  size_t out_stride_bytes, size_t frames, size_t ntiles, int CH, int swz, int fpt, float scale) {
    static_assert(kPcmThreads == kTileFrames, "one thread per frame of a tile row");
    extern __shared__ __align__(16) uint32_t deint_tile[];  // fpt * kTileFrames * CH * B bytes + 16 spare (the 24-bit reader looks one word ahead)
    const uint8_t* sin = in + (size_t)blockIdx.y * in_stride;
    uint32_t* sout = reinterpret_cast<uint32_t*>(out + (size_t)blockIdx.y * out_stride_bytes);
    const int tile_frames = fpt * kTileFrames;  // few channels: several frames per thread, so a tile is worth its two barriers
    const int nvec = tile_frames * CH * B / 16;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(sin + t * (size_t)tile_frames * (size_t)(CH * B));
        for (int v = threadIdx.x; v < nvec; v += kPcmThreads) {
            const uint4 q = ld_stream_v4_nv(p + v);
            const int w = 4 * v;
            if (B != 3 && swz) {
                const int k = (w >> 5) & swz;  // the four words of a vector share a 32-word row
                deint_tile[w ^ k] = q.x; deint_tile[(w + 1) ^ k] = q.y; deint_tile[(w + 2) ^ k] = q.z; deint_tile[(w + 3) ^ k] = q.w;
            } else {
                reinterpret_cast<uint4*>(deint_tile)[v] = q;
            }
        }
        __syncthreads();
        for (int j = 0; j < fpt; j++) {
        const int f = j * kTileFrames + threadIdx.x;
        uint32_t* o = sout + t * (size_t)tile_frames + f;
        for (int c = 0; c < CH; c++) {
            const int bo = (f * CH + c) * B;
            int w = bo >> 2;
            int32_t raw;
            if (B == 3) {
                const uint32_t u = __funnelshift_r(deint_tile[w], deint_tile[w + 1], 8 * (bo & 3));
                raw = ((int32_t)(BE ? __byte_perm(u, 0, 0x0124) : __byte_perm(u, 0, 0x2104))) >> 8;
            } else {
                w ^= (w >> 5) & swz;
                uint32_t u = deint_tile[w];
                if (B == 2) {
                    u = (u >> (8 * (bo & 2))) & 0xffffu;
                    if (BE) u = __byte_perm(u, 0, 0x4401);
                    raw = (int32_t)(int16_t)u;
                } else {
                    raw = (int32_t)(BE ? bswap32(u) : u);
                }
            }
            o[(size_t)c * frames] = planar_apply<CONV>(raw, scale);
        }
        }
        __syncthreads();  // the tile is free for the next copy
    }
}

// any channel count / alignment / tail frames: one thread per sample
__global__ void deint_scalar_kernel(const uint8_t* __restrict__ in, size_t in_stride, uint8_t* __restrict__ out, size_t out_stride_bytes,
                                    size_t frames, size_t plane, size_t first_frame, int B, int CH, int be, int conv, float scale) {
    const uint8_t* sin = in + (size_t)blockIdx.y * in_stride;
    uint8_t* sout = out + (size_t)blockIdx.y * out_stride_bytes;
    const size_t total = (frames - first_frame) * (size_t)CH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t f = first_frame + i / CH; int c = (int)(i % CH);
        const uint8_t* p = sin + (f * CH + c) * (size_t)B;
        uint32_t u = 0;
        if (!be) { for (int b = 0; b < B; b++) u |= (uint32_t)p[b] << (8 * b); }
        else     { for (int b = 0; b < B; b++) u = (u << 8) | p[b]; }
        int32_t raw = B == 2 ? (int32_t)(int16_t)u : (B == 3 ? ((int32_t)(u << 8)) >> 8 : (int32_t)u);
        uint32_t r;
        switch (conv) {
            case CONV_F32_SCALED: r = planar_apply<CONV_F32_SCALED>(raw, scale); break;
            case CONV_F32_BITS: r = planar_apply<CONV_F32_BITS>(raw, scale); break;
            case CONV_F32_BITS_SCRUB: r = planar_apply<CONV_F32_BITS_SCRUB>(raw, scale); break;
            case CONV_RAW_I32: r = planar_apply<CONV_RAW_I32>(raw, scale); break;
            default: r = planar_apply<CONV_RAW_I16>(raw, scale); break;
        }
        if (conv == CONV_RAW_I16) reinterpret_cast<uint16_t*>(sout)[(size_t)c * plane + f] = (uint16_t)r;
        else reinterpret_cast<uint32_t*>(sout)[(size_t)c * plane + f] = r;
    }
}

template <int B, int CH, bool BE, int CONV>
static size_t launch_deint_vec(sk_ctx* ctx, const uint8_t* in, size_t in_stride, uint8_t* out, size_t out_stride_bytes, size_t frames, size_t plane,
                               size_t n_streams, float scale) {  // `plane`: elements between channel planes (the kernels' `frames` argument is that stride)
    constexpr int OB = CONV == CONV_RAW_I16 ? 2 : 4;
    constexpr int G = (OB == 2) ? 8 : ((B == 3) ? 8 * (CH == 1 ? 2 : 1) : (B == 2 && CH == 1 ? 8 : 4));
    if (B == 3 && CH == 2 && OB == 4 && frames >= 256) {  // warp-tiled path for packed 24-bit stereo
        const size_t ntiles = frames / 256;
        const size_t ctas = (ntiles + kPcmThreads / 32 - 1) / (kPcmThreads / 32);
        dim3 grid((unsigned)std::max<size_t>(1, std::min<size_t>((ctas + 1) / 2, (size_t)1 << 22)), (unsigned)n_streams);  // short-lived CTAs (grid_for)
        deint_s24_stereo_kernel<BE, CONV><<<grid, kPcmThreads, 0, ctx->stream>>>(in, in_stride, out, out_stride_bytes, plane, ntiles, scale);
        return ntiles * 256;
    }
    size_t ngroups = frames / G;
    if (!ngroups) return 0;
    dim3 grid(grid_for(ctx, ngroups, kPcmThreads), (unsigned)n_streams);
    // keep total CTAs bounded for big batches: the grid-stride loop covers the rest
    if (n_streams > 1) grid.x = (unsigned)((ngroups + (size_t)kPcmThreads * 8 - 1) / ((size_t)kPcmThreads * 8));
    if (grid.x < 1) grid.x = 1;
    deint_kernel<B, CH, BE, CONV><<<grid, kPcmThreads, 0, ctx->stream>>>(in, in_stride, out, out_stride_bytes, plane, ngroups, scale);
    return ngroups * G;
}

// dispatch on (B, CH, BE, CONV); returns frames covered by the vector kernel (0 = none)
static size_t dispatch_deint_vec(sk_ctx* ctx, int B, int CH, bool be, int conv, const uint8_t* in, size_t in_stride, uint8_t* out,
                                 size_t out_stride_bytes, size_t frames, size_t plane, size_t n_streams, float scale) {
    if (CH == 1 && (B == 2 || B == 3) && conv == CONV_F32_SCALED) {  // mono 16- / 24-bit -> f32: warp-tiled stores
        const size_t tile = (size_t)32 * (B == 3 ? 16 : 8);
        const size_t ntiles = frames / tile;
        if (ntiles) {
            const size_t ctas = (ntiles + kPcmWarps - 1) / kPcmWarps;
            const dim3 grid((unsigned)std::max<size_t>(1, std::min<size_t>((ctas + 1) / 2, (size_t)1 << 22)), (unsigned)n_streams);
            if (B == 2) { if (be) deint_mono_kernel<2, true, CONV_F32_SCALED><<<grid, kPcmThreads, 0, ctx->stream>>>(in, in_stride, out, out_stride_bytes, ntiles, scale);
                          else deint_mono_kernel<2, false, CONV_F32_SCALED><<<grid, kPcmThreads, 0, ctx->stream>>>(in, in_stride, out, out_stride_bytes, ntiles, scale); }
            else { if (be) deint_mono_kernel<3, true, CONV_F32_SCALED><<<grid, kPcmThreads, 0, ctx->stream>>>(in, in_stride, out, out_stride_bytes, ntiles, scale);
                   else deint_mono_kernel<3, false, CONV_F32_SCALED><<<grid, kPcmThreads, 0, ctx->stream>>>(in, in_stride, out, out_stride_bytes, ntiles, scale); }
            return ntiles * tile;
        }
    }
    if (CH >= 3 && CH <= kTileMaxCh && conv != CONV_RAW_I16 && frames >= (size_t)kTileFrames) {
        const int row = CH * B;  // bytes between the same channel of consecutive frames
        const int fpt = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(4, 64 / row), frames / kTileFrames));  // <= 16 KB of tile
        const size_t ntiles = frames / ((size_t)fpt * kTileFrames);
        const int swz = (B != 3 && row % 16 == 0) ? 7 : 0;
        const size_t smem = (size_t)fpt * kTileFrames * row + 16;
        const dim3 grid((unsigned)std::min<size_t>(ntiles, (size_t)1 << 20), (unsigned)n_streams);
#define SK_DT(B_, BE_, CONV_) \
        if (B == B_ && be == BE_ && conv == CONV_) { \
            deint_tile_kernel<B_, BE_, CONV_><<<grid, kPcmThreads, smem, ctx->stream>>>(in, in_stride, out, out_stride_bytes, plane, ntiles, CH, swz, fpt, scale); \
            return ntiles * (size_t)fpt * kTileFrames; \
        }
        SK_DT(2, false, CONV_F32_SCALED) SK_DT(2, true, CONV_F32_SCALED) SK_DT(3, false, CONV_F32_SCALED) SK_DT(3, true, CONV_F32_SCALED)
        SK_DT(4, false, CONV_F32_SCALED) SK_DT(4, true, CONV_F32_SCALED) SK_DT(4, false, CONV_F32_BITS) SK_DT(4, true, CONV_F32_BITS)
        SK_DT(4, false, CONV_F32_BITS_SCRUB) SK_DT(4, true, CONV_F32_BITS_SCRUB) SK_DT(3, false, CONV_RAW_I32)
#undef SK_DT
    }
#define SK_DV(B_, CH_, BE_, CONV_) \
    if (B == B_ && CH == CH_ && be == BE_ && conv == CONV_) \
        return launch_deint_vec<B_, CH_, BE_, CONV_>(ctx, in, in_stride, out, out_stride_bytes, frames, plane, n_streams, scale);
#define SK_DV_CH(B_, BE_, CONV_) SK_DV(B_, 1, BE_, CONV_) SK_DV(B_, 2, BE_, CONV_)
    SK_DV_CH(2, false, CONV_F32_SCALED) SK_DV_CH(2, true, CONV_F32_SCALED)
    SK_DV_CH(3, false, CONV_F32_SCALED) SK_DV_CH(3, true, CONV_F32_SCALED)
    SK_DV_CH(4, false, CONV_F32_SCALED) SK_DV_CH(4, true, CONV_F32_SCALED)
    SK_DV_CH(4, false, CONV_F32_BITS) SK_DV_CH(4, true, CONV_F32_BITS)
    SK_DV_CH(4, false, CONV_F32_BITS_SCRUB) SK_DV_CH(4, true, CONV_F32_BITS_SCRUB)
    SK_DV_CH(3, false, CONV_RAW_I32)
    SK_DV_CH(2, false, CONV_RAW_I16)
#undef SK_DV_CH
#undef SK_DV
    return 0;
}

// common driver: vector kernel on the aligned body, scalar kernel on everything else
static int run_deint(sk_ctx* ctx, int B, int CH, bool be, int conv, const uint8_t* in, size_t in_stride, void* outv, size_t out_stride_bytes,
                     size_t frames, size_t n_streams, float scale, size_t plane = 0) {  // plane: elements between channel planes (0 = frames)
    if (!plane) plane = frames;
    if (!frames || !n_streams) return SK_OK;
    uint8_t* out = (uint8_t*)outv;
    const int OB = conv == CONV_RAW_I16 ? 2 : 4;
    size_t done = 0;
    bool vec_ok = aligned16(in) && aligned16(out) && (in_stride % 16 == 0) && (out_stride_bytes % 16 == 0) && ((plane * OB) % 16 == 0) &&
                  n_streams <= 65535;
    if (vec_ok) {
        done = dispatch_deint_vec(ctx, B, CH, be, conv, in, in_stride, out, out_stride_bytes, frames, plane, n_streams, scale);
        if (done) SK_LAUNCH_CHECK(ctx);
    }
    if (done < frames) {
        for (size_t s0 = 0; s0 < n_streams; s0 += 65535) {
            size_t ns = n_streams - s0 < 65535 ? n_streams - s0 : 65535;
            dim3 grid(grid_for(ctx, (frames - done) * CH, 256), (unsigned)ns);
            if (ns > 1 && grid.x > 64) grid.x = 64;
            deint_scalar_kernel<<<grid, 256, 0, ctx->stream>>>(in + s0 * in_stride, in_stride, out + s0 * out_stride_bytes, out_stride_bytes,
                                                               frames, plane, done, B, CH, be ? 1 : 0, conv, scale);
            SK_LAUNCH_CHECK(ctx);
        }
    }
    return SK_OK;
}

// (bits, fmt, mode) -> conversion + scale; returns 0 when the combination is not valid
static int planar_f32_conv(int bits, int fmt, int big_endian, int mode, int* conv, float* scale, bool* be) {
    *be = mode == SK_MODE_B ? big_endian != 0 : false;  // mode A ignores endianness() (audio_pipeline.rs:74-98)
    if (bits == 16) { *conv = CONV_F32_SCALED; *scale = 1.0f / 32768.0f; return 1; }
    if (bits == 24) { *conv = CONV_F32_SCALED; *scale = mode == SK_MODE_A ? 4.656612873077393e-10f /* 2^-31 (sic) */ : 1.0f / 8388608.0f; return 1; }
    if (bits == 32 && fmt == SK_FMT_FLOAT) { *conv = mode == SK_MODE_A ? CONV_F32_BITS : CONV_F32_BITS_SCRUB; *scale = 1.0f; return 1; }
    if (bits == 32) { *conv = CONV_F32_SCALED; *scale = 4.656612873077393e-10f; return 1; }
    return 0;
}

extern "C" int sk_pcm_to_planar_f32_dev(sk_ctx* ctx, const uint8_t* in, size_t in_stride, size_t n_streams, size_t frames, int bits, int fmt,
                                        int big_endian, int channels, int scale_mode, float* out, size_t out_stride) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    int conv; float scale; bool be;
    if (!planar_f32_conv(bits, fmt, big_endian, scale_mode, &conv, &scale, &be)) SK_FAIL(ctx, SK_E_ARG, "unsupported bits per sample %d", bits);
    if (fmt == SK_FMT_FLOAT && bits != 32) SK_FAIL(ctx, SK_E_ARG, "floating-point PCM must contain 32-bit samples");
    if (!frames || !n_streams) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_deint(ctx, bits / 8, channels, be, conv, in, in_stride, out, out_stride * sizeof(float), frames, n_streams, scale);
}

// internal (sk_resample.cu: the two-pass chain): as sk_pcm_to_planar_f32_dev, channel planes `plane` floats apart (>= frames) so that every
// plane can start on a 16-byte boundary whatever the stream length
int sk_pcm_to_planar_rows_dev(sk_ctx* ctx, const uint8_t* in, size_t in_stride, size_t n_streams, size_t frames, int bits, int fmt, int big_endian,
                              int channels, int scale_mode, float* out, size_t out_stride, size_t plane) {
    int conv; float scale; bool be;
    if (!planar_f32_conv(bits, fmt, big_endian, scale_mode, &conv, &scale, &be)) SK_FAIL(ctx, SK_E_ARG, "unsupported bits per sample %d", bits);
    if (!frames || !n_streams) return SK_OK;
    return run_deint(ctx, bits / 8, channels, be, conv, in, in_stride, out, out_stride * sizeof(float), frames, n_streams, scale, plane);
}

extern "C" int sk_pcm_to_planar_f32(sk_ctx* ctx, const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian, int channels, int scale_mode,
                                    float* out, size_t frames_cap, size_t* frames_out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "unsupported bits per sample %d", bits);
    size_t B = (size_t)bits / 8;
    if (scale_mode == SK_MODE_B && (bytes % B != 0 || (bytes / B) % (size_t)channels != 0))
        SK_FAIL(ctx, SK_E_ARG, "PCM data is unsupported or contains a partial frame");
    size_t frames = bytes / (B * channels);
    if (frames_out) *frames_out = frames;
    if (frames > frames_cap) SK_FAIL(ctx, SK_E_ARG, "output capacity %zu < %zu frames", frames_cap, frames);
    if (!frames) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t in_bytes = frames * B * channels, out_bytes = frames * channels * sizeof(float);
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, out_bytes))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, in_bytes))) return rc;
    if ((rc = sk_pcm_to_planar_f32_dev(ctx, (const uint8_t*)ctx->d_in, in_bytes, 1, frames, bits, fmt, big_endian, channels, scale_mode,
                                       (float*)ctx->d_out, frames * channels)))
        return rc;
    // planes are `frames` apart on the device; the caller's planes are frames_cap apart
    for (int c = 0; c < channels; c++)
        if ((rc = sk_d2h(ctx, out + (size_t)c * frames_cap, (float*)ctx->d_out + (size_t)c * frames, frames * sizeof(float)))) return rc;
    return SK_OK;
}

extern "C" int sk_deinterleave(sk_ctx* ctx, int kind, const uint8_t* in, size_t bytes, int channels, void* out, size_t* frames_out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    int B, conv;
    if (kind == 0) { B = 2; conv = CONV_RAW_I16; }
    else if (kind == 1) { B = 3; conv = CONV_RAW_I32; }
    else if (kind == 2) { B = 4; conv = CONV_F32_BITS; }
    else SK_FAIL(ctx, SK_E_ARG, "unknown deinterleave kind %d", kind);
    const int OB = conv == CONV_RAW_I16 ? 2 : 4;
    size_t frames = bytes / ((size_t)B * channels);  // chunks_exact: trailing partial frame dropped
    if (frames_out) *frames_out = frames;
    if (!frames) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t in_bytes = frames * B * channels, out_bytes = frames * channels * OB;
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, out_bytes))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, in_bytes))) return rc;
    if ((rc = run_deint(ctx, B, channels, false, conv, (const uint8_t*)ctx->d_in, in_bytes, ctx->d_out, out_bytes, frames, 1, 1.0f))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, out_bytes);
}

// ------------------------------------------------------------------------------------------------
// planar f32 -> interleaved PCM bytes (f32_channels_to_bytes) and the small helpers
// ------------------------------------------------------------------------------------------------
template <int BITS, bool ISFLOAT>
__device__ __forceinline__ uint32_t quantize(float x) {
    if (ISFLOAT) return __float_as_uint(x);
    if (BITS == 16) return (uint32_t)float_sample_to_i16_dev(x) & 0xffffu;
    if (BITS == 24) return (uint32_t)f32_to_s24_dev(x) & 0xffffffu;
    return (uint32_t)f32_to_i32_dev(x);
}

// lane = G frames x CH channels; output span G*CH*(BITS/8) bytes is whole 16-byte vectors; a warp owns 32 groups
template <int BITS, int CH, bool ISFLOAT>
__global__ void __launch_bounds__(kPcmThreads) interleave_kernel(const float* __restrict__ planar, size_t frames, uint8_t* __restrict__ out,
                                                                  size_t ntiles) {
    constexpr int OB = BITS / 8;
    constexpr int G = (OB == 3) ? (CH == 1 ? 16 : 8) : (OB == 2 ? (CH == 1 ? 16 : 8) : 4);
    constexpr int NW = G * CH * OB / 4, NV_OUT = NW / 4, NV_IN = G / 4;
    constexpr int NVX = NV_IN > NV_OUT ? NV_IN : NV_OUT;
    static_assert((G * CH * OB) % 16 == 0 && G % 4 == 0, "vector group");
    __shared__ uint4 xch_all[kPcmWarps][NVX * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* xch = xch_all[warp];
    for (size_t t = (size_t)blockIdx.x * kPcmWarps + warp; t < ntiles; t += (size_t)gridDim.x * kPcmWarps) {
        uint32_t xb[CH][G];
#pragma unroll
        for (int c = 0; c < CH; c++) warp_load_spans<NV_IN>(xch, reinterpret_cast<const uint4*>(planar + (size_t)c * frames + t * (size_t)(32 * G)), xb[c], lane);
        uint32_t o[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) o[i] = 0;
#pragma unroll
        for (int k = 0; k < G; k++)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const uint32_t r = quantize<BITS, ISFLOAT>(__uint_as_float(xb[c][k]));
                const int byte = (k * CH + c) * OB, wi = byte >> 2, sh = (byte & 3) * 8;
                o[wi] |= r << sh;
                if (sh + OB * 8 > 32) o[wi + 1] |= r >> (32 - sh);
            }
        warp_store_spans<NV_OUT>(xch, reinterpret_cast<uint4*>(out + t * (size_t)(32 * G * CH * OB)), o, lane);
    }
}

__global__ void interleave_scalar_kernel(const float* __restrict__ planar, size_t frames, size_t first_frame, int CH, int bits, int isfloat,
                                         uint8_t* __restrict__ out) {
    const size_t total = (frames - first_frame) * (size_t)CH;
    const int OB = bits / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t f = first_frame + i / CH; int c = (int)(i % CH);
        float x = planar[(size_t)c * frames + f];
        uint32_t r = isfloat ? quantize<32, true>(x) : (bits == 16 ? quantize<16, false>(x) : (bits == 24 ? quantize<24, false>(x) : quantize<32, false>(x)));
        uint8_t* p = out + (f * CH + c) * (size_t)OB;
        for (int b = 0; b < OB; b++) p[b] = (uint8_t)(r >> (8 * b));
    }
}

// 3 .. 16 channels: the mirror image of deint_tile_kernel.  Thread f reads sample f of every channel plane (each warp load is 128
// contiguous bytes of one plane), quantises and drops the bytes at their interleaved place in a shared-memory tile (XOR-swizzled
// where the frame size is a multiple of 16 bytes); the CTA then writes the tile with coalesced 128-bit stores.
template <int BITS, bool ISFLOAT>
__global__ void __launch_bounds__(kPcmThreads) interleave_tile_kernel(const float* __restrict__ planar, size_t frames, uint8_t* __restrict__ out, size_t ntiles,
                                                                      int CH, int swz, int fpt) {
    constexpr int OB = BITS / 8;
    extern __shared__ __align__(16) uint32_t il_tile[];
    uint8_t* tile8 = reinterpret_cast<uint8_t*>(il_tile);
    const int tile_frames = fpt * kTileFrames;
    const int nvec = tile_frames * CH * OB / 16;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int j = 0; j < fpt; j++) {
            const int f = j * kTileFrames + threadIdx.x;
            const float* x = planar + t * (size_t)tile_frames + f;
            for (int c = 0; c < CH; c++) {
                const uint32_t r = quantize<BITS, ISFLOAT>(__ldg(x + (size_t)c * frames));
                const int bo = (f * CH + c) * OB;
                if (OB == 3) {
                    tile8[bo] = (uint8_t)r; tile8[bo + 1] = (uint8_t)(r >> 8); tile8[bo + 2] = (uint8_t)(r >> 16);
                } else {
                    int w = bo >> 2;
                    w ^= (w >> 5) & swz;
                    if (OB == 4) il_tile[w] = r;
                    else reinterpret_cast<uint16_t*>(il_tile)[2 * w + ((bo >> 1) & 1)] = (uint16_t)r;
                }
            }
        }
        __syncthreads();
        uint4* q = reinterpret_cast<uint4*>(out + t * (size_t)tile_frames * (size_t)(CH * OB));
        for (int v = threadIdx.x; v < nvec; v += kPcmThreads) {
            const int w = 4 * v;
            const int k = OB == 3 ? 0 : ((w >> 5) & swz);
            st_stream_v4(q + v, make_uint4(il_tile[w ^ k], il_tile[(w + 1) ^ k], il_tile[(w + 2) ^ k], il_tile[(w + 3) ^ k]));
        }
        __syncthreads();  // the tile is free for the next frames
    }
}
template <int BITS, bool ISFLOAT>
static size_t launch_interleave_tile(sk_ctx* ctx, const float* planar, size_t frames, int channels, uint8_t* out) {
    constexpr int OB = BITS / 8;
    const int row = channels * OB;
    const int fpt = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(4, 64 / row), frames / kTileFrames));
    const size_t ntiles = frames / ((size_t)fpt * kTileFrames);
    if (!ntiles) return 0;
    const int swz = (OB != 3 && row % 16 == 0) ? 7 : 0;
    const size_t smem = (size_t)fpt * kTileFrames * row + 16;
    interleave_tile_kernel<BITS, ISFLOAT><<<(unsigned)std::min<size_t>(ntiles, (size_t)1 << 20), kPcmThreads, smem, ctx->stream>>>(planar, frames, out, ntiles, channels,
                                                                                                                                swz, fpt);
    return ntiles * (size_t)fpt * kTileFrames;
}

template <int BITS, int CH, bool ISFLOAT>
static size_t launch_interleave_vec(sk_ctx* ctx, const float* planar, size_t frames, uint8_t* out) {
    constexpr int OB = BITS / 8;
    constexpr int G = (OB == 3) ? (CH == 1 ? 16 : 8) : (OB == 2 ? (CH == 1 ? 16 : 8) : 4);
    const size_t ntiles = frames / (32 * (size_t)G);
    if (!ntiles) return 0;
    interleave_kernel<BITS, CH, ISFLOAT><<<grid_for(ctx, ntiles * 32, kPcmThreads), kPcmThreads, 0, ctx->stream>>>(planar, frames, out, ntiles);
    return ntiles * 32 * G;
}

extern "C" int sk_planar_f32_to_pcm_dev(sk_ctx* ctx, const float* planar, size_t frames, int channels, int bits, int fmt, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (fmt == SK_FMT_FLOAT && bits != 32) SK_FAIL(ctx, SK_E_ARG, "PCMFloat output requires 32-bit samples");
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "Unsupported output bits per sample: %d", bits);
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool isf = fmt == SK_FMT_FLOAT;
    size_t done = 0;
    if (aligned16(planar) && aligned16(out) && frames % 4 == 0 && channels <= 2) {
#define SK_IV(BITS_, CH_, F_) if (bits == BITS_ && channels == CH_ && isf == F_) done = launch_interleave_vec<BITS_, CH_, F_>(ctx, planar, frames, out);
        SK_IV(16, 1, false) SK_IV(16, 2, false) SK_IV(24, 1, false) SK_IV(24, 2, false)
        SK_IV(32, 1, false) SK_IV(32, 2, false) SK_IV(32, 1, true) SK_IV(32, 2, true)
#undef SK_IV
        if (done) SK_LAUNCH_CHECK(ctx);
    } else if (aligned16(out) && channels >= 3 && channels <= kTileMaxCh) {
        if (isf) done = launch_interleave_tile<32, true>(ctx, planar, frames, channels, out);
        else if (bits == 16) done = launch_interleave_tile<16, false>(ctx, planar, frames, channels, out);
        else if (bits == 24) done = launch_interleave_tile<24, false>(ctx, planar, frames, channels, out);
        else done = launch_interleave_tile<32, false>(ctx, planar, frames, channels, out);
        if (done) SK_LAUNCH_CHECK(ctx);
    }
    if (done < frames) {
        interleave_scalar_kernel<<<grid_for(ctx, (frames - done) * channels, 256), 256, 0, ctx->stream>>>(planar, frames, done, channels, bits, isf ? 1 : 0, out);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

extern "C" int sk_planar_f32_to_pcm(sk_ctx* ctx, const float* planar, size_t frames, int channels, int bits, int fmt, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (fmt == SK_FMT_FLOAT && bits != 32) SK_FAIL(ctx, SK_E_ARG, "PCMFloat output requires 32-bit samples");
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "Unsupported output bits per sample: %d", bits);
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t in_bytes = frames * channels * sizeof(float), out_bytes = frames * channels * (size_t)(bits / 8);
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, out_bytes))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, planar, in_bytes))) return rc;
    if ((rc = sk_planar_f32_to_pcm_dev(ctx, (const float*)ctx->d_in, frames, channels, bits, fmt, (uint8_t*)ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, out_bytes);
}

// exact_signed_pcm_to_i16 (decoder lib.rs:3458-3489) = flat narrowing ops on the interleaved bytes
extern "C" int sk_pcm_narrow_i16(sk_ctx* ctx, const uint8_t* in, size_t bytes, int bits, int big_endian, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "exact narrowing takes 24- or 32-bit signed PCM");
    if (bytes % (size_t)(bits / 8)) SK_FAIL(ctx, SK_E_ARG, "integer PCM contains a partial sample");
    int op = bits == 24 ? (big_endian ? SK_OP_S24BE_TO_I16 : SK_OP_S24LE_TO_I16) : (big_endian ? SK_OP_S32BE_TO_I16 : SK_OP_S32LE_TO_I16);
    return sk_pcm_flat_convert(ctx, op, in, bytes, out, nullptr);
}

// ---- mono mixdown ---------------------------------------------------------------------------------
// blockIdx.y = stream: a whole batch is one launch (in_stream / out_stream = floats between consecutive streams)
__global__ void __launch_bounds__(kPcmThreads) mixdown_kernel(const float* __restrict__ planar_all, size_t in_stream, size_t plane, size_t frames,
                                                               int channels, int mode, float scale, float* __restrict__ out_all, size_t out_stream, int vec) {
    const float* planar = planar_all + (size_t)blockIdx.y * in_stream;
    float* out = out_all + (size_t)blockIdx.y * out_stream;
    const size_t stride = (size_t)gridDim.x * kPcmThreads;
    if (vec) {
        const size_t n4 = frames / 4;
        for (size_t g = (size_t)blockIdx.x * kPcmThreads + threadIdx.x; g < n4; g += stride) {
            float m[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < channels; c++) {
                float4 q = __ldg(reinterpret_cast<const float4*>(planar + (size_t)c * plane) + g);
                float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; k++) m[k] = mode == SK_MIX_SCALE_THEN_SUM ? __fadd_rn(m[k], __fmul_rn(x[k], scale)) : __fadd_rn(m[k], x[k]);
            }
            if (mode == SK_MIX_SUM_THEN_SCALE)
#pragma unroll
                for (int k = 0; k < 4; k++) m[k] = __fmul_rn(m[k], scale);
            reinterpret_cast<float4*>(out)[g] = make_float4(m[0], m[1], m[2], m[3]);
        }
    }
    const size_t first = vec ? frames / 4 * 4 : 0;
    for (size_t i = first + (size_t)blockIdx.x * kPcmThreads + threadIdx.x; i < frames; i += stride) {
        float m = 0.f;
        for (int c = 0; c < channels; c++) {
            float x = planar[(size_t)c * plane + i];
            m = mode == SK_MIX_SCALE_THEN_SUM ? __fadd_rn(m, __fmul_rn(x, scale)) : __fadd_rn(m, x);
        }
        if (mode == SK_MIX_SUM_THEN_SCALE) m = __fmul_rn(m, scale);
        out[i] = m;
    }
}

// n_streams planar blocks -> n_streams mono rows in one launch (the unfused arms of sk_normalize_batch_dev)
int sk_mixdown_batch_dev(sk_ctx* ctx, const float* planar, size_t in_stream_stride, size_t plane_stride, size_t frames, int channels, int mix_mode,
                         float* out, size_t out_stream_stride, size_t n_streams) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (mix_mode != SK_MIX_SUM_THEN_SCALE && mix_mode != SK_MIX_SCALE_THEN_SUM) SK_FAIL(ctx, SK_E_ARG, "unknown mix mode %d", mix_mode);
    if (!frames || !n_streams) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (channels == 1 && mix_mode == SK_MIX_SUM_THEN_SCALE) {  // mixdown_to_mono_f32 returns channels[0].clone()
        SK_CUDA(ctx, cudaMemcpy2DAsync(out, out_stream_stride * sizeof(float), planar, in_stream_stride * sizeof(float), frames * sizeof(float),
                                       n_streams, cudaMemcpyDeviceToDevice, ctx->stream));
        return SK_OK;
    }
    const float scale = 1.0f / (float)channels;
    const int vec = aligned16(planar) && aligned16(out) && plane_stride % 4 == 0 && (n_streams == 1 || (in_stream_stride % 4 == 0 && out_stream_stride % 4 == 0));
    const unsigned gx = (unsigned)std::min<int>(grid_for(ctx, frames / 4 + 1, kPcmThreads), 1 << 20);
    for (size_t s0 = 0; s0 < n_streams; s0 += 65535) {
        const size_t ns = std::min<size_t>(65535, n_streams - s0);
        mixdown_kernel<<<dim3(gx, (unsigned)ns), kPcmThreads, 0, ctx->stream>>>(planar + s0 * in_stream_stride, in_stream_stride, plane_stride, frames, channels,
                                                                               mix_mode, scale, out + s0 * out_stream_stride, out_stream_stride, vec);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

int sk_mixdown_strided_dev(sk_ctx* ctx, const float* planar, size_t plane_stride, size_t frames, int channels, int mix_mode, float* out) {
    return sk_mixdown_batch_dev(ctx, planar, 0, plane_stride, frames, channels, mix_mode, out, 0, 1);
}

// mono f32 rows -> bytes as f32_channels_to_bytes writes one channel (soundkit-decoder/src/lib.rs:3619-3683), a whole batch per
// launch: blockIdx.y = stream, 4 samples per thread where the row pointers allow 128-bit loads
__global__ void __launch_bounds__(256) quantize_mono_batch_kernel(const float* __restrict__ in_all, size_t in_stream, size_t n, int bits, int isfloat,
                                                                   uint8_t* __restrict__ out_all, size_t out_stream_bytes, int vec) {
    const float* in = in_all + (size_t)blockIdx.y * in_stream;
    uint8_t* out = out_all + (size_t)blockIdx.y * out_stream_bytes;
    const int ob = bits >> 3;
    const size_t stride = (size_t)gridDim.x * 256;
    auto q1 = [&](float x) -> uint32_t {
        return isfloat ? quantize<32, true>(x) : (bits == 16 ? quantize<16, false>(x) : (bits == 24 ? quantize<24, false>(x) : quantize<32, false>(x)));
    };
    const size_t n4 = vec ? n / 4 : 0;
    for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < n4; g += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + g);
        const uint32_t q[4] = {q1(v.x), q1(v.y), q1(v.z), q1(v.w)};
        if (ob == 2) *reinterpret_cast<uint2*>(out + g * 8) = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
        else if (ob == 3) {
            uint32_t* w = reinterpret_cast<uint32_t*>(out + g * 12);
            w[0] = q[0] | (q[1] << 24); w[1] = (q[1] >> 8) | (q[2] << 16); w[2] = (q[2] >> 16) | (q[3] << 8);
        } else *reinterpret_cast<uint4*>(out + g * 16) = make_uint4(q[0], q[1], q[2], q[3]);
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const uint32_t r = q1(in[i]);
        for (int b = 0; b < ob; b++) out[i * ob + b] = (uint8_t)(r >> (8 * b));
    }
}

int sk_quantize_mono_batch_dev(sk_ctx* ctx, const float* in, size_t in_stream_stride, size_t n, int bits, int fmt, uint8_t* out, size_t out_stream_bytes,
                               size_t n_streams) {
    if (!n || !n_streams) return SK_OK;
    const int vec = aligned16(in) && in_stream_stride % 4 == 0 && ((uintptr_t)out % 16) == 0 && out_stream_bytes % 16 == 0;
    const unsigned gx = (unsigned)std::min<int>(grid_for(ctx, n / 4 + 1, 256), 1 << 20);
    for (size_t s0 = 0; s0 < n_streams; s0 += 65535) {
        const size_t ns = std::min<size_t>(65535, n_streams - s0);
        quantize_mono_batch_kernel<<<dim3(gx, (unsigned)ns), 256, 0, ctx->stream>>>(in + s0 * in_stream_stride, in_stream_stride, n, bits, fmt == SK_FMT_FLOAT,
                                                                                    out + s0 * out_stream_bytes, out_stream_bytes, vec);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

extern "C" int sk_mixdown_mono_dev(sk_ctx* ctx, const float* planar, size_t frames, int channels, int mix_mode, float* out) {
    return sk_mixdown_strided_dev(ctx, planar, frames, frames, channels, mix_mode, out);
}

extern "C" int sk_mixdown_mono(sk_ctx* ctx, const float* planar, size_t frames, int channels, int mix_mode, float* out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t in_bytes = frames * channels * sizeof(float);
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, frames * sizeof(float)))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, planar, in_bytes))) return rc;
    if ((rc = sk_mixdown_mono_dev(ctx, (const float*)ctx->d_in, frames, channels, mix_mode, (float*)ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, frames * sizeof(float));
}

// ---- surround -> stereo downmix with block-wide peak normalisation (decoder lib.rs:3512-3557) -----
__global__ void downmix_stereo_kernel(const float* __restrict__ planar, size_t frames, int channels, float* __restrict__ out, unsigned* peak_bits) {
    const float k = 0.707f;
    float mx = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += (size_t)gridDim.x * blockDim.x) {
        float l = planar[i], r = planar[frames + i];
        float c = planar[2 * frames + i];
        l = __fadd_rn(l, __fmul_rn(k, c));
        r = __fadd_rn(r, __fmul_rn(k, c));
        if (channels > 4) {
            l = __fadd_rn(l, __fmul_rn(k, planar[4 * frames + i]));
            if (channels > 5) r = __fadd_rn(r, __fmul_rn(k, planar[5 * frames + i]));
        }
        out[i] = l; out[frames + i] = r;
        float al = fabsf(l), ar = fabsf(r);
        if (al > mx) mx = al;  // NaN never wins, like f32::max in the reference's fold
        if (ar > mx) mx = ar;
    }
    for (int o = 16; o > 0; o >>= 1) { float t = __shfl_xor_sync(0xffffffffu, mx, o); if (t > mx) mx = t; }
    if ((threadIdx.x & 31) == 0) atomicMax(peak_bits, __float_as_uint(mx));  // non-negative floats order like their bits
}
__global__ void downmix_stereo_scale_kernel(float* __restrict__ out, size_t n, const unsigned* peak_bits) {
    float mx = __uint_as_float(*peak_bits);
    if (!(mx > 1.0f)) return;
    float scale = __fdiv_rn(1.0f, mx);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __fmul_rn(out[i], scale);
}

extern "C" int sk_downmix_stereo(sk_ctx* ctx, const float* planar, size_t frames, int channels, float* out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 2 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "stereo downmix needs more than 2 input channels");
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t in_bytes = frames * channels * sizeof(float), out_bytes = 2 * frames * sizeof(float);
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, in_bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, out_bytes + 256))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, planar, in_bytes))) return rc;
    unsigned* peak = (unsigned*)((char*)ctx->d_out + sk_round_up(out_bytes, 16));
    SK_CUDA(ctx, cudaMemsetAsync(peak, 0, sizeof(unsigned), ctx->stream));
    downmix_stereo_kernel<<<grid_for(ctx, frames, 256), 256, 0, ctx->stream>>>((const float*)ctx->d_in, frames, channels, (float*)ctx->d_out, peak);
    SK_LAUNCH_CHECK(ctx);
    downmix_stereo_scale_kernel<<<grid_for(ctx, 2 * frames, 256), 256, 0, ctx->stream>>>((float*)ctx->d_out, 2 * frames, peak);
    SK_LAUNCH_CHECK(ctx);
    return sk_d2h(ctx, out, ctx->d_out, out_bytes);
}

// ---- i16 helpers: interleave_vecs_i16, stereo_to_mono_* -----------------------------------------------
__global__ void interleave_i16_kernel(const int16_t* __restrict__ planar, int channels, size_t frames, int16_t* __restrict__ out) {
    const size_t total = frames * (size_t)channels;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = planar[(i % channels) * frames + i / channels];
}
extern "C" int sk_interleave_i16(sk_ctx* ctx, const int16_t* planar, int channels, size_t frames, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = frames * channels * 2;
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, bytes))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, planar, bytes))) return rc;
    interleave_i16_kernel<<<grid_for(ctx, frames * channels, 256), 256, 0, ctx->stream>>>((const int16_t*)ctx->d_in, channels, frames, (int16_t*)ctx->d_out);
    SK_LAUNCH_CHECK(ctx);
    return sk_d2h(ctx, out, ctx->d_out, bytes);
}

// WavStreamEncoder::push_planar_{i16,i32,f32} (soundkit/src/wav.rs:444-494): planar [channels][frames] elements of 2 or 4 bytes ->
// interleaved little-endian frames.  A bit copy: 4-byte elements take the PCMFloat interleaver (no arithmetic touches the words, so
// i32 samples and NaN payloads pass unchanged), 2-byte elements the i16 interleaver.
extern "C" int sk_interleave_planar_dev(sk_ctx* ctx, int elem_bytes, const void* planar, int channels, size_t frames, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (elem_bytes != 2 && elem_bytes != 4) SK_FAIL(ctx, SK_E_ARG, "planar elements must be 2 or 4 bytes wide, got %d", elem_bytes);
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    if (elem_bytes == 4) return sk_planar_f32_to_pcm_dev(ctx, (const float*)planar, frames, channels, 32, SK_FMT_FLOAT, out);
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    interleave_i16_kernel<<<grid_for(ctx, frames * channels, 256), 256, 0, ctx->stream>>>((const int16_t*)planar, channels, frames, (int16_t*)out);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}
extern "C" int sk_interleave_planar(sk_ctx* ctx, int elem_bytes, const void* planar, int channels, size_t frames, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (elem_bytes != 2 && elem_bytes != 4) SK_FAIL(ctx, SK_E_ARG, "planar elements must be 2 or 4 bytes wide, got %d", elem_bytes);
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (!frames) return SK_OK;
    if (!planar || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = frames * (size_t)channels * (size_t)elem_bytes;
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, bytes))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, bytes))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, planar, bytes))) return rc;
    if ((rc = sk_interleave_planar_dev(ctx, elem_bytes, ctx->d_in, channels, frames, (uint8_t*)ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, bytes);
}

__global__ void stereo_to_mono_kernel(const uint32_t* __restrict__ in, size_t frames, int avg, int16_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t fr = in[i];
        int l = (int16_t)(fr & 0xffff), r = (int16_t)(fr >> 16);
        out[i] = (int16_t)(avg ? (l + r) / 2 : l);  // C `/` truncates toward zero like Rust
    }
}
extern "C" int sk_stereo_to_mono_i16(sk_ctx* ctx, int avg, const int16_t* in, size_t n, int16_t* out) {
    if (!ctx) return SK_E_ARG;
    if (n % 2) SK_FAIL(ctx, SK_E_ARG, "Stereo buffer must contain an even number of samples");
    size_t frames = n / 2;
    if (!frames) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, n * 2))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, frames * 2))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, n * 2))) return rc;
    stereo_to_mono_kernel<<<grid_for(ctx, frames, 256), 256, 0, ctx->stream>>>((const uint32_t*)ctx->d_in, frames, avg, (int16_t*)ctx->d_out);
    SK_LAUNCH_CHECK(ctx);
    return sk_d2h(ctx, out, ctx->d_out, frames * 2);
}

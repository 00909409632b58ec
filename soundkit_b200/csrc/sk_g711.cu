// sk_g711.cu — kernel 2: G.711 mu-law / A-law transcode with shared-memory lookup tables.
//
// Replaces soundkit-g711/src/lib.rs:22-66 (encode_i16 / decode_i16) and the trait wrappers
// G711Encoder::encode_{i16,i32} (:120-138), G711Decoder::decode_{i16,i32,f32} (:183-232).
// Bit-exact: the tables are generated on the device from the same integer arithmetic the
// reference uses per sample (:235-322); the hot loop is 1 lookup per sample.
//
// HBM-bound (3 B/sample for the i16 side, 5 B for i32/f32 outputs).  Layout choices:
//   decode: the 256-entry table is replicated once per shared-memory bank ([256][32] u32 = 32 KiB)
//           so 32 lanes with 32 arbitrary codes never conflict; entries already hold the output
//           word (i16 | i32<<16 | f32 bits) so the loop is extract -> LDS -> pack -> 128-bit store.
//   encode: the full 65 536-entry u8 table (64 KiB) sits in shared memory; one byte lookup per
//           sample, 16 B loads and 8/16 B stores.
// Both are persistent grid-stride kernels sized to a multiple of the SM count.
#include "sk_common.cuh"

// ---- per-sample reference arithmetic (device) ------------------------------------------------
__device__ __forceinline__ int g711_dec_mulaw(uint32_t b) {  // lib.rs:256-269
    uint32_t u = (~b) & 0xff;
    int m = (int)((u & 0x0f) << 3) + 0x84;
    m <<= (u & 0x70) >> 4;
    return (u & 0x80) ? 0x84 - m : m - 0x84;
}
__device__ __forceinline__ int g711_dec_alaw(uint32_t b) {  // lib.rs:296-315
    uint32_t a = (b ^ 0x55) & 0xff;
    int seg = (a & 0x70) >> 4;
    int m = (int)(a & 0x0f) << 4;
    if (seg == 0) m += 8;
    else if (seg == 1) m += 0x108;
    else m = (m + 0x108) << (seg - 1);
    return (a & 0x80) ? m : -m;
}
__device__ __forceinline__ int g711_segment_for(int v) {  // lib.rs:317-322 (first i with v <= ends[i])
    // ends[i] = (0x100 << i) - 1  ->  i = max(0, bit_length(v) - 8), 8 when v > 0x7fff
    int seg = v <= 0xff ? 0 : (32 - __clz(v)) - 8;
    return seg > 8 ? 8 : seg;
}
__device__ __forceinline__ uint32_t g711_enc_mulaw(int s) {  // lib.rs:235-254
    int pcm = s; uint32_t mask;
    if (pcm < 0) { pcm = 0x84 - pcm; mask = 0x7f; } else { pcm += 0x84; mask = 0xff; }
    pcm = min(pcm, 32635);
    int seg = g711_segment_for(pcm);
    uint32_t enc = seg >= 8 ? 0x7f : (uint32_t)((seg << 4) | ((pcm >> (seg + 3)) & 0x0f));
    return (enc ^ mask) & 0xff;
}
__device__ __forceinline__ uint32_t g711_enc_alaw(int s) {  // lib.rs:271-294
    int pcm = s; uint32_t mask;
    if (pcm >= 0) mask = 0xd5; else { pcm = -pcm - 1; mask = 0x55; }
    int seg = g711_segment_for(pcm);
    uint32_t enc;
    if (seg >= 8) enc = 0x7f;
    else enc = (uint32_t)(seg << 4) | (uint32_t)((pcm >> (seg < 2 ? 4 : seg + 3)) & 0x0f);
    return (enc ^ mask) & 0xff;
}
// output word for one decoded sample: i16 (low half) | i32 = i16 << 16 (:206-208) | f32 = i16 / 32768 (:227-229)
__device__ __forceinline__ uint32_t g711_dec_word(int law, int kind, uint32_t code) {
    int v = law == SK_LAW_MULAW ? g711_dec_mulaw(code) : g711_dec_alaw(code);
    if (kind == SK_KIND_I16) return (uint32_t)v & 0xffffu;
    if (kind == SK_KIND_I32) return (uint32_t)v << 16;
    return __float_as_uint(__fmul_rn((float)v, 1.0f / 32768.0f));  // exact: power-of-two scale
}

// ---- decode ------------------------------------------------------------------------------------
constexpr int kDecThreads = 512;

template <int KIND>
__global__ void __launch_bounds__(kDecThreads) g711_decode_kernel(const uint8_t* __restrict__ in, void* __restrict__ outv,
                                                                   size_t n, int law) {
    __shared__ uint32_t lut[256 * 32];  // [code][bank]
    for (int i = threadIdx.x; i < 256 * 32; i += kDecThreads) lut[i] = g711_dec_word(law, KIND, (uint32_t)i >> 5);
    __syncthreads();
    const uint32_t* my = lut + (threadIdx.x & 31);
    const size_t nvec = n / 16;  // 16 codes per 128-bit load
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    const size_t stride = (size_t)gridDim.x * kDecThreads;
    for (size_t v = (size_t)blockIdx.x * kDecThreads + threadIdx.x; v < nvec; v += stride) {
        uint4 c = ld_stream_v4(in4 + v);
        uint32_t w[4] = {c.x, c.y, c.z, c.w};
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int k = 0; k < 4; k++) o[q * 4 + k] = my[((w[q] >> (8 * k)) & 0xff) << 5];
        if (KIND == SK_KIND_I16) {
            uint4* out4 = reinterpret_cast<uint4*>(outv) + v * 2;
            uint4 a, b;
            a.x = o[0] | (o[1] << 16);   a.y = o[2] | (o[3] << 16);   a.z = o[4] | (o[5] << 16);   a.w = o[6] | (o[7] << 16);
            b.x = o[8] | (o[9] << 16);   b.y = o[10] | (o[11] << 16); b.z = o[12] | (o[13] << 16); b.w = o[14] | (o[15] << 16);
            st_stream_v4(out4, a);
            st_stream_v4(out4 + 1, b);
        } else {
            uint4* out4 = reinterpret_cast<uint4*>(outv) + v * 4;
#pragma unroll
            for (int q = 0; q < 4; q++) st_stream_v4(out4 + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
        }
    }
    // tail (< 16 codes): one thread per sample
    if (blockIdx.x == 0) {
        for (size_t i = nvec * 16 + threadIdx.x; i < n; i += kDecThreads) {
            uint32_t wv = my[(uint32_t)in[i] << 5];
            if (KIND == SK_KIND_I16) reinterpret_cast<uint16_t*>(outv)[i] = (uint16_t)wv;
            else reinterpret_cast<uint32_t*>(outv)[i] = wv;
        }
    }
}

// unaligned caller buffers: correctness-only scalar path
__global__ void g711_decode_scalar_kernel(const uint8_t* __restrict__ in, void* __restrict__ outv, size_t n, int law, int kind) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t wv = g711_dec_word(law, kind, in[i]);
        if (kind == SK_KIND_I16) reinterpret_cast<uint16_t*>(outv)[i] = (uint16_t)wv;
        else reinterpret_cast<uint32_t*>(outv)[i] = wv;
    }
}

// ---- encode ------------------------------------------------------------------------------------
constexpr int kEncThreads = 512;

__global__ void g711_build_enc_table_kernel(uint8_t* table, int law) {  // 65536 entries, index = (u16)sample
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536) {
        int s = (int)(int16_t)(uint16_t)i;
        table[i] = (uint8_t)(law == SK_LAW_MULAW ? g711_enc_mulaw(s) : g711_enc_alaw(s));
    }
}

template <int KIND>  // SK_KIND_I16: 8 samples / 16 B; SK_KIND_I32: 4 samples / 16 B, (s >> 16) as i16 (:133-135)
__global__ void __launch_bounds__(kEncThreads) g711_encode_kernel(const void* __restrict__ inv, uint8_t* __restrict__ out, size_t n,
                                                                   const uint8_t* __restrict__ table) {
    extern __shared__ __align__(16) uint8_t lut[];  // 65536 bytes
    {
        const uint4* t4 = reinterpret_cast<const uint4*>(table);
        uint4* l4 = reinterpret_cast<uint4*>(lut);
        for (int i = threadIdx.x; i < 4096; i += kEncThreads) l4[i] = t4[i];
    }
    __syncthreads();
    constexpr int SPV = KIND == SK_KIND_I16 ? 8 : 4;  // samples per 128-bit load
    const size_t nvec = n / SPV;
    const uint4* in4 = reinterpret_cast<const uint4*>(inv);
    const size_t stride = (size_t)gridDim.x * kEncThreads;
    for (size_t v = (size_t)blockIdx.x * kEncThreads + threadIdx.x; v < nvec; v += stride) {
        uint4 c = ld_stream_v4(in4 + v);
        uint32_t w[4] = {c.x, c.y, c.z, c.w};
        if (KIND == SK_KIND_I16) {
            uint32_t o[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t a = lut[w[2 * h] & 0xffff], b = lut[w[2 * h] >> 16];
                uint32_t c2 = lut[w[2 * h + 1] & 0xffff], d = lut[w[2 * h + 1] >> 16];
                o[h] = a | (b << 8) | (c2 << 16) | (d << 24);
            }
            st_stream_v2(reinterpret_cast<uint2*>(out) + v, make_uint2(o[0], o[1]));
        } else {
            uint32_t o = lut[w[0] >> 16] | ((uint32_t)lut[w[1] >> 16] << 8) | ((uint32_t)lut[w[2] >> 16] << 16) |
                         ((uint32_t)lut[w[3] >> 16] << 24);
            reinterpret_cast<uint32_t*>(out)[v] = o;
        }
    }
    if (blockIdx.x == 0) {
        for (size_t i = nvec * SPV + threadIdx.x; i < n; i += kEncThreads) {
            uint32_t idx = KIND == SK_KIND_I16 ? (uint32_t)reinterpret_cast<const uint16_t*>(inv)[i]
                                               : (reinterpret_cast<const uint32_t*>(inv)[i] >> 16);
            out[i] = lut[idx];
        }
    }
}

__global__ void g711_encode_scalar_kernel(const void* __restrict__ inv, uint8_t* __restrict__ out, size_t n, int law, int kind) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int s = kind == SK_KIND_I16 ? (int)reinterpret_cast<const int16_t*>(inv)[i]
                                    : (int)(int16_t)(reinterpret_cast<const int32_t*>(inv)[i] >> 16);
        out[i] = (uint8_t)(law == SK_LAW_MULAW ? g711_enc_mulaw(s) : g711_enc_alaw(s));
    }
}

// ---- host side -----------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int grid_for(sk_ctx* ctx, size_t nvec, int threads, int per_sm) {
    size_t need = (nvec + threads - 1) / threads;
    size_t cap = (size_t)ctx->num_sms * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" int sk_g711_decode_dev(sk_ctx* ctx, int law, const uint8_t* in, size_t n, int out_kind, void* out) {
    if (!ctx) return SK_E_ARG;
    if (law != SK_LAW_MULAW && law != SK_LAW_ALAW) SK_FAIL(ctx, SK_E_ARG, "unknown G.711 law %d", law);
    if (out_kind < SK_KIND_I16 || out_kind > SK_KIND_F32) SK_FAIL(ctx, SK_E_ARG, "unknown output kind %d", out_kind);
    if (n == 0) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!aligned16(in) || !aligned16(out)) {
        g711_decode_scalar_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(in, out, n, law, out_kind);
        SK_LAUNCH_CHECK(ctx);
        return SK_OK;
    }
    int grid = grid_for(ctx, n / 16, kDecThreads, 4);
    if (out_kind == SK_KIND_I16) g711_decode_kernel<SK_KIND_I16><<<grid, kDecThreads, 0, ctx->stream>>>(in, out, n, law);
    else if (out_kind == SK_KIND_I32) g711_decode_kernel<SK_KIND_I32><<<grid, kDecThreads, 0, ctx->stream>>>(in, out, n, law);
    else g711_decode_kernel<SK_KIND_F32><<<grid, kDecThreads, 0, ctx->stream>>>(in, out, n, law);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

static int ensure_enc_table(sk_ctx* ctx, int law) {
    if (ctx->d_g711_enc[law]) return SK_OK;
    SK_CUDA(ctx, cudaMalloc((void**)&ctx->d_g711_enc[law], 65536));
    g711_build_enc_table_kernel<<<256, 256, 0, ctx->stream>>>(ctx->d_g711_enc[law], law);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

extern "C" int sk_g711_encode_dev(sk_ctx* ctx, int law, int in_kind, const void* in, size_t n, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (law != SK_LAW_MULAW && law != SK_LAW_ALAW) SK_FAIL(ctx, SK_E_ARG, "unknown G.711 law %d", law);
    if (in_kind != SK_KIND_I16 && in_kind != SK_KIND_I32) SK_FAIL(ctx, SK_E_ARG, "G.711 encode takes i16 or i32 input");
    if (n == 0) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!aligned16(in) || !aligned16(out)) {
        g711_encode_scalar_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(in, out, n, law, in_kind);
        SK_LAUNCH_CHECK(ctx);
        return SK_OK;
    }
    int rc = ensure_enc_table(ctx, law);
    if (rc) return rc;
    // the attribute is per device; setting it is cheap enough to do on every call
    SK_CUDA(ctx, cudaFuncSetAttribute(g711_encode_kernel<SK_KIND_I16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    SK_CUDA(ctx, cudaFuncSetAttribute(g711_encode_kernel<SK_KIND_I32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    if (in_kind == SK_KIND_I16) {
        int grid = grid_for(ctx, n / 8, kEncThreads, 3);
        g711_encode_kernel<SK_KIND_I16><<<grid, kEncThreads, 65536, ctx->stream>>>(in, out, n, ctx->d_g711_enc[law]);
    } else {
        int grid = grid_for(ctx, n / 4, kEncThreads, 3);
        g711_encode_kernel<SK_KIND_I32><<<grid, kEncThreads, 65536, ctx->stream>>>(in, out, n, ctx->d_g711_enc[law]);
    }
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

static size_t kind_bytes(int kind) { return kind == SK_KIND_I16 ? 2 : 4; }

extern "C" int sk_g711_decode(sk_ctx* ctx, int law, const uint8_t* in, size_t n, int out_kind, void* out) {
    if (!ctx) return SK_E_ARG;
    if (out_kind < SK_KIND_I16 || out_kind > SK_KIND_F32) SK_FAIL(ctx, SK_E_ARG, "unknown output kind %d", out_kind);
    if (n == 0) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, n))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, n * kind_bytes(out_kind)))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, n))) return rc;
    if ((rc = sk_g711_decode_dev(ctx, law, (const uint8_t*)ctx->d_in, n, out_kind, ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, n * kind_bytes(out_kind));
}

extern "C" int sk_g711_encode(sk_ctx* ctx, int law, int in_kind, const void* in, size_t n, uint8_t* out) {
    if (!ctx) return SK_E_ARG;
    if (in_kind != SK_KIND_I16 && in_kind != SK_KIND_I32) SK_FAIL(ctx, SK_E_ARG, "G.711 encode takes i16 or i32 input");
    if (n == 0) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = sk_ensure_dev_in(ctx, n * kind_bytes(in_kind)))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, n))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, n * kind_bytes(in_kind)))) return rc;
    if ((rc = sk_g711_encode_dev(ctx, law, in_kind, ctx->d_in, n, (uint8_t*)ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, n);
}

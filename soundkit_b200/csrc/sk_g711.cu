// sk_g711.cu — kernel 2: G.711 mu-law / A-law transcode.
//
// Replaces soundkit-g711/src/lib.rs:22-66 (encode_i16 / decode_i16) and the trait wrappers
// G711Encoder::encode_{i16,i32} (:120-138), G711Decoder::decode_{i16,i32,f32} (:183-232).
// Bit-exact: the decode tables are generated on the device from the same integer arithmetic the
// reference uses per sample (:235-322); the encoder evaluates that arithmetic in exact float operations.
//
// HBM-bound (3 B/sample for the i16 side, 5 B for i32/f32 outputs).  Layout choices:
//   decode: a shared-memory table lookup per sample.  16-bit outputs keep the 256-entry table replicated once per
//           shared-memory bank ([256][32] u32 = 32 KiB, 32 lanes with 32 arbitrary codes never conflict); 4-byte outputs
//           use the plain 256-entry table.  The loop is extract -> LDS -> pack -> 128-bit store.
//   encode: TABLE-FREE.  The segment is the exponent of the biased magnitude and the four mantissa bits it keeps are the
//           top fraction bits, so (seg << 4 | mantissa) falls out of the float's bit pattern after one exact power-of-two
//           scale (see g711_enc_fast).  A 64 KiB shared-memory table of the same function measured 0.51 of the HBM rate.
// Both directions launch a LARGE grid of short-lived CTAs (grid_for in sk_pcm.cu): persistent grid-stride forms of the
// same loops topped out at 0.85-0.89 of the copy rate, this form runs at 1.02-1.08.
#include <algorithm>

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
// Every global access is a full-width, contiguous warp access: per step a warp loads 512 codes as one 512-byte run
// (16 B per lane) and stores the 1024 / 2048 output bytes as 2 / 4 runs of 512 bytes (16 B per lane).  The lane that
// loaded codes [16t, 16t+16) is not the lane that stores their samples, so the codes cross lanes through 512 bytes of
// warp-private shared memory (one STS.128, then one LDS.64 / LDS.32 per store).
// The grid is LARGE and short-lived (one CTA = 8 warps x 2 steps = 8 Ki codes): a persistent grid whose CTAs stride through
// the buffers in lockstep reached 0.88 of the HBM copy rate, the same work as a big grid 1.0+ (see sk_pcm.cu: grid_for).
// Table: 4-byte outputs (i32 / f32) use the plain 256-entry table (1 KiB); random codes conflict ~3.5-way in it, which at
// full HBM rate is 0.4 of a shared-memory wavefront per clock.  16-bit outputs do twice the lookups per byte moved, so they
// keep the bank-replicated [256][32] table (32 lanes with arbitrary codes never conflict): each thread computes one entry
// and writes its 32 copies with a rotated column order (conflict-free), and a CTA owns 4x more steps to amortise that.
constexpr int kDecThreads = 256;
constexpr int kDecUnroll = 2;  // warp steps whose loads are issued together
template <int KIND> __host__ __device__ constexpr int dec_steps_per_warp() { return KIND == SK_KIND_I16 ? 8 : 2; }

template <int KIND>
__global__ void __launch_bounds__(kDecThreads) g711_decode_kernel(const uint8_t* __restrict__ in, void* __restrict__ outv,
                                                                   size_t n, int law) {
    constexpr bool REPL = KIND == SK_KIND_I16;
    constexpr int LS = REPL ? 5 : 0;  // log2 of the table's row stride
    __shared__ uint32_t lut[REPL ? 256 * 32 : 256];
    __shared__ __align__(16) uint8_t xch[kDecThreads / 32][kDecUnroll][512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const uint32_t e = g711_dec_word(law, KIND, (uint32_t)threadIdx.x);
        if (REPL) {
#pragma unroll
            for (int j = 0; j < 32; j++) lut[(threadIdx.x << 5) + ((j + lane) & 31)] = e;
        } else lut[threadIdx.x] = e;
    }
    __syncthreads();
    const uint32_t* my = lut + (REPL ? lane : 0);
    constexpr int OB = KIND == SK_KIND_I16 ? 2 : 4;
    constexpr int EPL = 16 / OB;        // elements per lane per store (8 or 4)
    constexpr int NST = 512 / (32 * EPL);  // stores per step (2 or 4)
    const size_t nsteps = n / 512;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    uint4* out4 = reinterpret_cast<uint4*>(outv);
    // a warp owns dec_steps_per_warp consecutive steps of its CTA's chunk, kDecUnroll at a time
    constexpr int SPW = dec_steps_per_warp<KIND>();
    const size_t cta_steps = (size_t)(kDecThreads / 32) * SPW;
    for (size_t cbase = (size_t)blockIdx.x * cta_steps; cbase < nsteps; cbase += (size_t)gridDim.x * cta_steps)
    for (size_t s0 = cbase + (size_t)warp * SPW, send = s0 + SPW; s0 < send && s0 < nsteps; s0 += kDecUnroll) {
        uint4 c[kDecUnroll];
#pragma unroll
        for (int u = 0; u < kDecUnroll; u++)
            if (s0 + u < nsteps) c[u] = ld_stream_v4_nv(in4 + (s0 + u) * 32 + lane);
        __syncwarp();  // the previous iteration's reads of xch are done
#pragma unroll
        for (int u = 0; u < kDecUnroll; u++)
            if (s0 + u < nsteps) reinterpret_cast<uint4*>(xch[warp][u])[lane] = c[u];
        __syncwarp();
#pragma unroll
        for (int u = 0; u < kDecUnroll; u++) {
            if (s0 + u >= nsteps) break;
#pragma unroll
            for (int i = 0; i < NST; i++) {
                uint32_t w[2];
                if (EPL == 8) { const uint2 t = reinterpret_cast<const uint2*>(xch[warp][u])[i * 32 + lane]; w[0] = t.x; w[1] = t.y; }
                else { w[0] = reinterpret_cast<const uint32_t*>(xch[warp][u])[i * 32 + lane]; w[1] = 0; }
                uint32_t o[EPL];
#pragma unroll
                for (int k = 0; k < EPL; k++) o[k] = my[((w[k >> 2] >> (8 * (k & 3))) & 0xff) << LS];
                uint4 r;
                if (KIND == SK_KIND_I16) {
                    r.x = __byte_perm(o[0], o[1], 0x5410); r.y = __byte_perm(o[2], o[3], 0x5410);
                    r.z = __byte_perm(o[4 % EPL], o[5 % EPL], 0x5410); r.w = __byte_perm(o[6 % EPL], o[7 % EPL], 0x5410);
                } else {
                    r = make_uint4(o[0], o[1], o[2], o[3]);
                }
                st_stream_v4(out4 + ((s0 + u) * NST + i) * 32 + lane, r);
            }
        }
    }
    // tail (< 512 codes): one thread per sample
    if (blockIdx.x == 0) {
        for (size_t i = nsteps * 512 + threadIdx.x; i < n; i += kDecThreads) {
            uint32_t wv = my[(uint32_t)in[i] << LS];
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
// Table-free: the segment search of lib.rs:317-322 is the exponent of the (biased) magnitude, and the 4 mantissa
// bits it keeps are the top 4 fraction bits, so one int->float conversion yields `(seg << 4) | mantissa` as bits
// [19, 27) of the float once the value is scaled so that segment 0 has the right exponent:
//   mu-law (lib.rs:235-254): p = min(|s| + 0x84, 32635) in [132, 32635]; p * 2^-133 has exponent field seg + 1
//           -> enc = (bits >> 19) - 16, out = enc ^ (s < 0 ? 0x7f : 0xff) = (0x10f - (bits >> 19)) ^ (neg << 7)
//   A-law  (lib.rs:271-294): p = s >= 0 ? s : -s - 1 = |s + 0.5| - 0.5; segments 0 and 1 share one linear step, which is
//           exactly an IEEE subnormal: p * 2^-134 has exponent field seg (0 for p < 256) -> enc = bits >> 19,
//           out = enc ^ (s >= 0 ? 0xd5 : 0x55)
// Every float operation is exact (integers < 2^16 scaled by powers of two), so the result is the reference's
// integer arithmetic bit for bit (checked over all 65 536 inputs in tests).  A 64 KiB shared-memory table of the
// same function ran at 0.51 of HBM (random 16-bit indices: ~3.5-way bank conflicts); this form is 7 issue slots
// per sample and streams.
constexpr int kEncThreads = 256;

template <int LAW>
__device__ __forceinline__ uint32_t g711_enc_fast(float f) {  // f = (float)sample; returns enc + 16 (mu) / enc (A), in [0, 0x8f]
    if (LAW == SK_LAW_MULAW) {
        float p = fminf(__fmaf_rn(fabsf(f), 0x1p-133f, 132.0f * 0x1p-133f), 32635.0f * 0x1p-133f);
        return __float_as_uint(p) >> 19;
    } else {
        float g = __fadd_rn(f, 0.5f);
        float p = __fmaf_rn(fabsf(g), 0x1p-134f, -0x1p-135f);
        return __float_as_uint(p) >> 19;
    }
}

// 4 samples held as two packed i16 pairs (KIND i16) or four i32 words whose high halves are the samples (KIND i32)
template <int LAW, int KIND>
__device__ __forceinline__ uint32_t g711_enc4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    float f0, f1, f2, f3;
    uint32_t signs;  // byte k = high byte of sample k (bit 7 = sign)
    if (KIND == SK_KIND_I16) {
        f0 = (float)(int16_t)(a & 0xffff); f1 = (float)(int16_t)(a >> 16);
        f2 = (float)(int16_t)(b & 0xffff); f3 = (float)(int16_t)(b >> 16);
        signs = __byte_perm(a, b, 0x7531);
    } else {
        f0 = (float)(int16_t)(a >> 16); f1 = (float)(int16_t)(b >> 16);
        f2 = (float)(int16_t)(c >> 16); f3 = (float)(int16_t)(d >> 16);
        signs = __byte_perm(__byte_perm(a, b, 0x7773), __byte_perm(c, d, 0x7773), 0x5410);
    }
    uint32_t e0 = g711_enc_fast<LAW>(f0), e1 = g711_enc_fast<LAW>(f1), e2 = g711_enc_fast<LAW>(f2), e3 = g711_enc_fast<LAW>(f3);
    uint32_t packed = __byte_perm(__byte_perm(e0, e1, 0x0040), __byte_perm(e2, e3, 0x0040), 0x5410);
    // mu: bytes hold enc + 16 >= 16, so one packed subtract removes the bias without borrows; then ^ 0xff ^ (neg << 7)
    // A : ^ 0x55 ^ (!neg << 7) = ^ 0xd5 ^ (neg << 7)
    if (LAW == SK_LAW_MULAW) packed = ~(packed - 0x10101010u);
    else packed ^= 0xd5d5d5d5u;
    return packed ^ (signs & 0x80808080u);
}

// Same access discipline as the decoder: per step a warp loads 512 samples as 2 (i16) or 4 (i32) contiguous 512-byte
// runs and stores their 512 codes as one 512-byte run (16 B per lane); the codes cross lanes through warp-private
// shared memory.  KIND i32: (s >> 16) as i16 (:133-135).
template <int LAW, int KIND>
__global__ void __launch_bounds__(kEncThreads) g711_encode_kernel(const void* __restrict__ inv, uint8_t* __restrict__ out, size_t n) {
    __shared__ __align__(16) uint8_t xch[kEncThreads / 32][512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NLD = KIND == SK_KIND_I16 ? 2 : 4;  // 128-bit loads per lane per step
    const size_t nsteps = n / 512;
    const uint4* in4 = reinterpret_cast<const uint4*>(inv);
    uint4* out4 = reinterpret_cast<uint4*>(out);
    const size_t wstride = (size_t)gridDim.x * (kEncThreads / 32);
    for (size_t s0 = (size_t)blockIdx.x * (kEncThreads / 32) + warp; s0 < nsteps; s0 += wstride) {
        uint4 c[NLD];
#pragma unroll
        for (int i = 0; i < NLD; i++) c[i] = ld_stream_v4_nv(in4 + (s0 * NLD + i) * 32 + lane);
        __syncwarp();  // the previous step's read of xch is done
#pragma unroll
        for (int i = 0; i < NLD; i++) {
            if (KIND == SK_KIND_I16) {
                const uint32_t lo = g711_enc4<LAW, KIND>(c[i].x, c[i].y, 0, 0), hi = g711_enc4<LAW, KIND>(c[i].z, c[i].w, 0, 0);
                reinterpret_cast<uint2*>(xch[warp])[i * 32 + lane] = make_uint2(lo, hi);
            } else {
                reinterpret_cast<uint32_t*>(xch[warp])[i * 32 + lane] = g711_enc4<LAW, KIND>(c[i].x, c[i].y, c[i].z, c[i].w);
            }
        }
        __syncwarp();
        st_stream_v4(out4 + s0 * 32 + lane, reinterpret_cast<const uint4*>(xch[warp])[lane]);
    }
    if (blockIdx.x == 0) {
        for (size_t i = nsteps * 512 + threadIdx.x; i < n; i += kEncThreads) {
            int s = KIND == SK_KIND_I16 ? (int)reinterpret_cast<const int16_t*>(inv)[i]
                                        : (int)(int16_t)(reinterpret_cast<const uint32_t*>(inv)[i] >> 16);
            out[i] = (uint8_t)(LAW == SK_LAW_MULAW ? g711_enc_mulaw(s) : g711_enc_alaw(s));
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

// as many short-lived CTAs as the work takes (see the decode kernel's note); the kernels keep a stride loop beyond 2^22 CTAs
static int grid_for(sk_ctx* ctx, size_t work_threads, int threads) {
    (void)ctx;
    size_t need = (work_threads + threads - 1) / threads;
    const size_t cap = (size_t)1 << 22;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

constexpr size_t kG711SmallN = 4096;  // below this a launch is latency, not bandwidth: one scalar CTA per 256 samples

extern "C" int sk_g711_decode_dev(sk_ctx* ctx, int law, const uint8_t* in, size_t n, int out_kind, void* out) {
    if (!ctx) return SK_E_ARG;
    if (law != SK_LAW_MULAW && law != SK_LAW_ALAW) SK_FAIL(ctx, SK_E_ARG, "unknown G.711 law %d", law);
    if (out_kind < SK_KIND_I16 || out_kind > SK_KIND_F32) SK_FAIL(ctx, SK_E_ARG, "unknown output kind %d", out_kind);
    if (n == 0) return SK_OK;
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!aligned16(in) || !aligned16(out) || n < kG711SmallN) {  // small calls: per-sample arithmetic, no table to build
        const unsigned sgrid = (unsigned)std::min<size_t>((size_t)ctx->num_sms * 8, (n + 255) / 256);
        g711_decode_scalar_kernel<<<sgrid, 256, 0, ctx->stream>>>(in, out, n, law, out_kind);
        SK_LAUNCH_CHECK(ctx);
        return SK_OK;
    }
    const int spw = out_kind == SK_KIND_I16 ? dec_steps_per_warp<SK_KIND_I16>() : dec_steps_per_warp<SK_KIND_F32>();
    int grid = grid_for(ctx, (n / 512 + spw - 1) / spw * 32, kDecThreads);
    if (out_kind == SK_KIND_I16) g711_decode_kernel<SK_KIND_I16><<<grid, kDecThreads, 0, ctx->stream>>>(in, out, n, law);
    else if (out_kind == SK_KIND_I32) g711_decode_kernel<SK_KIND_I32><<<grid, kDecThreads, 0, ctx->stream>>>(in, out, n, law);
    else g711_decode_kernel<SK_KIND_F32><<<grid, kDecThreads, 0, ctx->stream>>>(in, out, n, law);
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
    if (!aligned16(in) || !aligned16(out) || n < kG711SmallN) {
        const unsigned sgrid = (unsigned)std::min<size_t>((size_t)ctx->num_sms * 8, (n + 255) / 256);
        g711_encode_scalar_kernel<<<sgrid, 256, 0, ctx->stream>>>(in, out, n, law, in_kind);
        SK_LAUNCH_CHECK(ctx);
        return SK_OK;
    }
    int grid = grid_for(ctx, n / 512 * 32, kEncThreads);
    if (law == SK_LAW_MULAW) {
        if (in_kind == SK_KIND_I16) g711_encode_kernel<SK_LAW_MULAW, SK_KIND_I16><<<grid, kEncThreads, 0, ctx->stream>>>(in, out, n);
        else g711_encode_kernel<SK_LAW_MULAW, SK_KIND_I32><<<grid, kEncThreads, 0, ctx->stream>>>(in, out, n);
    } else {
        if (in_kind == SK_KIND_I16) g711_encode_kernel<SK_LAW_ALAW, SK_KIND_I16><<<grid, kEncThreads, 0, ctx->stream>>>(in, out, n);
        else g711_encode_kernel<SK_LAW_ALAW, SK_KIND_I32><<<grid, kEncThreads, 0, ctx->stream>>>(in, out, n);
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
    void *s_in, *s_out;
    if (sk_small_begin(ctx, in, n, n * kind_bytes(out_kind), &s_in, &s_out)) {  // a telephony frame: zero copy, one launch, one synchronize
        if ((rc = sk_g711_decode_dev(ctx, law, (const uint8_t*)s_in, n, out_kind, s_out))) return rc;
        return sk_small_end(ctx, out, s_out, n * kind_bytes(out_kind));
    }
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
    void *s_in, *s_out;
    if (sk_small_begin(ctx, in, n * kind_bytes(in_kind), n, &s_in, &s_out)) {
        if ((rc = sk_g711_encode_dev(ctx, law, in_kind, s_in, n, (uint8_t*)s_out))) return rc;
        return sk_small_end(ctx, out, s_out, n);
    }
    if ((rc = sk_ensure_dev_in(ctx, n * kind_bytes(in_kind)))) return rc;
    if ((rc = sk_ensure_dev_out(ctx, n))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, n * kind_bytes(in_kind)))) return rc;
    if ((rc = sk_g711_encode_dev(ctx, law, in_kind, ctx->d_in, n, (uint8_t*)ctx->d_out))) return rc;
    return sk_d2h(ctx, out, ctx->d_out, n);
}

// sk_resample.cu — kernels 3/4: batched polyphase windowed-sinc FIR on FP32 CUDA cores, optionally
// fused with the PCM->f32 conversion in front of it and the mono mixdown behind it.
//
// Replaces rubato 0.14.1 `SincFixedIn<f32>::process_into_buffer` as soundkit reaches it from
// soundkit/src/audio_pipeline.rs:438-493 (downsample_audio, one chunk = whole buffer) and
// soundkit-decoder/src/lib.rs:1917-2060, 3324-3456 (StreamingResampler, chunk 4096, + downmix).
//
// Numerics contract (what makes the output identical to the AVX+FMA rubato path):
//   * the [256][256] table is generated on the host with the same f32 expression tree as rubato's
//     make_sincs / blackman_harris (libm sinf/cosf);
//   * the f64 index walk `idx += t_ratio` is data independent, so it is executed ONCE on the host per
//     (ratio, length, semantics) and shared by every stream of a batch (the "plan");
//   * each dot product keeps rubato's 8 lane accumulators (lane l owns taps l, l+8, ...; fused
//     multiply-add in tap order) and its reduction tree ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7));
//   * out = p0 + frac * (p1 - p0) with separate mul / sub / add (file is built with -fmad=false).
//
// Two kernels:
//   resample_int_kernel<STEP,...>  integer t_ratio (48k->16k: STEP 3).  idx is always an integer, so
//       subindex 0 and frac 0: out == row-0 dot product exactly, one table row, which is passed BY
//       VALUE as a kernel parameter -> every FFMA takes its coefficient straight from the constant
//       bank (no register, no shared-memory traffic for taps).  Each thread owns R consecutive
//       outputs of one channel (8*R accumulators) and walks the input window once: every input
//       sample fetched from shared memory with LDS.128 feeds R FFMAs.
//   resample_gen_kernel            any ratio, both interpolation products; outputs of one tile are
//       visited phase-major (lane = period, warp = phase) so the two table rows are warp-uniform
//       broadcast loads and the input reads are conflict-free.
#include <math.h>

#include <algorithm>

#include "sk_common.cuh"

#define SINC_LEN 256
#define OVERSAMPLING 256
#define STREAM_CHUNK 4096  // RESAMPLE_CHUNK_SIZE, soundkit-decoder/src/lib.rs:79

// ================================================================================================
// host: table and index plan
// ================================================================================================

// rubato sinc.rs make_sincs::<f32> + windows.rs blackman_harris (squared: BlackmanHarris2), f32 throughout
static void build_sinc_table(double ratio, std::vector<float>& sincs) {
    const float f_cutoff = ratio >= 1.0 ? 0.95f : 0.95f * (float)ratio;  // make_interpolator
    const size_t tot = (size_t)SINC_LEN * OVERSAMPLING;
    std::vector<float> y(tot);
    const float pi = 3.14159265358979323846264338327950288f;
    const float pi2 = 2.0f * pi, pi4 = 4.0f * pi, pi6 = 6.0f * pi;
    const float npf = (float)tot;
    float sum = 0.0f;
    for (size_t x = 0; x < tot; x++) {
        const float xf = (float)x;
        float w = 0.35875f - 0.48829f * cosf(pi2 * xf / npf) + 0.14128f * cosf(pi4 * xf / npf) - 0.01168f * cosf(pi6 * xf / npf);
        w = w * w;
        const float v = (xf - (float)(tot / 2)) * f_cutoff / (float)OVERSAMPLING;
        const float s = v == 0.0f ? 1.0f : sinf(v * pi) / (v * pi);
        const float val = w * s;
        sum += val;
        y[x] = val;
    }
    sum /= (float)OVERSAMPLING;
    sincs.resize(tot);
    for (size_t p = 0; p < SINC_LEN; p++)
        for (size_t n = 0; n < OVERSAMPLING; n++) sincs[(OVERSAMPLING - n - 1) * SINC_LEN + p] = y[OVERSAMPLING * p + n] / sum;
}

static uint64_t rate_key(uint32_t in_rate, uint32_t out_rate) { return ((uint64_t)in_rate << 32) | out_rate; }

static int get_sinc_table(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, const float** d_table, const std::vector<float>** h_table) {
    const uint64_t key = rate_key(in_rate, out_rate);
    auto it = ctx->sinc_tables.find(key);
    if (it == ctx->sinc_tables.end()) {
        std::vector<float>& h = ctx->sinc_host[key];
        build_sinc_table((double)out_rate / (double)in_rate, h);
        float* d = nullptr;
        SK_CUDA(ctx, cudaMalloc((void**)&d, h.size() * sizeof(float)));
        SK_CUDA(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->sinc_tables.emplace(key, d).first;
    }
    if (d_table) *d_table = it->second;
    if (h_table) *h_table = &ctx->sinc_host[key];
    return SK_OK;
}

// one rubato chunk: walks idx from last_index, appends (chunk_start + floor(idx), subindex, frac)
struct WalkOut {
    std::vector<int32_t> i0; std::vector<uint8_t> s0; std::vector<float> frac;
    size_t count = 0; bool record = true;
};
static double walk_chunk(double last_index, double ratio, size_t chunk, long long chunk_start, WalkOut& w) {
    const double t_ratio = 1.0 / ratio;
    const long long end_idx = (long long)chunk - (SINC_LEN + 1) - (long long)ceil(t_ratio);
    double idx = last_index;
    while (idx < (double)end_idx) {
        idx += t_ratio;  // sequential f64 accumulation, reproduced bit for bit
        if (w.record) {
            const double fl = floor(idx);
            w.i0.push_back((int32_t)(chunk_start + (long long)fl));
            w.s0.push_back((uint8_t)(long long)floor((idx - fl) * (double)OVERSAMPLING));  // get_nearest_times_2
            const double f = idx * (double)OVERSAMPLING;
            w.frac.push_back((float)(f - floor(f)));
        }
        w.count++;
    }
    return idx - (double)chunk;
}

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
};

static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

static void free_plan(sk_plan* p) {
    if (!p) return;
    if (p->d_i0) cudaFree(p->d_i0);
    if (p->d_s0) cudaFree(p->d_s0);
    if (p->d_frac) cudaFree(p->d_frac);
    delete p;
}

void sk_resample_free_pipe(sk_ctx* ctx);
void sk_resample_free_ctx(sk_ctx* ctx) {
    sk_resample_free_pipe(ctx);
    for (auto& kv : ctx->sinc_tables) cudaFree(kv.second);
    ctx->sinc_tables.clear();
    for (auto& kv : ctx->plans) free_plan(kv.second);
    ctx->plans.clear();
}

static const int kGenMaxWindow = 14400;  // frames of shared memory window the general kernel may use (57.6 KB)
static const int kGen2MaxWindow = 40000;  // resample_gen2_kernel: up to 160 KB of window + 64 KB of staged table rows (512-thread form)

// finish a general plan from recorded walk arrays: choose the tile period and upload
static int finalize_general_plan(sk_ctx* ctx, sk_plan* p, uint32_t in_rate, uint32_t out_rate, const WalkOut& w) {
    p->n_out = w.i0.size();
    const uint32_t g = gcd_u32(in_rate, out_rate);
    const uint32_t L = out_rate / g;
    auto window_for = [&](int period) {
        int mw = 0;
        const size_t tile = (size_t)period * 32;
        for (size_t n0 = 0; n0 < p->n_out; n0 += tile) {
            size_t n1 = std::min(p->n_out, n0 + tile) - 1;
            int span = w.i0[n1] - w.i0[n0] + SINC_LEN + 2;
            mw = std::max(mw, span);
        }
        return mw;
    };
    p->period = 8;
    if (L >= 8 && L <= 512) {
        int mw = window_for((int)L);
        if (mw <= kGenMaxWindow) p->period = (int)L;
    }
    p->max_window = window_for(p->period);
    // gen2 mapping: 64 periods per tile, period = the smallest multiple of L that is >= 32 outputs (phases repeat every L
    // outputs, so any multiple keeps "same position in the period -> same table rows" true)
    {
        const uint32_t per = L * ((31 + L) / L);
        auto window2_for = [&](size_t period) {
            int mw = 0;
            const size_t tile = period * 64;
            for (size_t n0 = 0; n0 < p->n_out; n0 += tile) {
                size_t n1 = std::min(p->n_out, n0 + tile) - 1;
                mw = std::max(mw, w.i0[n1] - w.i0[n0] + SINC_LEN + 2);
            }
            return mw;
        };
        if (per <= 1024 && p->n_out >= (size_t)per * 64) {
            const int mw = window2_for(per);
            if (mw <= kGen2MaxWindow) { p->period2 = (int)per; p->max_window2 = mw; }
        }
    }
    if (p->n_out) {
        SK_CUDA(ctx, cudaMalloc((void**)&p->d_i0, p->n_out * sizeof(int32_t)));
        SK_CUDA(ctx, cudaMalloc((void**)&p->d_s0, p->n_out));
        SK_CUDA(ctx, cudaMalloc((void**)&p->d_frac, p->n_out * sizeof(float)));
        SK_CUDA(ctx, cudaMemcpyAsync(p->d_i0, w.i0.data(), p->n_out * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaMemcpyAsync(p->d_s0, w.s0.data(), p->n_out, cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaMemcpyAsync(p->d_frac, w.frac.data(), p->n_out * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // vectors die with the caller
    }
    return SK_OK;
}

static bool integer_step_of(uint32_t in_rate, uint32_t out_rate, int* step) {
    if (in_rate % out_rate) return false;
    uint32_t s = in_rate / out_rate;
    if (s == 2 || s == 3 || s == 4 || s == 6 || s == 12) { *step = (int)s; return true; }
    return false;
}

// whole-stream plan for the batch entry points
static int build_stream_plan(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, size_t frames, int semantics, sk_plan** out) {
    // frame and output indices travel as int32 inside the kernels (12.4 hours at 48 kHz): refuse longer streams loudly
    if (frames > 0x7fff0000ull || (double)frames * (double)out_rate / (double)in_rate > 2.0e9)
        SK_FAIL(ctx, SK_E_ARG, "stream of %zu frames is too long for one call (limit 2^31 frames in and out); split it", frames);
    if (semantics != SK_SEM_ONESHOT && semantics != SK_SEM_STREAMING) SK_FAIL(ctx, SK_E_ARG, "unknown resample semantics %d", semantics);
    char key[96];
    snprintf(key, sizeof(key), "%u:%u:%zu:%d", in_rate, out_rate, frames, semantics);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) { *out = it->second; return SK_OK; }
    const double ratio = (double)out_rate / (double)in_rate;
    sk_plan* p = new sk_plan();
    int step = 0;
    p->integer_step = integer_step_of(in_rate, out_rate, &step);
    WalkOut w;
    w.record = !p->integer_step;
    if (semantics == SK_SEM_ONESHOT) {
        walk_chunk(-(double)(SINC_LEN / 2), ratio, frames, 0, w);  // SincFixedIn::new(.., chunk = frames) + one process()
    } else if (frames > 0) {  // no AudioData ever arrives for an empty stream, so no resampler exists to flush
        // StreamingResampler::process + flush (decoder lib.rs:1970-2059)
        double last = -(double)(SINC_LEN / 2);
        size_t pos = 0;
        while (frames - pos >= STREAM_CHUNK) { last = walk_chunk(last, ratio, STREAM_CHUNK, (long long)pos, w); pos += STREAM_CHUNK; }
        const size_t remaining = frames - pos;
        const size_t before = w.count;
        walk_chunk(last, ratio, STREAM_CHUNK, (long long)pos, w);  // process_partial: zero-padded (or all-zero) chunk
        if (remaining > 0) {
            const size_t padded = STREAM_CHUNK - remaining;
            const size_t trim = (size_t)llround(((double)padded * (double)out_rate) / (double)in_rate);
            const size_t got = w.count - before;
            const size_t keep = got > trim ? got - trim : 0;
            w.count = before + keep;
            if (w.record) { w.i0.resize(w.count); w.s0.resize(w.count); w.frac.resize(w.count); }
        }
    }
    if (p->integer_step) {
        p->step = step;
        p->n_out = w.count;
        p->i0_first = -(long long)(SINC_LEN / 2) + step;
    } else {
        int rc = finalize_general_plan(ctx, p, in_rate, out_rate, w);
        if (rc) { free_plan(p); return rc; }
    }
    ctx->plans[key] = p;
    *out = p;
    return SK_OK;
}

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
        const float* p0 = reinterpret_cast<const float*>(s.base + stream * s.stream_stride + (long long)lane0 * s.lane_stride);
        const float* p1 = reinterpret_cast<const float*>(s.base + stream * s.stream_stride + (long long)(lane0 + 1) * s.lane_stride);
        for (int i = threadIdx.x; i < count; i += NTHREADS) {
            const long long f = w0 + i;
            const bool ok = f >= s.valid_lo && f < s.valid_hi;
            float a = ok ? __ldg(p0 + f + s.origin) : 0.0f;
            float b = (ok && lane1_valid) ? __ldg(p1 + f + s.origin) : 0.0f;
            big |= ((__float_as_uint(a) & 0x7f800000u) >= 0x78000000u) | ((__float_as_uint(b) & 0x7f800000u) >= 0x78000000u);
            win0[i] = a; if (win1) win1[i] = b;
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
// multiples of 4.
template <> struct IntCfg<3>  { static constexpr int R = 8, SKEW = 4;  };  // S = 24: {0,24,16,8} + 4
template <> struct IntCfg<2>  { static constexpr int R = 6, SKEW = 16; };  // S = 12: {0,12,24,4} + 16
template <> struct IntCfg<4>  { static constexpr int R = 7, SKEW = 16; };  // S = 28: {0,28,24,20} + 16
template <> struct IntCfg<6>  { static constexpr int R = 6, SKEW = 16; };  // S = 36: {0,4,8,12} + 16
template <> struct IntCfg<12> { static constexpr int R = 7, SKEW = 16; };  // S = 84: {0,20,8,28} + 16

constexpr int kIntThreads = 256;                 // 128 thread-blocks x 2 lanes
constexpr int kIntBlocksPerLane = kIntThreads / 2;

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
template <int R>
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
        if (dst.pcm_bits) {  // bytes out: quantise the mono samples in registers, 8 / 12 / 16 bytes per thread when aligned
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

template <int STEP>
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
    const bool big = stage_window<kIntThreads>(src, stream, lane0, lane1_valid, w0, W, smem, smem + WS);
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
    int_epilogue<R>(y, dst, stream, n0, b, lane, lane0, lane1_valid);
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
template <int STEP, int SRCW> __host__ __device__ constexpr int tma_smem_bytes() {
    return 2 * tma_lane_stride<STEP>() * (int)sizeof(float) + 2 * tma_raw_bytes<STEP, SRCW>() + 64;
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
                // sample bytes -> result bytes 1..3 (most significant byte on top), then arithmetic >> 8 sign-extends
                const uint32_t sel = !BE ? ((uint32_t)(sh + 2) << 12) | ((uint32_t)(sh + 1) << 8) | ((uint32_t)sh << 4)
                                         : ((uint32_t)sh << 12) | ((uint32_t)(sh + 1) << 8) | ((uint32_t)(sh + 2) << 4);
                f = __int2float_rn(((int32_t)__byte_perm(lo, hi, sel)) >> 8);
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

// per-tile geometry, computed once by the TMA-issuing thread and handed to the CTA through shared memory
struct TileGeom { int stream; int n0; int wa; int g_lo; int g_hi; int pad[3]; };

#ifndef SK_TMA_CTAS_PER_SM
#define SK_TMA_CTAS_PER_SM 3
#endif

template <int STEP, int SRCW>
__global__ void __launch_bounds__(kIntThreads, SK_TMA_CTAS_PER_SM)
resample_int_tma_kernel(const __grid_constant__ Taps taps, const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, long long i0_first,
                        unsigned tiles_per_stream, unsigned total_tiles, const float* __restrict__ row1) {
    constexpr int R = IntCfg<STEP>::R;
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int LEAD = tma_lead<STEP>();
    constexpr int WF = tma_frames<STEP>();
    constexpr int WS = tma_lane_stride<STEP>();
    constexpr int FB = 2 * SRCW;  // bytes per stereo frame
    constexpr int RAWB = tma_raw_bytes<STEP, SRCW>();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* win = reinterpret_cast<float*>(smem_raw);
    uint8_t* raw = smem_raw + 2 * WS * sizeof(float);
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + 2 * RAWB);
    __shared__ TileGeom geoms[4];

    const unsigned nk = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles this CTA owns

    // one elected thread: geometry of tile k -> geoms[k & 3], then the TMA request for its raw bytes
    auto issue = [&](unsigned k) {
        const unsigned t = blockIdx.x + k * gridDim.x;
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
        const int buf = (int)(k & 1);
        const uint32_t bar = smem_u32(&bars[buf]);
        const uint32_t bytes = (uint32_t)(b_hi - b_lo);
        mbar_expect_tx(bar, bytes);
        if (bytes) tma_load_1d(smem_u32(raw + buf * RAWB + (b_lo - wa * FB)), src.base + (long long)stream * src.stream_stride + b_lo, bytes, bar);
    };

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_fence_init();
        if (nk > 0) issue(0);
        if (nk > 1) issue(1);
    }
    __syncthreads();

    const int lane = threadIdx.x & 1, b = threadIdx.x >> 1;
    for (unsigned k = 0; k < nk; k++) {
        const TileGeom g = geoms[k & 3];
        const int buf = (int)(k & 1);
        mbar_wait(smem_u32(&bars[buf]), (k >> 1) & 1);

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
        if (threadIdx.x == 0 && k + 2 < nk) issue(k + 2);

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
        int_epilogue<R>(y, dst, g.stream, g.n0, b, lane, 0, true);
        __syncthreads();  // windows free for the next tile
    }
}

// ================================================================================================
// general kernel (any ratio): one lane per CTA, phase-major tile
// ================================================================================================
constexpr int kGenThreads = 256;

__global__ void __launch_bounds__(kGenThreads)
resample_gen_kernel(const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, const int32_t* __restrict__ pi0,
                    const uint8_t* __restrict__ ps0, const float* __restrict__ pfrac, const float* __restrict__ table, int period,
                    int lanes_per_stream) {
    extern __shared__ __align__(16) float win[];
    const long long stream = blockIdx.y / lanes_per_stream;
    const int lane_ch = blockIdx.y % lanes_per_stream;
    const long long tile = (long long)period * 32;
    const long long n0 = (long long)blockIdx.x * tile;
    const long long n1 = min(dst.n_out, n0 + tile) - 1;  // last output of this tile
    const int w0 = pi0[n0];
    const int count = pi0[n1] - w0 + SINC_LEN + 2;
    stage_window<kGenThreads>(src, stream, lane_ch, false, (long long)w0, count, win, nullptr);
    __syncthreads();

    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    float* op = dst.base + stream * dst.stream_stride + (long long)lane_ch * dst.lane_stride;
    for (int phase = warp; phase < period; phase += kGenThreads / 32) {
        const long long n = n0 + (long long)q * period + phase;
        if (n > n1) continue;
        const int i0 = pi0[n];
        const int s0 = ps0[n];
        const float frac = pfrac[n];
        int s1 = s0 + 1, d = 0;
        if (s1 >= OVERSAMPLING) { s1 -= OVERSAMPLING; d = 1; }  // get_nearest_times_2
        const float* x = win + (i0 - w0);
        const float* h0 = table + s0 * SINC_LEN;
        const float* h1 = table + s1 * SINC_LEN;
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int k = 0; k < SINC_LEN; k += 8) {
            float xs[9];
#pragma unroll
            for (int l = 0; l < 9; l++) xs[l] = x[k + l];
            const float4 ha = __ldg(reinterpret_cast<const float4*>(h0 + k)), hb = __ldg(reinterpret_cast<const float4*>(h0 + k + 4));
            const float4 hc = __ldg(reinterpret_cast<const float4*>(h1 + k)), hd = __ldg(reinterpret_cast<const float4*>(h1 + k + 4));
            const float g0[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
            const float g1[8] = {hc.x, hc.y, hc.z, hc.w, hd.x, hd.y, hd.z, hd.w};
#pragma unroll
            for (int l = 0; l < 8; l++) {
                a[l] = __fmaf_rn(xs[l], g0[l], a[l]);
                c[l] = __fmaf_rn(d ? xs[l + 1] : xs[l], g1[l], c[l]);
            }
        }
        const float p0 = reduce8(a), p1 = reduce8(c);
        op[n] = __fadd_rn(p0, __fmul_rn(frac, __fsub_rn(p1, p0)));  // interp_lin
    }
}

// ================================================================================================
// general kernel v2: register tile of 2 adjacent phases x 2 periods per thread
// ================================================================================================
// The v1 kernel computes one output per thread at a time: every 8 taps cost 9 shared-memory loads and 4 table
// loads for 16 FMAs, so the load/store unit, not the FMA pipe, sets the pace (13 % of the FP32 peak).  Here a thread
// owns a 2 x 2 tile: two ADJACENT outputs (phases p, p+1: their input windows differ by DELTA = 0 or 1 samples, so
// one register window serves both) of two PERIODS (k and k + 32: same phases, hence the same four table rows).
// Per 8 taps: 16 + 2*DELTA shared-memory loads and 8 table loads feed 64 FMAs (0.39 loads per FMA instead of 0.81).
// Each of the 8 dot products keeps rubato's 8 lane accumulators and reduction tree, and `p0 + frac * (p1 - p0)` uses the
// output's own frac, so results stay bit-identical to v1 (and to the CPU restatement the tests compare against).
// The tile mapping needs warp-uniform rows: all lanes of a warp sit at the same position inside the phase period, which
// fixes (subindex, DELTA) except where the f64 walk lands within rounding of an integer; such warps (and ragged tile
// ends, and subindex 255 whose second product starts one sample later) take the v1 arithmetic per output.
constexpr int kGen2Periods = 64;
constexpr int kGen2MaxDelta = 6;  // largest window offset between the two adjacent outputs of a tile

__device__ __forceinline__ float gen_single(const float* x, const float* __restrict__ h0, const float* __restrict__ h1, int d, float frac) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int k = 0; k < SINC_LEN; k += 8) {
        float xs[9];
#pragma unroll
        for (int l = 0; l < 9; l++) xs[l] = x[k + l];
        const float4 ha = __ldg(reinterpret_cast<const float4*>(h0 + k)), hb = __ldg(reinterpret_cast<const float4*>(h0 + k + 4));
        const float4 hc = __ldg(reinterpret_cast<const float4*>(h1 + k)), hd = __ldg(reinterpret_cast<const float4*>(h1 + k + 4));
        const float g0[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        const float g1[8] = {hc.x, hc.y, hc.z, hc.w, hd.x, hd.y, hd.z, hd.w};
#pragma unroll
        for (int l = 0; l < 8; l++) {
            a[l] = __fmaf_rn(xs[l], g0[l], a[l]);
            c[l] = __fmaf_rn(d ? xs[l + 1] : xs[l], g1[l], c[l]);
        }
    }
    const float p0 = reduce8(a), p1 = reduce8(c);
    return __fadd_rn(p0, __fmul_rn(frac, __fsub_rn(p1, p0)));  // interp_lin
}

// y[phase][period] for the 2 x 2 tile; x0 / x1: windows of the two periods at phase 0's first tap; hs: the warp's four
// table rows (subindex s_a, s_a + 1, s_b, s_b + 1) staged in shared memory, so the loop's coefficient reads are
// broadcast LDS.128 with shared-memory latency instead of L1/L2 round trips
template <int DELTA>
__device__ __forceinline__ void gen2_tile(const float* x0, const float* x1, const float* hs, const float (&frac)[2][2], float (&y)[2][2]) {
    float acc[2][2][2][8];  // [phase][period][product][lane]
#pragma unroll
    for (int i = 0; i < 64; i++) (&acc[0][0][0][0])[i] = 0.0f;
    // 4 taps per step (lanes 0-3, then 4-7 of rubato's accumulators): 4 broadcast LDS.128 + 8 (+2) LDS.32 feed 32 FMAs
    // with only 26 operand registers live, which keeps the kernel at 2 CTAs per SM
#pragma unroll 1
    for (int k = 0; k < SINC_LEN; k += 8) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int kk = k + 4 * half;
            // lanes are periods: consecutive lanes read words M apart (M odd for the common ratios) -> conflict-free LDS.32.
            // (An even/odd double copy of the window read with LDS.64 was measured slower: 64-bit bank conflicts.)
            float xs[2][4 + DELTA];
#pragma unroll
            for (int l = 0; l < 4 + DELTA; l++) { xs[0][l] = x0[kk + l]; xs[1][l] = x1[kk + l]; }
            float4 g[4];
#pragma unroll
            for (int r = 0; r < 4; r++) g[r] = *reinterpret_cast<const float4*>(hs + r * SINC_LEN + kk);
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const int ln = 4 * half + l;
                const float c0 = l == 0 ? g[0].x : (l == 1 ? g[0].y : (l == 2 ? g[0].z : g[0].w));
                const float c1 = l == 0 ? g[1].x : (l == 1 ? g[1].y : (l == 2 ? g[1].z : g[1].w));
                const float c2 = l == 0 ? g[2].x : (l == 1 ? g[2].y : (l == 2 ? g[2].z : g[2].w));
                const float c3 = l == 0 ? g[3].x : (l == 1 ? g[3].y : (l == 2 ? g[3].z : g[3].w));
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    acc[0][b][0][ln] = __fmaf_rn(xs[b][l], c0, acc[0][b][0][ln]);
                    acc[0][b][1][ln] = __fmaf_rn(xs[b][l], c1, acc[0][b][1][ln]);
                    acc[1][b][0][ln] = __fmaf_rn(xs[b][l + DELTA], c2, acc[1][b][0][ln]);
                    acc[1][b][1][ln] = __fmaf_rn(xs[b][l + DELTA], c3, acc[1][b][1][ln]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const float p0 = reduce8(acc[a][b][0]), p1 = reduce8(acc[a][b][1]);
            y[a][b] = __fadd_rn(p0, __fmul_rn(frac[a][b], __fsub_rn(p1, p0)));  // interp_lin
        }
}

template <int NT>  // 256 threads x 2 CTAs/SM when two windows fit an SM, else 512 threads x 1 CTA (same 16 warps per SM)
__global__ void __launch_bounds__(NT, 512 / NT)
resample_gen2_kernel(const __grid_constant__ SrcDesc src, const __grid_constant__ OutDesc dst, const int32_t* __restrict__ pi0,
                     const uint8_t* __restrict__ ps0, const float* __restrict__ pfrac, const float* __restrict__ table, int period,
                     int lanes_per_stream) {
    extern __shared__ __align__(16) float win_all[];
    float* rows_smem = win_all;                                   // [8 warps][4 rows][256]
    float* win = win_all + (NT / 32) * 4 * SINC_LEN;
    const long long stream = blockIdx.y / lanes_per_stream;
    const int lane_ch = blockIdx.y % lanes_per_stream;
    const long long tile = (long long)period * kGen2Periods;
    const long long n0 = (long long)blockIdx.x * tile;
    const long long n1 = min(dst.n_out, n0 + tile) - 1;  // last output of this tile
    const int w0 = pi0[n0];
    const int count = pi0[n1] - w0 + SINC_LEN + 2;
    stage_window<NT>(src, stream, lane_ch, false, (long long)w0, count, win, nullptr);
    __syncthreads();

    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    float* op = dst.base + stream * dst.stream_stride + (long long)lane_ch * dst.lane_stride;
    for (int phase = 2 * warp; phase < period; phase += 2 * (NT / 32)) {
        // the tile's four outputs of this lane: n[a][b] = n0 + (q + 32 b) * period + phase + a
        long long n[2][2]; int i0[2][2], s0[2][2]; float frac[2][2]; bool ok[2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                n[a][b] = n0 + (long long)(q + 32 * b) * period + phase + a;
                ok[a][b] = phase + a < period && n[a][b] <= n1;
                i0[a][b] = ok[a][b] ? pi0[n[a][b]] : w0;
                s0[a][b] = ok[a][b] ? (int)ps0[n[a][b]] : 0;
                frac[a][b] = ok[a][b] ? pfrac[n[a][b]] : 0.0f;
            }
        // warp-uniform rows and window offset?  (reference values: lane 0, period 0)
        const int sa = __shfl_sync(0xffffffffu, s0[0][0], 0), sb = __shfl_sync(0xffffffffu, s0[1][0], 0);
        const int dl = __shfl_sync(0xffffffffu, i0[1][0] - i0[0][0], 0);
        const bool all4 = ok[0][0] && ok[0][1] && ok[1][0] && ok[1][1];
        const bool none = !(ok[0][0] || ok[0][1] || ok[1][0] || ok[1][1]);
        const bool fits = all4 && s0[0][0] == sa && s0[0][1] == sa && s0[1][0] == sb && s0[1][1] == sb && sa < OVERSAMPLING - 1 &&
                          sb < OVERSAMPLING - 1 && i0[1][0] - i0[0][0] == dl && i0[1][1] - i0[0][1] == dl && dl >= 0 && dl <= kGen2MaxDelta;
        const bool lane0_ok = __shfl_sync(0xffffffffu, (int)all4, 0) != 0;
        if (lane0_ok && __all_sync(0xffffffffu, fits || none)) {
            float y[2][2];
            const float* x0 = win + (i0[0][0] - w0);
            const float* x1 = win + (i0[0][1] - w0);
            // stage the four rows for this warp: rows sa, sa + 1 are contiguous in the table, so are sb, sb + 1
            float* hs = rows_smem + warp * 4 * SINC_LEN;
            __syncwarp();  // the previous work item's reads of hs are done
#pragma unroll
            for (int j = 0; j < 4; j++) {
                reinterpret_cast<float4*>(hs)[q + 32 * j] = __ldg(reinterpret_cast<const float4*>(table + sa * SINC_LEN) + q + 32 * j);
                reinterpret_cast<float4*>(hs + 2 * SINC_LEN)[q + 32 * j] = __ldg(reinterpret_cast<const float4*>(table + sb * SINC_LEN) + q + 32 * j);
            }
            __syncwarp();
            switch (dl) {  // adjacent outputs are floor/ceil(t_ratio) input samples apart: 0-1 when upsampling, up to 6 for 44.1 -> 8 kHz
                case 0: gen2_tile<0>(x0, x1, hs, frac, y); break;
                case 1: gen2_tile<1>(x0, x1, hs, frac, y); break;
                case 2: gen2_tile<2>(x0, x1, hs, frac, y); break;
                case 3: gen2_tile<3>(x0, x1, hs, frac, y); break;
                case 4: gen2_tile<4>(x0, x1, hs, frac, y); break;
                case 5: gen2_tile<5>(x0, x1, hs, frac, y); break;
                default: gen2_tile<6>(x0, x1, hs, frac, y); break;
            }
            if (all4) { op[n[0][0]] = y[0][0]; op[n[1][0]] = y[1][0]; op[n[0][1]] = y[0][1]; op[n[1][1]] = y[1][1]; }
        } else {
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    if (!ok[a][b]) continue;
                    int s1 = s0[a][b] + 1, d = 0;
                    if (s1 >= OVERSAMPLING) { s1 -= OVERSAMPLING; d = 1; }  // get_nearest_times_2
                    op[n[a][b]] = gen_single(win + (i0[a][b] - w0), table + s0[a][b] * SINC_LEN, table + s1 * SINC_LEN, d, frac[a][b]);
                }
        }
    }
}

// ================================================================================================
// host: launch helpers
// ================================================================================================
static int src_desc_for(sk_ctx* ctx, SrcDesc* s, const void* in, size_t in_stride, int bits, int fmt, int big_endian, int channels, int scale_mode,
                        size_t frames) {
    memset(s, 0, sizeof(*s));
    s->base = (const uint8_t*)in;
    s->stream_stride = (long long)in_stride;
    s->channels = channels;
    s->origin = 0; s->valid_lo = 0; s->valid_hi = (long long)frames;
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "Unsupported bits_per_sample: %d", bits);
    if (fmt == SK_FMT_FLOAT && bits != 32) SK_FAIL(ctx, SK_E_ARG, "floating-point PCM must contain 32-bit samples");
    s->width = bits / 8;
    s->pair_ok = (((uintptr_t)in | (uintptr_t)in_stride) & 7) == 0;
    s->byte_loads = s->width != 3 && (((uintptr_t)in | (uintptr_t)in_stride) & (uintptr_t)(s->width - 1)) != 0;
    s->big_endian = scale_mode == SK_MODE_B ? (big_endian != 0) : 0;  // mode A is always little-endian
    s->is_float = fmt == SK_FMT_FLOAT;
    s->scrub = scale_mode == SK_MODE_B;
    if (bits == 16) s->scale = 1.0f / 32768.0f;
    else if (bits == 24) s->scale = scale_mode == SK_MODE_A ? 4.656612873077393e-10f : 1.0f / 8388608.0f;
    else s->scale = 4.656612873077393e-10f;
    return SK_OK;
}

template <int STEP, int SRCW>
static int launch_int_tma(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams) {
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int smem = tma_smem_bytes<STEP, SRCW>();
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_tma_kernel<STEP, SRCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long tiles_per_stream = (long long)((plan->n_out + T - 1) / T);
    const long long total = tiles_per_stream * (long long)n_streams;
    const int per_sm = std::max(1, std::min(SK_TMA_CTAS_PER_SM, (220 * 1024) / smem));
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
    resample_int_tma_kernel<STEP, SRCW><<<(unsigned)grid, kIntThreads, smem, ctx->stream>>>(t2, s2, dst, plan->i0_first, (unsigned)tiles_per_stream,
                                                                                          (unsigned)total, d_table + SINC_LEN);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

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
        if (src.width == 2) return launch_int_tma<STEP, 2>(ctx, taps, src, dst, plan, d_table, n_streams);
        if (src.width == 3) return launch_int_tma<STEP, 3>(ctx, taps, src, dst, plan, d_table, n_streams);
        return launch_int_tma<STEP, 4>(ctx, taps, src, dst, plan, d_table, n_streams);
    }
    constexpr int T = int_tile_outputs<STEP>();
    constexpr int smem = int_smem_bytes<STEP>();
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_int_kernel<STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
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
        resample_int_kernel<STEP><<<grid, kIntThreads, smem, ctx->stream>>>(taps, s2, d2, plan->i0_first, lanes_per_stream, d_table + SINC_LEN);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

static int launch_gen2(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams,
                       int lanes_per_stream) {
    const int smem256 = (plan->max_window2 + 8 + (256 / 32) * 4 * SINC_LEN) * (int)sizeof(float);
    const bool two_ctas = 2 * (smem256 + 1024) <= 227 * 1024;
    const int nt = two_ctas ? 256 : 512;
    const int smem = (plan->max_window2 + 8 + (nt / 32) * 4 * SINC_LEN) * (int)sizeof(float);
    if (two_ctas) SK_CUDA(ctx, cudaFuncSetAttribute(resample_gen2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else SK_CUDA(ctx, cudaFuncSetAttribute(resample_gen2_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const size_t tile = (size_t)plan->period2 * kGen2Periods;
    const size_t tiles = (plan->n_out + tile - 1) / tile;
    const size_t lanes = (size_t)lanes_per_stream * n_streams;
    const size_t slab = (size_t)lanes_per_stream * (65535 / lanes_per_stream);
    for (size_t l0 = 0; l0 < lanes; l0 += slab) {
        const size_t nl = std::min<size_t>(slab, lanes - l0);
        SrcDesc s2 = src; OutDesc d2 = dst;
        const size_t stream0 = l0 / lanes_per_stream;
        s2.base += (long long)stream0 * src.stream_stride;
        d2.base += (long long)stream0 * dst.stream_stride;
        dim3 grid((unsigned)tiles, (unsigned)nl);
        if (two_ctas) resample_gen2_kernel<256><<<grid, 256, smem, ctx->stream>>>(s2, d2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, plan->period2, lanes_per_stream);
        else resample_gen2_kernel<512><<<grid, 512, smem, ctx->stream>>>(s2, d2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, plan->period2, lanes_per_stream);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

static int launch_gen(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams,
                      int lanes_per_stream) {
    const int smem = (plan->max_window + 8) * (int)sizeof(float);
    SK_CUDA(ctx, cudaFuncSetAttribute(resample_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const size_t tile = (size_t)plan->period * 32;
    const size_t tiles = (plan->n_out + tile - 1) / tile;
    const size_t lanes = (size_t)lanes_per_stream * n_streams;
    for (size_t l0 = 0; l0 < lanes; l0 += (size_t)lanes_per_stream * (65535 / lanes_per_stream)) {
        const size_t nl = std::min<size_t>((size_t)lanes_per_stream * (65535 / lanes_per_stream), lanes - l0);
        SrcDesc s2 = src; OutDesc d2 = dst;
        const size_t stream0 = l0 / lanes_per_stream;
        s2.base += (long long)stream0 * src.stream_stride;
        d2.base += (long long)stream0 * dst.stream_stride;
        dim3 grid((unsigned)tiles, (unsigned)nl);
        resample_gen_kernel<<<grid, kGenThreads, smem, ctx->stream>>>(s2, d2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, plan->period, lanes_per_stream);
        SK_LAUNCH_CHECK(ctx);
    }
    return SK_OK;
}

// run one plan over a batch.  mix != NONE requires exactly two lanes per stream for the fused path;
// the caller handles other channel counts with a separate mixdown pass.
static int run_plan(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, uint32_t in_rate, uint32_t out_rate, size_t n_streams,
                    int lanes_per_stream) {
    if (!plan->n_out || !n_streams) return SK_OK;
    const float* d_table; const std::vector<float>* h_table;
    int rc = get_sinc_table(ctx, in_rate, out_rate, &d_table, &h_table);
    if (rc) return rc;
    if (plan->integer_step) {
        switch (plan->step) {
            case 2: return launch_int<2>(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 3: return launch_int<3>(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 4: return launch_int<4>(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 6: return launch_int<6>(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 12: return launch_int<12>(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
        }
        SK_FAIL(ctx, SK_E_STATE, "internal: no kernel for integer step %d", plan->step);
    }
    if (dst.mix != SK_MIX_NONE) SK_FAIL(ctx, SK_E_STATE, "internal: general kernel does not mix");
    if (plan->period2 > 0) return launch_gen2(ctx, src, dst, plan, d_table, n_streams, lanes_per_stream);
    return launch_gen(ctx, src, dst, plan, d_table, n_streams, lanes_per_stream);
}

// ================================================================================================
// C ABI: batch chain
// ================================================================================================
extern "C" int sk_normalize_batch_plan(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, size_t frames, int semantics, size_t* n_out) {
    if (!ctx) return SK_E_ARG;
    if (!in_rate || !out_rate) SK_FAIL(ctx, SK_E_ARG, "sampling_rate must be > 0");
    if (semantics != SK_SEM_ONESHOT && semantics != SK_SEM_STREAMING) SK_FAIL(ctx, SK_E_ARG, "unknown semantics %d", semantics);
    if (frames > 0x7fff0000ull) SK_FAIL(ctx, SK_E_ARG, "stream too long (%zu frames)", frames);
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_plan* p;
    int rc = build_stream_plan(ctx, in_rate, out_rate, frames, semantics, &p);
    if (rc) return rc;
    if (n_out) *n_out = p->n_out;
    return SK_OK;
}

int sk_mixdown_strided_dev(sk_ctx* ctx, const float* planar, size_t plane_stride, size_t frames, int channels, int mix_mode, float* out);  // sk_pcm.cu

extern "C" int sk_normalize_batch_dev(sk_ctx* ctx, const sk_batch_desc* d) {
    if (!ctx || !d) return SK_E_ARG;
    if (d->channels <= 0 || d->channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (!d->in_rate || !d->out_rate) SK_FAIL(ctx, SK_E_ARG, "sampling_rate must be > 0");
    if (d->mix_mode < SK_MIX_NONE || d->mix_mode > SK_MIX_SCALE_THEN_SUM) SK_FAIL(ctx, SK_E_ARG, "unknown mix mode %d", d->mix_mode);
    if (!d->n_streams) return SK_OK;
    if (!d->in || !d->out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_plan* plan;
    int rc = build_stream_plan(ctx, d->in_rate, d->out_rate, d->frames, d->semantics, &plan);
    if (rc) return rc;
    SrcDesc src;
    if ((rc = src_desc_for(ctx, &src, d->in, d->in_stride, d->bits, d->fmt, d->big_endian, d->channels, d->scale_mode, d->frames))) return rc;
    OutDesc dst;
    dst.base = d->out; dst.stream_stride = (long long)d->out_stride; dst.lane_stride = (long long)d->out_stride_ch;
    dst.n_out = (long long)plan->n_out;
    int mix = d->channels == 1 ? SK_MIX_NONE : d->mix_mode;  // a mono source is never downmixed (emit_resampled_chunks)
    if (mix != SK_MIX_NONE && (d->channels != 2 || !plan->integer_step)) {
        // unfused: resample to a planar scratch, then mix down stream by stream
        const size_t plane = sk_round_up(plan->n_out, 4);
        const size_t per_stream = plane * d->channels;
        // bound the scratch: process streams in slabs
        size_t slab = std::max<size_t>(1, (size_t)(1u << 28) / std::max<size_t>(1, per_stream));  // <= 1 GiB of floats
        slab = std::min(slab, d->n_streams);
        if ((rc = sk_ensure_dev_out(ctx, slab * per_stream * sizeof(float)))) return rc;
        float* scratch = (float*)ctx->d_out;
        for (size_t s0 = 0; s0 < d->n_streams; s0 += slab) {
            const size_t ns = std::min(slab, d->n_streams - s0);
            SrcDesc s2 = src; s2.base += (long long)s0 * src.stream_stride;
            OutDesc d2; d2.base = scratch; d2.stream_stride = (long long)per_stream; d2.lane_stride = (long long)plane; d2.mix = SK_MIX_NONE;
            d2.n_out = (long long)plan->n_out;
            if ((rc = run_plan(ctx, s2, d2, plan, d->in_rate, d->out_rate, ns, d->channels))) return rc;
            for (size_t s = 0; s < ns; s++) {
                if ((rc = sk_mixdown_strided_dev(ctx, scratch + s * per_stream, plane, plan->n_out, d->channels, mix, d->out + (s0 + s) * d->out_stride)))
                    return rc;
            }
        }
        return SK_OK;
    }
    dst.mix = mix;
    return run_plan(ctx, src, dst, plan, d->in_rate, d->out_rate, d->n_streams, d->channels);
}

// apply_output_options with bytes out for a whole batch (soundkit-decoder/src/lib.rs:3324-3456, rate-changing arm down to
// one channel): convert -> resample -> downmix -> f32_channels_to_bytes.  Two-channel integer-ratio sources leave the fused
// kernel as bytes directly; every other combination goes through an f32 scratch and the interleave kernel.
int sk_planar_f32_to_pcm_dev_impl(sk_ctx* ctx, const float* planar, size_t frames, int channels, int bits, int fmt, uint8_t* out);
extern "C" int sk_normalize_batch_pcm_dev(sk_ctx* ctx, const sk_batch_desc* d, int out_bits, int out_fmt, void* out, size_t out_stride_bytes) {
    if (!ctx || !d) return SK_E_ARG;
    if (out_fmt == SK_FMT_FLOAT && out_bits != 32) SK_FAIL(ctx, SK_E_ARG, "PCMFloat output requires 32-bit samples");
    if (out_bits != 16 && out_bits != 24 && out_bits != 32) SK_FAIL(ctx, SK_E_ARG, "Unsupported output bits per sample: %d", out_bits);
    if (d->channels <= 0 || d->channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (d->channels != 1 && d->mix_mode == SK_MIX_NONE) SK_FAIL(ctx, SK_E_ARG, "byte output needs a mono result: set mix_mode (or use sk_planar_f32_to_pcm_dev)");
    if (!d->in_rate || !d->out_rate) SK_FAIL(ctx, SK_E_ARG, "sampling_rate must be > 0");
    if (!d->n_streams) return SK_OK;
    if (!d->in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_plan* plan;
    int rc = build_stream_plan(ctx, d->in_rate, d->out_rate, d->frames, d->semantics, &plan);
    if (rc) return rc;
    if (!plan->n_out) return SK_OK;
    const size_t ob = (size_t)out_bits / 8;
    if (out_stride_bytes < plan->n_out * ob) SK_FAIL(ctx, SK_E_ARG, "out_stride_bytes %zu < %zu bytes per stream", out_stride_bytes, plan->n_out * ob);
    if (d->channels == 2 && plan->integer_step) {
        SrcDesc src;
        if ((rc = src_desc_for(ctx, &src, d->in, d->in_stride, d->bits, d->fmt, d->big_endian, d->channels, d->scale_mode, d->frames))) return rc;
        OutDesc dst;
        dst.base = nullptr; dst.stream_stride = 0; dst.lane_stride = 0; dst.mix = d->mix_mode; dst.n_out = (long long)plan->n_out;
        dst.pcm_bits = out_bits; dst.pcm_float = out_fmt == SK_FMT_FLOAT; dst.pcm_base = (uint8_t*)out; dst.pcm_stream_stride = (long long)out_stride_bytes;
        return run_plan(ctx, src, dst, plan, d->in_rate, d->out_rate, d->n_streams, d->channels);
    }
    // general case: mono f32 into a scratch (slabs of streams), then the interleave kernel with one channel
    const size_t plane = sk_round_up(plan->n_out, 4);
    size_t slab = std::max<size_t>(1, (size_t)(1u << 27) / plane);  // <= 512 MiB of floats
    slab = std::min(slab, d->n_streams);
    if (ctx->d_tmp_cap < slab * plane * sizeof(float)) {
        if (ctx->d_tmp) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_tmp); ctx->d_tmp = nullptr; ctx->d_tmp_cap = 0; }
        SK_CUDA(ctx, cudaMalloc(&ctx->d_tmp, slab * plane * sizeof(float)));
        ctx->d_tmp_cap = slab * plane * sizeof(float);
    }
    float* mono = (float*)ctx->d_tmp;
    for (size_t s0 = 0; s0 < d->n_streams; s0 += slab) {
        const size_t ns = std::min(slab, d->n_streams - s0);
        sk_batch_desc dd = *d;
        dd.in = (const char*)d->in + s0 * d->in_stride; dd.n_streams = ns; dd.out = mono; dd.out_stride = plane; dd.out_stride_ch = plane;
        if ((rc = sk_normalize_batch_dev(ctx, &dd))) return rc;
        for (size_t s = 0; s < ns; s++)
            if ((rc = sk_planar_f32_to_pcm_dev_impl(ctx, mono + s * plane, plan->n_out, 1, out_bits, out_fmt, (uint8_t*)out + (s0 + s) * out_stride_bytes))) return rc;
    }
    return SK_OK;
}

extern "C" int sk_normalize_stream(sk_ctx* ctx, const uint8_t* in, size_t bytes, int bits, int fmt, int big_endian, int channels, int scale_mode,
                                   uint32_t in_rate, uint32_t out_rate, int semantics, int mix_mode, float* out, size_t out_cap_per_plane,
                                   size_t* n_out) {
    if (!ctx) return SK_E_ARG;
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (bits != 16 && bits != 24 && bits != 32) SK_FAIL(ctx, SK_E_ARG, "Unsupported bits_per_sample: %d", bits);
    const size_t B = (size_t)bits / 8;
    if (scale_mode == SK_MODE_B && (bytes % B != 0 || (bytes / B) % (size_t)channels != 0))
        SK_FAIL(ctx, SK_E_ARG, "PCM data is unsupported or contains a partial frame");
    const size_t frames = bytes / (B * channels);
    size_t n = 0;
    int rc = sk_normalize_batch_plan(ctx, in_rate, out_rate, frames, semantics, &n);
    if (rc) return rc;
    if (n_out) *n_out = n;
    if (!n) return SK_OK;
    if (n > out_cap_per_plane) SK_FAIL(ctx, SK_E_ARG, "output capacity %zu < %zu frames", out_cap_per_plane, n);
    if (!in || !out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    const int mix = channels == 1 ? SK_MIX_NONE : mix_mode;
    const int planes = mix == SK_MIX_NONE ? channels : 1;
    const size_t plane = sk_round_up(n, 4);
    // device staging: input bytes in d_in; output planes in a private allocation (d_out may be used as mix scratch)
    const size_t in_bytes = frames * B * channels;
    if ((rc = sk_ensure_dev_in(ctx, in_bytes + 16))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, in_bytes))) return rc;
    float* d_res = nullptr;
    SK_CUDA(ctx, cudaMalloc((void**)&d_res, plane * planes * sizeof(float)));
    sk_batch_desc d;
    memset(&d, 0, sizeof(d));
    d.in = ctx->d_in; d.in_stride = in_bytes; d.n_streams = 1; d.frames = frames; d.bits = bits; d.fmt = fmt; d.big_endian = big_endian;
    d.channels = channels; d.scale_mode = scale_mode; d.in_rate = in_rate; d.out_rate = out_rate; d.semantics = semantics; d.mix_mode = mix;
    d.out = d_res; d.out_stride = plane * planes; d.out_stride_ch = plane;
    rc = sk_normalize_batch_dev(ctx, &d);
    for (int c = 0; c < planes && !rc; c++) rc = sk_d2h(ctx, out + (size_t)c * out_cap_per_plane, d_res + (size_t)c * plane, n * sizeof(float));
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_res);
    return rc;
}

// Host-pointer batch: the same fused chain for n_streams host-resident streams, pipelined in slabs —
// H2D of slab i+1 and D2H of slab i-1 overlap the kernel of slab i on three streams.  `in` / `out`
// of the descriptor are HOST pointers here (pinned memory gives full PCIe rate; pageable works).
struct BatchPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    void* d_in[2] = {nullptr, nullptr}; size_t in_cap = 0;
    float* d_out[2] = {nullptr, nullptr}; size_t out_cap = 0;
};
// one per context, owned by the context (contexts are thread-confined; no process-wide state, so pipelines on different
// threads never meet)
void sk_resample_free_pipe(sk_ctx* ctx) {
    if (!ctx->batch_pipe) return;
    BatchPipe& p = *static_cast<BatchPipe*>(ctx->batch_pipe);
    for (int b = 0; b < 2; b++) {
        if (p.d_in[b]) cudaFree(p.d_in[b]);
        if (p.d_out[b]) cudaFree(p.d_out[b]);
        if (p.ev_in[b]) cudaEventDestroy(p.ev_in[b]);
        if (p.ev_k[b]) cudaEventDestroy(p.ev_k[b]);
        if (p.ev_out[b]) cudaEventDestroy(p.ev_out[b]);
    }
    if (p.s_in) cudaStreamDestroy(p.s_in);
    if (p.s_out) cudaStreamDestroy(p.s_out);
    delete static_cast<BatchPipe*>(ctx->batch_pipe);
    ctx->batch_pipe = nullptr;
}

extern "C" int sk_normalize_batch(sk_ctx* ctx, const sk_batch_desc* d) {
    if (!ctx || !d) return SK_E_ARG;
    if (d->channels <= 0 || d->channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (d->bits != 16 && d->bits != 24 && d->bits != 32) SK_FAIL(ctx, SK_E_ARG, "Unsupported bits_per_sample: %d", d->bits);
    if (!d->n_streams) return SK_OK;
    if (!d->in || !d->out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_plan* plan;
    int rc = build_stream_plan(ctx, d->in_rate, d->out_rate, d->frames, d->semantics, &plan);
    if (rc) return rc;
    if (!plan->n_out) return SK_OK;
    const int mix = d->channels == 1 ? SK_MIX_NONE : d->mix_mode;
    const int planes = mix == SK_MIX_NONE ? d->channels : 1;
    const size_t in_bytes = sk_round_up(d->frames * (size_t)(d->bits / 8) * d->channels, 16);
    const size_t plane = sk_round_up(plan->n_out, 4);
    const size_t out_floats = plane * planes;
    // slab: as many streams as fit ~256 MiB of input (at least one)
    size_t slab = std::max<size_t>(1, ((size_t)256 << 20) / std::max<size_t>(1, in_bytes));
    slab = std::min(slab, d->n_streams);
    if (!ctx->batch_pipe) ctx->batch_pipe = new BatchPipe();
    BatchPipe& p = *static_cast<BatchPipe*>(ctx->batch_pipe);
    if (!p.s_in) {
        SK_CUDA(ctx, cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
        SK_CUDA(ctx, cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            SK_CUDA(ctx, cudaEventCreateWithFlags(&p.ev_in[b], cudaEventDisableTiming));
            SK_CUDA(ctx, cudaEventCreateWithFlags(&p.ev_k[b], cudaEventDisableTiming));
            SK_CUDA(ctx, cudaEventCreateWithFlags(&p.ev_out[b], cudaEventDisableTiming));
        }
    }
    if (p.in_cap < slab * in_bytes || p.out_cap < slab * out_floats) {
        cudaDeviceSynchronize();
        for (int b = 0; b < 2; b++) {
            if (p.d_in[b]) cudaFree(p.d_in[b]);
            if (p.d_out[b]) cudaFree(p.d_out[b]);
            p.d_in[b] = nullptr; p.d_out[b] = nullptr;
        }
        p.in_cap = slab * in_bytes; p.out_cap = slab * out_floats;
        for (int b = 0; b < 2; b++) {
            SK_CUDA(ctx, cudaMalloc(&p.d_in[b], p.in_cap));
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_out[b], p.out_cap * sizeof(float)));
        }
    }
    const size_t raw_bytes = d->frames * (size_t)(d->bits / 8) * d->channels;
    size_t islab = 0;
    for (size_t s0 = 0; s0 < d->n_streams; s0 += slab, islab++) {
        const int b = (int)(islab & 1);
        const size_t ns = std::min(slab, d->n_streams - s0);
        if (islab >= 2) {
            SK_CUDA(ctx, cudaStreamWaitEvent(p.s_in, p.ev_k[b], 0));        // kernel of slab-2 finished reading d_in[b]
            SK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p.ev_out[b], 0)); // D2H of slab-2 finished reading d_out[b]
        }
        for (size_t s = 0; s < ns; s++)
            SK_CUDA(ctx, cudaMemcpyAsync((char*)p.d_in[b] + s * in_bytes, (const char*)d->in + (s0 + s) * d->in_stride, raw_bytes,
                                         cudaMemcpyHostToDevice, p.s_in));
        SK_CUDA(ctx, cudaEventRecord(p.ev_in[b], p.s_in));
        SK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p.ev_in[b], 0));
        sk_batch_desc dd = *d;
        dd.in = p.d_in[b]; dd.in_stride = in_bytes; dd.n_streams = ns; dd.mix_mode = mix;
        dd.out = p.d_out[b]; dd.out_stride = out_floats; dd.out_stride_ch = plane;
        if ((rc = sk_normalize_batch_dev(ctx, &dd))) return rc;
        SK_CUDA(ctx, cudaEventRecord(p.ev_k[b], ctx->stream));
        SK_CUDA(ctx, cudaStreamWaitEvent(p.s_out, p.ev_k[b], 0));
        for (size_t s = 0; s < ns; s++)
            for (int c = 0; c < planes; c++)
                SK_CUDA(ctx, cudaMemcpyAsync(d->out + (s0 + s) * d->out_stride + (size_t)c * d->out_stride_ch, p.d_out[b] + s * out_floats + (size_t)c * plane,
                                             plan->n_out * sizeof(float), cudaMemcpyDeviceToHost, p.s_out));
        SK_CUDA(ctx, cudaEventRecord(p.ev_out[b], p.s_out));
    }
    SK_CUDA(ctx, cudaStreamSynchronize(p.s_out));
    SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SK_OK;
}

// ================================================================================================
// C ABI: stateful per-stream resampler (rubato SincFixedIn)
// ================================================================================================
struct sk_resampler {
    sk_ctx* ctx;
    uint32_t in_rate, out_rate;
    double ratio;
    int channels;
    size_t chunk;
    double last_index;
    float* d_buf = nullptr;   // [channels][2*SINC_LEN + chunk]: history then the current chunk
    float* d_out = nullptr;   // [channels][out_max rounded up]
    size_t out_plane = 0;
    sk_plan plan;             // reused per chunk (general path arrays sized for out_max)
    size_t plan_cap = 0;
};

// buffer.copy_within(chunk .. chunk + 2*sinc_len, 0): ranges may overlap when chunk < 2*sinc_len
__global__ void history_shift_kernel(float* buf, size_t plane, size_t chunk) {
    float* b = buf + (size_t)blockIdx.x * plane;
    const float v = b[chunk + threadIdx.x];
    __syncthreads();
    b[threadIdx.x] = v;
}

extern "C" int sk_resampler_create(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, int channels, size_t chunk, const sk_resample_params* params,
                                   sk_resampler** out) {
    if (!ctx || !out) return SK_E_ARG;
    *out = nullptr;
    if (!in_rate || !out_rate) SK_FAIL(ctx, SK_E_ARG, "sampling_rate must be > 0");
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (params && (params->sinc_len != SINC_LEN || params->oversampling != OVERSAMPLING || params->f_cutoff != 0.95f || params->interp != 0 ||
                   params->window != 0))
        SK_FAIL(ctx, SK_E_ARG, "only soundkit's parameters are implemented: sinc_len 256, oversampling 256, f_cutoff 0.95, Linear, BlackmanHarris2");
    if (chunk > 0x7fff0000ull) SK_FAIL(ctx, SK_E_ARG, "chunk too long");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_resampler* r = new (std::nothrow) sk_resampler();
    if (!r) SK_FAIL(ctx, SK_E_ALLOC, "out of memory");
    r->ctx = ctx; r->in_rate = in_rate; r->out_rate = out_rate; r->ratio = (double)out_rate / (double)in_rate;
    r->channels = channels; r->chunk = chunk; r->last_index = -(double)(SINC_LEN / 2);
    const size_t plane = 2 * SINC_LEN + chunk;
    const size_t omax = (size_t)((double)chunk * r->ratio * 2.0 + 10.0);
    r->out_plane = sk_round_up(omax, 4);
    if (cudaMalloc((void**)&r->d_buf, plane * channels * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void**)&r->d_out, r->out_plane * channels * sizeof(float)) != cudaSuccess) {
        sk_resampler_destroy(r);
        SK_FAIL(ctx, SK_E_ALLOC, "device allocation failed");
    }
    cudaMemsetAsync(r->d_buf, 0, plane * channels * sizeof(float), ctx->stream);
    int step;
    r->plan.integer_step = integer_step_of(in_rate, out_rate, &step);
    r->plan.step = r->plan.integer_step ? step : 0;
    int rc = get_sinc_table(ctx, in_rate, out_rate, nullptr, nullptr);
    if (rc) { sk_resampler_destroy(r); return rc; }
    *out = r;
    return SK_OK;
}

extern "C" void sk_resampler_destroy(sk_resampler* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    if (r->d_buf) cudaFree(r->d_buf);
    if (r->d_out) cudaFree(r->d_out);
    if (r->plan.d_i0) cudaFree(r->plan.d_i0);
    if (r->plan.d_s0) cudaFree(r->plan.d_s0);
    if (r->plan.d_frac) cudaFree(r->plan.d_frac);
    delete r;
}

extern "C" size_t sk_resampler_out_max(const sk_resampler* r) { return r ? (size_t)((double)r->chunk * r->ratio * 2.0 + 10.0) : 0; }
extern "C" double sk_resampler_last_index(const sk_resampler* r) { return r ? r->last_index : 0.0; }

extern "C" int sk_resampler_copy_table(const sk_resampler* r, float* out65536) {
    if (!r || !out65536) return SK_E_ARG;
    const std::vector<float>* h;
    int rc = get_sinc_table(r->ctx, r->in_rate, r->out_rate, nullptr, &h);
    if (rc) return rc;
    memcpy(out65536, h->data(), h->size() * sizeof(float));
    return SK_OK;
}

static int resampler_run(sk_resampler* r, const float* const* in, size_t frames, float* const* out, size_t* n_out) {
    sk_ctx* ctx = r->ctx;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t plane = 2 * SINC_LEN + r->chunk;
    // history shift, then the new chunk (zero padded to `chunk`)
    history_shift_kernel<<<r->channels, 2 * SINC_LEN, 0, ctx->stream>>>(r->d_buf, plane, r->chunk);
    SK_LAUNCH_CHECK(ctx);
    int rc;
    for (int c = 0; c < r->channels; c++) {
        float* dst = r->d_buf + (size_t)c * plane + 2 * SINC_LEN;
        if (in && frames) {
            if ((rc = sk_h2d(ctx, dst, in[c], frames * sizeof(float)))) return rc;
        }
        if (frames < r->chunk) SK_CUDA(ctx, cudaMemsetAsync(dst + frames, 0, (r->chunk - frames) * sizeof(float), ctx->stream));
    }
    // index walk of this chunk
    WalkOut w;
    w.record = !r->plan.integer_step;
    const double next_last = walk_chunk(r->last_index, r->ratio, r->chunk, 0, w);
    const size_t n = w.count;
    sk_plan& p = r->plan;
    p.n_out = n;
    if (p.integer_step) {
        // idx is an exact integer on this path
        p.i0_first = (long long)r->last_index + p.step;
    } else if (n) {
        if (n > r->plan_cap) {
            if (p.d_i0) cudaFree(p.d_i0);
            if (p.d_s0) cudaFree(p.d_s0);
            if (p.d_frac) cudaFree(p.d_frac);
            r->plan_cap = n + 64;
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_i0, r->plan_cap * sizeof(int32_t)));
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_s0, r->plan_cap));
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_frac, r->plan_cap * sizeof(float)));
        }
        SK_CUDA(ctx, cudaMemcpyAsync(p.d_i0, w.i0.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaMemcpyAsync(p.d_s0, w.s0.data(), n, cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaMemcpyAsync(p.d_frac, w.frac.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // tile mapping: same rule as the batch plans
        const uint32_t g = gcd_u32(r->in_rate, r->out_rate);
        const uint32_t L = r->out_rate / g;
        auto window_for = [&](int period) {
            int mw = 0;
            const size_t tile = (size_t)period * 32;
            for (size_t a = 0; a < n; a += tile) {
                size_t b = std::min(n, a + tile) - 1;
                mw = std::max(mw, w.i0[b] - w.i0[a] + SINC_LEN + 2);
            }
            return mw;
        };
        p.period = 8;
        if (L >= 8 && L <= 512 && window_for((int)L) <= kGenMaxWindow) p.period = (int)L;
        p.max_window = window_for(p.period);
    }
    if (n) {
        SrcDesc src;
        memset(&src, 0, sizeof(src));
        src.base = (const uint8_t*)r->d_buf; src.stream_stride = 0; src.lane_stride = (long long)(plane * sizeof(float));
        src.width = 0; src.channels = r->channels; src.scale = 1.0f;
        src.origin = 2 * SINC_LEN; src.valid_lo = -(long long)(2 * SINC_LEN); src.valid_hi = (long long)r->chunk;
        OutDesc dst;
        dst.base = r->d_out; dst.stream_stride = 0; dst.lane_stride = (long long)r->out_plane; dst.mix = SK_MIX_NONE; dst.n_out = (long long)n;
        if ((rc = run_plan(ctx, src, dst, &p, r->in_rate, r->out_rate, 1, r->channels))) return rc;
        for (int c = 0; c < r->channels; c++)
            if ((rc = sk_d2h(ctx, out[c], r->d_out + (size_t)c * r->out_plane, n * sizeof(float)))) return rc;
    }
    r->last_index = next_last;
    if (n_out) *n_out = n;
    return SK_OK;
}

extern "C" int sk_resampler_process(sk_resampler* r, const float* const* in, float* const* out, size_t* n_out) {
    if (!r) return SK_E_ARG;
    if (!in || !out) SK_FAIL(r->ctx, SK_E_ARG, "null buffer");
    return resampler_run(r, in, r->chunk, out, n_out);
}

extern "C" int sk_resampler_process_partial(sk_resampler* r, const float* const* in_or_null, size_t frames, float* const* out, size_t* n_out) {
    if (!r) return SK_E_ARG;
    if (!out) SK_FAIL(r->ctx, SK_E_ARG, "null buffer");
    if (frames > r->chunk) frames = r->chunk;
    return resampler_run(r, in_or_null, in_or_null ? frames : 0, out, n_out);
}

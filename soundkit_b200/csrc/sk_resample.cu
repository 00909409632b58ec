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

#include "sk_resample_dev.cuh"


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

// four word-shifted, zero-padded copies of the table for 16-byte row staging (layout: sk_resample_gen3.cu)
int sk_get_shifted_sinc_table(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, int pad_l, int row_len, const float** d_table4) {
    const uint64_t key = rate_key(in_rate, out_rate);
    auto it = ctx->sinc_tables4.find(key);
    if (it == ctx->sinc_tables4.end()) {
        const std::vector<float>* h;
        int rc = get_sinc_table(ctx, in_rate, out_rate, nullptr, &h);
        if (rc) return rc;
        std::vector<float> g((size_t)4 * OVERSAMPLING * row_len, 0.0f);
        for (int sh = 0; sh < 4; sh++)
            for (int r = 0; r < OVERSAMPLING; r++)
                memcpy(&g[((size_t)sh * OVERSAMPLING + r) * row_len + pad_l + sh], h->data() + (size_t)r * SINC_LEN, SINC_LEN * sizeof(float));
        float* d = nullptr;
        SK_CUDA(ctx, cudaMalloc((void**)&d, g.size() * sizeof(float)));
        SK_CUDA(ctx, cudaMemcpyAsync(d, g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->sinc_tables4.emplace(key, d).first;
    }
    *d_table4 = it->second;
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
    for (auto& kv : ctx->sinc_tables4) cudaFree(kv.second);
    ctx->sinc_tables4.clear();
    for (auto& kv : ctx->plans) free_plan(kv.second);
    ctx->plans.clear();
}

static const int kGenMaxWindow = 14400;  // frames of shared memory window the general kernel may use (57.6 KB)
static const int kGen2MaxWindow = 40000;  // resample_gen2_kernel: up to 160 KB of window + 64 KB of staged table rows (512-thread form)

// finish a general plan from recorded walk arrays: choose the tile period and upload
static int finalize_general_plan(sk_ctx* ctx, sk_plan* p, uint32_t in_rate, uint32_t out_rate, const WalkOut& w) {
    p->n_out = w.i0.size();
    p->max_step3 = 0;
    for (size_t n = 0; n + 3 < p->n_out; n++) p->max_step3 = std::max(p->max_step3, w.i0[n + 3] - w.i0[n]);
    p->max_step7 = 0;
    for (size_t n = 0; n + 7 < p->n_out; n++) p->max_step7 = std::max(p->max_step7, w.i0[n + 7] - w.i0[n]);
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
    if (!in_rate || !out_rate || in_rate % out_rate) return false;
    uint32_t s = in_rate / out_rate;
    if (s == 2 || s == 3 || s == 4 || s == 6 || s == 12) { *step = (int)s; return true; }
    return false;
}

// The plan cache is bounded: a general-ratio plan keeps 9 bytes per output frame on the device (0.5 GB for an hour at
// 16 kHz) and callers with variable-length inputs (one plan per distinct length) would otherwise grow it for the life of
// the context.  Least-recently-used plans go first, never the one just built; kernels still reading an evicted plan's
// arrays are drained first.
static const size_t kPlanCacheMaxEntries = 64;
static const size_t kPlanCacheMaxBytes = (size_t)768 << 20;
static size_t plan_dev_bytes(const sk_plan* p) { return p->d_i0 ? p->n_out * 9 : 0; }
static void evict_plans(sk_ctx* ctx, const sk_plan* keep) {
    bool drained = false;
    for (;;) {
        size_t bytes = 0;
        for (auto& kv : ctx->plans) bytes += plan_dev_bytes(kv.second);
        if (ctx->plans.size() <= kPlanCacheMaxEntries && bytes <= kPlanCacheMaxBytes) return;
        auto victim = ctx->plans.end();
        for (auto it = ctx->plans.begin(); it != ctx->plans.end(); ++it)
            if (it->second != keep && (victim == ctx->plans.end() || it->second->last_use < victim->second->last_use)) victim = it;
        if (victim == ctx->plans.end()) return;  // only the plan in use is left
        if (!drained && victim->second->d_i0) { cudaStreamSynchronize(ctx->stream); drained = true; }
        free_plan(victim->second);
        ctx->plans.erase(victim);
    }
}

// whole-stream plan for the batch entry points
static int build_stream_plan(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, size_t frames, int semantics, sk_plan** out) {
    // every entry point funnels through here: no caller can reach the ratio arithmetic with a zero rate
    // (downsample_audio: "sampling_rate must be > 0", audio_pipeline.rs:449; apply_output_options: "Output sample rate must be > 0")
    if (!in_rate || !out_rate) SK_FAIL(ctx, SK_E_ARG, "sampling_rate must be > 0");
    // frame and output indices travel as int32 inside the kernels (12.4 hours at 48 kHz): refuse longer streams loudly
    if (frames > 0x7fff0000ull || (double)frames * (double)out_rate / (double)in_rate > 2.0e9)
        SK_FAIL(ctx, SK_E_ARG, "stream of %zu frames is too long for one call (limit 2^31 frames in and out); split it", frames);
    if (semantics != SK_SEM_ONESHOT && semantics != SK_SEM_STREAMING) SK_FAIL(ctx, SK_E_ARG, "unknown resample semantics %d", semantics);
    char key[96];
    snprintf(key, sizeof(key), "%u:%u:%zu:%d", in_rate, out_rate, frames, semantics);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) { it->second->last_use = ++ctx->plan_clock; *out = it->second; return SK_OK; }
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
    p->last_use = ++ctx->plan_clock;
    ctx->plans[key] = p;
    *out = p;
    evict_plans(ctx, p);
    return SK_OK;
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
                     int lanes_per_stream, unsigned long long* __restrict__ stats) {
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
            const int nfb = (int)ok[0][0] + (int)ok[0][1] + (int)ok[1][0] + (int)ok[1][1];
            if (nfb) atomicAdd(stats, (unsigned long long)nfb);
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

static int launch_gen2(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams,
                       int lanes_per_stream) {
    if (int rc = sk_ensure_stats(ctx)) return rc;
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
        if (two_ctas) resample_gen2_kernel<256><<<grid, 256, smem, ctx->stream>>>(s2, d2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, plan->period2, lanes_per_stream, ctx->d_stats);
        else resample_gen2_kernel<512><<<grid, 512, smem, ctx->stream>>>(s2, d2, plan->d_i0, plan->d_s0, plan->d_frac, d_table, plan->period2, lanes_per_stream, ctx->d_stats);
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
// one translation unit per integer step (sk_resample_int<STEP>.cu)
#define SK_DECL_INT(N) int sk_launch_int_##N(sk_ctx*, const SrcDesc&, const OutDesc&, const sk_plan*, const float*, const float*, size_t, int);
SK_DECL_INT(2) SK_DECL_INT(3) SK_DECL_INT(4) SK_DECL_INT(6) SK_DECL_INT(12)
#undef SK_DECL_INT

int sk_launch_gen3(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, uint32_t in_rate, uint32_t out_rate,
                   size_t n_streams, int lanes_per_stream);

static int run_plan(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, uint32_t in_rate, uint32_t out_rate, size_t n_streams,
                    int lanes_per_stream) {
    if (!plan->n_out || !n_streams) return SK_OK;
    const float* d_table; const std::vector<float>* h_table;
    int rc = get_sinc_table(ctx, in_rate, out_rate, &d_table, &h_table);
    if (rc) return rc;
    if (plan->integer_step) {
        switch (plan->step) {
            case 2: return sk_launch_int_2(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 3: return sk_launch_int_3(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 4: return sk_launch_int_4(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 6: return sk_launch_int_6(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
            case 12: return sk_launch_int_12(ctx, src, dst, plan, h_table->data(), d_table, n_streams, lanes_per_stream);
        }
        SK_FAIL(ctx, SK_E_STATE, "internal: no kernel for integer step %d", plan->step);
    }
    if (dst.mix != SK_MIX_NONE) SK_FAIL(ctx, SK_E_STATE, "internal: general kernel does not mix");
    if (!ctx->disable_gen3) {  // packed 4 x 2 tile kernel (sk_resample_gen3.cu); 1 = does not apply to this plan / batch
        rc = sk_launch_gen3(ctx, src, dst, plan, d_table, in_rate, out_rate, n_streams, lanes_per_stream);
        if (rc <= 0) return rc;
    }
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

// sk_pcm.cu: interleaved PCM -> planar f32 with an explicit plane stride; whole-batch mixdown / mono quantiser (blockIdx.y = stream)
int sk_pcm_to_planar_rows_dev(sk_ctx* ctx, const uint8_t* in, size_t in_stride, size_t n_streams, size_t frames, int bits, int fmt, int big_endian,
                              int channels, int scale_mode, float* out, size_t out_stride, size_t plane);
int sk_mixdown_batch_dev(sk_ctx* ctx, const float* planar, size_t in_stream_stride, size_t plane_stride, size_t frames, int channels, int mix_mode,
                         float* out, size_t out_stream_stride, size_t n_streams);
int sk_quantize_mono_batch_dev(sk_ctx* ctx, const float* in, size_t in_stream_stride, size_t n, int bits, int fmt, uint8_t* out, size_t out_stream_bytes,
                               size_t n_streams);

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
    // Mono and multi-channel sources at integer ratios: TWO passes instead of the scalar-staging FIR (0.16 / 0.29 of FP32 peak for
    // mono / 6 channels) — (1) interleaved PCM -> planar f32 rows with the HBM-rate conversion kernels, (2) the persistent TMA FIR
    // for planar rows (resample_int_planar_kernel), two lanes per tile: the channel pairs of a stream, or pairs of mono streams —
    // in slabs that bound the scratch.  Same values: both forms convolve the same converted f32 samples in rubato's order.
    if (plan->integer_step && d->channels != 2 && d->frames >= 4 && plan->n_out > 0) {
        const size_t in_plane = sk_round_up(d->frames, 4);           // floats between the planes of a stream: every plane 16-byte aligned
        const size_t row_floats = in_plane * (size_t)d->channels;   // one stream's planes
        const size_t plane = sk_round_up(plan->n_out, 4), per_stream_out = plane * d->channels;
        size_t slab = std::max<size_t>(1, ((size_t)1 << 28) / std::max(row_floats, per_stream_out));  // <= 1 GiB per scratch
        slab = std::min(slab, d->n_streams);
        if (slab > 1 && d->channels == 1) slab &= ~(size_t)1;  // mono streams are convolved in pairs
        if (ctx->d_rows_cap < slab * row_floats * sizeof(float)) {
            if (ctx->d_rows) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_rows); ctx->d_rows = nullptr; ctx->d_rows_cap = 0; }
            const size_t cap = sk_round_up(slab * row_floats * sizeof(float), (size_t)1 << 20);
            if (cudaMalloc(&ctx->d_rows, cap) != cudaSuccess) SK_FAIL(ctx, SK_E_ALLOC, "device alloc of %zu bytes failed", cap);
            ctx->d_rows_cap = cap;
        }
        float* rows = (float*)ctx->d_rows;
        float* scratch = nullptr;
        if (mix != SK_MIX_NONE) {
            if ((rc = sk_ensure_dev_out(ctx, slab * per_stream_out * sizeof(float)))) return rc;
            scratch = (float*)ctx->d_out;
        }
        for (size_t s0 = 0; s0 < d->n_streams; s0 += slab) {
            const size_t ns = std::min(slab, d->n_streams - s0);
            if ((rc = sk_pcm_to_planar_rows_dev(ctx, (const uint8_t*)d->in + s0 * d->in_stride, d->in_stride, ns, d->frames, d->bits, d->fmt, d->big_endian,
                                                d->channels, d->scale_mode, rows, row_floats, in_plane)))
                return rc;
            SrcDesc s2;
            memset(&s2, 0, sizeof(s2));
            s2.base = (const uint8_t*)rows; s2.stream_stride = (long long)(row_floats * sizeof(float)); s2.lane_stride = (long long)(in_plane * sizeof(float));
            s2.width = 0; s2.channels = d->channels; s2.scale = 1.0f; s2.origin = 0; s2.valid_lo = 0; s2.valid_hi = (long long)d->frames;
            OutDesc d2;
            d2.n_out = (long long)plan->n_out; d2.mix = SK_MIX_NONE;
            if (mix != SK_MIX_NONE) { d2.base = scratch; d2.stream_stride = (long long)per_stream_out; d2.lane_stride = (long long)plane; }
            else { d2.base = d->out + s0 * d->out_stride; d2.stream_stride = (long long)d->out_stride; d2.lane_stride = (long long)d->out_stride_ch; }
            if (d->channels == 1 && ns >= 2) {
                // pairs of mono streams as the two lanes of a tile (an odd last stream goes alone)
                SrcDesc sp = s2; OutDesc dp = d2;
                sp.stream_stride = 2 * s2.stream_stride; sp.lane_stride = s2.stream_stride; sp.channels = 2;
                dp.stream_stride = 2 * d2.stream_stride; dp.lane_stride = d2.stream_stride;
                if ((rc = run_plan(ctx, sp, dp, plan, d->in_rate, d->out_rate, ns / 2, 2))) return rc;
                if (ns & 1) {
                    SrcDesc sl = s2; OutDesc dl = d2;
                    sl.base += (long long)(ns - 1) * s2.stream_stride; dl.base += (long long)(ns - 1) * d2.stream_stride;
                    if ((rc = run_plan(ctx, sl, dl, plan, d->in_rate, d->out_rate, 1, 1))) return rc;
                }
            } else if ((rc = run_plan(ctx, s2, d2, plan, d->in_rate, d->out_rate, ns, d->channels))) {
                return rc;
            }
            if (mix != SK_MIX_NONE &&
                (rc = sk_mixdown_batch_dev(ctx, scratch, per_stream_out, plane, plan->n_out, d->channels, mix, d->out + s0 * d->out_stride, d->out_stride, ns)))
                return rc;
        }
        return SK_OK;
    }
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
            // one launch mixes the whole slab (blockIdx.y = stream)
            if ((rc = sk_mixdown_batch_dev(ctx, scratch, per_stream, plane, plan->n_out, d->channels, mix, d->out + s0 * d->out_stride, d->out_stride, ns)))
                return rc;
        }
        return SK_OK;
    }
    dst.mix = mix;
    return run_plan(ctx, src, dst, plan, d->in_rate, d->out_rate, d->n_streams, d->channels);
}

// apply_output_options with bytes out for a whole batch (soundkit-decoder/src/lib.rs:3324-3456, rate-changing arm down to
// one channel): convert -> resample -> downmix -> f32_channels_to_bytes.  Two-channel integer-ratio sources leave the fused
// kernel as bytes directly; every other combination goes through an f32 scratch and the interleave kernel.
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
    if (d->channels == 2 && plan->integer_step && (plan->step == 2 || plan->step == 3 || plan->step == 6)) {
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
        if ((rc = sk_quantize_mono_batch_dev(ctx, mono, plane, plan->n_out, out_bits, out_fmt, (uint8_t*)out + s0 * out_stride_bytes, out_stride_bytes, ns))) return rc;
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
    // short clips (an utterance, a decoder's AudioData block) are latency-bound calls: the chain reads the clip from the pinned staging
    // buffer and writes its result planes straight back into it (zero copy): one synchronize, no staging DMA
    char* const s_in = sk_stage_half(ctx, 0);
    char* const s_out = sk_stage_half(ctx, 1);
    if (s_in && in_bytes <= ((size_t)1 << 20) && plane * planes * sizeof(float) <= kSkStageHalf) {
        memcpy(s_in, in, in_bytes);
        sk_batch_desc d;
        memset(&d, 0, sizeof(d));
        d.in = s_in; d.in_stride = in_bytes; d.n_streams = 1; d.frames = frames; d.bits = bits; d.fmt = fmt; d.big_endian = big_endian;
        d.channels = channels; d.scale_mode = scale_mode; d.in_rate = in_rate; d.out_rate = out_rate; d.semantics = semantics; d.mix_mode = mix;
        d.out = reinterpret_cast<float*>(s_out); d.out_stride = plane * planes; d.out_stride_ch = plane;
        if ((rc = sk_normalize_batch_dev(ctx, &d))) return rc;
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < planes; c++) memcpy(out + (size_t)c * out_cap_per_plane, reinterpret_cast<const float*>(s_out) + (size_t)c * plane, n * sizeof(float));
        return SK_OK;
    }
    if ((rc = sk_ensure_dev_in(ctx, in_bytes + 16))) return rc;
    if ((rc = sk_h2d(ctx, ctx->d_in, in, in_bytes))) return rc;
    // result planes in the context's grow-only d_res (d_out may serve as the mix scratch of the unfused arms): this is the
    // call downsample_audio / the decoder's per-AudioData stage bind to, so it must not pay cudaMalloc + cudaFree per call
    if (ctx->d_res_cap < plane * planes * sizeof(float)) {
        if (ctx->d_res) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_res); ctx->d_res = nullptr; ctx->d_res_cap = 0; }
        const size_t cap = sk_round_up(plane * planes * sizeof(float), (size_t)1 << 20);
        if (cudaMalloc(&ctx->d_res, cap) != cudaSuccess) SK_FAIL(ctx, SK_E_ALLOC, "device alloc of %zu bytes failed", cap);
        ctx->d_res_cap = cap;
    }
    float* d_res = (float*)ctx->d_res;
    sk_batch_desc d;
    memset(&d, 0, sizeof(d));
    d.in = ctx->d_in; d.in_stride = in_bytes; d.n_streams = 1; d.frames = frames; d.bits = bits; d.fmt = fmt; d.big_endian = big_endian;
    d.channels = channels; d.scale_mode = scale_mode; d.in_rate = in_rate; d.out_rate = out_rate; d.semantics = semantics; d.mix_mode = mix;
    d.out = d_res; d.out_stride = plane * planes; d.out_stride_ch = plane;
    rc = sk_normalize_batch_dev(ctx, &d);
    for (int c = 0; c < planes && !rc; c++) rc = sk_d2h(ctx, out + (size_t)c * out_cap_per_plane, d_res + (size_t)c * plane, n * sizeof(float));
    return rc;
}

// Host-pointer batch: the same fused chain for n_streams host-resident streams, pipelined in slabs —
// H2D of slab i+1 and D2H of slab i-1 overlap the kernel of slab i on three streams.  `in` / `out`
// of the descriptor are HOST pointers here (pinned memory gives full PCIe rate; pageable works).
constexpr int kPipeDepth = 3;  // slabs in flight: H2D of slab i+1, kernel of slab i and D2H of slab i-1 overlap with one slab of slack
struct BatchPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kPipeDepth] = {}, ev_k[kPipeDepth] = {}, ev_out[kPipeDepth] = {};
    void* d_in[kPipeDepth] = {}; size_t in_cap = 0;
    float* d_out[kPipeDepth] = {}; size_t out_cap = 0;
};
// one per context, owned by the context (contexts are thread-confined; no process-wide state, so pipelines on different
// threads never meet)
static void pipe_free_slabs(BatchPipe& p) {
    for (int b = 0; b < kPipeDepth; b++) {
        if (p.d_in[b]) cudaFree(p.d_in[b]);
        if (p.d_out[b]) cudaFree(p.d_out[b]);
        p.d_in[b] = nullptr; p.d_out[b] = nullptr;
    }
    p.in_cap = 0; p.out_cap = 0;
}
void sk_resample_free_pipe(sk_ctx* ctx) {
    if (!ctx->batch_pipe) return;
    BatchPipe& p = *static_cast<BatchPipe*>(ctx->batch_pipe);
    pipe_free_slabs(p);
    for (int b = 0; b < kPipeDepth; b++) {
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
    if (!d->in_rate || !d->out_rate) SK_FAIL(ctx, SK_E_ARG, "sampling_rate must be > 0");
    if (d->mix_mode < SK_MIX_NONE || d->mix_mode > SK_MIX_SCALE_THEN_SUM) SK_FAIL(ctx, SK_E_ARG, "unknown mix mode %d", d->mix_mode);
    if (!d->n_streams) return SK_OK;
    if (!d->in || !d->out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_plan* plan;
    int rc = build_stream_plan(ctx, d->in_rate, d->out_rate, d->frames, d->semantics, &plan);
    if (rc) return rc;
    if (!plan->n_out) return SK_OK;
    const int mix = d->channels == 1 ? SK_MIX_NONE : d->mix_mode;
    const int planes = mix == SK_MIX_NONE ? d->channels : 1;
    const size_t raw_bytes = d->frames * (size_t)(d->bits / 8) * d->channels;
    const size_t in_bytes = sk_round_up(raw_bytes, 16);
    const size_t plane = sk_round_up(plan->n_out, 4);
    const size_t out_floats = plane * planes;
    // slab: as many streams as fit the configured slab size (default 256 MiB of input; at least one)
    size_t slab = std::max<size_t>(1, ctx->batch_slab_bytes / std::max<size_t>(1, in_bytes));
    slab = std::min(slab, d->n_streams);
    if (!ctx->batch_pipe) ctx->batch_pipe = new BatchPipe();
    BatchPipe& p = *static_cast<BatchPipe*>(ctx->batch_pipe);
    if (!p.s_in) SK_CUDA(ctx, cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
    if (!p.s_out) SK_CUDA(ctx, cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
    for (int b = 0; b < kPipeDepth; b++) {
        if (!p.ev_in[b]) SK_CUDA(ctx, cudaEventCreateWithFlags(&p.ev_in[b], cudaEventDisableTiming));
        if (!p.ev_k[b]) SK_CUDA(ctx, cudaEventCreateWithFlags(&p.ev_k[b], cudaEventDisableTiming));
        if (!p.ev_out[b]) SK_CUDA(ctx, cudaEventCreateWithFlags(&p.ev_out[b], cudaEventDisableTiming));
    }
    if (p.in_cap < slab * in_bytes || p.out_cap < slab * out_floats) {
        // grow: the capacities are committed only after every allocation succeeded, so a failed call leaves an empty
        // (not a half-allocated) pipeline behind
        cudaDeviceSynchronize();
        const size_t want_in = std::max(p.in_cap, slab * in_bytes), want_out = std::max(p.out_cap, slab * out_floats);
        pipe_free_slabs(p);
        for (int b = 0; b < kPipeDepth; b++) {
            if (cudaMalloc(&p.d_in[b], want_in) != cudaSuccess || cudaMalloc((void**)&p.d_out[b], want_out * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                pipe_free_slabs(p);
                SK_FAIL(ctx, SK_E_ALLOC, "device allocation of the batch pipeline failed (%d x (%zu + %zu) bytes)", kPipeDepth, want_in,
                        want_out * sizeof(float));
            }
        }
        p.in_cap = want_in; p.out_cap = want_out;
    }
    // a stream's rows are copied with one call when both sides are dense (the common case), else plane by plane
    const bool out_dense = planes == 1 || d->out_stride_ch == plan->n_out;
    // the slab loop; whatever it returns, no copy may still be reading d->in or writing d->out when this call returns
    auto pump = [&]() -> int {
        size_t islab = 0;
        for (size_t s0 = 0; s0 < d->n_streams; s0 += slab, islab++) {
            const int b = (int)(islab % kPipeDepth);
            const size_t ns = std::min(slab, d->n_streams - s0);
            if (islab >= (size_t)kPipeDepth) {
                SK_CUDA(ctx, cudaStreamWaitEvent(p.s_in, p.ev_k[b], 0));        // the kernel that last read d_in[b] has finished
                SK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p.ev_out[b], 0)); // the D2H that last read d_out[b] has finished
            }
            if (d->in_stride == raw_bytes && in_bytes == raw_bytes)
                SK_CUDA(ctx, cudaMemcpyAsync(p.d_in[b], (const char*)d->in + s0 * d->in_stride, ns * raw_bytes, cudaMemcpyHostToDevice, p.s_in));
            else
                SK_CUDA(ctx, cudaMemcpy2DAsync(p.d_in[b], in_bytes, (const char*)d->in + s0 * d->in_stride, d->in_stride, raw_bytes, ns,
                                               cudaMemcpyHostToDevice, p.s_in));
            SK_CUDA(ctx, cudaEventRecord(p.ev_in[b], p.s_in));
            SK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p.ev_in[b], 0));
            sk_batch_desc dd = *d;
            dd.in = p.d_in[b]; dd.in_stride = in_bytes; dd.n_streams = ns; dd.mix_mode = mix;
            dd.out = p.d_out[b]; dd.out_stride = out_floats; dd.out_stride_ch = plane;
            if (int rc2 = sk_normalize_batch_dev(ctx, &dd)) return rc2;
            SK_CUDA(ctx, cudaEventRecord(p.ev_k[b], ctx->stream));
            SK_CUDA(ctx, cudaStreamWaitEvent(p.s_out, p.ev_k[b], 0));
            if (out_dense && planes == 1) {
                SK_CUDA(ctx, cudaMemcpy2DAsync(d->out + s0 * d->out_stride, d->out_stride * sizeof(float), p.d_out[b], out_floats * sizeof(float),
                                               plan->n_out * sizeof(float), ns, cudaMemcpyDeviceToHost, p.s_out));
            } else {
                for (size_t s = 0; s < ns; s++)
                    SK_CUDA(ctx, cudaMemcpy2DAsync(d->out + (s0 + s) * d->out_stride, d->out_stride_ch * sizeof(float), p.d_out[b] + s * out_floats,
                                                   plane * sizeof(float), plan->n_out * sizeof(float), (size_t)planes, cudaMemcpyDeviceToHost, p.s_out));
            }
            SK_CUDA(ctx, cudaEventRecord(p.ev_out[b], p.s_out));
        }
        return SK_OK;
    };
    rc = pump();
    const cudaError_t e_in = cudaStreamSynchronize(p.s_in), e_k = cudaStreamSynchronize(ctx->stream), e_out = cudaStreamSynchronize(p.s_out);
    if (rc) return rc;
    SK_CUDA(ctx, e_in);
    SK_CUDA(ctx, e_k);
    SK_CUDA(ctx, e_out);
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

static int rows_copy(sk_ctx* ctx, float* dst, size_t dst_pitch, const float* src, size_t src_pitch, size_t count, size_t rows);

static int resampler_run(sk_resampler* r, const float* const* in, size_t frames, float* const* out, size_t* n_out) {
    sk_ctx* ctx = r->ctx;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t plane = 2 * SINC_LEN + r->chunk;
    // history shift, then the new chunk (zero padded to `chunk`)
    history_shift_kernel<<<r->channels, 2 * SINC_LEN, 0, ctx->stream>>>(r->d_buf, plane, r->chunk);
    SK_LAUNCH_CHECK(ctx);
    int rc;
    // A 4096-frame chunk is a latency-bound call: the chunk goes through the pinned staging buffer, ONE copy kernel reads it from
    // there (zero copy) into every channel's row, the FIR writes its outputs straight back into pinned memory and the call
    // synchronises once (history: an H2D and a D2H staging copy per channel, each with its own synchronize: 75-87 us per call).
    char* const s_in = sk_stage_half(ctx, 0);
    char* const s_out = sk_stage_half(ctx, 1);
    const bool zero_copy = s_in && (size_t)r->channels * r->chunk * sizeof(float) <= kSkStageHalf &&
                           (size_t)r->channels * r->out_plane * sizeof(float) <= kSkStageHalf;
    if (zero_copy) {
        for (int c = 0; c < r->channels; c++) {
            float* row = reinterpret_cast<float*>(s_in) + (size_t)c * r->chunk;
            if (in && frames) memcpy(row, in[c], frames * sizeof(float));
            if (frames < r->chunk) memset(row + (in ? frames : 0), 0, (r->chunk - (in ? frames : 0)) * sizeof(float));
        }
        if ((rc = rows_copy(ctx, r->d_buf + 2 * SINC_LEN, plane, reinterpret_cast<const float*>(s_in), r->chunk, r->chunk, (size_t)r->channels))) return rc;
    } else {
        for (int c = 0; c < r->channels; c++) {
            float* dst = r->d_buf + (size_t)c * plane + 2 * SINC_LEN;
            if (in && frames) {
                if ((rc = sk_h2d(ctx, dst, in[c], frames * sizeof(float)))) return rc;
            }
            if (frames < r->chunk) SK_CUDA(ctx, cudaMemsetAsync(dst + frames, 0, (r->chunk - frames) * sizeof(float), ctx->stream));
        }
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
        dst.base = zero_copy ? reinterpret_cast<float*>(s_out) : r->d_out;
        dst.stream_stride = 0; dst.lane_stride = (long long)r->out_plane; dst.mix = SK_MIX_NONE; dst.n_out = (long long)n;
        if ((rc = run_plan(ctx, src, dst, &p, r->in_rate, r->out_rate, 1, r->channels))) return rc;
        if (zero_copy) {
            SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int c = 0; c < r->channels; c++) memcpy(out[c], reinterpret_cast<const float*>(s_out) + (size_t)c * r->out_plane, n * sizeof(float));
        } else {
            for (int c = 0; c < r->channels; c++)
                if ((rc = sk_d2h(ctx, out[c], r->d_out + (size_t)c * r->out_plane, n * sizeof(float)))) return rc;
        }
    } else if (zero_copy) {
        SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the staged chunk must be consumed before the next call overwrites it
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

// ================================================================================================
// C ABI: batched stateful streaming resampler — N StreamingResamplers in lock step
// ================================================================================================
// soundkit-decoder/src/lib.rs:1917-2060: `StreamingResampler { resampler: SincFixedIn(chunk 4096), accum, accum_start }`.
// One handle carries that state for n_streams live streams on the device: per (stream, channel) lane a row
// [2 * SINC_LEN history | pending input]; the f64 index of rubato's walk is data independent, so ONE host-side walk serves
// every stream of the batch.  A process call appends the pushed frames, runs every whole 4096-frame chunk that is pending in
// ONE launch for all streams (the chunks' outputs are contiguous, like StreamingResampler::process returning its pushes), and
// moves the new history + remainder to the front of the second row buffer (ping-pong: no overlapping copy).
struct sk_resampler_batch {
    sk_ctx* ctx;
    uint32_t in_rate, out_rate;
    double ratio;
    int channels;
    size_t n_streams, max_push, row;   // row = 4 + 2 * SINC_LEN + STREAM_CHUNK + max_push floats per lane
    float* d_rows[2] = {nullptr, nullptr};
    int cur = 0;
    size_t pad = 0;                    // the history starts at element `pad` (0..3) of a row: integer ratios keep every tile window of the
                                       // next run on a 16-byte boundary of the row, so the planar TMA kernel can fetch it
    size_t pending = 0;                // frames waiting behind the history in d_rows[cur]
    double last_index;
    sk_plan plan;                      // per call (general ratios: arrays re-uploaded per call)
    size_t plan_cap = 0;
};

extern "C" int sk_resampler_batch_create(sk_ctx* ctx, uint32_t in_rate, uint32_t out_rate, int channels, size_t n_streams, size_t max_push_frames,
                                         sk_resampler_batch** out) {
    if (!ctx || !out) return SK_E_ARG;
    *out = nullptr;
    if (!in_rate) SK_FAIL(ctx, SK_E_ARG, "Input sample rate must be > 0");
    if (!out_rate) SK_FAIL(ctx, SK_E_ARG, "Output sample rate must be > 0");
    if (channels <= 0 || channels > 255) SK_FAIL(ctx, SK_E_ARG, "Channel count must be > 0");
    if (!n_streams || !max_push_frames) SK_FAIL(ctx, SK_E_ARG, "n_streams and max_push_frames must be > 0");
    if (max_push_frames > ((size_t)1 << 26) || n_streams * (size_t)channels > ((size_t)1 << 30)) SK_FAIL(ctx, SK_E_ARG, "batch too large");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    sk_resampler_batch* b = new (std::nothrow) sk_resampler_batch();
    if (!b) SK_FAIL(ctx, SK_E_ALLOC, "out of memory");
    b->ctx = ctx; b->in_rate = in_rate; b->out_rate = out_rate; b->ratio = (double)out_rate / (double)in_rate;
    b->channels = channels; b->n_streams = n_streams; b->max_push = max_push_frames;
    b->row = sk_round_up(4 + 2 * SINC_LEN + STREAM_CHUNK + max_push_frames, 4);
    b->last_index = -(double)(SINC_LEN / 2);
    const size_t bytes = n_streams * (size_t)channels * b->row * sizeof(float);
    for (int i = 0; i < 2; i++) {
        if (cudaMalloc((void**)&b->d_rows[i], bytes) != cudaSuccess) {
            cudaGetLastError();
            sk_resampler_batch_destroy(b);
            SK_FAIL(ctx, SK_E_ALLOC, "device allocation of %zu bytes failed", bytes);
        }
    }
    // SincFixedIn::new zero-fills its buffer: the initial history
    cudaMemsetAsync(b->d_rows[0], 0, bytes, ctx->stream);
    int step;
    b->plan.integer_step = integer_step_of(in_rate, out_rate, &step);
    b->plan.step = b->plan.integer_step ? step : 0;
    int rc = get_sinc_table(ctx, in_rate, out_rate, nullptr, nullptr);
    if (rc) { sk_resampler_batch_destroy(b); return rc; }
    *out = b;
    return SK_OK;
}

extern "C" void sk_resampler_batch_destroy(sk_resampler_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (int i = 0; i < 2; i++)
        if (b->d_rows[i]) cudaFree(b->d_rows[i]);
    if (b->plan.d_i0) cudaFree(b->plan.d_i0);
    if (b->plan.d_s0) cudaFree(b->plan.d_s0);
    if (b->plan.d_frac) cudaFree(b->plan.d_frac);
    delete b;
}

// frames per channel one process / flush call can produce at most (every call's `out` rows must hold this many)
extern "C" size_t sk_resampler_batch_out_cap(const sk_resampler_batch* b) {
    if (!b) return 0;
    const size_t chunks = (STREAM_CHUNK + b->max_push) / STREAM_CHUNK;
    return chunks * (size_t)((double)STREAM_CHUNK * b->ratio * 2.0 + 10.0);
}
extern "C" size_t sk_resampler_batch_pending(const sk_resampler_batch* b) { return b ? b->pending : 0; }
extern "C" double sk_resampler_batch_last_index(const sk_resampler_batch* b) { return b ? b->last_index : 0.0; }

// run `chunks` rubato chunks over d_rows[cur] (frames [0, valid) of the pending region are data, the rest reads as zero), write
// `keep` of the produced frames per lane, advance last_index
static int batch_run_chunks(sk_resampler_batch* b, size_t chunks, size_t valid, float* out, size_t out_stream_stride, size_t out_ch_stride, size_t drop_tail,
                            size_t* n_out) {
    sk_ctx* ctx = b->ctx;
    sk_plan& p = b->plan;
    WalkOut w;
    w.record = !p.integer_step;
    double last = b->last_index;
    const double first_last = last;
    for (size_t c = 0; c < chunks; c++) last = walk_chunk(last, b->ratio, STREAM_CHUNK, (long long)(c * STREAM_CHUNK), w);
    size_t n = w.count;
    n = n > drop_tail ? n - drop_tail : 0;
    p.n_out = n;
    if (p.integer_step) {
        p.i0_first = (long long)first_last + p.step;  // idx is an exact integer on this path and advances by `step` across chunk borders
    } else if (n) {
        if (n > b->plan_cap) {
            cudaStreamSynchronize(ctx->stream);
            if (p.d_i0) cudaFree(p.d_i0);
            if (p.d_s0) cudaFree(p.d_s0);
            if (p.d_frac) cudaFree(p.d_frac);
            p.d_i0 = nullptr; p.d_s0 = nullptr; p.d_frac = nullptr;
            b->plan_cap = 0;
            const size_t cap = n + n / 8 + 64;
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_i0, cap * sizeof(int32_t)));
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_s0, cap));
            SK_CUDA(ctx, cudaMalloc((void**)&p.d_frac, cap * sizeof(float)));
            b->plan_cap = cap;
        }
        int rc;
        if ((rc = sk_h2d(ctx, p.d_i0, w.i0.data(), n * sizeof(int32_t)))) return rc;
        if ((rc = sk_h2d(ctx, p.d_s0, w.s0.data(), n))) return rc;
        if ((rc = sk_h2d(ctx, p.d_frac, w.frac.data(), n * sizeof(float)))) return rc;
        p.max_step3 = 0;
        for (size_t i = 0; i + 3 < n; i++) p.max_step3 = std::max(p.max_step3, w.i0[i + 3] - w.i0[i]);
        p.max_step7 = 0;
        for (size_t i = 0; i + 7 < n; i++) p.max_step7 = std::max(p.max_step7, w.i0[i + 7] - w.i0[i]);
        // tile mapping of the one-output-per-thread kernel (batches below 64 lanes): same rule as the batch plans
        const uint32_t g = gcd_u32(b->in_rate, b->out_rate);
        const uint32_t L = b->out_rate / g;
        auto window_for = [&](int period) {
            int mw = 0;
            const size_t tile = (size_t)period * 32;
            for (size_t a = 0; a < n; a += tile) mw = std::max(mw, w.i0[std::min(n, a + tile) - 1] - w.i0[a] + SINC_LEN + 2);
            return mw;
        };
        p.period = 8;
        if (L >= 8 && L <= 512 && window_for((int)L) <= kGenMaxWindow) p.period = (int)L;
        p.max_window = window_for(p.period);
        p.period2 = 0;
    }
    if (n) {
        SrcDesc src;
        memset(&src, 0, sizeof(src));
        src.base = (const uint8_t*)b->d_rows[b->cur];
        src.stream_stride = (long long)((size_t)b->channels * b->row * sizeof(float));
        src.lane_stride = (long long)(b->row * sizeof(float));
        src.width = 0; src.channels = b->channels; src.scale = 1.0f;
        src.origin = 2 * SINC_LEN + (long long)b->pad; src.valid_lo = -(long long)(2 * SINC_LEN); src.valid_hi = (long long)valid;
        OutDesc dst;
        dst.base = out; dst.stream_stride = (long long)out_stream_stride; dst.lane_stride = (long long)out_ch_stride; dst.mix = SK_MIX_NONE;
        dst.n_out = (long long)n;
        int rc;
        if (p.integer_step && b->channels == 1 && b->n_streams >= 2) {
            // mono streams: the integer-step kernels convolve two lanes per tile, so pairs of streams go as the two lanes of one
            // (an odd last stream alone) instead of leaving the second lane of every tile empty
            SrcDesc sp = src; OutDesc dp = dst;
            sp.stream_stride = 2 * src.stream_stride; sp.lane_stride = src.stream_stride; sp.channels = 2;
            dp.stream_stride = 2 * dst.stream_stride; dp.lane_stride = dst.stream_stride;
            rc = run_plan(ctx, sp, dp, &p, b->in_rate, b->out_rate, b->n_streams / 2, 2);
            if (!rc && (b->n_streams & 1)) {
                SrcDesc sl = src; OutDesc dl = dst;
                sl.base += (long long)(b->n_streams - 1) * src.stream_stride; dl.base += (long long)(b->n_streams - 1) * dst.stream_stride;
                rc = run_plan(ctx, sl, dl, &p, b->in_rate, b->out_rate, 1, 1);
            }
        } else {
            rc = run_plan(ctx, src, dst, &p, b->in_rate, b->out_rate, b->n_streams, b->channels);
        }
        if (rc) return rc;
    }
    b->last_index = last;
    if (n_out) *n_out = n;
    return SK_OK;
}

// rows x `count` floats between pitched planes, 4 floats per thread (cudaMemcpy2DAsync moves these 16 KB rows at a third of the HBM
// rate): 128-bit accesses where the row starts allow it, scalar otherwise — coalesced either way
__global__ void __launch_bounds__(256) rows_copy_kernel(float* __restrict__ dst, size_t dst_pitch, const float* __restrict__ src, size_t src_pitch, size_t count,
                                                        size_t rows, int vec_ld, int vec_st) {
    const size_t per_row = (count + 3) / 4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_row * rows) return;
    const size_t r = i / per_row, c = 4 * (i - r * per_row);
    const float* sp = src + r * src_pitch + c;
    float* dp = dst + r * dst_pitch + c;
    if (c + 4 <= count) {
        float4 v;
        if (vec_ld) v = __ldg(reinterpret_cast<const float4*>(sp));
        else { v.x = __ldg(sp); v.y = __ldg(sp + 1); v.z = __ldg(sp + 2); v.w = __ldg(sp + 3); }
        if (vec_st) *reinterpret_cast<float4*>(dp) = v;
        else { dp[0] = v.x; dp[1] = v.y; dp[2] = v.z; dp[3] = v.w; }
    } else {
        for (size_t k = c; k < count; k++) dst[r * dst_pitch + k] = __ldg(src + r * src_pitch + k);
    }
}
static int rows_copy(sk_ctx* ctx, float* dst, size_t dst_pitch, const float* src, size_t src_pitch, size_t count, size_t rows) {
    if (!count || !rows) return SK_OK;
    const size_t threads = ((count + 3) / 4) * rows;
    const int vec_ld = ((uintptr_t)src % 16) == 0 && (src_pitch % 4) == 0, vec_st = ((uintptr_t)dst % 16) == 0 && (dst_pitch % 4) == 0;
    rows_copy_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(dst, dst_pitch, src, src_pitch, count, rows, vec_ld, vec_st);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

// new history (the last 2 * SINC_LEN frames consumed) + the unconsumed remainder -> front of the other row buffer
static int batch_carry(sk_resampler_batch* b, size_t consumed, size_t total) {
    sk_ctx* ctx = b->ctx;
    const size_t lanes = b->n_streams * (size_t)b->channels;
    const size_t keep = 2 * SINC_LEN + (total - consumed);
    // integer ratios: the next run's first window frame is (long long)last_index + step - step % 8 before frame 0, i.e. congruent to
    // last_index modulo 4 — start the history where that lands on a 16-byte boundary
    const size_t new_pad = b->plan.integer_step ? (size_t)((-(long long)b->last_index) & 3) : 0;
    if (int rc = rows_copy(ctx, b->d_rows[b->cur ^ 1] + new_pad, b->row, b->d_rows[b->cur] + b->pad + consumed, b->row, keep, lanes)) return rc;
    b->pad = new_pad;
    b->cur ^= 1;
    b->pending = total - consumed;
    return SK_OK;
}

// StreamingResampler::process (soundkit-decoder/src/lib.rs:1970-2015) for every stream of the batch: in[s][c][0 .. frames) are
// device f32 planes (strides in floats); out[s][c][0 .. *n_out) receives what the reference's returned pushes concatenate to
extern "C" int sk_resampler_batch_process_dev(sk_resampler_batch* b, const float* in, size_t in_stream_stride, size_t in_ch_stride, size_t frames, float* out,
                                              size_t out_stream_stride, size_t out_ch_stride, size_t* n_out) {
    if (!b) return SK_E_ARG;
    sk_ctx* ctx = b->ctx;
    if (n_out) *n_out = 0;
    if (frames > b->max_push) SK_FAIL(ctx, SK_E_ARG, "push of %zu frames exceeds max_push_frames %zu", frames, b->max_push);
    if (!frames) return SK_OK;
    if (!out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    // append behind the pending frames: one 2-D copy per channel plane set (in == NULL: the producer has written the frames in
    // place, see sk_resampler_batch_input_dev)
    float* row0 = b->d_rows[b->cur] + b->pad + 2 * SINC_LEN + b->pending;
    if (!in) {
    } else if (in_ch_stride * (size_t)b->channels == in_stream_stride || b->channels == 1 || b->n_streams == 1) {
        // lanes are equally spaced in the source: a single 2-D copy
        const size_t lane_pitch = b->channels == 1 ? in_stream_stride : in_ch_stride;
        if (int rc = rows_copy(ctx, row0, b->row, in, lane_pitch, frames, b->n_streams * (size_t)b->channels)) return rc;
    } else {
        for (int c = 0; c < b->channels; c++)
            if (int rc = rows_copy(ctx, row0 + (size_t)c * b->row, (size_t)b->channels * b->row, in + (size_t)c * in_ch_stride, in_stream_stride, frames, b->n_streams))
                return rc;
    }
    const size_t total = b->pending + frames;
    const size_t chunks = total / STREAM_CHUNK;
    if (!chunks) { b->pending = total; return SK_OK; }
    int rc = batch_run_chunks(b, chunks, chunks * STREAM_CHUNK, out, out_stream_stride, out_ch_stride, 0, n_out);
    if (rc) return rc;
    return batch_carry(b, chunks * STREAM_CHUNK, total);
}

// Where the next push's frames go inside the handle's row buffer: lane (stream s, channel c) starts at *in + (s * channels + c) *
// *lane_pitch floats.  A device-side producer that writes there and then calls process_dev with in == NULL saves the append copy
// (a quarter of the time of a 4096-frame push).  Valid until the next process / flush call on this handle.
extern "C" int sk_resampler_batch_input_dev(sk_resampler_batch* b, float** in, size_t* lane_pitch) {
    if (!b || !in || !lane_pitch) return SK_E_ARG;
    *in = b->d_rows[b->cur] + b->pad + 2 * SINC_LEN + b->pending;
    *lane_pitch = b->row;
    return SK_OK;
}

// StreamingResampler::flush (soundkit-decoder/src/lib.rs:2017-2059): a pending remainder is zero padded to one chunk and the
// padding's share of the output (round(padded * out / in) frames) is dropped; with nothing pending one chunk of zeros is run
extern "C" int sk_resampler_batch_flush_dev(sk_resampler_batch* b, float* out, size_t out_stream_stride, size_t out_ch_stride, size_t* n_out) {
    if (!b) return SK_E_ARG;
    sk_ctx* ctx = b->ctx;
    if (n_out) *n_out = 0;
    if (!out) SK_FAIL(ctx, SK_E_ARG, "null buffer");
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t remaining = b->pending;
    const size_t lanes = b->n_streams * (size_t)b->channels;
    // the zero padding becomes part of the next history: make it real
    SK_CUDA(ctx, cudaMemset2DAsync(b->d_rows[b->cur] + b->pad + 2 * SINC_LEN + remaining, b->row * sizeof(float), 0, (STREAM_CHUNK - remaining) * sizeof(float), lanes,
                                   ctx->stream));
    size_t trim = 0;
    if (remaining > 0) trim = (size_t)llround(((double)(STREAM_CHUNK - remaining) * (double)b->out_rate) / (double)b->in_rate);
    int rc = batch_run_chunks(b, 1, STREAM_CHUNK, out, out_stream_stride, out_ch_stride, trim, n_out);
    if (rc) return rc;
    return batch_carry(b, STREAM_CHUNK, STREAM_CHUNK);
}

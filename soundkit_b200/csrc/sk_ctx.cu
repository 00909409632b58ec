// sk_ctx.cu — context, pinned staging, error plumbing, FFMA peak microbenchmark.
#include <stdlib.h>

#include "sk_common.cuh"

extern void sk_resample_free_ctx(sk_ctx* ctx);  // sk_resample.cu: frees cached tables / plans

extern "C" int sk_abi_version(void) { return SK_ABI_VERSION; }

extern "C" int sk_ctx_create(int device, sk_ctx** out) {
    if (!out) return SK_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return SK_E_NODEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return SK_E_NODEVICE;
    sk_ctx* ctx = new (std::nothrow) sk_ctx();
    if (!ctx) return SK_E_ALLOC;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        ctx->num_sms = prop.multiProcessorCount;
        if (prop.major != 10) {  // sm_100a cubin only: fail loudly instead of a confusing launch error
            delete ctx;
            return SK_E_NODEVICE;
        }
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SK_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    { const char* e = getenv("SK_DISABLE_GEN3"); ctx->disable_gen3 = e && e[0] == '1'; }  // A/B switch for measurements
    *out = ctx;
    return SK_OK;
}

extern "C" void sk_ctx_destroy(sk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    sk_resample_free_ctx(ctx);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    if (ctx->d_res) cudaFree(ctx->d_res);
    if (ctx->d_rows) cudaFree(ctx->d_rows);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    for (int i = 0; i < 2; i++)
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    if (ctx->d_tickets) cudaFree(ctx->d_tickets);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char* sk_last_error(const sk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int sk_ctx_set_stream(sk_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SK_E_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;  // NULL is CUDA's legacy default stream, a valid choice
    return SK_OK;
}

extern "C" int sk_ctx_reset_stream(sk_ctx* ctx) {
    if (!ctx) return SK_E_ARG;
    ctx->stream = ctx->own_stream;
    return SK_OK;
}

extern "C" int sk_ctx_synchronize(sk_ctx* ctx) {
    if (!ctx) return SK_E_ARG;
    SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SK_OK;
}

extern "C" uint64_t sk_ctx_launch_count(const sk_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int sk_ctx_set_batch_slab_bytes(sk_ctx* ctx, size_t bytes) {
    if (!ctx) return SK_E_ARG;
    if (bytes < ((size_t)1 << 20) || bytes > ((size_t)16 << 30)) SK_FAIL(ctx, SK_E_ARG, "slab size must be between 1 MiB and 16 GiB");
    ctx->batch_slab_bytes = bytes;
    return SK_OK;
}

extern "C" int sk_ctx_set_resample_mode(sk_ctx* ctx, int mode) {
    if (!ctx) return SK_E_ARG;
    if (mode != SK_RESAMPLE_BIT_EXACT && mode != SK_RESAMPLE_FAST) SK_FAIL(ctx, SK_E_ARG, "unknown resample mode %d", mode);
    ctx->resample_fast = mode == SK_RESAMPLE_FAST;
    return SK_OK;
}

int sk_ensure_stats(sk_ctx* ctx) {
    if (ctx->d_stats) return SK_OK;
    SK_CUDA(ctx, cudaMalloc((void**)&ctx->d_stats, 8 * sizeof(unsigned long long)));
    SK_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
    return SK_OK;
}

extern "C" int sk_ctx_resample_fallback_outputs(sk_ctx* ctx, uint64_t* count) {
    if (!ctx || !count) return SK_E_ARG;
    *count = 0;
    if (!ctx->d_stats) return SK_OK;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long v = 0;
    SK_CUDA(ctx, cudaMemcpyAsync(&v, ctx->d_stats, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    SK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count = v;
    return SK_OK;
}

// ---- staging -------------------------------------------------------------------------------
int sk_ensure_host(sk_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_cap) return SK_OK;
    if (ctx->h_stage) { cudaFreeHost(ctx->h_stage); ctx->h_stage = nullptr; ctx->h_cap = 0; }
    size_t cap = sk_round_up(bytes, 1 << 20);
    if (cudaMallocHost(&ctx->h_stage, cap) != cudaSuccess) SK_FAIL(ctx, SK_E_ALLOC, "pinned alloc of %zu bytes failed", cap);
    ctx->h_cap = cap;
    return SK_OK;
}
static int ensure_dev(sk_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return SK_OK;
    if (*p) { cudaStreamSynchronize(ctx->stream); cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t ncap = sk_round_up(bytes + 256, 1 << 20);
    if (cudaMalloc(p, ncap) != cudaSuccess) SK_FAIL(ctx, SK_E_ALLOC, "device alloc of %zu bytes failed", ncap);
    *cap = ncap;
    return SK_OK;
}
int sk_ensure_dev_in(sk_ctx* ctx, size_t bytes) { return ensure_dev(ctx, &ctx->d_in, &ctx->d_in_cap, bytes); }
int sk_ensure_dev_out(sk_ctx* ctx, size_t bytes) { return ensure_dev(ctx, &ctx->d_out, &ctx->d_out_cap, bytes); }

// Host buffers of the drop-in API are pageable (Rust Vec / slices): stage through pinned memory in
// 8 MiB pieces so the DMA engine runs at full PCIe rate and the pinned footprint stays bounded.
static const size_t kStagePiece = kSkStageHalf;
// the two slot events live in the context (created once): a small call costs no event create / destroy
static int stage_events(sk_ctx* ctx) {
    for (int i = 0; i < 2; i++)
        if (!ctx->stage_ev[i]) SK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    return SK_OK;
}
int sk_h2d(sk_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (!bytes) return SK_OK;
    int rc = sk_ensure_host(ctx, 2 * kStagePiece);
    if (rc) return rc;
    if ((rc = stage_events(ctx))) return rc;
    char* stage = (char*)ctx->h_stage;
    cudaEvent_t* ev = ctx->stage_ev;
    size_t off = 0; int slot = 0; bool used[2] = {false, false};
    cudaError_t e = cudaSuccess;
    while (off < bytes && e == cudaSuccess) {
        size_t n = bytes - off < kStagePiece ? bytes - off : kStagePiece;
        if (used[slot]) e = cudaEventSynchronize(ev[slot]);
        if (e != cudaSuccess) break;
        memcpy(stage + slot * kStagePiece, (const char*)src_host + off, n);
        e = cudaMemcpyAsync((char*)dst_dev + off, stage + slot * kStagePiece, n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[slot], ctx->stream);
        used[slot] = true;
        off += n; slot ^= 1;
    }
    const cudaError_t es = cudaStreamSynchronize(ctx->stream);  // the stage must be idle when this returns, error or not
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) SK_FAIL(ctx, SK_E_CUDA, "h2d copy failed: %s", cudaGetErrorString(e));
    return SK_OK;
}
int sk_d2h(sk_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (!bytes) return SK_OK;
    int rc = sk_ensure_host(ctx, 2 * kStagePiece);
    if (rc) return rc;
    if ((rc = stage_events(ctx))) return rc;
    char* stage = (char*)ctx->h_stage;
    cudaEvent_t* ev = ctx->stage_ev;
    size_t off = 0; int slot = 0;
    size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    while ((off < bytes || pend_n[0] || pend_n[1]) && e == cudaSuccess) {
        if (pend_n[slot]) {
            e = cudaEventSynchronize(ev[slot]);
            if (e != cudaSuccess) break;
            memcpy((char*)dst_host + pend_off[slot], stage + slot * kStagePiece, pend_n[slot]);
            pend_n[slot] = 0;
        }
        if (off < bytes) {
            size_t n = bytes - off < kStagePiece ? bytes - off : kStagePiece;
            e = cudaMemcpyAsync(stage + slot * kStagePiece, (const char*)src_dev + off, n, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaEventRecord(ev[slot], ctx->stream);
            pend_off[slot] = off; pend_n[slot] = n;
            off += n;
        }
        slot ^= 1;
    }
    const cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) SK_FAIL(ctx, SK_E_CUDA, "d2h copy failed: %s", cudaGetErrorString(e));
    return SK_OK;
}

char* sk_stage_half(sk_ctx* ctx, int half) {
    if (sk_ensure_host(ctx, 2 * kStagePiece) != SK_OK) return nullptr;
    return (char*)ctx->h_stage + (half ? kStagePiece : 0);
}
bool sk_small_begin(sk_ctx* ctx, const void* in_host, size_t in_bytes, size_t out_bytes, void** in_dev, void** out_dev) {
    if (in_bytes > kSkSmallCallBytes || out_bytes > 4 * kSkSmallCallBytes) return false;
    if (sk_ensure_host(ctx, 2 * kStagePiece) != SK_OK) return false;  // (the large path reports the allocation failure)
    char* stage = (char*)ctx->h_stage;
    memcpy(stage, in_host, in_bytes);
    *in_dev = stage;
    *out_dev = stage + kStagePiece;
    return true;
}
int sk_small_end(sk_ctx* ctx, void* out_host, const void* out_dev, size_t out_bytes) {
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) SK_FAIL(ctx, SK_E_CUDA, "small call failed: %s", cudaGetErrorString(e));
    memcpy(out_host, out_dev, out_bytes);
    return SK_OK;
}

// ---- FFMA peak microbenchmark ----------------------------------------------------------------
// 16 independent accumulator chains per thread, operands in registers; the measured rate is the
// denominator of the resampler's FP32 roofline (BASELINE.md section 2 asks for it "in the same run").
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* sink, float x, float y, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmaf_rn(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123.456f) sink[0] = s;  // never true; keeps the chains alive
}

// same with the packed sm_100 FFMA2 (two fused multiply-adds per instruction)
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float* sink, float x, float y, int iters) {
    unsigned long long a[16], xx, yy;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
#pragma unroll
    for (int i = 0; i < 16; i++) { float f = (float)(threadIdx.x + i); asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(f)); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(xx), "l"(yy));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= a[i];
    if (s == 0x123456789abcdefull) sink[0] = 1.0f;
}

static int measure_peak(sk_ctx* ctx, int packed, double* tflops) {
    const int iters = 4096, blocks = ctx->num_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    SK_CUDA(ctx, cudaEventCreate(&e0));
    SK_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, ctx->stream);
        if (packed) ffma2_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((float*)ctx->d_out, 0.999f, 0.001f, iters);
        else ffma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((float*)ctx->d_out, 0.999f, 0.001f, iters);
        cudaEventRecord(e1, ctx->stream);
        SK_LAUNCH_CHECK(ctx);
        SK_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * (packed ? 2.0 : 1.0) * (double)blocks * threads * (double)iters * 8.0 * 16.0;
        double t = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return SK_OK;
}

extern "C" int sk_ctx_measure_ffma_peaks(sk_ctx* ctx, double* scalar_tflops, double* packed_tflops) {
    if (!ctx) return SK_E_ARG;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = sk_ensure_dev_out(ctx, 256);
    if (rc) return rc;
    double a = 0, b = 0;
    if ((rc = measure_peak(ctx, 0, &a))) return rc;
    if ((rc = measure_peak(ctx, 1, &b))) return rc;
    if (scalar_tflops) *scalar_tflops = a;
    if (packed_tflops) *packed_tflops = b;
    return SK_OK;
}

// FP32 FMA peak of this device: the better of the scalar (FFMA) and packed (FFMA2) microbenchmarks
extern "C" int sk_ctx_measure_ffma_peak(sk_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return SK_E_ARG;
    double a = 0, b = 0;
    int rc = sk_ctx_measure_ffma_peaks(ctx, &a, &b);
    if (rc) return rc;
    *tflops = a > b ? a : b;
    return SK_OK;
}

// ---- diagnostic: poison shared memory -----------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) poison_smem_kernel(uint32_t pattern, int words, unsigned* sink) {
    extern __shared__ uint32_t poison_sm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) poison_sm[i] = pattern;
    __syncthreads();
    if (threadIdx.x == 0 && poison_sm[words - 1] != pattern) *sink = 1;  // keeps the stores alive
}

extern "C" int sk_ctx_poison_shared_memory(sk_ctx* ctx, uint32_t pattern) {
    if (!ctx) return SK_E_ARG;
    SK_CUDA(ctx, cudaSetDevice(ctx->device));
    const int bytes = 226 * 1024;  // one CTA owns (nearly) the whole shared memory of its SM, so the first wave covers every SM
    SK_CUDA(ctx, cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    int rc = sk_ensure_dev_out(ctx, 256);
    if (rc) return rc;
    poison_smem_kernel<<<ctx->num_sms * 2, 1024, bytes, ctx->stream>>>(pattern, bytes / 4, (unsigned*)ctx->d_out);
    SK_LAUNCH_CHECK(ctx);
    return SK_OK;
}

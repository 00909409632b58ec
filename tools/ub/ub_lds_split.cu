// tools/ub/ub_lds_split.cu — (1) shared-memory wavefront cost of broadcast / partially-broadcast LDS.32/.64/.128, (2) the FIR walk of
// the general-ratio resampler with the 8 accumulator lanes of an output split over a lane PAIR (LDS.128, 4 columns per thread)
// against a lane QUAD (LDS.64, 8 columns per thread): 128 accumulator registers either way.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ub_lds_split ub_lds_split.cu && ./ub_lds_split
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- (1) LDS patterns: every thread issues `iters` x 8 independent loads of width W bytes at address base + f(lane)
template <int W, int PAT>
__global__ void __launch_bounds__(256, 1) lds_k(float* sink, int iters) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int q = threadIdx.x & 31;
    int off;  // word offset of this lane
    if (PAT == 0) off = 0;                              // all lanes the same address
    else if (PAT == 1) off = (q >> 4) * (W / 4);        // one address per half-warp, contiguous
    else if (PAT == 2) off = (q >> 3) * (W / 4);        // one address per quarter-warp, contiguous
    else off = q * (W / 4);                             // all distinct, contiguous
    uint32_t a = smem_u32(sm + off + (threadIdx.x >> 5) * 512);
    unsigned long long s0 = 0, s1 = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (W == 4) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a + u * 1024)); s0 += v; }
            if (W == 8) { unsigned long long v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a + u * 1024)); s0 ^= v; }
            if (W == 16) { unsigned long long v, w; asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(w) : "r"(a + u * 1024)); s0 ^= v; s1 ^= w; }
        }
    }
    if ((s0 ^ s1) == 0x1234567ull) sink[0] = 1.0f;
}

template <int W, int PAT>
static void run_lds(const char* name) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(lds_k<W, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        lds_k<W, PAT><<<sms, 256, 65536>>>(sink, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * khz * 1e3;
    printf("LDS.%-3d %-34s %.2f SM cycles per warp instruction\n", W * 8, name, cyc / ((double)iters * 8 * 8));
    cudaFree(sink);
}

// ---- (2) the walk.  SPLIT = 2: lane pair, 4 columns x 4 outputs x 2 products x 4 virtual lanes per thread (LDS.128);
//                     SPLIT = 4: lane quad, 8 columns x 4 outputs x 2 products x 2 virtual lanes per thread (LDS.64).
constexpr int MLEN = 264, ROWS = 8, TS = MLEN, CSTRIDE = 548;  // CSTRIDE / 4 odd
template <int SPLIT, int LOADS, int UNROLL>
__global__ void __launch_bounds__(256, 1) walk_k(float* sink, int groups) {
    extern __shared__ __align__(16) float sm[];
    float* taps = sm;                     // [8 warps][8 rows][TS]
    float* win = sm + 8 * ROWS * TS;      // [64 columns][CSTRIDE]
    for (int i = threadIdx.x; i < 8 * ROWS * TS + 64 * CSTRIDE; i += blockDim.x) sm[i] = 1.0f / (float)(1 + (i & 1023));
    __syncthreads();
    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    const float* T = taps + warp * ROWS * TS;
    unsigned long long total = 0;
    if (SPLIT == 2) {
        const int ql = q & 15, hh = q >> 4;
        const float* xb[4];
        for (int b = 0; b < 4; b++) xb[b] = win + (ql + 16 * (b ^ (2 * hh))) * CSTRIDE + 4 * hh;
        const float* Th = T + 4 * hh;
        for (int g = 0; g < groups; g++) {
            unsigned long long acc[4][4][2][2];
#pragma unroll
            for (int i = 0; i < 64; i++) (&acc[0][0][0][0])[i] = 0ull;
            const int xo = (g & 7) * 4;
            unsigned long long xp[4][2], tt[8][2];
#pragma unroll UNROLL
            for (int m = 0; m < MLEN; m += 8) {
#pragma unroll
                for (int b = 0; b < 4; b++) if (LOADS == 2 || m == 0) asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(xp[b][0]), "=l"(xp[b][1]) : "r"(smem_u32(xb[b] + xo + m)));
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    if (LOADS >= 1 || m == 0) asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(tt[r][0]), "=l"(tt[r][1]) : "r"(smem_u32(Th + r * TS + m)));
                    const unsigned long long t0 = tt[r][0], t1 = tt[r][1];
#pragma unroll
                    for (int b = 0; b < 4; b++) { ffma2(acc[r >> 1][b][r & 1][0], xp[b][0], t0); ffma2(acc[r >> 1][b][r & 1][1], xp[b][1], t1); }
                }
            }
#pragma unroll
            for (int i = 0; i < 64; i++) total ^= (&acc[0][0][0][0])[i];
        }
    } else {
        const int ql = q & 7, hq = q >> 3;
        const float* xb[8];
        for (int b = 0; b < 8; b++) xb[b] = win + (ql + 8 * (b ^ (2 * hq))) * CSTRIDE + 2 * hq;
        const float* Th = T + 2 * hq;
        for (int g = 0; g < groups; g++) {
            unsigned long long acc[4][8][2];
#pragma unroll
            for (int i = 0; i < 64; i++) (&acc[0][0][0])[i] = 0ull;
            const int xo = (g & 7) * 4;
            unsigned long long xp[8], tt[8];
#pragma unroll UNROLL
            for (int m = 0; m < MLEN; m += 8) {
#pragma unroll
                for (int b = 0; b < 8; b++) if (LOADS == 2 || m == 0) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(xp[b]) : "r"(smem_u32(xb[b] + xo + m)));
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    if (LOADS >= 1 || m == 0) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(tt[r]) : "r"(smem_u32(Th + r * TS + m)));
                    const unsigned long long t0 = tt[r];
#pragma unroll
                    for (int b = 0; b < 8; b++) ffma2(acc[r >> 1][b][r & 1], xp[b], t0);
                }
            }
#pragma unroll
            for (int i = 0; i < 64; i++) total ^= (&acc[0][0][0])[i];
        }
    }
    if (total == 0x1234567ull) sink[0] = 1.0f;
}

template <int SPLIT, int LOADS, int UNROLL = 33>
static void run_walk(int threads = 256) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* sink; cudaMalloc(&sink, 4);
    const int smem = (8 * ROWS * TS + 64 * CSTRIDE) * 4;
    cudaFuncSetAttribute(walk_k<SPLIT, LOADS, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int groups = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        walk_k<SPLIT, LOADS, UNROLL><<<sms, threads, smem>>>(sink, groups);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // per thread and group: (MLEN / 8) blocks x 64 FFMA2 x 4 flop
    const double flops = (double)groups * (MLEN / 8) * 64 * 4 * threads * sms;
    printf("walk, unroll %2d blocks, %d warps/SM, accumulator lanes split over %d threads, loads in the loop: %s: %.2f TFLOP/s executed (%.3f ms)\n", UNROLL, threads / 32, SPLIT, LOADS == 2 ? "x + taps" : LOADS == 1 ? "taps only" : "none", flops / (best * 1e-3) / 1e12, best);
    cudaFree(sink);
}

// ---- (3) 16 warps per SM, quad split, 4 tap rows (2 outputs x 2 products) x 8 columns per thread: 32 accumulator pairs, <= 128 registers
template <int UNROLL>
__global__ void __launch_bounds__(512, 1) walk16_k(float* sink, int groups) {
    extern __shared__ __align__(16) float sm[];
    float* taps = sm;                       // [16 warps][4 rows][TS]
    float* win = sm + 16 * 4 * TS;          // [64 columns][CSTRIDE]
    for (int i = threadIdx.x; i < 16 * 4 * TS + 64 * CSTRIDE; i += blockDim.x) sm[i] = 1.0f / (float)(1 + (i & 1023));
    __syncthreads();
    const int warp = threadIdx.x >> 5, q = threadIdx.x & 31;
    const float* T = taps + warp * 4 * TS;
    unsigned long long total = 0;
    const int ql = q & 7, hq = q >> 3;
    const float* xb[8];
    for (int b = 0; b < 8; b++) xb[b] = win + (ql + 8 * (b ^ (2 * hq))) * CSTRIDE + 2 * hq;
    const float* Th = T + 2 * hq;
    for (int g = 0; g < groups; g++) {
        unsigned long long acc[4][8];
#pragma unroll
        for (int i = 0; i < 32; i++) (&acc[0][0])[i] = 0ull;
        const int xo = (g & 7) * 4;
#pragma unroll UNROLL
        for (int m = 0; m < MLEN; m += 8) {
            unsigned long long xp[8];
#pragma unroll
            for (int b = 0; b < 8; b++) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(xp[b]) : "r"(smem_u32(xb[b] + xo + m)));
#pragma unroll
            for (int r = 0; r < 4; r++) {
                unsigned long long t0;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t0) : "r"(smem_u32(Th + r * TS + m)));
#pragma unroll
                for (int b = 0; b < 8; b++) ffma2(acc[r][b], xp[b], t0);
            }
        }
#pragma unroll
        for (int i = 0; i < 32; i++) total ^= (&acc[0][0])[i];
    }
    if (total == 0x1234567ull) sink[0] = 1.0f;
}
template <int UNROLL>
static void run_walk16() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* sink; cudaMalloc(&sink, 4);
    const int smem = (16 * 4 * TS + 64 * CSTRIDE) * 4;
    cudaFuncSetAttribute(walk16_k<UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int groups = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        walk16_k<UNROLL><<<sms, 512, smem>>>(sink, groups);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double flops = (double)groups * (MLEN / 8) * 32 * 4 * 512 * sms;
    printf("walk, unroll %2d blocks, 16 warps/SM, quad split, 4 rows x 8 columns per thread: %.2f TFLOP/s executed (%.3f ms)\n", UNROLL, flops / (best * 1e-3) / 1e12, best);
    cudaFree(sink);
}

int main() {
    run_lds<4, 0>("all lanes one address");
    run_lds<4, 3>("32 distinct words");
    run_lds<8, 0>("all lanes one address");
    run_lds<8, 1>("one address per half-warp");
    run_lds<8, 2>("one address per quarter-warp");
    run_lds<8, 3>("32 distinct (256 B)");
    run_lds<16, 0>("all lanes one address");
    run_lds<16, 1>("one address per half-warp");
    run_lds<16, 2>("one address per quarter-warp");
    run_lds<16, 3>("32 distinct (512 B)");
    run_walk<2, 2, 33>(); run_walk<2, 2, 11>(); run_walk<2, 2, 3>(); run_walk<2, 2, 1>();
    run_walk<4, 2, 33>(); run_walk<4, 2, 11>(); run_walk<4, 2, 3>(); run_walk<4, 2, 1>();
    run_walk<4, 0, 33>(); run_walk<4, 0, 3>();
    run_walk<2, 2, 3>(128); run_walk<4, 2, 3>(128); run_walk<4, 0, 3>(128);
    run_walk16<33>(); run_walk16<11>(); run_walk16<3>();
    return 0;
}

// tools/ub/ub_ffma2_warps.cu — how many warps per SM sub-partition does it take to keep the FP32 pipe busy with packed FFMA2 whose
// three operands are all distinct register pairs (the operand pattern of a register-tiled FIR: acc += x * tap)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ub_ffma2_warps ub_ffma2_warps.cu && ./ub_ffma2_warps
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024, 1) k(float* sink, int iters, float seed) {
    unsigned long long acc[NACC], x[4], t[8];
    for (int i = 0; i < NACC; i++) acc[i] = 0ull;
    for (int i = 0; i < 4; i++) { float f = seed + i + threadIdx.x; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(f)); }
    for (int i = 0; i < 8; i++) { float f = seed * 0.5f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(t[i]) : "f"(f)); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < NACC; a++) ffma2(acc[a], x[a & 3], t[(a >> 2) & 7]);
    }
    unsigned long long s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    if (s == 0x1234567ull) sink[0] = 1.0f;
}

template <int NACC>
static void run(int warps_per_sm) {
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* sink; cudaMalloc(&sink, 4);
    const int iters = 20000, threads = warps_per_sm * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<NACC><<<sms, threads>>>(sink, iters, 1.0f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double flops = 4.0 * NACC * (double)iters * threads * sms;
    printf("acc pairs %3d  warps/SM %2d (%.1f per scheduler): %.2f TFLOP/s\n", NACC, warps_per_sm, warps_per_sm / 4.0, flops / (best * 1e-3) / 1e12);
    cudaFree(sink);
}

int main() {
    for (int w : {4, 8, 12, 16, 32}) run<64>(w);
    for (int w : {4, 8, 16}) run<16>(w);
    return 0;
}

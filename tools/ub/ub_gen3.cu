// Microbenchmark for a constant-bank-tap general-ratio FIR inner loop (sm_100a).
// Question it answers: can LDCU.128 feed 2T rows of warp-uniform taps fast enough that an FFMA2 loop with one LDS.128 of
// inputs per 2 tap pairs runs near the FP32 dispatch bound, at the 6 warps/SM the per-window shared-memory layout allows?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ub_gen3 ub_gen3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NP = 132;          // tap pairs per tile (256 taps + ragged ends)
constexpr int PITCH = 292;
constexpr int KB = 66;           // LDS.128 per loop trip (65 = 13 x 5)       // words per window, = 4 mod 32: LDS.128 conflict-free across a quarter warp
__constant__ float4 c_taps[4096];   // 64 KiB

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// T outputs per thread (2T rows). MODE 0: taps from the constant bank; MODE 1: taps from shared memory (broadcast LDS.128)
template <int T, int MODE>
__global__ void __launch_bounds__(192, 1) fir_kernel(const float* __restrict__ gx, float* __restrict__ out, int ntiles, int reps) {
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float* xw = sm + (warp * 32 + lane) * PITCH;
    float4* staps = sm4 + nwarp * 32 * PITCH / 4;   // MODE 1: [NP][2T/2] float4 for one tile
    constexpr int R = 2 * T;
    float total = 0.f;
    for (int rep = 0; rep < reps; rep++) {
        // stage 32 windows (lane = k inside a window)
        for (int q = 0; q < 32; q++)
            for (int k = lane; k < PITCH; k += 32) sm[(warp * 32 + q) * PITCH + k] = gx[(size_t)((blockIdx.x * nwarp + warp) * 32 + q) * 64 + k + rep];
        __syncwarp();
        for (int tile = 0; tile < ntiles; tile++) {
            if (MODE == 1) {
                __syncthreads();
                for (int i = threadIdx.x; i < NP * R / 2; i += blockDim.x) staps[i] = c_taps[tile * NP * R / 2 + i];
                __syncthreads();
            }
            unsigned long long acc[R][4];
            const unsigned xs = (unsigned)__cvta_generic_to_shared(xw);
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int p = 0; p < 4; p++) acc[r][p] = 0ull;
            const float4* tp = MODE == 0 ? c_taps + tile * (NP * R / 2) : staps;
#pragma unroll 1
            for (int kb = 0; kb < NP / 2; kb += KB)
#pragma unroll
            for (int kq = kb; kq < kb + KB; kq++) {
                unsigned long long x0, x1;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x0), "=l"(x1) : "r"(xs + 16 * kq));
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int kp = 2 * kq + h;
#pragma unroll
                    for (int r2 = 0; r2 < R / 2; r2++) {
                        unsigned long long t0, t1;
                        if (MODE == 0) {
                            asm("ld.const.v2.b64 {%0, %1}, [%2];" : "=l"(t0), "=l"(t1) : "l"(tp + kp * (R / 2) + r2));
                        } else {
                            const unsigned long long* tq = reinterpret_cast<const unsigned long long*>(tp + kp * (R / 2) + r2);
                            t0 = tq[0]; t1 = tq[1];
                        }
                        ffma2(acc[2 * r2][kp & 3], h ? x1 : x0, t0);
                        ffma2(acc[2 * r2 + 1][kp & 3], h ? x1 : x0, t1);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                float l[8];
#pragma unroll
                for (int p = 0; p < 4; p++) { l[2 * p] = __uint_as_float((unsigned)acc[r][p]); l[2 * p + 1] = __uint_as_float((unsigned)(acc[r][p] >> 32)); }
                total += ((l[0] + l[4]) + (l[1] + l[5])) + ((l[2] + l[6]) + (l[3] + l[7]));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

template <int T, int MODE>
void run(const char* name, int warps, int ctas_per_sm, const float* gx, float* out) {
    const int R = 2 * T;
    const int ntiles = 4096 / (NP * R / 2);
    const int reps = 8;
    size_t smem = (size_t)warps * 32 * PITCH * 4 + (MODE == 1 ? NP * R / 2 * 16 : 0);
    CK(cudaFuncSetAttribute(fir_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fir_kernel<T, MODE>, warps * 32, smem));
    const int grid = 148 * ctas_per_sm * 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) fir_kernel<T, MODE><<<grid, warps * 32, smem>>>(gx, out, ntiles, reps);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int L = 5;
    for (int i = 0; i < L; i++) fir_kernel<T, MODE><<<grid, warps * 32, smem>>>(gx, out, ntiles, reps);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= L;
    const double fma = (double)grid * warps * 32 * reps * ntiles * NP * R * 2;
    printf("%-28s warps/CTA %d occ %d CTA/SM smem %zu  ntiles %d  %.3f ms  %.2f TFLOP/s  (%.3f of 74.4)\n", name, warps, occ, smem, ntiles, ms, fma * 2 / ms / 1e9, fma * 2 / ms / 1e9 / 74.4);
}

int main() {
    float *gx, *out;
    const size_t nx = (size_t)148 * 16 * 8 * 32 * 64 + 4096;
    CK(cudaMalloc(&gx, nx * 4)); CK(cudaMalloc(&out, 148 * 64 * 256 * 4));
    std::vector<float> h(nx); for (size_t i = 0; i < nx; i++) h[i] = (float)((i * 2654435761u) & 0xffff) / 65536.f - 0.5f;
    CK(cudaMemcpy(gx, h.data(), nx * 4, cudaMemcpyHostToDevice));
    std::vector<float> t(16384); for (int i = 0; i < 16384; i++) t[i] = (float)((i * 40503u) & 0xfff) / 4096.f - 0.5f;
    CK(cudaMemcpyToSymbol(c_taps, t.data(), 65536));
    run<4, 0>("T=4 const taps", 6, 1, gx, out);
    run<4, 0>("T=4 const taps", 3, 2, gx, out);
    run<4, 0>("T=4 const taps", 2, 3, gx, out);
    run<2, 0>("T=2 const taps", 6, 1, gx, out);
    run<2, 0>("T=2 const taps", 2, 3, gx, out);
    run<3, 0>("T=3 const taps", 6, 1, gx, out);
    run<4, 1>("T=4 smem taps", 6, 1, gx, out);
    run<2, 1>("T=2 smem taps", 6, 1, gx, out);
    return 0;
}

// sk_resample_int3_w2.cu — the persistent TMA kernels of t_ratio = 3 for 2-byte samples (one translation unit per sample
// width: these are the longest compiles of the build; see launch_int in sk_resample_dev.cuh)
#include "sk_resample_dev.cuh"

int sk_launch_int3_tma_w2(sk_ctx* ctx, const Taps& taps, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* d_table, size_t n_streams) {
    return launch_int_tma<3, 2>(ctx, taps, src, dst, plan, d_table, n_streams);
}

// sk_resample_int3.cu — instantiation of the integer-ratio resampler kernels for t_ratio = 3 (see sk_resample_dev.cuh)
#include "sk_resample_dev.cuh"

int sk_launch_int_3(sk_ctx* ctx, const SrcDesc& src, const OutDesc& dst, const sk_plan* plan, const float* h_table, const float* d_table,
                     size_t n_streams, int lanes_per_stream) {
    return launch_int<3>(ctx, src, dst, plan, h_table, d_table, n_streams, lanes_per_stream);
}

#ifdef SK_TRACE_CTAS
extern "C" int sk_debug_cta_trace(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, sk_cta_trace, sizeof(sk_cta_trace)); }
#endif

#!/bin/bash
# tools/build_variant.sh <name> [nvcc defines...] — builds soundkit_b200/variants/libsk_<name>.so with extra -D flags for
# A/B runs (select at run time with SOUNDKIT_CUDA_LIB=<path>).  Development aid; the product is csrc/Makefile's build.
set -e
cd "$(dirname "$0")/../soundkit_b200/csrc"
name=$1; shift
mkdir -p ../variants
nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 --threads 4 "$@" \
     -shared -o ../variants/libsk_${name}.so sk_ctx.cu sk_g711.cu sk_pcm.cu sk_resample.cu sk_resample_gen3.cu sk_resample_int2.cu sk_resample_int3.cu sk_resample_int3_w2.cu sk_resample_int3_w3.cu sk_resample_int3_w4.cu sk_resample_int4.cu sk_resample_int6.cu sk_resample_int12.cu sk_norm48k.cu sk_g726.cu sk_g722.cu -cudart static
echo built ../variants/libsk_${name}.so

set -x
mkdir -p gpurun_out/r02z
for v in n48o16 n48o8s48; do SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_$v.so timeout 120 python tools/kbench.py norm48k > gpurun_out/r02z/kbench_$v.json 2>&1; done
SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_n48o16.so timeout 300 python -m pytest tests -m gpu -q -k "48k or norm48" 2>&1 | tail -2
grep -A3 "normalizer" gpurun_out/r02z/kbench*.json | grep -v median

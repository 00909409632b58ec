set -x
mkdir -p gpurun_out/r02y
(time timeout 300 python -m pytest tests -m gpu -q -k "48k or norm48 or poison or unaligned or guard" 2>&1 | tail -4) > gpurun_out/r02y/pytest_n48.log 2>&1
timeout 120 python tools/kbench.py norm48k > gpurun_out/r02y/kbench_n48_o8.json 2>&1
SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_n48o4.so timeout 120 python tools/kbench.py norm48k > gpurun_out/r02y/kbench_n48_o4.json 2>&1
tail -3 gpurun_out/r02y/pytest_n48.log; grep -A3 "normalizer" gpurun_out/r02y/kbench*.json | grep -v median

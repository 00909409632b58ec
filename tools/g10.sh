set -x
mkdir -p gpurun_out/r02ad
(time timeout 600 python -m pytest tests -m gpu -q -k "g711 or G711 or latency or small" 2>&1 | tail -4) > gpurun_out/r02ad/pytest_g711.log 2>&1
timeout 200 python tools/kbench.py latency > gpurun_out/r02ad/kbench_latency.json 2>&1
timeout 300 python tools/kbench_int.py > gpurun_out/r02ad/kbench_int.json 2>&1
SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_int2r12.so timeout 300 python tools/kbench_int.py > gpurun_out/r02ad/kbench_int_int2r12.json 2>&1
SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_int2r12.so timeout 600 python -m pytest tests -m gpu -q -k "24000 or 22050 or resampl or chain or normalize" 2>&1 | tail -3 > gpurun_out/r02ad/pytest_int2r12.log
tail -3 gpurun_out/r02ad/pytest_g711.log; cat gpurun_out/r02ad/pytest_int2r12.log; grep -A2 "g711" gpurun_out/r02ad/kbench_latency.json; cat gpurun_out/r02ad/kbench_int.json gpurun_out/r02ad/kbench_int_int2r12.json

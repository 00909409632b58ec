set -x
mkdir -p gpurun_out/r02ae
(time timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/r02ae/pytest_gpu.log 2>&1
timeout 200 python tools/kbench.py latency > gpurun_out/r02ae/kbench_latency.json 2>&1
tail -3 gpurun_out/r02ae/pytest_gpu.log; grep -A2 "resampler_chunk\|160B" gpurun_out/r02ae/kbench_latency.json

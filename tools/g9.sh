set -x
mkdir -p gpurun_out/r02ac
(time timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r02ac/pytest_gpu.log 2>&1
timeout 200 python tools/kbench.py latency > gpurun_out/r02ac/kbench_latency.json 2>&1
tail -4 gpurun_out/r02ac/pytest_gpu.log; cat gpurun_out/r02ac/kbench_latency.json

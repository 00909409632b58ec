set -x
mkdir -p gpurun_out/r02aa
(time timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/r02aa/pytest_gpu.log 2>&1
timeout 120 python tools/kbench.py norm48k > gpurun_out/r02aa/kbench_n48.json 2>&1
tail -3 gpurun_out/r02aa/pytest_gpu.log; grep -A3 "normalizer" gpurun_out/r02aa/kbench*.json

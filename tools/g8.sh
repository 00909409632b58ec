set -x
mkdir -p gpurun_out/r02ab
(time timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r02ab/pytest_gpu.log 2>&1
timeout 120 python tools/kbench.py g722 g726 > gpurun_out/r02ab/kbench_g722.json 2>&1
tail -4 gpurun_out/r02ab/pytest_gpu.log; grep -A1 "g72" gpurun_out/r02ab/kbench_g722.json | grep -v median

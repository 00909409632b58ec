set -x
mkdir -p gpurun_out/r02v
(time python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02v/pytest_gpu.log 2>&1
bash tools/ab_kbench.sh "g726,bytes" g726u4 > gpurun_out/r02v/ab_g726_bytes.log 2>&1
python bench.py --extras bytes,g726 --no-cpu > gpurun_out/r02v/bench_bytes.json 2>&1
cat gpurun_out/r02v/pytest_gpu.log | tail -4; cat gpurun_out/r02v/ab_g726_bytes.log

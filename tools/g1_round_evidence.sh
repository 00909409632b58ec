set -x
mkdir -p gpurun_out/r02u
(time python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r02u/pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u/smoke.log 2>&1
python tools/kbench.py g726 > gpurun_out/r02u/kbench_g726.json 2>&1
python bench.py > gpurun_out/r02u/bench_default.json 2> gpurun_out/r02u/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u/bench_reference.json 2> gpurun_out/r02u/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02u/launches_bench_steps2_warmup3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r02u/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resample_int_tma_kernel -c 1 -s 3 -o gpurun_out/r02u/ncu_full_int3 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r02u/ncu_full.log 2>&1
ncu -i gpurun_out/r02u/ncu_full_int3.ncu-rep --page raw --csv > gpurun_out/r02u/ncu_full_int3.csv 2>/dev/null
KB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:resample_gen3_kernel -c 2 -s 1 -o gpurun_out/r02u/ncu_full_gen3 python tools/kbench_gen.py 64 > gpurun_out/r02u/ncu_full_gen3.log 2>&1
ncu -i gpurun_out/r02u/ncu_full_gen3.ncu-rep --page raw --csv > gpurun_out/r02u/ncu_full_gen3.csv 2>/dev/null
python tools/kbench_gen.py 64 > gpurun_out/r02u/kbench_gen3.json 2>&1
ncu --set full --clock-control none -k regex:g726 -c 2 -o gpurun_out/r02u/ncu_full_g726 python tools/ncu_targets.py > gpurun_out/r02u/ncu_full_g726.log 2>&1
ncu -i gpurun_out/r02u/ncu_full_g726.ncu-rep --page raw --csv > gpurun_out/r02u/ncu_full_g726.csv 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02u/smi.txt
tail -3 gpurun_out/r02u/pytest_gpu.log; cat gpurun_out/r02u/smoke.log | tail -2; head -c 1500 gpurun_out/r02u/bench_default.json

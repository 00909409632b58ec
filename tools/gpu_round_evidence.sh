# tools/gpu_round_evidence.sh [name] — on the GPU box: the driver's checks in order (pytest -m gpu -x, smoke, bench both arms) plus the ncu launch list of the
# bench command and one ncu --set full pass over every HBM-bound / codec kernel (tools/ncu_targets.py); outputs under gpurun_out/<name>/
set -x
D=gpurun_out/${1:-evidence}
mkdir -p $D
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $D/pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
python bench.py > $D/bench_default.json 2> $D/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $D/launches_bench_steps2_warmup3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > $D/ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:'g711|deint|flat_kernel|interleave_kernel|norm48k|g726|g722' -o $D/ncu_full_hbm_codec python tools/ncu_targets.py > $D/ncu_targets.log 2>&1
ncu -i $D/ncu_full_hbm_codec.ncu-rep --page raw --csv > $D/ncu_full_hbm_codec.csv 2>/dev/null
rm -f $D/ncu_full_hbm_codec.ncu-rep
tail -3 $D/pytest_gpu.log; tail -2 $D/smoke.log; head -c 600 $D/bench_default.json

set -x
mkdir -p gpurun_out/r02x
(time python -m pytest tests -m gpu -q -k "g726 or G726" 2>&1 | tail -4) > gpurun_out/r02x/pytest_g726.log 2>&1
python tools/kbench.py g726 > gpurun_out/r02x/kbench_g726_u4.json 2>&1
SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_g726e8.so python tools/kbench.py g726 > gpurun_out/r02x/kbench_g726_e8.json 2>&1
ncu --set full --clock-control none -k regex:norm48k -c 1 -o gpurun_out/r02x/ncu_full_norm48k python tools/ncu_targets.py > gpurun_out/r02x/ncu_norm48k.log 2>&1
ncu -i gpurun_out/r02x/ncu_full_norm48k.ncu-rep --page raw --csv > gpurun_out/r02x/ncu_full_norm48k.csv 2>/dev/null
tail -3 gpurun_out/r02x/pytest_g726.log; grep -A1 "g726" gpurun_out/r02x/kbench*.json | grep -v median

set -x
mkdir -p gpurun_out/r02w
(time python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02w/pytest_gpu.log 2>&1
python tools/kbench.py norm48k flat > gpurun_out/r02w/kbench_norm48k_flat.json 2>&1
SOUNDKIT_CUDA_LIB=$PWD/soundkit_b200/variants/libsk_g726u4.so python tools/kbench.py g726 > gpurun_out/r02w/kbench_g726_u4.json 2>&1
tail -4 gpurun_out/r02w/pytest_gpu.log; grep -A3 "normalizer\|s16_bytes\|g726" gpurun_out/r02w/*.json | grep -v median | head -40

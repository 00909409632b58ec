# tools/gpu_multi.sh <N> — on an N-GPU box: the driver's multi-GPU launch of bench.py (both arms), outputs under gpurun_out/multi_<N>/
set -x
N=${1:-8}
D=gpurun_out/multi_$N
mkdir -p $D
nvidia-smi topo -m > $D/topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $D/bench.json 2> $D/bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
tail -c 400 $D/bench.err; head -c 900 $D/bench.json; echo; head -c 400 $D/bench_reference.json

# usage: bash tools/multi_r02.sh N   (inside gpurun --gpus N)
N=$1
set -x
python -m pytest tests -m gpu -x -q -k "multi or cpp" 2>&1 | tail -3 > gpurun_out/gputest_multi_${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --no-sub-benchmarks > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 400 gpurun_out/bench_${N}gpu.err
cat gpurun_out/gputest_multi_${N}.log

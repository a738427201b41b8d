#!/bin/bash
# last multi-GPU session of round 2 (run under gpurun --gpus N): the bench line at N GPUs (configs[1] + the other
# configurations) and configuration 5 at its named size (16384 streams over 8 GPUs) with the split resampling pipeline
N=${N:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run $N 29543 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2b_bench_c2_${N}gpu.json 2> gpurun_out/r2b_bench_c2_${N}gpu.err
tail -c 400 gpurun_out/r2b_bench_c2_${N}gpu.json
run $N 29544 bench.py --gpus $N --steps 20 --warmup 3 --config c5 --no-other-configs > gpurun_out/r2b_bench_c5_${N}gpu.json 2> gpurun_out/r2b_bench_c5_${N}gpu.err
tail -c 400 gpurun_out/r2b_bench_c5_${N}gpu.json

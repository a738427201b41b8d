#!/bin/bash
# bench line at N GPUs of one box (run under gpurun --gpus N)
N=${N:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 3 \
  > gpurun_out/r2_bench_c2_${N}gpu.json 2> gpurun_out/r2_bench_c2_${N}gpu.err
tail -c 300 gpurun_out/r2_bench_c2_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus $N --steps 3 --warmup 1 \
  > gpurun_out/r2_bench_ref_${N}gpu.json 2> gpurun_out/r2_bench_ref_${N}gpu.err
tail -c 300 gpurun_out/r2_bench_ref_${N}gpu.json

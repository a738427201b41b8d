#!/bin/bash
# closing multi-GPU session of round 2 (run under gpurun --gpus N): configuration 5 at its named size (16384 streams over
# 8 GPUs) with the final block shape of k_resample_ls
N=${N:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 \
  --config c5 --no-other-configs > gpurun_out/r2c_bench_c5_${N}gpu.json 2> gpurun_out/r2c_bench_c5_${N}gpu.err
tail -c 600 gpurun_out/r2c_bench_c5_${N}gpu.json

#!/bin/bash
# multi-GPU session (run under gpurun --gpus N): copy ceilings at 1/2/4/N GPUs, the bench line at N GPUs (configs[1]) and
# configuration 5 at its named size (16384 streams over 8 GPUs)
N=${N:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 1 2 4 $N; do
  [ $n -gt $N ] && continue
  run $n 29541 tools/pcie_peak.py > gpurun_out/pcie_ceiling_${n}gpu.json 2> gpurun_out/pcie_ceiling_${n}gpu.err
  run $n 29542 tools/pcie_peak.py numa > gpurun_out/pcie_ceiling_${n}gpu_numa.json 2>> gpurun_out/pcie_ceiling_${n}gpu.err
done
run $N 29543 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_${N}gpu.json 2> gpurun_out/r2_bench_c2_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_c2_${N}gpu.json
run $N 29544 bench.py --gpus $N --steps 20 --warmup 3 --config c5 --no-other-configs > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_c5_${N}gpu.json
nproc; numactl -H 2>/dev/null | head -5; lscpu | grep -i "numa\|model name\|^CPU(s)" | head -8

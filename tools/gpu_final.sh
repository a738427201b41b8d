#!/bin/bash
# end-of-round GPU session (run under gpurun): parity tests, the two bench arms, the ncu launch list of the bench
# command and one full ncu capture of the dominant kernel.  TAG names the output files.
TAG=${TAG:-r1_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_$TAG.json
# (only our kernels: the input synthesis of bench.py launches hundreds of torch kernels first)
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(void )?k_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_stream --launch-skip 3 -c 1 -f \
  -o gpurun_out/prof_$TAG python bench.py --quick --steps 2 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
for cfg in c1 c3 c4 c5; do echo "== $cfg"; timeout 300 python bench.py --quick --config $cfg --steps 20 --warmup 3 2>&1 | tail -1; done

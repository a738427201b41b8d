#!/bin/bash
# one GPU session: parity tests, quick device-resident timings of the named configurations, optionally an ncu capture
# of one kernel (run under gpurun):  CFGS="c2 c1" STREAM_MODES="1 0" NCU=k_stream TAG=v2 bash tools/gpu_round.sh
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log; fi
for cfg in ${CFGS:-c2}; do
  for st in ${STREAM_MODES:-1 0}; do
    echo "== $cfg IAMFB_STREAM=$st"
    IAMFB_STREAM=$st timeout 300 python bench.py --quick --config $cfg --steps 20 --warmup 3 2>&1 | tail -1
  done
done
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$NCU --launch-skip 3 -c 1 -f \
    -o gpurun_out/prof_${NCU}_${TAG:-x} python bench.py --quick --config ${NCU_CFG:-c2} --steps 2 --warmup 3 > gpurun_out/ncu_${NCU}_${TAG:-x}.log 2>&1
  tail -2 gpurun_out/ncu_${NCU}_${TAG:-x}.log
fi

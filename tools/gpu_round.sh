#!/bin/bash
# one GPU session: parity tests, then quick device-resident timings of the named configurations (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for cfg in ${CFGS:-c2}; do
  for st in ${STREAM_MODES:-1 0}; do
    echo "== $cfg IAMFB_STREAM=$st"
    IAMFB_STREAM=$st timeout 300 python bench.py --quick --config $cfg --steps 20 --warmup 3 2>&1 | tail -1
  done
done

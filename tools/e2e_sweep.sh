#!/bin/bash
# sweep of the host-resident (e2e) path: stream groups per submit x frames per submit (run under gpurun)
for ch in ${CHUNKS:-4 8}; do for fr in ${FRAMES:-8 16}; do
  IAMFB_HOST_CHUNKS=$ch timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-frames $fr 2>/dev/null | CH=$ch FR=$fr python -c '
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("chunks",os.environ["CH"],"frames",os.environ["FR"],"e2e",round(d["e2e"]["value"]),"ms",round(d["e2e"]["ms_per_step"],2),"value",round(d["value"]))'
done; done

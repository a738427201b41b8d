#!/bin/bash
# resampling pipeline forms: parity + C5 timing per form (run under gpurun)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resampling_pipeline_forms or pipe_rs" 2>&1 | tail -15 | tee gpurun_out/pytest_rs.log; fi
for m in ${MODES:-0 1}; do
  for ck in ${CHUNKS:-auto}; do
    if [ "$ck" = auto ]; then unset IAMFB_LS_CHUNK; else export IAMFB_LS_CHUNK=$ck; fi
    echo "== IAMFB_RS_SPLIT=$m IAMFB_LS_CHUNK=$ck"
    IAMFB_RS_SPLIT=$m IAMFB_BENCH_KERNELS=1 timeout 300 python bench.py --quick --config c5 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'value': round(d.get('value')), 'ms_per_submit': d.get('ms_per_submit'), 'kernels': d.get('kernels')}))
" || true
  done
done

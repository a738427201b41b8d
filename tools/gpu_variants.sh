#!/bin/bash
# A/B of prebuilt variants of libiamf_b200.so (tmp_variants/*.so; run under gpurun): resampling parity tests + C5 timing per variant
mkdir -p gpurun_out
cp iac_b200/libiamf_b200.so /tmp/orig.so
for v in ${VARIANTS:-$(ls tmp_variants/*.so)}; do
  n=$(basename $v .so); echo "== $n"
  cp $v iac_b200/libiamf_b200.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resampling_pipeline_forms or pipe_rs" 2>&1 | tail -1
  for rep in 1 2; do
  IAMFB_BENCH_KERNELS=1 timeout 300 python bench.py --quick --config c5 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels') or {}
print(json.dumps({'value': round(d.get('value')), 'ms_per_submit': round(d.get('ms_per_submit'), 4), 'ls': k.get('k_resample_ls')}))
" || true
  done
done 2>&1 | tee gpurun_out/variants.log
cp /tmp/orig.so iac_b200/libiamf_b200.so

#!/bin/bash
# round-2 closing session (run under gpurun): GPU parity tests, both bench arms, a sweep of the resampler's work-item
# size, the full ncu capture of k_resample_ls in its final block shape
TAG=${TAG:-s5}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
for ck in 28 32 36 40 44 48 52 56; do
  echo "== IAMFB_LS_CHUNK=$ck"
  IAMFB_LS_CHUNK=$ck IAMFB_BENCH_KERNELS=1 timeout 300 python bench.py --quick --config c5 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get('kernels') or {}
print(json.dumps({'value': round(d.get('value')), 'ms_per_submit': round(d.get('ms_per_submit'), 4), 'ls': k.get('k_resample_ls')}))
" || true
done 2>&1 | tee gpurun_out/chunk_sweep_$TAG.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_resample_ls --launch-skip 4 -c 1 -f -o gpurun_out/r2_ncu_resample_ls_c5_t4 \
    python bench.py --quick --config c5 --steps 1 --warmup 3 --submits-per-step 2 > gpurun_out/r2_ncu_resample_ls_c5_t4.log 2>&1
tail -1 gpurun_out/r2_ncu_resample_ls_c5_t4.log | cut -c1-200

#!/bin/bash
# round-2, last session: full ncu captures of the split resampling pipeline's kernels (configuration 5) and the launch
# list of the bench command (run under gpurun)
mkdir -p gpurun_out
cap() {  # cap <tag> <kernel regex> <config>
  local tag=$1 k=$2 cfg=$3; shift 3
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 -c 1 -f -o gpurun_out/r2_ncu_$tag \
    python bench.py --quick --config $cfg --steps 1 --warmup 3 --submits-per-step 2 "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2_ncu_$tag.log
}
cap resample_ls_c5 k_resample_ls c5
cap pipe_rs_lim_c5 k_pipe_rs c5
cap prerender_c5 k_pipe_prerender c5
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(void )?k_' -c 700 --csv --log-file gpurun_out/r2_launches_final.csv \
  python bench.py --steps 1 --warmup 3 --submits-per-step 2 --no-cpu-baseline > gpurun_out/r2_launches_final.log 2>&1
tail -2 gpurun_out/r2_launches_final.log | cut -c1-300

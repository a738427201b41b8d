#!/bin/bash
# round-2 profile session (run under gpurun): ncu launch list of the bench command + one full capture per dominant kernel
mkdir -p gpurun_out
cap() {  # cap <tag> <kernel regex> <config> [extra bench args]
  local tag=$1 k=$2 cfg=$3; shift 3
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 -c 1 -f -o gpurun_out/r2_ncu_$tag \
    python bench.py --quick --config $cfg --steps 1 --warmup 3 --submits-per-step 2 "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2_ncu_$tag.log
}
cap hrtf_gemm_c4h k_hrtf_gemm c4h
cap hrtf_prep_c4h k_hrtf_prep c4h
cap pipe_c3 'k_pipe' c3
cap pipe_c4 'k_pipe' c4
cap stream_c2 'k_stream' c2
cap pipe_rs_c5 'k_pipe_rs' c5
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(void )?k_' -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 3 --submits-per-step 2 --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
tail -2 gpurun_out/r2_launches.log | cut -c1-300

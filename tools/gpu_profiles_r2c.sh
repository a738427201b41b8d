#!/bin/bash
# round-2 closing session: full ncu captures of k_pipe at the 16-frame submits bench.py now runs C3 / C4 with (run under gpurun)
mkdir -p gpurun_out
cap() {  # cap <tag> <kernel regex> <config>
  local tag=$1 k=$2 cfg=$3; shift 3
  timeout 120 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 -c 1 -f -o gpurun_out/r2_ncu_$tag \
    python bench.py --quick --config $cfg --steps 1 --warmup 3 --submits-per-step 2 "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  tail -1 gpurun_out/r2_ncu_$tag.log | cut -c1-160
}
cap pipe_c4_f16 k_pipe c4
cap pipe_c3_f16 k_pipe c3

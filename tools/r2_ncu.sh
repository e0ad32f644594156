#!/bin/bash
# ncu --set full captures of the trace kernel for one scene / schedule (run under gpurun, one GPU).
# usage: tools/r2_ncu.sh <tag> <scene: config2|soup1m|hf4m> <sched 1..4> <tri_threshold> [mode]
tag=$1; scene=$2; sched=$3; thr=$4; mode=${5:-closest}
mkdir -p gpurun_out
TRIRO_SCHED=$sched TRIRO_TRI_THRESHOLD=$thr timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/$tag python tools/profile_closest.py $scene 2 $mode > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log

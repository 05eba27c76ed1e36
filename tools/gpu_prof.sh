#!/bin/bash
# ncu evidence for one bench step: launch list (time shares) + one --set full capture of the dominant GEMM.
# usage: tools/gpu_prof.sh <tag> [kernel-regex] [skip] [count]
TAG=${1:-r01}; KREG=${2:-gemm_tn}; SKIP=${3:-179}; CNT=${4:-2}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KREG -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG}_$KREG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/plain_$TAG.log | cut -c1-600

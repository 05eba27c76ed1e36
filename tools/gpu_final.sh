#!/bin/bash
# end-of-round evidence: full gpu suite, bench line, per-kernel roofline table, ncu launch list of one step
TAG=${1:-final}
bash tools/gpu_quick.sh $TAG
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 420 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"

#!/bin/bash
# Round-2 ncu evidence (one GPU): launch list of one step, --set full captures of gemm_wgrad / layernorm_bwd / patch_embed /
# mlp_fwd launches, DRAM traffic of every gemm_tn launch of one step.  Each ncu run follows a plain run that exited 0.
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
# launches per step: count them from a first short list, then skip the 10 warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_wgrad_kernel" -s 270 -c 3 -f -o gpurun_out/prof_${TAG}_wgrad $CMD > gpurun_out/ncu_wgrad_$TAG.log 2>&1; echo "wgrad rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd_kernel" -s 150 -c 3 -f -o gpurun_out/prof_${TAG}_lnbwd $CMD > gpurun_out/ncu_lnbwd_$TAG.log 2>&1; echo "lnbwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"patch_embed_kernel|mlp_fwd_kernel" -s 50 -c 3 -f -o gpurun_out/prof_${TAG}_fused $CMD > gpurun_out/ncu_fused_$TAG.log 2>&1; echo "fused rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm_tn_kernel" -s 490 -c 60 --csv --log-file gpurun_out/gemm_tn_traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1; echo "traffic rc=$?"
tail -n 2 gpurun_out/plain_$TAG.log | cut -c1-300

#!/bin/bash
# Round-2 final evidence (one GPU): whole GPU suite, default bench line + per-kernel table, ncu launch list of one step, --set
# full captures of the short-row attention kernels and of the decoder attention backward, DRAM traffic of the gemm_tn launches.
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/t_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_all_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/roofline_table.py gpurun_out/bench_$TAG.log > gpurun_out/kernel_roofline_$TAG.md
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm_tn_kernel" -s 440 -c 100 --csv --log-file gpurun_out/gemm_tn_traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1; echo "traffic rc=$?"
timeout 120 python tools/attn_enc_one.py > gpurun_out/plain_enc_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_.*enc" -s 2 -c 2 -f -o gpurun_out/prof_${TAG}_attn_enc python tools/attn_enc_one.py > gpurun_out/ncu_enc_$TAG.log 2>&1; echo "enc capture rc=$?"
timeout 120 python tools/attn_one_bwd.py > gpurun_out/plain_tcb_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tc" -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_attn_bwd_tc python tools/attn_one_bwd.py > gpurun_out/ncu_tcb_$TAG.log 2>&1; echo "tc bwd capture rc=$?"
tail -n 1 gpurun_out/bench_$TAG.log | cut -c1-400

#!/bin/bash
# short-row tcgen05 attention: parity tests, then device timing against the mma.sync kernels
TAG=${1:-enc}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "attention" > gpurun_out/t_attn_$TAG.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/t_attn_$TAG.log
timeout 120 python tools/attn_time.py > gpurun_out/attn_time_$TAG.log 2>&1; echo "time rc=$?"; cat gpurun_out/attn_time_$TAG.log
MAE_ATTN_NO_TC=1 timeout 120 python tools/attn_time.py > gpurun_out/attn_time_${TAG}_notc.log 2>&1; cat gpurun_out/attn_time_${TAG}_notc.log

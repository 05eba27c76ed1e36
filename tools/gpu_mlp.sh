#!/bin/bash
# fused-MLP iteration loop: kernel test + micro timing of mae_mlp_fwd vs the three launches it replaces
TAG=${1:-m}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "mlp_fwd" 2>&1 | tail -3
timeout 300 python tools/mlp_time.py 2>&1 | tee gpurun_out/mlp_time_$TAG.log

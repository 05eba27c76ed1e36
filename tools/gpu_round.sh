#!/bin/bash
# One gpurun call: per-file GPU tests in separate processes (own timeouts), smoke, short bench.
mkdir -p gpurun_out; rm -f gpurun_out/metrics.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "token_mask or gather or layernorm or scatter or mse or casts" > gpurun_out/t_elem.log 2>&1; echo "elem rc=$?"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "gemm_tn" > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "wgrad" > gpurun_out/t_wgrad.log 2>&1; echo "wgrad rc=$?"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "attention" > gpurun_out/t_attn.log 2>&1; echo "attn rc=$?"
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 5 gpurun_out/t_*.log gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench.log

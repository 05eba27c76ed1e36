#!/bin/bash
# quick check: full gpu test-suite + bench summary
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/t_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/t_all_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.log").read().strip().splitlines()[-1])
print("img/s",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]))
print({k.replace("mae_",""):v["ms"] for k,v in d["kernel_share"].items() if v["ms"]>0.03})
print("gemm_tn GB/s",round(d["roofline"]["achieved"],1),{k:v for k,v in d["extra"].items() if k!="kernel_roofline"})
PY
python tools/roofline_table.py gpurun_out/bench_$TAG.log > gpurun_out/kernel_roofline_$TAG.md; head -30 gpurun_out/kernel_roofline_$TAG.md

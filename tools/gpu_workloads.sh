#!/bin/bash
# extra bench lines: BASELINE.json configs[3] (classifier) and configs[4] (ViT-B/16 stress) on one GPU
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py --workload classifier --steps 10 --warmup 5 > gpurun_out/bench_classifier_$TAG.log 2>&1; echo "classifier rc=$?"; tail -n 1 gpurun_out/bench_classifier_$TAG.log | cut -c1-1200
timeout 900 python bench.py --workload vitb16 --steps 10 --warmup 5 --no-cpu > gpurun_out/bench_vitb16_$TAG.log 2>&1; echo "vitb16 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_vitb16_$TAG.log").read().strip().splitlines()[-1])
print("vitb16 img/s",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]), "step_roofline", round(d["step_roofline"]["frac"],3), "gemm_tn hbm frac", round(d["roofline"]["frac"],3), "tensor", round(d["roofline"]["tensor"]["frac"],3))
print({k.replace("mae_",""):v["ms"] for k,v in d["kernel_share"].items() if v["ms"]>0.2})
PY
python tools/roofline_table.py gpurun_out/bench_vitb16_$TAG.log > gpurun_out/kernel_roofline_vitb16_$TAG.md

#!/bin/bash
# 1/2/4/8-GPU weak-scaling run on ONE box (gpurun --gpus 8): same procedure as the driver's SCALE run
TAG=${1:-r02}
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then CMD="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu";
  else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 20 --warmup 5"; fi
  $CMD > gpurun_out/scale_n${N}_$TAG.log 2>&1; echo "N=$N rc=$?"
done
python - <<PY
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_n%d_$TAG.log"%N).read().strip().splitlines()[-1])
    except Exception as e:
        print(N, "failed", e); continue
    if base is None: base=d["value"]
    print("N",N,"img/s",round(d["value"]),"ms",round(d["ms_per_step"],3),"eff",round(d["value"]/(N*base),4),"e2e",round(d["e2e"]["value"]),"e2e_u8",round(d["extra"]["e2e_uint8_gpu_augment"]["value"]),"replica_maxdiff",d["extra"]["replica_param_maxdiff"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/ddp_parity.py --batch 500 --steps 3 > gpurun_out/ddp_parity_n8_b500_$TAG.log 2>&1; echo "parity n8 rc=$?"; grep ddp_parity gpurun_out/ddp_parity_n8_b500_$TAG.log | cut -c1-800

#!/bin/bash
# N-GPU checks (run with gpurun --gpus N): N-rank parity vs one process, the CLI driver under torchrun, bench at N
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/ddp_parity.py --batch 500 --steps 3 > gpurun_out/ddp_parity_n${N}_b500_$TAG.log 2>&1; echo "parity b500 rc=$?"; grep ddp_parity gpurun_out/ddp_parity_n${N}_b500_$TAG.log | cut -c1-600
$TR --master-port 29512 tools/ddp_parity.py --batch 2000 --steps 3 > gpurun_out/ddp_parity_n${N}_b2000_$TAG.log 2>&1; echo "parity b2000 rc=$?"; grep ddp_parity gpurun_out/ddp_parity_n${N}_b2000_$TAG.log | cut -c1-600
export MAE_SYNTHETIC_DATA=1 MAE_SYNTHETIC_SIZE=2048
python - <<PY
import yaml
cfg = yaml.safe_load(open("configs/mae.yaml"))
cfg["pretrain"].update(batch_size=128, num_workers=0, total_epochs=2, warmup_epochs=1, val_split=0.125)
cfg["logging"]["output_dir_base"] = "gpurun_out/ddp_outputs"
yaml.safe_dump(cfg, open("gpurun_out/ddp_cfg.yaml", "w"))
PY
$TR --master-port 29513 -m scripts.training.pretrain_mae --config gpurun_out/ddp_cfg.yaml --output_dir_suffix ddp > gpurun_out/ddp_driver_n${N}_$TAG.log 2>&1; echo "driver rc=$?"; grep -E "trainer\]|complete|data\]" gpurun_out/ddp_driver_n${N}_$TAG.log | tail -6
ls gpurun_out/ddp_outputs/pretrain/ddp/checkpoints 2>/dev/null; rm -rf gpurun_out/ddp_outputs
$TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n${N}_$TAG.log").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "img/s", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "replica_maxdiff", d["extra"]["replica_param_maxdiff"])
PY

TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in p2p nccl; do
  MAE_GRAD_EXCHANGE=$mode $TR --master-port 29547 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_n2_$mode.log 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/bench_n2_$mode.log').read().strip().splitlines()[-1]); print('$mode', round(d['ms_per_step'],3), d['extra'].get('rank_ms_per_step'))"
done
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_n1_same.log 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_same.log').read().strip().splitlines()[-1]); print('n1', round(d['ms_per_step'],3))"

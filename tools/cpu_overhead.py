"""How long does the host take to ENQUEUE one fused step (no sync) vs the device time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssrl_vit_mae_jepa_b200 as P
import bench
mod = P.MAEPretrainModule(bench.CFG, dict(bench.PRETRAIN)).cuda()
tr = P.FusedTrainer(mod); tr.set_epoch(0)
x = torch.randn(2000, 3, 96, 96, device="cuda")
for _ in range(5): tr.train_step(x)
torch.cuda.synchronize()
for n in (1, 20):
    t0 = time.perf_counter()
    for _ in range(n): tr.train_step(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n}: enqueue {1e3*(t1-t0)/n:.2f} ms/step, total {1e3*(t2-t0)/n:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): tr.train_step(x)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

"""Experiment: forward + backward of one step captured in a CUDA graph (static images / noise) vs eager launches.
Tells how much of the step is launch gaps rather than kernel time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssrl_vit_mae_jepa_b200 as P
from bench import CFG, PRETRAIN
B = 2000
dev = torch.device("cuda", 0)
torch.manual_seed(73)
mod = P.MAEPretrainModule(CFG, dict(PRETRAIN, batch_size=B)).to(dev)
tr = P.FusedTrainer(mod); tr.set_epoch(0)
eng = tr.engine
images = torch.randn(B, 3, 96, 96, device=dev)
noise = torch.rand(B, 145, device=dev)
for _ in range(3): tr.train_step(images, noise)
torch.cuda.synchronize()
def fb():
    eng.forward(images, noise, 0.75)
    eng.backward()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager fwd+bwd ms:", round(timeit(fb), 3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(dev)
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    fb()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    fb()
torch.cuda.synchronize()
print("graph fwd+bwd ms:", round(timeit(g.replay), 3))

"""-m gpu: records the 'stock PyTorch on Blackwell' bar of BASELINE.md §5 -- the oracle restatement (identical ATen op
sequence to the reference) run eagerly under torch.autocast(bf16) with cuBLAS / SDPA / fused AdamW on the same
B200, B=2000, device-timed -- next to the same step through this framework.  Informational (written to
gpurun_out/metrics.log); the only assertion is that the framework is not slower than eager PyTorch."""
import pytest
import torch
import torch.nn.functional as F

import ssrl_vit_mae_jepa_b200 as P
from oracle import mae_oracle as O
from tests.gpu_util import report

pytestmark = pytest.mark.gpu
CFG = {"general": O.MAE_YAML["general"], "encoder": O.MAE_YAML["encoder"], "decoder": O.MAE_YAML["decoder"]}


def _time(fn, warm=3, iters=8):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_eager_pytorch_bar_vs_framework():
    B = 2000
    torch.manual_seed(0)
    images = torch.randn(B, 3, 96, 96, device="cuda")
    model = O.build_mae(seed=73).cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05, fused=True)

    def eager_step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred, target = model(images)
            loss = F.mse_loss(pred, target)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()

    ms_eager = _time(eager_step)
    del model, opt
    torch.cuda.empty_cache()
    mod = P.MAEPretrainModule(CFG, dict(O.MAE_YAML_PRETRAIN)).cuda()
    tr = P.FusedTrainer(mod)
    tr.set_epoch(0)
    ms_ours = _time(lambda: tr.train_step(images))
    report("eager_bar", eager_ms=ms_eager, eager_img_s=B / ms_eager * 1e3, ours_ms=ms_ours, ours_img_s=B / ms_ours * 1e3,
           speedup=ms_eager / ms_ours)
    assert ms_ours < ms_eager

"""N-rank data-parallel parity: `torchrun --nproc-per-node N tools/ddp_parity.py [--batch B] [--steps S]`.

Every rank trains its own shard with `FusedTrainer` (bucketed NCCL all-reduce(avg) overlapped with backward, then
clip + AdamW redundantly per rank).  Rank 0 also trains ONE non-distributed replica on the concatenated batch with the
ranks' noise concatenated in rank order.  Because the loss is a mean over equally sized shards, the averaged gradient
equals the single-process gradient (reference `src/training/mae.py:45-50`, `configs/mae.yaml:27`), so after S steps:
  * all ranks hold bit-identical parameters (max |diff| must be exactly 0.0),
  * they equal the single-process parameters up to the summation order of the fp32 gradient reductions
    (rel-L2 per tensor: <= 1e-5 for every weight / token / LayerNorm-scale tensor; <= 5e-3 for the zero-initialised biases,
    which after a few steps consist of nothing but the accumulated update; the zero-gradient K third of attn.qkv.bias excluded),
  * the mean of the per-rank losses equals the single-process loss (rel <= 1e-5).
Prints one JSON line on rank 0 and exits non-zero on failure."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=500, help="per-rank batch")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    import ssrl_vit_mae_jepa_b200 as P
    import yaml
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "mae.yaml")))
    mcfg, pcfg = cfg["model"], dict(cfg["pretrain"], batch_size=a.batch)

    def make(distributed):
        torch.manual_seed(73)
        mod = P.MAEPretrainModule(mcfg, pcfg).to("cuda")
        tr = P.FusedTrainer(mod, distributed=distributed)
        tr.set_epoch(0)
        return mod, tr

    mod, tr = make(True)
    if rank == 0:
        mod1, tr1 = make(False)
    g = torch.Generator(device="cuda").manual_seed(1234)
    losses, losses1 = [], []
    for s in range(a.steps):
        # the same global batch on every rank (same generator seed); each rank takes its slice
        images = torch.randn(world * a.batch, 3, 96, 96, device="cuda", generator=g)
        noise = torch.rand(world * a.batch, 145, device="cuda", generator=g)
        sl = slice(rank * a.batch, (rank + 1) * a.batch)
        l = tr.train_step(images[sl].contiguous(), noise[sl].contiguous())
        dist.all_reduce(l, op=dist.ReduceOp.AVG)
        losses.append(l.item())
        if rank == 0:
            losses1.append(tr1.train_step(images, noise).item())
    # replicas identical?
    flat = mod.model.engine.flat
    mx, mn = flat.clone(), flat.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    replica_maxdiff = (mx - mn).abs().max().item()
    ok = True
    if rank == 0:
        sd, sd1 = mod.model.state_dict(), mod1.model.state_dict()
        worst, worst_name, worst_b, worst_b_name = 0.0, "", 0.0, ""
        for n in sd:
            wa, wb = sd[n].double().flatten(), sd1[n].double().flatten()
            if n.endswith("attn.qkv.bias"):
                # softmax is invariant to a bias on the keys: the K third of qkv.bias has a mathematically ZERO gradient,
                # what both runs compute there is rounding noise, which AdamW normalises to +-lr.  Compare Q and V only.
                k = wa.numel() // 3
                wa, wb = torch.cat([wa[:k], wa[2 * k:]]), torch.cat([wb[:k], wb[2 * k:]])
            if wb.norm().item() <= 1e-6:
                continue
            d = (wa - wb).norm().item() / wb.norm().item()
            if n.endswith(".bias"):
                # zero-initialised: after a few steps the whole tensor IS the accumulated update (|w| ~ steps x lr), so the
                # elements whose gradient is at rounding-noise level weigh in fully; tracked with its own bound
                if d > worst_b:
                    worst_b, worst_b_name = d, n
            elif d > worst:
                worst, worst_name = d, n
        loss_rel = max(abs(x - y) / abs(y) for x, y in zip(losses, losses1))
        ok = replica_maxdiff == 0.0 and worst <= 1e-5 and worst_b <= 5e-3 and loss_rel <= 1e-5
        print(json.dumps({"tool": "ddp_parity", "world": world, "per_rank_batch": a.batch, "steps": a.steps,
                          "replica_param_maxdiff": replica_maxdiff, "weights_worst_rel_vs_single": worst,
                          "worst_tensor": worst_name, "zero_init_biases_worst_rel_vs_single": worst_b,
                          "worst_bias": worst_b_name, "loss_ddp": losses, "loss_single": losses1,
                          "loss_worst_rel": loss_rel, "grad_exchange": "p2p" if tr._peer is not None else "nccl", "ok": ok}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""MAE pre-training step benchmark (BASELINE.json metric: images/sec, fwd+bwd+clip+AdamW, device-timed).

  python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels via the C ABI)
  python bench.py --impl reference ...                       # the reference's CPU path (oracle port) on host cores

Workload at every N: configs/mae.yaml at batch 2000 per GPU (BASELINE.json configs[1]; weak scaling), synthetic
STL-10-shaped fp32 images, mask 0.75, bf16 tensor-core GEMMs with fp32 accumulation / fp32 master weights.
One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MAE pretrain images/sec (fwd+bwd+AdamW, device-timed)"  # BASELINE.json's metric; the reference arm is host-timed
GEMM_TN_DRAM_BYTES_PER_STEP = 10.332e9  # profiles/r02z_gemm_tn_dram_traffic.csv (B=2000, the 44 gemm_tn launches of one step: 6.37 GB read + 3.96 GB written)
FLOPS_PER_IMAGE = 1_124_504_064  # fwd+bwd, patch-embed on kept patches only, no patch-embed dgrad
CFG = {
    "general": {"image_size": 96, "patch_size": 8, "in_chans": 3},
    "encoder": {"embed_dim": 144, "depth": 4, "num_heads": 6},
    "decoder": {"decoder_embed_dim": 192, "decoder_depth": 2, "decoder_num_heads": 6},
}
PRETRAIN = {"mask_ratio_start": 0.75, "mask_ratio_end": 0.75, "mask_ramp_epochs": 5, "total_epochs": 800,
            "warmup_epochs": 20, "batch_size": 2000, "base_learning_rate": 1.5e-4, "weight_decay": 0.05}


VITB_CFG = {  # BASELINE.json configs[4]: ViT-B/16 MAE stress config (decoder = MAE paper's 512-d x 8 x 16 heads)
    "general": {"image_size": 224, "patch_size": 16, "in_chans": 3},
    "encoder": {"embed_dim": 768, "depth": 12, "num_heads": 12},
    "decoder": {"decoder_embed_dim": 512, "decoder_depth": 8, "decoder_num_heads": 16},
}
WORKLOADS = {
    # name: (model cfg, image size, default batch / GPU, algorithmic FLOPs per image fwd+bwd, label)
    "mae": (CFG, 96, 2000, FLOPS_PER_IMAGE,
            "configs/mae.yaml MAE pretrain step (96px, 8px patches, 144-d x4 enc, 192-d x2 dec)"),
    "vitb16": (VITB_CFG, 224, 256, 57_468_000_000,
               "ViT-B/16 MAE stress config (224px, 196 patches, 768-d x12 x12h enc, 512-d x8 x16h dec)"),
}
# classifier (BASELINE.json configs[3]): encoder forward_features over all 145 tokens + CLS head; fwd 345.3 MFLOP / image
# (4 blocks x 84.3 M + patch embed 8.0 M), x3 with the backward of a full fine-tune
CLS_FLOPS_FWD = 4 * (2 * 145 * 144 * 432 + 4 * 6 * 145 * 145 * 24 + 2 * 145 * 144 * 144 + 4 * 145 * 144 * 576) + 2 * 144 * 192 * 144


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], 0, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the upper half = clock while the GPU was busy (idle samples drag a plain median down)
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps: int, warmup: int, batch: int = 64, time_budget_s: float = 25.0):
    """The reference's CPU implementation of the step (oracle port: identical ATen op sequence, fp32 "highest"),
    on all host cores.  Returns (img/s, cores, ms_per_step, steps_done)."""
    import torch

    from oracle import mae_oracle as O

    torch.set_float32_matmul_precision("highest")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build_mae(seed=73)
    tr = O.OracleTrainer(model, batch_size=batch)
    g = torch.Generator().manual_seed(73)
    images = torch.randn(batch, 3, 96, 96, generator=g)
    for _ in range(warmup):
        tr.step(images)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        tr.step(images)
        done += 1
        if time.perf_counter() - t0 > time_budget_s:
            break
    dt = time.perf_counter() - t0
    return batch * done / dt, cores, 1000 * dt / done, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, cores, ms, done = cpu_reference_run(args.steps, max(args.warmup, 1), batch=64, time_budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s",
        # same metric string as the GPU arm (the driver pairs the two lines on it); this arm has no device: it is the
        # reference's CPU path timed with the host clock around whole steps
        "timed_with": "host perf_counter around whole CPU steps (no device in this arm)",
        "n_gpus": args.gpus, "steps": done, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs/mae.yaml MAE pretrain step, mask 0.75; CPU sample = batch 64 per step"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{done} steps of batch 64 (oracle port of src/models/mae.py + lightly/timm, fp32 highest)"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def algorithmic_cost(name, a):
    """(bytes, flops, shape label) one C-ABI call has to move / compute at minimum (every operand once; the argument order
    is include/mae_b200.h's).  Used for the per-kernel roofline table; unknown entry points return (None, None, "")."""
    nz = lambda p: 1 if p else 0
    if name == "mae_gemm_tn":
        M, N, K, epi = a[2], a[3], a[4], a[5]
        out_b = {0: 2, 1: 4, 2: 4, 4: 2}[epi]            # bf16 | bf16 h + bf16 gelu' | fp32 | bf16
        aux_b = {0: 0, 1: 0, 2: 4, 4: 2}[epi]            # - | - | fp32 residual | bf16 gelu'
        return M * K * 2.0 + N * K * 2.0 + M * N * (out_b + aux_b), 2.0 * M * N * K, f"M={M} N={N} K={K} epi={epi}"
    if name == "mae_mlp_fwd":
        rows, D = a[4], a[5]
        # read x, write xn + stats, h, gelu', x_out; weights once
        return rows * D * (4.0 + 2 + 4) + rows * 8.0 + rows * 4 * D * 4.0 + 2 * 4 * D * D * 2.0, 2.0 * rows * D * 4 * D * 2, \
            f"rows={rows} D={D}" + (" (gather)" if a[1] else "")
    if name == "mae_gemm_wgrad":
        Pc, Qc, M = a[1], a[3], a[4]
        return M * (Pc + Qc) * 2.0 + Pc * Qc * 4.0, 2.0 * M * Pc * Qc, f"M={M} P={Pc} Q={Qc}"
    if name in ("mae_attention_fwd", "mae_attention_fwd_q"):
        B, L, H, hd = a[1], a[2], a[3], a[4]
        return B * L * H * hd * 2.0 * 4 + B * H * L * 4.0, 4.0 * B * H * L * L * hd, f"B={B} L={L} H={H} hd={hd}"
    if name in ("mae_attention_bwd", "mae_attention_bwd_ws", "mae_attention_bwd_q"):
        B, L, H, hd = a[4], a[5], a[6], a[7]
        return B * L * H * hd * 2.0 * 8 + B * H * L * 4.0, 10.0 * B * H * L * L * hd, f"B={B} L={L} H={H} hd={hd}"
    if name == "mae_layernorm_fwd":
        rows, D = a[6], a[7]
        return rows * D * 6.0 + rows * 8.0, 0.0, f"rows={rows} D={D}"
    if name == "mae_layernorm_bwd":
        rows, D = a[9], a[10]
        return rows * D * (2.0 + 4 + 4 * nz(a[5]) + 4 + 2 * nz(a[12])) + rows * 8.0, 0.0, f"rows={rows} D={D}"
    if name == "mae_layernorm_fwd_add":
        rows, D = a[7], a[8]
        return rows * D * 12.0 + rows * 8.0, 0.0, f"rows={rows} D={D} (+delta)"
    if name == "mae_layernorm_bwd_scatter":
        rows, D = a[9], a[10]
        return rows * D * (2.0 + 4 + 4 + 4 + 2 * nz(a[12])) + rows * 8.0, 0.0, f"rows={rows} D={D} (scatter)"
    if name == "mae_random_token_mask":
        B, L = a[1], a[2]
        return B * L * 12.0, 0.0, f"B={B} L={L}"
    if name == "mae_gather_patches":
        B, C, patch, nk = a[2], a[3], a[5], a[6]
        return B * (nk - 1) * C * patch * patch * 6.0, 0.0, f"B={B} kept={nk - 1}"
    if name == "mae_patch_embed_tokens":
        B, C, patch, nk, D = a[2], a[3], a[5], a[6], a[7]
        pd = C * patch * patch
        rows = B * (nk - 1)
        # read the kept patches (fp32), write the im2col rows (bf16) and the fp32 stream incl. class-token rows; weight once
        return rows * pd * 6.0 + B * nk * D * 4.0 + D * pd * 2.0, 2.0 * rows * D * pd, f"B={B} kept={nk - 1} D={D}"
    if name == "mae_assemble_tokens":
        B, nk, D = a[4], a[5], a[6]
        return B * nk * D * 6.0, 0.0, f"B={B} kept={nk} D={D}"
    if name in ("mae_scatter_decoder_tokens", "mae_scatter_decoder_tokens_grouped"):
        B, L, nk, D = a[5], a[6], a[7], a[8]
        return B * nk * D * 2.0 + B * L * D * 4.0, 0.0, f"B={B} L={L} D={D}"
    if name == "mae_scatter_decoder_tokens_ln":
        B, L, nk, D = a[5], a[6], a[7], a[8]
        return B * nk * D * 2.0 + B * L * D * 6.0 + B * L * 8.0, 0.0, f"B={B} L={L} D={D} (+norm1)"
    if name == "mae_masked_mse":
        B, C, patch, nm = a[3], a[4], a[6], a[7]
        return B * nm * C * patch * patch * (2.0 + 4 + 2 * nz(a[11]) + 4 * nz(a[12])), 0.0, f"B={B} masked={nm}"
    if name == "mae_colsum_bf16":
        return a[1] * a[2] * 2.0, 0.0, f"M={a[1]} N={a[2]}"
    if name == "mae_gather_rows_bf16":
        return a[2] * a[3] * a[6] * 6.0, 0.0, f"rows={a[2] * a[3]} D={a[6]}"
    if name == "mae_zero_rows":
        return a[3] * a[4] * a[6] * (4.0 + 2 * nz(a[1])), 0.0, f"rows={a[3] * a[4]} D={a[6]}"
    if name == "mae_colsum_rows_f32":
        return a[2] * a[3] * a[6] * 4.0, 0.0, f"rows={a[2] * a[3]} D={a[6]}"
    if name == "mae_cast_bf16":
        return a[2] * 6.0, 0.0, f"n={a[2]}"
    if name == "mae_allreduce_avg_sumsq":
        return a[1] * a[3] * 4.0 + a[3] * 4.0, 0.0, f"world={a[1]} n={a[3]}"  # (world - 1) of the reads cross NVLink
    if name == "mae_grad_sumsq":
        return a[1] * 4.0, 0.0, f"n={a[1]}"
    if name == "mae_clip_adamw":
        return a[5] * 30.0, 0.0, f"n={a[5]}"  # read p, g, m, v; write p, m, v and the bf16 shadow
    if name == "mae_augment_params":
        return a[1] * 20.0, 0.0, f"B={a[1]}"
    if name == "mae_augment_u8":
        B, C, S = a[2], a[3], a[6]
        return B * C * S * S * 5.0, 0.0, f"B={B} S={S}"  # lower bound: a crop at least as large as the output
    return None, None, ""


def kernel_breakdown(trainer, images, noise, pk):
    """One instrumented step: CUDA events around every C-ABI call -> time share per kernel family, the roofline of the
    dominant one (gemm_tn: algorithmic bytes of its launches / their summed duration) and a per-kernel, per-shape roofline
    table (achieved GB/s and TFLOP/s against the measured peaks)."""
    import torch

    from ssrl_vit_mae_jepa_b200._lib import calls

    lib_calls = calls
    records = []
    orig = {}

    def wrap(name):
        fn = getattr(lib_calls, name)

        def timed(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(*a)
            e.record()
            records.append((name, s, e, a))

        return fn, timed

    from ssrl_vit_mae_jepa_b200 import _lib
    for name in _lib.SIGNATURES:
        orig[name], t = wrap(name)
        setattr(lib_calls, name, t)
    was_distributed = trainer.distributed
    trainer.distributed = False  # rank-0-only diagnostic step: it must not enter a collective
    was_side = trainer.engine.wgrad_side_stream
    trainer.engine.wgrad_side_stream = False  # per-call events need every launch on the timed stream
    try:
        trainer.train_step(images, noise)
        torch.cuda.synchronize()
    finally:
        trainer.distributed = was_distributed
        trainer.engine.wgrad_side_stream = was_side
        for name, fn in orig.items():
            setattr(lib_calls, name, fn)
    fam = {}
    table = {}
    gemm_flops = 0.0
    gemm_bytes = 0.0
    wgrad_flops = 0.0
    for name, s, e, a in records:
        ms = s.elapsed_time(e)
        d = fam.setdefault(name, [0.0, 0])
        d[0] += ms; d[1] += 1
        nbytes, flops, label = algorithmic_cost(name, a)
        if nbytes is not None:
            t = table.setdefault((name, label), [0.0, 0, 0.0, 0.0])
            t[0] += ms; t[1] += 1; t[2] += nbytes; t[3] += flops
        if name == "mae_gemm_tn":
            gemm_flops += flops
            gemm_bytes += nbytes
        if name == "mae_gemm_wgrad":
            wgrad_flops += flops
    rows = []
    for (name, label), (ms, n, nbytes, flops) in sorted(table.items(), key=lambda kv: -kv[1][0]):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        tfs = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        rows.append({"kernel": name[4:], "shape": label, "launches": n, "us_per_launch": round(ms * 1e3 / n, 1),
                     "mb_per_launch": round(nbytes / n / 1e6, 2), "gbs": round(gbs, 1), "hbm_frac": round(gbs / pk["hbm_gbs"], 3),
                     "tflops": round(tfs, 1), "tensor_frac": round(tfs / pk["tf_sustained"], 3)})
    total = sum(v[0] for v in fam.values())
    share = {k: {"ms": round(v[0], 4), "launches": v[1], "share": round(v[0] / total, 4)} for k, v in
             sorted(fam.items(), key=lambda kv: -kv[1][0])}
    g_ms, n_gemm = fam.get("mae_gemm_tn", [0.0, 0])
    roof = None
    if g_ms > 0:
        gbs = gemm_bytes / (g_ms * 1e-3) / 1e9
        tfs = gemm_flops / (g_ms * 1e-3) / 1e12
        roof = {"kernel": f"gemm_tn_kernel (tcgen05 + TMA; the {n_gemm} forward/dgrad GEMM launches of one step)",
                "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                # ncu dram__bytes_read.sum + dram__bytes_write.sum over the 44 gemm_tn launches of one step of this build
                # (profiles/r02z_gemm_tn_dram_traffic.csv: 6.37 GB read + 3.96 GB written), per launch like the figure below
                "traffic": (GEMM_TN_DRAM_BYTES_PER_STEP / max(1, n_gemm)) if images.shape[0] == 2000 else None,
                "algorithmic_bytes_per_launch": gemm_bytes / max(1, n_gemm),
                "peak_source": pk["source"] + " (copy bandwidth)",
                "traffic_note": "measured once with ncu (cold L2 per launch), not live: 10.33 GB per step vs the algorithmic "
                                "bytes below -> no wasted re-reads (outputs still resident in L2 at kernel end are not counted)",
                "algorithmic_bytes_per_step": gemm_bytes, "ms_per_step": g_ms,
                "tensor": {"achieved_tflops": tfs, "peak_tflops": pk["tf_sustained"], "frac": tfs / pk["tf_sustained"],
                           "flops_per_step": gemm_flops}}
    w_ms = fam.get("mae_gemm_wgrad", [0.0, 0])[0]
    extra = {"gemm_wgrad_tflops": (wgrad_flops / (w_ms * 1e-3) / 1e12) if w_ms > 0 else None,
             "instrumented_step_ms": total, "kernel_roofline": rows}
    return share, roof, extra


def run_ours(args):
    import torch
    import torch.distributed as dist

    import ssrl_vit_mae_jepa_b200 as P
    from ssrl_vit_mae_jepa_b200._lib import calls

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MAE hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl_cfg, IMG, wl_batch, wl_flops, wl_label = WORKLOADS[args.workload]
    B = args.batch or wl_batch
    torch.manual_seed(73 + rank)  # rank-distinct noise / data streams
    mod = P.MAEPretrainModule(wl_cfg, dict(PRETRAIN, batch_size=B)).to(dev)
    if world > 1:  # identical replicas
        for p in mod.parameters():
            dist.broadcast(p.data, 0)
    tr = P.FusedTrainer(mod)
    tr.set_epoch(0)
    pool = [torch.randn(B, 3, IMG, IMG, device=dev) for _ in range(3)]  # mae: 221 MB each (> 126 MB L2 between reuses)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # covers warm-up + timed region: the GPU is under the same load in both
    for i in range(max(args.warmup, 10)):  # >= 10 untimed steps also give nvidia-smi time to produce samples
        tr.train_step(pool[i % len(pool)])
    barrier()
    l0 = calls.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = tr.train_step(pool[i % len(pool)])
    ev1.record()
    barrier()
    launches = calls.launches - l0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    rank_ms = [ms / args.steps]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [x.item() / args.steps for x in allt]  # per-rank device time per step: the skew the MAX is made of
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms * 1e-3)

    # end to end through the public API: pinned host batches -> (prefetching) H2D -> step -> D2H of the loss,
    # every step.  The copy of batch i+1 overlaps step i on a side stream; it is still inside the timed region.
    host = [torch.randn(B, 3, IMG, IMG).pin_memory() for _ in range(2)]
    e2e_steps = max(3, min(args.steps, 10))
    pf = P.training.DevicePrefetcher((host[i % 2] for i in range(e2e_steps + 2)), dev)
    for _ in range(2):
        tr.train_step(next(pf)).item()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        lv = tr.train_step(next(pf)).item()
    ev1.record()
    barrier()
    e_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (t.item() * 1e-3)

    # the same end-to-end loop fed with raw uint8 batches (what torchvision's STL10 stores): 4x less H2D traffic, the
    # reference's RandomResizedCrop + flip + ToTensor + Normalize run on the GPU (mae_augment_u8)
    aug = P.data.GpuAugment(size=IMG, train=True)  # boxes / flips drawn on the device (mae_augment_params)
    host8 = [torch.randint(0, 256, (B, 3, IMG, IMG), dtype=torch.uint8).pin_memory() for _ in range(2)]
    pf8 = P.training.DevicePrefetcher((host8[i % 2] for i in range(e2e_steps + 2)), dev, transform=aug)
    for _ in range(2):
        tr.train_step(next(pf8)).item()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        tr.train_step(next(pf8)).item()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_u8_value = world * B * e2e_steps / (t.item() * 1e-3)

    # replicas must stay bit-identical: every rank applied the same averaged gradient with the same optimizer state
    replica_maxdiff = None
    if world > 1:
        flat = mod.model.engine.flat
        mx, mn = flat.clone(), flat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        replica_maxdiff = (mx - mn).abs().max().item()

    if rank == 0:
        pk = peaks()
        share, roof, extra = kernel_breakdown(tr, pool[0], None, pk)
        extra["rank_ms_per_step"] = [round(x, 4) for x in rank_ms]
        extra["grad_exchange"] = ("p2p: own kernel over NVLink peer memory (mae_allreduce_avg_sumsq)" if getattr(tr, "_peer", None)
                                  is not None else ("nccl buckets" if world > 1 else None))
        extra["replica_param_maxdiff"] = replica_maxdiff  # N > 1: max |param(rank i) - param(rank j)| after all steps; must be 0.0
        extra["e2e_uint8_gpu_augment"] = {"value": e2e_u8_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * IMG * IMG,
                                          "note": "uint8 host batches; crop boxes / flips drawn on the device, RandomResizedCrop/flip/normalize on the prefetch stream, all inside the timed loop"}
        step_tf = value / world * wl_flops / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "timed_with": "CUDA events on the launching stream, max over ranks",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{wl_label}, batch {B}/GPU, mask 0.75, fwd+bwd+clip+AdamW", "global_batch": B * world,
                       "parallelism": f"dp{world}",
                       "l2_policy": f"3 rotating {B * 3 * IMG * IMG * 4 / 1e6:.0f} MB input batches (> 126 MB L2)"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * IMG * IMG * 4,
                    "d2h_bytes_per_step": 4, "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clocks, "final_loss": float(loss.item()),
            "roofline": roof,
            "step_roofline": {"bound": "tensor", "achieved": step_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                              "frac": step_tf / pk["tf_sustained"], "flops_per_image": wl_flops,
                              "peak_source": pk["source"]},
            "kernel_share": share, "extra": extra,
        }
        if world == 1 and not args.no_cpu and args.workload == "mae":
            ips, cores, cms, done = cpu_reference_run(10, 2, batch=64, time_budget_s=20.0)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{done} steps of batch 64, oracle port, fp32 highest, {cms:.0f} ms/step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_classifier(args):
    """BASELINE.json configs[3]: downstream classifier on the 144-d ViT (reference src/models/classifier.py:47-57,
    scripts/training/train_mae.py:205-215), batch 2000: full fine-tune (value) and frozen-encoder probe (extra), each step =
    training_step + backward + clip 1.0 + torch AdamW through the product classes.  Extra bench line; single GPU."""
    import torch
    import torch.nn.functional as F  # noqa: F401

    import ssrl_vit_mae_jepa_b200 as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = args.batch or 2000
    torch.manual_seed(73)
    cfg = dict(CFG, head={"embed_dim": 144, "pool": "cls"})
    pool = [torch.randn(B, 3, 96, 96, device=dev) for _ in range(3)]
    labels = torch.randint(0, 10, (B,), device=dev)
    out = {}
    for mode, freeze in (("full_finetune", False), ("frozen_probe", True)):
        mae = P.MaskedAutoencoder(CFG["general"], CFG["encoder"], CFG["decoder"])
        mod = P.ViTClassifierTrainModule(pretrained_encoder=mae.encoder.vit, model_cfg=cfg,
                                         training_cfg={"freeze_encoder": freeze, "learning_rate": 1e-4}).to(dev)
        opt = mod.configure_optimizers()["optimizer"]
        params = [p for p in mod.parameters() if p.requires_grad]

        def step(x):
            opt.zero_grad(set_to_none=True)
            loss = mod.training_step((x, labels), 0)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return loss

        for i in range(max(args.warmup, 5)):
            step(pool[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(pool[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[mode] = {"images_per_s": B / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss.item())}
    pk = peaks()
    v = out["full_finetune"]["images_per_s"]
    tf_ft, tf_probe = v * 3 * CLS_FLOPS_FWD / 1e12, out["frozen_probe"]["images_per_s"] * CLS_FLOPS_FWD / 1e12
    print(json.dumps({
        "metric": "ViT classifier images/sec (fwd+bwd+AdamW, device-timed)", "value": v, "unit": "images/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": out["full_finetune"]["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "timed_with": "CUDA events on the launching stream",
        "config": {"workload": f"train_mae full fine-tune of the 144-d ViT classifier (all 145 tokens, CLS head), batch {B}",
                   "global_batch": B, "parallelism": "dp1"},
        "step_roofline": {"bound": "tensor", "achieved": tf_ft, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                          "frac": tf_ft / pk["tf_sustained"], "flops_per_image": 3 * CLS_FLOPS_FWD, "peak_source": pk["source"]},
        "extra": {"frozen_probe": dict(out["frozen_probe"], tflops=tf_probe, tensor_frac=tf_probe / pk["tf_sustained"],
                                       flops_per_image=CLS_FLOPS_FWD),
                  "full_finetune": out["full_finetune"]},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's, 2000 for configs/mae.yaml)")
    ap.add_argument("--workload", default="mae", choices=["mae", "vitb16", "classifier"],
                    help="mae = the headline (BASELINE.json configs[1..2]); vitb16 / classifier = extra lines for configs[4] / [3]")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "classifier":
        run_classifier(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""CPU-only checks of the drop-in boundary: Python surface, state_dict layout, C-ABI symbol export, schedule
helpers and the world_size-2 (gloo) gradient-bucket all-reduce.  No kernel is executed here."""
import ctypes
import os
import re

import pytest
import torch

import ssrl_vit_mae_jepa_b200 as P
from oracle import mae_oracle as O
from ssrl_vit_mae_jepa_b200 import _lib, training

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    return P.MaskedAutoencoder(O.MAE_YAML["general"], O.MAE_YAML["encoder"], O.MAE_YAML["decoder"])


def test_constructor_defaults_and_attributes():
    m = _build()
    assert (m.mask_ratio, m.image_size, m.patch_size, m.in_chans, m.sequence_length) == (0.75, 96, 8, 3, 145)
    assert m.encoder.vit.embed_dim == 144 and len(m.encoder.vit.blocks) == 4
    assert m.decoder.mask_token.shape == (1, 1, 192)
    # reference defaults (src/models/mae.py:23-26,32-34,49-51) are decoder 512-d / 6 heads, which timm's own
    # Attention rejects (512 % 6 != 0); dims outside the kernels' range raise NotImplementedError
    with pytest.raises(AssertionError):
        P.MaskedAutoencoder({}, {}, {})
    with pytest.raises(NotImplementedError):
        P.MaskedAutoencoder({"patch_size": 8}, {"embed_dim": 160, "depth": 2, "num_heads": 4},  # head_dim 40
                            {"decoder_embed_dim": 192, "decoder_depth": 1})
    vitb = P.MaskedAutoencoder({"image_size": 224, "patch_size": 16}, {"embed_dim": 768, "depth": 1, "num_heads": 12},
                               {"decoder_embed_dim": 512, "decoder_depth": 1, "decoder_num_heads": 16})
    assert vitb.sequence_length == 197 and P.num_keep_tokens(197, 0.75) == 49


def test_state_dict_is_interchangeable_with_reference_layout():
    m, o = _build(), O.build_mae()
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    for k, v in o.state_dict().items():
        assert m.state_dict()[k].shape == v.shape and m.state_dict()[k].dtype == v.dtype
    m.load_state_dict(o.state_dict())
    o.load_state_dict(m.state_dict())
    assert sum(p.numel() for p in m.parameters()) == 2_035_104
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 1_986_384
    assert not m.encoder.vit.pos_embed.requires_grad and not m.decoder.decoder_pos_embed.requires_grad
    assert torch.equal(m.encoder.vit.pos_embed, o.encoder.vit.pos_embed)
    assert torch.equal(m.decoder.decoder_pos_embed, o.decoder.decoder_pos_embed)
    assert not any("head" in n for n, _ in m.encoder.named_parameters())  # src/training/classifier.py:129-131


def test_seeded_init_matches_oracle_distribution():
    torch.manual_seed(73)
    m = _build()
    o = O.build_mae(seed=73)
    for (n, a), (_, b) in zip(m.state_dict().items(), o.state_dict().items()):
        assert abs(a.std().item() - b.std().item()) <= 0.15 * b.std().item() + 1e-6, n


def test_cpu_tensors_raise_no_fallback():
    m = _build()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 3, 96, 96))
    with pytest.raises(RuntimeError):
        m.forward_loss(torch.randn(2, 3, 96, 96))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mae_b200.h")).read()
    declared = set(re.findall(r"\b(mae_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(P.build())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mae_b200.h but not exported"
    assert declared - {"mae_version", "mae_error_string"} == set(_lib.SIGNATURES)
    lib.mae_version.restype = ctypes.c_int
    assert lib.mae_version() == 100
    lib.mae_error_string.restype = ctypes.c_char_p
    assert b"shape" in lib.mae_error_string(-1)


def test_pretrain_module_mirrors_reference():
    cfg = {"general": O.MAE_YAML["general"], "encoder": O.MAE_YAML["encoder"], "decoder": O.MAE_YAML["decoder"]}
    mod = P.MAEPretrainModule(cfg, dict(O.MAE_YAML_PRETRAIN))
    assert mod.batch_size == 2000 and abs(mod.effective_lr - 1.5e-4 * 2000 / 256) < 1e-12
    oc = mod.configure_optimizers()
    opt, sch = oc["optimizer"], oc["lr_scheduler"]["scheduler"]
    assert oc["lr_scheduler"]["interval"] == "epoch"
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["weight_decay"] == 0.05
    assert len(opt.param_groups) == 1 and len(opt.param_groups[0]["params"]) == 87
    assert abs(sch.get_last_lr()[0] - mod.effective_lr / 20) < 1e-12
    for e in (0, 3, 19, 20, 400, 799):
        assert training.lr_factor(e, 20, 800) == O.lr_lambda(e, 20, 800)
        assert training.mask_ratio_for_epoch(e, 0.5, 0.85, 200) == O.mask_ratio_at_epoch(e, 0.5, 0.85, 200)
    mod.current_epoch = 3
    mod.on_train_epoch_start()
    assert mod.model.mask_ratio == 0.75 and mod.logged["mask_ratio"] == 0.75
    # defaults when keys are missing (src/training/mae.py:31-39)
    d = P.MAEPretrainModule(cfg, {})
    assert (d.mask_start, d.mask_end, d.ramp_epochs, d.batch_size, d.total_epochs) == (0.5, 0.85, 200, 512, 200)


def test_num_keep_matches_lightly_rule():
    for L, r in [(145, 0.75), (145, 0.5), (145, 0.85), (197, 0.75), (145, 0.6123)]:
        assert P.num_keep_tokens(L, r) == O.num_keep_tokens(L, r)
    assert P.num_keep_tokens(145, 0.75) == 36


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    b = training.GradBuckets(flat, [0, 4, 10])
    b.launch(1)  # "decoder" bucket first, like the backward pass
    b.launch(0)
    b.finish()
    q.put((rank, flat.tolist()))
    dist.destroy_process_group()


def test_grad_buckets_world_size_2_gloo():
    import socket
    import torch.multiprocessing as mp

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in ps]
    expect = [i * 1.5 for i in range(10)]  # mean of x and 2x
    assert out[0] == expect and out[1] == expect


def test_classifier_module_mirrors_reference():
    from src.models.classifier import ViTClassifier
    from src.training.classifier import ViTClassifierTrainModule
    cfg = {"general": O.MAE_YAML["general"], "encoder": O.MAE_YAML["encoder"], "decoder": O.MAE_YAML["decoder"],
           "head": {"embed_dim": 144, "pool": "cls"}}
    mae = _build()
    mod = ViTClassifierTrainModule(pretrained_encoder=mae.encoder.vit, model_cfg=cfg, training_cfg={"freeze_encoder": True})
    # notebook.ipynb:1689-1698: 1.4 K trainable (1,450) / 1.1 M frozen (1,051,920)
    assert sum(p.numel() for p in mod.parameters() if p.requires_grad) == 1450
    assert sum(p.numel() for p in mod.parameters() if not p.requires_grad) == 1_051_920
    assert isinstance(mod.model, ViTClassifier) and mod.model.pool_type == "cls"
    assert list(mod.model.state_dict())[-2:] == ["head.classification.weight", "head.classification.bias"]
    assert list(mod.model.encoder.state_dict()) == list(O.VisionTransformer().state_dict())
    mod.unfreeze_last_layers(1)
    tr = {n for n, p in mod.model.named_parameters() if p.requires_grad}
    assert all(n.startswith(("encoder.blocks.3.", "encoder.norm.", "head.")) for n in tr) and len(tr) == 12 + 2 + 2
    with pytest.raises(ValueError):
        mod.unfreeze_last_layers(5)
    oc = mod.configure_optimizers()
    assert oc["lr_scheduler"]["monitor"] == "val_loss" and len(oc["optimizer"].param_groups[0]["params"]) == 16
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.randn(2, 3, 96, 96))


def test_data_module_synthetic_fallback(monkeypatch):
    import yaml
    from src.data import get_pretrain_dataloaders, get_train_dataloaders
    monkeypatch.setenv("MAE_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("MAE_SYNTHETIC_SIZE", "120")
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "mae.yaml")))
    assert cfg["pretrain"]["batch_size"] == 2000 and cfg["model"]["encoder"]["embed_dim"] == 144  # reference values
    cfg["pretrain"].update(batch_size=8, num_workers=0, val_split=0.25)
    cfg["train"].update(batch_size=8, num_workers=0, samples_per_class=3)
    tr, va = get_pretrain_dataloaders(cfg)
    assert len(tr.dataset) == 90 and len(va.dataset) == 30
    x, y = next(iter(tr))
    assert x.shape == (8, 3, 96, 96) and x.dtype == torch.float32 and -1.0 <= x.min() and x.max() <= 1.0
    x2, _ = next(iter(get_pretrain_dataloaders(cfg)[1]))
    x3, _ = next(iter(get_pretrain_dataloaders(cfg)[1]))
    assert torch.equal(x2, x3)  # deterministic
    tr, va = get_train_dataloaders(cfg)
    labels = torch.cat([y for _, y in tr])
    assert len(tr.dataset) == 30 and torch.bincount(labels, minlength=10).tolist() == [3] * 10


def test_encoder_checkpoint_prefix_detection(tmp_path):
    from scripts.utils import load_encoder_weights
    src = _build()
    for prefix, wrap in [("encoder.", False), ("model.encoder.", True), ("module.encoder.", False)]:
        sd = {(("model." if prefix.startswith("model.") else "module." if prefix.startswith("module.") else "") + k): v
              for k, v in src.state_dict().items()}
        path = tmp_path / f"w_{prefix.strip('.')}.pt"
        torch.save({"state_dict": sd} if wrap else sd, path)
        dst = _build()
        assert load_encoder_weights(dst.encoder, str(path)) == prefix
        assert torch.equal(dst.encoder.vit.blocks[2].mlp.fc1.weight, src.encoder.vit.blocks[2].mlp.fc1.weight)
    torch.save({"foo.bar": torch.zeros(1)}, tmp_path / "bad.pt")
    with pytest.raises(RuntimeError, match="no encoder weights"):
        load_encoder_weights(_build().encoder, str(tmp_path / "bad.pt"))


def test_cli_flags_match_reference():
    from scripts.training import pretrain_mae, train_mae
    a = pretrain_mae.parse_args([])
    assert (a.config, a.resume_from, a.output_dir_suffix) == ("configs/mae.yaml", None, "mae_pretrain")
    b = train_mae.parse_args(["--encoder_ckpt", "x.pt"])
    assert (b.config, b.encoder_ckpt, b.classifier_ckpt, b.output_dir_suffix) == ("configs/mae.yaml", "x.pt", None, "mae_finetune")


def test_bench_algorithmic_cost_covers_every_entry_point_and_renders(tmp_path):
    """bench.py's per-kernel roofline table: every compute entry point of the C ABI has a byte / FLOP formula (a new kernel
    without one would silently drop out of the table), spot values follow SURVEY §8d, and tools/roofline_table.py renders a
    bench line."""
    import importlib.util
    import json
    import subprocess
    import sys

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name, sig in _lib.SIGNATURES.items():
        if name in ("mae_version", "mae_error_string", "mae_transpose_cast_bf16"):
            continue
        args = [3 if t is ctypes.c_int or t is ctypes.c_longlong else (1.0 if t is ctypes.c_float else 4096) for t in sig]
        if name == "mae_gemm_tn":
            args[5] = 0  # a valid epilogue id
        nbytes, flops, label = bench.algorithmic_cost(name, args)
        assert nbytes is not None and nbytes > 0 and label, name
    # decoder qkv GEMM of configs/mae.yaml at B=2000: A + W + C bytes, 2 M N K flops
    nbytes, flops, _ = bench.algorithmic_cost("mae_gemm_tn", [0, 0, 290000, 576, 192, 0, 0, 0, 0, 0, 0])
    assert nbytes == 290000 * 192 * 2 + 576 * 192 * 2 + 290000 * 576 * 2 and flops == 2 * 290000 * 576 * 192
    # decoder attention: 4 L^2 hd flops per (sample, head) forward
    _, flops, _ = bench.algorithmic_cost("mae_attention_fwd", [0, 2000, 145, 6, 32, 0, 0, 0])
    assert flops == 4.0 * 2000 * 6 * 145 * 145 * 32
    line = {"ms_per_step": 8.0, "value": 250000.0, "unit": "images/s",
            "roofline": {"peak": 6553.0, "peak_source": "measured"}, "step_roofline": {"peak": 1393.7},
            "extra": {"instrumented_step_ms": 8.1, "kernel_roofline": [
                {"kernel": "gemm_tn", "shape": "M=1 N=1 K=1 epi=0", "launches": 2, "us_per_launch": 10.0, "mb_per_launch": 1.0,
                 "gbs": 100.0, "hbm_frac": 0.02, "tflops": 1.0, "tensor_frac": 0.001},
                {"kernel": "layernorm_fwd", "shape": "rows=1 D=4", "launches": 1, "us_per_launch": 5.0, "mb_per_launch": 0.5,
                 "gbs": 100.0, "hbm_frac": 0.02, "tflops": 0.0, "tensor_frac": 0.0}]}}
    path = tmp_path / "b.json"
    path.write_text("noise\n" + json.dumps(line) + "\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "roofline_table.py"), str(path)], capture_output=True,
                         text=True, check=True).stdout
    assert "| gemm_tn | M=1 N=1 K=1 epi=0 | 2 |" in out and "| layernorm_fwd |" in out and "– | –" in out


def test_gradient_bucket_boundaries_follow_the_blocks():
    """One all-reduce bucket per transformer block (reverse completion order in backward): contiguous, covering exactly
    the optimised part of the flat gradient buffer, every parameter inside one bucket, stem + block 0 in the first."""
    import ssrl_vit_mae_jepa_b200 as P
    from oracle import mae_oracle as O

    m = P.MaskedAutoencoder(O.MAE_YAML["general"], O.MAE_YAML["encoder"], O.MAE_YAML["decoder"])
    eng = m.engine
    eng._compute_layout()
    b = eng.grad_bucket_boundaries()
    assert b[0] == 0 and b[-1] == eng.n_opt and b == sorted(b) and len(set(b)) == len(b) and len(b) - 1 == 4 + 2
    which = {}
    for n, p_ in eng._param_list:
        o = eng.offsets[n]
        if o >= eng.n_opt:
            assert n in eng.NOGRAD
            continue
        i = max(k for k in range(len(b) - 1) if b[k] <= o)
        assert o + p_.numel() <= b[i + 1], n
        which[n] = i
    assert which["encoder.vit.patch_embed.proj.weight"] == 0 and which["encoder.vit.blocks.0.mlp.fc2.bias"] == 0
    assert which["encoder.vit.blocks.3.norm1.weight"] == 3 and which["encoder.vit.norm.bias"] == 3
    assert which["decoder.mask_token"] == 4 and which["decoder.decoder_blocks.0.mlp.fc2.bias"] == 4
    assert which["decoder.decoder_blocks.1.norm1.weight"] == 5 and which["decoder.decoder_pred.bias"] == 5


def test_data_module_needs_explicit_opt_in_for_synthetic_data(monkeypatch, tmp_path):
    """No STL-10 on disk, no network, no MAE_SYNTHETIC_DATA=1: the loaders raise instead of silently training on noise;
    with the opt-in, validation subsets read through the evaluation (no-augmentation) dataset object."""
    import yaml
    from ssrl_vit_mae_jepa_b200 import data as D
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "mae.yaml")))
    cfg["pretrain"].update(batch_size=8, num_workers=0)
    monkeypatch.delenv("MAE_SYNTHETIC_DATA", raising=False)
    monkeypatch.setattr(D, "DATA_DIR", str(tmp_path / "nowhere"))
    with pytest.raises(RuntimeError, match="MAE_SYNTHETIC_DATA=1"):
        D.get_pretrain_dataloaders(cfg)
    monkeypatch.setenv("MAE_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("MAE_SYNTHETIC_SIZE", "100")
    tr, va = D.get_pretrain_dataloaders(cfg)
    assert tr.dataset.dataset.train is True and va.dataset.dataset.train is False
    assert set(tr.dataset.indices).isdisjoint(va.dataset.indices) and len(tr.dataset) + len(va.dataset) == 100
    cfg["train"].update(batch_size=8, num_workers=0, samples_per_class=4)
    tr, va = D.get_train_dataloaders(cfg)
    assert tr.dataset.dataset.train is True and va.dataset.dataset.train is False
    assert D.get_train_dataloaders(cfg)[0].dataset.indices == tr.dataset.indices  # per-class default_rng(seed) shuffle


REF = "/root/reference"


def _import_reference_script(name):
    """Import ``/root/reference/scripts/training/<name>.py`` UNMODIFIED as ``scripts.training.<name>`` of a package rooted
    at the reference checkout, while ``src.*`` resolves to this repo and ``pytorch_lightning`` / ``timm`` to the real
    packages or, when they are not installed (this image), to the stand-ins under ``ssrl-vit-mae-jepa_b200/shims``."""
    import importlib.util
    import sys
    import types

    import ssrl_vit_mae_jepa_b200  # noqa: F401  (appends the shim directory to sys.path)

    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "scripts" or k.startswith("scripts.")}
    try:
        for pkg, sub in (("scripts", "scripts"), ("scripts.training", "scripts/training")):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF, sub)]
            sys.modules[pkg] = m
        for mod, rel in (("scripts.utils", "scripts/utils.py"), (f"scripts.training.{name}", f"scripts/training/{name}.py")):
            spec = importlib.util.spec_from_file_location(mod, os.path.join(REF, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[mod] = m
            spec.loader.exec_module(m)
        return sys.modules[f"scripts.training.{name}"]
    finally:
        for k in [k for k in sys.modules if k == "scripts" or k.startswith("scripts.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scripts", "training", "pretrain_mae.py")),
                    reason="the reference checkout only exists in the build container")
def test_reference_scripts_run_unmodified_up_to_fit(monkeypatch, tmp_path):
    """the reference's own ``scripts/training/pretrain_mae.py`` and ``train_mae.py`` execute unmodified against
    this repo's ``src.*`` and the pytorch_lightning / timm stand-ins: argument parsing, YAML, data loaders, module
    construction, callbacks, logger and Trainer.  ``Trainer.fit`` itself needs a GPU (the kernels have no CPU fallback),
    so here it is intercepted and its arguments are checked; tests/test_gpu_drivers.py runs the same loop on the GPU."""
    import sys

    import pytorch_lightning as pl
    import yaml
    monkeypatch.setenv("MAE_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("MAE_SYNTHETIC_SIZE", "64")
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "mae.yaml")))  # the reference's own config file
    cfg["pretrain"].update(batch_size=16, num_workers=0)
    cfg["train"].update(batch_size=16, num_workers=0, samples_per_class=2)
    cfg["logging"]["output_dir_base"] = str(tmp_path / "outputs")
    cfg_path = tmp_path / "cfg.yaml"
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    seen = {}

    def fake_fit(self, model, train_dataloaders=None, val_dataloaders=None, datamodule=None, ckpt_path=None):
        seen["fit"] = (self, model, train_dataloaders, val_dataloaders, ckpt_path)

    monkeypatch.setattr(pl.Trainer, "fit", fake_fit)
    # the reference's scripts switch float32 matmuls to "medium" at import (scripts/utils.py:22): undo it afterwards, the
    # oracle tests of this process need "highest"
    prec = torch.get_float32_matmul_precision()
    monkeypatch.setattr(torch, "set_float32_matmul_precision", lambda *_a, **_k: None)
    assert torch.get_float32_matmul_precision() == prec
    for script, argv in (("pretrain_mae", ["--output_dir_suffix", "p"]), ("train_mae", ["--output_dir_suffix", "f"])):
        mod = _import_reference_script(script)
        assert mod.__file__.startswith(REF)
        monkeypatch.setattr(sys, "argv", [script, "--config", str(cfg_path)] + argv)
        mod.main()
        trainer, module, tl, vl, ckpt = seen.pop("fit")
        assert trainer.max_epochs == cfg["pretrain" if script == "pretrain_mae" else "train"]["total_epochs"]
        assert trainer.gradient_clip_val == 1.0 and len(trainer.callbacks) in (3, 4)
        assert isinstance(module, pl.LightningModule) and module.__class__.__module__.startswith("ssrl_vit_mae_jepa_b200")
        x, _ = next(iter(tl))
        assert x.shape == (16, 3, 96, 96) and vl is not None and ckpt is None
        out = tmp_path / "outputs" / ("pretrain/p" if script == "pretrain_mae" else "train/f")
        assert (out / "config.yaml").exists() and (out / "vit-mae.pt").exists()  # the script's own torch.save ran
    sd = torch.load(tmp_path / "outputs" / "pretrain" / "p" / "vit-mae.pt")
    assert len(sd) == 87


def test_torch_library_ops_are_registered_cuda_only():
    """torch.ops.mae_b200.* exist with inferred schemas, are traceable through their fake implementations, and have no CPU
    kernel: a CPU tensor is refused by the dispatcher (there is no fallback path to fall into)."""
    import ssrl_vit_mae_jepa_b200  # noqa: F401
    ns = torch.ops.mae_b200
    for name in ("random_token_mask", "gemm_tn", "gemm_wgrad", "patch_embed_tokens", "mlp_fwd", "linear", "layernorm",
                 "layernorm_bwd", "attention", "attention_bwd", "masked_mse"):
        assert hasattr(ns, name), name
    assert "Tensor x, Tensor weight, Tensor bias" in str(ns.linear.default._schema)
    with pytest.raises(NotImplementedError):
        ns.linear(torch.zeros(4, 16, dtype=torch.bfloat16), torch.zeros(16, 16, dtype=torch.bfloat16), torch.zeros(16))
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(256, 144, dtype=torch.bfloat16, device="cuda")
        y = ns.linear(x, torch.empty(432, 144, dtype=torch.bfloat16, device="cuda"), torch.empty(432, device="cuda"))
        assert y.shape == (256, 432) and y.dtype == torch.bfloat16
        o, lse = ns.attention(y.view(8, 32, 432), 6)
        assert o.shape == (8, 32, 144) and lse.shape == (8, 6, 32)

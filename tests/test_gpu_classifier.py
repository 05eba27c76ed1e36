"""-m gpu: BASELINE config #4 -- the downstream classifier (frozen-encoder probe, unfreeze-last-N, full fine-tune)
through the product classes against the fp32 CPU oracle.  Tolerances as in test_gpu_e2e.py (bf16 GEMMs vs fp32)."""
import pytest
import torch
import torch.nn.functional as F

import ssrl_vit_mae_jepa_b200 as P
from oracle import mae_oracle as O
from tests.gpu_util import cos_sim, max_err, rel_err, report

pytestmark = pytest.mark.gpu
torch.set_float32_matmul_precision("highest")
CFG = {"general": O.MAE_YAML["general"], "encoder": O.MAE_YAML["encoder"], "decoder": O.MAE_YAML["decoder"],
       "head": {"embed_dim": 144, "pool": "cls"}}


def _pair(pool="cls", freeze=False, seed=21):
    mae_o = O.build_mae(seed=seed)
    oc = O.ViTClassifier(mae_o.encoder.vit, 10, pool)
    mae = P.MaskedAutoencoder(CFG["general"], CFG["encoder"], CFG["decoder"])
    mae.load_state_dict(mae_o.state_dict())
    cfg = dict(CFG, head={"embed_dim": 144, "pool": pool})
    mod = P.ViTClassifierTrainModule(pretrained_encoder=mae.encoder.vit, model_cfg=cfg,
                                     training_cfg={"freeze_encoder": freeze}).to("cuda")
    mod.model.head.load_state_dict(oc.head.state_dict())
    return oc, mod


def _data(B=24, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, 96, 96, generator=g), torch.randint(0, 10, (B,), generator=g)


def _check(mod, oc, names, tag):
    worst = 0.0
    for n in names:
        a = dict(mod.model.named_parameters())[n].grad
        b = dict(oc.named_parameters())[n].grad
        assert a is not None and b is not None, n
        c, r = cos_sim(a.cpu(), b), rel_err(a.cpu(), b)
        worst = max(worst, r)
        if b.norm().item() > 1e-4:
            assert c > 0.999 and r < 3e-2, (n, c, r)
    report(tag, tensors=len(names), worst_rel=worst)


def test_forward_features_and_full_finetune():
    oc, mod = _pair(freeze=False)
    for p_ in oc.parameters():  # reference unfreeze_encoder() (src/training/classifier.py:134-137) unfreezes EVERYTHING,
        p_.requires_grad = True  # including the sin-cos pos_embed the MAE wrapper had frozen
    x, y = _data()
    feats = mod.model.encoder.forward_features(x.cuda())
    with torch.no_grad():
        f_ref = oc.encoder.forward_features(x)
    assert feats.shape == (24, 145, 144) and max_err(feats.cpu(), f_ref) < 5e-2 and rel_err(feats.cpu(), f_ref) < 1e-2
    loss = mod.training_step((x.cuda(), y.cuda()), 0)
    loss.backward()
    l_ref = F.cross_entropy(oc(x), y)
    l_ref.backward()
    assert abs(loss.item() - l_ref.item()) < 3e-3 * l_ref.item()
    assert abs(mod.logged["train_acc"].item() - (oc(x).argmax(1) == y).float().mean().item()) < 0.1
    names = [n for n, p in oc.named_parameters() if p.requires_grad and p.grad is not None]
    assert "encoder.pos_embed" in names and len(names) == 56
    _check(mod, oc, names, "cls_full_ft")


def test_frozen_probe_and_unfreeze_last_layers():
    oc, mod = _pair(freeze=True)
    x, y = _data(seed=6)
    assert sum(p.numel() for p in mod.parameters() if p.requires_grad) == 1450  # notebook.ipynb:1693-1698
    loss = mod.training_step((x.cuda(), y.cuda()), 0)
    loss.backward()
    for p in oc.encoder.parameters():
        p.requires_grad = False
    l_ref = F.cross_entropy(oc(x), y)
    l_ref.backward()
    assert abs(loss.item() - l_ref.item()) < 3e-3 * l_ref.item()
    _check(mod, oc, ["head.classification.weight", "head.classification.bias"], "cls_probe")
    assert all(p.grad is None for n, p in mod.model.named_parameters() if "head" not in n)
    # unfreeze the last two blocks (+ final norm + head), reference src/training/classifier.py:139-171
    mod.unfreeze_last_layers(2)
    for p in mod.parameters():
        p.grad = None
    for n, p in oc.named_parameters():
        p.grad = None
        p.requires_grad = n.startswith(("encoder.blocks.2.", "encoder.blocks.3.", "encoder.norm.", "head."))
    mod.training_step((x.cuda(), y.cuda()), 0).backward()
    F.cross_entropy(oc(x), y).backward()
    names = [n for n, p in oc.named_parameters() if p.requires_grad]
    _check(mod, oc, names, "cls_last2")
    assert dict(mod.model.named_parameters())["encoder.blocks.1.attn.qkv.weight"].grad is None


def test_mean_pool_random_init_vit_with_trainable_pos_embed():
    torch.manual_seed(0)
    cfg = dict(CFG, head={"embed_dim": 144, "pool": "mean"})
    mod = P.ViTClassifierTrainModule(model_cfg=cfg, training_cfg={"freeze_encoder": False}).to("cuda")
    ov = O.VisionTransformer()
    ov.load_state_dict(mod.model.encoder.state_dict())
    oc = O.ViTClassifier(ov, 10, "mean")
    oc.head.load_state_dict({k: v.cpu() for k, v in mod.model.head.state_dict().items()})
    x, y = _data(seed=7)
    loss = mod.training_step((x.cuda(), y.cuda()), 0)
    loss.backward()
    l_ref = F.cross_entropy(oc(x), y)
    l_ref.backward()
    assert abs(loss.item() - l_ref.item()) < 3e-3 * l_ref.item()
    _check(mod, oc, [n for n, _ in oc.named_parameters()], "cls_mean_random_vit")
    # one optimizer step through torch AdamW on the flat-buffer views, then the kernels see the new weights
    opt = mod.configure_optimizers()["optimizer"]
    opt.step()
    l2 = mod.training_step((x.cuda(), y.cuda()), 0)
    assert torch.isfinite(l2) and l2.item() != loss.item()


def test_torch_optimizer_updates_reach_the_kernels():
    """Weights changed through torch (AdamW steps on the flat-buffer views, then ``load_state_dict``) must be what the
    kernels compute with: the bf16 W / W^T shadows are rebuilt unless the fused optimizer just refreshed them.  The head
    stays frozen, so every change of the logits comes from the encoder weights; the oracle takes the same AdamW steps."""
    oc, mod = _pair(freeze=False)
    for p_ in oc.parameters():
        p_.requires_grad = True
    for p_ in list(mod.model.head.parameters()) + list(oc.head.parameters()):
        p_.requires_grad = False
    x, y = _data(seed=8)
    xc, yc = x.cuda(), y.cuda()
    opt = torch.optim.AdamW([p_ for p_ in mod.model.parameters() if p_.requires_grad], lr=2e-3, weight_decay=0.0)
    opt_o = torch.optim.AdamW([p_ for p_ in oc.parameters() if p_.requires_grad], lr=2e-3, weight_decay=0.0)
    with torch.no_grad():
        f0 = mod.model.encoder.forward_features(xc).clone()
    for _ in range(4):
        opt.zero_grad(set_to_none=True); opt_o.zero_grad(set_to_none=True)
        mod.training_step((xc, yc), 0).backward()
        F.cross_entropy(oc(x), y).backward()
        opt.step(); opt_o.step()
    with torch.no_grad():
        f1 = mod.model.encoder.forward_features(xc).clone()
        f_ref = oc.encoder.forward_features(x)
    moved = rel_err(f1.cpu(), f0.cpu())
    report("cls_torch_adamw", moved=moved, rel_vs_oracle=rel_err(f1.cpu(), f_ref))
    assert moved > 5e-2                                   # 4 steps at lr 2e-3 move the features a lot ...
    assert rel_err(f1.cpu(), f_ref) < 0.5 * moved         # ... and the kernels follow the oracle, not the step-0 weights
    # load_state_dict after a forward: the next forward must use the loaded weights
    sd0 = {k: v.clone() for k, v in oc.encoder.state_dict().items()}
    mod.model.encoder.load_state_dict({k: torch.zeros_like(v) if k == "blocks.3.mlp.fc2.weight" else v for k, v in sd0.items()})
    with torch.no_grad():
        f2 = mod.model.encoder.forward_features(xc)
    assert rel_err(f2.cpu(), f1.cpu()) > 1e-3

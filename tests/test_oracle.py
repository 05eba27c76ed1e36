"""The oracle against everything that pins it: notebook model-summary integers, the committed
golden vectors, hand-computed known answers, and an independent implementation (HF ViTMAE)."""
import numpy as np
import pytest
import torch

from oracle import mae_oracle as O

torch.set_float32_matmul_precision("highest")


@pytest.fixture(scope="module")
def model():
    return O.build_mae(seed=73)


def test_param_and_module_counts_match_notebook_pins(model):
    # notebook.ipynb:986-995 -> 2.0 M trainable, 48.7 K non-trainable, 152 modules (151 + MSELoss)
    total = sum(p.numel() for p in model.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert total == 2_035_104 and trainable == 1_986_384 and total - trainable == 48_720
    assert sum(p.numel() for p in model.encoder.parameters()) == 1_052_064
    assert sum(p.numel() for p in model.decoder.parameters()) == 983_040
    assert len(list(model.modules())) == 151
    # notebook.ipynb:1689-1698 -> bare timm ViT: 100 modules, 1,051,920 params
    assert len(list(model.encoder.vit.modules())) == 100
    assert sum(p.numel() for p in model.encoder.vit.parameters()) == 1_051_920


def test_state_dict_layout(model):
    sd = model.state_dict()
    assert len(sd) == 87
    keys = list(sd)
    assert keys[0] == "encoder.mask_token" and keys[1] == "encoder.vit.cls_token" and keys[2] == "encoder.vit.pos_embed"
    assert sd["encoder.vit.patch_embed.proj.weight"].shape == (144, 3, 8, 8)
    assert sd["encoder.vit.blocks.3.attn.qkv.weight"].shape == (432, 144)
    assert sd["decoder.decoder_blocks.1.mlp.fc2.weight"].shape == (192, 768)
    assert sd["decoder.decoder_pos_embed"].shape == (1, 145, 192)
    assert sd["decoder.decoder_pred.weight"].shape == (192, 192)
    assert not any("head" in k for k in keys if k.startswith("encoder"))


def test_patchify_known_answer():
    # 1x3x16x16 ramp, patch 8: element (n, c, y, x) = c*1000 + y*16 + x
    c, y, x = torch.meshgrid(torch.arange(3), torch.arange(16), torch.arange(16), indexing="ij")
    img = (c * 1000 + y * 16 + x).float().unsqueeze(0)
    p = O.patchify(img, 8)
    assert p.shape == (1, 4, 192)
    # patch 3 = (row 1, col 1); element (py=2, px=5, c=1) sits at (py*8+px)*3+c
    assert p[0, 3, (2 * 8 + 5) * 3 + 1].item() == 1000 + (8 + 2) * 16 + (8 + 5)
    assert p[0, 1, 0].item() == 8 and p[0, 2, 0].item() == 8 * 16 and p[0, 0, 2].item() == 2000


def test_sincos_known_values():
    pe = O.sincos_2d_pos_embed(144, 12)[0]
    assert pe.shape == (145, 144) and pe[0].abs().max() == 0
    np.testing.assert_allclose(pe[2, :4].numpy(), [0.8415, 0.6992, 0.5642, 0.4477], atol=1e-4)
    assert pe[13, :4].abs().max() < 1e-7  # row 1 / col 0: column half encodes 0
    np.testing.assert_allclose(pe[13, 72:76].numpy(), [0.8415, 0.6992, 0.5642, 0.4477], atol=1e-4)


def test_random_token_mask_invariants_and_ties(golden):
    noise = torch.from_numpy(golden["noise"])
    k, m = O.random_token_mask_from_noise(noise, 0.75)
    assert k.shape == (4, 36) and m.shape == (4, 109) and k.dtype == torch.int64
    assert (k[:, 0] == 0).all() and (m > 0).all()
    assert (torch.cat([k, m], 1).sort(1).values == torch.arange(145)).all()
    assert np.array_equal(k.numpy(), golden["idx_keep"]) and np.array_equal(m.numpy(), golden["idx_mask"])
    kn, mn = O.random_token_mask_numpy(golden["noise"], 0.75)
    assert np.array_equal(kn, golden["idx_keep"]) and np.array_equal(mn, golden["idx_mask"])
    tk, tm = O.random_token_mask_from_noise(torch.from_numpy(golden["tie_noise"]), 0.75)
    assert np.array_equal(tk.numpy(), golden["tie_idx_keep"]) and np.array_equal(tm.numpy(), golden["tie_idx_mask"])
    assert tk[1].tolist() == list(range(36))  # all-equal noise -> identity order (stable)
    assert O.num_keep_tokens(145, 0.75) == 36 and O.num_keep_tokens(197, 0.75) == 49 and O.num_keep_tokens(5, 0.99) == 1


def test_golden_forward_backward(model, golden):
    w = sum(v.double().abs().sum().item() for v in model.state_dict().values())
    assert abs(w - float(golden["weight_checksum"])) < 1e-6 * w, "seeded init drifted; regenerate goldens"
    images, noise = torch.from_numpy(golden["images"]), torch.from_numpy(golden["noise"])
    loss, pred, target, grads = O.loss_and_grads(model, images, noise)
    assert abs(loss.item() - float(golden["loss"])) < 1e-5
    np.testing.assert_allclose(pred.numpy(), golden["pred"], atol=2e-5)
    assert abs(target.double().abs().sum().item() - float(golden["target_checksum"])) < 1e-3
    assert "encoder.mask_token" not in grads and "encoder.vit.pos_embed" not in grads
    assert len(grads) == 84
    for k in golden.files:
        if k.startswith("grad:"):
            np.testing.assert_allclose(grads[k[5:]].numpy(), golden[k], atol=1e-6, rtol=1e-4)


def test_cross_check_against_hf_vitmae(model, golden):
    pytest.importorskip("transformers")
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rel, err = mg.cross_check(model, torch.from_numpy(golden["images"]), torch.from_numpy(golden["noise"]))
    assert rel < 1e-6 and err < 2e-5


def test_schedule_and_clip_adamw_restatement(model):
    assert O.mask_ratio_at_epoch(0, 0.5, 0.85, 200) == 0.5 and O.mask_ratio_at_epoch(500, 0.5, 0.85, 200) == 0.85
    assert O.mask_ratio_at_epoch(3, 0.75, 0.75, 5) == 0.75
    assert abs(O.lr_lambda(0, 20, 800) - 1 / 20) < 1e-12 and abs(O.lr_lambda(400, 20, 800) - 0.5) < 1e-12
    # explicit clip+AdamW restatement == torch.optim.AdamW + clip_grad_norm_
    torch.manual_seed(0)
    ps = [torch.randn(7, 5), torch.randn(11)]
    gs = [torch.randn(7, 5) * 3, torch.randn(11) * 3]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    mine, m, v = [p.clone() for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    for step in (1, 2, 3):
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        O.clip_adamw_reference(mine, gs, m, v, step, lr=1e-2)
    for a, b in zip(mine, ref):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), atol=1e-6)

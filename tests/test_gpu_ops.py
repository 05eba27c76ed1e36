"""-m gpu: the torch.library surface (torch.ops.mae_b200.*) -- forward values and autograd formulas of the differentiable
building blocks against plain fp32 torch on the same inputs (bf16 tensor-core tolerances as in test_gpu_kernels.py), a
small module composed of the ops trained through loss.backward(), and torch.library.opcheck on the registrations."""
import math

import pytest
import torch
import torch.nn.functional as F

import ssrl_vit_mae_jepa_b200  # noqa: F401  (registers the ops)
from oracle import mae_oracle as O
from tests.gpu_util import bf, cos_sim, max_err, rel_err, report

pytestmark = pytest.mark.gpu
dev = "cuda"
ops = torch.ops.mae_b200


def test_linear_layernorm_attention_autograd():
    g = torch.Generator().manual_seed(0)
    B, L, D, H = 6, 36, 144, 6
    x = torch.randn(B * L, D, generator=g).to(dev).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).to(dev).requires_grad_(True)
    beta = (0.1 * torch.randn(D, generator=g)).to(dev).requires_grad_(True)
    w = bf(torch.randn(3 * D, D, generator=g) / math.sqrt(D)).to(dev).requires_grad_(True)
    b = (0.1 * torch.randn(3 * D, generator=g)).to(dev).requires_grad_(True)
    xn, mean, rstd = ops.layernorm(x, gamma, beta, 1e-6)
    qkv = ops.linear(xn, w, b)
    out, lse = ops.attention(qkv.view(B, L, 3 * D), H)
    tgt = torch.randn(B, L, D, generator=g).to(dev)
    loss = ((out.float() - tgt) ** 2).mean()
    loss.backward()
    # fp32 torch reference of the same chain
    xr, gr, br = x.detach().clone().requires_grad_(True), gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    wr, b_r = w.detach().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    qkv_r = F.linear(F.layer_norm(xr, (D,), gr, br, 1e-6), wr, b_r).view(B, L, 3, H, D // H).permute(2, 0, 3, 1, 4)
    out_r = F.scaled_dot_product_attention(qkv_r[0], qkv_r[1], qkv_r[2]).transpose(1, 2).reshape(B, L, D)
    loss_r = ((out_r - tgt) ** 2).mean()
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) < 5e-3 * loss_r.item() and max_err(out, out_r) < 3e-2
    for name, a, r in (("dx", x.grad, xr.grad), ("dgamma", gamma.grad, gr.grad), ("dbeta", beta.grad, br.grad),
                       ("dw", w.grad, wr.grad), ("db", b.grad, b_r.grad)):
        c, e = cos_sim(a, r), rel_err(a, r)
        report("ops_autograd", tensor=name, cos=c, rel=e)
        assert c > 0.999 and e < 3e-2, (name, c, e)


def test_masked_mse_and_mask_ops():
    g = torch.Generator().manual_seed(1)
    B = 5
    images, noise = torch.randn(B, 3, 96, 96, generator=g).to(dev), torch.rand(B, 145, generator=g).to(dev)
    keep, mask = ops.random_token_mask(noise, 36)
    k_ref, m_ref = O.random_token_mask_from_noise(noise.cpu(), 0.75)
    assert torch.equal(keep.cpu(), k_ref) and torch.equal(mask.cpu(), m_ref)
    pred = bf(torch.randn(B, 109, 192, generator=g)).to(dev).requires_grad_(True)
    loss, _ = ops.masked_mse(pred, images, mask, 8)
    (3.0 * loss).backward()
    target = O.get_at_index(O.patchify(images, 8), mask - 1)
    pr = pred.detach().float().requires_grad_(True)
    lr = F.mse_loss(pr, target)
    (3.0 * lr).backward()
    assert abs(loss.item() - lr.item()) < 1e-5 * lr.item()
    assert rel_err(pred.grad, pr.grad) < 1e-2  # bf16 gradient tensor


def test_patch_embed_and_fused_mlp_ops_match_the_module_path():
    g = torch.Generator().manual_seed(2)
    o = O.build_mae(seed=3)
    vit = o.encoder.vit
    B = 4
    images, noise = torch.randn(B, 3, 96, 96, generator=g).to(dev), torch.rand(B, 145, generator=g).to(dev)
    keep, _ = ops.random_token_mask(noise, 36)
    x, patches = ops.patch_embed_tokens(images, keep, bf(vit.patch_embed.proj.weight.detach()).to(dev),
                                        vit.patch_embed.proj.bias.detach().to(dev), vit.cls_token.detach().to(dev),
                                        vit.pos_embed.detach().to(dev), 8)
    with torch.no_grad():
        ref = O.get_at_index(o.encoder.preprocess(images.cpu()), keep.cpu())
    assert x.shape == (B, 36, 144) and patches.shape == (B * 35, 192) and max_err(x.cpu(), ref) < 3e-2
    blk = vit.blocks[0]
    y, xn, mean, rstd, h, dg = ops.mlp_fwd(x.view(-1, 144), blk.norm2.weight.detach().to(dev), blk.norm2.bias.detach().to(dev),
                                           bf(blk.mlp.fc1.weight.detach()).to(dev), blk.mlp.fc1.bias.detach().to(dev),
                                           bf(blk.mlp.fc2.weight.detach()).to(dev), blk.mlp.fc2.bias.detach().to(dev), 1e-6)
    with torch.no_grad():
        xc = x.view(-1, 144).cpu()
        y_ref = xc + blk.mlp(blk.norm2(xc))
    assert max_err(y.cpu(), y_ref) < 3e-2 and rel_err(y.cpu(), y_ref) < 5e-3


def test_cpu_tensors_are_refused_and_opcheck():
    with pytest.raises(NotImplementedError):
        ops.layernorm(torch.zeros(4, 16), torch.ones(16), torch.zeros(16), 1e-6)
    x = bf(torch.randn(256, 144)).to(dev).requires_grad_(True)
    w = bf(torch.randn(144, 144) / 12).to(dev).requires_grad_(True)
    b = torch.zeros(144, device=dev, requires_grad=True)
    torch.library.opcheck(ops.linear.default, (x, w, b), test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(ops.random_token_mask.default, (torch.rand(8, 145, device=dev), 36),
                          test_utils=("test_schema", "test_faketensor"))

"""GPU: the head's training path (setok_b200/training.py over csrc/train.cu + the tcgen05 GEMM) against
torch.autograd of the fp32 oracle on the CPU.  Forward values and every parameter gradient are compared.

Tolerances: forward and backward contractions take bf16 operands with f32 accumulation and activation gradients hop between
kernels in bf16, so gradients are asserted to 3e-2 relative Frobenius against the fp32 oracle (measured ~5e-3 to 1.5e-2);
single row kernels against torch on identical inputs: 1e-5 (f32) / one bf16 ulp."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from oracle import setok_oracle as O  # noqa: E402
from setok_b200 import SetokTokenizer, build_vision_projector, ops, training  # noqa: E402

DEV = torch.device("cuda:0")


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_row_kernels_vs_torch():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 200, generator=g)
    t = training.transpose_to_bf16(x.to(DEV))
    assert t.shape == (200, 300) and torch.equal(t.float().cpu(), x.t().to(torch.bfloat16).float())
    xb = torch.randn(3, 70, 52, generator=g).to(torch.bfloat16)
    assert torch.equal(training.transpose_to_bf16(xb.to(DEV)).cpu(), xb.transpose(1, 2))
    assert _rel(training.colsum(x.to(DEV)), x.sum(0)) < 1e-5
    # LayerNorm backward
    C = 192
    xx = (torch.randn(77, C, generator=g) * 2 + 0.5).requires_grad_()
    gam, bet = (1 + 0.1 * torch.randn(C, generator=g)).requires_grad_(), (0.1 * torch.randn(C, generator=g)).requires_grad_()
    dy = torch.randn(77, C, generator=g)
    torch.nn.functional.layer_norm(xx, (C,), gam, bet, 1e-5).backward(dy)
    xd, gd, bd = xx.detach().to(DEV).requires_grad_(), gam.detach().to(DEV).requires_grad_(), bet.detach().to(DEV).requires_grad_()
    training.LayerNormFn.apply(xd, gd, bd, 1e-5).backward(dy.to(DEV, torch.bfloat16))
    dyb = dy.to(torch.bfloat16).float()
    xr = xx.detach().clone().requires_grad_()
    g2, b2 = gam.detach().clone().requires_grad_(), bet.detach().clone().requires_grad_()
    torch.nn.functional.layer_norm(xr, (C,), g2, b2, 1e-5).backward(dyb)
    assert _rel(xd.grad, xr.grad) < 1e-4 and _rel(gd.grad, g2.grad) < 1e-4 and _rel(bd.grad, b2.grad) < 1e-4
    # GELU
    pre = torch.randn(50, 64, generator=g).requires_grad_()
    d2 = torch.randn(50, 64, generator=g)
    torch.nn.functional.gelu(pre).backward(d2)
    pd = pre.detach().to(DEV).requires_grad_()
    act = training.GeluFn.apply(pd)
    assert torch.equal(act.float().cpu(), torch.nn.functional.gelu(pre.detach()).to(torch.bfloat16).float())
    act.backward(d2.to(DEV, torch.bfloat16))
    pr = pre.detach().clone().requires_grad_()
    torch.nn.functional.gelu(pr).backward(d2.to(torch.bfloat16).float())
    assert _rel(pd.grad, pr.grad) < 1e-5


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (1000, 192, 320), (37, 64, 64)])
def test_linear_fn_vs_torch(M, N, K):
    g = torch.Generator().manual_seed(M)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * K ** -0.5, torch.randn(N, generator=g)
    res, dy = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    xb, wb, dyb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), dy.to(torch.bfloat16).float()
    xd, wd, bd, rd = [t.to(DEV).requires_grad_() for t in (x, w, b, res)]
    y = training.LinearFn.apply(xd, wd, bd, rd)
    assert _rel(y, xb @ wb.t() + b + res) < 1e-5
    y.backward(dy.to(DEV))
    assert _rel(xd.grad, dyb @ wb) < 1e-5 and _rel(wd.grad, dyb.t() @ xb) < 1e-4
    assert _rel(bd.grad, dy.sum(0)) < 1e-5 and torch.equal(rd.grad.cpu(), dy)


def _tok(C, Ctok, Fd, mcn, hp):
    cfg = dict(hidden_size=C, intermediate_size=2 * C, num_hidden_layers=1, num_attention_heads=max(1, C // 64), image_size=32, patch_size=4)
    tok = SetokTokenizer("siglip-synthetic", hidden_dim=C, token_feat_dim=Ctok, min_cluster_num=mcn, threshold=0.5, dim_feedforward=Fd, vision_config=cfg)
    tok.load_state_dict(hp, strict=False)
    return tok.to(DEV)


@pytest.mark.parametrize("N,C,Ctok,B", [(64, 128, 64, 3), (256, 256, 128, 2)])
def test_head_training_path_vs_oracle_autograd(N, C, Ctok, B):
    """Forward tokens and d(loss)/d(every head parameter) against torch.autograd of the oracle's per-image head loop (the
    reference's computation, tokenizer.py:123-155, 178-180), feature-injected mixtures so that K is dynamic."""
    mcn, k, Fd = 16, 8, 256
    hp = O.make_head_params(C, Ctok, Fd, seed=N)
    pairs = [O.well_posed_features(N, C, 5 + 4 * b, k, mcn, 0.5, seed0=700 + 100 * b) for b in range(B)]
    feats, noise = torch.stack([p_[0] for p_ in pairs]), torch.stack([p_[1] for p_ in pairs])
    # oracle with autograd
    hp_g = {n_: v.clone().requires_grad_(v.dtype.is_floating_point) for n_, v in hp.items()}
    toks_ref, w_loss = [], []
    gw = torch.Generator().manual_seed(5)
    loss_ref = 0.0
    for b in range(B):
        t = O.tokenizer_head(feats[b], noise[b], hp_g, min_cluster_num=mcn, threshold=0.5, k=k)[0]
        wl = torch.randn(t.shape, generator=gw)
        toks_ref.append(t); w_loss.append(wl)
        loss_ref = loss_ref + (t * wl).sum()
    loss_ref.backward()
    # ours
    tok = _tok(C, Ctok, Fd, mcn, hp)
    tok.train()
    h = int(N ** 0.5)
    pos = tok.position_embedding.table(h, h, DEV)
    x_pos = (feats.to(DEV) + pos[None]).contiguous()
    _, idx, score, down, numc, offs = ops.dpc_cluster(x_pos, noise.to(DEV), (h, h), k, 0.5, mcn, embedded=True)
    rt = training.head_forward_train(tok, x_pos, idx, numc, offs)
    assert rt.data.requires_grad and rt.counts == [t.shape[0] for t in toks_ref]
    for b in range(B):
        assert _rel(rt[b], toks_ref[b]) < 1e-2, b
    loss = (rt.packed() * torch.cat(w_loss, 0).to(DEV)).sum()
    loss.backward()
    checked = 0
    for name, p in tok.named_parameters():
        if name.startswith("image_feature_encoder") or ".layers." in name and ".0." in name:
            continue                                              # the tower is frozen; layers.i.0 aliases norm1 (same Parameter object)
        ref = hp_g[name].grad
        assert p.grad is not None and ref is not None, name
        r = _rel(p.grad, ref)
        assert r < 3e-2, (name, r)
        checked += 1
    assert checked >= 20


def test_tokenizer_forward_dispatches_to_training_path():
    C, L, H, P, IMG, B = 128, 2, 2, 4, 32, 3
    cfg = dict(hidden_size=C, intermediate_size=4 * C, num_hidden_layers=L, num_attention_heads=H, image_size=IMG, patch_size=P)
    torch.manual_seed(3)
    tok = SetokTokenizer("siglip-synthetic", hidden_dim=C, token_feat_dim=64, min_cluster_num=8, threshold=0.5, dim_feedforward=256, vision_config=cfg).to(DEV)
    imgs = torch.randn(B, 3, IMG, IMG, device=DEV)
    noise = torch.rand(B, (IMG // P) ** 2, device=DEV)
    tok.eval()
    rt_e, idx_e, _ = tok(imgs, k=8, noise=noise)
    assert not rt_e.data.requires_grad
    tok.train()
    rt_t, idx_t, _ = tok(imgs, k=8, noise=noise)
    assert rt_t.data.requires_grad and torch.equal(idx_t, idx_e) and rt_t.counts == rt_e.counts
    assert _rel(rt_t.packed(), rt_e.packed()) < 1e-2                 # same computation, unfused kernels
    rt_t.packed().square().mean().backward()
    assert tok.out.weight.grad is not None and float(tok.out.weight.grad.abs().sum()) > 0
    assert tok.inner_encoder.layers[0][1].qkv.weight.grad is not None
    assert all(p.grad is None for p in tok.image_feature_encoder.parameters())      # frozen tower (clip_encoder.py:36, :50)
    with torch.no_grad():
        assert not tok(imgs, k=8, noise=noise)[0].data.requires_grad


@pytest.mark.parametrize("kind", ["mlp2x_gelu", "linear", "mlp2x_gelu_Norm"])
def test_projector_training_path(kind):
    pp = O.make_projector_params(64, 128, kind, seed=9)
    proj = build_vision_projector(kind, mm_hidden_size=64, hidden_size=128)
    proj.load_state_dict(pp)
    proj = proj.to(DEV)
    g = torch.Generator().manual_seed(1)
    x, wl = torch.randn(45, 64, generator=g), torch.randn(45, 128, generator=g)
    pg = {n_: v.clone().requires_grad_() for n_, v in pp.items()}
    (O.projector(x, pg, kind) * wl).sum().backward()
    y = training.projector_forward_train(proj, x.to(DEV))
    assert _rel(y, O.projector(x, pp, kind)) < 1e-2
    (y * wl.to(DEV)).sum().backward()
    for name, p in proj.named_parameters():
        assert _rel(p.grad, pg[name].grad) < 3e-2, name

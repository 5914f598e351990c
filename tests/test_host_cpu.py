"""CPU: host-side logic — the reference's plugin surface (ctor kwargs, state_dict keys, factory errors),
the ragged container, sharding helpers."""
import dataclasses

import pytest
import torch

import setok_b200
from setok_b200 import RaggedTokens, SetokTokenizer, build_vision_projector, build_vision_tower
from setok_b200.dist import shard_batch
from oracle import ref_loader

VC = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=16, patch_size=4)


def test_ctor_kwargs_and_extras_are_accepted():
    tok = SetokTokenizer("siglip-x", hidden_dim=128, token_feat_dim=64, min_cluster_num=4, threshold=0.55, nheads=2,
                         dim_feedforward=256, proj_drop=0.2, drop_path=0.0, inner_cluster_layers=1, intra_cluster_layers=3,
                         attn_drop=0.0, pretrain_vision_tokenizer="", some_future_flag=1, vision_config=VC)
    assert tok.inner_encoder.depth == 1 and tok.inter_encoder.depth == 3 and tok.threshold == 0.55
    assert tok.is_loaded and tok.hidden_dim == 128 and tok.token_feat_dim == 64


def test_state_dict_keys_match_reference_layout():
    tok = SetokTokenizer("siglip-x", hidden_dim=128, token_feat_dim=64, min_cluster_num=4, dim_feedforward=256, vision_config=VC)
    ours = {k for k in tok.state_dict() if not k.startswith("image_feature_encoder")}
    expect = {"position_embedding.inv_freq", "out.weight", "out.bias"}
    for enc in ("inner_encoder", "inter_encoder"):
        for n in ("norm1", "norm2"):
            expect |= {f"{enc}.{n}.weight", f"{enc}.{n}.bias"}
        for i in range(2):
            expect |= {f"{enc}.layers.{i}.0.weight", f"{enc}.layers.{i}.0.bias"}          # alias of norm1 (module.py:88)
            for n in ("qkv", "proj"):
                expect |= {f"{enc}.layers.{i}.1.{n}.weight", f"{enc}.layers.{i}.1.{n}.bias"}
        for n in ("fc1", "fc2"):
            expect |= {f"{enc}.mlp.{n}.weight", f"{enc}.mlp.{n}.bias"}
    assert ours == expect
    assert tok.inner_encoder.layers[0][0] is tok.inner_encoder.norm1
    # xavier init / zero bias as tokenizer.py:59-72
    assert float(tok.out.bias.abs().max()) == 0.0 and float(tok.inner_encoder.norm1.weight.min()) == 1.0


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_state_dict_keys_equal_the_live_reference():
    ns = ref_loader.load_reference()
    ref = ref_loader.build_reference_tokenizer(ns, None, hidden_dim=128, token_feat_dim=64, min_cluster_num=4, threshold=0.5, dim_feedforward=256)
    tok = SetokTokenizer("siglip-x", hidden_dim=128, token_feat_dim=64, min_cluster_num=4, dim_feedforward=256, vision_config=VC)
    rk = {k: tuple(v.shape) for k, v in ref.state_dict().items() if not k.startswith("image_feature_encoder")}
    ok = {k: tuple(v.shape) for k, v in tok.state_dict().items() if not k.startswith("image_feature_encoder")}
    assert rk == ok
    tok.load_state_dict(ref.state_dict(), strict=False)
    pr = ns.projector_builder.build_vision_projector("mlp2x_gelu_Norm", mm_hidden_size=64, hidden_size=128)
    po = build_vision_projector("mlp2x_gelu_Norm", mm_hidden_size=64, hidden_size=128)
    assert {k: tuple(v.shape) for k, v in pr.state_dict().items()} == {k: tuple(v.shape) for k, v in po.state_dict().items()}


def test_factory_accepts_dataclass_dict_namespace_and_rejects_non_siglip():
    @dataclasses.dataclass
    class Args:
        vision_tower: str = "google/siglip-so400m-patch14-384"
        hidden_dim: int = 128
        token_feat_dim: int = 64
        min_cluster_num: int = 4
        dim_feedforward: int = 256
        pretrain_vision_tokenizer: str = ""
    t1 = build_vision_tower(Args(), vision_config=VC)
    t2 = build_vision_tower(dataclasses.asdict(Args()), vision_config=VC)
    assert isinstance(t1, SetokTokenizer) and isinstance(t2, SetokTokenizer)
    with pytest.raises(ValueError, match="Unknown vision tower"):
        build_vision_tower(Args(vision_tower="openai/clip-vit-large-patch14"), vision_config=VC)
    with pytest.raises(ValueError, match="Unknown projector type"):
        build_vision_projector("resampler")


def test_ragged_tokens_container():
    data = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    rt = RaggedTokens(data, torch.tensor([0, 3, 3, 7], dtype=torch.int32))
    assert len(rt) == 3 and rt.dim() == 3 and rt.counts == [3, 0, 4] and rt.total == 7
    assert torch.equal(rt[0], data[:3]) and rt[1].shape == (0, 4) and torch.equal(rt[-1], data[3:7])
    assert torch.equal(rt.packed(), data[:7])
    pad, mask = rt.to_padded()
    assert pad.shape == (3, 4, 4) and mask.sum().item() == 7 and torch.equal(pad[2], data[3:7])
    assert [t.shape[0] for t in rt] == [3, 0, 4]
    with pytest.raises(IndexError):
        rt[3]


def test_shard_batch_covers_everything_once():
    for n in (1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_u8_normalisation_constants_reproduce_the_hf_processor():
    """The uint8 tower path (setok_vit_forward_u8) applies lut[u8] -> (x - mean) / std in float32 with separately rounded
    steps.  Evaluated in numpy with the constants `CLIPVisionTower.u8_norm()` hands to the kernel, that arithmetic must
    reproduce, bit for bit, `transformers.image_transforms.rescale` + `normalize` -- the numpy functions the reference's
    pinned transformers==4.46.3 CLIPImageProcessor.preprocess calls (image_processing_clip.py: self.rescale / self.normalize).
    The torchvision-backed processor class of the installed transformers 5.x fuses the two steps and lands within 1 ulp."""
    import numpy as np
    from transformers import CLIPImageProcessor
    from transformers.image_transforms import normalize, rescale
    from setok_b200 import CLIPVisionTower
    cfg = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, image_size=16, patch_size=4)
    tower = CLIPVisionTower("siglip-synthetic", vision_config=cfg)
    n = tower.u8_norm()
    lut = np.array(list(n.lut), dtype=np.float32)
    mean = np.array(list(n.mean), dtype=np.float32)
    std = np.array(list(n.std), dtype=np.float32)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(3, 16, 16, 3), dtype=np.uint8)                 # HWC, as PIL / numpy images arrive
    chw = np.transpose(img, (0, 3, 1, 2))
    ours = (lut[chw] - mean[None, :, None, None]) / std[None, :, None, None]
    assert ours.dtype == np.float32
    ip = CLIPImageProcessor(do_resize=False, do_center_crop=False, do_convert_rgb=False)
    for i in range(img.shape[0]):
        x = rescale(img[i], scale=ip.rescale_factor, input_data_format="channels_last")
        y = normalize(x, mean=ip.image_mean, std=ip.image_std, input_data_format="channels_last")
        assert y.dtype == np.float32 and np.array_equal(ours[i], np.transpose(y, (2, 0, 1)))
    fast = ip.preprocess(list(img), return_tensors="np", input_data_format="channels_last")["pixel_values"]
    assert np.abs(ours - fast).max() <= 2.4e-7 * np.abs(fast).max()


def test_layernorm_fold_algebra_on_cpu():
    """The host half of SETOK_VIT_LN_FOLD (setok_b200/_pack.py:fold_layernorm_into_linear) and the epilogue formulas of
    include/setok_b200.h, evaluated in torch on the CPU: with ANY lagged statistics (c, r) the consuming-side formula reproduces
    LayerNorm(x) W^T + b up to the bf16 rounding of xhat and W' (no GPU involved: this pins the algebra, the GPU tests pin the kernels)."""
    import torch
    from setok_b200._pack import fold_layernorm_into_linear
    g = torch.Generator().manual_seed(5)
    M, C, N, eps = 64, 256, 96, 1e-5
    x = torch.randn(M, C, generator=g) * (0.5 + 3 * torch.rand(M, 1, generator=g)) + 2 * torch.randn(M, 1, generator=g)
    gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.3 * torch.randn(C, generator=g)
    w, b = torch.randn(N, C, generator=g) * C ** -0.5, 0.1 * torch.randn(N, generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (C,), gamma.double(), beta.double(), eps) @ w.double().t() + b.double()
    wg, s, t = fold_layernorm_into_linear(w, b, gamma, beta)
    assert wg.dtype == torch.bfloat16 and torch.equal(s, wg.float().sum(1))
    for lag in (0.0, 0.3, -1.0):                    # how far the lagged mean / scale are from the row's true ones
        c = x.mean(1) + lag * x.std(1)
        r = (x.var(1, unbiased=False) * (1 + lag) ** 2 + eps).rsqrt()
        d = x - c[:, None]
        xhat = (d * r[:, None]).to(torch.bfloat16)                       # producing side
        parts = torch.stack([d[:, i:i + 128].sum(1) for i in range(0, C, 128)], 1), torch.stack([(d[:, i:i + 128] ** 2).sum(1) for i in range(0, C, 128)], 1)
        m = parts[0].sum(1) / C                                          # consuming side
        var = parts[1].sum(1) / C - m * m
        rho = (var + eps).rsqrt()
        acc = xhat.float() @ wg.float().t()
        y = (rho / r)[:, None] * acc + (-rho * m)[:, None] * s[None, :] + t[None, :]
        err = float((y.double() - ref).abs().max() / ref.abs().max())
        assert err < 8e-3, (lag, err)                                    # bf16 operands: 2^-9 relative per element
        exact = (rho[:, None] * (d - m[:, None])) * gamma + beta        # the same identity without any rounding
        assert torch.allclose(exact.double() @ w.double().t() + b.double(), ref, rtol=1e-4, atol=1e-4)

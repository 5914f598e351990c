"""CPU, build-container only: the oracle restatement against the reference's own source executed
live from /root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import math

import pytest
import torch

from oracle import ref_loader
from oracle import setok_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ns():
    return ref_loader.load_reference()


@pytest.mark.parametrize("N,C,G,k,thr,mcn", [(64, 32, 5, 8, 0.5, 16), (144, 80, 12, 16, 0.5, 32), (256, 1024, 32, 16, 0.5, 64),
                                              (100, 64, 0, 16, 0.5, 16), (576, 128, 64, 16, 0.5, 64)])
def test_dpc_live(ns, N, C, G, k, thr, mcn):
    tok = ref_loader.build_reference_tokenizer(ns, None, hidden_dim=64, token_feat_dim=64, min_cluster_num=mcn, threshold=thr)
    x = O.mog_features(N, C, G, 0.05, N + C) if G else torch.randn(N, C)
    torch.manual_seed(77)
    r_down, r_idx, r_score = tok.cluster_dpc_knn(x, k, None, thr)
    torch.manual_seed(77)
    noise = torch.rand(N)
    o_down, o_idx, o_score = O.dpc_knn(x, k, noise, thr, mcn)
    assert torch.equal(r_down, o_down) and torch.equal(r_idx, o_idx) and torch.equal(r_score, o_score)


def test_posenc_live_full_size(ns):
    for (h, C) in [(16, 1024), (24, 1024), (14, 768), (32, 1024)]:
        pe = ns.module.PositionalEncoding2D(C)
        assert torch.equal(pe(torch.zeros(1, h, h, C))[0], O.pos_encoding_2d(h, h, C))


def test_head_live(ns):
    C, Ctok, N = 128, 96, 100
    torch.manual_seed(1)
    tok = ref_loader.build_reference_tokenizer(ns, None, hidden_dim=C, token_feat_dim=Ctok, min_cluster_num=16,
                                               threshold=0.5, dim_feedforward=256)
    p = {k: v for k, v in tok.state_dict().items()}
    feat = O.mog_features(N, C, 9, 0.05, 4)
    with torch.no_grad():
        x = feat[None]
        pos = tok.position_embedding(x.reshape(1, 10, 10, C)).reshape(1, N, C)
        x = (x + pos)[0]
        torch.manual_seed(5)
        down, idx, score = tok.cluster_dpc_knn(x, 8, None, 0.5)
        gf = tok.group_encoding(x, x[down], idx)
        ref = tok.out(tok.inter_encoder(gf[None])[0])
    torch.manual_seed(5)
    noise = torch.rand(N)
    toks, oidx, oscore = O.tokenizer_head(feat, noise, p, min_cluster_num=16, threshold=0.5, k=8)
    assert torch.equal(idx, oidx)
    torch.testing.assert_close(toks, ref, rtol=1e-4, atol=1e-5)


def test_tower_live_vit_b16_shape(ns):
    """HF CLIPVisionModel (seeded) vs the functional restatement at a ViT-B/16-like width."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=192, intermediate_size=768, num_hidden_layers=3, num_attention_heads=3,
                           image_size=64, patch_size=16)
    cfg._attn_implementation = "eager"
    hf = CLIPVisionModel(cfg).eval()
    tok = ref_loader.build_reference_tokenizer(ns, hf, hidden_dim=192, token_feat_dim=192, min_cluster_num=4, threshold=0.5)
    img = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        ref = tok.image_feature_encoder(img)
    f = O.tower_features(img, dict(hf.state_dict()), patch=16, heads=3, layers=3, select_layer=-2)
    torch.testing.assert_close(f, ref, rtol=1e-4, atol=1e-5)


def test_interpolated_pos_encoding_live(ns):
    """Oracle's bicubic position-table resize vs HF's `interpolate_pos_encoding=True` (BASELINE config 5)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=32, patch_size=8)
    cfg._attn_implementation = "eager"
    hf = CLIPVisionModel(cfg).eval()
    img = torch.randn(2, 3, 48, 48)
    with torch.no_grad():
        ref = hf(img, output_hidden_states=True, interpolate_pos_encoding=True).hidden_states[-1][:, 1:]
    f = O.tower_features(img, dict(hf.state_dict()), patch=8, heads=2, layers=2, select_layer=-1, interpolate_pos_encoding=True)
    torch.testing.assert_close(f, ref, rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        O.tower_features(img, dict(hf.state_dict()), patch=8, heads=2, layers=2, select_layer=-1)

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


def test_detok_qformer_live(ns):
    """The detokenizer oracle's Q-Former against the reference's own BertEmbeddings / BertEncoder executed live, at BERT-base
    head geometry (heads of 64), three layers with cross-attention in layers 0 and 2, ragged K_b via the additive mask."""
    import torch.nn.functional as F
    from transformers.models.bert import BertConfig
    from oracle import detok_oracle as D
    d = dict(token_dim=32, hidden=128, q_heads=2, q_inter=256, q_layers=3, cross_freq=2, grid=3, dec_dim=64, dec_depth=0, dec_mlp=128)
    p = D.make_detok_params(**d, seed=5)
    cfg = BertConfig()
    cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size = d["hidden"], d["q_heads"], d["q_inter"]
    cfg.encoder_width, cfg.add_cross_attention, cfg.cross_attention_freq = d["hidden"], True, d["cross_freq"]
    cfg.query_length, cfg.num_hidden_layers = d["grid"] ** 2, d["q_layers"]
    M = ns.module
    emb, enc = M.BertEmbeddings(cfg).eval(), M.BertEncoder(cfg).eval()
    for layer in enc.layer:                       # detokenizer.py:94-96
        layer.output = None
        layer.intermediate = None
    emb.load_state_dict({k[len("mapper.embeddings."):]: v for k, v in p.items() if k.startswith("mapper.embeddings.")}, strict=False)
    missing, unexpected = enc.load_state_dict({k[len("mapper.encoder."):]: v for k, v in p.items() if k.startswith("mapper.encoder.")}, strict=False)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(6)
    x = torch.randn(4, 7, d["token_dim"], generator=g)
    mask = torch.tensor([[1, 1, 1, 1, 1, 1, 1], [1, 1, 0, 0, 0, 0, 0], [1, 1, 1, 1, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0]], dtype=torch.float32)
    with torch.no_grad():
        enc_in = F.linear(x, p["mapper_fc_in.weight"], p["mapper_fc_in.bias"])
        q = p["mask_tokens"].expand(4, -1, -1)
        h0 = emb(query_embeds=q)
        inv = (1.0 - mask[:, None, None, :]) * torch.finfo(torch.float32).min
        ref = enc(h0, attention_mask=torch.zeros(4, 1, 1, q.shape[1]), head_mask=[None] * d["q_layers"], encoder_hidden_states=enc_in,
                  encoder_attention_mask=inv, query_length=q.shape[1], return_dict=True).last_hidden_state
        got = D.qformer(p, q, enc_in, mask, heads=d["q_heads"], layers=d["q_layers"], cross_freq=d["cross_freq"])
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)


def test_splice_live():
    """The splice oracle against the reference's own prepare_inputs_labels_for_multimodal executed live on a stub self, on a
    random ragged batch (0..3 placeholders per sample, holes in the attention mask, left and right padding, truncation)."""
    import types
    from oracle import splice_oracle as S
    fn = ref_loader.load_reference_splice()
    g = torch.Generator().manual_seed(71)
    V, H, B, L = 80, 8, 9, 20
    emb = torch.nn.Embedding(V, H)
    emb.weight.data = torch.randn(V, H, generator=g)
    ids = torch.randint(0, V, (B, L), generator=g)
    am = (torch.rand(B, L, generator=g) > 0.15).long()
    n_img = 0
    for b in range(B):
        k = int(torch.randint(0, 4, (1,), generator=g))
        pos = torch.randperm(L, generator=g)[:k]
        ids[b, pos] = -200
        n_img += max(int(((ids[b] == -200) & am[b].bool()).sum()), 1)
    feats = [torch.randn(int(torch.randint(1, 6, (1,), generator=g)), H, generator=g) for _ in range(n_img)]
    labels = ids.clone()
    labels[labels == -200] = -100
    labels[:, 1] = -300
    for side, maxlen in (("right", None), ("left", None), ("right", 17), ("left", 11)):
        cfg = types.SimpleNamespace(tokenizer_padding_side=side)
        if maxlen is not None:
            cfg.tokenizer_model_max_length = maxlen
        me = types.SimpleNamespace(get_vision_tower=lambda: object(), encode_images=lambda images: feats,
                                   get_model=lambda: types.SimpleNamespace(embed_tokens=emb), config=cfg, device=torch.device("cpu"))
        with torch.no_grad():
            _, pos_r, mask_r, _, emb_r, lab_r = fn(me, ids.clone(), torch.zeros(B, L, dtype=torch.long), am.clone(), None, labels.clone(), torch.zeros(1, 3, 2, 2))
            e, l, m, p = S.splice(ids, am, labels, emb.weight.data, feats, maxlen, side)
        assert torch.equal(e, emb_r) and torch.equal(l, lab_r) and torch.equal(m, mask_r.bool()) and torch.equal(p, pos_r), (side, maxlen)

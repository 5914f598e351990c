"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

What runs: the reference's `SetokTokenizer.cluster_dpc_knn`, `group_encoding`, `Block`,
`PositionalEncoding2D`, `build_vision_projector` loaded unmodified through `oracle/ref_loader.py`,
HF `CLIPVisionModel` (seeded, from config, eager attention) behind the reference's
`CLIPVisionTower.forward`, and the repaired per-image forward R1-R4 composed
from those reference sub-modules.  The tie-break noise (R3) is reproduced by seeding the global
RNG right before each `cluster_dpc_knn` call: its only RNG draw is `torch.rand(N)`
(tokenizer.py:91), so `torch.manual_seed(s); torch.rand(N)` is the identical tensor.

Everything is fp32 on CPU with a fixed thread count so the files are reproducible.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import setok_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def _save(name, **d):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **_np(d))
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def ref_noise(seed: int, N: int) -> torch.Tensor:
    torch.manual_seed(seed)
    return torch.rand(N)


def golden_posenc(ns):
    out = {}
    for (h, w, C) in [(16, 16, 64), (7, 7, 50), (14, 14, 96), (4, 6, 30)]:
        pe = ns.module.PositionalEncoding2D(C)
        out[f"pe_{h}_{w}_{C}"] = pe(torch.zeros(1, h, w, C))[0]
    _save("posenc", **out)


def _dpc_case(tok, x, k, thr, seed, token_mask=None):
    torch.manual_seed(seed)
    index_down, idx_cluster, score = tok.cluster_dpc_knn(x, k, token_mask, thr)
    return index_down, idx_cluster, score


def golden_dpc(ns):
    """cluster_dpc_knn known answers: thresholded, fallback, masked, K<min_cluster_num, C=1024."""
    cases = []
    tok = ref_loader.build_reference_tokenizer(ns, None, hidden_dim=64, token_feat_dim=64, min_cluster_num=16, threshold=0.5)
    # (name, N, C, G(0=iid), sigma, k, thr, min_cluster_num, mask?)
    spec = [
        ("mog_n64_c64", 64, 64, 6, 0.05, 8, 0.5, 16, False),
        ("mog_n100_c48", 100, 48, 9, 0.05, 8, 0.5, 16, False),      # N not a multiple of 32, C not of 32
        ("iid_n64_c64_fallback", 64, 64, 0, 0.0, 16, 0.5, 16, False),
        ("iid_n196_c96_fallback32", 196, 96, 0, 0.0, 32, 1e9, 32, False),   # BASELINE config 1 head shape (K=32)
        ("mog_n256_c64", 256, 64, 32, 0.05, 16, 0.5, 64, False),
        ("mog_n64_c64_masked", 64, 64, 6, 0.05, 8, 0.5, 16, True),
        ("mog_n256_c64_small_clusters", 256, 64, 100, 0.05, 16, 0.5, 64, False),   # K < min_cluster_num
        ("mog_n576_c32", 576, 32, 40, 0.05, 16, 0.5, 64, False),
        ("mog_n256_c1024", 256, 1024, 32, 0.05, 16, 0.5, 64, False),
    ]
    out = {}
    names = []
    for i, (name, N, C, G, sigma, k, thr, mcn, masked) in enumerate(spec):
        seed = 100 + i
        x = O.mog_features(N, C, G, sigma, seed) if G else torch.randn(N, C, generator=torch.Generator().manual_seed(seed))
        tok.min_cluster_num = mcn
        tm = None
        if masked:
            tm = (torch.rand(N, generator=torch.Generator().manual_seed(seed + 1)) > 0.25).float()
        index_down, idx_cluster, score = _dpc_case(tok, x, k, thr, seed + 7, tm)
        noise = ref_noise(seed + 7, N)
        names.append(name)
        out[name + "/x"] = x
        out[name + "/noise"] = noise
        out[name + "/params"] = np.array([k, thr, mcn], dtype=np.float64)
        if tm is not None:
            out[name + "/token_mask"] = tm
        out[name + "/index_down"] = index_down
        out[name + "/idx_cluster"] = idx_cluster
        out[name + "/score"] = score
        print(f"  {name}: K={index_down.numel()}  score[{float(score.min()):.4f},{float(score.max()):.4f}]")
    out["names"] = np.array(names)
    _save("dpc_knn", **out)


def golden_block_and_head(ns):
    """Block, group_encoding and the repaired per-image head forward at C=64."""
    C, Ctok, Fd, N = 64, 48, 128, 64
    torch.manual_seed(3)
    tok = ref_loader.build_reference_tokenizer(ns, None, hidden_dim=C, token_feat_dim=Ctok, min_cluster_num=16,
                                               threshold=0.5, dim_feedforward=Fd)
    # perturb biases/LN so every term is exercised (the reference init is zero bias / unit LN)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in tok.named_parameters():
            if n.endswith("bias") or "norm" in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    sd = {k: v.clone() for k, v in tok.state_dict().items() if not k.startswith("image_feature_encoder")}
    out = {"sd/" + k: v for k, v in sd.items()}
    out["sd_keys"] = np.array(list(sd.keys()))
    # Block
    xb = torch.randn(3, 10, C, generator=g)
    with torch.no_grad():
        out["block_in"] = xb
        out["block_out"] = tok.inner_encoder(xb)
    # group_encoding + head for two images (one thresholded, one fallback)
    feats = torch.stack([O.mog_features(N, C, 6, 0.05, 21), torch.randn(N, C, generator=g)])
    out["feats"] = feats
    for b in range(2):
        with torch.no_grad():
            x = feats[b].unsqueeze(0)
            h = w = int(math.sqrt(N))
            pos = tok.position_embedding(x.reshape(1, h, w, C)).reshape(1, h * w, C)
            x = (x + pos).squeeze(0)
            index_down, idx_cluster, score = _dpc_case(tok, x, 8, 0.5, 500 + b)
            gf = tok.group_encoding(x, x[index_down, :], idx_cluster)
            gi = tok.inter_encoder(gf[None])[0]                      # R2
            tokens = tok.out(gi)
        out[f"img{b}/noise"] = ref_noise(500 + b, N)
        out[f"img{b}/index_down"] = index_down
        out[f"img{b}/idx_cluster"] = idx_cluster
        out[f"img{b}/score"] = score
        out[f"img{b}/group_features"] = gf
        out[f"img{b}/tokens"] = tokens
        print(f"  head img{b}: K={index_down.numel()}")
    out["cfg"] = np.array([C, Ctok, Fd, N, 8, 16], dtype=np.int64)   # C, C_tok, F, N, k, min_cluster_num
    _save("head", **out)


def golden_tower_and_e2e(ns):
    """Tiny seeded HF CLIPVisionModel behind the reference CLIPVisionTower, then the whole repaired
    forward + mlp2x_gelu projector (encode_images, setokim_arch.py:206-211)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    C, L, H, P, IMG = 64, 3, 4, 4, 32            # N = 64 patches
    torch.manual_seed(5)
    cfg = CLIPVisionConfig(hidden_size=C, intermediate_size=4 * C, num_hidden_layers=L, num_attention_heads=H,
                           image_size=IMG, patch_size=P, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    cfg._attn_implementation = "eager"
    hf = CLIPVisionModel(cfg).eval()
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():    # HF init leaves biases at zero: perturb so they are exercised
        for n, p in hf.named_parameters():
            if n.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    Ctok, Hllm = 48, 80
    out = {}
    images = torch.randn(2, 3, IMG, IMG, generator=g)
    out["images"] = images
    for sl in (-2, -1):
        tok = ref_loader.build_reference_tokenizer(ns, hf, hidden_dim=C, token_feat_dim=Ctok, min_cluster_num=8,
                                                   threshold=0.5, dim_feedforward=2 * C, select_layer=sl)
        with torch.no_grad():
            out[f"feats_sl{sl}"] = tok.image_feature_encoder(images)
    # list input path of CLIPVisionTower.forward (clip_encoder.py:52-57)
    with torch.no_grad():
        lst = tok.image_feature_encoder([images[0], images[1]])
    out["feats_list0"] = lst[0]
    sd = {k: v.clone() for k, v in hf.state_dict().items()}
    for k_, v in sd.items():
        out["tower/" + k_] = v
    out["tower_keys"] = np.array(list(sd.keys()))
    out["tower_cfg"] = np.array([C, L, H, P, IMG], dtype=np.int64)

    # end-to-end with select_layer=-2
    torch.manual_seed(9)
    tok = ref_loader.build_reference_tokenizer(ns, hf, hidden_dim=C, token_feat_dim=Ctok, min_cluster_num=8,
                                               threshold=0.5, dim_feedforward=2 * C, select_layer=-2)
    with torch.no_grad():
        for n, p in tok.named_parameters():
            if not n.startswith("image_feature_encoder") and (n.endswith("bias") or "norm" in n):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    hsd = {k: v.clone() for k, v in tok.state_dict().items() if not k.startswith("image_feature_encoder")}
    for k_, v in hsd.items():
        out["head/" + k_] = v
    out["head_keys"] = np.array(list(hsd.keys()))
    torch.manual_seed(23)
    proj = ns.projector_builder.build_vision_projector("mlp2x_gelu", mm_hidden_size=Ctok, hidden_size=Hllm).eval()
    for k_, v in proj.state_dict().items():
        out["proj/" + k_] = v.clone()
    out["proj_keys"] = np.array(list(proj.state_dict().keys()))
    N = (IMG // P) ** 2
    with torch.no_grad():
        feats = tok.image_feature_encoder(images)                    # R1: batched tower
        for b in range(images.shape[0]):
            x = feats[b].unsqueeze(0)
            h = w = int(math.sqrt(N))
            pos = tok.position_embedding(x.reshape(1, h, w, C)).reshape(1, h * w, C)
            x = (x + pos).squeeze(0)
            index_down, idx_cluster, score = _dpc_case(tok, x, 4, 0.4, 700 + b)
            gf = tok.group_encoding(x, x[index_down, :], idx_cluster)
            tokens = tok.out(tok.inter_encoder(gf[None])[0])
            out[f"e2e{b}/noise"] = ref_noise(700 + b, N)
            out[f"e2e{b}/idx_cluster"] = idx_cluster
            out[f"e2e{b}/score"] = score
            out[f"e2e{b}/tokens"] = tokens
            out[f"e2e{b}/projected"] = proj(tokens)
            print(f"  e2e img{b}: K={index_down.numel()}")
    out["e2e_cfg"] = np.array([Ctok, Hllm, 2 * C, 4, 8], dtype=np.int64)   # C_tok, H, F, k, min_cluster_num
    out["e2e_thr"] = np.array([0.4])
    _save("tower_e2e", **out)


def golden_projectors(ns):
    out = {}
    g = torch.Generator().manual_seed(31)
    x = torch.randn(7, 24, generator=g)
    out["x"] = x
    for t in ("linear", "mlp2x_gelu", "mlp3x_gelu", "mlp2x_gelu_Norm", "identity"):
        torch.manual_seed(41)
        m = ns.projector_builder.build_vision_projector(t, mm_hidden_size=24, hidden_size=40).eval()
        sd = m.state_dict() if hasattr(m, "state_dict") else {}
        for k_, v in sd.items():
            out[f"{t}/{k_}"] = v.clone()
        out[f"{t}/keys"] = np.array(list(sd.keys()))
        with torch.no_grad():
            out[f"{t}/y"] = m(x)
    _save("projectors", **out)


DETOK_DIMS = dict(token_dim=48, hidden=64, q_heads=4, q_inter=128, q_layers=3, cross_freq=2, grid=4, dec_dim=64, dec_depth=2, dec_mlp=256)
DETOK_DEC_HEADS = 4
DETOK_K = [5, 2, 7]


def golden_detok(ns):
    """SetokDeTokenizer.forward (detokenizer.py:101-120) composed from the reference's own BertEmbeddings / BertEncoder /
    PositionalEncoding2D classes (executed unmodified) and HF ViTLayer standing in for timm's Block (same pre-LN block;
    timm is not installed).  `SetokDeTokenizer.__init__` itself cannot run here: it needs timm, diffusers and a local
    bert-base-uncased config; BertModel.__init__ does not construct under transformers 5.x (see oracle/detok_oracle.py)."""
    import torch.nn.functional as F
    from transformers.models.bert import BertConfig
    from transformers.models.vit.configuration_vit import ViTConfig
    from transformers.models.vit.modeling_vit import ViTLayer
    from oracle import detok_oracle as D
    M = ns.module
    d = DETOK_DIMS
    p = D.make_detok_params(**d, seed=11)
    Q = d["grid"] ** 2
    cfg = BertConfig()                               # == bert-base-uncased's config.json (detokenizer.py:80)
    cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size = d["hidden"], d["q_heads"], d["q_inter"]
    cfg.encoder_width, cfg.add_cross_attention, cfg.cross_attention_freq = d["hidden"], True, d["cross_freq"]   # :82-86
    cfg.query_length, cfg.num_hidden_layers = Q, d["q_layers"]                                                  # :87-88
    emb, enc = M.BertEmbeddings(cfg).eval(), M.BertEncoder(cfg).eval()
    for layer in enc.layer:                                                                                     # :94-96
        layer.output = None
        layer.intermediate = None
    emb.load_state_dict({k[len("mapper.embeddings."):]: v for k, v in p.items() if k.startswith("mapper.embeddings.")}, strict=False)
    missing, unexpected = enc.load_state_dict({k[len("mapper.encoder."):]: v for k, v in p.items() if k.startswith("mapper.encoder.")}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    pe = M.PositionalEncoding2D(d["hidden"])

    g = torch.Generator().manual_seed(12)
    tokens = torch.randn(sum(DETOK_K), d["token_dim"], generator=g)
    offsets = [0]
    for k_ in DETOK_K:
        offsets.append(offsets[-1] + k_)
    x, mask = D.pad_ragged(tokens, offsets)
    B = x.shape[0]
    with torch.no_grad():
        enc_in = F.linear(x, p["mapper_fc_in.weight"], p["mapper_fc_in.bias"])                                  # :104
        h0 = emb(query_embeds=p["mask_tokens"].expand(B, -1, -1))
        ext = torch.zeros(B, 1, 1, Q)                                               # get_extended_attention_mask(ones)
        inv = (1.0 - mask[:, None, None, :]) * torch.finfo(torch.float32).min       # invert_attention_mask
        h = enc(h0, attention_mask=ext, head_mask=[None] * d["q_layers"], encoder_hidden_states=enc_in, encoder_attention_mask=inv,
                query_length=Q, return_dict=True).last_hidden_state                                             # :105-109
        y = F.linear(h, p["decoder_fc_in.weight"], p["decoder_fc_in.bias"])                                     # :111
        pos = pe(y.reshape(B, d["grid"], d["grid"], -1)).reshape(B, Q, -1)                                      # :112-114
        y = y + pos
        vcfg = ViTConfig(hidden_size=d["dec_dim"], num_attention_heads=DETOK_DEC_HEADS, intermediate_size=d["dec_mlp"], hidden_act="gelu",
                         layer_norm_eps=1e-5, qkv_bias=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        vcfg._attn_implementation = "eager"
        C = d["dec_dim"]
        for i in range(d["dec_depth"]):
            blk = ViTLayer(vcfg).eval()
            b = f"pixel_decoder.{i}."
            sd = {"layernorm_before.weight": p[b + "norm1.weight"], "layernorm_before.bias": p[b + "norm1.bias"],
                  "layernorm_after.weight": p[b + "norm2.weight"], "layernorm_after.bias": p[b + "norm2.bias"],
                  "attention.output.dense.weight": p[b + "attn.proj.weight"], "attention.output.dense.bias": p[b + "attn.proj.bias"],
                  "intermediate.dense.weight": p[b + "mlp.fc1.weight"], "intermediate.dense.bias": p[b + "mlp.fc1.bias"],
                  "output.dense.weight": p[b + "mlp.fc2.weight"], "output.dense.bias": p[b + "mlp.fc2.bias"]}
            for j, n in enumerate(("query", "key", "value")):
                sd[f"attention.attention.{n}.weight"] = p[b + "attn.qkv.weight"][j * C:(j + 1) * C]
                sd[f"attention.attention.{n}.bias"] = p[b + "attn.qkv.bias"][j * C:(j + 1) * C]
            missing, unexpected = blk.load_state_dict(sd, strict=False)
            assert not unexpected and not missing, (missing, unexpected)
            o = blk(y)
            y = o[0] if isinstance(o, tuple) else o
        out = F.layer_norm(y, (C,), p["decoder_norm.weight"], p["decoder_norm.bias"], 1e-5)                     # :120
    keys = sorted(p.keys())
    save = {f"param/{k}": p[k] for k in keys}
    save.update(dict(keys=np.array(keys), tokens=tokens, offsets=np.array(offsets, dtype=np.int32), x=x, mask=mask, qformer=h, pos=pos[0], out=out,
                     dims=np.array([d[k] for k in ("token_dim", "hidden", "q_heads", "q_inter", "q_layers", "cross_freq", "grid", "dec_dim", "dec_depth", "dec_mlp")] + [DETOK_DEC_HEADS])))
    _save("detok", **save)


def golden_splice():
    """prepare_inputs_labels_for_multimodal (setokim_arch.py:213-354), the reference's own function object, on a stub self."""
    import types
    fn = ref_loader.load_reference_splice()
    g = torch.Generator().manual_seed(51)
    V, H, B, L = 50, 16, 5, 12
    emb = torch.nn.Embedding(V, H)
    emb.weight.data = torch.randn(V, H, generator=g)
    K = [3, 1, 4, 2, 5, 2]
    feats = [torch.randn(k, H, generator=g) for k in K]
    ids = torch.randint(0, V, (B, L), generator=g)
    am = torch.ones(B, L, dtype=torch.long)
    ids[0, 2] = -200                                    # one image
    ids[1, 0] = -200; ids[1, 7] = -200                  # two images, one at the very start
    am[1, 9:] = 0                                       # right-padded text
    # sample 2: no placeholder -> still consumes one image index (:262-269)
    ids[3, 11] = -200; am[3, :3] = 0                    # left-padded text, image at the very end
    ids[4, 5] = -200; am[4, 6] = 0                      # a masked hole in the middle
    labels = ids.clone()
    labels[labels == -200] = -100
    labels[0, 5] = -300                                 # TARGET_TOKEN_INDEX -> IGNORE (:345)
    out = dict(V=np.array(V), H=np.array(H), embed=emb.weight.data.clone(), input_ids=ids, attention_mask=am, labels=labels,
               feats=torch.cat(feats, 0), offsets=np.cumsum([0] + K).astype(np.int32))
    cases = {"right": dict(side="right", maxlen=None, with_labels=True, with_mask=True),
             "left": dict(side="left", maxlen=None, with_labels=True, with_mask=True),
             "trunc": dict(side="right", maxlen=13, with_labels=True, with_mask=True),
             "nolabels_nomask": dict(side="right", maxlen=None, with_labels=False, with_mask=False)}
    for name, c in cases.items():
        cfg = types.SimpleNamespace(tokenizer_padding_side=c["side"])
        if c["maxlen"] is not None:
            cfg.tokenizer_model_max_length = c["maxlen"]
        model = types.SimpleNamespace(embed_tokens=emb)
        me = types.SimpleNamespace(get_vision_tower=lambda: object(), encode_images=lambda images: feats, get_model=lambda: model, config=cfg,
                                   device=torch.device("cpu"))
        pos_in = torch.arange(L)[None].expand(B, L)
        with torch.no_grad():
            r = fn(me, ids.clone(), pos_in, am.clone() if c["with_mask"] else None, None, labels.clone() if c["with_labels"] else None,
                   torch.zeros(len(K), 3, 4, 4))
        _, pos, mask, _, embeds, lab = r
        out[name + "/embeds"] = embeds
        out[name + "/pos"] = pos
        out[name + "/mask"] = mask if mask is not None else np.zeros(0)
        out[name + "/labels"] = lab if lab is not None else np.zeros(0)
        out[name + "/cfg"] = np.array([1 if c["side"] == "left" else 0, c["maxlen"] or 0, int(c["with_labels"]), int(c["with_mask"])])
    out["names"] = np.array(list(cases.keys()))
    _save("splice", **out)


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not present: goldens can only be generated in the build container")
    torch.set_num_threads(4)
    ns = ref_loader.load_reference()
    golden_posenc(ns)
    golden_dpc(ns)
    golden_block_and_head(ns)
    golden_tower_and_e2e(ns)
    golden_projectors(ns)
    golden_detok(ns)
    golden_splice()


if __name__ == "__main__":
    main()

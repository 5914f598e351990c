"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (torch fp32) restatement of the SeTok tokenizer hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
this module, and only as the *checker* (or as the timed CPU baseline).  Nothing under
`setok_b200/` imports it; the product path fails loudly when the CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors for this path, so
this restatement is pinned against *outputs of the reference itself run in the build
container*: `oracle/make_golden.py` executes the reference's own
`src/model/setok/{utils,module,tokenizer}.py` (through `oracle/ref_loader.py`) and HF
`CLIPVisionModel`, and stores inputs + outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function below against them (bit-exact for the
integer outputs).  `tests/test_oracle_vs_reference.py` repeats that live when /root/reference
is present.

Each function cites the reference file:line it restates (paths relative to the reference
root).  The closed list of repairs applied to make the committed reference executable
(SURVEY.md §8c): R1 per-image loop over a batched tower output, R2 `inter_encoder(gf[None])[0]`,
R3 explicit tie-break noise tensor instead of the global-RNG `torch.rand`, R4 CLIP-style tower
(CLS token) built from config with seeded weights.

Parameters are plain ``dict[str, Tensor]`` keyed with the reference ``state_dict`` names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------
# a3: PositionalEncoding2D  (src/model/setok/module.py:105-146, src/model/setok/utils.py:5-10)
# ----------------------------------------------------------------------------------------
def pos_encoding_2d(h: int, w: int, C: int, dtype=torch.float32) -> torch.Tensor:
    """(h, w, C) table.  module.py:112 ch = ceil(C/4)*2; :114 inv_freq; :133-136 outer products,
    sin/cos interleaved (utils.py:9-10); :142-143 first ch channels <- row index, next ch <-
    column index; :145 sliced to C."""
    ch = int(math.ceil(C / 4) * 2)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
    pos_x = torch.arange(h, dtype=inv_freq.dtype)
    pos_y = torch.arange(w, dtype=inv_freq.dtype)
    sx = pos_x[:, None] * inv_freq[None, :]
    sy = pos_y[:, None] * inv_freq[None, :]
    ex = torch.stack((sx.sin(), sx.cos()), dim=-1).flatten(-2, -1)  # (h, ch)
    ey = torch.stack((sy.sin(), sy.cos()), dim=-1).flatten(-2, -1)  # (w, ch)
    emb = torch.zeros(h, w, 2 * ch, dtype=dtype)
    emb[:, :, :ch] = ex[:, None, :].to(dtype)
    emb[:, :, ch:2 * ch] = ey[None, :, :].to(dtype)
    return emb[:, :, :C].contiguous()


# ----------------------------------------------------------------------------------------
# a4: cluster_dpc_knn  (src/model/setok/tokenizer.py:78-121)
# ----------------------------------------------------------------------------------------
def dpc_knn(x: torch.Tensor, k: int, noise: torch.Tensor, threshold: float, min_cluster_num: int,
            token_mask: Optional[torch.Tensor] = None, return_intermediates: bool = False):
    """x (N, C) fp32, noise (N,) = the `torch.rand(N)` draw of tokenizer.py:91 *before* the 1e-6
    scale (repair R3).  Returns (index_down int64 (K,), idx_cluster int64 (N,), score (1, N))."""
    N, C = x.shape
    dist_matrix = torch.cdist(x, x) / (C ** 0.5)                                   # :82
    if token_mask is not None:                                                      # :84-86
        token_mask = token_mask > 0
        dist_matrix = dist_matrix * token_mask[None, :] + (dist_matrix.max() + 1) * (~token_mask[None, :])
    dist_nearest, _ = torch.topk(dist_matrix, k=k, dim=-1, largest=False)          # :88
    density = (-(dist_nearest ** 2).mean(dim=-1)).exp()                            # :90
    density = density + noise.to(density.dtype) * 1e-6                             # :91 (R3)
    if token_mask is not None:                                                      # :93-94
        density = density * token_mask
    mask = (density[None, :] > density[:, None]).type(x.dtype)                     # :96-97
    # :98 -- dist_max has shape (1, 1, N): it broadcasts along the LAST axis, i.e. the fill value
    # for entry (i, j) is rowmax[j], and the min below returns shape (1, N).
    dist_max = dist_matrix.flatten(1).max(dim=-1)[0][None, None]
    dist, index_parent = (dist_matrix * mask + dist_max * (1 - mask)).min(dim=-1)  # :99
    score = dist * density                                                          # :101  (1, N)
    index_down = torch.nonzero(score.reshape(-1) > threshold).reshape(-1)          # :103
    if index_down.numel() == 0:                                                     # :104-107
        _, index_down = torch.topk(score, k=min_cluster_num, dim=-1)
        index_down = torch.sort(index_down).values.reshape(-1)
    sel = dist_matrix[index_down, :]                                                # :111
    idx_cluster = sel.argmin(dim=0)                                                 # :113
    idx_cluster[index_down] = torch.arange(index_down.size(0))                      # :117-119
    if return_intermediates:
        return index_down, idx_cluster, score, dict(dist_matrix=dist_matrix, density=density,
                                                    parent_dist=dist.reshape(-1))
    return index_down, idx_cluster, score


def dpc_margins(x: torch.Tensor, k: int, noise: torch.Tensor, threshold: float, min_cluster_num: int,
                token_mask: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """Decision margins of one clustering problem (test helper): the smallest gap by which any
    integer decision of `dpc_knn` was taken.  An implementation that reproduces the float
    intermediates to better than these margins must reproduce the integer outputs bit-exactly."""
    index_down, idx_cluster, score, im = dpc_knn(x, k, noise, threshold, min_cluster_num, token_mask, True)
    D = im["dist_matrix"]
    dens = torch.sort(im["density"]).values
    s = score.reshape(-1)
    out = {"density_gap": float((dens[1:] - dens[:-1]).min()) if dens.numel() > 1 else float("inf")}
    if bool((s > threshold).any()):
        out["threshold_margin"] = float((s - threshold).abs().min())
    else:
        ss = torch.sort(s, descending=True).values
        kk = min_cluster_num
        out["threshold_margin"] = float(ss[kk - 1] - ss[kk]) if kk < ss.numel() else float("inf")
    sel = D[index_down, :]
    if sel.shape[0] > 1:
        two = torch.topk(sel, 2, dim=0, largest=False).values
        gap = two[1] - two[0]
        gap[index_down] = float("inf")     # centres are overwritten (:117-119): no decision there
        out["argmin_margin"] = float(gap.min())
    else:
        gap = torch.full((x.shape[0],), float("inf"))
        out["argmin_margin"] = float("inf")
    out["token_gap"] = gap                 # per-token gap between the two nearest centres
    return out


# ----------------------------------------------------------------------------------------
# module.py:29-100  Mlp / Attention / Block
# ----------------------------------------------------------------------------------------
def attention(x: torch.Tensor, p: Params, prefix: str, nheads: int) -> torch.Tensor:
    """module.py:61-73.  x (B, N, C).  Fused qkv Linear, `nheads` heads of C/nheads, softmax of
    q k^T * head_dim^-0.5, proj.  Dropouts are identity (eval)."""
    B, N, C = x.shape
    hd = C // nheads
    qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])
    qkv = qkv.reshape(B, N, 3, nheads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(y, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def block(x: torch.Tensor, p: Params, prefix: str, depth: int, nheads: int) -> torch.Tensor:
    """module.py:95-100.  `depth` x [x += Attn_i(norm1(x))] with ONE shared norm1 (module.py:81,88),
    then x += Mlp(norm2(x)); GELU(erf) (module.py:30,35), LayerNorm eps 1e-5."""
    C = x.shape[-1]
    for i in range(depth):
        h = F.layer_norm(x, (C,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-5)
        x = x + attention(h, p, f"{prefix}layers.{i}.1.", nheads)
    h = F.layer_norm(x, (C,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-5)
    h = F.linear(h, p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])
    return x + h


# ----------------------------------------------------------------------------------------
# a5: group_encoding  (src/model/setok/tokenizer.py:123-155)
# ----------------------------------------------------------------------------------------
def group_encoding(x: torch.Tensor, labels: torch.Tensor, p: Params, depth: int, nheads: int) -> torch.Tensor:
    """x (N, C), labels (N,) -> (K, C).  For each label in sorted `labels.unique()` (:141; always
    0..K-1 because every centre owns its own label, :117-119): Block over the member tokens as a
    batch of one (:150), then mean over tokens (:151)."""
    outs = []
    for lab in labels.unique():
        m = labels == lab
        y = block(x[m].unsqueeze(0), p, "inner_encoder.", depth, nheads)
        outs.append(y.squeeze(0).mean(dim=0))
    return torch.stack(outs, dim=0)


# ----------------------------------------------------------------------------------------
# a6/a7: head of SetokTokenizer.forward for ONE image  (tokenizer.py:162-182, repairs R1-R3)
# ----------------------------------------------------------------------------------------
def tokenizer_head(feat: torch.Tensor, noise: torch.Tensor, p: Params, *, min_cluster_num: int,
                   threshold: float, k: Optional[int] = None, thr: Optional[float] = None,
                   token_mask: Optional[torch.Tensor] = None, nheads: int = 2,
                   inner_depth: int = 2, inter_depth: int = 2, return_intermediates: bool = False):
    """feat (N, C): one image's tower features (R1).  Returns (tokens (K, C_tok), idx_cluster (N,),
    score (1, N))."""
    N, C = feat.shape
    h = w = int(math.sqrt(N))                                                       # :164
    x = feat + pos_encoding_2d(h, w, C, feat.dtype).reshape(h * w, C)               # :165-169
    _thr = thr if thr else threshold                                                # :171 (0 is falsy)
    _k = k if k else min_cluster_num                                                # :172
    index_down, idx_cluster, score = dpc_knn(x, _k, noise, _thr, min_cluster_num, token_mask)   # :174
    gf = group_encoding(x, idx_cluster, p, inner_depth, nheads)                     # :178
    gi = block(gf[None], p, "inter_encoder.", inter_depth, nheads)[0]               # :179 (R2)
    tokens = F.linear(gi, p["out.weight"], p["out.bias"])                           # :180
    if return_intermediates:
        return tokens, idx_cluster, score, dict(x=x, index_down=index_down, group_features=gf, inter=gi)
    return tokens, idx_cluster, score


# ----------------------------------------------------------------------------------------
# a1/a2: CLIP ViT tower (third-party: transformers==4.46.3 pinned by the reference's
# pyproject.toml:18; transformers 5.5.0 is what the container has.  Published algorithm of
# transformers/models/clip/modeling_clip.py: CLIPVisionEmbeddings :138-220, CLIPAttention
# :261-336, CLIPMLP :339-351, CLIPEncoderLayer :354-386, CLIPVisionTransformer :647-690),
# reached from src/model/setok/clip_encoder.py:50-62 with feature_select :40-48.
# ----------------------------------------------------------------------------------------
def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


def interpolated_pos_embedding(pos: torch.Tensor, new_grid: int) -> torch.Tensor:
    """CLIPVisionEmbeddings.interpolate_pos_encoding (modeling_clip.py:160-196): class row kept, the patch grid resized
    with bicubic interpolation (align_corners=False).  pos (1+g*g, C) -> (1+new_grid^2, C)."""
    n = pos.shape[0] - 1
    g = int(n ** 0.5)
    if g == new_grid:
        return pos
    C = pos.shape[1]
    grid = pos[1:].reshape(1, g, g, C).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(new_grid, new_grid), mode="bicubic", align_corners=False)
    return torch.cat([pos[:1], grid.permute(0, 2, 3, 1).reshape(-1, C)], dim=0)


def clip_vit_hidden_states(images: torch.Tensor, p: Params, *, patch: int, heads: int, layers: int,
                           prefix: str = "vision_model.", n_layers_run: Optional[int] = None,
                           eps: float = 1e-5, interpolate_pos_encoding: bool = False) -> List[torch.Tensor]:
    """Returns HF's `hidden_states` tuple: [pre_layrnorm(embeddings), layer_1 out, ..., layer_L out].
    `n_layers_run` stops early (only hidden_states[:n+1] are produced)."""
    B = images.shape[0]
    Wp = p[prefix + "embeddings.patch_embedding.weight"]                # (C, 3, p, p), no bias
    C = Wp.shape[0]
    pe = F.conv2d(images.to(Wp.dtype), Wp, None, stride=patch).flatten(2).transpose(1, 2)   # (B, N, C)
    cls = p[prefix + "embeddings.class_embedding"].expand(B, 1, C)
    pos = p[prefix + "embeddings.position_embedding.weight"]
    if pos.shape[0] != pe.shape[1] + 1:
        if not interpolate_pos_encoding:
            raise ValueError(f"Input image size ({images.shape[2]}*{images.shape[3]}) doesn't match model.")
        pos = interpolated_pos_embedding(pos, images.shape[2] // patch)
    x = torch.cat([cls, pe], dim=1) + pos[None]
    x = F.layer_norm(x, (C,), p[prefix + "pre_layrnorm.weight"], p[prefix + "pre_layrnorm.bias"], eps)
    hs = [x]
    hd = C // heads
    L = layers if n_layers_run is None else n_layers_run
    for i in range(L):
        q = f"{prefix}encoder.layers.{i}."
        h = F.layer_norm(x, (C,), p[q + "layer_norm1.weight"], p[q + "layer_norm1.bias"], eps)
        T = h.shape[1]
        qh = F.linear(h, p[q + "self_attn.q_proj.weight"], p[q + "self_attn.q_proj.bias"]).view(B, T, heads, hd).transpose(1, 2)
        kh = F.linear(h, p[q + "self_attn.k_proj.weight"], p[q + "self_attn.k_proj.bias"]).view(B, T, heads, hd).transpose(1, 2)
        vh = F.linear(h, p[q + "self_attn.v_proj.weight"], p[q + "self_attn.v_proj.bias"]).view(B, T, heads, hd).transpose(1, 2)
        a = torch.matmul(qh, kh.transpose(-1, -2)) * (hd ** -0.5)
        a = F.softmax(a, dim=-1, dtype=torch.float32).to(qh.dtype)
        o = torch.matmul(a, vh).transpose(1, 2).reshape(B, T, C)
        x = x + F.linear(o, p[q + "self_attn.out_proj.weight"], p[q + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (C,), p[q + "layer_norm2.weight"], p[q + "layer_norm2.bias"], eps)
        h = quick_gelu(F.linear(h, p[q + "mlp.fc1.weight"], p[q + "mlp.fc1.bias"]))
        x = x + F.linear(h, p[q + "mlp.fc2.weight"], p[q + "mlp.fc2.bias"])
        hs.append(x)
    return hs


def layers_needed(select_layer: int, layers: int) -> int:
    """hidden_states has layers+1 entries; entry i needs i encoder layers."""
    idx = select_layer if select_layer >= 0 else layers + 1 + select_layer
    if not 0 <= idx <= layers:
        raise IndexError(f"select_layer {select_layer} out of range for {layers} layers")
    return idx


def tower_features(images: torch.Tensor, p: Params, *, patch: int, heads: int, layers: int,
                   select_layer: int = -2, select_feature: str = "patch",
                   prefix: str = "vision_model.", interpolate_pos_encoding: bool = False) -> torch.Tensor:
    """clip_encoder.py:50-62 + feature_select :40-48: hidden_states[select_layer], CLS dropped for
    'patch', kept for 'cls_patch', anything else raises ValueError (:47)."""
    if select_feature not in ("patch", "cls_patch"):
        raise ValueError(f"Unexpected select feature: {select_feature}")
    n = layers_needed(select_layer, layers)
    hs = clip_vit_hidden_states(images, p, patch=patch, heads=heads, layers=layers, prefix=prefix, n_layers_run=n,
                                interpolate_pos_encoding=interpolate_pos_encoding)
    f = hs[n]
    if select_feature == "patch":
        f = f[:, 1:]
    return f.to(images.dtype)


# ----------------------------------------------------------------------------------------
# a8: mm_in_projector  (src/model/multimodal_projector/builder.py:33-64) + encode_images
# (src/model/setokim_arch.py:206-211)
# ----------------------------------------------------------------------------------------
def projector(x: torch.Tensor, p: Params, projector_type: str = "mlp2x_gelu", prefix: str = "") -> torch.Tensor:
    """'linear' -> one Linear (key `weight`/`bias`); 'mlpNx_gelu' -> Sequential indices 0,2,4,..
    with GELU(erf) between; '_Norm' inserts LayerNorm at index 1 (then Linear layers sit at 0,3,5..);
    'identity' returns x."""
    import re
    if projector_type == "identity":
        return x
    if projector_type == "linear":
        return F.linear(x, p[prefix + "weight"], p[prefix + "bias"])
    use_norm = "_Norm" in projector_type
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type.replace("_Norm", ""))
    if not m:
        raise ValueError(f"Unknown projector type: {projector_type}")
    depth = int(m.group(1))
    idx = 0
    x = F.linear(x, p[f"{prefix}{idx}.weight"], p[f"{prefix}{idx}.bias"])
    idx += 1
    if use_norm:
        x = F.layer_norm(x, (x.shape[-1],), p[f"{prefix}{idx}.weight"], p[f"{prefix}{idx}.bias"], 1e-5)
        idx += 1
    for _ in range(1, depth):
        x = F.gelu(x)
        idx += 1
        x = F.linear(x, p[f"{prefix}{idx}.weight"], p[f"{prefix}{idx}.bias"])
        idx += 1
    return x


def setok_forward(images: torch.Tensor, noise: torch.Tensor, tower_p: Params, head_p: Params, *,
                  patch: int, heads: int, layers: int, select_layer: int, min_cluster_num: int,
                  threshold: float, k: Optional[int] = None, thr: Optional[float] = None,
                  nheads: int = 2, inner_depth: int = 2, inter_depth: int = 2,
                  feats: Optional[torch.Tensor] = None):
    """Whole tokenizer for a batch (R1: tower batched, head per image).  noise (B, N).
    Returns lists (tokens_b, idx_cluster_b, score_b)."""
    if feats is None:
        feats = tower_features(images, tower_p, patch=patch, heads=heads, layers=layers, select_layer=select_layer)
    toks, idxs, scores = [], [], []
    for b in range(feats.shape[0]):
        t, i, s = tokenizer_head(feats[b], noise[b], head_p, min_cluster_num=min_cluster_num, threshold=threshold,
                                 k=k, thr=thr, nheads=nheads, inner_depth=inner_depth, inter_depth=inter_depth)
        toks.append(t); idxs.append(i); scores.append(s)
    return toks, idxs, scores


# ----------------------------------------------------------------------------------------
# Seeded parameter factories (shared by tests / bench so that CPU and GPU see identical weights)
# ----------------------------------------------------------------------------------------
def make_head_params(C: int, C_tok: int, F_dim: int = 4096, inner_depth: int = 2, inter_depth: int = 2,
                     seed: int = 0, randomize_norm_bias: bool = True) -> Params:
    """Head parameters with the reference's init (tokenizer.py:59-72: xavier_uniform weights, zero
    bias, LN = (1, 0)).  `randomize_norm_bias` perturbs biases / LN affine slightly so parity tests
    exercise every term (zero biases would hide a missing bias add)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, out_f, in_f):
        bound = math.sqrt(6.0 / (in_f + out_f))
        p[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.randn(out_f, generator=g) * 0.02) if randomize_norm_bias else torch.zeros(out_f)

    def ln(name):
        p[name + ".weight"] = 1.0 + (torch.randn(C, generator=g) * 0.05 if randomize_norm_bias else torch.zeros(C))
        p[name + ".bias"] = torch.randn(C, generator=g) * 0.02 if randomize_norm_bias else torch.zeros(C)

    for enc, depth in (("inner_encoder", inner_depth), ("inter_encoder", inter_depth)):
        ln(f"{enc}.norm1"); ln(f"{enc}.norm2")
        for i in range(depth):
            lin(f"{enc}.layers.{i}.1.qkv", 3 * C, C)
            lin(f"{enc}.layers.{i}.1.proj", C, C)
        lin(f"{enc}.mlp.fc1", F_dim, C)
        lin(f"{enc}.mlp.fc2", C, F_dim)
    lin("out", C_tok, C)
    return p


def make_tower_params(C: int, layers: int, heads: int, patch: int, image: int, mlp: Optional[int] = None,
                      seed: int = 0, prefix: str = "vision_model.") -> Params:
    """Seeded CLIP-ViT parameters with HF key names (std 0.02-style init, enough for parity work;
    the goldens use HF's own initialisation instead)."""
    g = torch.Generator().manual_seed(seed)
    mlp = mlp or 4 * C
    n_pos = (image // patch) ** 2 + 1
    p: Params = {}
    r = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    p[prefix + "embeddings.class_embedding"] = r(C, std=C ** -0.5)
    p[prefix + "embeddings.patch_embedding.weight"] = r(C, 3, patch, patch, std=0.02)
    p[prefix + "embeddings.position_embedding.weight"] = r(n_pos, C, std=0.02)
    for n in ("pre_layrnorm", "post_layernorm"):
        p[f"{prefix}{n}.weight"] = 1.0 + r(C, std=0.05)
        p[f"{prefix}{n}.bias"] = r(C, std=0.02)
    for i in range(layers):
        q = f"{prefix}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            p[q + f"self_attn.{n}.weight"] = r(C, C, std=C ** -0.5)
            p[q + f"self_attn.{n}.bias"] = r(C, std=0.02)
        for n in ("layer_norm1", "layer_norm2"):
            p[q + n + ".weight"] = 1.0 + r(C, std=0.05)
            p[q + n + ".bias"] = r(C, std=0.02)
        p[q + "mlp.fc1.weight"] = r(mlp, C, std=C ** -0.5)
        p[q + "mlp.fc1.bias"] = r(mlp, std=0.02)
        p[q + "mlp.fc2.weight"] = r(C, mlp, std=mlp ** -0.5)
        p[q + "mlp.fc2.bias"] = r(C, std=0.02)
    return p


def make_projector_params(C_tok: int, H: int, projector_type: str = "mlp2x_gelu", seed: int = 0) -> Params:
    """nn.Linear default init (kaiming_uniform(a=sqrt 5) == U(-1/sqrt(in), 1/sqrt(in)))."""
    import re
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        p[name + "weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
        p[name + "bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * b

    if projector_type == "identity":
        return p
    if projector_type == "linear":
        lin("", H, C_tok)
        return p
    use_norm = "_Norm" in projector_type
    depth = int(re.match(r"^mlp(\d+)x_gelu$", projector_type.replace("_Norm", "")).group(1))
    idx = 0
    lin(f"{idx}.", H, C_tok); idx += 1
    if use_norm:
        p[f"{idx}.weight"] = torch.ones(H); p[f"{idx}.bias"] = torch.zeros(H); idx += 1
    for _ in range(1, depth):
        idx += 1
        lin(f"{idx}.", H, H); idx += 1
    return p


# ----------------------------------------------------------------------------------------
# Synthetic inputs: feature-injected mixtures and "Mondrian" images
# ----------------------------------------------------------------------------------------
def mog_features(N: int, C: int, G: int, sigma: float = 0.05, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(G, C, generator=g)
    lab = torch.randint(0, G, (N,), generator=g)
    return centres[lab] + sigma * torch.randn(N, C, generator=g)


def tie_noise(N: int, seed: int) -> torch.Tensor:
    """R3: the explicit stand-in for `torch.rand(density.shape)` at tokenizer.py:91."""
    return torch.rand(N, generator=torch.Generator().manual_seed(seed))


def well_posed_features(N: int, C: int, G: int, k: int, min_cluster_num: int, threshold: float = 0.5, seed0: int = 0,
                        min_margin: float = 2e-4, tries: int = 60):
    """Test helper: seeded (features, noise) for one image whose every integer decision in `dpc_knn` (after the
    position embedding is added) has a margin above `min_margin`, so that an implementation with a different
    but equally valid fp32 summation order must reproduce the integers exactly."""
    h = int(math.sqrt(N))
    pos = pos_encoding_2d(h, h, C).reshape(N, C)
    for t in range(tries):
        seed = seed0 + t
        f = mog_features(N, C, G, 0.05, seed) if G else torch.randn(N, C, generator=torch.Generator().manual_seed(seed))
        noise = tie_noise(N, seed + 5000)
        m = dpc_margins(f + pos, k, noise, threshold, min_cluster_num)
        if m["threshold_margin"] > min_margin and m["argmin_margin"] > min_margin:
            return f, noise
    raise RuntimeError("no well-posed seed found")

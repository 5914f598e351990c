"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32) of the reference's SetokDeTokenizer forward
(`/root/reference/src/model/setok/detokenizer.py:101-120`): `mapper_fc_in` -> Q-Former
(`module.py:151-207` embeddings, `:209-373` attention, `:376-388` self-output, `:447-474` FFN, `:476-583` layer,
`:586-690` encoder) -> `decoder_fc_in` -> + PositionalEncoding2D (`module.py:105-146`) -> timm ViT blocks -> LayerNorm.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this file; the product path
(`setok_b200/`) never does.

Pinning (tests/golden/detok.npz, written by oracle/make_golden.py): the Q-Former half is produced by the reference's
own `BertEmbeddings` + `BertEncoder` classes executed unmodified; `BertModel.forward`'s glue around them
(`module.py:852-1013`: all-ones self mask -> additive zeros; `invert_attention_mask` -> (1 - m) * finfo.min) is
restated because `BertModel.__init__` does not construct under transformers 5.x.  The decoder blocks are
`timm.models.vision_transformer.Block` (timm==0.9.16, `pyproject.toml:22`, absent from this image): the published
pre-LN block x += proj(MHSA(LN(x))); x += fc2(GELU(fc1(LN(x)))) is restated here and cross-checked in the golden
generator against HF `ViTLayer`, an independent implementation of the same block.

Deviations from the committed text, kept closed: (D1) the reference `forward` has no `return`; the oracle returns the
normalised decoder states.  (D2) `PositionalEncoding2D(hidden_dim)` emits 2*ceil(hidden_dim/4)*2 channels and the
reference adds it to a `decoder_embed_dim`-wide tensor, which only broadcasts when decoder_embed_dim <= that width;
the oracle requires it (the reference raises otherwise)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import setok_oracle as O

Params = Dict[str, torch.Tensor]


def make_detok_params(*, token_dim: int, hidden: int, q_heads: int, q_inter: int, q_layers: int, cross_freq: int, grid: int,
                      dec_dim: int, dec_depth: int, dec_mlp: int, seed: int = 0, init_range: float = 0.02) -> Params:
    """Seeded parameters under the reference's state_dict keys.  Linear layers xavier-uniform with zero bias
    (detokenizer.py:58-70) except the Q-Former, which keeps BERT's normal(0, 0.02) init; LayerNorms get a small random
    perturbation so that a test cannot pass with gamma/beta ignored."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, out_f, in_f, xavier=True):
        if xavier:
            a = math.sqrt(6.0 / (in_f + out_f))
            p[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a
        else:
            p[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * init_range
        p[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def ln(name, dim):
        p[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
        p[name + ".bias"] = 0.05 * torch.randn(dim, generator=g)

    Q = grid * grid
    p["mask_tokens"] = torch.randn(1, Q, hidden, generator=g) * init_range
    lin("mapper_fc_in", hidden, token_dim)
    lin("decoder_fc_in", dec_dim, hidden)
    ln("decoder_norm", dec_dim)
    ln("mapper.embeddings.LayerNorm", hidden)
    for i in range(q_layers):
        b = f"mapper.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(b + "attention.self." + n, hidden, hidden, xavier=False)
        lin(b + "attention.output.dense", hidden, hidden, xavier=False)
        ln(b + "attention.output.LayerNorm", hidden)
        if i % cross_freq == 0:
            for n in ("query", "key", "value"):
                lin(b + "crossattention.self." + n, hidden, hidden, xavier=False)
            lin(b + "crossattention.output.dense", hidden, hidden, xavier=False)
            ln(b + "crossattention.output.LayerNorm", hidden)
        lin(b + "intermediate_query.dense", q_inter, hidden, xavier=False)
        lin(b + "output_query.dense", hidden, q_inter, xavier=False)
        ln(b + "output_query.LayerNorm", hidden)
    for i in range(dec_depth):
        b = f"pixel_decoder.{i}."
        ln(b + "norm1", dec_dim)
        lin(b + "attn.qkv", 3 * dec_dim, dec_dim)
        lin(b + "attn.proj", dec_dim, dec_dim)
        ln(b + "norm2", dec_dim)
        lin(b + "mlp.fc1", dec_mlp, dec_dim)
        lin(b + "mlp.fc2", dec_dim, dec_mlp)
    return p


def _lin(x, p, name):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _ln(x, p, name, eps):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def _mha(q, k, v, heads, add_mask=None):
    """softmax(q k^T / sqrt(hd) + mask) v with (B, T, C) operands (module.py:267-373 / timm Attention)."""
    B, Tq, C = q.shape
    hd = C // heads
    qh = q.view(B, Tq, heads, hd).transpose(1, 2)
    kh = k.view(B, k.shape[1], heads, hd).transpose(1, 2)
    vh = v.view(B, v.shape[1], heads, hd).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(hd)
    if add_mask is not None:
        s = s + add_mask
    a = torch.softmax(s, dim=-1)
    return torch.matmul(a, vh).transpose(1, 2).reshape(B, Tq, C)


def qformer(p: Params, query_embeds: torch.Tensor, enc: torch.Tensor, enc_mask: torch.Tensor, *, heads: int, layers: int,
            cross_freq: int, eps: float = 1e-12) -> torch.Tensor:
    """BertModel(query_embeds=..., encoder_hidden_states=enc, encoder_attention_mask=enc_mask) (module.py:852-1013) with
    the text FFN deleted (detokenizer.py:94-96).  query_embeds (B, Q, H), enc (B, K, H), enc_mask (B, K) in {0, 1}."""
    h = _ln(query_embeds, p, "mapper.embeddings.LayerNorm", eps)                         # module.py:203-205
    inv = (1.0 - enc_mask[:, None, None, :].to(h.dtype)) * torch.finfo(h.dtype).min       # invert_attention_mask
    for i in range(layers):
        b = f"mapper.encoder.layer.{i}."
        ctx = _mha(_lin(h, p, b + "attention.self.query"), _lin(h, p, b + "attention.self.key"), _lin(h, p, b + "attention.self.value"), heads)
        a = _ln(_lin(ctx, p, b + "attention.output.dense") + h, p, b + "attention.output.LayerNorm", eps)      # module.py:383-387
        if i % cross_freq == 0:                                                                                 # module.py:484-493
            ctx = _mha(_lin(a, p, b + "crossattention.self.query"), _lin(enc, p, b + "crossattention.self.key"),
                       _lin(enc, p, b + "crossattention.self.value"), heads, inv)
            a = _ln(_lin(ctx, p, b + "crossattention.output.dense") + a, p, b + "crossattention.output.LayerNorm", eps)
        u = F.gelu(_lin(a, p, b + "intermediate_query.dense"))                                                  # module.py:579-582
        h = _ln(_lin(u, p, b + "output_query.dense") + a, p, b + "output_query.LayerNorm", eps)
    return h


def timm_block(p: Params, prefix: str, x: torch.Tensor, heads: int, eps: float = 1e-5) -> torch.Tensor:
    """timm 0.9.16 vision_transformer.Block with init_values=None, qk_norm=False, drop = 0 (eval)."""
    C = x.shape[-1]
    qkv = _lin(_ln(x, p, prefix + "norm1", eps), p, prefix + "attn.qkv")
    x = x + _lin(_mha(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads), p, prefix + "attn.proj")
    return x + _lin(F.gelu(_lin(_ln(x, p, prefix + "norm2", eps), p, prefix + "mlp.fc1")), p, prefix + "mlp.fc2")


def decoder_pos_table(hidden: int, grid: int, dec_dim: int) -> torch.Tensor:
    """PositionalEncoding2D(hidden_dim) evaluated on a (1, grid, grid, dec_dim) tensor (detokenizer.py:53, :113-116):
    the 2*ceil(hidden/4)*2 channels it builds, sliced to dec_dim (module.py:144)."""
    full = _pos_full(hidden, grid)
    if dec_dim > full.shape[1]:
        raise ValueError(f"decoder_embed_dim {dec_dim} exceeds the {full.shape[1]} channels PositionalEncoding2D({hidden}) emits (D2)")
    return full[:, :dec_dim].contiguous()


def _pos_full(hidden: int, grid: int) -> torch.Tensor:
    ch = int(math.ceil(hidden / 4) * 2)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
    pos = torch.arange(grid, dtype=inv_freq.dtype)
    s = torch.einsum("i,j->ij", pos, inv_freq)
    e = torch.flatten(torch.stack((s.sin(), s.cos()), dim=-1), -2, -1)
    emb = torch.zeros(grid, grid, 2 * ch)
    emb[:, :, :ch] = e.unsqueeze(1)
    emb[:, :, ch:] = e
    return emb.reshape(grid * grid, 2 * ch)


def detok_forward(p: Params, x: torch.Tensor, attention_masks: torch.Tensor, *, q_heads: int, q_layers: int, cross_freq: int, grid: int,
                  dec_heads: int, dec_depth: int, hidden: int, return_intermediates: bool = False):
    """SetokDeTokenizer.forward (detokenizer.py:101-120) + the missing return (D1).  x (B, K, C_tok) padded tokens,
    attention_masks (B, K) in {0, 1}.  Returns (B, grid^2, decoder_embed_dim)."""
    B = x.shape[0]
    mask_tokens = p["mask_tokens"].expand(B, -1, -1)                                       # :103
    enc = _lin(x, p, "mapper_fc_in")                                                      # :104
    h = qformer(p, mask_tokens, enc, attention_masks, heads=q_heads, layers=q_layers, cross_freq=cross_freq)   # :105-109
    y = _lin(h, p, "decoder_fc_in")                                                       # :111
    y = y + decoder_pos_table(hidden, grid, y.shape[-1])[None].to(y.dtype)                            # :112-115
    for i in range(dec_depth):                                                            # :117-118
        y = timm_block(p, f"pixel_decoder.{i}.", y, dec_heads)
    out = _ln(y, p, "decoder_norm", 1e-5)                                                 # :120
    if return_intermediates:
        return out, dict(qformer=h, enc=enc)
    return out


def pad_ragged(tokens: torch.Tensor, offsets: List[int]):
    """packed (sum K, C) + offsets -> the reference's padded (B, K_max, C) + mask (B, K_max)."""
    B = len(offsets) - 1
    kmax = max(offsets[b + 1] - offsets[b] for b in range(B))
    x = torch.zeros(B, kmax, tokens.shape[1], dtype=tokens.dtype)
    m = torch.zeros(B, kmax)
    for b in range(B):
        n = offsets[b + 1] - offsets[b]
        x[b, :n] = tokens[offsets[b]:offsets[b + 1]]
        m[b, :n] = 1
    return x, m

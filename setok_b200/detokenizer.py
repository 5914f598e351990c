"""Host-side mirror of the reference's ``SetokDeTokenizer`` (src/model/setok/detokenizer.py:14-120) on libsetok_b200.

Same constructor kwargs, ``forward(x, attention_masks)`` signature and ``state_dict`` keys as the reference
(``mask_tokens``, ``mapper_fc_in.*``, ``mapper.embeddings.LayerNorm.*``, ``mapper.encoder.layer.{i}.attention.self.{query,key,value}.*``,
``...attention.output.{dense,LayerNorm}.*``, ``...crossattention.*``, ``...intermediate_query.dense.*``, ``...output_query.{dense,LayerNorm}.*``,
``decoder_fc_in.*``, ``pixel_decoder.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*``, ``decoder_norm.*``,
``position_embedding.inv_freq``).  The torch modules only hold parameters; ``forward`` packs them once and calls
``setok_detok_forward``.  There is no PyTorch or CPU fallback.

The detokenizer is a *ragged consumer*: ``forward`` also accepts the tokenizer's ``RaggedTokens`` directly
(``detok(ragged)``), in which case nothing is padded; the reference's padded ``(B, K_max, C_tok)`` + ``attention_masks`` pair is
converted to packed rows + offsets on the device.

Differences from the committed reference text, kept closed: D1 ``forward`` returns the normalised
decoder states (the reference has no ``return``); D2 ``decoder_embed_dim`` must not exceed the channels
``PositionalEncoding2D(hidden_dim)`` emits (the reference raises a broadcast error there); the Q-Former config is BERT-base's
(``BertConfig()`` == bert-base-uncased's config.json, detokenizer.py:80) with ``hidden_size = hidden_dim`` heads of 64."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import SetokError
from ._pack import PackedParams
from .ragged import RaggedTokens
from .tokenizer import Attention, Mlp, PositionalEncoding2D, _bf16, _f32


class _TimmBlock(nn.Module):
    """Parameter container with timm 0.9.16 ``vision_transformer.Block``'s key names."""

    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=True)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))


class _BertSelfAttention(nn.Module):
    def __init__(self, hidden, kv_width):
        super().__init__()
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(kv_width, hidden)
        self.value = nn.Linear(kv_width, hidden)


class _BertSelfOutput(nn.Module):
    def __init__(self, hidden, eps):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)


class _BertAttention(nn.Module):
    def __init__(self, hidden, kv_width, eps):
        super().__init__()
        self.self = _BertSelfAttention(hidden, kv_width)
        self.output = _BertSelfOutput(hidden, eps)


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _BertOutput(nn.Module):
    def __init__(self, inter, hidden, eps):
        super().__init__()
        self.dense = nn.Linear(inter, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)


class _BertLayer(nn.Module):
    def __init__(self, hidden, inter, eps, has_cross):
        super().__init__()
        self.attention = _BertAttention(hidden, hidden, eps)
        self.has_cross_attention = has_cross
        if has_cross:
            self.crossattention = _BertAttention(hidden, hidden, eps)       # encoder_width = vision_width = hidden_dim (detokenizer.py:82)
        self.intermediate_query = _Dense(hidden, inter)
        self.output_query = _BertOutput(inter, hidden, eps)


class _BertEncoder(nn.Module):
    def __init__(self, hidden, inter, eps, layers, freq):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(hidden, inter, eps, i % freq == 0) for i in range(layers)])


class _BertEmbeddings(nn.Module):
    def __init__(self, hidden, eps, max_pos=512):
        super().__init__()
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)
        self.register_buffer("position_ids", torch.arange(max_pos).expand((1, -1)))


class _Mapper(nn.Module):
    def __init__(self, hidden, inter, eps, layers, freq):
        super().__init__()
        self.embeddings = _BertEmbeddings(hidden, eps)
        self.encoder = _BertEncoder(hidden, inter, eps, layers, freq)


class SetokDeTokenizer(PackedParams, nn.Module):
    def __init__(self, token_feat_dim: Optional[int] = 4096, hidden_dim: Optional[int] = 4096, patch_size: Optional[int] = 14,
                 image_size: Optional[int] = 256, decoder_embed_dim: Optional[int] = 4096, decoder_nheads: Optional[int] = 16,
                 proj_drop: Optional[float] = 0.2, attn_drop: Optional[float] = 0.2, decoder_depth: Optional[int] = 16,
                 norm_layer: nn.Module = nn.LayerNorm, mlp_ratio: Optional[float] = 4.0,
                 feature_mapper_path_or_name: Optional[str] = "bert-base-uncased", num_hidden_layers: Optional[int] = 6,
                 cross_attention_freq: Optional[int] = 2, initializer_range: Optional[float] = 0.02,
                 mapper_num_attention_heads: Optional[int] = None, mapper_intermediate_size: Optional[int] = None, **kwargs) -> None:
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise SetokError("setok_b200 implements the reference default only: norm_layer=nn.LayerNorm")
        self.token_feat_dim = token_feat_dim
        self.patch_size = patch_size
        self.height = self.weight = image_size // patch_size
        self.num_mask_token = self.height * self.weight
        self.hidden_dim = hidden_dim
        self.decoder_embed_dim = decoder_embed_dim
        self.decoder_nheads = decoder_nheads
        self.decoder_depth = decoder_depth
        self.mlp_ratio = mlp_ratio
        self.cross_attention_freq = cross_attention_freq
        # BertConfig() defaults = bert-base-uncased: 12 heads over hidden 768 (head_dim 64), intermediate 3072, eps 1e-12
        self.mapper_heads = mapper_num_attention_heads or max(1, hidden_dim // 64)
        self.mapper_inter = mapper_intermediate_size or 4 * hidden_dim
        if hidden_dim % self.mapper_heads:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)" % (hidden_dim, self.mapper_heads))
        pos_channels = int(math.ceil(hidden_dim / 4) * 2) * 2
        if decoder_embed_dim > pos_channels:
            raise ValueError(f"decoder_embed_dim {decoder_embed_dim} exceeds the {pos_channels} channels PositionalEncoding2D({hidden_dim}) "
                             "emits: the reference's `x + pos_emb` (detokenizer.py:115) cannot broadcast")
        query_tokens = nn.Parameter(torch.zeros(1, self.num_mask_token, self.hidden_dim))
        query_tokens.data.normal_(mean=0.0, std=initializer_range)
        self.mask_tokens = query_tokens
        self.mapper_fc_in = nn.Linear(self.token_feat_dim, self.hidden_dim)
        self.decoder_fc_in = nn.Linear(self.hidden_dim, self.decoder_embed_dim)
        self.decoder_norm = norm_layer(self.decoder_embed_dim)
        self.pixel_decoder = nn.ModuleList([_TimmBlock(self.decoder_embed_dim, decoder_nheads, mlp_ratio, norm_layer) for _ in range(decoder_depth)])
        self.position_embedding = PositionalEncoding2D(self.hidden_dim)
        self.initialize_weights()
        self.mapper = _Mapper(hidden_dim, self.mapper_inter, 1e-12, num_hidden_layers, cross_attention_freq)
        for m in self.mapper.modules():                       # BertPreTrainedModel._init_weights: normal(0, initializer_range)
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=initializer_range)
                m.bias.data.zero_()
        self._packed = None

    def initialize_weights(self):
        self.apply(self._init_weights)

    def _init_weights(self, m):                               # detokenizer.py:58-70
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def load_model(self):
        pass

    # -- packing -------------------------------------------------------------------------------
    @property
    def device(self):
        return self.mask_tokens.device

    def _pack(self):
        dev = self.device
        if dev.type != "cuda":
            raise SetokError("the detokenizer must live on a CUDA device (setok_b200 has no CPU path)")
        keep: Dict[str, torch.Tensor] = {}

        def put(name, t):
            keep[name] = t
            return t.data_ptr()

        nq = len(self.mapper.encoder.layer)
        qlayers = (_lib.QFormerLayer * max(nq, 1))()
        for i, L in enumerate(self.mapper.encoder.layer):
            s, o = L.attention.self, L.attention.output
            f = dict(w_qkv=put(f"q{i}.w_qkv", _bf16(torch.cat([s.query.weight, s.key.weight, s.value.weight], 0), dev)),
                     b_qkv=put(f"q{i}.b_qkv", _f32(torch.cat([s.query.bias, s.key.bias, s.value.bias], 0), dev)),
                     w_so=put(f"q{i}.w_so", _bf16(o.dense.weight, dev)), b_so=put(f"q{i}.b_so", _f32(o.dense.bias, dev)),
                     ln_s_g=put(f"q{i}.ln_s_g", _f32(o.LayerNorm.weight, dev)), ln_s_b=put(f"q{i}.ln_s_b", _f32(o.LayerNorm.bias, dev)),
                     has_cross=int(L.has_cross_attention),
                     w_f1=put(f"q{i}.w_f1", _bf16(L.intermediate_query.dense.weight, dev)), b_f1=put(f"q{i}.b_f1", _f32(L.intermediate_query.dense.bias, dev)),
                     w_f2=put(f"q{i}.w_f2", _bf16(L.output_query.dense.weight, dev)), b_f2=put(f"q{i}.b_f2", _f32(L.output_query.dense.bias, dev)),
                     ln_f_g=put(f"q{i}.ln_f_g", _f32(L.output_query.LayerNorm.weight, dev)), ln_f_b=put(f"q{i}.ln_f_b", _f32(L.output_query.LayerNorm.bias, dev)))
            if L.has_cross_attention:
                cs, co = L.crossattention.self, L.crossattention.output
                f.update(w_cq=put(f"q{i}.w_cq", _bf16(cs.query.weight, dev)), b_cq=put(f"q{i}.b_cq", _f32(cs.query.bias, dev)),
                         w_ckv=put(f"q{i}.w_ckv", _bf16(torch.cat([cs.key.weight, cs.value.weight], 0), dev)),
                         b_ckv=put(f"q{i}.b_ckv", _f32(torch.cat([cs.key.bias, cs.value.bias], 0), dev)),
                         w_co=put(f"q{i}.w_co", _bf16(co.dense.weight, dev)), b_co=put(f"q{i}.b_co", _f32(co.dense.bias, dev)),
                         ln_c_g=put(f"q{i}.ln_c_g", _f32(co.LayerNorm.weight, dev)), ln_c_b=put(f"q{i}.ln_c_b", _f32(co.LayerNorm.bias, dev)))
            for k_, v in f.items():
                setattr(qlayers[i], k_, v)
        nb = len(self.pixel_decoder)
        blocks = (_lib.VitLayer * max(nb, 1))()
        for i, Bk in enumerate(self.pixel_decoder):
            f = dict(w_qkv=put(f"d{i}.w_qkv", _bf16(Bk.attn.qkv.weight, dev)), b_qkv=put(f"d{i}.b_qkv", _f32(Bk.attn.qkv.bias, dev)),
                     w_o=put(f"d{i}.w_o", _bf16(Bk.attn.proj.weight, dev)), b_o=put(f"d{i}.b_o", _f32(Bk.attn.proj.bias, dev)),
                     w_fc1=put(f"d{i}.w_fc1", _bf16(Bk.mlp.fc1.weight, dev)), b_fc1=put(f"d{i}.b_fc1", _f32(Bk.mlp.fc1.bias, dev)),
                     w_fc2=put(f"d{i}.w_fc2", _bf16(Bk.mlp.fc2.weight, dev)), b_fc2=put(f"d{i}.b_fc2", _f32(Bk.mlp.fc2.bias, dev)),
                     ln1_g=put(f"d{i}.ln1_g", _f32(Bk.norm1.weight, dev)), ln1_b=put(f"d{i}.ln1_b", _f32(Bk.norm1.bias, dev)),
                     ln2_g=put(f"d{i}.ln2_g", _f32(Bk.norm2.weight, dev)), ln2_b=put(f"d{i}.ln2_b", _f32(Bk.norm2.bias, dev)))
            for k_, v in f.items():
                setattr(blocks[i], k_, v)
        # PositionalEncoding2D(hidden_dim) on a (1, h, w, decoder_embed_dim) tensor: all 2*channels columns, sliced (module.py:144)
        pe = self.position_embedding
        saved = pe.org_channels
        pe.org_channels = self.decoder_embed_dim
        try:
            pe._tables.pop((self.height, self.weight, str(dev)), None)
            pos = pe.table(self.height, self.weight, dev).clone()
            pe._tables.pop((self.height, self.weight, str(dev)), None)
        finally:
            pe.org_channels = saved
        d = _lib.Detok(token_dim=self.token_feat_dim, hidden=self.hidden_dim, q_heads=self.mapper_heads, q_inter=self.mapper_inter,
                       q_layers=nq, grid=self.height, dec_dim=self.decoder_embed_dim, dec_heads=self.decoder_nheads,
                       dec_mlp=int(self.decoder_embed_dim * self.mlp_ratio), dec_depth=nb, q_ln_eps=1e-12, dec_ln_eps=1e-5,
                       w_map_in=put("w_map_in", _bf16(self.mapper_fc_in.weight, dev)), b_map_in=put("b_map_in", _f32(self.mapper_fc_in.bias, dev)),
                       mask_tokens=put("mask_tokens", _f32(self.mask_tokens[0], dev)),
                       emb_ln_g=put("emb_ln_g", _f32(self.mapper.embeddings.LayerNorm.weight, dev)),
                       emb_ln_b=put("emb_ln_b", _f32(self.mapper.embeddings.LayerNorm.bias, dev)), qlayer=qlayers,
                       w_dec_in=put("w_dec_in", _bf16(self.decoder_fc_in.weight, dev)), b_dec_in=put("b_dec_in", _f32(self.decoder_fc_in.bias, dev)),
                       pos=put("pos", pos), block=blocks, norm_g=put("norm_g", _f32(self.decoder_norm.weight, dev)),
                       norm_b=put("norm_b", _f32(self.decoder_norm.bias, dev)))
        self._packed = (d, qlayers, blocks, keep)
        return self._packed

    # -- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, attention_masks: Optional[torch.Tensor] = None, out_dtype=None) -> torch.Tensor:
        """``x``: the reference's padded tokens (B, K_max, C_tok) with ``attention_masks`` (B, K_max) in {0, 1}, or a
        ``RaggedTokens`` (packed rows + offsets, nothing padded).  Returns (B, (image_size/patch)^2, decoder_embed_dim)."""
        d = self._packed_get()[0]
        dev = self.device
        if isinstance(x, RaggedTokens):
            tokens, offsets = x.data, x.offsets
            B = int(offsets.numel()) - 1
        else:
            if attention_masks is None:
                attention_masks = torch.ones(x.shape[:2], device=x.device)
            x = x.to(dev)
            m = attention_masks.to(dev) > 0
            B = x.shape[0]
            counts = m.sum(dim=1).to(torch.int32)
            offsets = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            offsets[1:] = torch.cumsum(counts, 0)
            tokens = x[m]                                   # packed rows in image order (a valid prefix per image is not assumed)
        if tokens.dtype not in (torch.float32, torch.bfloat16):
            tokens = tokens.float()
        tokens = tokens.contiguous()
        if not tokens.is_cuda:
            raise SetokError("setok_b200 kernels need CUDA tensors; there is no CPU fallback")
        if tokens.shape[1] != self.token_feat_dim:
            raise SetokError(f"token width {tokens.shape[1]} != token_feat_dim {self.token_feat_dim}")
        cap = max(int(tokens.shape[0]), 1)
        if tokens.shape[0] == 0:
            tokens = torch.zeros(1, self.token_feat_dim, dtype=tokens.dtype, device=dev)
        offsets = offsets.to(device=dev, dtype=torch.int32).contiguous()
        out_dtype = out_dtype or (torch.bfloat16 if tokens.dtype == torch.bfloat16 else torch.float32)
        out = torch.empty(B, self.num_mask_token, self.decoder_embed_dim, dtype=out_dtype, device=dev)
        lib = _lib.load()
        nbytes = lib.setok_detok_workspace_bytes(C.byref(d), B, cap)
        ws = ops.workspace(dev, nbytes, "detok")
        with torch.cuda.device(dev):
            st = lib.setok_detok_forward(C.byref(d), tokens.data_ptr(), ops._dt(tokens), offsets.data_ptr(), B, cap, out.data_ptr(),
                                         ops._dt(out), ws.data_ptr(), ws.numel(), ops._stream(dev))
        _lib.check(st, "setok_detok_forward")
        return out

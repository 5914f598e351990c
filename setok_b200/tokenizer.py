"""Host-side mirror of the reference's tokenizer plugin interface, running on libsetok_b200.

Same names, constructor kwargs, call signatures and ``state_dict`` keys as the reference:

* ``SetokTokenizer``          <- src/model/setok/tokenizer.py:13-182
* ``CLIPVisionTower``         <- src/model/setok/clip_encoder.py:8-93
* ``Block/Attention/Mlp``     <- src/model/setok/module.py:29-100 (parameter containers; the math
  runs in the CUDA library)
* ``PositionalEncoding2D``    <- src/model/setok/module.py:105-146

The torch modules below only *hold parameters* under the reference's names so that the Setokim
pipeline's ``load_state_dict`` / ``.to()`` / ``requires_grad_`` calls work unchanged; ``forward`` packs
them once into the layouts the kernels want (bf16 matrices, fp32 vectors, fused qkv) and calls the
C ABI.  There is no PyTorch or CPU fallback: calling ``forward`` off-GPU raises ``SetokError``.

Repairs to the committed reference that this interface bakes in: R1 the batch is
processed as a batch (the committed forward only works per image), R2 the inter-cluster encoder sees
each image's tokens as one sequence, R3 the tie-break noise of tokenizer.py:91 can be passed in
(``noise=``; default: ``torch.rand`` on the module's device, seeded by the global CUDA RNG), R4 the
tower is a CLIP-style ViT with a CLS token.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import SetokError
from ._pack import PackedParams, fold_layernorm_into_linear, stamp
from .ragged import RaggedTokens


# --------------------------------------------------------------------------------------------
# parameter containers with the reference's module tree / key names
# --------------------------------------------------------------------------------------------
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(nn.Module):
    """depth x [shared norm1 -> Attention] then norm2 -> Mlp.  ``layers.i.0`` aliases ``norm1`` exactly as
    in the reference (module.py:86-91), so the state_dict carries the same duplicate keys."""

    def __init__(self, dim, num_heads, mlp_hidden_dim, qkv_bias=True, qk_scale=None, proj_drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, depth=0):
        super().__init__()
        if act_layer is not nn.GELU or norm_layer is not nn.LayerNorm:
            raise SetokError("setok_b200 implements the reference defaults only: act_layer=nn.GELU, norm_layer=nn.LayerNorm")
        self.dim, self.num_heads, self.mlp_hidden_dim, self.depth = dim, num_heads, mlp_hidden_dim, depth
        self.norm1 = norm_layer(dim)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.layers = nn.ModuleList()
        for _ in range(depth):
            self.layers.append(nn.Sequential(self.norm1, Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                                                   attn_drop=attn_drop, proj_drop=proj_drop), self.drop_path))
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=proj_drop)


class PositionalEncoding2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.org_channels = channels
        channels = int(np.ceil(channels / 4) * 2)
        self.channels = channels
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
        self.register_buffer("inv_freq", inv_freq)
        self._tables: Dict[tuple, torch.Tensor] = {}

    def table(self, h: int, w: int, device) -> torch.Tensor:
        """(h*w, C) float32 table on `device`, computed once per shape with the reference's own torch ops on the
        CPU (module.py:131-145) so that it is bit-identical to the reference's ``cached_penc``."""
        key = (h, w, str(device))
        t = self._tables.get(key)
        if t is None:
            inv = self.inv_freq.detach().float().cpu()
            sx = torch.einsum("i,j->ij", torch.arange(h, dtype=inv.dtype), inv)
            sy = torch.einsum("i,j->ij", torch.arange(w, dtype=inv.dtype), inv)
            ex = torch.flatten(torch.stack((sx.sin(), sx.cos()), dim=-1), -2, -1).unsqueeze(1)
            ey = torch.flatten(torch.stack((sy.sin(), sy.cos()), dim=-1), -2, -1)
            emb = torch.zeros((h, w, self.channels * 2), dtype=torch.float32)
            emb[:, :, :self.channels] = ex
            emb[:, :, self.channels:2 * self.channels] = ey
            t = emb[:, :, :self.org_channels].reshape(h * w, self.org_channels).contiguous().to(device)
            self._tables[key] = t
        return t


# --------------------------------------------------------------------------------------------
# vision tower
# --------------------------------------------------------------------------------------------
def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _bf16(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


class CLIPVisionTower(PackedParams, nn.Module):
    """Frozen CLIP-style ViT behind the reference wrapper's interface (clip_encoder.py:8-93).

    ``vision_tower`` is a checkpoint name/path for ``transformers`` (as in the reference) or, when no
    checkpoint is reachable, ``vision_config`` (a ``CLIPVisionConfig`` or dict) builds the same architecture
    with seeded random weights.  HF's ``CLIPVisionModel`` is used purely as the parameter container (HF key
    names); its forward is never called."""

    def __init__(self, vision_tower: Optional[str], unfreeze_mm_vision_tower: Optional[bool] = False,
                 mm_vision_select_feature: Optional[str] = "patch", mm_vision_select_layer: Optional[int] = -2,
                 delay_load=False, vision_config=None, residual_f32: bool = True, ln_fold: Optional[bool] = None):
        super().__init__()
        self.is_loaded = False
        # residual stream of the tower kept in float32 between layers (SETOK_VIT_RESIDUAL_F32); False = bf16 stream
        self.residual_f32 = bool(residual_f32)
        # LayerNorms folded into the GEMMs around them (SETOK_VIT_LN_FOLD, include/setok_b200.h): on by default with the f32
        # residual stream when the widths allow it; SETOK_VIT_LN_FOLD=0 in the environment keeps the separate LayerNorm passes
        self.ln_fold = (os.environ.get("SETOK_VIT_LN_FOLD", "1") != "0") if ln_fold is None else bool(ln_fold)
        self.vision_tower_name = vision_tower
        self.select_layer = mm_vision_select_layer
        self.select_feature = mm_vision_select_feature
        self._vision_config = vision_config
        self._packed = None
        self.image_processor = None
        if not delay_load or unfreeze_mm_vision_tower:
            self.load_model()
        else:
            self.cfg_only = self._resolve_config()

    def _resolve_config(self):
        from transformers import CLIPVisionConfig
        cfg = self._vision_config
        if cfg is None:
            from transformers import AutoConfig
            cfg = AutoConfig.from_pretrained(self.vision_tower_name)
            cfg = getattr(cfg, "vision_config", cfg)
        elif isinstance(cfg, dict):
            cfg = CLIPVisionConfig(**cfg)
        return cfg

    def load_model(self, device_map=None):
        if self.is_loaded:
            print("{} is already loaded, `load_model` called again, skipping.".format(self.vision_tower_name))
            return
        from transformers import CLIPVisionModel
        if self._vision_config is not None:
            self.vision_tower = CLIPVisionModel(self._resolve_config())
            self.image_processor = _default_image_processor(self.vision_tower.config)
        else:
            from transformers import AutoProcessor
            self.image_processor = AutoProcessor.from_pretrained(self.vision_tower_name)
            self.vision_tower = CLIPVisionModel.from_pretrained(self.vision_tower_name, device_map=device_map)
        cfg = self.vision_tower.config
        if getattr(cfg, "hidden_act", "quick_gelu") != "quick_gelu":
            raise SetokError(f"tower activation {cfg.hidden_act!r} unsupported (CLIP quick_gelu only)")
        self.vision_tower.requires_grad_(False)
        self.vision_tower.eval()
        self.is_loaded = True
        self._packed = None

    # -- the packed copy is validated against the live parameters before every forward (setok_b200/_pack.py)
    def _pack_sources(self):
        return (self.vision_tower,)

    def _pack(self):
        cfg = self.config
        dev = self.device
        if dev.type != "cuda":
            raise SetokError("the vision tower must live on a CUDA device (setok_b200 has no CPU path)")
        sd = self.vision_tower.state_dict()
        pre = "vision_model."
        Cc, p = cfg.hidden_size, cfg.patch_size
        Kp = (3 * p * p + 63) // 64 * 64
        keep: Dict[str, torch.Tensor] = {}
        w32 = _f32(sd[pre + "embeddings.patch_embedding.weight"].reshape(Cc, -1), dev)
        if self.residual_f32:
            # high-precision mode: [w_hi | w_hi | w_lo] for the 3-term split patch embedding (SETOK_VIT_PATCH_SPLIT)
            wp = torch.zeros(Cc, 3 * Kp, dtype=torch.bfloat16, device=dev)
            w_hi = w32.to(torch.bfloat16)
            wp[:, :3 * p * p] = w_hi
            wp[:, Kp:Kp + 3 * p * p] = w_hi
            wp[:, 2 * Kp:2 * Kp + 3 * p * p] = (w32 - w_hi.float()).to(torch.bfloat16)
        else:
            wp = torch.zeros(Cc, Kp, dtype=torch.bfloat16, device=dev)
            wp[:, :3 * p * p] = w32.to(torch.bfloat16)
        keep["w_patch"] = wp
        keep["cls"] = _f32(sd[pre + "embeddings.class_embedding"], dev)
        keep["pos"] = _f32(sd[pre + "embeddings.position_embedding.weight"], dev)
        keep["pre_g"] = _f32(sd[pre + "pre_layrnorm.weight"], dev)
        keep["pre_b"] = _f32(sd[pre + "pre_layrnorm.bias"], dev)
        L = cfg.num_hidden_layers
        layers = (_lib.VitLayer * max(L, 1))()
        fold = self.ln_fold and self.residual_f32 and Cc % 32 == 0 and cfg.intermediate_size % 32 == 0
        for i in range(L):
            q = f"{pre}encoder.layers.{i}."
            t = {
                "w_qkv": _f32(torch.cat([sd[q + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0), dev),
                "b_qkv": _f32(torch.cat([sd[q + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0), dev),
                "w_o": _bf16(sd[q + "self_attn.out_proj.weight"], dev), "b_o": _f32(sd[q + "self_attn.out_proj.bias"], dev),
                "w_fc1": _f32(sd[q + "mlp.fc1.weight"], dev), "b_fc1": _f32(sd[q + "mlp.fc1.bias"], dev),
                "w_fc2": _bf16(sd[q + "mlp.fc2.weight"], dev), "b_fc2": _f32(sd[q + "mlp.fc2.bias"], dev),
                "ln1_g": _f32(sd[q + "layer_norm1.weight"], dev), "ln1_b": _f32(sd[q + "layer_norm1.bias"], dev),
                "ln2_g": _f32(sd[q + "layer_norm2.weight"], dev), "ln2_b": _f32(sd[q + "layer_norm2.bias"], dev),
            }
            for w_, b_, g_, be_, s_ in (("w_qkv", "b_qkv", "ln1_g", "ln1_b", "s_qkv"), ("w_fc1", "b_fc1", "ln2_g", "ln2_b", "s_fc1")):
                w32_ = t[w_]
                if fold:
                    # LN(x) W^T + b = rho (x - mu) (gamma (.) W)^T + (W beta + b): the GEMM runs on W' = bf16(gamma (.) W), its
                    # epilogue needs s_n = sum_k W'_nk (of the ROUNDED matrix: the mean correction cancels exactly) and t = W beta + b
                    t[w_], t[s_], t[b_] = fold_layernorm_into_linear(w32_, t[b_], t[g_], t[be_])
                else:
                    t[w_] = w32_.to(torch.bfloat16).contiguous()
            for k_, v in t.items():
                keep[f"l{i}.{k_}"] = v
                setattr(layers[i], k_, v.data_ptr())
        vit = _lib.Vit(image_size=cfg.image_size, patch=p, hidden=Cc, heads=cfg.num_attention_heads, layers=L,
                       mlp=cfg.intermediate_size, ln_eps=float(cfg.layer_norm_eps), w_patch=wp.data_ptr(),
                       cls=keep["cls"].data_ptr(), pos=keep["pos"].data_ptr(), pre_ln_g=keep["pre_g"].data_ptr(),
                       pre_ln_b=keep["pre_b"].data_ptr(), layer=layers,
                       flags=((_lib.VIT_RESIDUAL_F32 | _lib.VIT_PATCH_SPLIT) if self.residual_f32 else 0) | (_lib.VIT_LN_FOLD if fold else 0))
        self._packed = (vit, layers, keep)
        self._resized = {}
        return self._packed

    def _vit_for_size(self, size: int):
        """The packed tower for `size`^2 inputs.  Other than the native size needs `interpolate_pos_encoding`: the position
        table is resized once per size exactly as HF does (bicubic on the patch grid, class row kept;
        modeling_clip.py:160-196) and cached; all kernels take the grid size as a run-time argument."""
        vit, layers, keep = self._packed_get()
        if size == vit.image_size:
            return vit
        hit = self._resized.get(size)
        if hit is None:
            if size % vit.patch != 0:
                raise SetokError(f"image size {size} is not a multiple of the patch size {vit.patch}")
            pos = keep["pos"].detach().float().cpu()
            g = int((pos.shape[0] - 1) ** 0.5)
            ng = size // vit.patch
            grid = pos[1:].reshape(1, g, g, -1).permute(0, 3, 1, 2)
            grid = torch.nn.functional.interpolate(grid, size=(ng, ng), mode="bicubic", align_corners=False)
            new_pos = torch.cat([pos[:1], grid.permute(0, 2, 3, 1).reshape(ng * ng, -1)], 0).contiguous().to(keep["pos"].device)
            v2 = _lib.Vit(image_size=size, patch=vit.patch, hidden=vit.hidden, heads=vit.heads, layers=vit.layers, mlp=vit.mlp,
                          ln_eps=vit.ln_eps, w_patch=vit.w_patch, cls=vit.cls, pos=new_pos.data_ptr(), pre_ln_g=vit.pre_ln_g,
                          pre_ln_b=vit.pre_ln_b, layer=layers, flags=vit.flags)
            hit = (v2, new_pos)
            self._resized[size] = hit
        return hit[0]

    def u8_norm(self) -> "_lib.U8Norm":
        """Constants of the tower's image processor for the uint8 path (CLIPImageProcessor.rescale / .normalize of
        transformers 4.46.3): lut[v] = float32(float64(v) * rescale_factor), float32 mean / std."""
        ip = self.image_processor
        mean = list(getattr(ip, "image_mean", None) or (0.48145466, 0.4578275, 0.40821073))
        std = list(getattr(ip, "image_std", None) or (0.26862954, 0.26130258, 0.27577711))
        rescale = float(getattr(ip, "rescale_factor", None) or 1.0 / 255.0)
        n = _lib.U8Norm()
        lut = (np.arange(256, dtype=np.float64) * rescale).astype(np.float32)
        for v in range(256):
            n.lut[v] = float(lut[v])
        for c in range(3):
            n.mean[c] = float(np.float32(mean[c]))
            n.std[c] = float(np.float32(std[c]))
        return n

    def layers_to_run(self) -> int:
        L = self.config.num_hidden_layers
        idx = self.select_layer if self.select_layer >= 0 else L + 1 + self.select_layer
        if not 0 <= idx <= L:
            raise IndexError(f"mm_vision_select_layer {self.select_layer} out of range for {L} layers")
        return idx

    @torch.no_grad()
    def forward(self, images, interpolate_pos_encoding: bool = False, pos_embedding=None):
        """clip_encoder.py:50-62.  ``pos_embedding`` (a PositionalEncoding2D, 'patch' features only): the position-embedding
        add of tokenizer.py:164-169 is fused into the tower's last row pass and the result is the float32 tensor
        ``features + pos`` (what tokenizer.py:168 calls x) instead of the features."""
        if self.select_feature not in ("patch", "cls_patch"):
            raise ValueError(f"Unexpected select feature: {self.select_feature}")
        if type(images) is list:       # clip_encoder.py:52-57: per-image loop -> list of (1, N, C)
            return [self.forward(im.unsqueeze(0), interpolate_pos_encoding, pos_embedding) for im in images]
        if not self.is_loaded:
            raise SetokError("vision tower not loaded: call load_model() first")
        vit, _, _ = self._packed_get()
        dev = self.device
        if images.dim() != 4 or images.shape[1] != 3:
            raise SetokError(f"images must be (B, 3, H, W); got {tuple(images.shape)}")
        if images.shape[2] != vit.image_size or images.shape[3] != vit.image_size:
            if not interpolate_pos_encoding or images.shape[2] != images.shape[3]:      # HF raises likewise (modeling_clip.py:204-207)
                raise ValueError(f"Input image size ({images.shape[2]}*{images.shape[3]}) doesn't match model "
                                 f"({vit.image_size}*{vit.image_size}).")
            vit = self._vit_for_size(int(images.shape[2]))
        out_dtype = images.dtype
        x = images.to(device=dev)
        is_u8 = x.dtype == torch.uint8
        if is_u8:
            out_dtype = torch.float32       # raw pixels: the processor's float32 output dtype (mm_utils.py:166-182)
        elif x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.float32)
        x = x.contiguous()
        B = x.shape[0]
        keep_cls = 1 if self.select_feature == "cls_patch" else 0
        N = (vit.image_size // vit.patch) ** 2 + keep_cls
        feat_dtype = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32
        lib = _lib.load()
        nbytes = lib.setok_vit_workspace_bytes(C.byref(vit), B)
        ws = ops.workspace(dev, nbytes, "vit")
        if is_u8:
            # uint8 pixels (B, 3, H, W) already at the tower's resolution: rescale + normalize run inside the patch-embedding
            # im2col with the processor's constants, so the upload is 1 byte per pixel
            if pos_embedding is not None and keep_cls:
                raise SetokError("the fused position-embedding add needs mm_vision_select_feature='patch'")
            norm = self.u8_norm()
            pos = None
            if pos_embedding is not None:
                g = vit.image_size // vit.patch
                pos = pos_embedding.table(g, g, dev)
            feats = torch.empty(B, N, vit.hidden, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                st = lib.setok_vit_forward_u8(C.byref(vit), x.data_ptr(), C.byref(norm), B, self.layers_to_run(), keep_cls, ops._p(pos),
                                              feats.data_ptr(), ops._dt(feats), ws.data_ptr(), ws.numel(), ops._stream(dev))
            _lib.check(st, "setok_vit_forward_u8")
            return feats
        if pos_embedding is not None:
            if keep_cls:
                raise SetokError("the fused position-embedding add needs mm_vision_select_feature='patch'")
            g = vit.image_size // vit.patch
            pos = pos_embedding.table(g, g, dev)
            x_pos = torch.empty(B, N, vit.hidden, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                st = lib.setok_vit_forward_pos(C.byref(vit), x.data_ptr(), ops._dt(x), B, self.layers_to_run(), pos.data_ptr(),
                                               x_pos.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(dev))
            _lib.check(st, "setok_vit_forward_pos")
            return x_pos
        feats = torch.empty(B, N, vit.hidden, dtype=feat_dtype, device=dev)
        with torch.cuda.device(dev):
            st = lib.setok_vit_forward(C.byref(vit), x.data_ptr(), ops._dt(x), B, self.layers_to_run(), keep_cls, feats.data_ptr(),
                                       ops._dt(feats), ws.data_ptr(), ws.numel(), ops._stream(dev))
        _lib.check(st, "setok_vit_forward")
        return feats if feats.dtype == out_dtype else feats.to(out_dtype)

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self.vision_tower.dtype

    @property
    def device(self):
        return self.vision_tower.device

    @property
    def config(self):
        return self.vision_tower.config if self.is_loaded else self.cfg_only

    @property
    def hidden_size(self):
        return self.config.hidden_size

    @property
    def num_patches_per_side(self):
        return self.config.image_size // self.config.patch_size

    @property
    def num_patches(self):
        return (self.config.image_size // self.config.patch_size) ** 2


def _default_image_processor(cfg):
    try:
        from transformers import CLIPImageProcessor
        return CLIPImageProcessor(size={"shortest_edge": cfg.image_size}, crop_size={"height": cfg.image_size, "width": cfg.image_size})
    except Exception:   # pragma: no cover - optional vision deps
        return None


# --------------------------------------------------------------------------------------------
# tokenizer
# --------------------------------------------------------------------------------------------
class SetokTokenizer(PackedParams, nn.Module):
    def __init__(self, vision_tower: str = "google/siglip-so400m-patch14-384", unfreeze_mm_vision_tower: Optional[bool] = False,
                 mm_vision_select_feature: Optional[str] = "patch", mm_vision_select_layer: Optional[int] = -2,
                 delay_load: Optional[bool] = False, hidden_dim: Optional[int] = 4096, token_feat_dim: Optional[int] = 4096,
                 min_cluster_num: Optional[int] = 64, threshold: Optional[float] = 0.5, nheads: Optional[int] = 2,
                 dim_feedforward: Optional[int] = 4096, proj_drop: Optional[float] = 0.2, drop_path: Optional[float] = 0.0,
                 inner_cluster_layers: Optional[int] = 2, intra_cluster_layers: Optional[int] = 2, attn_drop: Optional[float] = 0.0,
                 act_layer: nn.Module = nn.GELU, norm_layer: nn.Module = nn.LayerNorm, **kwargs) -> None:
        super().__init__()
        self.hidden_dim = hidden_dim
        self.token_feat_dim = token_feat_dim
        self.nheads = nheads
        self.dim_feedforward = dim_feedforward
        self.inner_encoder = Block(hidden_dim, nheads, dim_feedforward, proj_drop=proj_drop, attn_drop=attn_drop, drop_path=drop_path,
                                   act_layer=act_layer, norm_layer=norm_layer, depth=inner_cluster_layers)
        self.inter_encoder = Block(hidden_dim, nheads, dim_feedforward, proj_drop=proj_drop, attn_drop=attn_drop, drop_path=drop_path,
                                   act_layer=act_layer, norm_layer=norm_layer, depth=intra_cluster_layers)
        self.position_embedding = PositionalEncoding2D(hidden_dim)
        self.out = nn.Linear(hidden_dim, token_feat_dim)
        self.min_cluster_num = min_cluster_num
        self.threshold = threshold
        self.initialize_weights()
        self.image_feature_encoder = CLIPVisionTower(vision_tower, unfreeze_mm_vision_tower=unfreeze_mm_vision_tower,
                                                     mm_vision_select_feature=mm_vision_select_feature,
                                                     mm_vision_select_layer=mm_vision_select_layer, delay_load=delay_load,
                                                     vision_config=kwargs.get("vision_config"),
                                                     residual_f32=kwargs.get("tower_residual_f32", True),
                                                     ln_fold=kwargs.get("tower_ln_fold"))
        self.image_processor = self.image_feature_encoder.image_processor
        self.eval()

    # -- reference surface -------------------------------------------------------------------
    def initialize_weights(self):
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
            if m.weight is not None:
                nn.init.constant_(m.weight, 1.0)

    @property
    def dtype(self):          # the reference's property is broken (tokenizer.py:74-76 reads self.Linear)
        return self.out.weight.dtype

    @property
    def device(self):
        return self.out.weight.device

    @property
    def is_loaded(self):
        return self.image_feature_encoder.is_loaded

    def load_model(self, device_map=None):
        self.image_feature_encoder.load_model(device_map=device_map)
        self.image_processor = self.image_feature_encoder.image_processor

    def _pack_sources(self):
        return (self.inner_encoder, self.inter_encoder, self.out)

    def invalidate(self):
        """Forces a re-pack of the kernels' parameter copies.  Not needed after ordinary in-place updates, ``.to()`` or any
        ``load_state_dict`` (the copies are validated against the parameters' version counters before every forward)."""
        self._packed = None
        self.image_feature_encoder.invalidate()

    # -- packing -----------------------------------------------------------------------------
    def _pack_block(self, blk: Block, keep: dict, tag: str, dev):
        attn = (_lib.Attn * max(blk.depth, 1))()
        for i in range(blk.depth):
            a = blk.layers[i][1]
            t = {"w_qkv": _bf16(a.qkv.weight, dev), "b_qkv": _f32(a.qkv.bias, dev), "w_proj": _bf16(a.proj.weight, dev), "b_proj": _f32(a.proj.bias, dev)}
            for k_, v in t.items():
                keep[f"{tag}.a{i}.{k_}"] = v
                setattr(attn[i], k_, v.data_ptr())
        t = {"n1_g": _f32(blk.norm1.weight, dev), "n1_b": _f32(blk.norm1.bias, dev), "n2_g": _f32(blk.norm2.weight, dev),
             "n2_b": _f32(blk.norm2.bias, dev), "w_fc1": _bf16(blk.mlp.fc1.weight, dev), "b_fc1": _f32(blk.mlp.fc1.bias, dev),
             "w_fc2": _bf16(blk.mlp.fc2.weight, dev), "b_fc2": _f32(blk.mlp.fc2.bias, dev)}
        for k_, v in t.items():
            keep[f"{tag}.{k_}"] = v
        keep[f"{tag}.attn"] = attn
        return _lib.Block(depth=blk.depth, attn=attn, **{k_: v.data_ptr() for k_, v in t.items()})

    def _pack(self):
        dev = self.device
        if dev.type != "cuda":
            raise SetokError("SetokTokenizer must live on a CUDA device (setok_b200 has no CPU path)")
        keep: dict = {}
        inner = self._pack_block(self.inner_encoder, keep, "inner", dev)
        inter = self._pack_block(self.inter_encoder, keep, "inter", dev)
        keep["w_out"], keep["b_out"] = _bf16(self.out.weight, dev), _f32(self.out.bias, dev)
        head = _lib.Head(hidden=self.hidden_dim, heads=self.nheads, mlp=self.dim_feedforward, token_dim=self.token_feat_dim,
                         inner=inner, inter=inter, w_out=keep["w_out"].data_ptr(), b_out=keep["b_out"].data_ptr())
        self._packed = (head, keep)
        return self._packed

    # -- forward -----------------------------------------------------------------------------
    @torch.no_grad()
    def encode_features(self, feats: torch.Tensor, k=None, threshold=None, token_mask=None, noise: Optional[torch.Tensor] = None,
                        token_dtype=None, return_group_features: bool = False, embedded: bool = False):
        """The head of tokenizer.py:162-182 for a batch of tower features (B, N, C).  Returns
        (RaggedTokens, idx_cluster (B, N) int64, score (B, 1, N)).  ``embedded``: `feats` is already
        ``features + pos`` in float32 (``image_feature_encoder(images, pos_embedding=...)``)."""
        head, _ = self._packed_get()
        dev = self.device
        if feats.dim() == 2:
            feats = feats.unsqueeze(0)
        B, N, Cc = feats.shape
        if Cc != self.hidden_dim:
            raise SetokError(f"tower width {Cc} != hidden_dim {self.hidden_dim} (the reference has no projection between them)")
        h = w = int(math.sqrt(N))                                        # tokenizer.py:164
        if h * w != N:
            raise SetokError(f"N={N} is not a square grid")
        _threshold = threshold if threshold else self.threshold            # tokenizer.py:171 (0 is falsy, as in the reference)
        _k = k if k else self.min_cluster_num                              # tokenizer.py:172
        feats = feats.to(dev)
        if feats.dtype not in (torch.float32, torch.bfloat16):
            feats = feats.float()
        feats = feats.contiguous()
        if noise is None:
            noise = torch.rand(B, N, device=dev, dtype=torch.float32)      # tokenizer.py:91
        else:
            noise = noise.to(device=dev, dtype=torch.float32).reshape(B, N).contiguous()
        if token_mask is not None:
            token_mask = token_mask.to(dev).reshape(B, N)
        if embedded:
            if feats.dtype != torch.float32:
                raise SetokError("embedded features must be float32 (the head's residual stream starts from them)")
            x_pos, idx, score, down, numc, offs = ops.dpc_cluster(feats, noise, (h, w), int(_k), float(_threshold), int(self.min_cluster_num),
                                                                  token_mask=token_mask, embedded=True)
        else:
            pos = self.position_embedding.table(h, w, dev)
            x_pos, idx, score, down, numc, offs = ops.dpc_cluster(feats, noise, (h, w), int(_k), float(_threshold), int(self.min_cluster_num),
                                                                  pos_table=pos, token_mask=token_mask)
        token_dtype = token_dtype or (torch.bfloat16 if feats.dtype == torch.bfloat16 else torch.float32)
        tokens = torch.empty(B * N, self.token_feat_dim, dtype=token_dtype, device=dev)
        gf = torch.empty(B * N, Cc, dtype=torch.float32, device=dev) if return_group_features else None
        lib = _lib.load()
        nbytes = lib.setok_head_workspace_bytes(C.byref(head), B, N)
        ws = ops.workspace(dev, nbytes, "head")
        with torch.cuda.device(dev):
            st = lib.setok_head_forward(C.byref(head), x_pos.data_ptr(), idx.data_ptr(), numc.data_ptr(), offs.data_ptr(), B, N,
                                        tokens.data_ptr(), ops._dt(tokens), None if gf is None else gf.data_ptr(), ws.data_ptr(),
                                        ws.numel(), ops._stream(dev))
        _lib.check(st, "setok_head_forward")
        rt = RaggedTokens(tokens, offs, index_down=down)
        if return_group_features:
            return rt, idx, score.unsqueeze(1), RaggedTokens(gf, offs)
        return rt, idx, score.unsqueeze(1)

    def forward(self, x, k=None, threshold=None, token_mask=None, noise=None, interpolate_pos_encoding: bool = False,
                token_dtype: Optional[torch.dtype] = None):
        """Inference (`eval()` or grad disabled): the fused path below.  `train()` with gradients enabled and a trainable
        head: `forward_train` (autograd through group_encoding / inter_encoder / out, as in the reference where only the
        tower and the clustering are under no_grad: clip_encoder.py:50, tokenizer.py:79)."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.out.parameters()) and torch.is_tensor(x):
            return self.forward_train(x, k=k, threshold=threshold, token_mask=token_mask, noise=noise, interpolate_pos_encoding=interpolate_pos_encoding)
        return self._forward_inference(x, k=k, threshold=threshold, token_mask=token_mask, noise=noise,
                                       interpolate_pos_encoding=interpolate_pos_encoding, token_dtype=token_dtype)

    def forward_train(self, x, k=None, threshold=None, token_mask=None, noise=None, interpolate_pos_encoding: bool = False):
        """Images (B, 3, H, W) -> the reference's 3-tuple with float32 tokens that carry a graph back to the head's parameters
        (setok_b200/training.py).  Tower and clustering run without gradients, exactly where the reference disables them."""
        from . import training
        tower = self.image_feature_encoder
        with torch.no_grad():
            if tower.select_feature != "patch":
                raise SetokError("forward_train needs mm_vision_select_feature='patch'")
            x_pos = tower(x, interpolate_pos_encoding, pos_embedding=self.position_embedding)
            B, N, _ = x_pos.shape
            h = w = int(math.sqrt(N))
            _threshold = threshold if threshold else self.threshold
            _k = k if k else self.min_cluster_num
            dev = self.device
            noise = torch.rand(B, N, device=dev, dtype=torch.float32) if noise is None else noise.to(device=dev, dtype=torch.float32).reshape(B, N).contiguous()
            if token_mask is not None:
                token_mask = token_mask.to(dev).reshape(B, N)
            _, idx, score, down, numc, offs = ops.dpc_cluster(x_pos, noise, (h, w), int(_k), float(_threshold), int(self.min_cluster_num),
                                                              token_mask=token_mask, embedded=True)
        rt = training.head_forward_train(self, x_pos, idx, numc, offs)
        rt.index_down = down
        return rt, idx, score.unsqueeze(1)

    @torch.no_grad()
    def _forward_inference(self, x, k=None, threshold=None, token_mask=None, noise=None, interpolate_pos_encoding: bool = False,
                           token_dtype: Optional[torch.dtype] = None):
        """x: images (B, 3, H, W) (or a list of (3, H, W)).  Returns the reference's 3-tuple
        ``(group_features, idx_cluster, score)`` (tokenizer.py:182) for the whole batch: ``group_features`` is a
        RaggedTokens whose ``[b]`` is image b's (K_b, C_tok) tensor, ``idx_cluster`` is (B, N) int64 and
        ``score[b]`` has the reference's (1, N) shape.

        A list whose images differ in resolution (BASELINE config 5; needs ``interpolate_pos_encoding=True``) is
        processed as one batch per resolution and re-packed in the original image order; ``idx_cluster`` / ``score``
        are then per-image lists because N differs, and ``noise`` is a list of (N_i,) tensors."""
        if isinstance(x, (list, tuple)) and len({tuple(im.shape) for im in x}) > 1:
            return self._forward_mixed(list(x), k, threshold, noise, interpolate_pos_encoding, token_mask, token_dtype)
        if isinstance(x, (list, tuple)):
            x = torch.stack(list(x), dim=0)
            if isinstance(noise, (list, tuple)):
                noise = torch.stack(list(noise), dim=0)
        tower = self.image_feature_encoder
        if tower.select_feature == "patch" and torch.is_tensor(x):
            # a1..a3 in one call: the tower's last row pass drops CLS and adds the position embedding (tokenizer.py:164-169)
            x_pos = tower(x, interpolate_pos_encoding, pos_embedding=self.position_embedding)
            token_dtype = token_dtype or (torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32)
            return self.encode_features(x_pos, k=k, threshold=threshold, token_mask=token_mask, noise=noise, token_dtype=token_dtype,
                                        embedded=True)
        feats = tower(x, interpolate_pos_encoding)
        return self.encode_features(feats, k=k, threshold=threshold, token_mask=token_mask, noise=noise, token_dtype=token_dtype)

    def _forward_mixed(self, images, k, threshold, noise, interpolate_pos_encoding, token_mask=None, token_dtype=None):
        if token_mask is not None and (not isinstance(token_mask, (list, tuple)) or len(token_mask) != len(images)):
            raise SetokError("a mixed-resolution batch takes token_mask as a list with one (N_i,) mask per image")
        groups: Dict[tuple, List[int]] = {}
        for i, im in enumerate(images):
            groups.setdefault(tuple(im.shape), []).append(i)
        B = len(images)
        per_tokens: List[Optional[torch.Tensor]] = [None] * B
        idxs: List[Optional[torch.Tensor]] = [None] * B
        scores: List[Optional[torch.Tensor]] = [None] * B
        for shape, members in groups.items():
            batch = torch.stack([images[i] for i in members], dim=0)
            nz = None if noise is None else torch.stack([noise[i] for i in members], dim=0)
            tm = None if token_mask is None else torch.stack([token_mask[i].reshape(-1) for i in members], dim=0)
            rt, idx, score = self.forward(batch, k=k, threshold=threshold, token_mask=tm, noise=nz,
                                          interpolate_pos_encoding=interpolate_pos_encoding, token_dtype=token_dtype)
            for j, i in enumerate(members):
                per_tokens[i], idxs[i], scores[i] = rt[j], idx[j], score[j]
        counts = torch.tensor([0] + [t.shape[0] for t in per_tokens], dtype=torch.int32)
        offsets = torch.cumsum(counts, 0).to(dtype=torch.int32, device=self.device)
        return RaggedTokens(torch.cat(per_tokens, dim=0), offsets), idxs, scores

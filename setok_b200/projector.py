"""mm_in_projector on the packed ragged layout (reference src/model/multimodal_projector/builder.py:33-64,
applied at src/model/setokim_arch.py:210).

`build_vision_projector` keeps the reference's signature and returns an nn.Module whose parameters carry
the reference's state_dict keys ('weight'/'bias' for 'linear'; Sequential indices '0', '2', ... for
'mlpNx_gelu'; '0','1','3',.. with '_Norm').  Its forward takes a RaggedTokens (or a (rows, C) / (B, K, C)
CUDA tensor) and runs the Linear/GELU chain as packed-row tcgen05 GEMMs with the GELU fused in the epilogue."""
from __future__ import annotations

import ctypes as C
import re
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import SetokError
from ._pack import PackedParams
from .ragged import RaggedTokens


class IdentityMap(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_projector_type": "identity"}


class _PackedProjector(PackedParams):
    """Mixin: packs the Linear chain and calls setok_project."""

    def _linears(self):
        raise NotImplementedError

    def _norm(self) -> Optional[nn.LayerNorm]:
        return None

    def _pack(self):
        lins = self._linears()
        dev = lins[0].weight.device
        if dev.type != "cuda":
            raise SetokError("the projector must live on a CUDA device (setok_b200 has no CPU path)")
        n = len(lins)
        ws = [l.weight.detach().to(torch.bfloat16).contiguous() for l in lins]
        bs = [l.bias.detach().to(torch.float32).contiguous() for l in lins]
        wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
        dims = (C.c_int * (n + 1))(*([lins[0].in_features] + [l.out_features for l in lins]))
        norm = self._norm()
        ng = norm.weight.detach().float().contiguous() if norm is not None else None
        nb = norm.bias.detach().float().contiguous() if norm is not None else None
        proj = _lib.Projector(n_linear=n, w=wp, b=bp, dims=dims, norm_g=None if ng is None else ng.data_ptr(),
                              norm_b=None if nb is None else nb.data_ptr())
        self._packed = (proj, (ws, bs, wp, bp, dims, ng, nb))
        return self._packed

    @torch.no_grad()
    def forward(self, x, out_dtype=None):
        proj, keep = self._packed_get()
        ragged = isinstance(x, RaggedTokens)
        data = x.data if ragged else x
        lead = None
        if not ragged and data.dim() == 3:
            lead = data.shape[:2]
            data = data.reshape(-1, data.shape[-1])
        dev = ops._dev(data)
        if data.dtype not in (torch.float32, torch.bfloat16):
            data = data.float()
        rows = data.shape[0]
        out_dtype = out_dtype or data.dtype
        out = torch.empty(rows, keep[4][proj.n_linear], dtype=out_dtype, device=dev)
        lib = _lib.load()
        nbytes = lib.setok_project_workspace_bytes(C.byref(proj), rows)
        ws = ops.workspace(dev, nbytes, "proj")
        m_dev = x.offsets[-1:] if ragged else None
        with torch.cuda.device(dev):
            st = lib.setok_project(C.byref(proj), data.data_ptr(), ops._dt(data), rows, None if m_dev is None else m_dev.data_ptr(),
                                   out.data_ptr(), ops._dt(out), ws.data_ptr(), ws.numel(), ops._stream(dev))
        _lib.check(st, "setok_project")
        if ragged:
            return x.with_data(out)
        return out.reshape(*lead, -1) if lead is not None else out


class LinearProjector(_PackedProjector, nn.Linear):
    def __init__(self, in_features, out_features):
        nn.Linear.__init__(self, in_features, out_features)

    def _linears(self):
        return [self]


class MlpProjector(_PackedProjector, nn.Sequential):
    def __init__(self, *modules):
        nn.Sequential.__init__(self, *modules)

    def _linears(self):
        return [m for m in self if isinstance(m, nn.Linear)]

    def _norm(self):
        for m in self:
            if isinstance(m, nn.LayerNorm):
                return m
        return None


def build_vision_projector(projector_type="linear", mm_hidden_size=4096, hidden_size=3078, delay_load=False, **kwargs):
    if projector_type == "linear":
        return LinearProjector(mm_hidden_size, hidden_size)
    use_norm = False
    if "_Norm" in projector_type:
        use_norm = True
        projector_type = projector_type.replace("_Norm", "")
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m:
        depth = int(m.group(1))
        modules = [nn.Linear(mm_hidden_size, hidden_size)]
        if use_norm:
            modules.append(nn.LayerNorm(hidden_size))
        for _ in range(1, depth):
            modules.append(nn.GELU())
            modules.append(nn.Linear(hidden_size, hidden_size))
        return MlpProjector(*modules)
    if projector_type == "identity":
        return IdentityMap()
    raise ValueError(f"Unknown projector type: {projector_type}")

"""Validity stamp for the packed (bf16 matrix / f32 vector) copies of a module's parameters.

The kernels read packed copies made once; the copies must be rebuilt whenever the parameters they came from change.
nn.Module offers no single hook that sees every way a parameter can change (a parent's ``load_state_dict`` recurses
through ``_load_from_state_dict`` and never calls a child's ``load_state_dict`` override; optimizers and ``copy_`` write
in place), so the cache is validated instead: every tensor carries a version counter that any in-place write bumps, and
a storage address that ``.to()`` / re-assignment changes.  ``stamp()`` folds both over the source parameters (~0.1 ms
for a ViT-L tower) and the packed copy is reused only while the stamp is unchanged."""
from __future__ import annotations

from typing import Iterable

import torch


def stamp(modules: Iterable[torch.nn.Module]) -> int:
    h = 0x345678
    for m in modules:
        for t in m.parameters():
            h = (h * 1000003) ^ t.data_ptr() ^ (t._version << 48)
            h &= (1 << 63) - 1
        for t in m.buffers():
            h = (h * 1000003) ^ t.data_ptr() ^ (t._version << 48)
            h &= (1 << 63) - 1
    return h


class PackedParams:
    """Mixin: ``self._packed_get()`` returns the packed copy, re-packing (``self._pack()``) when the source parameters
    moved or were written since the last pack.  ``_pack_sources()`` names the modules whose parameters are packed."""

    _packed = None
    _packed_stamp = None

    def _pack_sources(self):
        return (self,)

    def _packed_get(self):
        s = stamp(self._pack_sources())
        if self._packed is None or s != self._packed_stamp:
            self._packed = None
            self._pack()
            self._packed_stamp = s
        return self._packed

    def invalidate(self):
        self._packed = None


def fold_layernorm_into_linear(w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """What SETOK_VIT_LN_FOLD packs for a Linear that follows a LayerNorm (include/setok_b200.h): ``LN(x) W^T + b`` is evaluated by the
    GEMM as ``(rho / r) * (xhat W'^T)_n - rho * m * s_n + t_n`` with xhat = (x - c) r the stream normalised by lagged statistics
    (c, r), m = mean(x - c), rho = rsqrt(var + eps).  Returns (W' = bf16(W diag(gamma)), s = row sums of the ROUNDED W' -- the mean
    correction must cancel exactly against what the tensor cores multiply --, t = W beta + b), all computed in float32."""
    w = w.float()
    wg = (w * gamma.float()[None, :]).to(torch.bfloat16).contiguous()
    s = wg.float().sum(1).contiguous()
    t = ((w * beta.float()[None, :]).sum(1) + b.float()).contiguous()
    return wg, s, t

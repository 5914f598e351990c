"""Training path of the tokenizer head and the projector.

In the reference, gradients reach `inner_encoder`, `inter_encoder`, `out` (src/model/setok/tokenizer.py:123-155, 179-180;
`Block` / `Attention` / `Mlp` module.py:29-100) and `mm_in_projector` (multimodal_projector/builder.py:33-64) through
torch.autograd; the tower runs under `@torch.no_grad()` (clip_encoder.py:50) and so does the clustering (tokenizer.py:79), so
nothing upstream of the position-embedded features needs a gradient.  Here the same computation is expressed as
`torch.autograd.Function`s over libsetok_b200:

* every contraction, forward and backward, is the tcgen05 GEMM: y = x W^T, dgrad dx = dy W (W as the MN-major operand),
  wgrad dW = dy^T x (x as the MN-major operand, dy^T from the bf16 transpose kernel); the attention backward is the same four
  products per head around the masked-softmax backward kernel;
* LayerNorm / GELU / segment mean / padding have their own forward + backward row kernels (csrc/train.cu).

bf16 operands, f32 accumulation, f32 residual stream and f32 gradients.  The inference path (`setok_head_forward`) is untouched:
this path trades its fusion for saved activations.  `all_gather_with_grad` is the differentiable all-gather the contrastive loss
uses (src/model/loss/multilabel_constrastive.py:14-23, diffdist.functional.all_gather): forward all-gather, backward reduce-scatter."""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import SetokError
from .ragged import RaggedTokens

BF16, F32 = torch.bfloat16, torch.float32


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _call(name, *args):
    _lib.check(getattr(_lib.load(), name)(*args), name)


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x if x.dtype == BF16 else x.to(BF16)


# --------------------------------------------------------------------------------------------------------------------
# thin wrappers over the row kernels
# --------------------------------------------------------------------------------------------------------------------
def transpose_to_bf16(x: torch.Tensor) -> torch.Tensor:
    """x (rows, cols) or (batch, rows, cols), f32|bf16, unit inner stride -> bf16 (.., cols, rows); the result's leading
    dimension is padded to a multiple of 8 (the view returned has the exact shape)."""
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(0)
    Bt, rows, cols = x.shape
    if x.stride(2) != 1:
        raise SetokError("transpose_to_bf16 needs a unit inner stride")
    ld = (rows + 7) // 8 * 8
    out = torch.empty(Bt, cols, ld, dtype=BF16, device=x.device)
    with torch.cuda.device(x.device):
        _call("setok_transpose_to_bf16", x.data_ptr(), ops._dt(x), x.stride(1), x.stride(0), out.data_ptr(), ld, cols * ld, rows, cols, Bt, _st(x.device))
    out = out[:, :, :rows]
    return out[0] if squeeze else out


def colsum(x: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(x.shape[1], dtype=F32, device=x.device)
    with torch.cuda.device(x.device):
        _call("setok_colsum_add", x.data_ptr(), ops._dt(x), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), _st(x.device))
    return out


def _gemm_mn(a: torch.Tensor, w_kn: torch.Tensor, out: Optional[torch.Tensor] = None, out_dtype=F32) -> torch.Tensor:
    """a (M, K) bf16 @ w_kn (K, N) bf16 given row-major (the MN-major operand form); 2-D convenience over gemm_batched."""
    return ops.gemm_batched(a.unsqueeze(0), w_kn.unsqueeze(0), w_mn_major=True, out_dtype=out_dtype, out=None if out is None else out.unsqueeze(0))[0]


# --------------------------------------------------------------------------------------------------------------------
# autograd functions
# --------------------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b (+ residual): nn.Linear on the tcgen05 GEMM, f32 out."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        xb, wb = _bf16(x).contiguous(), _bf16(weight).contiguous()
        y = ops.gemm(xb, wb, bias.detach().float().contiguous(), residual=residual, out_dtype=F32)
        ctx.save_for_backward(xb, wb)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        dy = dy.contiguous()
        dyb = _bf16(dy)
        dx = _gemm_mn(dyb, wb) if ctx.needs_input_grad[0] else None             # (M, N) @ W (N, K)
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = _gemm_mn(transpose_to_bf16(dy), xb)                             # dy^T (N, M) @ x (M, K)
        if ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dx, dw, db, (dy if ctx.has_res else None)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        y = ops.layernorm(x, g, b, eps, out_dtype=BF16)
        ctx.save_for_backward(x, g)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        dy = dy.contiguous()
        rows, Cc = x.shape
        dx = torch.empty_like(x)
        dg, db = torch.zeros(Cc, dtype=F32, device=x.device), torch.zeros(Cc, dtype=F32, device=x.device)
        with torch.cuda.device(x.device):
            _call("setok_layernorm_bwd", x.data_ptr(), dy.data_ptr(), ops._dt(dy), g.data_ptr(), ctx.eps, rows, Cc, dx.data_ptr(), dg.data_ptr(),
                  db.data_ptr(), _st(x.device))
        return dx, dg, db, None


class GeluFn(torch.autograd.Function):
    """act = GELU_erf(pre) as bf16 (nn.GELU, module.py:38)."""

    @staticmethod
    def forward(ctx, pre):
        pre = pre.contiguous()
        act = torch.empty(pre.shape, dtype=BF16, device=pre.device)
        with torch.cuda.device(pre.device):
            _call("setok_gelu_fwd", pre.data_ptr(), ops._dt(pre), act.data_ptr(), pre.numel(), _st(pre.device))
        ctx.save_for_backward(pre)
        return act

    @staticmethod
    def backward(ctx, dy):
        (pre,) = ctx.saved_tensors
        dy = dy.contiguous()
        dpre = torch.empty(pre.shape, dtype=F32, device=pre.device)
        with torch.cuda.device(pre.device):
            _call("setok_gelu_bwd", pre.data_ptr(), ops._dt(pre), dy.data_ptr(), ops._dt(dy), dpre.data_ptr(), pre.numel(), _st(pre.device))
        return dpre.to(pre.dtype) if pre.dtype != F32 else dpre


class SegmentedAttentionFn(torch.autograd.Function):
    """Attention.forward (module.py:61-73) over rows that are N-token groups sorted by segment: softmax(q k^T scale) v within
    each row's segment.  qkv (R, 3C) laid out [q | k | v] (rounded to bf16 here); returns bf16 (R, C); the gradient comes back
    in f32.  Dense per group on the tensor cores."""

    @staticmethod
    def forward(ctx, qkv_in, seg_off, row_seg, N, heads, scale):
        qkv = _bf16(qkv_in).contiguous()
        R, C3 = qkv.shape
        Cc = C3 // 3
        hd = Cc // heads
        G = R // N
        dev = qkv.device
        q3 = qkv.view(G, N, C3)
        ldP = (N + 7) // 8 * 8
        ao = torch.empty(R, Cc, dtype=BF16, device=dev)
        Ps: List[torch.Tensor] = []
        for h in range(heads):
            qh, kh, vh = q3[:, :, h * hd:(h + 1) * hd], q3[:, :, Cc + h * hd:Cc + (h + 1) * hd], q3[:, :, 2 * Cc + h * hd:2 * Cc + (h + 1) * hd]
            S = ops.gemm_batched(qh, kh, out_dtype=F32)                                  # (G, N, N)
            P = torch.empty(R, ldP, dtype=BF16, device=dev)
            with torch.cuda.device(dev):
                _call("setok_masked_softmax", S.data_ptr(), N, seg_off.data_ptr(), row_seg.data_ptr(), R, N, float(scale), P.data_ptr(), ldP, _st(dev))
            ops.gemm_batched(P.view(G, N, ldP)[:, :, :N], vh, w_mn_major=True, out=ao.view(G, N, Cc)[:, :, h * hd:(h + 1) * hd])
            Ps.append(P)
        ctx.save_for_backward(qkv, seg_off, row_seg, *Ps)
        ctx.cfg = (N, heads, float(scale))
        return ao

    @staticmethod
    def backward(ctx, dao):
        qkv, seg_off, row_seg, *Ps = ctx.saved_tensors
        N, heads, scale = ctx.cfg
        R, C3 = qkv.shape
        Cc = C3 // 3
        hd = Cc // heads
        G = R // N
        dev = qkv.device
        ldP = (N + 7) // 8 * 8
        q3 = qkv.view(G, N, C3)
        dao3 = _bf16(dao.contiguous()).view(G, N, Cc)
        dqkv = torch.empty(R, C3, dtype=F32, device=dev)
        d3 = dqkv.view(G, N, C3)
        for h in range(heads):
            qh, kh, vh = q3[:, :, h * hd:(h + 1) * hd], q3[:, :, Cc + h * hd:Cc + (h + 1) * hd], q3[:, :, 2 * Cc + h * hd:2 * Cc + (h + 1) * hd]
            dOh = dao3[:, :, h * hd:(h + 1) * hd]
            P3 = Ps[h].view(G, N, ldP)[:, :, :N]
            ops.gemm_batched(transpose_to_bf16(P3), dOh, w_mn_major=True, out=d3[:, :, 2 * Cc + h * hd:2 * Cc + (h + 1) * hd])      # dV = P^T dO
            dP = ops.gemm_batched(dOh, vh, out_dtype=F32)                                                                             # dP = dO V^T
            dS = torch.empty(R, ldP, dtype=BF16, device=dev)
            with torch.cuda.device(dev):
                _call("setok_masked_softmax_bwd", Ps[h].data_ptr(), ldP, dP.data_ptr(), N, seg_off.data_ptr(), row_seg.data_ptr(), R, N, scale,
                      dS.data_ptr(), ldP, _st(dev))
            dS3 = dS.view(G, N, ldP)[:, :, :N]
            ops.gemm_batched(dS3, kh, w_mn_major=True, out=d3[:, :, h * hd:(h + 1) * hd])                                            # dQ = dS K
            ops.gemm_batched(transpose_to_bf16(dS3), qh, w_mn_major=True, out=d3[:, :, Cc + h * hd:Cc + (h + 1) * hd])              # dK = dS^T Q
        return dqkv, None, None, None, None, None


class SegmentMeanFn(torch.autograd.Function):
    """tokenizer.py:151: mean over each segment of the sorted rows -> (n_segments, C)."""

    @staticmethod
    def forward(ctx, x, seg_off, row_seg, n_segments):
        x = x.contiguous()
        out = torch.empty(n_segments, x.shape[1], dtype=F32, device=x.device)
        with torch.cuda.device(x.device):
            _call("setok_segment_mean", x.data_ptr(), seg_off.data_ptr(), n_segments, x.shape[1], out.data_ptr(), _st(x.device))
        ctx.save_for_backward(seg_off, row_seg)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, dg):
        seg_off, row_seg = ctx.saved_tensors
        dg = dg.contiguous()
        dx = torch.empty(ctx.rows, dg.shape[1], dtype=F32, device=dg.device)
        with torch.cuda.device(dg.device):
            _call("setok_segment_mean_bwd", dg.data_ptr(), seg_off.data_ptr(), row_seg.data_ptr(), ctx.rows, dg.shape[1], dx.data_ptr(), _st(dg.device))
        return dx, None, None, None


def _gather(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    out = torch.empty(index.numel(), x.shape[1], dtype=F32, device=x.device)
    with torch.cuda.device(x.device):
        _call("setok_gather_rows_f32", x.data_ptr(), index.data_ptr(), index.numel(), x.shape[1], out.data_ptr(), _st(x.device))
    return out


class GatherRowsFn(torch.autograd.Function):
    """out[r] = x[index[r]] (0 where index < 0); `inverse` maps x's rows to out's (-1: the row is not gathered)."""

    @staticmethod
    def forward(ctx, x, index, inverse):
        ctx.save_for_backward(inverse)
        return _gather(x.contiguous(), index)

    @staticmethod
    def backward(ctx, dy):
        (inverse,) = ctx.saved_tensors
        return _gather(dy.contiguous(), inverse), None, None


# --------------------------------------------------------------------------------------------------------------------
# the reference's modules, differentiable
# --------------------------------------------------------------------------------------------------------------------
def block_train(blk, x: torch.Tensor, seg_off: torch.Tensor, row_seg: torch.Tensor, N: int, heads: int) -> torch.Tensor:
    """Block.forward (module.py:95-100): depth x [x += Attn_i(norm1(x))], then x += Mlp(norm2(x)); x f32 (R, C), rows in groups
    of N sorted by segment."""
    Cc = x.shape[1]
    scale = (Cc // heads) ** -0.5
    for i in range(blk.depth):
        at = blk.layers[i][1]
        h = LayerNormFn.apply(x, blk.norm1.weight, blk.norm1.bias, 1e-5)
        qkv = LinearFn.apply(h, at.qkv.weight, at.qkv.bias, None)
        ao = SegmentedAttentionFn.apply(qkv, seg_off, row_seg, N, heads, scale)
        x = LinearFn.apply(ao, at.proj.weight, at.proj.bias, x)
    h = LayerNormFn.apply(x, blk.norm2.weight, blk.norm2.bias, 1e-5)
    pre = LinearFn.apply(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias, None)
    act = GeluFn.apply(pre)
    return LinearFn.apply(act, blk.mlp.fc2.weight, blk.mlp.fc2.bias, x)


def head_forward_train(tok, x_pos: torch.Tensor, idx_cluster: torch.Tensor, num_clusters: torch.Tensor, offsets: torch.Tensor) -> RaggedTokens:
    """group_encoding -> inter_encoder -> out (tokenizer.py:123-155, 178-180) with autograd through the head's parameters.
    x_pos (B, N, C) f32 = features + position embedding; clustering outputs from `ops.dpc_cluster` (no gradient, :79)."""
    B, N, Cc = x_pos.shape
    dev = x_pos.device
    if N % 4 != 0 or N < 8:
        raise SetokError(f"the training path needs N % 4 == 0 (N = {N})")
    R = B * N
    perm = torch.empty(R, dtype=torch.int32, device=dev)
    row_seg = torch.empty(R, dtype=torch.int32, device=dev)
    seg_off = torch.empty(R + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _call("setok_sort_by_cluster", idx_cluster.data_ptr(), num_clusters.data_ptr(), offsets.data_ptr(), B, N, perm.data_ptr(), row_seg.data_ptr(),
              seg_off.data_ptr(), _st(dev))
    xs = _gather(x_pos.reshape(R, Cc).float().contiguous(), perm)                      # tokens sorted by cluster (no gradient upstream)
    x = block_train(tok.inner_encoder, xs, seg_off, row_seg, N, tok.nheads)
    offs = [int(v) for v in offsets.tolist()]                                          # one host read per training step
    total, counts = offs[-1], np.diff(np.asarray(offs))
    g = SegmentMeanFn.apply(x, seg_off, row_seg, total)                                # (sum K, C): tokenizer.py:151-153
    # inter_encoder (tokenizer.py:179, repair R2): each image's K_b cluster tokens attend to each other.  Padded to Kpad rows per
    # image so that the dense attention applies; two segments per image (live rows, padding rows).
    Kpad = max(8, int(math.ceil(counts.max() / 8.0)) * 8)
    pad_index = np.full(B * Kpad, -1, dtype=np.int32)
    unpad_index = np.empty(total, dtype=np.int32)
    seg2 = np.empty(2 * B + 1, dtype=np.int32)
    rseg2 = np.empty(B * Kpad, dtype=np.int32)
    for b in range(B):
        k = int(counts[b])
        pad_index[b * Kpad:b * Kpad + k] = np.arange(offs[b], offs[b] + k)
        unpad_index[offs[b]:offs[b] + k] = b * Kpad + np.arange(k)
        seg2[2 * b], seg2[2 * b + 1] = b * Kpad, b * Kpad + k
        rseg2[b * Kpad:b * Kpad + k] = 2 * b
        rseg2[b * Kpad + k:(b + 1) * Kpad] = 2 * b + 1
    seg2[2 * B] = B * Kpad
    to_dev = lambda a: torch.from_numpy(a).to(dev)
    pad_index, unpad_index, seg2, rseg2 = to_dev(pad_index), to_dev(unpad_index), to_dev(seg2), to_dev(rseg2)
    gp = GatherRowsFn.apply(g, pad_index, unpad_index)
    y = block_train(tok.inter_encoder, gp, seg2, rseg2, Kpad, tok.nheads)
    t = GatherRowsFn.apply(y, unpad_index, pad_index)
    tokens = LinearFn.apply(t, tok.out.weight, tok.out.bias, None)                     # tokenizer.py:180
    rt = RaggedTokens(tokens, offsets)
    rt._host = offs
    return rt


def projector_forward_train(proj, rows: torch.Tensor) -> torch.Tensor:
    """mm_in_projector (multimodal_projector/builder.py:33-64) over packed rows with autograd: Linear [-> LayerNorm] -> GELU -> Linear ..."""
    lins = proj._linears()
    norm = proj._norm()
    x = rows
    for i, lin in enumerate(lins):
        x = LinearFn.apply(x, lin.weight, lin.bias, None)
        if i == 0 and norm is not None:
            x = LayerNormFn.apply(x, norm.weight, norm.bias, 1e-5).float()
        if i < len(lins) - 1:
            x = GeluFn.apply(x)
    return x


# --------------------------------------------------------------------------------------------------------------------
# differentiable all-gather (contrastive loss)
# --------------------------------------------------------------------------------------------------------------------
class _AllGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        world = dist.get_world_size(ctx.group)
        rank = dist.get_rank(ctx.group)
        grad = grad.contiguous()
        n = grad.shape[0] // world
        mine = torch.empty((n,) + tuple(grad.shape[1:]), dtype=grad.dtype, device=grad.device)
        if grad.is_cuda:
            dist.reduce_scatter_tensor(mine, grad, op=dist.ReduceOp.SUM, group=ctx.group)
        else:                                   # gloo has no reduce_scatter: all-reduce, then keep this rank's slice
            g = grad.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
            mine = g[rank * n:(rank + 1) * n].clone()
        return mine, None


def all_gather_with_grad(x: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """`dist_collect` of multilabel_constrastive.py:14-23 (diffdist.functional.all_gather + cat): every rank gets the
    concatenation of all ranks' (B, C) embeddings, and the gradient of each rank's slice is the SUM over ranks of the gradients
    that flowed into it (a reduce-scatter), so the contrastive loss trains as if the global batch sat on one device."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllGatherFn.apply(x, group)

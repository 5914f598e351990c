"""Tensor-level wrappers over the C ABI.  torch is used for device memory and streams only; every
function here requires CUDA tensors and raises otherwise (no CPU path)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_QUICK_GELU, BF16, F32, SetokError, check  # noqa: F401


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise SetokError(f"unsupported dtype {t.dtype} (float32 or bfloat16 expected)")


def _dev(*ts: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise SetokError("setok_b200 kernels need CUDA tensors; there is no CPU fallback")
        if not t.is_contiguous():
            raise SetokError("setok_b200 kernels need contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise SetokError("tensors live on different devices")
    return dev


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


_ws_cache = {}


def workspace(dev: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    """Grow-only scratch buffer per (device, CUDA stream, tag), so steady-state steps allocate nothing and two streams
    (or threads, each on its own stream) running the same op never share scratch."""
    stream = torch.cuda.current_stream(dev)
    key = (dev.index, stream.cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        buf.record_stream(stream)
        _ws_cache[key] = buf
    return buf


def ln_records(rows: int, C: int, dev) -> torch.Tensor:
    """Row records of the folded LayerNorms (setok_gemm_bf16_ln): (rows, 2 + 2 * ceil(C / 128)) float32."""
    return torch.empty(rows, 2 + 2 * ((C + 127) // 128), dtype=torch.float32, device=dev)


def ln_fold_init(x: torch.Tensor, eps: float = 1e-5):
    """First record and xhat of a float32 residual stream x (rows, C): returns (xhat bf16, records)."""
    dev = _dev(x)
    rows, C_ = x.shape
    xhat = torch.empty(rows, C_, dtype=torch.bfloat16, device=dev)
    rec = ln_records(rows, C_, dev)
    with torch.cuda.device(dev):
        st = _lib.load().setok_ln_fold_init(x.data_ptr(), xhat.data_ptr(), rec.data_ptr(), float(eps), rows, C_, _stream(dev))
    check(st, "setok_ln_fold_init")
    return xhat, rec


def gemm_ln(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, rec_in: torch.Tensor, *, ln_C: int, eps: float = 1e-5, act: int = ACT_NONE,
            ln_s: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, rec_out: Optional[torch.Tensor] = None,
            xhat: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The LayerNorm-fold GEMMs (include/setok_b200.h: setok_gemm_bf16_ln).  Consuming side: ``ln_s`` given, a = xhat, returns
    act(LN(x) W0^T + b0) in bf16.  Producing side: ``residual`` (the f32 stream, updated in place when ``out`` is it), ``rec_out`` and
    ``xhat`` given."""
    dev = _dev(a, w, bias, rec_in, ln_s, residual, rec_out, xhat, out)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if residual is not None else torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().setok_gemm_bf16_ln(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), _dt(out), _p(bias),
                                            _p(residual), residual.stride(0) if residual is not None else 0, _dt(residual) if residual is not None else 0,
                                            act, M, N, K, rec_in.data_ptr(), _p(rec_out), _p(ln_s), _p(xhat), xhat.stride(0) if xhat is not None else 0,
                                            float(eps), int(ln_C), _stream(dev))
    check(st, "setok_gemm_bf16_ln")
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16,
         m_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) + residual   (tcgen05 GEMM)."""
    dev = _dev(a, w, bias, residual, out, m_dev)
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise SetokError("gemm operands must be bfloat16")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise SetokError(f"gemm shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().setok_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), _dt(out),
                                         _p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
                                         _dt(residual) if residual is not None else 0, act, M, N, K, _p(m_dev), _stream(dev))
    check(st, "setok_gemm_bf16")
    return out


def gemm_batched(a: torch.Tensor, w: torch.Tensor, *, w_mn_major: bool = False, out_dtype=torch.bfloat16, bias: Optional[torch.Tensor] = None,
                 act: int = ACT_NONE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a (B, M, K) bf16 (row stride may exceed K); w (B, N, K) or, with w_mn_major, (B, K, N).  Returns (B, M, N); `out` may
    be a strided view (unit inner stride), e.g. one head's columns of a packed buffer."""
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or a.dim() != 3 or w.dim() != 3:
        raise SetokError("gemm_batched needs 3-D bfloat16 operands")
    if not a.is_cuda or a.stride(2) != 1 or w.stride(2) != 1:
        raise SetokError("gemm_batched needs CUDA operands with a unit innermost stride")
    dev = a.device
    Bt, M, K = a.shape
    N = w.shape[2] if w_mn_major else w.shape[1]
    if out is None:
        out = torch.empty(Bt, M, N, dtype=out_dtype, device=dev)
    elif tuple(out.shape) != (Bt, M, N) or out.stride(2) != 1 or not out.is_cuda:
        raise SetokError(f"gemm_batched: out must be a CUDA ({Bt}, {M}, {N}) tensor with a unit inner stride")
    with torch.cuda.device(dev):
        st = _lib.load().setok_gemm_bf16_batched(a.data_ptr(), a.stride(1), a.stride(0), w.data_ptr(), w.stride(1), w.stride(0), int(w_mn_major),
                                                 out.data_ptr(), out.stride(1), out.stride(0), _dt(out), _p(bias), act, Bt, M, N, K, _stream(dev))
    check(st, "setok_gemm_bf16_batched")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, *, out_dtype=torch.bfloat16,
              gather: Optional[torch.Tensor] = None, m_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _dev(x, gamma, beta, gather, m_dev, out)
    rows = x.shape[0] if gather is None else gather.shape[0]
    Cc = x.shape[1]
    if out is None:
        out = torch.empty(rows, Cc, dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().setok_layernorm(x.data_ptr(), _dt(x), out.data_ptr(), _dt(out), gamma.data_ptr(), beta.data_ptr(), eps, rows, Cc,
                                         _p(gather), _p(m_dev), _stream(dev))
    check(st, "setok_layernorm")
    return out


def attention(qkv: torch.Tensor, heads: int, scale: float, *, seg_off: Optional[torch.Tensor] = None,
              row_seg: Optional[torch.Tensor] = None, uniform_T: int = 0, m_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _dev(qkv, seg_off, row_seg, m_dev)
    if qkv.dtype != torch.bfloat16:
        raise SetokError("attention needs bfloat16 qkv")
    rows, C3 = qkv.shape
    Cc = C3 // 3
    out = torch.empty(rows, Cc, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().setok_attention(qkv.data_ptr(), out.data_ptr(), rows, Cc, heads, scale, _p(seg_off), _p(row_seg), uniform_T,
                                         _p(m_dev), _stream(dev))
    check(st, "setok_attention")
    return out


def dpc_cluster(feats: torch.Tensor, noise: torch.Tensor, hw: Tuple[int, int], k: int, threshold: float, min_cluster_num: int,
                pos_table: Optional[torch.Tensor] = None, token_mask: Optional[torch.Tensor] = None, embedded: bool = False):
    """feats (B, N, C) f32|bf16, noise (B, N) f32.  Returns x_pos (B,N,C) f32, idx_cluster (B,N) i64, score (B,N) f32,
    index_down (B,N) i64 (-1 padded), num_clusters (B,) i32, offsets (B+1,) i32 — all on the device, no sync.
    embedded=True: `feats` already carries the position embedding (tokenizer.py:168); it is read once and returned
    as x_pos unchanged (no table, no copy)."""
    dev = _dev(feats, noise, pos_table, token_mask)
    B, N, Cc = feats.shape
    h, w = hw
    if h * w != N:
        raise SetokError(f"h*w ({h}x{w}) != N ({N})")
    if noise.dtype != torch.float32 or tuple(noise.shape) != (B, N):
        raise SetokError("noise must be float32 (B, N)")
    if pos_table is not None and (pos_table.dtype != torch.float32 or pos_table.numel() != N * Cc):
        raise SetokError("pos_table must be float32 with N*C elements")
    if token_mask is not None:
        token_mask = token_mask.to(torch.float32).contiguous()
    x_pos = feats if embedded else torch.empty(B, N, Cc, dtype=torch.float32, device=dev)
    idx = torch.empty(B, N, dtype=torch.int64, device=dev)
    score = torch.empty(B, N, dtype=torch.float32, device=dev)
    down = torch.empty(B, N, dtype=torch.int64, device=dev)
    numc = torch.empty(B, dtype=torch.int32, device=dev)
    offs = torch.empty(B + 1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.setok_dpc_workspace_bytes(B, N, Cc)
    ws = workspace(dev, nbytes, "dpc")
    with torch.cuda.device(dev):
        if embedded:
            st = lib.setok_dpc_cluster_embedded(feats.data_ptr(), _dt(feats), noise.data_ptr(), _p(token_mask), B, N, Cc, k, float(threshold),
                                                min_cluster_num, idx.data_ptr(), score.data_ptr(), down.data_ptr(), numc.data_ptr(),
                                                offs.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev))
        else:
            st = lib.setok_dpc_cluster_pos(feats.data_ptr(), _dt(feats), _p(pos_table), noise.data_ptr(), _p(token_mask), B, h, w, Cc, k,
                                           float(threshold), min_cluster_num, x_pos.data_ptr(), idx.data_ptr(), score.data_ptr(),
                                           down.data_ptr(), numc.data_ptr(), offs.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev))
    check(st, "setok_dpc_cluster")
    return x_pos, idx, score, down, numc, offs

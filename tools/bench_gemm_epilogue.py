#!/usr/bin/env python
"""A/B of the tcgen05 GEMM's two epilogues at the ViT-L layer shapes (M = 256 x 257 rows): the transposing one (TMEM ->
registers -> shared-memory transpose -> coalesced stores) against the row-owner one (TMEM -> registers -> 256-bit stores,
no shared memory).  The variants alternate inside every round and the median over the rounds is reported, so that the
box's power / thermal drift hits both alike.  Also checks that both epilogues produce identical bits.  GPU only."""
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import _lib, ops

dev = torch.device("cuda:0")
M, C, F = 256 * 257, 1024, 4096
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    lib = _lib.load()
    lib.setok_debug_set_gemm_epi_direct.argtypes = [ctypes.c_int]
    lib.setok_debug_set_gemm_epi_direct.restype = None
    a = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    u = torch.randn(M, F, device=dev, generator=g).to(torch.bfloat16)
    xb = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    xf = torch.randn(M, C, device=dev, generator=g)
    shapes = [("qkv", a, 3 * C, C, ops.ACT_NONE, None, torch.bfloat16), ("fc1", a, F, C, ops.ACT_QUICK_GELU, None, torch.bfloat16),
              ("fc1/gelu_erf", a, F, C, ops.ACT_GELU_ERF, None, torch.bfloat16),
              ("out_proj/bf16", a, C, C, ops.ACT_NONE, xb, torch.bfloat16), ("out_proj/f32", a, C, C, ops.ACT_NONE, xf, torch.float32),
              ("fc2/bf16", u, C, F, ops.ACT_NONE, xb, torch.bfloat16), ("fc2/f32", u, C, F, ops.ACT_NONE, xf, torch.float32),
              ("plain/f32out", a, C, C, ops.ACT_NONE, None, torch.float32)]
    print(f"{'shape':14s} {'transpose_us':>12s} {'rowowner_us':>12s} {'ratio':>6s}  TFLOP/s(t)  TFLOP/s(r)  identical")
    tot = [0.0, 0.0]
    for name, inp, n, k, act, res, odt in shapes:
        w = (torch.randn(n, k, device=dev, generator=g) * k ** -0.5).to(torch.bfloat16)
        b = torch.randn(n, device=dev, generator=g)
        outs = []
        for mode in (0, 1):
            lib.setok_debug_set_gemm_epi_direct(mode)
            out = torch.empty(M, n, dtype=odt, device=dev)
            ops.gemm(inp, w, b, act=act, residual=res, out=out)
            outs.append(out)
        same = bool(torch.equal(outs[0], outs[1]))
        ts = ([], [])
        out = torch.empty(M, n, dtype=odt, device=dev)
        for _ in range(5):
            for mode in (0, 1):
                lib.setok_debug_set_gemm_epi_direct(mode)
                ts[mode].append(timeit(lambda: ops.gemm(inp, w, b, act=act, residual=res, out=out)))
        t0, t1 = statistics.median(ts[0]), statistics.median(ts[1])
        fl = 2.0 * M * n * k
        if name in ("qkv", "fc1", "out_proj/f32", "fc2/f32"):
            tot[0] += t0
            tot[1] += t1
        print(f"{name:14s} {t0 * 1e3:12.1f} {t1 * 1e3:12.1f} {t1 / t0:6.3f}  {fl / t0 / 1e9:9.1f}  {fl / t1 / 1e9:9.1f}  {same}")
    lib.setok_debug_set_gemm_epi_direct(-1)
    print(f"ViT layer (qkv + out_proj/f32 + fc1 + fc2/f32): transpose {tot[0] * 1e3:.1f} us, row-owner {tot[1] * 1e3:.1f} us")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Decomposes the tcgen05 GEMM's time at the ViT-L layer shapes by switching parts of the epilogue off
(setok_debug_set_gemm_epi_mode): 0 normal, 1 drain only (MMA + TMA + TMEM loads), 2 + smem transpose + math (no global traffic),
3 normal minus the residual loads, 4 drain only + the producer stages A only, 5 drain only + no operand loads at all
(stale shared memory: the tensor pipe's own rate at the clocks the power cap allows).  bf16 and f32 residual streams.  GPU only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import _lib, ops

dev = torch.device("cuda:0")
M, C, F = 256 * 257, 1024, 4096
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, reps=20, warm=3):
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
    lib.setok_debug_set_gemm_epi_mode.argtypes = [ctypes.c_int]
    a = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    u = torch.randn(M, F, device=dev, generator=g).to(torch.bfloat16)
    xb = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    xf = torch.randn(M, C, device=dev, generator=g)
    shapes = [("qkv", a, 3 * C, C, ops.ACT_NONE, None), ("fc1", a, F, C, ops.ACT_QUICK_GELU, None),
              ("out_proj/bf16", a, C, C, ops.ACT_NONE, xb), ("out_proj/f32", a, C, C, ops.ACT_NONE, xf),
              ("fc2/bf16", u, C, F, ops.ACT_NONE, xb), ("fc2/f32", u, C, F, ops.ACT_NONE, xf)]
    print(f"{'shape':14s} " + " ".join(f"mode{m:d}_us" for m in range(6)) + "   TFLOP/s(mode0)  TFLOP/s(mode1)  TFLOP/s(mode5)")
    for name, inp, n, k, act, res in shapes:
        w = (torch.randn(n, k, device=dev, generator=g) * k ** -0.5).to(torch.bfloat16)
        b = torch.zeros(n, device=dev)
        out = res if res is not None else torch.empty(M, n, dtype=torch.bfloat16, device=dev)
        ts = []
        for mode in range(6):
            lib.setok_debug_set_gemm_epi_mode(mode)
            ts.append(timeit(lambda: ops.gemm(inp, w, b, act=act, residual=res, out=out)))
        lib.setok_debug_set_gemm_epi_mode(0)
        fl = 2.0 * M * n * k
        print(f"{name:14s} " + " ".join(f"{t * 1e3:8.1f}" for t in ts) + f"   {fl / ts[0] / 1e9:8.1f}  {fl / ts[1] / 1e9:8.1f}  {fl / ts[5] / 1e9:8.1f}")
    # LayerNorm passes: bf16 -> bf16 and f32 -> bf16
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for nm, src in (("bf16", xb), ("f32", xf)):
        ms = timeit(lambda: ops.layernorm(src, gam, bet, out=a))
        print(f"layernorm {nm}->bf16 {M}x{C}: {ms * 1e3:8.1f} us  {M * C * (src.element_size() + 2) / ms / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    main()

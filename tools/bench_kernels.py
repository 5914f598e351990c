#!/usr/bin/env python
"""Per-kernel timings at the BASELINE config-2 shapes (CUDA events, L2-exceeding operands).  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import ops

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
    a = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    u = torch.randn(M, F, device=dev, generator=g).to(torch.bfloat16)
    x = torch.randn(M, C, device=dev, generator=g).to(torch.bfloat16)
    shapes = {"qkv": (a, 3 * C, C, ops.ACT_NONE, False), "out_proj": (a, C, C, ops.ACT_NONE, True),
              "fc1": (a, F, C, ops.ACT_QUICK_GELU, False), "fc2": (u, C, F, ops.ACT_NONE, True)}
    import ctypes
    from setok_b200 import _lib
    lib = _lib.load()
    cg = int(os.environ.get("SETOK_GEMM_CG", "0"))
    lib.setok_debug_set_gemm_cta_group.argtypes = [ctypes.c_int]
    lib.setok_debug_set_gemm_cta_group(cg)
    print(f"gemm cta_group setting: {cg} (0 = auto: pairs when M >= 256)")
    tot_ms = tot_fl = 0
    for name, (inp, n, k, act, res) in shapes.items():
        w = (torch.randn(n, k, device=dev, generator=g) * k ** -0.5).to(torch.bfloat16)
        b = torch.zeros(n, device=dev)
        out = x if res else torch.empty(M, n, dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.gemm(inp, w, b, act=act, residual=out if res else None, out=out))
        fl = 2.0 * M * n * k
        tot_ms += ms; tot_fl += fl
        print(f"gemm {name:9s} M={M} N={n} K={k}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s")
    print(f"gemm layer mix: {tot_ms * 1e3:8.1f} us  {tot_fl / tot_ms / 1e9:8.1f} TFLOP/s")
    qkv = torch.randn(M, 3 * C, device=dev, generator=g).to(torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, 16, 0.125, uniform_T=257))
    print(f"attention ViT T=257 hd=64 B=256: {ms * 1e3:8.1f} us  {4.0 * 257 * 257 * C * 256 / ms / 1e9:8.1f} TFLOP/s")
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    ms = timeit(lambda: ops.layernorm(x, gam, bet, out=a))
    print(f"layernorm bf16 {M}x{C}: {ms * 1e3:8.1f} us  {M * C * 4 / ms / 1e6:8.1f} GB/s")
    from setok_b200.synth import mog_features
    feats = mog_features(256, 256, C, 7, dev)
    noise = torch.rand(256, 256, device=dev)
    ms = timeit(lambda: ops.dpc_cluster(feats, noise, (16, 16), 16, 0.5, 64), reps=10)
    print(f"dpc_cluster B=256 N=256 C=1024: {ms * 1e3:8.1f} us")


if __name__ == "__main__":
    main()

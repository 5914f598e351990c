#!/usr/bin/env python
"""ViT attention at the bench shape (B = 256, T = 257, 16 heads of 64): whole-row kernel vs the chunked kernel, CUDA events."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import _lib, ops

dev = torch.device("cuda:0")


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
    lib.setok_debug_set_attention_fullrow.argtypes = [ctypes.c_int]
    for T, B, H in ((257, 256, 16), (197, 256, 12), (257, 64, 16)):
        C = 64 * H
        qkv = torch.randn(B * T, 3 * C, device=dev).to(torch.bfloat16)
        outs = {}
        for mode, name in ((2, "fullrow"), (0, "chunked")):
            lib.setok_debug_set_attention_fullrow(mode)
            ms = timeit(lambda: ops.attention(qkv, H, 0.125, uniform_T=T))
            outs[name] = ops.attention(qkv, H, 0.125, uniform_T=T).float()
            print(f"attention T={T} B={B} heads={H} {name:8s}: {ms * 1e3:8.1f} us  {4.0 * T * T * C * B / ms / 1e9:8.1f} TFLOP/s")
        d = (outs["fullrow"] - outs["chunked"]).abs().max().item()
        print(f"   max |fullrow - chunked| = {d:.3e}")
        if T == 257 and B == 256:
            lib.setok_debug_set_attention_fullrow_dbg.argtypes = [ctypes.c_int]
            lib.setok_debug_set_attention_fullrow(2)
            for flags, what in ((1, "no pass 1"), (2, "no exp2"), (3, "no pass 1, no exp2"), (4, "no P store"), (8, "no O store"), (15, "all off"), (16, "packed bf16x2 exp2")):
                lib.setok_debug_set_attention_fullrow_dbg(flags)
                ms = timeit(lambda: ops.attention(qkv, H, 0.125, uniform_T=T))
                extra = ""
                if flags == 16:
                    d16 = (ops.attention(qkv, H, 0.125, uniform_T=T).float() - outs["fullrow"]).abs().max().item()
                    extra = f"   max |packed - f32 exp| = {d16:.3e}"
                print(f"   fullrow with {what:20s}: {ms * 1e3:8.1f} us{extra}")
            lib.setok_debug_set_attention_fullrow_dbg(0)
        lib.setok_debug_set_attention_fullrow(1)


if __name__ == "__main__":
    main()

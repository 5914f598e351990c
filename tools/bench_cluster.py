#!/usr/bin/env python
"""Clustering (a3+a4) A/B: multi-kernel path vs the fused persistent kernel (IEEE-half 3-term split = default, bf16 4-term, bf16 3-term), agreement of
their integer outputs, and timings at the BASELINE config-2 shape.  GPU only."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import _lib, ops
from setok_b200.synth import mog_features

dev = torch.device("cuda:0")
lib = _lib.load()
lib.setok_debug_set_dpc_fused.argtypes = [ctypes.c_int]


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


def score_torch(x, noise, k, dtype):
    """tokenizer.py:82-101 in plain torch at `dtype` (float64 = ground truth with the exact difference form;
    float32 = the reference's own arithmetic, cdist in its matmul form)."""
    x = x.to(dtype)
    C = x.shape[-1]
    if dtype == torch.float64:
        D = torch.cdist(x, x, compute_mode="donot_use_mm_for_euclid_dist") / (C ** 0.5)
    else:
        D = torch.cdist(x, x) / (C ** 0.5)
    dn, _ = torch.topk(D, k=k, dim=-1, largest=False)
    dens = (-(dn ** 2).mean(dim=-1)).exp() + noise.to(dtype) * 1e-6
    mask = (dens[:, None, :] > dens[:, :, None]).to(dtype)
    dmax = D.max(dim=-1)[0][:, None, :]
    dist = (D * mask + dmax * (1 - mask)).min(dim=-1)[0]
    return dist * dens


def accuracy():
    """Score error of every path against float64 ground truth (and the reference's own fp32 error, for scale)."""
    B, N, C, k = 32, 256, 1024, 16
    feats = mog_features(B, N, C, 11, dev)
    noise = torch.rand(B, N, device=dev)
    zero_pos = torch.zeros(N, C, device=dev)
    truth = score_torch(feats, noise, k, torch.float64)
    ref32 = score_torch(feats, noise, k, torch.float32)
    rel = lambda s: ((s.double() - truth).abs() / truth.abs().clamp_min(1e-9))
    e = rel(ref32)
    print(f"accuracy vs fp64 truth (B={B}, k={k}):  torch fp32 reference formula: max {float(e.max()):.2e} mean {float(e.mean()):.2e}")
    for mode in (0, 1, 2, 3):
        lib.setok_debug_set_dpc_fused(mode)
        out = ops.dpc_cluster(feats, noise, (N, 1), k, 0.5, 64, pos_table=zero_pos)
        e = rel(out[2])
        print(f"  mode {mode}: score rel err max {float(e.max()):.2e} mean {float(e.mean()):.2e}")
    lib.setok_debug_set_dpc_fused(1)


def main():
    accuracy()
    B, N, C = 256, 256, 1024
    for dtype in (torch.float32, torch.bfloat16):
        feats = mog_features(B, N, C, 7, dev).to(dtype)
        noise = torch.rand(B, N, device=dev)
        outs = {}
        for mode, name in ((0, "multi-kernel"), (1, "fused f16 3t"), (2, "fused bf16 4t"), (3, "fused bf16 3t")):
            lib.setok_debug_set_dpc_fused(mode)
            for k in (16, 64):
                run = lambda: ops.dpc_cluster(feats, noise, (16, 16), k, 0.5, 64)
                out = run()
                torch.cuda.synchronize()
                ms = timeit(run)
                outs[(mode, k)] = out
                Kc = out[4].float()
                nbytes = B * (N * C * feats.element_size() + N * 16) + float(Kc.sum()) * 8
                print(f"{str(dtype):15s} {name:13s} k={k:2d}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s algorithmic  K min/mean/max "
                      f"{int(Kc.min())}/{float(Kc.mean()):.1f}/{int(Kc.max())}")
        for k in (16, 64):
            ref = outs[(0, k)]
            for mode in (1, 2, 3):
                o = outs[(mode, k)]
                same_k = int((o[4] == ref[4]).sum())
                same_lab = float((o[1] == ref[1]).float().mean())
                same_down = float((o[3] == ref[3]).float().mean())
                serr = float(((o[2] - ref[2]).abs() / ref[2].abs().clamp_min(1e-6)).max())
                print(f"  {str(dtype):15s} k={k} mode {mode} vs multi-kernel: K equal {same_k}/{B}, labels equal {same_lab:.6f}, "
                      f"index_down equal {same_down:.6f}, score max rel diff {serr:.2e}, x_pos equal {bool(torch.equal(o[0], ref[0]))}")
    lib.setok_debug_set_dpc_fused(1)


if __name__ == "__main__":
    main()

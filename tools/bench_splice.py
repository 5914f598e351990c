#!/usr/bin/env python
"""Device splice (setokim_arch.py:241-354) at BASELINE config 4's consumer shape: 64 samples x 512 text tokens + one image of
K in 8..128 rows each, H = 4096 bf16 (Vicuna-7B).  Reports achieved HBM GB/s (rows read + rows written).  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import RaggedTokens, prepare_inputs_labels_for_multimodal

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(62)
B, L, V, H = 64, 512, 32000, 4096
emb = torch.randn(V, H, device=dev, generator=gen).to(torch.bfloat16)
ids = torch.randint(0, V, (B, L), device=dev, generator=gen)
ids[torch.arange(B, device=dev), torch.randint(0, L, (B,), device=dev, generator=gen)] = -200
K = torch.randint(8, 129, (B,), device=dev, generator=gen)
offsets = torch.zeros(B + 1, dtype=torch.int32, device=dev)
offsets[1:] = torch.cumsum(K, 0)
rows = torch.randn(int(offsets[-1]), H, device=dev, generator=gen).to(torch.bfloat16)
mask = torch.ones(B, L, dtype=torch.bool, device=dev)
rt = RaggedTokens(rows, offsets)
run = lambda: prepare_inputs_labels_for_multimodal(ids, None, mask, None, ids, rt, emb)
for _ in range(3):
    out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
T = out[4].shape[1]
moved = (float(mask.sum()) - B + float(K.sum())) * H * 2 + B * T * H * 2       # rows gathered + the (B, T, H) output written
print(f"splice B={B} L={L} H={H} bf16: {ms * 1e3:.1f} us per call (two phases incl. the one host read of max_len), T={T}; "
      f"{moved / ms / 1e6:.0f} GB/s over {moved / 1e6:.0f} MB moved")

#!/usr/bin/env python
"""BASELINE config 3's decoder half: the detokenizer over a ragged batch of 128 images at 336^2 (576 queries), hidden 768 /
12 heads, Q-Former 6 layers with cross-attention every 2, decoder 768 x 16 blocks x 12 heads.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from setok_b200 import RaggedTokens, SetokDeTokenizer

dev = torch.device("cuda:0")
B, Q, H, Dd, DEPTH, QL, CT = 128, 576, 768, 768, 16, 6, 1024
torch.manual_seed(0)
det = SetokDeTokenizer(token_feat_dim=CT, hidden_dim=H, patch_size=14, image_size=336, decoder_embed_dim=Dd, decoder_nheads=12,
                       decoder_depth=DEPTH, num_hidden_layers=QL, cross_attention_freq=2).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
K = torch.randint(8, 129, (B,), device=dev, generator=g)
offsets = torch.zeros(B + 1, dtype=torch.int32, device=dev)
offsets[1:] = torch.cumsum(K, 0)
total = int(offsets[-1])
tokens = torch.randn(total, CT, device=dev, generator=g).to(torch.bfloat16)
rt = RaggedTokens(tokens, offsets)
for _ in range(3):
    out = det(rt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    out = det(rt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
I = 4 * H
ncross = (QL + 1) // 2
fl_q = QL * (2 * Q * H * 3 * H + 4 * Q * Q * H + 2 * Q * H * H + 4 * Q * H * I) + ncross * (2 * Q * H * H + 2 * Q * H * H)
fl_d = DEPTH * (24 * Q * Dd * Dd + 4 * Q * Q * Dd) + 2 * Q * H * Dd
flops = B * (fl_q + fl_d) + ncross * total * (4 * H * H + 4 * Q / 1 * 0) + 2 * total * CT * H
print(f"detokenizer B={B} Q={Q} hidden={H} dec={Dd}x{DEPTH}: {ms:.2f} ms  {B / ms * 1e3:.0f} images/s  {flops / ms / 1e9:.0f} TFLOP/s "
      f"(sum K = {total}, K mean {total / B:.1f})")

# A/B: the CUDA-core cross-attention kernel against the tensor-core one (default), same inputs, outputs compared
import ctypes
from setok_b200 import _lib
lib = _lib.load()
lib.setok_debug_set_cross_attention_tc.argtypes = [ctypes.c_int]
lib.setok_debug_set_cross_attention_tc.restype = None


def timed():
    for _ in range(2):
        o = det(rt)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        o = det(rt)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, o.float()


lib.setok_debug_set_cross_attention_tc(0)
ms_cc, o_cc = timed()
lib.setok_debug_set_cross_attention_tc(1)
ms_tc, o_tc = timed()
err = float((o_tc - o_cc).norm() / o_cc.norm())
print(f"cross-attention A/B: CUDA-core {ms_cc:.2f} ms, tcgen05 {ms_tc:.2f} ms per detokenizer call; outputs differ by {err:.2e} rel-Frobenius")

#!/usr/bin/env python
"""Per-step wall-clock trace of the config-4 end-to-end pipeline (stream_tokenize + projector) to locate stalls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, setok_b200
from setok_b200.pipeline import stream_tokenize

dev = torch.device("cuda:0")
B = 64
tok = bench.build_model(dev)
torch.manual_seed(2)
proj = setok_b200.build_vision_projector("mlp2x_gelu", mm_hidden_size=1024, hidden_size=4096).to(dev)
u8, imgs_h, noise_h = bench.host_batch(4, 0, B)
imgs_h = imgs_h.to(torch.bfloat16)
h_img, h_noise = imgs_h.pin_memory(), noise_h.pin_memory()
post = lambda r, i, s_: (proj(r), i, s_)
for rnd in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = []
    for res in stream_tokenize(tok, ((h_img, h_noise) for _ in range(6)), post=post, k=16):
        ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print(f"round {rnd}: results at (ms) " + " ".join(f"{1e3 * t:.1f}" for t in ts) + f" | total {1e3 * (time.perf_counter() - t0):.1f} ms; "
          f"allocated {torch.cuda.memory_allocated() / 1e9:.2f} GB reserved {torch.cuda.memory_reserved() / 1e9:.2f} GB", flush=True)
# the same without the projector
for rnd in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = []
    for res in stream_tokenize(tok, ((h_img, h_noise) for _ in range(6)), k=16):
        ts.append(time.perf_counter() - t0)
    print(f"no projector round {rnd}: " + " ".join(f"{1e3 * t:.1f}" for t in ts), flush=True)
# profile one round with the projector
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for res in stream_tokenize(tok, ((h_img, h_noise) for _ in range(6)), post=post, k=16):
    pass
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

#!/usr/bin/env python
"""Why did bench.py's `parity` leg report O(1) tower error while tests/test_gpu_model.py::test_bench_config_full_depth_parity
measures 2e-3 on the same architecture?  Replays the bench's legs one at a time and re-measures after each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from oracle import setok_oracle as O

dev = torch.device("cuda:0")
tok = bench.build_model(dev)
tower = tok.image_feature_encoder
u8, imgs_h, noise_h = bench.host_batch(2, 0)
n = 8
images, noise = imgs_h.to(dev), noise_h.to(dev)


def check(tag):
    sd = {k: v.detach().float().cpu() for k, v in tower.vision_tower.state_dict().items()}
    h = sum(float(v.double().sum()) for v in sd.values())
    with torch.no_grad():
        ref = O.tower_features(imgs_h[:n], sd, patch=14, heads=16, layers=24, select_layer=-2)
    got = tower(images[:n]).float().cpu()
    print(f"{tag}: rel-Frobenius {float((got - ref).norm() / ref.norm()):.4e}  weight checksum {h:.6f}  |ref| {float(ref.norm()):.2f} |got| {float(got.norm()):.2f}", flush=True)


check("fresh")
bench.k32_variant(tok, images, noise)
check("after k32_variant")
bench.gpu_eager_baseline(dev, tok, images, noise, head_sample=2)
check("after gpu_eager_baseline")
orc = bench.CpuOracle(2, n)
orc.tp = {k: v.detach().float().cpu() for k, v in tower.vision_tower.state_dict().items()}
secs, st = orc.run(n)
got = tower(images[:n]).float().cpu()
print("bench's own expression:", float((got - st["feats"]).norm() / st["feats"].norm()))

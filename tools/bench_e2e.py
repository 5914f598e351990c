#!/usr/bin/env python
"""Where does the end-to-end (host buffers in, host result out) step spend its time?  Times, on one GPU: the pinned H2D copy
alone, the resident step, and the streamed step of pipeline.stream_tokenize for float32 and uint8 images.

    python tools/bench_e2e.py [--steps 10]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from setok_b200.pipeline import stream_tokenize
    from setok_b200.synth import mondrian_images
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    images = mondrian_images(bench.BATCH, 224, 1234, "cpu")
    noise = torch.rand(bench.BATCH, 256, generator=torch.Generator().manual_seed(99))
    h_img, h_noise = images.pin_memory(), noise.pin_memory()
    d_img, d_noise = images.to(dev), noise.to(dev)
    out = {}

    def timed(fn, n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    def h2d(n):
        for _ in range(n):
            h_img.to(dev, non_blocking=True)
    h2d(2)
    out["h2d_ms_154MB"] = timed(h2d, 5)

    def resident(n):
        for _ in range(n):
            tok(d_img, k=bench.KNN_K, noise=d_noise)
    resident(3)
    out["resident_ms"] = timed(resident, args.steps)

    def streamed(src):
        def run(n):
            for _ in stream_tokenize(tok, ((src, h_noise) for _ in range(n)), k=bench.KNN_K):
                pass
        return run
    streamed(h_img)(3)
    out["streamed_f32_ms"] = timed(streamed(h_img), args.steps)
    out["streamed_f32_ms_40"] = timed(streamed(h_img), 4 * args.steps)
    u8 = torch.randint(0, 256, (bench.BATCH, 3, 224, 224), dtype=torch.uint8).pin_memory()
    streamed(u8)(3)
    out["streamed_u8_ms"] = timed(streamed(u8), args.steps)
    # host-side cost of one forward call (launch overhead), GPU idle-synchronised before
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tok(d_img, k=bench.KNN_K, noise=d_noise)
    out["host_launch_ms"] = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

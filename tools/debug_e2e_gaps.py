#!/usr/bin/env python
"""Where does the streamed (host buffers in / out) step of config 2 lose time against the resident one?  Wraps the tokenizer with
CUDA events on the compute stream: per batch the device-side duration of its kernels and the idle gap before the next batch's
first kernel, for the resident loop and for pipeline.stream_tokenize.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from setok_b200.pipeline import stream_tokenize


class Timed:
    def __init__(self, tok):
        self.tok, self.ev = tok, []
        self.device = tok.device

    def __call__(self, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = self.tok(*a, **k)
        e.record()
        self.ev.append((s, e))
        return out

    def report(self, what):
        torch.cuda.synchronize()
        ev = self.ev[4:]                                     # skip the warm-up batches
        dur = [s.elapsed_time(e) for s, e in ev]
        gap = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)]
        total = ev[0][0].elapsed_time(ev[-1][1]) / len(ev)
        print(f"{what:28s}: kernels of a batch {sum(dur) / len(dur):7.3f} ms (min {min(dur):.3f} max {max(dur):.3f}), idle gap before the next batch "
              f"{sum(gap) / len(gap):6.3f} ms (max {max(gap):.3f}), per batch overall {total:7.3f} ms", flush=True)
        self.ev = []


def main():
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    images, noise = imgs_h.to(dev), noise_h.to(dev)
    h_u8, h_noise = u8.pin_memory(), noise_h.pin_memory()
    T = Timed(tok)
    n = 16
    for rnd in range(2):
        for _ in range(n):
            T(images, k=bench.KNN_K, noise=noise)
        T.report("resident, float images")
        d_u8 = u8.to(dev)
        for _ in range(n):
            T(d_u8, k=bench.KNN_K, noise=noise)
        T.report("resident, uint8 images")
        for _ in stream_tokenize(T, ((h_u8, h_noise) for _ in range(n)), k=bench.KNN_K):
            pass
        T.report("streamed, uint8 images")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Is the tokenizer step launch-bound anywhere?  Times the BASELINE config-2 step (256 images) eagerly (161 launches from the host
per step, programmatic dependent launch between them) and as a CUDA-graph replay of the same launches.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    u8, imgs_h, noise_h = bench.host_batch(2, 0)
    images, noise = imgs_h.to(dev), noise_h.to(dev)

    def step():
        return tok(images, k=bench.KNN_K, noise=noise)
    for rnd in range(3):
        ms_e = bench.cuda_time(step, 8, warm=3)
        ms_g, how = bench.cuda_time_graph(step, 8, warm=3)
        print(f"round {rnd}: eager {ms_e:7.3f} ms per step, {how} {ms_g:7.3f} ms per step", flush=True)


if __name__ == "__main__":
    main()

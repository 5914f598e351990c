#!/usr/bin/env python
"""Does the ViT tower run faster when the 256-image batch of BASELINE config 2 walks the layers in micro-batches whose
activations stay L2-resident (126 MB) between consecutive kernels?  Times the tower on the same 256 images split into
chunks of B_mb images (whole tower per chunk, one after the other on the same stream), CUDA events, back to back so that the
power cap settles.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    dev = torch.device("cuda:0")
    tok = bench.build_model(dev)
    tower = tok.image_feature_encoder
    from setok_b200.synth import mondrian_images
    images = mondrian_images(256, 224, 1234, "cpu").to(dev)
    sizes = [int(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else "256,128,72,64,32".split(","))]
    reps = int(os.environ.get("REPS", "6"))
    full = tower(images)
    for mb in sizes + sizes[:1]:
        chunks = [images[i:i + mb] for i in range(0, 256, mb)]

        def run():
            return [tower(c) for c in chunks]
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        same = torch.equal(torch.cat(out, 0), full)
        print(f"micro-batch {mb:4d} ({[c.shape[0] for c in chunks]}): tower {ms:7.2f} ms per 256 images = {256 / ms * 1e3:7.1f} images/s; "
              f"bit-identical to the one-batch tower: {same}", flush=True)


if __name__ == "__main__":
    main()

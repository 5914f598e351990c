#!/usr/bin/env python
"""A/B of SETOK_VIT_LN_FOLD on the BASELINE config-2 tower (256 images, ViT-L/14, 23 layers): the LayerNorms folded into the
GEMMs around them against the separate LayerNorm passes.  Alternates the two towers so that both see the same power state;
reports time per 256 images, the difference between the two outputs and per-GEMM times of the folded variants.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import setok_b200


def main():
    dev = torch.device("cuda:0")
    towers = {}
    for name, fold in (("folded", True), ("separate", False)):
        torch.manual_seed(0)
        tok = setok_b200.SetokTokenizer("siglip-synthetic-vit-l-14", vision_config=dict(bench.VIT, image_size=224), tower_ln_fold=fold, **bench.HEAD)
        towers[name] = tok.to(dev).image_feature_encoder
    from setok_b200.synth import mondrian_images
    nimg = int(os.environ.get("IMAGES", "256"))
    images = mondrian_images(nimg, 224, 1234, "cpu").to(dev)
    out = {k: t(images) for k, t in towers.items()}
    d = (out["folded"] - out["separate"]).float()
    print(f"folded vs separate: rel-Frobenius {float(d.norm() / out['separate'].float().norm()):.3e}, "
          f"max abs / max {float(d.abs().max() / out['separate'].float().abs().max()):.3e}", flush=True)
    reps = int(os.environ.get("REPS", "8"))
    if os.environ.get("EPI_DIRECT"):      # -1 automatic, 0 / 1: force the transposing / row-owner epilogue where both exist (A/B of the producers)
        import ctypes
        from setok_b200 import _lib
        lib = _lib.load()
        lib.setok_debug_set_gemm_epi_direct.argtypes = [ctypes.c_int]
        lib.setok_debug_set_gemm_epi_direct(int(os.environ["EPI_DIRECT"]))
        print("gemm epilogue override:", os.environ["EPI_DIRECT"])
    for rnd in range(3):
        for name, t in towers.items():
            for _ in range(2):
                t(images)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                t(images)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"round {rnd} {name:9s}: tower {ms:7.2f} ms per {nimg} images = {nimg / ms * 1e3:7.1f} images/s", flush=True)


if __name__ == "__main__":
    main()

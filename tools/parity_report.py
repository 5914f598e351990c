#!/usr/bin/env python
"""Float parity of the tower at the BENCH architecture (24-layer ViT-L/14, BASELINE config 2 / config 3), per depth and for
both residual-stream precisions, against the fp32 oracle on the host; plus the A/B step time of the two precisions.

    python tools/parity_report.py [--images 8] [--size 224] [--out gpurun_out/parity_r02.json]

Prints one JSON document: for each depth n in --depths the relative Frobenius / normalised max error of
hidden_states[n][:, 1:] (ours vs oracle), the same for torch's own bf16 evaluation of the oracle formula, and ms per
256-image tower pass with the bf16 and the f32 residual stream."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return {"max": float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6)), "fro": float((got - ref).norm() / ref.norm().clamp_min(1e-6))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--depths", default="1,6,12,18,23,24")
    ap.add_argument("--time-batch", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_r02.json"))
    args = ap.parse_args()
    from oracle import setok_oracle as O
    import setok_b200
    from setok_b200.synth import mondrian_images
    dev = torch.device("cuda:0")
    C, L, H, P, IMG = 1024, 24, 16, 14, args.size
    vc = dict(hidden_size=C, intermediate_size=4 * C, num_hidden_layers=L, num_attention_heads=H, image_size=IMG, patch_size=P)
    # the bench's own weights: HF CLIPVisionModel initialisation under torch.manual_seed(0) (bench.py:build_model)
    torch.manual_seed(0)
    tok0 = setok_b200.SetokTokenizer("siglip-parity", hidden_dim=C, token_feat_dim=C, min_cluster_num=64, dim_feedforward=4096,
                                     mm_vision_select_layer=-2, vision_config=vc)
    tp = {k: v.detach().clone() for k, v in tok0.image_feature_encoder.vision_tower.state_dict().items()}
    del tok0
    imgs = mondrian_images(args.images, IMG, 1234, "cpu")
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        hs = O.clip_vit_hidden_states(imgs, tp, patch=P, heads=H, layers=L, n_layers_run=L)
        pb = {k: v.to(torch.bfloat16) for k, v in tp.items()}
        hs_b = O.clip_vit_hidden_states(imgs.to(torch.bfloat16), pb, patch=P, heads=H, layers=L, n_layers_run=L)
    depths = [int(d) for d in args.depths.split(",")]
    doc = {"config": f"ViT-L/14 @{IMG}, {args.images} Mondrian images (seed 1234), hidden_states[n][:, 1:]", "depths": {}}
    for n in depths:
        doc["depths"][str(n)] = {"torch_bf16": err(hs_b[n][:, 1:], hs[n][:, 1:])}
    for mode, f32 in (("bf16_residual", False), ("f32_residual", True)):
        tok = setok_b200.SetokTokenizer("siglip-parity", hidden_dim=C, token_feat_dim=C, min_cluster_num=64, dim_feedforward=4096,
                                        mm_vision_select_layer=-2, vision_config=vc, tower_residual_f32=f32)
        tok.image_feature_encoder.vision_tower.load_state_dict(tp)
        tok = tok.to(dev)
        tower = tok.image_feature_encoder
        for n in depths:
            tower.select_layer = n
            got = tower(imgs.to(dev))
            doc["depths"][str(n)][mode] = err(got, hs[n][:, 1:])
        tower.select_layer = -2
        big = mondrian_images(args.time_batch, IMG, 99, dev)
        for _ in range(3):
            tower(big)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            tower(big)
        e1.record()
        torch.cuda.synchronize()
        doc[mode + "_ms_per_tower_pass"] = e0.elapsed_time(e1) / 5
        del tok, tower, big
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()

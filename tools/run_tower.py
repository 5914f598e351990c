#!/usr/bin/env python
"""Runs the BASELINE config-2 tower (256 images, ViT-L/14) a few times: the target of ncu launch lists.  FOLD=0 keeps the separate
LayerNorm passes (SETOK_VIT_LN_FOLD off); LAYERS limits the depth."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import setok_b200

dev = torch.device("cuda:0")
torch.manual_seed(0)
vit = dict(bench.VIT, image_size=224)
if os.environ.get("LAYERS"):
    vit["num_hidden_layers"] = int(os.environ["LAYERS"])
tok = setok_b200.SetokTokenizer("siglip-synthetic-vit-l-14", vision_config=vit, tower_ln_fold=os.environ.get("FOLD", "1") != "0", **bench.HEAD).to(dev)
from setok_b200.synth import mondrian_images
images = mondrian_images(256, 224, 1234, "cpu").to(dev)
for _ in range(int(os.environ.get("CALLS", "3"))):
    tok.image_feature_encoder(images)
torch.cuda.synchronize()
print("done")

#!/usr/bin/env python
"""Generates tests/golden/preprocess.npz: inputs + outputs of the reference's image preprocessing for a handful of image
sizes, computed with PIL itself (the dependency whose arithmetic the path uses) arranged exactly as the pinned
transformers 4.46.3 CLIPImageProcessor.preprocess + the reference's expand2square arrange it
(/root/reference/src/mm_utils.py:152-182; transformers image_transforms.resize -> PIL.Image.resize(BICUBIC);
center_crop; rescale; normalize).  The transformers installed in this image (5.x) resizes with torchvision instead of PIL and
differs by one grey level on ~1 % of the pixels, so it is NOT used for the fixture; the expand2square function object is the
reference's own, loaded from /root/reference when present.

    python oracle/make_golden_preprocess.py [out.npz]"""
import importlib.util
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def reference_expand2square():
    """The reference's own function object (mm_utils.py:152-163), extracted without importing the package's heavy deps."""
    path = "/root/reference/src/mm_utils.py"
    if not os.path.exists(path):
        return None
    src = open(path).read()
    start = src.index("def expand2square")
    end = src.index("\ndef ", start + 1)
    ns = {"Image": Image}
    exec(src[start:end], ns)
    return ns["expand2square"]


def pipeline(img: np.ndarray, size: int, pad: bool, expand2square):
    pil = Image.fromarray(img)
    if pad:
        pil = expand2square(pil, tuple(int(x * 255) for x in MEAN))
    W, H = pil.size
    short, long = (W, H) if W <= H else (H, W)
    new_short, new_long = size, int(size * long / short)
    oh, ow = (new_long, new_short) if W <= H else (new_short, new_long)
    r = np.asarray(pil.resize((ow, oh), resample=Image.BICUBIC, reducing_gap=None))
    top, left = (oh - size) // 2, (ow - size) // 2
    c = r[top:top + size, left:left + size]
    u8 = np.ascontiguousarray(c.transpose(2, 0, 1))
    x = (u8.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(MEAN, np.float32).reshape(3, 1, 1)) / np.asarray(STD, np.float32).reshape(3, 1, 1)
    return u8, x


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    e2s = reference_expand2square()
    assert e2s is not None, "needs /root/reference for the reference's expand2square"
    rng = np.random.default_rng(2024)
    cases = [(60, 90, 56, False), (60, 90, 56, True), (97, 41, 56, True), (150, 200, 112, False), (33, 33, 56, False), (240, 180, 112, True),
             (56, 56, 56, False), (40, 300, 28, False)]
    d = {"cases": np.asarray(cases, dtype=np.int32)}
    for i, (H, W, S, pad) in enumerate(cases):
        # structured content (gradients + blocks + noise) so that resampling errors cannot hide in flat regions
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.stack([(yy * 255 // max(H - 1, 1)), (xx * 255 // max(W - 1, 1)), ((yy // 7 + xx // 5) % 2) * 255], -1).astype(np.int64)
        img = np.clip(img + rng.integers(-40, 41, img.shape), 0, 255).astype(np.uint8)
        u8, x = pipeline(img, S, bool(pad), e2s)
        d[f"in{i}"], d[f"u8_{i}"], d[f"f32_{i}"] = img, u8, x
    np.savez_compressed(out, **d)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

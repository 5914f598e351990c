"""TEST INFRASTRUCTURE — CPU restatement of the reference's image preprocessing, numpy only.

Path restated: `process_images` (reference src/mm_utils.py:166-182) = optional `expand2square` (:152-163) followed by
`CLIPImageProcessor.preprocess` of the pinned transformers 4.46.3 (pyproject.toml:18): resize (shortest edge, PIL BICUBIC on
the uint8 image: image_transforms.resize -> PIL.Image.resize), center crop, rescale (float32(float64(u8) * 1/255)), normalize
((x - mean) / std in float32).  The arithmetic of the resize lives in the third-party dependency Pillow (not vendored in the
reference): `ImagingResample` of Pillow's libImaging/Resample.c, whose published algorithm is restated in `resample_u8`:
separable convolution, double-precision filter coefficients normalised per output pixel and converted to 22-bit fixed point
(PRECISION_BITS = 32 - 8 - 2), horizontal pass then vertical pass, each rounding to uint8 with clipping.

Pinned: `tests/test_preprocess_cpu.py` checks `resample_u8` bit-for-bit against `PIL.Image.resize` itself (Pillow is installed
here and on the GPU box) and the whole pipeline against the committed fixture `tests/golden/preprocess.npz`
(`oracle/make_golden_preprocess.py`, generated with PIL + the pipeline of transformers 4.46.3).  Only tests/, smoke() and
bench.py's cpu_baseline leg may import this module."""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: float) -> float:
    """Pillow Resample.c:bicubic_filter (a = -0.5)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Resample.c:precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): returns the fixed-point taps
    kk[out_size, ksize] (int32) and bounds[out_size, 2] = (first source index, number of taps)."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def _pass(x: np.ndarray, kk: np.ndarray, bounds: np.ndarray, axis: int) -> np.ndarray:
    x = np.moveaxis(x.astype(np.int64), axis, 0)
    out = np.empty((kk.shape[0],) + x.shape[1:], dtype=np.int64)
    for i in range(kk.shape[0]):
        lo, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.tensordot(kk[i, :n].astype(np.int64), x[lo:lo + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resample_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BICUBIC) of a uint8 (H, W, C) image: horizontal pass, then vertical pass; a pass
    whose size does not change is skipped (Resample.c:ImagingResampleInner)."""
    H, W = img.shape[:2]
    x = img
    if out_w != W:
        x = _pass(x, *precompute_coeffs(W, out_w), axis=1)
    if out_h != H:
        x = _pass(x, *precompute_coeffs(H, out_h), axis=0)
    return x


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """mm_utils.py:152-163 on a (H, W, 3) array: pad to a square canvas of the background colour, the image centred along
    the short side (integer division of the slack, as Image.paste is called there)."""
    H, W = img.shape[:2]
    if W == H:
        return img
    D = max(H, W)
    out = np.empty((D, D, 3), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    if W > H:
        y0 = (W - H) // 2
        out[y0:y0 + H, :, :] = img
    else:
        x0 = (H - W) // 2
        out[:, x0:x0 + W, :] = img
    return out


def resize_output_size(H: int, W: int, shortest_edge: int) -> Tuple[int, int]:
    """transformers 4.46.3 image_transforms.get_resize_output_image_size(size=int, default_to_square=False)."""
    short, long = (W, H) if W <= H else (H, W)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if W <= H else (new_short, new_long)


def preprocess_u8(img: np.ndarray, size: int, pad: bool, mean: Sequence[float] = CLIP_MEAN) -> np.ndarray:
    """uint8 (H, W, 3) -> uint8 (3, size, size): [expand2square with int(255 * mean)] -> resize shortest edge -> center crop."""
    if pad:
        img = expand2square(img, tuple(int(m * 255) for m in mean))
    H, W = img.shape[:2]
    oh, ow = resize_output_size(H, W, size)
    r = resample_u8(img, oh, ow)
    top, left = (oh - size) // 2, (ow - size) // 2
    return np.ascontiguousarray(r[top:top + size, left:left + size].transpose(2, 0, 1))


def rescale_normalize(u8_chw: np.ndarray, mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD) -> np.ndarray:
    """CLIPImageProcessor.rescale + normalize (4.46.3): float32(float64(u8) * (1/255)), then (x - mean) / std in float32."""
    x = (u8_chw.astype(np.float64) * (1.0 / 255.0)).astype(np.float32)
    m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return (x - m) / s

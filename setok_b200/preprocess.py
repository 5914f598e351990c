"""Image preprocessing on the device: the host-side mirror of the reference's `process_images`
(src/mm_utils.py:166-182) and `expand2square` (:152-163) over libsetok_b200's `setok_preprocess_u8`.

The reference pads / resizes / crops every image on the host with PIL, converts it to float32 and uploads 12 bytes per pixel
of the *resized* image.  Here the decoded uint8 pixels go up as they are (3 bytes per source pixel), one batched launch pair
resizes them with Pillow's exact fixed-point bicubic arithmetic, and the tower's patch-embedding pass does rescale +
normalize (`CLIPVisionTower.forward(uint8)`), so nothing float ever crosses PCIe.  The only host arithmetic is the per-size
tap table (Pillow's `precompute_coeffs`, double precision -> 22-bit fixed point), cached per (source size, target size)."""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import SetokError

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@lru_cache(maxsize=512)
def _taps(in_size: int, out_size: int, first: int, count: int) -> Tuple[np.ndarray, int]:
    """Pillow Resample.c:precompute_coeffs + normalize_coeffs_8bpc for output pixels [first, first + count): an int32 array
    [count, 2 + ksize] of (first source index, tap count, taps...) and ksize.  float64 throughout, as in C."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xx = np.arange(first, first + count, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast truncates; arguments are >= -support
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    k = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((k + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(k < xmax[:, None], w, 0.0)
    ww = np.zeros(count, dtype=np.float64)
    for j in range(ksize):                                                    # same left-to-right summation order as the C loop
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww[:, None] != 0.0, ww[:, None], 1.0), w)
    fixed = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64))
    out = np.zeros((count, 2 + ksize), dtype=np.int32)
    out[:, 0], out[:, 1], out[:, 2:] = xmin, xmax, fixed
    return out, ksize


def resize_output_size(H: int, W: int, shortest_edge: int) -> Tuple[int, int]:
    """transformers 4.46.3 get_resize_output_image_size(size=int, default_to_square=False) -> (height, width)."""
    short, long = (W, H) if W <= H else (H, W)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if W <= H else (new_short, new_long)


def _to_u8_hwc(img) -> torch.Tensor:
    if isinstance(img, torch.Tensor):
        t = img
    else:
        if hasattr(img, "convert"):                      # PIL image
            img = img.convert("RGB")
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(img)))
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise SetokError(f"images must be uint8 (H, W, 3); got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def preprocess_images(images: Sequence, size: int, pad: bool = False, image_mean=(0.48145466, 0.4578275, 0.40821073),
                      device: Optional[torch.device] = None) -> torch.Tensor:
    """Decoded images (uint8 (H, W, 3) tensors / arrays / PIL images, any sizes) -> uint8 (B, 3, size, size) on the device:
    [expand2square with the mean colour] -> bicubic resize of the shortest edge to `size` -> center crop."""
    imgs = [_to_u8_hwc(im) for im in images]
    if not imgs:
        raise SetokError("preprocess_images: empty batch")
    dev = torch.device(device) if device is not None else (imgs[0].device if imgs[0].is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if dev.type != "cuda":
        raise SetokError("setok_b200 kernels need a CUDA device; there is no CPU fallback")
    S = int(size)
    B = len(imgs)
    descs = (_lib.ResizeDesc * B)()
    tables, t_off, tmp_off, max_rows, keep = [], 0, 0, 1, []
    for i, im in enumerate(imgs):
        H, W = int(im.shape[0]), int(im.shape[1])
        d_im = im.to(dev, non_blocking=True)
        keep.append(d_im)
        Hp, Wp, px, py = H, W, 0, 0
        if pad and H != W:                               # mm_utils.py:152-163
            Dq = max(H, W)
            Hp = Wp = Dq
            if W > H:
                py = (W - H) // 2
            else:
                px = (H - W) // 2
        oh, ow = resize_output_size(Hp, Wp, S)
        if oh < S or ow < S:
            raise SetokError(f"image {i}: resized size {oh}x{ow} is smaller than the crop {S} (cannot happen for shortest-edge resizing)")
        top, left = (oh - S) // 2, (ow - S) // 2
        idx, idy = int(ow == Wp), int(oh == Hp)
        kx, ksx = _taps(Wp, ow, left, S)
        ky, ksy = _taps(Hp, oh, top, S)
        if idy:
            y0, y1 = top, top + S
        else:
            y0, y1 = int(ky[0, 0]), int((ky[:, 0] + ky[:, 1]).max())
        descs[i] = _lib.ResizeDesc(src=d_im.data_ptr(), H=H, W=W, pad_x=px, pad_y=py, y0=y0, y1=y1, top=top, left=left, identity_x=idx,
                                   identity_y=idy, ksize_x=ksx, ksize_y=ksy, kx_off=t_off, ky_off=t_off + kx.size, tmp_off=tmp_off)
        tables += [kx.reshape(-1), ky.reshape(-1)]
        t_off += kx.size + ky.size
        tmp_off += (y1 - y0) * S * 3
        max_rows = max(max_rows, y1 - y0)
    tab = torch.from_numpy(np.concatenate(tables)).to(dev, non_blocking=True)
    dsc = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev, non_blocking=True)
    bg = (C.c_uint8 * 3)(*[int(m * 255) for m in image_mean])      # mm_utils.py:172: tuple(int(x*255) for x in image_mean)
    out = torch.empty(B, 3, S, S, dtype=torch.uint8, device=dev)
    ws = ops.workspace(dev, max(tmp_off, 256), "preprocess")
    with torch.cuda.device(dev):
        st = _lib.load().setok_preprocess_u8(dsc.data_ptr(), B, max_rows, tab.data_ptr(), S, C.addressof(bg), out.data_ptr(), ws.data_ptr(),
                                             ws.numel(), tmp_off, ops._stream(dev))
    _lib.check(st, "setok_preprocess_u8")
    stream = torch.cuda.current_stream(dev)
    for t in keep + [tab, dsc]:
        t.record_stream(stream)
    return out


def process_images(images: Sequence, image_processor, model_cfg=None, device: Optional[torch.device] = None) -> torch.Tensor:
    """Device-side `process_images` (src/mm_utils.py:166-182).  `image_aspect_ratio == 'pad'` pads to a square of the mean
    colour first; the default resizes the shortest edge and center-crops; 'anyres' (multi-crop grids) is not on the tokenizer's
    path and raises.  Returns uint8 (B, 3, S, S) on the device: feed it to `SetokTokenizer` / `CLIPVisionTower`, which apply
    rescale + normalize themselves (the float32 tensor the reference returns is `normalize(out)`)."""
    aspect = getattr(model_cfg, "image_aspect_ratio", None) if model_cfg is not None else None
    if aspect == "anyres":
        raise SetokError("image_aspect_ratio='anyres' is not supported by the device preprocessing path")
    def field(obj, key):                                 # dict (transformers 4.x) or SizeDict (5.x)
        if obj is None:
            return None
        return obj.get(key) if isinstance(obj, dict) else getattr(obj, key, None)
    crop, size = getattr(image_processor, "crop_size", None), getattr(image_processor, "size", None)
    S = int(field(crop, "height") or field(size, "shortest_edge") or 0)
    if S <= 0 or (field(size, "shortest_edge") or S) != S or (field(crop, "width") or S) != S:
        raise SetokError("device preprocessing needs a square crop equal to the processor's shortest_edge")
    mean = tuple(getattr(image_processor, "image_mean", None) or (0.48145466, 0.4578275, 0.40821073))
    return preprocess_images(images, S, pad=(aspect == "pad"), image_mean=mean, device=device)

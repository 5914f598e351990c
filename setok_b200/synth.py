"""Synthetic inputs of the benchmark configurations: "Mondrian" images for the end-to-end
path and mixture-of-Gaussians features for the clustering head in isolation.  Pure tensor plumbing."""
from __future__ import annotations

import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def mondrian_images(B: int, size: int, seed: int, device="cpu", g_min: int = 8, g_max: int = 128, dtype=torch.float32) -> torch.Tensor:
    """B images of `size`^2: G_b ~ U{g_min..g_max} random axis-aligned constant-colour rectangles painted over a
    base colour, plus N(0, 0.02) noise, CLIP-normalised.  Deterministic in (seed, B, size) on a given device type."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    img = torch.rand(B, 3, 1, 1, generator=gen, device=dev).expand(B, 3, size, size).clone()
    G = torch.randint(g_min, g_max + 1, (B,), generator=gen, device=dev)
    ys = torch.arange(size, device=dev)[None, :, None]
    xs = torch.arange(size, device=dev)[None, None, :]
    for r in range(g_max):
        live = (r < G)[:, None, None]
        y0 = torch.randint(0, size, (B, 1, 1), generator=gen, device=dev)
        x0 = torch.randint(0, size, (B, 1, 1), generator=gen, device=dev)
        hh = torch.randint(size // 16, size // 2, (B, 1, 1), generator=gen, device=dev)
        ww = torch.randint(size // 16, size // 2, (B, 1, 1), generator=gen, device=dev)
        col = torch.rand(B, 3, 1, 1, generator=gen, device=dev)
        m = (live & (ys >= y0) & (ys < y0 + hh) & (xs >= x0) & (xs < x0 + ww))[:, None]
        img = torch.where(m, col, img)
    img = img + 0.02 * torch.randn(B, 3, size, size, generator=gen, device=dev)
    mean = torch.tensor(CLIP_MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, device=dev).view(1, 3, 1, 1)
    return ((img - mean) / std).to(dtype).contiguous()


def mog_features(B: int, N: int, C: int, seed: int, device="cpu", g_min: int = 8, g_max: int = 128, sigma: float = 0.05) -> torch.Tensor:
    """Feature-injected input at the ViT-output boundary: per image G_b ~ U{g_min..g_max} centres ~ N(0, I_C), labels
    uniform, noise sigma.  Gives K ~ G."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    centres = torch.randn(B, g_max, C, generator=gen, device=dev)
    G = torch.randint(g_min, g_max + 1, (B,), generator=gen, device=dev)
    lab = (torch.rand(B, N, generator=gen, device=dev) * G[:, None]).long()
    x = torch.gather(centres, 1, lab[..., None].expand(B, N, C))
    return x + sigma * torch.randn(B, N, C, generator=gen, device=dev)


def mondrian_u8(B: int, size: int, seed: int, g_min: int = 8, g_max: int = 128) -> torch.Tensor:
    """The same "Mondrian" images as raw uint8 pixels (B, 3, size, size) on the host: what an image decoder hands to the
    processor.  `normalize_u8` gives the float32 tensor CLIPImageProcessor.rescale + normalize makes of them."""
    x = mondrian_images(B, size, seed, "cpu", g_min, g_max)
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    return (x * std + mean).clamp_(0.0, 1.0).mul_(255.0).round_().to(torch.uint8).contiguous()


def normalize_u8(u8: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """float32(float64(u8) / 255) then (x - mean) / std in float32: transformers' CLIPImageProcessor arithmetic, the same
    values `CLIPVisionTower.forward(uint8)` computes on the device."""
    lut = (torch.arange(256, dtype=torch.float64) * (1.0 / 255.0)).to(torch.float32)
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return ((lut[u8.long()] - mean) / std).to(dtype).contiguous()

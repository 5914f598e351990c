"""Plugin factories with the reference's names and behaviour (the drop-in seam):
`build_vision_tower` <- src/model/multimodal_encoder/builder.py:6-22, `encode_images` <-
src/model/setokim_arch.py:206-211."""
from __future__ import annotations

from dataclasses import asdict, is_dataclass

from .projector import build_vision_projector  # noqa: F401  (re-export)
from .tokenizer import SetokTokenizer


def build_vision_tower(vision_tower_cfg, **kwargs):
    vision_tower = getattr(vision_tower_cfg, "vision_tokenizer", getattr(vision_tower_cfg, "vision_tower", None))
    if isinstance(vision_tower_cfg, dict):
        vision_tower = vision_tower_cfg.get("vision_tokenizer", vision_tower_cfg.get("vision_tower"))
    elif is_dataclass(vision_tower_cfg):
        vision_tower_cfg = asdict(vision_tower_cfg)
    else:
        vision_tower_cfg = dict(vars(vision_tower_cfg))
    if vision_tower is not None and "siglip" in vision_tower:      # builder.py:19 — the only accepted family name
        return SetokTokenizer(**vision_tower_cfg, **kwargs)
    raise ValueError(f"Unknown vision tower: {vision_tower}")


def encode_images(vision_tower, mm_in_projector, images, **tower_kwargs):
    """SetokimMetaForCausalLM.encode_images (setokim_arch.py:206-211): unpack the tokenizer's 3-tuple and apply
    mm_in_projector.  Returns a RaggedTokens: `out[i]` is image i's (K_i, H) rows."""
    image_features, _, _ = vision_tower(images, **tower_kwargs)
    return mm_in_projector(image_features)

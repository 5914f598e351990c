"""setok_b200 — B200-native SeTok vision tokenizer hot path.

Importing the package does not need a GPU; running anything does, and needs libsetok_b200.so
(`python -m setok_b200.build`).  There is no CPU fallback."""
from ._lib import SetokError  # noqa: F401
from .builder import build_vision_projector, build_vision_tower, encode_images  # noqa: F401
from .ragged import RaggedTokens  # noqa: F401
from .tokenizer import CLIPVisionTower, SetokTokenizer  # noqa: F401
from .detokenizer import SetokDeTokenizer  # noqa: F401
from .splice import prepare_inputs_labels_for_multimodal  # noqa: F401
from .preprocess import preprocess_images, process_images  # noqa: F401

__all__ = ["SetokTokenizer", "SetokDeTokenizer", "CLIPVisionTower", "RaggedTokens", "build_vision_tower", "build_vision_projector",
           "encode_images", "prepare_inputs_labels_for_multimodal", "process_images", "preprocess_images", "SetokError"]

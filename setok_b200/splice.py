"""Device-side splice of ``SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal`` (src/model/setokim_arch.py:213-354):
text embeddings + the ragged image-token batch -> padded ``inputs_embeds`` / ``labels`` / ``attention_mask`` / ``position_ids``.

Same call shape and return tuple as the reference method, with the two things it reaches for through ``self`` passed in:
``image_features`` (what ``encode_images`` returned: a ``RaggedTokens`` or a list of (K_i, H) tensors) and ``embed_tokens``
(the LLM's embedding module or its weight).  One host read of the batch's maximum length sizes the outputs (the reference
synchronises once per sample); the work itself is four launches (``setok_splice``).  CUDA only, no fallback."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from . import _lib, ops
from ._lib import SetokError
from .ragged import RaggedTokens

IGNORE_INDEX, IMAGE_TOKEN_INDEX, TARGET_TOKEN_INDEX = -100, -200, -300


def prepare_inputs_labels_for_multimodal(input_ids: torch.Tensor, position_ids: Optional[torch.Tensor], attention_mask: Optional[torch.Tensor],
                                         past_key_values, labels: Optional[torch.Tensor],
                                         image_features: Union[RaggedTokens, Sequence[torch.Tensor]], embed_tokens,
                                         tokenizer_model_max_length: Optional[int] = None, tokenizer_padding_side: str = "right"):
    """Returns ``(None, position_ids, attention_mask, past_key_values, new_input_embeds, new_labels)`` (setokim_arch.py:354)."""
    weight = embed_tokens.weight if hasattr(embed_tokens, "weight") else embed_tokens
    dev = weight.device
    if dev.type != "cuda":
        raise SetokError("setok_b200 kernels need CUDA tensors; there is no CPU fallback")
    if isinstance(image_features, RaggedTokens):
        rows, offsets = image_features.data, image_features.offsets
    else:
        feats = list(image_features)
        counts = torch.tensor([0] + [int(f.shape[0]) for f in feats], dtype=torch.int32)
        offsets = torch.cumsum(counts, 0).to(dtype=torch.int32, device=dev)
        rows = torch.cat(feats, 0) if feats else weight.new_zeros(0, weight.shape[1])
    n_images = int(offsets.numel()) - 1
    dtype = weight.dtype
    if dtype not in (torch.float32, torch.bfloat16):
        raise SetokError(f"unsupported embedding dtype {dtype} (float32 or bfloat16 expected)")
    weight = weight.detach().contiguous()
    offsets = offsets.to(device=dev, dtype=torch.int32).contiguous()
    ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
    B, L = ids.shape
    V, H = weight.shape
    mask8 = None if attention_mask is None else attention_mask.to(dev).bool().to(torch.uint8).contiguous()
    lab = None if labels is None else labels.to(device=dev, dtype=torch.int64).contiguous()
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    max_len = torch.empty(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    limit = int(tokenizer_model_max_length) if tokenizer_model_max_length is not None else 2 ** 30
    # phase 1 needs the three plan slices only (the workspace query for out_cap = 1)
    plan_bytes = lib.setok_splice_workspace_bytes(B, L, 1)
    ws_plan = ops.workspace(dev, plan_bytes, "splice_plan")
    with torch.cuda.device(dev):
        st = lib.setok_splice_plan(ids.data_ptr(), ops._p(mask8), B, L, offsets.data_ptr(), n_images, int(tokenizer_model_max_length or 0), limit,
                                   lens.data_ptr(), max_len.data_ptr(), ws_plan.data_ptr(), ws_plan.numel(), ops._stream(dev))
    _lib.check(st, "setok_splice_plan")
    # the one host read: the reference's max(x.shape[0] ...) (:313), together with the live row count of the ragged batch
    T, live = (int(v) for v in torch.cat([max_len, offsets[-1:]]).tolist())
    T = max(T, 1)
    # only the live rows are cast / copied (a tokenizer-produced RaggedTokens has B*N rows of capacity)
    rows = rows[:live].to(device=dev, dtype=dtype).contiguous()
    if rows.shape[0] == 0:
        rows = weight.new_zeros(1, weight.shape[1])
    # phase 2: exactly-sized workspace (4 bytes per output column); the plan is carried over in its first slices
    ws = ops.workspace(dev, lib.setok_splice_workspace_bytes(B, L, T), "splice")
    ws[:plan_bytes].copy_(ws_plan[:plan_bytes])
    embeds = torch.empty(B, T, H, dtype=dtype, device=dev)
    labels_out = torch.empty(B, T, dtype=torch.int64, device=dev)
    mask_out = torch.empty(B, T, dtype=torch.uint8, device=dev)
    pos_out = torch.empty(B, T, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = lib.setok_splice_fill(ids.data_ptr(), ops._p(mask8), ops._p(lab), B, L, weight.data_ptr(), ops._dt(weight), V, H, rows.data_ptr(),
                                   offsets.data_ptr(), n_images, int(tokenizer_padding_side == "left"), T, lens.data_ptr(), max_len.data_ptr(),
                                   embeds.data_ptr(), labels_out.data_ptr(), mask_out.data_ptr(), pos_out.data_ptr(), ws.data_ptr(), ws.numel(),
                                   ops._stream(dev))
    _lib.check(st, "setok_splice_fill")
    new_input_embeds = embeds[:, :T]
    new_labels = None if labels is None else labels_out[:, :T]
    new_mask = None if attention_mask is None else mask_out[:, :T].to(attention_mask.dtype)
    new_pos = None if position_ids is None else pos_out[:, :T].to(position_ids.dtype)
    return None, new_pos, new_mask, past_key_values, new_input_embeds, new_labels

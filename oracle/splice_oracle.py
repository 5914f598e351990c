"""TEST INFRASTRUCTURE ONLY — CPU restatement of the splice in the reference's
`SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal` (`/root/reference/src/model/setokim_arch.py:241-354`).
Pure index arithmetic in plain Python/torch loops.  Only `tests/` may import it.

Pinned: `oracle/make_golden.py:golden_splice` executes the reference's own method (the unmodified function object of
`setokim_arch.py`, bound to a stub `self` that supplies `get_vision_tower`, `encode_images`, `get_model().embed_tokens`,
`config` and `device`) and commits inputs + outputs to `tests/golden/splice.npz`."""
from __future__ import annotations

from typing import List, Optional

import torch

IGNORE_INDEX, IMAGE_TOKEN_INDEX, TARGET_TOKEN_INDEX = -100, -200, -300


def splice(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], labels: Optional[torch.Tensor], embed: torch.Tensor,
           image_features: List[torch.Tensor], max_length: Optional[int] = None, padding_side: str = "right"):
    """Returns (inputs_embeds (B, T, H), labels (B, T) or None, attention_mask bool (B, T), position_ids (B, T))."""
    B, L = input_ids.shape
    mask = torch.ones_like(input_ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()       # :245-248
    lab_in = torch.full_like(input_ids, IGNORE_INDEX) if labels is None else labels                                 # :251-252
    seqs, labs = [], []
    cur = 0
    for b in range(B):
        ids = input_ids[b][mask[b]]                                                                                 # :256
        lb = lab_in[b][mask[b]]
        e_rows, l_rows = [], []
        n_img = int((ids == IMAGE_TOKEN_INDEX).sum())
        if n_img == 0:                                                                                              # :262-269
            e_rows = [embed[ids]]
            l_rows = [lb]
            cur += 1
        else:
            for t in range(ids.numel()):                                                                            # :271-297
                if int(ids[t]) == IMAGE_TOKEN_INDEX:
                    f = image_features[cur]
                    cur += 1
                    e_rows.append(f.to(embed.dtype))
                    l_rows.append(torch.full((f.shape[0],), IGNORE_INDEX, dtype=lb.dtype))
                else:
                    e_rows.append(embed[ids[t]][None])
                    l_rows.append(lb[t][None])
        e = torch.cat(e_rows, 0) if e_rows else embed.new_zeros(0, embed.shape[1])
        l_ = torch.cat(l_rows, 0) if l_rows else lb.new_zeros(0)
        if max_length is not None:                                                                                  # :307-310
            e, l_ = e[:max_length], l_[:max_length]
        seqs.append(e)
        labs.append(l_)
    T = max(s.shape[0] for s in seqs)                                                                               # :313
    H = embed.shape[1]
    out = embed.new_zeros(B, T, H)
    lab = torch.full((B, T), IGNORE_INDEX, dtype=lab_in.dtype)
    am = torch.zeros(B, T, dtype=torch.bool)
    pos = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):                                                                                              # :320-341
        n = seqs[b].shape[0]
        if n == 0:
            continue
        sl = slice(T - n, T) if padding_side == "left" else slice(0, n)
        out[b, sl] = seqs[b]
        lab[b, sl] = labs[b]
        am[b, sl] = True
        pos[b, sl] = torch.arange(n)
    if labels is None:
        lab = None
    else:
        lab[lab == TARGET_TOKEN_INDEX] = IGNORE_INDEX                                                               # :345
    return out, lab, am, pos

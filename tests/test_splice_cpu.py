"""CPU: the splice oracle against the golden vectors written by the reference's own
`prepare_inputs_labels_for_multimodal` (setokim_arch.py:213-354) executed on a stub self."""
import numpy as np
import torch

from conftest import load_golden
from oracle import splice_oracle as S

T = lambda a: torch.from_numpy(np.asarray(a))


def test_splice_oracle_matches_reference_function():
    g = load_golden("splice")
    offs = [int(v) for v in g["offsets"]]
    feats = [T(g["feats"])[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    for name in g["names"]:
        name = str(name)
        left, maxlen, wl, wm = [int(v) for v in g[name + "/cfg"]]
        e, l, m, p = S.splice(T(g["input_ids"]), T(g["attention_mask"]) if wm else None, T(g["labels"]) if wl else None, T(g["embed"]), feats,
                              maxlen or None, "left" if left else "right")
        assert torch.equal(e, T(g[name + "/embeds"])), name
        assert torch.equal(p, T(g[name + "/pos"])), name
        if wl:
            assert torch.equal(l, T(g[name + "/labels"])), name
        else:
            assert l is None and g[name + "/labels"].size == 0
        if wm:
            assert torch.equal(m, T(g[name + "/mask"]).bool()), name

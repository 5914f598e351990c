"""GPU parity of the device-side splice (`setok_splice`) through the C ABI: bit-exact against the golden
vectors of the reference's own `prepare_inputs_labels_for_multimodal` and against the CPU oracle on random ragged batches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from conftest import load_golden  # noqa: E402
from oracle import splice_oracle as S  # noqa: E402
from setok_b200 import RaggedTokens, prepare_inputs_labels_for_multimodal  # noqa: E402

DEV = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.asarray(a))


def test_splice_golden_bit_exact():
    g = load_golden("splice")
    offs = [int(v) for v in g["offsets"]]
    feats = [T(g["feats"])[offs[i]:offs[i + 1]].to(DEV) for i in range(len(offs) - 1)]
    emb = torch.nn.Embedding(int(g["V"]), int(g["H"])).to(DEV)
    emb.weight.data = T(g["embed"]).to(DEV)
    ids, am, labels = T(g["input_ids"]).to(DEV), T(g["attention_mask"]).to(DEV), T(g["labels"]).to(DEV)
    for name in g["names"]:
        name = str(name)
        left, maxlen, wl, wm = [int(v) for v in g[name + "/cfg"]]
        pos_in = torch.arange(ids.shape[1], device=DEV)[None].expand_as(ids)
        for imgs in (feats, RaggedTokens(T(g["feats"]).to(DEV), T(g["offsets"]).to(DEV))):       # list of tensors or the ragged container
            r = prepare_inputs_labels_for_multimodal(ids, pos_in, am if wm else None, "pkv", labels if wl else None, imgs, emb,
                                                     tokenizer_model_max_length=maxlen or None, tokenizer_padding_side="left" if left else "right")
            assert r[0] is None and r[3] == "pkv"
            assert torch.equal(r[4].cpu(), T(g[name + "/embeds"])), name
            assert torch.equal(r[1].cpu(), T(g[name + "/pos"])), name
            if wl:
                assert torch.equal(r[5].cpu(), T(g[name + "/labels"])), name
            else:
                assert r[5] is None
            if wm:
                assert r[2].dtype == am.dtype and torch.equal(r[2].cpu(), T(g[name + "/mask"])), name
            else:
                assert r[2] is None
        assert prepare_inputs_labels_for_multimodal(ids, None, am, None, labels, feats, emb)[1] is None     # position_ids None in -> None out (:351-352)


@pytest.mark.parametrize("dtype,side,maxlen", [(torch.float32, "right", None), (torch.bfloat16, "left", None), (torch.bfloat16, "right", 300)])
def test_splice_vs_oracle_random(dtype, side, maxlen):
    """BASELINE config 4's consumer shape, scaled down in width: 24 samples, 0..3 images each with K in 1..48, ragged text."""
    gen = torch.Generator().manual_seed(61)
    B, L, V, H = 24, 256, 1000, 128
    emb = torch.randn(V, H, generator=gen).to(dtype)
    ids = torch.randint(0, V, (B, L), generator=gen)
    am = torch.ones(B, L, dtype=torch.bool)
    n_img_total = 0
    for b in range(B):
        n = int(torch.randint(40, L + 1, (1,), generator=gen))
        am[b, n:] = False
        k = int(torch.randint(0, 4, (1,), generator=gen))
        for p_ in torch.randperm(n, generator=gen)[:k].tolist():
            ids[b, p_] = -200
        n_img_total += max(k, 1)
    K = torch.randint(1, 49, (n_img_total,), generator=gen).tolist()
    feats = [torch.randn(k, H, generator=gen).to(dtype) for k in K]
    labels = ids.clone()
    labels[labels == -200] = -100
    labels[:, 3] = -300
    e, l, m, p = S.splice(ids, am, labels, emb, feats, maxlen, side)
    r = prepare_inputs_labels_for_multimodal(ids.to(DEV), torch.zeros(B, L, dtype=torch.long, device=DEV), am.to(DEV), None, labels.to(DEV),
                                             [f.to(DEV) for f in feats], emb.to(DEV), tokenizer_model_max_length=maxlen, tokenizer_padding_side=side)
    assert torch.equal(r[4].cpu(), e) and torch.equal(r[5].cpu(), l) and torch.equal(r[2].cpu(), m) and torch.equal(r[1].cpu(), p)


def test_splice_properties_config4_size():
    """64 samples x 512 text tokens x one image of K in 8..128 rows, H = 4096 bf16 (Vicuna-7B width): every output row is
    either an embedding-table row, an image row or zero; lengths = text + K; labels IGNORE over image rows."""
    gen = torch.Generator(device=DEV).manual_seed(62)
    B, L, V, H = 64, 512, 32000, 4096
    emb = torch.randn(V, H, device=DEV, generator=gen).to(torch.bfloat16)
    ids = torch.randint(0, V, (B, L), device=DEV, generator=gen)
    where = torch.randint(0, L, (B,), device=DEV, generator=gen)
    ids[torch.arange(B, device=DEV), where] = -200
    K = torch.randint(8, 129, (B,), device=DEV, generator=gen)
    offsets = torch.zeros(B + 1, dtype=torch.int32, device=DEV)
    offsets[1:] = torch.cumsum(K, 0)
    rows = torch.randn(int(offsets[-1]), H, device=DEV, generator=gen).to(torch.bfloat16)
    labels = ids.clone()
    r = prepare_inputs_labels_for_multimodal(ids, None, torch.ones(B, L, dtype=torch.bool, device=DEV), None, labels, RaggedTokens(rows, offsets), emb)
    embeds, lab, mask = r[4], r[5], r[2]
    lens = mask.sum(1)
    assert torch.equal(lens.cpu(), (L - 1 + K).cpu()) and embeds.shape == (B, int(lens.max()), H)
    for b in (0, 31, 63):
        w, k = int(where[b]), int(K[b])
        assert torch.equal(embeds[b, :w], emb[ids[b, :w]])
        assert torch.equal(embeds[b, w:w + k], rows[int(offsets[b]):int(offsets[b + 1])])
        assert torch.equal(embeds[b, w + k:w + k + (L - 1 - w)], emb[ids[b, w + 1:]])
        assert bool((embeds[b, int(lens[b]):] == 0).all())
        assert bool((lab[b, w:w + k] == -100).all()) and torch.equal(lab[b, :w], ids[b, :w])

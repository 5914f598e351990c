"""CPU checks of the detokenizer row: the oracle restatement against the golden vectors produced by
the reference's own BertEmbeddings / BertEncoder / PositionalEncoding2D classes (+ HF ViTLayer for timm's Block), and
the host mirror's parameter surface.  No GPU, no compute through the C ABI."""
import numpy as np
import torch

from conftest import load_golden
from oracle import detok_oracle as D

T = lambda a: torch.from_numpy(np.asarray(a))


def _golden():
    g = load_golden("detok")
    p = {str(k): T(g["param/" + str(k)]) for k in g["keys"]}
    names = ("token_dim", "hidden", "q_heads", "q_inter", "q_layers", "cross_freq", "grid", "dec_dim", "dec_depth", "dec_mlp", "dec_heads")
    d = {n: int(v) for n, v in zip(names, g["dims"])}
    return g, p, d


def test_oracle_matches_reference_modules_bit_exact():
    g, p, d = _golden()
    out, im = D.detok_forward(p, T(g["x"]), T(g["mask"]), q_heads=d["q_heads"], q_layers=d["q_layers"], cross_freq=d["cross_freq"], grid=d["grid"],
                              dec_heads=d["dec_heads"], dec_depth=d["dec_depth"], hidden=d["hidden"], return_intermediates=True)
    assert torch.equal(D.decoder_pos_table(d["hidden"], d["grid"], d["dec_dim"]), T(g["pos"]))
    torch.testing.assert_close(im["qformer"], T(g["qformer"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out, T(g["out"]), rtol=1e-5, atol=1e-6)


def test_padding_is_inert_and_ragged_equals_padded():
    """Extra padded (masked) rows must not change the result (module.py:849 additive mask), and packing/unpacking round-trips."""
    g, p, d = _golden()
    x, m = T(g["x"]), T(g["mask"])
    kw = dict(q_heads=d["q_heads"], q_layers=d["q_layers"], cross_freq=d["cross_freq"], grid=d["grid"], dec_heads=d["dec_heads"],
              dec_depth=d["dec_depth"], hidden=d["hidden"])
    ref = D.detok_forward(p, x, m, **kw)
    x2 = torch.cat([x, torch.randn(x.shape[0], 3, x.shape[2])], 1)
    m2 = torch.cat([m, torch.zeros(m.shape[0], 3)], 1)
    torch.testing.assert_close(D.detok_forward(p, x2, m2, **kw), ref, rtol=1e-5, atol=1e-6)
    xp, mp = D.pad_ragged(T(g["tokens"]), [int(v) for v in g["offsets"]])
    assert torch.equal(xp, x) and torch.equal(mp, m)


def test_host_mirror_state_dict_surface():
    """SetokDeTokenizer exposes the reference's parameter names (detokenizer.py:41-54, 72-96; module.py Bert* classes)."""
    from setok_b200 import SetokDeTokenizer
    g, p, d = _golden()
    det = SetokDeTokenizer(token_feat_dim=d["token_dim"], hidden_dim=d["hidden"], patch_size=4, image_size=4 * d["grid"],
                           decoder_embed_dim=d["dec_dim"], decoder_nheads=d["dec_heads"], decoder_depth=d["dec_depth"],
                           mlp_ratio=d["dec_mlp"] / d["dec_dim"], num_hidden_layers=d["q_layers"], cross_attention_freq=d["cross_freq"],
                           mapper_num_attention_heads=d["q_heads"], mapper_intermediate_size=d["q_inter"])
    keys = set(det.state_dict().keys())
    assert set(p.keys()) <= keys, sorted(set(p.keys()) - keys)
    assert keys - set(p.keys()) <= {"position_embedding.inv_freq", "mapper.embeddings.position_ids"}
    missing, unexpected = det.load_state_dict(p, strict=False)
    assert not unexpected
    assert det.num_mask_token == d["grid"] ** 2 and det.mask_tokens.shape == (1, d["grid"] ** 2, d["hidden"])
    # layer i carries cross-attention iff i % cross_attention_freq == 0 (module.py:484-493)
    assert [L.has_cross_attention for L in det.mapper.encoder.layer] == [i % d["cross_freq"] == 0 for i in range(d["q_layers"])]
    # defaults are BERT-base's: heads of 64, intermediate 4x
    big = SetokDeTokenizer(token_feat_dim=64, hidden_dim=768, image_size=28, decoder_embed_dim=768, decoder_nheads=12, decoder_depth=1, num_hidden_layers=2)
    assert big.mapper_heads == 12 and big.mapper_inter == 3072
    import pytest
    with pytest.raises(ValueError):                       # D2: the reference's x + pos_emb cannot broadcast
        SetokDeTokenizer(token_feat_dim=64, hidden_dim=768, image_size=28, decoder_embed_dim=4096, decoder_depth=1, num_hidden_layers=1)
    from setok_b200 import SetokError
    with pytest.raises(SetokError):                       # no CPU path
        det(torch.zeros(1, 2, d["token_dim"]), torch.ones(1, 2))

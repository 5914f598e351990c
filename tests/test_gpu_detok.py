"""GPU parity of the detokenizer consumer (BASELINE config 3's decoder half) through the C ABI
(`setok_detok_forward`) against the golden vectors of the reference's own modules and against the CPU oracle.

Tolerance: the stream is bf16 with fp32 accumulation (as in the tower); against the fp32 oracle we assert relative
Frobenius error <= 1e-2 and normalised max error <= 3e-2 on the LayerNorm-ed output, and that the error does not exceed
1.5x the error torch's own bf16 evaluation of the same formula makes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from conftest import load_golden  # noqa: E402
from oracle import detok_oracle as D  # noqa: E402
from setok_b200 import RaggedTokens, SetokDeTokenizer  # noqa: E402

DEV = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.asarray(a))
NAMES = ("token_dim", "hidden", "q_heads", "q_inter", "q_layers", "cross_freq", "grid", "dec_dim", "dec_depth", "dec_mlp", "dec_heads")


def _err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6)), float((got - ref).norm() / ref.norm().clamp_min(1e-6))


def _build(d, p, patch=4):
    det = SetokDeTokenizer(token_feat_dim=d["token_dim"], hidden_dim=d["hidden"], patch_size=patch, image_size=patch * d["grid"],
                           decoder_embed_dim=d["dec_dim"], decoder_nheads=d["dec_heads"], decoder_depth=d["dec_depth"],
                           mlp_ratio=d["dec_mlp"] / d["dec_dim"], num_hidden_layers=d["q_layers"], cross_attention_freq=d["cross_freq"],
                           mapper_num_attention_heads=d["q_heads"], mapper_intermediate_size=d["q_inter"])
    missing, unexpected = det.load_state_dict(p, strict=False)
    assert not unexpected
    return det.to(DEV)


def _oracle(p, d, x, m, dtype=torch.float32):
    pp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in p.items()}
    return D.detok_forward(pp, x.to(dtype), m, q_heads=d["q_heads"], q_layers=d["q_layers"], cross_freq=d["cross_freq"], grid=d["grid"],
                           dec_heads=d["dec_heads"], dec_depth=d["dec_depth"], hidden=d["hidden"])


def test_detok_golden():
    g = load_golden("detok")
    p = {str(k): T(g["param/" + str(k)]) for k in g["keys"]}
    d = {n: int(v) for n, v in zip(NAMES, g["dims"])}
    det = _build(d, p)
    ref = T(g["out"])
    out = det(T(g["x"]).to(DEV), T(g["mask"]).to(DEV))
    assert out.shape == ref.shape and out.dtype == torch.float32
    emax, efro = _err(out, ref)
    assert efro < 1e-2 and emax < 3e-2, (emax, efro)
    # ragged input (the tokenizer's own container) = padded input, bit for bit; extra masked padding is inert
    rt = RaggedTokens(T(g["tokens"]).to(DEV), T(g["offsets"]).to(DEV))
    assert torch.equal(det(rt), out)
    x2 = torch.cat([T(g["x"]), torch.randn(3, 4, d["token_dim"])], 1)
    m2 = torch.cat([T(g["mask"]), torch.zeros(3, 4)], 1)
    assert torch.equal(det(x2.to(DEV), m2.to(DEV)), out)
    # masks need not be prefixes: scatter image 1's two tokens to the end of its row
    x3, m3 = T(g["x"]).clone(), T(g["mask"]).clone()
    x3[1, -2:] = x3[1, :2]; m3[1, -2:] = 1; m3[1, :2] = 0
    assert torch.equal(det(x3.to(DEV), m3.to(DEV)), out)


@pytest.mark.parametrize("hidden,q_heads,dec_dim,dec_heads,grid,B", [(128, 2, 128, 2, 8, 4), (768, 12, 768, 12, 24, 2), (256, 4, 192, 4, 6, 3)])
def test_detok_vs_oracle(hidden, q_heads, dec_dim, dec_heads, grid, B):
    """Head dims 64 (tcgen05 attention path; 768/12 at Q = 576 is BASELINE config 3's decoder geometry) and 48
    (CUDA-core path), ragged K_b in 1..40, bf16 token input as the tokenizer emits it in config 3."""
    d = dict(token_dim=64, hidden=hidden, q_heads=q_heads, q_inter=2 * hidden, q_layers=2, cross_freq=2, grid=grid, dec_dim=dec_dim,
             dec_depth=2, dec_mlp=2 * dec_dim, dec_heads=dec_heads)
    p = D.make_detok_params(**{k: v for k, v in d.items() if k != "dec_heads"}, seed=21)
    det = _build(d, p)
    gen = torch.Generator().manual_seed(22)
    K = torch.randint(1, 41, (B,), generator=gen).tolist()
    offsets = [0]
    for k_ in K:
        offsets.append(offsets[-1] + k_)
    tokens = torch.randn(offsets[-1], d["token_dim"], generator=gen).to(torch.bfloat16)
    x, m = D.pad_ragged(tokens.float(), offsets)
    ref = _oracle(p, d, x, m)
    out = det(RaggedTokens(tokens.to(DEV), torch.tensor(offsets, dtype=torch.int32, device=DEV)))
    assert out.dtype == torch.bfloat16 and out.shape == (B, grid * grid, dec_dim)
    emax, efro = _err(out, ref)
    ref_bf16 = _oracle(p, d, x, m, torch.bfloat16)
    bmax, bfro = _err(ref_bf16, ref)
    assert efro < 1e-2 and emax < 3e-2, (emax, efro)
    assert efro <= 1.5 * bfro + 1e-3, (efro, bfro)

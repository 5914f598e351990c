"""CPU: the oracle restatement (oracle/setok_oracle.py) against golden vectors produced by the
reference's own source (oracle/make_golden.py).  Integer outputs bit-exact; floats to 1e-5."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import setok_oracle as O

T = lambda a: torch.from_numpy(np.asarray(a))


def test_posenc_golden():
    g = load_golden("posenc")
    for key in g.files:
        _, h, w, C = key.split("_")
        pe = O.pos_encoding_2d(int(h), int(w), int(C))
        assert torch.equal(pe, T(g[key])), key


def test_dpc_knn_golden_bit_exact():
    g = load_golden("dpc_knn")
    for name in g["names"]:
        name = str(name)
        k, thr, mcn = g[name + "/params"]
        tm = T(g[name + "/token_mask"]) if (name + "/token_mask") in g.files else None
        idx_down, idx_cluster, score = O.dpc_knn(T(g[name + "/x"]), int(k), T(g[name + "/noise"]), float(thr), int(mcn), tm)
        assert torch.equal(idx_down, T(g[name + "/index_down"])), name
        assert torch.equal(idx_cluster, T(g[name + "/idx_cluster"])), name
        assert score.shape == (1, g[name + "/x"].shape[0])
        torch.testing.assert_close(score, T(g[name + "/score"]), rtol=1e-5, atol=1e-7)


def test_dpc_properties_on_goldens():
    """Every centre owns its own label; labels in [0,K); K>=1 (tokenizer.py:117-119)."""
    g = load_golden("dpc_knn")
    for name in g["names"]:
        name = str(name)
        idx_down, idx_cluster = T(g[name + "/index_down"]), T(g[name + "/idx_cluster"])
        K = idx_down.numel()
        assert K >= 1 and int(idx_cluster.min()) >= 0 and int(idx_cluster.max()) < K
        assert torch.equal(idx_cluster[idx_down], torch.arange(K))
        assert torch.equal(torch.sort(idx_down).values, idx_down)


def _head_params(g, prefix, keys):
    return {str(k): T(g[prefix + str(k)]) for k in g[keys]}


def test_block_and_head_golden():
    g = load_golden("head")
    p = _head_params(g, "sd/", "sd_keys")
    C, Ctok, Fd, N, k, mcn = [int(v) for v in g["cfg"]]
    y = O.block(T(g["block_in"]), p, "inner_encoder.", 2, 2)
    torch.testing.assert_close(y, T(g["block_out"]), rtol=1e-5, atol=1e-5)
    feats = T(g["feats"])
    for b in range(2):
        toks, idx, score, im = O.tokenizer_head(feats[b], T(g[f"img{b}/noise"]), p, min_cluster_num=mcn, threshold=0.5,
                                                k=k, return_intermediates=True)
        assert torch.equal(idx, T(g[f"img{b}/idx_cluster"]))
        assert torch.equal(im["index_down"], T(g[f"img{b}/index_down"]))
        torch.testing.assert_close(score, T(g[f"img{b}/score"]), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(im["group_features"], T(g[f"img{b}/group_features"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(toks, T(g[f"img{b}/tokens"]), rtol=1e-4, atol=1e-5)


def test_tower_and_e2e_golden():
    g = load_golden("tower_e2e")
    C, L, H, P, IMG = [int(v) for v in g["tower_cfg"]]
    tp = {str(k): T(g["tower/" + str(k)]) for k in g["tower_keys"]}
    images = T(g["images"])
    for sl in (-2, -1):
        f = O.tower_features(images, tp, patch=P, heads=H, layers=L, select_layer=sl)
        torch.testing.assert_close(f, T(g[f"feats_sl{sl}"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(O.tower_features(images[:1], tp, patch=P, heads=H, layers=L, select_layer=-1),
                               T(g["feats_list0"]), rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        O.tower_features(images, tp, patch=P, heads=H, layers=L, select_feature="bogus")
    hp = {str(k): T(g["head/" + str(k)]) for k in g["head_keys"]}
    pp = {str(k): T(g["proj/" + str(k)]) for k in g["proj_keys"]}
    Ctok, Hllm, Fd, k, mcn = [int(v) for v in g["e2e_cfg"]]
    thr = float(g["e2e_thr"][0])
    noise = torch.stack([T(g[f"e2e{b}/noise"]) for b in range(2)])
    toks, idxs, scores = O.setok_forward(images, noise, tp, hp, patch=P, heads=H, layers=L, select_layer=-2,
                                         min_cluster_num=mcn, threshold=0.5, k=k, thr=thr)
    for b in range(2):
        assert torch.equal(idxs[b], T(g[f"e2e{b}/idx_cluster"]))
        torch.testing.assert_close(scores[b], T(g[f"e2e{b}/score"]), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(toks[b], T(g[f"e2e{b}/tokens"]), rtol=1e-3, atol=1e-4)
        torch.testing.assert_close(O.projector(toks[b], pp, "mlp2x_gelu"), T(g[f"e2e{b}/projected"]), rtol=1e-3, atol=1e-4)


def test_projector_golden():
    g = load_golden("projectors")
    x = T(g["x"])
    for t in ("linear", "mlp2x_gelu", "mlp3x_gelu", "mlp2x_gelu_Norm", "identity"):
        p = {str(k): T(g[f"{t}/{k}"]) for k in g[f"{t}/keys"]}
        torch.testing.assert_close(O.projector(x, p, t), T(g[f"{t}/y"]), rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        O.projector(x, {}, "conv")


def test_known_answers_from_survey():
    """SURVEY.md §8c observed behaviour: iid features never pass 0.5 -> fallback K == min_cluster_num;
    K may be < min_cluster_num when something passes; score has shape (1, N)."""
    x = torch.randn(128, 256, generator=torch.Generator().manual_seed(0))
    idx_down, idx_cluster, score = O.dpc_knn(x, 16, O.tie_noise(128, 1), 0.5, 24)
    assert idx_down.numel() == 24 and score.shape == (1, 128)
    x = O.mog_features(256, 256, 100, 0.05, 3)
    idx_down, _, _ = O.dpc_knn(x, 16, O.tie_noise(256, 2), 0.5, 64)
    assert 1 <= idx_down.numel() < 64

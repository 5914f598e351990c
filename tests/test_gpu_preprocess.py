"""GPU parity of the device image preprocessing (setok_preprocess_u8) against the oracle and the
committed PIL fixture: integer work, so everything is bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from conftest import load_golden  # noqa: E402
from oracle import preprocess_oracle as P  # noqa: E402
from setok_b200 import SetokTokenizer  # noqa: E402
from setok_b200.preprocess import preprocess_images, process_images  # noqa: E402

DEV = torch.device("cuda:0")


def test_golden_cases_bit_exact():
    """Every fixture case (PIL itself, arranged as transformers 4.46.3 + the reference's expand2square arrange it), one launch
    pair per padding mode over images of different sizes."""
    g = load_golden("preprocess")
    by_mode = {}
    for i, (H, W, S, pad) in enumerate(g["cases"]):
        by_mode.setdefault((int(S), bool(pad)), []).append(i)
    for (S, pad), idx in by_mode.items():
        out = preprocess_images([torch.from_numpy(g[f"in{i}"]) for i in idx], S, pad=pad, device=DEV).cpu().numpy()
        for j, i in enumerate(idx):
            assert np.array_equal(out[j], g[f"u8_{i}"]), (i, S, pad)


@pytest.mark.parametrize("S,pad", [(224, False), (224, True), (336, False)])
def test_random_sizes_vs_oracle(S, pad):
    """A ragged batch: landscape / portrait / square, up- and down-scaling, extreme aspect ratios, sizes equal to the target."""
    rng = np.random.default_rng(S + pad)
    sizes = [(300, 400), (500, 375), (100, 80), (640, 640), (S, S), (231, 517), (S, 2 * S), (1000, 60), (61, 1000), (S + 1, S - 1 + 2)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    out = preprocess_images(imgs, S, pad=pad, device=DEV)
    assert out.shape == (len(sizes), 3, S, S) and out.dtype == torch.uint8
    for j, im in enumerate(imgs):
        assert np.array_equal(out[j].cpu().numpy(), P.preprocess_u8(im, S, pad)), (sizes[j], S, pad)


def test_process_images_feeds_the_tokenizer():
    """process_images (mm_utils.py:166-182) on the device, then the tokenizer on the uint8 result: the same tokens as the
    tokenizer on the float32 tensor the reference's host pipeline (oracle) produces."""
    C, L, H, Pp, IMG = 128, 2, 2, 4, 32
    cfg = dict(hidden_size=C, intermediate_size=4 * C, num_hidden_layers=L, num_attention_heads=H, image_size=IMG, patch_size=Pp)
    torch.manual_seed(5)
    tok = SetokTokenizer("siglip-synthetic", hidden_dim=C, token_feat_dim=64, min_cluster_num=8, threshold=0.5, dim_feedforward=256,
                         vision_config=cfg).to(DEV)
    rng = np.random.default_rng(7)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((40, 52), (70, 33), (32, 32), (90, 90))]

    class Cfg:
        image_aspect_ratio = "pad"
    u8 = process_images(imgs, tok.image_processor, Cfg(), device=DEV)
    assert u8.shape == (4, 3, IMG, IMG)
    ref = torch.from_numpy(np.stack([P.rescale_normalize(P.preprocess_u8(im, IMG, True)) for im in imgs]))
    noise = torch.rand(4, (IMG // Pp) ** 2, generator=torch.Generator().manual_seed(1)).to(DEV)
    rt_a, idx_a, sc_a = tok(u8, k=8, noise=noise)
    rt_b, idx_b, sc_b = tok(ref.to(DEV), k=8, noise=noise)
    assert torch.equal(idx_a, idx_b) and torch.equal(sc_a, sc_b) and torch.equal(rt_a.offsets, rt_b.offsets)
    n = int(rt_a.offsets[-1])
    assert torch.equal(rt_a.data[:n], rt_b.data[:n])
    with pytest.raises(Exception):
        class Any:
            image_aspect_ratio = "anyres"
        process_images(imgs, tok.image_processor, Any(), device=DEV)

"""CPU: the preprocessing oracle (oracle/preprocess_oracle.py) against PIL itself and against the committed fixture, and the
product's host-side tap tables (setok_b200/preprocess.py:_taps, Pillow's precompute_coeffs vectorised) against the oracle's
scalar restatement."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import preprocess_oracle as P


@pytest.mark.parametrize("H,W,oh,ow", [(300, 400, 224, 298), (500, 375, 298, 224), (100, 80, 280, 224), (640, 640, 224, 224),
                                       (224, 224, 336, 336), (37, 91, 224, 550), (1200, 900, 448, 336), (224, 300, 224, 300)])
def test_resample_matches_pil(H, W, oh, ow):
    from PIL import Image
    img = np.random.default_rng(H * 1000 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(P.resample_u8(img, oh, ow), ref)


def test_pipeline_matches_golden():
    g = load_golden("preprocess")
    for i, (H, W, S, pad) in enumerate(g["cases"]):
        u8 = P.preprocess_u8(g[f"in{i}"], int(S), bool(pad))
        assert np.array_equal(u8, g[f"u8_{i}"]), (i, H, W, S, pad)
        assert np.array_equal(P.rescale_normalize(u8), g[f"f32_{i}"]), i


def test_resize_output_size_rule():
    assert P.resize_output_size(300, 400, 224) == (224, 298) and P.resize_output_size(500, 375, 224) == (298, 224)
    assert P.resize_output_size(224, 224, 336) == (336, 336) and P.resize_output_size(97, 41, 56) == (132, 56)


@pytest.mark.parametrize("n_in,n_out", [(400, 298), (375, 224), (80, 224), (640, 224), (224, 336), (91, 550), (3000, 224), (224, 224), (41, 56)])
def test_product_tap_tables_equal_oracle(n_in, n_out):
    from setok_b200.preprocess import _taps, resize_output_size
    kk, bounds = P.precompute_coeffs(n_in, n_out)
    for first, count in ((0, n_out), (n_out // 3, n_out - n_out // 3), (max(0, (n_out - 56) // 2), min(56, n_out))):
        tab, ksize = _taps(n_in, n_out, first, count)
        assert ksize == kk.shape[1] and tab.shape == (count, 2 + ksize)
        assert np.array_equal(tab[:, :2], bounds[first:first + count])
        assert np.array_equal(tab[:, 2:], kk[first:first + count])
    assert resize_output_size(97, 41, 56) == P.resize_output_size(97, 41, 56)

"""CPU: the C-ABI shared library builds/loads and exports every symbol include/setok_b200.h declares; argument
validation answers without touching a GPU; host-side containers behave like the reference's outputs."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from setok_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "setok_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(setok_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from setok_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/setok_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in setok_b200/_lib.py"
    assert lib.setok_abi_version() == 3


def test_argument_validation_without_gpu(lib):
    from setok_b200 import _lib
    # null operands / bad shapes are rejected before any CUDA call
    assert lib.setok_gemm_bf16(None, 8, None, 8, None, 8, 0, None, None, 0, 0, 0, 4, 8, 8, None, None) == -1
    assert b"null" in lib.setok_last_error()
    buf = (C.c_char * 64)()
    p = C.addressof(buf) // 16 * 16 + 16
    assert lib.setok_gemm_bf16(p, 8, p, 8, p, 8, 0, None, None, 0, 0, 0, 4, 6, 8, None, None) == -2     # N % 4
    assert lib.setok_layernorm(p, 0, p, 0, p, p, 1e-5, 4, 6, None, None, None) == -2                  # C % 4
    assert lib.setok_attention(p, p, 4, 24, 5, 1.0, None, None, 0, None, None) == -1                  # C % heads
    assert lib.setok_vit_workspace_bytes(None, 4) == 0
    with pytest.raises(_lib.SetokError):
        _lib.check(-2, "probe")


def test_struct_layouts_match_header():
    from setok_b200 import _lib
    assert C.sizeof(_lib.VitLayer) == 14 * 8
    assert C.sizeof(_lib.Attn) == 4 * 8
    assert C.sizeof(_lib.Vit) == 6 * 4 + 4 + 4 + 6 * 8 + 8      # 6 ints, float, pad, 6 pointers, flags + pad
    assert _lib.Vit.flags.offset == 80
    assert _lib.Vit.w_patch.offset == 32
    assert C.sizeof(_lib.Block) == 8 + 4 * 8 + 8 + 4 * 8
    assert C.sizeof(_lib.Head) == 16 + 2 * C.sizeof(_lib.Block) + 16
    assert C.sizeof(_lib.Projector) == 8 + 3 * 8 + 2 * 8


def test_workspace_queries(lib):
    from setok_b200 import _lib
    layers = (_lib.VitLayer * 24)()
    vit = _lib.Vit(image_size=224, patch=14, hidden=1024, heads=16, layers=24, mlp=4096, ln_eps=1e-5, layer=layers)
    n256 = lib.setok_vit_workspace_bytes(C.byref(vit), 256)
    n128 = lib.setok_vit_workspace_bytes(C.byref(vit), 128)
    R = 256 * 257
    expect = 256 * 256 * 640 * 2 + R * 1024 * 4 + R * 1024 * 2 * 3 + R * 3072 * 2 + R * 4096 * 2
    assert n256 >= expect and n256 < expect * 1.01 and abs(n256 - 2 * n128) < 1 << 16
    assert lib.setok_dpc_workspace_bytes(256, 256, 1024) >= 256 * 256 * 256 * 4


def test_no_cpu_fallback():
    import setok_b200
    from setok_b200 import ops
    with pytest.raises(setok_b200.SetokError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    vc = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, image_size=16, patch_size=4)
    tok = setok_b200.SetokTokenizer("siglip-x", hidden_dim=128, token_feat_dim=64, min_cluster_num=4, dim_feedforward=256, vision_config=vc)
    with pytest.raises(setok_b200.SetokError):
        tok(torch.zeros(1, 3, 16, 16))
    with pytest.raises(setok_b200.SetokError):
        setok_b200.build_vision_projector("mlp2x_gelu", 64, 128)(torch.zeros(4, 64))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under setok_b200/ may reference it."""
    pkg = os.path.join(ROOT, "setok_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in src, f

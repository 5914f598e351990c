"""TEST INFRASTRUCTURE ONLY — loads the *actual* reference source from /root/reference.

This file exists so that `oracle/make_golden.py` and the container-only test
`tests/test_oracle_vs_reference.py` can execute the reference's own
`src/model/setok/{utils,module,clip_encoder,tokenizer}.py` unmodified, without running
`src/__init__.py` (which eagerly imports training code whose dependencies are absent).

It cannot travel to the GPU box (/root/reference does not exist there): nothing under
`setok_b200/`, `bench.py` or the `-m gpu` tests imports it.  `available()` says whether the
reference tree is present.

Shim:
  * `timm`, `timm.models`, `timm.models.layers` stub exposing an identity `DropPath`
    (reference: tokenizer.py:7, module.py:7 import it; it is identity in eval mode);
  * three helpers that `module.py:16-21` imports from `transformers.modeling_utils`
    but which moved/vanished in transformers 5.x are aliased back.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SETOK_REFERENCE_ROOT", "/root/reference")
_PKG = "_setok_reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src/model/setok/tokenizer.py"))


def _install_shims() -> None:
    import torch.nn as nn

    if "timm" not in sys.modules:
        class DropPath(nn.Identity):
            def __init__(self, drop_prob: float = 0.0, *a, **k):
                super().__init__()

        def _mk(name):
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
            m.__path__ = []
            return m

        timm = _mk("timm")
        models = _mk("timm.models")
        layers = _mk("timm.models.layers")
        layers.DropPath = DropPath
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers

    import transformers.modeling_utils as mu
    try:
        import transformers.pytorch_utils as pu
    except Exception:  # pragma: no cover
        pu = None
    for name in ("apply_chunking_to_forward", "prune_linear_layer", "find_pruneable_heads_and_indices"):
        if not hasattr(mu, name):
            if pu is not None and hasattr(pu, name):
                setattr(mu, name, getattr(pu, name))
            else:
                def _missing(*a, _n=name, **k):
                    raise NotImplementedError(f"{_n} is not available in this transformers version")
                setattr(mu, name, _missing)


def _load(modname: str, relpath: str):
    full = f"{_PKG}.{modname}"
    if full in sys.modules:
        return sys.modules[full]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with the reference's `tokenizer`, `module`, `utils`, `clip_encoder`
    and `projector_builder` modules, executed from /root/reference unmodified."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_shims()
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src/model/setok")]
        pkg.__spec__ = importlib.machinery.ModuleSpec(_PKG, loader=None, is_package=True)
        sys.modules[_PKG] = pkg
    ns = types.SimpleNamespace()
    ns.utils = _load("utils", "src/model/setok/utils.py")
    ns.module = _load("module", "src/model/setok/module.py")
    ns.clip_encoder = _load("clip_encoder", "src/model/setok/clip_encoder.py")
    ns.tokenizer = _load("tokenizer", "src/model/setok/tokenizer.py")
    ns.projector_builder = _load("projector_builder", "src/model/multimodal_projector/builder.py")
    return ns


def build_reference_tokenizer(ns, hf_tower, *, hidden_dim, token_feat_dim, min_cluster_num, threshold,
                              nheads=2, dim_feedforward=4096, inner_cluster_layers=2,
                              intra_cluster_layers=2, select_layer=-2, select_feature="patch"):
    """Constructs the reference `SetokTokenizer` without touching the network.

    `SetokTokenizer.__init__` (tokenizer.py:50-56) calls `AutoModel.from_pretrained`, which needs
    a checkpoint; we bypass `__init__` via `__new__` and assign exactly the sub-modules the
    constructor would (tokenizer.py:37-48), then attach a `CLIPVisionTower` whose `vision_tower`
    is the given seeded HF model (clip_encoder.py:29-38 minus the download).
    """
    import torch.nn as nn
    T = ns.tokenizer
    tok = T.SetokTokenizer.__new__(T.SetokTokenizer)
    nn.Module.__init__(tok)
    tok.hidden_dim = hidden_dim
    tok.token_feat_dim = token_feat_dim
    Block = ns.module.Block
    tok.inner_encoder = Block(hidden_dim, nheads, dim_feedforward, proj_drop=0.2, attn_drop=0.0, drop_path=0.0,
                              act_layer=nn.GELU, norm_layer=nn.LayerNorm, depth=inner_cluster_layers)
    tok.inter_encoder = Block(hidden_dim, nheads, dim_feedforward, proj_drop=0.2, attn_drop=0.0, drop_path=0.0,
                              act_layer=nn.GELU, norm_layer=nn.LayerNorm, depth=intra_cluster_layers)
    tok.position_embedding = ns.module.PositionalEncoding2D(hidden_dim)
    tok.out = nn.Linear(hidden_dim, token_feat_dim)
    tok.min_cluster_num = min_cluster_num
    tok.threshold = threshold
    tok.initialize_weights()
    tower = ns.clip_encoder.CLIPVisionTower.__new__(ns.clip_encoder.CLIPVisionTower)
    nn.Module.__init__(tower)
    tower.is_loaded = True
    tower.vision_tower_name = "seeded-clip"
    tower.select_layer = select_layer
    tower.select_feature = select_feature
    tower.vision_tower = hf_tower
    if hf_tower is not None:
        hf_tower.requires_grad_(False)
    tok.image_feature_encoder = tower
    tok.eval()
    return tok


def load_reference_splice():
    """The reference's `SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal` function object, from the unmodified
    `src/model/setokim_arch.py`.  The module's imports of sibling builders / the diffusion loss / `src.constants` are
    satisfied by stand-in modules (only the three token constants are real: they are read from `src/constants.py`)."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    pkg = "_setok_reference_model"
    if pkg + ".setokim_arch" in sys.modules:
        return sys.modules[pkg + ".setokim_arch"].SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal

    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    consts = {}
    exec(open(os.path.join(REFERENCE_ROOT, "src/constants.py")).read(), consts)
    stub = lambda *a, **k: None
    mk(pkg)
    mk(pkg + ".multimodal_encoder"); mk(pkg + ".multimodal_encoder.builder", build_vision_tower=stub)
    mk(pkg + ".multimodal_projector"); mk(pkg + ".multimodal_projector.builder", build_vision_projector=stub)
    mk(pkg + ".multimodal_generator"); mk(pkg + ".multimodal_generator.builder", build_vision_generator=stub)
    mk(pkg + ".loss", DiffLoss=object)
    had_src = "src" in sys.modules
    if not had_src:
        mk("src")
    mk("src.constants", **{k: v for k, v in consts.items() if k.isupper()})
    path = os.path.join(REFERENCE_ROOT, "src/model/setokim_arch.py")
    spec = importlib.util.spec_from_file_location(pkg + ".setokim_arch", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[pkg + ".setokim_arch"] = mod
    spec.loader.exec_module(mod)
    return mod.SetokimMetaForCausalLM.prepare_inputs_labels_for_multimodal

"""ctypes binding of libsetok_b200.so (the C ABI declared in include/setok_b200.h).

There is no fallback: if the shared library is missing or a call fails, a `SetokError` is raised.
Build it with `python -m setok_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SETOK_B200_LIB") or os.path.join(HERE, "libsetok_b200.so")   # override: A/B builds under tools/

F32, BF16 = 0, 1
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF = 0, 1, 2

c_void_p, c_int, c_float, c_size_t, c_int64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64


class SetokError(RuntimeError):
    pass


class VitLayer(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("w_qkv", "b_qkv", "w_o", "b_o", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
                                        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "s_qkv", "s_fc1")]


class Vit(C.Structure):
    _fields_ = [("image_size", c_int), ("patch", c_int), ("hidden", c_int), ("heads", c_int), ("layers", c_int),
                ("mlp", c_int), ("ln_eps", c_float), ("w_patch", c_void_p), ("cls", c_void_p), ("pos", c_void_p),
                ("pre_ln_g", c_void_p), ("pre_ln_b", c_void_p), ("layer", C.POINTER(VitLayer)), ("flags", c_int)]


VIT_RESIDUAL_F32 = 1
VIT_PATCH_SPLIT = 2
VIT_LN_FOLD = 4
ABI_VERSION = 3


class Attn(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("w_qkv", "b_qkv", "w_proj", "b_proj")]


class Block(C.Structure):
    _fields_ = [("depth", c_int), ("n1_g", c_void_p), ("n1_b", c_void_p), ("n2_g", c_void_p), ("n2_b", c_void_p),
                ("attn", C.POINTER(Attn)), ("w_fc1", c_void_p), ("b_fc1", c_void_p), ("w_fc2", c_void_p), ("b_fc2", c_void_p)]


class Head(C.Structure):
    _fields_ = [("hidden", c_int), ("heads", c_int), ("mlp", c_int), ("token_dim", c_int), ("inner", Block),
                ("inter", Block), ("w_out", c_void_p), ("b_out", c_void_p)]


class Projector(C.Structure):
    _fields_ = [("n_linear", c_int), ("w", C.POINTER(c_void_p)), ("b", C.POINTER(c_void_p)), ("dims", C.POINTER(c_int)),
                ("norm_g", c_void_p), ("norm_b", c_void_p)]


class U8Norm(C.Structure):
    _fields_ = [("lut", c_float * 256), ("mean", c_float * 3), ("std", c_float * 3)]


class QFormerLayer(C.Structure):
    _fields_ = [("w_qkv", c_void_p), ("b_qkv", c_void_p), ("w_so", c_void_p), ("b_so", c_void_p), ("ln_s_g", c_void_p), ("ln_s_b", c_void_p),
                ("has_cross", c_int), ("w_cq", c_void_p), ("b_cq", c_void_p), ("w_ckv", c_void_p), ("b_ckv", c_void_p), ("w_co", c_void_p),
                ("b_co", c_void_p), ("ln_c_g", c_void_p), ("ln_c_b", c_void_p), ("w_f1", c_void_p), ("b_f1", c_void_p), ("w_f2", c_void_p),
                ("b_f2", c_void_p), ("ln_f_g", c_void_p), ("ln_f_b", c_void_p)]


class Detok(C.Structure):
    _fields_ = [("token_dim", c_int), ("hidden", c_int), ("q_heads", c_int), ("q_inter", c_int), ("q_layers", c_int), ("grid", c_int),
                ("dec_dim", c_int), ("dec_heads", c_int), ("dec_mlp", c_int), ("dec_depth", c_int), ("q_ln_eps", c_float), ("dec_ln_eps", c_float),
                ("w_map_in", c_void_p), ("b_map_in", c_void_p), ("mask_tokens", c_void_p), ("emb_ln_g", c_void_p), ("emb_ln_b", c_void_p),
                ("qlayer", C.POINTER(QFormerLayer)), ("w_dec_in", c_void_p), ("b_dec_in", c_void_p), ("pos", c_void_p),
                ("block", C.POINTER(VitLayer)), ("norm_g", c_void_p), ("norm_b", c_void_p)]


class ResizeDesc(C.Structure):
    _fields_ = [("src", c_void_p), ("H", c_int), ("W", c_int), ("pad_x", c_int), ("pad_y", c_int), ("y0", c_int), ("y1", c_int), ("top", c_int),
                ("left", c_int), ("identity_x", c_int), ("identity_y", c_int), ("ksize_x", c_int), ("ksize_y", c_int), ("kx_off", c_int),
                ("ky_off", c_int), ("tmp_off", C.c_longlong)]


# name -> (restype, argtypes); every symbol include/setok_b200.h declares
SIGNATURES = {
    "setok_last_error": (C.c_char_p, []),
    "setok_abi_version": (c_int, []),
    "setok_launch_count": (C.c_uint64, []),
    "setok_gemm_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "setok_gemm_bf16_ln": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p]),
    "setok_ln_fold_init": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p]),
    "setok_gemm_bf16_batched": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int,
                                        c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "setok_layernorm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p,
                                c_void_p, c_void_p]),
    "setok_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "setok_vit_workspace_bytes": (c_size_t, [C.POINTER(Vit), c_int]),
    "setok_vit_forward": (c_int, [C.POINTER(Vit), c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "setok_vit_forward_u8": (c_int, [C.POINTER(Vit), c_void_p, C.POINTER(U8Norm), c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "setok_vit_forward_pos": (c_int, [C.POINTER(Vit), c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_dpc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "setok_dpc_cluster_embedded": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_dpc_cluster": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_dpc_cluster_pos": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_head_workspace_bytes": (c_size_t, [C.POINTER(Head), c_int, c_int]),
    "setok_head_forward": (c_int, [C.POINTER(Head), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_detok_workspace_bytes": (c_size_t, [C.POINTER(Detok), c_int, c_int]),
    "setok_detok_forward": (c_int, [C.POINTER(Detok), c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "setok_splice_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "setok_splice_plan": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_splice_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_splice": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "setok_preprocess_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "setok_transpose_to_bf16": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "setok_colsum_add": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "setok_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "setok_gelu_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "setok_gelu_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "setok_masked_softmax": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_int64, c_void_p]),
    "setok_masked_softmax_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_int64, c_void_p]),
    "setok_sort_by_cluster": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "setok_segment_mean": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "setok_segment_mean_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "setok_gather_rows_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "setok_project_workspace_bytes": (c_size_t, [C.POINTER(Projector), c_int]),
    "setok_project": (c_int, [C.POINTER(Projector), c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the library (once).  Raises SetokError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SetokError(f"{LIB_PATH} not found: build the CUDA library first (python -m setok_b200.build). "
                         "setok_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SetokError(f"{LIB_PATH} does not export {name}; stale build?") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().setok_last_error().decode("utf-8", "replace")
        raise SetokError(f"{what} failed with status {status}: {msg}")


def launch_count() -> int:
    return int(load().setok_launch_count())

"""A few launches of the clustering path at the BASELINE config-2 shape, for ncu captures."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import _lib, ops
from setok_b200.synth import mog_features
dev = torch.device("cuda:0")
lib = _lib.load()
lib.setok_debug_set_dpc_fused.argtypes = [ctypes.c_int]
lib.setok_debug_set_dpc_fused(int(os.environ.get("SETOK_DPC_FUSED", "1")))
feats = mog_features(256, 256, 1024, 7, dev)
if os.environ.get("SETOK_DPC_BF16"):
    feats = feats.to(torch.bfloat16)
noise = torch.rand(256, 256, device=dev)
emb = bool(os.environ.get("SETOK_DPC_EMBEDDED"))
for _ in range(4):
    ops.dpc_cluster(feats, noise, (16, 16), int(os.environ.get("SETOK_DPC_K", "16")), 0.5, 64, embedded=emb)
torch.cuda.synchronize()

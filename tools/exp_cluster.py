"""Timing experiments on the fused clustering kernel (debug flag bits: 8 = no x_pos store, 16 = no pos load)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import _lib, ops
from setok_b200.synth import mog_features
dev = torch.device("cuda:0")
lib = _lib.load()
lib.setok_debug_set_dpc_fused.argtypes = [ctypes.c_int]
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
noise = torch.rand(256, 256, device=dev)
for dt in (torch.float32, torch.bfloat16):
    for B in (148, 256):
        feats = mog_features(B, 256, 1024, 7, dev).to(dt)
        nz = torch.rand(B, 256, device=dev)
        for mode in (1, 2, 25):
            lib.setok_debug_set_dpc_fused(mode)
            ms = timeit(lambda: ops.dpc_cluster(feats, nz, (16, 16), 16, 0.5, 64))
            print(f"{str(dt):15s} B={B} mode={mode:2d}: {ms*1e3:7.1f} us")
lib.setok_debug_set_dpc_fused(1)

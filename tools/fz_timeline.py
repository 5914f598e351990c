"""Timeline of the fused clustering kernel (CTA 0, first image).  Build with SETOK_NVCC_EXTRA=-DSETOK_FZ_TRACE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import _lib, ops
from setok_b200.synth import mog_features
dev = torch.device("cuda:0")
lib = _lib.load()
lib.setok_debug_set_dpc_fused.argtypes = [ctypes.c_int]
lib.setok_debug_set_fz_trace.argtypes = [ctypes.c_void_p]
mode = int(os.environ.get("SETOK_DPC_FUSED", "25"))
lib.setok_debug_set_dpc_fused(mode)
dt = torch.bfloat16 if os.environ.get("SETOK_DPC_BF16") else torch.float32
feats = mog_features(256, 256, 1024, 7, dev).to(dt)
noise = torch.rand(256, 256, device=dev)
for _ in range(3):
    ops.dpc_cluster(feats, noise, (16, 16), 16, 0.5, 64)
torch.cuda.synchronize()
buf = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
lib.setok_debug_set_fz_trace(buf.data_ptr())
ops.dpc_cluster(feats, noise, (16, 16), 16, 0.5, 64)
torch.cuda.synchronize()
lib.setok_debug_set_fz_trace(None)
t = buf.cpu().view(4, 256)
t0 = int(t[t > 0].min())
rel = lambda x: (int(x) - t0) / 1e3 if int(x) > 0 else float("nan")
print(f"mode {mode} dtype {dt}: times in us since the first event")
print("MMA: tempty", rel(t[0, 0]))
for hk in range(32):
    print(f"  hk {hk:2d}: full seen {rel(t[0,1+2*hk]):7.2f}  issued {rel(t[0,2+2*hk]):7.2f} | row0: empty ok {rel(t[1,3*hk]):7.2f} raw ok {rel(t[1,3*hk+1]):7.2f} published {rel(t[1,3*hk+2]):7.2f} | tma tile wait-done {rel(t[3,hk]):7.2f}")
print("select: convert done", rel(t[2, 0]), "tfull", rel(t[2, 1]), "rowpass done", rel(t[2, 2]), "parent done", rel(t[2, 3]), "end", rel(t[2, 4]))

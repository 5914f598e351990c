"""In-kernel timeline (clock64) of one persistent CTA of the whole-row attention kernel, local pair 2 (steady state).
Needs the trace build:  SETOK_NVCC_EXTRA=-DSETOK_ATTN_TRACE SETOK_BUILD_OUT=.../libsetok_b200_trace.so python -m setok_b200.build
then  SETOK_B200_LIB=.../libsetok_b200_trace.so python tools/attn_fullrow_timeline.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import ops, _lib
dev = torch.device("cuda:0")
B = 256
qkv = torch.randn(B * 257, 3072, device=dev).to(torch.bfloat16)
buf = torch.zeros(3 * 64, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.setok_debug_set_fullrow_trace.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attention(qkv, 16, 0.125, uniform_T=257)
torch.cuda.synchronize()
lib.setok_debug_set_fullrow_trace(buf.data_ptr())
ops.attention(qkv, 16, 0.125, uniform_T=257)
torch.cuda.synchronize()
lib.setok_debug_set_fullrow_trace(None)
t = buf.cpu().tolist()
names = {}
for role, nm in ((0, "A"), (1, "B")):
    for slot, what in enumerate(("o_full seen", "o_read arrived (TMEM drained)", "O stored, waiting s_full", "s_full seen", "pass 1 done", "p_full arrived (pass 2 done)", "257th-row slice done")):
        names[role * 64 + slot] = f"softmax {nm}: {what}"
for tt, nm in ((0, "A"), (1, "B")):
    for slot, what in enumerate(("p_full seen", "PV (+extras) issued, o_full committed", "o_read seen", "next S issued")):
        names[128 + 8 * tt + slot] = f"mma {nm}: {what}"
ev = sorted((v, names.get(i, str(i))) for i, v in enumerate(t) if v)
t0 = ev[0][0]
prev = 0
for c, n in ev:
    print(f"{c - t0:8d} (+{c - t0 - prev:6d})  {n}")
    prev = c - t0

import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import ops, _lib
dev = torch.device("cuda:0")
B = 256
qkv = torch.randn(B * 257, 3072, device=dev).to(torch.bfloat16)
buf = torch.zeros(128, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.setok_debug_set_attention_trace.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attention(qkv, 16, 0.125, uniform_T=257)
lib.setok_debug_set_attention_trace(buf.data_ptr())
ops.attention(qkv, 16, 0.125, uniform_T=257)
torch.cuda.synchronize()
lib.setok_debug_set_attention_trace(None)
t = buf.cpu().tolist()
t0 = t[0]
names = {0: "start", 1: "after setup sync", 2: "mma: q_full", 3: "mma: S0 issued", 80: "sm: o_final", 81: "sm: stored", 82: "dealloc done"}
for j in range(5):
    names[10 + 4 * j] = f"mma: S{j+1} issued"; names[11 + 4 * j] = f"mma: p_full {j}"; names[12 + 4 * j] = f"mma: PV{j} issued"
    names[40 + 4 * j] = f"sm: s_full {j}"; names[41 + 4 * j] = f"sm: max {j}"; names[42 + 4 * j] = f"sm: exps done {j}"; names[43 + 4 * j] = f"sm: arrived {j}"
ev = sorted((v - t0, names.get(i, str(i))) for i, v in enumerate(t) if v)
prev = 0
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
for c, n in ev:
    print(f"{c:8d} (+{c - prev:6d})  {n}")
    prev = c

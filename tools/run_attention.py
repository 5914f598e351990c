import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
qkv = torch.randn(B * 257, 3072, device=dev).to(torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, 16, 0.125, uniform_T=257)
torch.cuda.synchronize()

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from setok_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 65792, 3072, 1024
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = torch.randn(N, K, device=dev).to(torch.bfloat16)
b = torch.zeros(N, device=dev)
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.gemm(a, w, b, out=out)
torch.cuda.synchronize()

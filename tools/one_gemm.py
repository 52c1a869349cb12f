"""One GEMM shape, a few launches (for ncu): python tools/one_gemm.py M N K tA tB [det]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from d3feat.pytorch_b200 import ops
M, N, K, ta, tb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
det = len(sys.argv) > 6
dev = torch.device("cuda:0")
a = torch.randn((K, M) if ta else (M, K), device=dev)
b = torch.randn((N, K) if tb else (K, N), device=dev)
for _ in range(4):
    c = ops.gemm(a, b, bool(ta), bool(tb), deterministic=det)
torch.cuda.synchronize()
print(float(c.abs().sum()))

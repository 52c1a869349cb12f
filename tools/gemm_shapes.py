"""Representative GEMM shapes of one pair step (for ncu): python tools/gemm_shapes.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from d3feat.pytorch_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [  # (M, N, K, ta, tb)  as seen in bench.py's op breakdown
    (40000, 32, 384, False, True),    # last_unary forward
    (40000, 384, 32, False, False),   # last_unary dx
    (32, 384, 40000, True, False),    # last_unary dW (split-K)
    (40000, 128, 32, False, True),    # unary2 forward at level 0
    (693, 1024, 3072, False, True),   # decoder unary forward
    (7680, 512, 189, True, False),    # KPConv dW at level 4
    (12064, 64, 960, False, False),   # KPConv contraction at level 1
    (12064, 960, 64, False, True),    # KPConv dwf at level 1
]
for rep in range(2):
    for (M, N, K, ta, tb) in shapes:
        a = torch.randn((K, M) if ta else (M, K), device=dev)
        b = torch.randn((N, K) if tb else (K, N), device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = ops.gemm(a, b, ta, tb); e1.record(); torch.cuda.synchronize()
        if rep:
            fl = 2.0 * M * N * K; by = 4.0 * (M * K + N * K + M * N)
            print("M=%6d N=%5d K=%6d ta=%d tb=%d  %.1f us  %.1f TFLOP/s  %.0f GB/s(min traffic)" % (M, N, K, ta, tb, e0.elapsed_time(e1) * 1e3, fl / e0.elapsed_time(e1) / 1e9, by / e0.elapsed_time(e1) / 1e6))

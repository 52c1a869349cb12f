"""Per-phase cycle breakdown of the tcgen05 GEMM main loop (diagnostic build with -DD3F_TC5_TIMING).
  python tools/tc5_timing.py build     (here: nvcc -> tools/_diag/libd3feat_b200_timing.so)
  python tools/tc5_timing.py           (on the GPU box)
The diagnostic library is a separate .so; the product library is never compiled with the timing code."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3feat.pytorch_b200 import _lib
DIAG = os.path.join(ROOT, "tools", "_diag", "libd3feat_b200_timing.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    os.makedirs(os.path.dirname(DIAG), exist_ok=True)
    cmd = ["nvcc", "-DD3F_TC5_TIMING"] + _lib.NVCC_FLAGS + ["-o", DIAG] + _lib.sources()
    subprocess.check_call(cmd)
    print("built", DIAG)
    sys.exit(0)

import torch
_lib.LIB_PATH = DIAG
lib = _lib.load()
lib.d3f_tc5_timing.restype = ctypes.c_int
lib.d3f_tc5_timing.argtypes = [ctypes.c_void_p]
from d3feat.pytorch_b200 import ops
dev = torch.device("cuda:0")
names = ["wait MMAs of previous tile (mbarrier)", "store_tile (incl. waiting for the global loads)", "issue next global loads",
         "fence.proxy.async + tcgen05 fence", "__syncthreads", "MMA issue + commit (thread 0)", "prologue", "epilogue: shared C tile -> global", "", "",
         "epilogue: TMEM -> registers -> shared C tile", "epilogue: __syncthreads"]
shapes = [(40000, 32, 480, False, False, "L0 contraction"), (13312, 64, 960, False, False, "L1 contraction"),
          (40000, 128, 32, False, True, "unary 32->128"), (480, 32, 40000, True, False, "L0 dW"), (768, 1024, 3072, False, True, "decoder unary"),
          (768, 1024, 256, False, False, "small dx"), (2816, 512, 128, False, True, "small unary"), (256, 1024, 768, True, False, "small dW"),
          (13312, 128, 64, False, False, "tiny dx")]
for (M, N, K, ta, tb, label) in shapes:
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev)
    for _ in range(3):
        ops.gemm(a, b, ta, tb, deterministic=not ta)
    torch.cuda.synchronize()
    torch.cuda._sleep(300000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, b, ta, tb, deterministic=not ta); e1.record(); torch.cuda.synchronize()
    print("%s: whole op %.1f us" % (label, e0.elapsed_time(e1) * 1e3))
    buf = (ctypes.c_ulonglong * 32)()
    assert lib.d3f_tc5_timing(buf) == 0
    for who, off in (("thread 0 (MMA issuer)", 0), ("thread 255", 16)):
        tot, nk = buf[off + 8], buf[off + 9]
        print("%s  M=%d N=%d K=%d ta=%d tb=%d | %s: CTA total %d cycles, %d K tiles, %.0f cycles per tile in the loop"
              % (label, M, N, K, ta, tb, who, tot, nk, sum(buf[off + i] for i in range(6)) / max(nk, 1)))
        for i, nme in enumerate(names):
            if not nme: continue
            print("    %-52s %9d cycles  %5.1f %%" % (nme, buf[off + i], 100.0 * buf[off + i] / max(tot, 1)))

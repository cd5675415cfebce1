"""GPU probe: the tcgen05 tf32 GEMM (C = A B') at the two config-5 product shapes, both tile widths."""
import json, os, subprocess, sys
CODE = r"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
from stochqn_b200 import _lib
lib = _lib.load(np.float32).lib
for name, M, N, K, ldc in (("Z = X W' (B=1024)", 1024, 4096, 8192, 4096), ("G = D' X (B=1024)", 4096, 8192, 1024, 8193),
                           ("Z = X W' (B=4096)", 4096, 4096, 8192, 4096), ("G = D' X (B=4096)", 4096, 8192, 4096, 8193)):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M * ldc, device="cuda")
    for _ in range(3): lib.stochqn_b200_gemm_tn(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), ldc, M, N, K, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lib.stochqn_b200_gemm_tn(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), ldc, M, N, K, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps(dict(bn=os.environ.get("STOCHQN_B200_GEMM_BN", "auto"), product=name, M=M, N=N, K=K, ms=ms, tflops=2.0 * M * N * K / ms / 1e9)), flush=True)
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for bn in ("128", "256", "0"):
    subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, STOCHQN_B200_GEMM_BN=bn), check=True)

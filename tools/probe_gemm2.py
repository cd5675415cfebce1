"""GPU probe: the cta_group::2 tf32 GEMM (csrc/gemm_tf32_sm100.cuh, gemm_tf32_2sm_kernel) against the one-CTA kernel: correctness on
ragged shapes against an fp64 product (stated tf32 tolerance: 1.5e-3 of sum |a||b|), then timing at the config-5 product shapes.
Every variant runs in its own process under a time limit (a protocol mistake between the two CTAs of a pair shows up as a hang)."""
import json, os, subprocess, sys
CODE = r"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
from stochqn_b200 import _lib
lib = _lib.load(np.float32).lib
mode = sys.argv[1]
torch.manual_seed(0)
if mode == "check":
    for (M, N, K, ldc) in ((256, 256, 64, 256), (512, 512, 256, 512), (1024, 4096, 2048, 4096), (1000, 520, 1032, 523), (260, 300, 100, 301), (4096, 8192, 1024, 8193)):
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.full((M * ldc,), float("nan"), device="cuda")
        rc = lib.stochqn_b200_gemm_tn(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), ldc, M, N, K, None)
        torch.cuda.synchronize()
        ref = (A.double() @ B.double().T)
        got = C.view(M, ldc)[:, :N].double()
        bound = (A.double().abs() @ B.double().abs().T)
        err = float(((got - ref).abs() / bound).max())
        print(json.dumps(dict(mode=os.environ.get("STOCHQN_B200_GEMM_2SM", "1"), shape=[M, N, K], rc=rc, max_err_over_bound=err, ok=bool(err <= 1.5e-3))), flush=True)
else:
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
        print(json.dumps(dict(mode=os.environ.get("STOCHQN_B200_GEMM_2SM", "1"), product=name, M=M, N=N, K=K, ms=ms, tflops=2.0 * M * N * K / ms / 1e9)), flush=True)
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for what in ("check", "time"):
    for mode in ("2", "0"):          # 2: CTA pairs forced for every shape with M, N >= 256; 0: the one-CTA kernel
        try:
            subprocess.run([sys.executable, "-c", CODE, what], env=dict(os.environ, STOCHQN_B200_GEMM_2SM=mode), check=False, timeout=90)
        except subprocess.TimeoutExpired:
            print(json.dumps(dict(mode=mode, what=what, error="timeout (hang)")), flush=True)

"""GPU: multinomial-logistic loss / gradient / Hessian-vector callbacks against the NumPy restatement of the
scikit-learn (<= 1.0) arithmetic the reference's Python layer calls (oracle/multinomial_np.py; parity unpinned -
see its header).  fp64 and fp32 CUDA-core paths to rounding level; the fp32 tensor-core (tcgen05, tf32) path within the
stated tf32 tolerance, and against the same library with tensor cores switched off."""
import os

import numpy as np
import pytest

from oracle import multinomial_np as M
from stochqn_b200 import _lib

pytestmark = pytest.mark.gpu


def _t(a, dtype):
    import torch
    return torch.tensor(np.ascontiguousarray(a, dtype=dtype), device="cuda")


def _run(dtype, B, d, K, fit_intercept, use_labels, weighted, seed=0, scale=1.0):
    import torch
    abi = _lib.load(dtype)
    lib = abi.lib
    rng = np.random.default_rng(seed + B + 3 * d + 7 * K)
    X = (rng.standard_normal((B, d)) * scale / np.sqrt(d)).astype(dtype)
    lab = rng.integers(0, K, B).astype(np.int32)
    Y = np.eye(K)[lab].astype(dtype)
    nw = K * (d + int(fit_intercept))
    w = (rng.standard_normal(nw) * 0.5).astype(dtype)
    v = rng.standard_normal(nw).astype(dtype)
    sw = (rng.random(B) + 0.5).astype(dtype) if weighted else None
    alpha = 1e-3
    Xd, wd, vd = _t(X, dtype), _t(w, dtype), _t(v, dtype)
    Yd = None if use_labels else _t(Y, dtype)
    labd = torch.tensor(lab, device="cuda") if use_labels else None
    swd = _t(sw, dtype) if weighted else None
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(B, d, K), device="cuda", dtype=torch.uint8)
    g = torch.zeros_like(wd)
    hv = torch.zeros_like(wd)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    yp = Yd.data_ptr() if Yd is not None else None
    lp = labd.data_ptr() if labd is not None else None
    sp = swd.data_ptr() if swd is not None else None
    assert lib.stochqn_b200_multinomial_loss_grad(Xd.data_ptr(), d, yp, K, lp, sp, B, d, K, int(fit_intercept), wd.data_ptr(), alpha,
                                                  g.data_ptr(), loss.data_ptr(), work.data_ptr(), None) == 0
    assert lib.stochqn_b200_multinomial_hess_vec(Xd.data_ptr(), d, yp, K, lp, sp, B, d, K, int(fit_intercept), wd.data_ptr(), vd.data_ptr(),
                                                 alpha, hv.data_ptr(), work.data_ptr(), None) == 0
    torch.cuda.synchronize()
    X64, w64, v64 = X.astype(np.float64), w.astype(np.float64), v.astype(np.float64)
    sw64 = None if sw is None else sw.astype(np.float64)
    lr, gr, _ = M.multinomial_loss_grad(w64, X64, Y.astype(np.float64), alpha, sw64)
    hr = M.multinomial_hess_vec(w64, v64, X64, Y.astype(np.float64), alpha, sw64)
    return (g.cpu().numpy().astype(np.float64), hv.cpu().numpy().astype(np.float64), float(loss.item())), (gr, hr, lr)


SHAPES = [(1, 1, 2), (20, 7, 5), (257, 64, 33), (50, 1836, 159), (300, 130, 70), (600, 40, 450)]   # the last one takes the stats + tiled row kernels (B*K >= 2^18)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 3e-5)])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("fit_intercept,use_labels,weighted", [(True, False, False), (False, True, True), (True, True, True)])
def test_multinomial_cuda_cores(dtype, tol, shape, fit_intercept, use_labels, weighted):
    B, d, K = shape
    os.environ["STOCHQN_B200_NO_TENSOR_CORES"] = "1"
    try:
        (g, hv, loss), (gr, hr, lr) = _run(dtype, B, d, K, fit_intercept, use_labels, weighted)
    finally:
        os.environ.pop("STOCHQN_B200_NO_TENSOR_CORES", None)
    assert np.max(np.abs(g - gr)) <= tol * max(np.max(np.abs(gr)), 1e-3)
    assert np.max(np.abs(hv - hr)) <= tol * max(np.max(np.abs(hr)), 1e-3)
    assert abs(loss - lr) <= max(tol, 1e-12) * abs(lr)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 3e-5)])
@pytest.mark.parametrize("shape", [(1, 1, 2), (20, 7, 5), (50, 1836, 159), (128, 96, 96), (64, 2300, 130), (24, 33, 512), (3, 2368, 17)])
@pytest.mark.parametrize("fit_intercept,use_labels,weighted", [(True, False, False), (False, True, True), (True, True, True)])
def test_small_batch_gradient_in_one_launch(dtype, tol, shape, fit_intercept, use_labels, weighted):
    """Gradient-only requests on small batches take mn_grad_small (one cooperative launch, csrc/multinomial.cu): same result as
    the oracle and as the five-launch route (which the loss + gradient call takes)."""
    import torch
    B, d, K = shape
    abi = _lib.load(dtype)
    lib = abi.lib
    rng = np.random.default_rng(5 + B + 3 * d + 7 * K)
    X = (rng.standard_normal((B, d)) / np.sqrt(d)).astype(dtype)
    lab = rng.integers(0, K, B).astype(np.int32)
    Y = np.eye(K)[lab].astype(dtype)
    w = (rng.standard_normal(K * (d + int(fit_intercept))) * 0.5).astype(dtype)
    sw = (rng.random(B) + 0.5).astype(dtype) if weighted else None
    alpha = 1e-2
    Xd, wd = _t(X, dtype), _t(w, dtype)
    Yd = None if use_labels else _t(Y, dtype)
    labd = torch.tensor(lab, device="cuda") if use_labels else None
    swd = _t(sw, dtype) if weighted else None
    work = torch.empty(lib.stochqn_b200_multinomial_work_size(B, d, K), device="cuda", dtype=torch.uint8)
    g1, g5 = torch.zeros_like(wd), torch.zeros_like(wd)
    loss = torch.zeros(1, device="cuda", dtype=torch.float64)
    yp = Yd.data_ptr() if Yd is not None else None
    lp = labd.data_ptr() if labd is not None else None
    sp = swd.data_ptr() if swd is not None else None
    n0 = _lib.launch_count()
    for _ in range(3):                         # (the barrier words are re-armed by every call)
        assert lib.stochqn_b200_multinomial_loss_grad(Xd.data_ptr(), d, yp, K, lp, sp, B, d, K, int(fit_intercept), wd.data_ptr(), alpha,
                                                      g1.data_ptr(), None, work.data_ptr(), None) == 0
    n1 = _lib.launch_count()
    assert n1 - n0 == 3, "the gradient-only call did not take the one-launch route"
    assert lib.stochqn_b200_multinomial_loss_grad(Xd.data_ptr(), d, yp, K, lp, sp, B, d, K, int(fit_intercept), wd.data_ptr(), alpha,
                                                  g5.data_ptr(), loss.data_ptr(), work.data_ptr(), None) == 0
    torch.cuda.synchronize()
    _, gr, _ = M.multinomial_loss_grad(w.astype(np.float64), X.astype(np.float64), Y.astype(np.float64), alpha,
                                       None if sw is None else sw.astype(np.float64))
    scale = max(np.max(np.abs(gr)), 1e-3)
    assert np.max(np.abs(g1.cpu().numpy().astype(np.float64) - gr)) <= tol * scale
    assert np.max(np.abs(g1.cpu().numpy().astype(np.float64) - g5.cpu().numpy().astype(np.float64))) <= tol * scale


@pytest.mark.parametrize("shape", [(512, 1024, 512), (1000, 1032, 520), (256, 2048, 1030)])
@pytest.mark.parametrize("fit_intercept", [True, False])
def test_multinomial_tensor_cores_fp32(shape, fit_intercept):
    """tcgen05 tf32 products: inputs rounded to 10-bit mantissas, fp32 accumulation.  Stated tolerance: 2e-3 of the
    largest entry against the fp64 oracle, and the same against this library's own CUDA-core fp32 path."""
    B, d, K = shape
    (g, hv, loss), (gr, hr, lr) = _run(np.float32, B, d, K, fit_intercept, True, True, scale=3.0)
    assert np.all(np.isfinite(g)) and np.all(np.isfinite(hv))
    assert np.max(np.abs(g - gr)) <= 2e-3 * np.max(np.abs(gr))
    assert np.max(np.abs(hv - hr)) <= 2e-3 * np.max(np.abs(hr))
    assert abs(loss - lr) <= 2e-3 * abs(lr)
    os.environ["STOCHQN_B200_NO_TENSOR_CORES"] = "1"
    try:
        (g2, hv2, loss2), _ = _run(np.float32, B, d, K, fit_intercept, True, True, scale=3.0)
    finally:
        os.environ.pop("STOCHQN_B200_NO_TENSOR_CORES", None)
    assert np.max(np.abs(g - g2)) <= 2e-3 * np.max(np.abs(g2))
    assert np.max(np.abs(g - g2)) > 0.0        # the two paths really are different arithmetic


@pytest.mark.parametrize("shape", [(128, 128, 64), (1024, 4096, 2048), (1000, 520, 1032), (130, 2050, 516), (4096, 300, 1024), (512, 256, 4100)])
@pytest.mark.parametrize("bn", ["0", "128", "256"])
def test_gemm_tn_tensor_cores_vs_fp64(shape, bn):
    """C = A B' on the tcgen05 path (both tile widths, ragged edges, K not a multiple of the k-block) against an fp64
    product of the same fp32 inputs: tf32 rounding only (|err| <= 2^-10 * sum |a||b| per entry, checked as 1.5e-3 of
    the row-column magnitude bound)."""
    import subprocess, sys, json
    M, N, K = shape
    code = """
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
from stochqn_b200 import _lib
lib = _lib.load(np.float32).lib
M, N, K = %d, %d, %d
g = torch.Generator(device="cuda"); g.manual_seed(M + N + K)
A = torch.randn(M, K, device="cuda", generator=g); B = torch.randn(N, K, device="cuda", generator=g)
C = torch.full((M, N), float("nan"), device="cuda")
assert lib.stochqn_b200_gemm_tn(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, M, N, K, None) == 0
torch.cuda.synchronize()
ref = A.double() @ B.double().T
bound = A.abs().double() @ B.abs().double().T
err = ((C.double() - ref).abs() / bound).max().item()
exact = (C.double() - ref).abs().max().item()
print(json.dumps(dict(err=err, finite=bool(torch.isfinite(C).all().item()), exact=exact)))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), M, N, K)
    env = dict(os.environ, STOCHQN_B200_GEMM_BN=bn)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["finite"], res
    assert res["err"] <= 1.5e-3, res

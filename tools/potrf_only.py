"""Factor-only Cholesky (gpk_potrf_lower_dev) and the blocked triangular solve alone, device-resident, CUDA-event timed.
  [GPK_POTRF_NB=256|512|1024] python tools/potrf_only.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_algos_b200 import _lib, synthetic, MatrixUtils as MU

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
X, y, th = synthetic.make_c2(n=n, D=8)
ts = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(ts)
h = _lib.Handle(0, ts.cuda_stream)
dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()
K0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(K0)
info = torch.zeros(1, dtype=torch.int32, device="cuda")
thc = np.ascontiguousarray(th)
h.check(h.lib.gpk_cov_se_ard_dev(h.h, dX.data_ptr(), n, 8, n, _lib.ptr(thc), K0.data_ptr(), n))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts_ = []
for rep in range(6):
    A.copy_(K0); torch.cuda.synchronize(); e0.record()
    h.check(h.lib.gpk_potrf_lower_dev(h.h, A.data_ptr(), n, n, info.data_ptr()))
    e1.record(); torch.cuda.synchronize(); ts_.append(e0.elapsed_time(e1))
ms = min(ts_[1:])
print(f"potrf_lower_dev n={n} nb={os.environ.get('GPK_POTRF_NB', '512')}: {ms:.3f} ms = {n**3 / 3 / ms * 1e-9:.2f} TFLOP/s   all={['%.3f' % t for t in ts_]}  info={int(info.item())}")
if n <= 8192 and not os.environ.get("SKIP_TRSM"):
    L = A.cpu().numpy().reshape(n, n).T.copy()
    b = np.random.default_rng(0).standard_normal(n)
    MU.forwardSolve(L, b)
    t0 = time.perf_counter(); x = MU.forwardSolve(L, b); t1 = time.perf_counter() - t0
    print(f"forwardSolve(L, y) host API n={n}: {t1 * 1e3:.2f} ms  (includes the {n * n * 8 / 1e6:.0f} MB upload of L); resid {np.abs(L @ x - b).max():.2e}")

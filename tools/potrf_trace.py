"""GPK_TRACE=1 timeline of ONE factor-only Cholesky (gpk_potrf_lower_dev, n = 8192): ms since the start per milestone and stream."""
import os, sys
os.environ["GPK_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_algos_b200 import _lib, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
X, y, th = synthetic.make_c2(n=n, D=8)
h = _lib.Handle(0)
dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()
K0 = torch.empty(n * n, dtype=torch.float64, device="cuda"); A = torch.empty_like(K0)
thc = np.ascontiguousarray(th)
h.check(h.lib.gpk_cov_se_ard_dev(h.h, dX.data_ptr(), n, 8, n, _lib.ptr(thc), K0.data_ptr(), n))
for rep in range(2):
    A.copy_(K0); torch.cuda.synchronize()
    if rep == 1: print("---- second run ----", file=sys.stderr)
    h.check(h.lib.gpk_potrf_lower_dev(h.h, A.data_ptr(), n, n, None)); h.synchronize()

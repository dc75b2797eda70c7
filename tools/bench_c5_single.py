"""BASELINE.json config 5 on ONE B200: n = 65536, D = 8 -- K build, FP64 Cholesky (+ L^-1) and alpha = K^-1 y, all resident
(2 x 32 GiB matrices + 13 GiB scratch of the 180 GB).  Checks ||K alpha - y|| / ||y|| with K regenerated block by block."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_algos_b200 as gp
from gp_algos_b200 import synthetic

n = int(os.environ.get("C5_N", 65536))
X, y, th = synthetic.make_c2(n=n, D=8, seed=5)
kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
pred = gp.GpPredictor(kf)
t0 = time.perf_counter(); fitted = pred.fit(X, None, y, th); t_first = time.perf_counter() - t0
fitted.close()
t0 = time.perf_counter(); fitted = pred.fit(X, None, y, th); t_fit = time.perf_counter() - t0
alpha = fitted.alphaVec
# residual with K regenerated in 2048-column blocks (cross-covariance kernel) + sn^2 alpha
r = th[-1] ** 2 * alpha - y
blk = 2048
for j0 in range(0, n, blk):
    Kb = gp.MatrixUtils.buildKernelMatrix(kf, X, X[j0:j0 + blk])      # n x blk, no noise
    r += Kb @ alpha[j0:j0 + blk]
res = float(np.linalg.norm(r) / np.linalg.norm(y))
print(json.dumps({"config": f"C5 single GPU: n={n}, D=8, factor + L^-1 + alpha", "fit_seconds": t_fit, "first_call_seconds": t_first,
                  "potrf_plus_trtri_tflops": 2 * float(n) ** 3 / 3 / t_fit * 1e-12, "ll": fitted.logLikelihood,
                  "residual_Kalpha_minus_y_over_y": res}))

"""Two EP sweeps at n = 4096, D = 4 (BASELINE.json config 3) through the host API -- the short command profiled by ncu."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_algos_b200 as gp

n = int(os.environ.get("C3_N", 4096))
rng = np.random.default_rng(3)
X = rng.standard_normal((n, 4)); w = rng.standard_normal(4)
t = np.where(X @ w + 0.3 * rng.standard_normal(n) >= 0, 1, -1).astype(np.int32)
kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(1.0, [1.0] * 4, 0.0))
K = gp.MatrixUtils.buildKernelMatrix(kf, X)
for i in range(int(os.environ.get("C3_REPS", 2))):
    t0 = time.perf_counter()
    site, L = gp.EpParameterEstimator(K, t, gp.FixedSweeps(2)).estimateSiteParams
    print(f"run {i}: {1e3 * (time.perf_counter() - t0) / 2:.2f} ms/sweep  logZ={site.marginalLogLikelihood:.10g}")

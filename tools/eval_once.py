"""Three MLE objective+gradient evaluations at n = 8192, D = 8 (BASELINE.json config 2) through the host API -- the short
command profiled by ncu (launch list and --set full captures under profiles/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_algos_b200 as gp

n = int(os.environ.get("EVAL_N", 8192))
rng = np.random.default_rng(2)
X = rng.uniform(0.0, 1.0, size=(n, 8)); w = rng.standard_normal(8)
y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
theta = np.concatenate([[1.0], np.full(8, 0.7), [0.1]])
pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(theta[0], theta[1:-1], theta[-1])))
for i in range(int(os.environ.get("EVAL_REPS", 3))):
    t0 = time.perf_counter()
    ll, g = pred.logLikelihoodWithDerivatives(gp.PredictionTrainingInput(X, None, y), theta, 10)
    print(f"eval {i}: {1e3 * (time.perf_counter() - t0):.2f} ms  ll={ll:.10g}  |g|max={np.abs(g).max():.6g}")

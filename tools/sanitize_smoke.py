"""One pass over every C-ABI entry point at small (and one pipelined) sizes: a coverage smoke run.  (Written as the target of
`compute-sanitizer --tool memcheck`; that tool is closed on this GPU pool, so out-of-bounds protection rests on the padded
layouts, the parity tests at ragged sizes (n = 1, 2, 127, 129, 255, 4300, ...) and the oracle comparison.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_algos_b200 as gp
from gp_algos_b200 import batched

rng = np.random.default_rng(0)
def data(n, D):
    X = rng.uniform(size=(n, D)); y = np.sin(X @ rng.standard_normal(D)) + 0.1 * rng.standard_normal(n)
    th = np.concatenate([[1.0], np.full(D, 0.7), [0.1]])
    return X, y, th
big = int(os.environ.get("SAN_BIG", 4200))
for n, D in ((333, 3), (big, 8)):
    X, y, th = data(n, D)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
    pred = gp.GpPredictor(kf)
    K = gp.MatrixUtils.buildKernelMatrix(kf, X)
    Kc = gp.MatrixUtils.buildKernelMatrix(kf, X[:17], X)
    dK = gp.MatrixUtils.buildKernelDerMatrix(kf, X[:200], 2)
    L = gp.MatrixUtils.cholesky(K)
    if n < 1000:
        Li = gp.MatrixUtils.invTriangular(L)
        z = gp.MatrixUtils.forwardSolve(L, y); a = gp.MatrixUtils.backSolve(L.T, z)
        Z = gp.MatrixUtils.forwardSolve(L, K[:, :5])
    ll, g = pred.logLikelihoodWithDerivatives(gp.PredictionTrainingInput(X, None, y), th, D + 2)
    ll2, _ = pred.logLikelihoodWithDerivatives(gp.PredictionTrainingInput(X, 0.05, y), th, 0)
    dist, ll3 = pred.predict(gp.PredictionInput(X, X[:9] + 0.01, None, y), th)
    m = pred.fit(X, None, y, th)
    d2, V = m.computePosterior(X[:3] + 0.02)
    ucb, gu, _, _ = gp.ucb_with_gradient(m, X[:4] + 0.03, 1.5)
    m.close()
    print(f"n={n}: ll={ll:.6f} |g|={np.abs(g).max():.4g} ucb0={ucb[0]:.4f}", flush=True)
# batched
B, n, D = 3, 260, 4
Xb = rng.uniform(size=(B, n, D)); yb = rng.standard_normal((B, n)); thb = np.tile(np.concatenate([[1.0], np.full(D, 0.7), [0.2]]), (B, 1))
llb, gb, info = batched.log_likelihood_with_derivatives_batched(Xb, yb, thb)
mb, vb, _, _ = batched.predict_batched(Xb, yb, thb, rng.uniform(size=(B, 5, D)))
print("batched", llb, flush=True)
# EP
n, D = 300, 3
X = rng.standard_normal((n, D)); t = np.where(X @ rng.standard_normal(D) + 0.3 * rng.standard_normal(n) >= 0, 1, -1).astype(np.int32)
kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(1.0, [1.0] * D, 0.1))
K = gp.MatrixUtils.buildKernelMatrix(kf, X)
site, L = gp.EpParameterEstimator(K, t, gp.FixedSweeps(2)).estimateSiteParams
p = gp.GpClassifier(gp.FixedSweeps(2)).classify(gp.AfterEstimationClassifierInput(t, (site, L), None, K, K[:7], K[:7, :7]))
ev = gp.MarginalLikelihoodEvaluator(gp.FixedSweeps(2), kf)
lz, ge = ev.logLikelihood(X, t, kf.theta)
g2 = ev.logLikelihoodDerivativesAfterHyperParams(gp.HyperParameterOptimInput(site, L, K, X), kf)
print("ep", site.marginalLogLikelihood, lz, ge[:2], flush=True)
# block-cyclic solver on a 1 x 1 grid
from gp_algos_b200.distributed import DistributedGp
X, y, th = data(700, 5)
s = DistributedGp(nb=256, device=0)
f = s.fit(X, y, th)
print("distributed", f.logLikelihood, s.residual(y, f.alphaVec), flush=True)
print("sanitize_smoke done")

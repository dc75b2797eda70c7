"""BASELINE.json config 3: EP binary GP classification, n=4096, D=4 (synthetic seed 3): s/sweep for fixed 5 sweeps and for
the shipped criterion (eps = 0.01), through the public host API (K on the host in, site parameters out)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_algos_b200 as gp
from gp_algos_b200 import synthetic

n = int(os.environ.get("C3_N", 4096))
X, t, th = synthetic.make_c3(n=n)
kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
t0 = time.perf_counter(); K = gp.MatrixUtils.buildKernelMatrix(kf, X); t_k = time.perf_counter() - t0
gp.EpParameterEstimator(K, t, gp.FixedSweeps(1)).estimateSiteParams  # warm-up
res = {}
for name, stop in (("fixed5", gp.FixedSweeps(5)), ("eps0.01", gp.AvgBasedStopCriterion(0.01))):
    est = gp.EpParameterEstimator(K, t, stop)
    t0 = time.perf_counter(); site, L = est.estimateSiteParams; dt = time.perf_counter() - t0
    res[name] = {"sweeps": est.sweeps, "seconds": dt, "s_per_sweep": dt / est.sweeps, "logZ": site.marginalLogLikelihood}
# size-independent properties at full size: tau >= 0 growth, posterior B = I + S^1/2 K S^1/2 = L L^t, probabilities in (0,1)
st = np.sqrt(site.tauSiteParams)
Bm = np.eye(n) + (st[:, None] * st[None, :]) * K
res["L_backward_error"] = float(np.linalg.norm(L @ L.T - Bm) / np.linalg.norm(Bm))
# (the CPU port of one sweep is timed by the tests' oracle, not here: tools never execute oracle/)
res["kernel_matrix_build_e2e_s"] = t_k
# fused route (MarginalLikelihoodEvaluator.logLikelihood, MarginalLikelihoodEvaluator.scala:33-45): X in, (logZ, gradient) out; K, L
# and the site parameters never cross PCIe
thg = th.copy(); thg[-1] = 0.1
kfg = gp.GaussianRbfKernel(gp.GaussianRbfParams(thg[0], thg[1:-1], thg[-1]))
ev = gp.MarginalLikelihoodEvaluator(gp.AvgBasedStopCriterion(0.01), kfg)
ev.logLikelihood(X, t, thg)
t0 = time.perf_counter(); lz, gr = ev.logLikelihood(X, t, thg); dt = time.perf_counter() - t0
res["fused_logZ_and_gradient"] = {"sweeps": ev.sweeps, "seconds": dt, "s_per_sweep": dt / ev.sweeps, "logZ": lz, "grad_inf_norm": float(np.abs(gr).max())}
print(json.dumps({"config": f"C3: EP classification n={n}, D=4", **res}))

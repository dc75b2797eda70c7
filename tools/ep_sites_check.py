"""EP site kernel variants (GPK_EP_SITES: 4 = register tile + W helper warps, 5 = warp-specialised with the branch-free scalar
update): site parameters / log Z agree to rounding (the scalar update is evaluated with other primitives), plus seconds per sweep.
No torch: ctypes only, starts in a second."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gp_algos_b200 as gp
from gp_algos_b200 import synthetic, _lib

h = _lib.Handle(0)
h.set_graph_mode(0)
out = {}
for n, sweeps in ((300, 4), (1000, 3), (int(os.environ.get("C3_N", 4096)), 4)):
    X, t, th = synthetic.make_c3(n=n)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
    K = gp.MatrixUtils.buildKernelMatrix(kf, X, handle=h)
    ref = None
    for sites in (4, 5):
        os.environ["GPK_EP_SITES"] = str(sites)
        gp.EpParameterEstimator(K, t, gp.FixedSweeps(1), h).estimateSiteParams
        est = gp.EpParameterEstimator(K, t, gp.FixedSweeps(sweeps), h)
        t0 = time.perf_counter(); site, L = est.estimateSiteParams; dt = time.perf_counter() - t0
        cur = (site.tauSiteParams, site.niSiteParams, site.marginalLogLikelihood)
        if ref is None:
            ref = cur
        d = max(np.abs(cur[0] - ref[0]).max() / np.abs(ref[0]).max(), np.abs(cur[1] - ref[1]).max() / np.abs(ref[1]).max(),
                abs(cur[2] - ref[2]) / abs(ref[2]))
        out[f"n{n}_sites{sites}"] = {"s_per_sweep_e2e": dt / sweeps, "logZ": cur[2], "max_rel_diff_vs_sites4": float(d)}
        print(n, sites, out[f"n{n}_sites{sites}"], flush=True)
print(json.dumps(out))

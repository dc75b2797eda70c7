"""One-shot GPU diagnostic: runs every libgpk entry point against the oracle and PRINTS errors
(no asserts), plus rough timings.  Development aid; the graded checks live in tests/."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
from oracle import gp_oracle as orc
import gp_algos_b200 as gp
from gp_algos_b200 import MatrixUtils as MU, _lib


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    h = _lib.default_handle()
    rng = np.random.default_rng(0)
    for n, D in ((3, 3), (100, 2), (128, 8), (300, 5), (1000, 1), (1100, 8)):
        X = rng.uniform(0, 1, size=(n, D)); y = np.sin(X @ rng.standard_normal(D)) + 0.1 * rng.standard_normal(n)
        th = orc.pack_theta(1.1, rng.uniform(0.5, 1.0, size=D), 0.15)
        kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
        K = MU.buildKernelMatrix(kf, X)
        Ko = orc.fast_build_kernel_matrix(X, th)
        print(f"n={n} D={D} cov rel {rel(K, Ko):.2e} sym {np.array_equal(K, K.T)}", flush=True)
        Xs = rng.uniform(0, 1, size=(17, D))
        Kc = MU.buildKernelMatrix(kf, Xs, X)
        print(f"   cross rel {rel(Kc, orc.fast_build_kernel_matrix(Xs, th, X)):.2e}")
        L = MU.cholesky(Ko)
        import scipy.linalg as sla
        Lo = sla.cholesky(Ko, lower=True)
        print(f"   potrf rel {rel(L, Lo):.2e}  resid {np.linalg.norm(L @ L.T - Ko) / np.linalg.norm(Ko):.2e} upper0 {np.all(np.triu(L, 1) == 0)}")
        Li = MU.invTriangular(Lo)
        print(f"   trtri rel {rel(Li, np.linalg.inv(Lo)):.2e}")
        Ui = MU.invTriangular(Lo.T.copy(), True)
        print(f"   trtri upper rel {rel(Ui, np.linalg.inv(Lo.T)):.2e}")
        b = rng.standard_normal(n); B = rng.standard_normal((n, 5))
        print(f"   fwd vec {rel(MU.forwardSolve(Lo, b), sla.solve_triangular(Lo, b, lower=True)):.2e}"
              f" back(L.t) {rel(MU.backSolve(Lo, b, transposed=True), sla.solve_triangular(Lo, b, lower=True, trans='T')):.2e}"
              f" fwd mat {rel(MU.forwardSolve(Lo, B), sla.solve_triangular(Lo, B, lower=True)):.2e}"
              f" back upper {rel(MU.backSolve(Lo.T.copy(), B), sla.solve_triangular(Lo.T, B, lower=False)):.2e}")
        pred = gp.GpPredictor(kf)
        for s in (None, 0.05):
            ll, g = pred.logLikelihoodWithDerivatives(gp.PredictionTrainingInput(X, s, y), th, D + 2)
            llo, go = orc.fast_loglik_with_derivs(X, y, th, s)
            print(f"   s={s} ll {ll:.12g} vs {llo:.12g} rel {abs(ll - llo) / abs(llo):.2e}  grad rel(max) {np.max(np.abs(g - go) / np.maximum(np.abs(go), 1e-9 * np.abs(go).max())):.2e}")
            Lg, ag, _ = pred.preComputeComponents(X, s, y, th)
            Lf, af = orc.fast_precompute(X, y, th, s)
            print(f"      fit L rel {rel(Lg, Lf):.2e} alpha rel {rel(ag, af):.2e}")
            dist, llp = pred.predict(gp.PredictionInput(X, Xs, s, y), th)
            mo, So, llo2 = orc.fast_predict(X, y, Xs, th, s)
            print(f"      predict mean rel {rel(dist.mean, mo):.2e} sigma rel {rel(dist.sigma, So):.2e} ll rel {abs(llp - llo2) / abs(llo2):.2e}")
        fitted = pred.fit(X, None, y, th)
        d2, V = fitted.computePosterior(Xs[:1], full_cov=False)
        Lf, af = orc.fast_precompute(X, y, th, None)
        mo, so, Vo = orc.fast_compute_posterior(X, Xs[:1], Lf, af, th, full_cov=False)
        print(f"   m=1 posterior mean rel {rel(d2.mean, mo):.2e} var rel {rel(d2.sigma, so):.2e} V rel {rel(V, Vo):.2e}")
        d3, V3 = pred.computePosterior(X, Xs, Lf, af)
        mo, So, Vo = orc.fast_compute_posterior(X, Xs, Lf, af, th)
        print(f"   computePosterior(from factor) mean {rel(d3.mean, mo):.2e} sigma {rel(d3.sigma, So):.2e} V {rel(V3, Vo):.2e}")
    # error paths
    try:
        MU.cholesky(np.array([[1., 2.], [3., 4.]]))
    except gp.MatrixNotSymmetricError as e:
        print("notsym OK")
    try:
        MU.cholesky(np.array([[1., 2.], [2., 1.]]))
    except gp.NotPositiveDefiniteError as e:
        print("notpd OK minor", e.minor)
    # timing at C2 size
    for n in (2048, 4096, 8192):
        X, y, th = orc.make_c2(n=n)
        kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
        pred = gp.GpPredictor(kf)
        inp = gp.PredictionTrainingInput(X, None, y)
        pred.logLikelihoodWithDerivatives(inp, th, 10)
        t0 = time.time(); reps = 3
        for _ in range(reps):
            ll, g = pred.logLikelihoodWithDerivatives(inp, th, 10)
        dt = (time.time() - t0) / reps
        print(f"n={n}: nll_grad e2e {dt * 1e3:.2f} ms  ({n**3 / dt * 1e-12:.2f} TFLOP/s eff) launches {h.launch_count()} ll={ll:.10g}", flush=True)
        if n <= 4096:
            t0 = time.time(); llo, go = orc.fast_loglik_with_derivs(X, y, th); tcpu = time.time() - t0
            print(f"   oracle(fast) {tcpu:.2f}s ll rel {abs(ll - llo) / abs(llo):.2e} grad rel {np.max(np.abs(g - go) / np.maximum(np.abs(go), 1e-9 * np.abs(go).max())):.2e}")


if __name__ == "__main__":
    main()

"""Defend the "parity unpinned" parts of the oracle: literal C restatement vs the LAPACK-backed
flavour vs a 50-digit mpmath evaluation, plus finite-difference gradient checks."""
import numpy as np
import pytest

from oracle import gp_oracle as orc


def _small(n=40, D=3, seed=7):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, size=(n, D))
    y = np.sin(X @ rng.standard_normal(D)) + 0.1 * rng.standard_normal(n)
    th = orc.pack_theta(1.3, rng.uniform(0.4, 1.1, size=D), 0.2)
    return X, y, th


def test_literal_vs_fast_kernel_matrix_2ulp():
    # same association of the distance sum; only exp() differs (libm vs NumPy SIMD), <= 2 ulp
    X, y, th = _small(64, 5)
    a, b = orc.lit_build_kernel_matrix(X, th), orc.fast_build_kernel_matrix(X, th)
    assert np.all(np.abs(a - b) <= 2 * np.spacing(np.abs(a)))
    Xs = X[:9] + 0.01
    a, b = orc.lit_build_kernel_matrix(Xs, th, X), orc.fast_build_kernel_matrix(Xs, th, X)
    assert np.all(np.abs(a - b) <= 2 * np.spacing(np.abs(a)))


@pytest.mark.parametrize("sigma_noise", [None, 0.05])
def test_literal_vs_fast_loglik_grad(sigma_noise):
    X, y, th = _small(150, 4)
    ll1, g1 = orc.lit_loglik_with_derivs(X, y, th, sigma_noise)
    ll2, g2 = orc.fast_loglik_with_derivs(X, y, th, sigma_noise, block=64)
    assert abs(ll1 - ll2) <= 1e-11 * abs(ll1)
    assert np.all(np.abs(g1 - g2) <= 1e-10 * np.maximum(np.abs(g1), 1e-9 * np.abs(g1).max()))


def test_gradient_matches_finite_differences():
    X, y, th = _small(60, 3)
    ll, g = orc.lit_loglik_with_derivs(X, y, th)
    for p in range(len(th)):
        h = 1e-6 * max(1.0, abs(th[p]))
        tp = th.copy(); tp[p] += h
        tm = th.copy(); tm[p] -= h
        fd = (orc.lit_loglik_with_derivs(X, y, tp)[0] - orc.lit_loglik_with_derivs(X, y, tm)[0]) / (2 * h)
        assert abs(fd - g[p]) <= 1e-5 * max(1.0, abs(g[p])), (p, fd, g[p])


def test_mpmath_arbiter_loglik_and_gradient():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    X, y, th = _small(14, 2, seed=11)
    n, D = X.shape
    sf, ls, sn = mp.mpf(th[0]), [mp.mpf(v) for v in th[1:D + 1]], mp.mpf(th[D + 1])
    Xm = [[mp.mpf(float(v)) for v in row] for row in X]

    def r(i, j):
        return sum((Xm[i][d] - Xm[j][d]) ** 2 / ls[d] ** 2 for d in range(D))

    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = sf ** 2 * mp.e ** (-r(i, j) / 2) + (sn ** 2 if i == j else 0)
    ym = mp.matrix([mp.mpf(float(v)) for v in y])
    Kinv = K ** -1
    alpha = Kinv * ym
    ll = -(ym.T * alpha)[0] / 2 - mp.log(mp.det(K)) / 2 - mp.mpf(n) / 2 * mp.log(2 * mp.pi)
    W = alpha * alpha.T - Kinv
    g = []
    for p in range(D + 2):
        tot = mp.mpf(0)
        for i in range(n):
            for j in range(n):
                e = mp.e ** (-r(i, j) / 2)
                if p == 0:
                    dk = 2 * sf * e
                elif p <= D:
                    dk = sf ** 2 * e * (Xm[i][p - 1] - Xm[j][p - 1]) ** 2 / ls[p - 1] ** 3
                else:
                    dk = 2 * sn if i == j else 0
                tot += W[i, j] * dk
        g.append(tot / 2)
    for flavour in (orc.lit_loglik_with_derivs, orc.fast_loglik_with_derivs):
        ll_o, g_o = flavour(X, y, th)
        assert abs(ll_o - float(ll)) <= 1e-11 * abs(float(ll))
        for p in range(D + 2):
            assert abs(g_o[p] - float(g[p])) <= 1e-9 * max(abs(float(g[p])), 1e-6), (p, g_o[p], float(g[p]))


def test_literal_vs_fast_predict():
    X, y, th = _small(120, 3)
    Xs = np.random.default_rng(3).uniform(0, 1, size=(17, 3))
    for s in (None, 0.3):
        m1, S1, ll1 = orc.lit_predict(X, y, Xs, th, s)
        m2, S2, ll2 = orc.fast_predict(X, y, Xs, th, s)
        assert np.allclose(m1, m2, rtol=1e-10, atol=1e-12)
        assert np.allclose(S1, S2, rtol=1e-9, atol=1e-12)
        assert abs(ll1 - ll2) <= 1e-11 * abs(ll1)
    # predictive covariance carries sn^2 on the diagonal (MatrixUtils.scala:63 via GpPredictor.scala:36)
    L, alpha = orc.lit_precompute(X, y, th)
    m, S, V = orc.lit_compute_posterior(X, X[:5], L, alpha, th)
    assert np.all(np.diag(S) > th[-1] ** 2 * 0.999)


def test_sigma_noise_added_unsquared():  # GpPredictor.scala:116-117
    X, y, th = _small(30, 2)
    L, _ = orc.lit_precompute(X, y, th, sigma_noise=0.5)
    K = orc.lit_build_kernel_matrix(X, th) + 0.5 * np.eye(30)
    assert np.allclose(L @ L.T, K, rtol=1e-12, atol=1e-13)


def test_ep_literal_vs_fast():
    X, t, th = orc.make_c3(n=96, D=3, seed=3)
    K = orc.lit_build_kernel_matrix(X, th)
    a = orc.lit_ep_estimate(K, t, fixed_sweeps=3)
    b = orc.fast_ep_estimate(K, t, fixed_sweeps=3)
    assert a["sweeps"] == b["sweeps"] == 3
    for k in ("tau", "nu", "mu"):
        assert np.allclose(a[k], b[k], rtol=1e-9, atol=1e-12), k
    assert abs(a["logZ"] - b["logZ"]) <= 1e-9 * abs(a["logZ"])
    # shipped criterion (spring-context.xml:53-55, eps = 0.01): same sweep count in both flavours
    a = orc.lit_ep_estimate(K, t, eps=0.01)
    b = orc.fast_ep_estimate(K, t, eps=0.01)
    assert a["sweeps"] == b["sweeps"] >= 1
    Ks = orc.lit_build_kernel_matrix(X[:7] + 0.05, th, X)
    Kss = orc.lit_build_kernel_matrix(X[:7] + 0.05, th)
    p1, fm1, fv1 = orc.lit_ep_classify(K, Ks, Kss, a["tau"], a["nu"], a["L"])
    p2, fm2, fv2 = orc.fast_ep_classify(K, Ks, Kss, b["tau"], b["nu"], b["L"])
    assert np.allclose(p1, p2, rtol=1e-9, atol=1e-12)
    assert np.all((p1 > 0) & (p1 < 1))


def test_ep_linebreak_quirk_switch():  # EpParameterEstimator.scala:91-92
    X, t, th = orc.make_c3(n=40, D=2, seed=5)
    K = orc.lit_build_kernel_matrix(X, th)
    q = orc.lit_ep_estimate(K, t, fixed_sweeps=2, keep_quirk=True)
    f = orc.lit_ep_estimate(K, t, fixed_sweeps=2, keep_quirk=False)
    extra = np.sum(0.5 * np.log(1 + q["tau"] / q["cav_tau"]) - np.log(np.diag(q["L"])))
    assert abs((f["logZ"] - q["logZ"]) - extra) <= 1e-10 * max(1.0, abs(extra))


def test_ep_hyperparameter_gradient_literal_vs_fast_and_quirk():  # MarginalLikelihoodEvaluator.scala:47-66
    X, t, th = orc.make_c3(n=90, D=3, seed=17)
    th = th.copy(); th[-1] = 0.25
    logz, g, o = orc.lit_ep_loglik_with_derivs(X, t, th, fixed_sweeps=3)
    K = orc.lit_build_kernel_matrix(X, th)
    gf = orc.fast_ep_loglik_derivs(X, th, K, o["tau"], o["nu"], o["L"])
    assert np.allclose(g, gf, rtol=1e-10, atol=1e-10 * np.abs(g).max())
    # as compiled rMatrix = b b^t, so every g_p is the quadratic form 1/2 b^t C_p b
    st = np.sqrt(o["tau"])
    b = o["nu"] - np.linalg.solve(np.diag(st) @ o["L"], np.linalg.solve(o["L"].T, st * (K @ o["nu"])))
    for p in range(5):
        Cm = orc.lit_build_der_matrix(p + 1, X, th)
        assert abs(g[p] - 0.5 * b @ Cm @ b) <= 1e-9 * max(abs(g[p]), 1e-9 * np.abs(g).max())
    # the discarded `- backSolve(...)` line changes the result when it is not discarded
    g2 = orc.lit_ep_loglik_derivs(X, th, K, o["tau"], o["nu"], o["L"], keep_quirk=False)
    assert np.abs(g2 - g).max() > 1e-3 * np.abs(g).max()

"""GPU parity for the GP-UCB inner loop (gp/optimization/GPOptimizer.scala:82-109) through gpk_gp_model_ucb, against the
oracle's statement-by-statement restatement, plus an end-to-end maximisation through the GPOptimizer mirror."""
import numpy as np
import pytest

import gp_algos_b200 as gp
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), np.abs(b).max() * 1e-6))


@pytest.mark.parametrize("n,D,m,k", [(40, 1, 3, 2.0), (150, 3, 5, 1.0), (300, 8, 17, 0.5)])
def test_ucb_objective_and_gradient_vs_literal_oracle(n, D, m, k):
    X, y, th = orc.make_c2(n=n, D=D, seed=40 + n)
    pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1])))
    model = pred.fit(X, None, y, th)
    pts = np.random.default_rng(n).uniform(-0.2, 1.2, size=(m, D))
    ucb, grad, mean, var = gp.ucb_with_gradient(model, pts, k)
    L, alpha = orc.lit_precompute(X, y, th)
    for i in range(m):
        u_o, g_o, m_o, s_o = orc.lit_ucb_with_grad(X, L, alpha, th, pts[i], k)
        assert abs(ucb[i] - u_o) <= 1e-9 * max(abs(u_o), 1e-3)
        assert abs(mean[i] - m_o) <= 1e-9 * max(abs(m_o), 1e-3) and abs(var[i] - s_o) <= 1e-9 * abs(s_o)
        assert close(grad[i], g_o)
    # one point per call (the reference's call pattern) gives the same numbers as the batch
    u1, g1, _, _ = gp.ucb_with_gradient(model, pts[2], k)
    assert abs(u1[0] - ucb[2]) <= 1e-12 * max(abs(ucb[2]), 1e-3) and np.allclose(g1[0], grad[2], rtol=1e-10, atol=1e-12)
    model.close()


def test_ucb_gradient_matches_finite_differences():
    X, y, th = orc.make_c2(n=200, D=4, seed=77)
    pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1])))
    model = pred.fit(X, None, y, th)
    x = np.array([0.4, 0.6, 0.1, 0.9])
    _, g, _, _ = gp.ucb_with_gradient(model, x, 1.5)
    eps = 1e-6
    P = np.vstack([x + eps * e for e in np.eye(4)] + [x - eps * e for e in np.eye(4)])
    u, _, _, _ = gp.ucb_with_gradient(model, P, 1.5)
    fd = (u[:4] - u[4:]) / (2 * eps)
    assert np.allclose(g[0], fd, rtol=1e-5, atol=1e-7)
    model.close()


def test_gp_optimizer_finds_the_maximum_of_a_smooth_function():
    # a positive bump: away from the data the zero-mean prior gives UCB ~ k sqrt(sf^2 + sn^2) = 1 < the bump's height, so
    # the (unbounded, like the reference) UCB search stays near the data
    f = lambda p: float(2.0 * np.exp(-((p[0] - 0.3) ** 2 + (p[1] + 0.2) ** 2) / 0.5))
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(1.0, [0.6, 0.6], 0.05))
    opt = gp.GPOptimizer(gp.GpPredictor(kf), None, gp.BreezeLbfgsOptimizer(10), seed=3)
    best, val = opt.maximize(f, gp.GPOInput(ranges=[(-1.0, 1.0), (-1.0, 1.0)], mParam=15, cParam=3, kParam=1.0))
    assert val > 1.5 and np.linalg.norm(best - np.array([0.3, -0.2])) < 0.4
    best2, val2 = opt.minimize(lambda p: -f(p), gp.GPOInput(ranges=[(-1.0, 1.0), (-1.0, 1.0)], mParam=3, cParam=2, kParam=1.0))
    assert val2 <= 0.0 and best2.shape == (2,)
    with pytest.raises(ValueError):
        opt.maximize(f, gp.GPOInput(ranges=[(-1.0, 1.0)], mParam=0, cParam=1, kParam=1.0))


# ---- GPOptimizer.scala:48-71: one evaluated point joins the training set per iteration --------------------------------
@pytest.mark.parametrize("n0,D,sigma_noise", [(30, 2, None), (381, 8, None), (254, 3, 0.02)])
def test_append_equals_refit(n0, D, sigma_noise):
    """The resident model after k appends equals the reference's refit on the enlarged set (preComputeComponents,
    GpPredictor.scala:104-124): alpha, posterior mean / variance and the log-likelihood, 1e-9.  n0 = 381 / 254 walk through a
    full 128-tile (n == N: the factor moves to a larger buffer) while appending."""
    k = 6
    X, y, th = orc.make_c2(n=n0 + k, D=D, seed=11 + n0)
    pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1])))
    model = pred.fit(X[:n0], sigma_noise, y[:n0], th)
    Xs = np.random.default_rng(n0).uniform(0, 1, size=(9, D))
    for j in range(k):
        model.append(X[n0 + j], y[n0 + j])
        n = n0 + j + 1
        assert model.n == n
        if j in (0, 2, k - 1):
            L, alpha = orc.lit_precompute(X[:n], y[:n], th, sigma_noise)
            a = model.alphaVec
            assert a.shape == (n,)
            assert np.all(np.abs(a - alpha) <= 1e-9 * np.abs(alpha).max())
            dist, _ = model.computePosterior(Xs)
            m_o, S_o, _ = orc.lit_compute_posterior(X[:n], Xs, L, alpha, th)
            assert np.all(np.abs(dist.mean - m_o) <= 1e-9 * np.maximum(np.abs(m_o), 1e-3))
            assert np.all(np.abs(np.diag(dist.sigma) - np.diag(S_o)) <= 1e-9 * np.abs(np.diag(S_o)))
            ll_o = orc.lit_loglik(alpha, L, y[:n])
            assert abs(model.logLikelihood - ll_o) <= 1e-9 * abs(ll_o)
            u, g, _, _ = gp.ucb_with_gradient(model, Xs[:2], 1.5)
            u_o, g_o, _, _ = orc.lit_ucb_with_grad(X[:n], L, alpha, th, Xs[1], 1.5)
            assert abs(u[1] - u_o) <= 1e-9 * max(abs(u_o), 1e-3) and close(g[1], g_o)
    model.close()


def test_append_not_positive_definite_leaves_the_model_unchanged():
    th = orc.pack_theta(1.0, [0.7, 0.7], 0.0)           # no noise: a repeated input makes the enlarged K exactly singular
    X, y = np.array([[0.3, 0.8]]), np.array([0.5])
    pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1])))
    model = pred.fit(X, None, y, th)
    a0 = model.alphaVec
    with pytest.raises(gp.NotPositiveDefiniteError) as ei:   # k = 1, l = 1, d^2 = 1 - 1 = 0: `cholesky` of the refit throws
        model.append(X[0], 0.3)
    assert ei.value.minor == 2 and model.n == 1
    assert np.array_equal(model.alphaVec, a0)
    with pytest.raises(ValueError):
        model.append(np.zeros(3), 0.0)
    model.append(np.array([1.1, -0.4]), 0.3)            # a fresh point still goes in
    L, alpha = orc.lit_precompute(np.vstack([X, [[1.1, -0.4]]]), np.append(y, 0.3), th)
    assert np.all(np.abs(model.alphaVec - alpha) <= 1e-9 * np.abs(alpha).max())
    model.close()


def test_append_rejects_a_sigma_noise_other_than_the_fitted_one():
    """include/gpk.h: the bordered row must carry the Option sigmaNoise the factor was built with; anything else is GPK_EINVAL
    (a C caller that forgets it would otherwise get an inconsistent factor without an error)."""
    import ctypes as C
    X, y, th = orc.make_c2(n=100, D=2, seed=9)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
    model = gp.GpPredictor(kf).fit(X[:90], 0.05, y[:90], th)
    h = model.handle
    x = np.ascontiguousarray(X[90]); d = C.c_double()
    from gp_algos_b200 import _lib
    for has_s, s in ((0, 0.0), (1, 0.06)):
        rc = h.lib.gpk_gp_model_append(h.h, model._m, _lib.ptr(x), float(y[90]), has_s, s, C.addressof(d))
        assert rc == _lib.GPK_EINVAL
    model.append(X[90], y[90])                           # the Python mirror replays the stored value
    L, alpha = orc.lit_precompute(X[:91], y[:91], th, 0.05)
    assert np.allclose(model.alphaVec, alpha, rtol=1e-9, atol=1e-12)
    model.close()

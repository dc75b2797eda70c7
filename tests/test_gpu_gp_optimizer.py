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

"""CPU: host logic of the GP-UKF mirror (gp_algos_b200/gp_ukf.py) -- the unscented transform against the oracle's restatement
of UnscentedKalmanFilter.scala:82-118, and the recursion on a linear-Gaussian model, where the UKF must reproduce the
Kalman filter exactly (a property the reference's own UKF tests rely on).  No GPU: the SSM functions here are plain NumPy."""
import numpy as np

import gp_algos_b200 as gp
from gp_algos_b200.gp_ukf import logGaussianDensity, nllOfHiddenData
from oracle import gp_oracle as orc


def test_unscented_transform_matches_oracle_restatement():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 3)); cov = A @ A.T + 0.5 * np.eye(3); mean = rng.standard_normal(3)
    f = lambda p: np.array([np.sin(p[0]) + p[1], p[1] * p[2], p[0] - p[2] ** 2, 1.0 + p[0]])
    ukf = gp.UnscentedKalmanFilter()
    for (a, b, k) in ((1.0, 0.0, 2.0), (0.5, 2.0, 1.0)):
        out = ukf.unscentedTransform(gp.GaussianDistribution(mean, cov), gp.UnscentedTransformParams(a, b, k),
                                     lambda pts: np.stack([f(p) for p in pts]))
        fm, fc, w, sp, tsp = orc.ukf_unscented_transform(mean, cov, a, b, k, f)
        assert np.array_equal(out.sigmaPoints, sp) and np.array_equal(out.transformedSigmaPoints, tsp)
        assert np.allclose(out.distribution.mean, fm, rtol=1e-14, atol=0) and np.allclose(out.distribution.sigma, fc, rtol=1e-13, atol=1e-15)
        assert out.weights == w
        assert abs(w[0] + 6 * w[2] - 1.0) < 1e-14          # mean weights sum to one (2d = 6 non-central points)


def test_ukf_on_a_linear_gaussian_model_is_the_kalman_filter():
    rng = np.random.default_rng(1)
    F = np.array([[0.9, 0.1], [-0.2, 0.8]]); H = np.array([[1.0, 0.5], [0.0, 1.0], [1.0, -1.0]])
    Q = 0.05 * np.eye(2); R = 0.1 * np.eye(3)
    T = 40
    z = np.zeros((2, T)); y = np.zeros((3, T)); z[:, 0] = [1.0, -1.0]
    for t in range(T):
        y[:, t] = H @ z[:, t] + rng.multivariate_normal(np.zeros(3), R)
        if t + 1 < T:
            z[:, t + 1] = F @ z[:, t] + rng.multivariate_normal(np.zeros(2), Q)
    model = gp.SsmModel(transitionFuncImpl=lambda u, pts, t: np.atleast_2d(pts) @ F.T, observationFuncImpl=lambda pts, t: np.atleast_2d(pts) @ H.T)
    inp = gp.UnscentedFilteringInput(model, y, None, np.array([1.0, -1.0]), 0.2 * np.eye(2), lambda ctx: Q, lambda ctx: R)
    out = gp.UnscentedKalmanFilter().inferHiddenState(inp, None, True)
    m, P, ll = np.array([1.0, -1.0]), 0.2 * np.eye(2), 0.0
    for t in range(1, T):                                   # textbook Kalman filter
        mp, Pp = F @ m, F @ P @ F.T + Q
        S = H @ Pp @ H.T + R
        K = Pp @ H.T @ np.linalg.inv(S)
        ll += logGaussianDensity(y[:, t], H @ mp, S)
        m, P = mp + K @ (y[:, t] - H @ mp), Pp - K @ S @ K.T
        # the reference forms the cross-covariance from two DIFFERENT sigma-point sets (UnscentedKalmanFilter.scala:50-60:
        # transformed points of the first transform against those of the second).  For a linear model that is
        # F L1 L2^t H^t with L1 = chol(P_{t-1}), L2 = chol(F P_{t-1} F^t + Q), not (F P F^t + Q) H^t -- reproduce the quirk
        Pprev = out.hiddenCovs[t - 1]
        L1, L2 = np.linalg.cholesky(Pprev), np.linalg.cholesky(F @ Pprev @ F.T + Q)
        Kq = (F @ L1 @ L2.T @ H.T) @ np.linalg.inv(H @ (F @ Pprev @ F.T + Q) @ H.T + R)
        mq = F @ out.hiddenMeans[:, t - 1]
        assert np.allclose(out.hiddenMeans[:, t], mq + Kq @ (y[:, t] - H @ mq), rtol=1e-9, atol=1e-12)
    assert out.logLikelihood is not None and np.isfinite(out.logLikelihood)
    assert np.isfinite(nllOfHiddenData(z, out.hiddenMeans, out.hiddenCovs))

"""CPU: the host UKF recursion of the test harness (tests/host_callers/ukf_host.py) -- the unscented transform against the oracle's restatement
of UnscentedKalmanFilter.scala:82-118, and the recursion on a linear-Gaussian model, where the UKF must reproduce the
Kalman filter exactly (a property the reference's own UKF tests rely on).  No GPU: the SSM functions here are plain NumPy."""
import numpy as np

import gp_algos_b200 as gp
from tests.host_callers.stats_utils import logGaussianDensity, nllOfHiddenData
from tests.host_callers import stats_utils as StatsUtils
from tests.host_callers.ukf_host import UnscentedKalmanFilter
from oracle import gp_oracle as orc


def test_unscented_transform_matches_oracle_restatement():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 3)); cov = A @ A.T + 0.5 * np.eye(3); mean = rng.standard_normal(3)
    f = lambda p: np.array([np.sin(p[0]) + p[1], p[1] * p[2], p[0] - p[2] ** 2, 1.0 + p[0]])
    ukf = UnscentedKalmanFilter()
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
    out = UnscentedKalmanFilter().inferHiddenState(inp, None, True)
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


# ---- the facts the reference's own src/test/scala/dynamicalsystems/filtering/UnscentedKalmanFilterTest.scala holds -----------
def test_reference_unscented_transform_facts():
    ukf = UnscentedKalmanFilter()
    # :44-56  1-d Gaussian through the identity
    tr = ukf.unscentedTransform(gp.GaussianDistribution(np.array([2.]), np.array([[2.]])), gp.UnscentedTransformParams(), lambda pts: pts)
    assert tr.distribution.dim == 1 and tr.distribution.sigma[0, 0] != 0.0
    assert tr.sigmaPoints.shape[0] == 3 and tr.transformedSigmaPoints.shape[0] == 3
    assert np.array_equal(tr.transformedSigmaPoints, tr.sigmaPoints)
    assert abs(tr.distribution.mean[0] - 2.0) < 1e-14 and abs(tr.distribution.sigma[0, 0] - 2.0) < 1e-14   # identity keeps the moments
    # :58-71  3-d Gaussian through the elementwise square
    g3 = gp.GaussianDistribution(np.array([1., 2., 3.]), np.eye(3))
    tr = ukf.unscentedTransform(g3, gp.UnscentedTransformParams(), lambda pts: pts * pts)
    assert tr.distribution.dim == 3 and tr.distribution.sigma.shape == (3, 3)
    assert tr.sigmaPoints.shape[0] == 7 and tr.transformedSigmaPoints.shape[0] == 7
    np.linalg.cholesky(tr.distribution.sigma)                                       # `cholesky(transformedDistr.sigma)` must not throw
    assert np.allclose(tr.distribution.mean, np.array([1., 4., 9.]) + 1.0, rtol=1e-13)   # E[x^2] = mu^2 + sigma^2, exact for the UT


def test_reference_filter_facts_on_the_sinusoidal_and_kitagawa_models():
    from tests.host_callers.ssm_examples import SinusoidalSsm, KitagawaSsm, generateSeries
    rng = np.random.default_rng(11)
    seq = 50
    cases = ((SinusoidalSsm(), gp.GaussianDistribution(np.array([0.]), np.array([[1.]]))),          # GaussianDistribution.standard (:29)
             (KitagawaSsm(), gp.GaussianDistribution(np.array([0.]), np.array([[0.5 * 0.5]]))))     # initHiddenStateDistrKit (:30-31)
    for model, init in cases:
        hidden, obs = generateSeries(model, seq, init, rng)
        assert hidden.shape == (1, seq) and obs.shape == (1, seq)
        inp = gp.UnscentedFilteringInput(model, obs, None, init.mean, init.sigma, lambda ctx, m=model: m.latentNoise,
                                         lambda ctx, m=model: m.obsNoise)
        for params in (gp.UnscentedTransformParams(alpha=1.0), gp.UnscentedTransformParams(2.012, 0.24, 0.4871)):   # :82-86, :97-99
            out = UnscentedKalmanFilter().inferHiddenState(inp, params, True)
            assert out.hiddenMeans.shape[1] == seq and out.hiddenMeans.shape[0] == hidden.shape[0] and len(out.hiddenCovs) == seq
            assert np.all(np.isfinite(out.hiddenMeans))
            # (the reference asserts shapes only: on the bimodal Kitagawa model the filter may lose track and the likelihood,
            #  log of an underflowed density as in StatsUtils.scala:57-59, is then -inf)
            assert model is cases[1][0] or np.isfinite(out.logLikelihood)
    # the sinusoidal model is easy: the filter tracks the hidden state better than the prior mean does
    model, init = cases[0]
    hidden, obs = generateSeries(model, 200, init, np.random.default_rng(12))
    inp = gp.UnscentedFilteringInput(model, obs, None, init.mean, init.sigma, lambda ctx: model.latentNoise, lambda ctx: model.obsNoise)
    out = UnscentedKalmanFilter().inferHiddenState(inp, gp.UnscentedTransformParams(alpha=1.0), True)
    assert StatsUtils.mse(out.hiddenMeans[:, 1:].T, hidden[:, 1:].T) < StatsUtils.mse(np.zeros_like(hidden[:, 1:].T), hidden[:, 1:].T)

"""GPU parity for the device-resident GP-UKF (csrc/gpk_ukf.cu; dynamicalsystems/filtering/GPUnscentedKalmanFilter.scala:63-147 on
top of UnscentedKalmanFilter.scala:24-118): the whole filter run is one `gpk_gpukf_filter` call -- against the oracle's per-point
restatement of the same recursion, and against the test harness's host recursion over the same device-resident GPs."""
import numpy as np
import pytest

import gp_algos_b200 as gp
from oracle import gp_oracle as orc
from tests.host_callers.ukf_host import HostRecursionGpUkf

pytestmark = pytest.mark.gpu


def _kernel():
    theta = orc.pack_theta(1.0, [1.0, 1.0], 0.1)
    return theta, gp.GaussianRbfKernel(gp.GaussianRbfParams(theta[0], theta[1:-1], theta[-1]))


@pytest.mark.parametrize("tMax,params", [(40, (1.0, 0.0, 2.0)), (90, (0.8, 2.0, 1.0))])
def test_device_gp_ukf_matches_oracle(tMax, params):
    hidden, obs = orc.make_ssm_series(tMax)
    theta, kf = _kernel()
    ukf = gp.GPUnscentedKalmanFilter(None, gp.GpPredictor(kf))
    inp = gp.UnscentedFilteringInput(None, obs, None, hidden[:, 0].copy(), 0.1 * np.eye(2), None, None)
    out = ukf.inferHiddenStateFromSamples(inp, hidden, gp.UnscentedTransformParams(*params), computeLL=True)
    m_o, c_o, ll_o = orc.gpukf_infer(hidden, obs, theta, hidden[:, 0], 0.1 * np.eye(2), *params)
    assert np.allclose(out.hiddenMeans, m_o, rtol=1e-9, atol=1e-9 * np.abs(m_o).max())
    for t in range(tMax):
        assert np.allclose(out.hiddenCovs[t], c_o[t], rtol=1e-8, atol=1e-9 * np.abs(c_o[t]).max())
    assert abs(out.logLikelihood - ll_o) <= 1e-9 * abs(ll_o)
    # the filter tracks the sampled trajectory (sanity of the whole pipeline, not a parity statement)
    assert np.abs(out.hiddenMeans - hidden).mean() < 0.2
    assert ukf._models == []                                       # device models released


def test_many_filters_in_one_call_equal_the_host_recursion_over_the_same_device_gps():
    """B = 5 series (different observations and priors) through ONE device call; each must equal the harness's host recursion
    (one device GP call per unscented transform) run on that series alone."""
    hidden, obs = orc.make_ssm_series(60)
    theta, kf = _kernel()
    pred = gp.GpPredictor(kf)
    ukf = gp.GPUnscentedKalmanFilter(None, pred).learn(obs, hidden)
    rng = np.random.default_rng(4)
    series = [obs] + [obs + 0.05 * rng.standard_normal(obs.shape) for _ in range(4)]
    m0 = [hidden[:, 0] + 0.1 * rng.standard_normal(2) for _ in series]
    c0 = []
    for _ in series:
        A = rng.standard_normal((2, 2)); c0.append(0.05 * (A @ A.T) + 0.05 * np.eye(2))   # non-diagonal priors: layout check
    params = gp.UnscentedTransformParams(0.9, 1.0, 1.5)
    outs = ukf.filter_many(series, m0, c0, params, computeLL=True)
    host = HostRecursionGpUkf(pred)
    model, q, r = host.learnNewSsmModelWithNoises(obs, hidden, False)
    for b, y in enumerate(series):
        inp = gp.UnscentedFilteringInput(model, y, None, m0[b], c0[b], q, r)
        ref = host.inferHiddenState(inp, params, True)
        assert np.allclose(outs[b].hiddenMeans, ref.hiddenMeans, rtol=1e-9, atol=1e-9 * np.abs(ref.hiddenMeans).max())
        for t in range(y.shape[1]):
            assert np.allclose(outs[b].hiddenCovs[t], ref.hiddenCovs[t], rtol=1e-8, atol=1e-10)
        assert abs(outs[b].logLikelihood - ref.logLikelihood) <= 1e-9 * abs(ref.logLikelihood)
    # batch consistency: a series filtered alone gives bit-identical results to the same series inside the batch
    alone = ukf.filter_many([series[3]], [m0[3]], [c0[3]], params, True)[0]
    assert np.array_equal(alone.hiddenMeans, outs[3].hiddenMeans) and alone.logLikelihood == outs[3].logLikelihood
    host.close(); ukf.close()


def test_a_covariance_that_is_not_positive_definite_is_reported():
    hidden, obs = orc.make_ssm_series(20)
    theta, kf = _kernel()
    ukf = gp.GPUnscentedKalmanFilter(None, gp.GpPredictor(kf)).learn(obs, hidden)
    with pytest.raises(gp.NotPositiveDefiniteError):                # breeze cholesky (UnscentedKalmanFilter.scala:85) would throw
        ukf.filter_many([obs], [hidden[:, 0]], [np.array([[1.0, 2.0], [2.0, 1.0]])])
    ukf.close()


def test_models_mean_var_equals_one_compute_posterior_per_model():
    from gp_algos_b200.gp_predictor import models_mean, models_mean_var
    hidden, obs = orc.make_ssm_series(30)
    theta, kf = _kernel()
    host = HostRecursionGpUkf(gp.GpPredictor(kf))
    model, q, r = host.learnNewSsmModelWithNoises(obs, hidden, False)
    pts = np.random.default_rng(0).standard_normal((5, 2))
    batch = model.transitionFuncImpl(None, pts, 1)
    single = np.stack([model.transitionFuncImpl(None, p, 1)[0] for p in pts])
    assert np.allclose(batch, single, rtol=1e-12, atol=1e-14)      # 2d+1 sigma points in one call == one call per point
    assert len(host._models) == 4
    mean, var = models_mean_var(host._models, pts)                  # every dimension, means and variances, ONE call
    assert np.array_equal(mean[:, :2], models_mean(host._models[:2], pts))
    for j, mdl in enumerate(host._models):
        dist = mdl.computePosterior(pts, full_cov=False, want_v=False)[0]
        assert np.array_equal(mean[:, j], dist.mean) and np.allclose(var[:, j], dist.sigma, rtol=1e-13, atol=1e-16)
        assert np.array_equal(mdl.mean(pts), dist.mean)
    host.close()
    assert host._models == []

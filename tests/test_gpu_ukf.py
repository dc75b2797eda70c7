"""GPU parity for the GP-UKF call pattern (dynamicalsystems/filtering/GPUnscentedKalmanFilter.scala:63-147 on top of
UnscentedKalmanFilter.scala:24-118): one device-resident GP per state / observation dimension, all sigma points of a transform
in one predict call per dimension -- against the oracle's per-point restatement of the same recursion."""
import numpy as np
import pytest

import gp_algos_b200 as gp
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tMax,params", [(40, (1.0, 0.0, 2.0)), (90, (0.8, 2.0, 1.0))])
def test_gp_ukf_matches_oracle(tMax, params):
    hidden, obs = orc.make_ssm_series(tMax)
    theta = orc.pack_theta(1.0, [1.0, 1.0], 0.1)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(theta[0], theta[1:-1], theta[-1]))
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


def test_gp_ukf_models_are_released_and_predictions_batch_consistently():
    hidden, obs = orc.make_ssm_series(30)
    theta = orc.pack_theta(1.0, [1.0, 1.0], 0.1)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(theta[0], theta[1:-1], theta[-1]))
    ukf = gp.GPUnscentedKalmanFilter(None, gp.GpPredictor(kf))
    model, q, r = ukf.learnNewSsmModelWithNoises(obs, hidden, False)
    pts = np.random.default_rng(0).standard_normal((5, 2))
    batch = model.transitionFuncImpl(None, pts, 1)
    single = np.stack([model.transitionFuncImpl(None, p, 1)[0] for p in pts])
    assert np.allclose(batch, single, rtol=1e-12, atol=1e-14)      # 2d+1 sigma points in one call == one call per point
    assert len(ukf._models) == 4
    from gp_algos_b200.gp_predictor import models_mean
    allm = models_mean(ukf._models[:2], pts)                        # every dimension in one call == one call per model
    assert np.array_equal(allm, np.stack([mdl.mean(pts) for mdl in ukf._models[:2]], axis=1))
    for mdl in ukf._models:                                         # mean-only path == mean of the full posterior call
        assert np.array_equal(mdl.mean(pts), mdl.computePosterior(pts, full_cov=False, want_v=False)[0].mean)
    ukf.close()
    assert ukf._models == []

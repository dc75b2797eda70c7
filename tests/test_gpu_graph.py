"""CUDA-graph replay of GpPredictor.logLikelihoodWithDerivatives (gp/regression/GpPredictor.scala:60-80 called over and over by
obtainOptimalHyperParams, GpPredictor.scala:126-142): the second call with one signature is captured, later calls are replays
whose hyper-parameters travel through device memory.  Replays must be bit-identical to eager launches, agree with the oracle
to 1e-9, report a failed factorisation like an eager call and survive a change of the problem in between."""
import numpy as np
import pytest

import gp_algos_b200 as gp
from gp_algos_b200 import _lib, synthetic
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _pred(th, h):
    return gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1])), handle=h)


def _thetas(theta0, k, seed):
    rng = np.random.default_rng(seed)
    return [theta0 * 10 ** rng.uniform(-0.15, 0.15, size=theta0.size) for _ in range(k)]


def assert_grad(g, go):
    floor = RTOL * np.abs(go).max()
    assert np.all(np.abs(g - go) <= RTOL * np.maximum(np.abs(go), floor)), (g, go)


@pytest.mark.parametrize("n", [300, 1100, 4200])   # recursive path (one and several levels) and the look-ahead driver (N >= 4096)
def test_replay_equals_eager_and_oracle(n):
    X, y, theta0 = synthetic.make_c2(n=n, D=8)
    hg, he = _lib.Handle(0), _lib.Handle(0)
    hg.set_graph_mode(True)
    he.set_graph_mode(False)
    inp = gp.PredictionTrainingInput(X, None, y)
    pg, pe = _pred(theta0, hg), _pred(theta0, he)
    l_eager = l_graph = None
    for i, th in enumerate(_thetas(theta0, 5, n)):
        c0, c1 = hg.launch_count(), he.launch_count()
        ll_g, g_g = pg.logLikelihoodWithDerivatives(inp, th, 10)
        ll_e, g_e = pe.logLikelihoodWithDerivatives(inp, th, 10)
        l_graph, l_eager = hg.launch_count() - c0, he.launch_count() - c1
        assert ll_g == ll_e and np.array_equal(g_g, g_e), (i, ll_g, ll_e)
        if i in (0, 1, 4):   # eager first call, the capturing call, a later replay
            ll_o, g_o = orc.fast_loglik_with_derivs(X, y, th, None, 10)
            assert abs(ll_g - ll_o) <= RTOL * abs(ll_o)
            assert_grad(g_g, g_o)
    assert l_graph == l_eager + 1      # the replay's kernel nodes + the parameter store are what gpk_launch_count reports
    # sigmaNoise (GpPredictor.scala:116, added un-squared) rides in the same device record
    inp_s = gp.PredictionTrainingInput(X, 0.05, y)
    ll_g, g_g = pg.logLikelihoodWithDerivatives(inp_s, theta0, 10)
    ll_e, g_e = pe.logLikelihoodWithDerivatives(inp_s, theta0, 10)
    assert ll_g == ll_e and np.array_equal(g_g, g_e)
    assert ll_g != pg.logLikelihoodWithDerivatives(inp, theta0, 10)[0]
    hg.close(); he.close()


def test_replay_reports_not_positive_definite_and_recovers():
    X, y, theta0 = synthetic.make_c2(n=400, D=8)
    X[:] = 0.25                                   # identical inputs: K = sf^2 11^t + sn^2 I, exactly singular without noise
    h = _lib.Handle(0)
    p = _pred(theta0, h)
    inp = gp.PredictionTrainingInput(X, None, y)
    for _ in range(3):                            # eager, capture, replay
        p.logLikelihoodWithDerivatives(inp, theta0, 10)
    bad = theta0.copy(); bad[-1] = 0.0
    with pytest.raises(_lib.NotPositiveDefiniteError) as ei:
        p.logLikelihoodWithDerivatives(inp, bad, 10)
    assert ei.value.minor == 2             # second pivot = 1 - 1*1 = 0 exactly
    ll, g = p.logLikelihoodWithDerivatives(inp, theta0, 10)
    ll_o, g_o = orc.fast_loglik_with_derivs(X, y, theta0, None, 10)
    assert abs(ll - ll_o) <= RTOL * abs(ll_o)
    assert_grad(g, g_o)
    h.close()


def test_signature_changes_between_calls():
    """Alternating problems (other n, other D, other nParams) drop the cached graph; a bigger problem regrows the workspace."""
    h = _lib.Handle(0)
    cases = []
    for n, D, P in ((200, 3, 5), (520, 8, 10), (200, 3, 2), (900, 8, 10)):
        X, y, th = synthetic.make_c2(n=n, D=D)
        cases.append((X, y, th, P, orc.fast_loglik_with_derivs(X, y, th, None, P)))
    for rep in range(3):
        for X, y, th, P, (ll_o, g_o) in cases:
            p = _pred(th, h)
            for _ in range(rep + 1):
                ll, g = p.logLikelihoodWithDerivatives(gp.PredictionTrainingInput(X, None, y), th, P)
                assert abs(ll - ll_o) <= RTOL * abs(ll_o)
                assert_grad(g, g_o)
    h.set_graph_mode(False)
    X, y, th, P, (ll_o, g_o) = cases[1]
    ll, g = _pred(th, h).logLikelihoodWithDerivatives(gp.PredictionTrainingInput(X, None, y), th, P)
    assert abs(ll - ll_o) <= RTOL * abs(ll_o)
    h.close()


def test_other_entry_points_between_replays():
    """fit / predict on the same handle reuse the workspace arenas between two replays of a cached evaluation graph."""
    X, y, theta0 = synthetic.make_c2(n=640, D=8)
    h = _lib.Handle(0)
    p = _pred(theta0, h)
    inp = gp.PredictionTrainingInput(X, None, y)
    ref = [p.logLikelihoodWithDerivatives(inp, theta0, 10) for _ in range(3)][-1]
    dist, _ = p.predict(gp.PredictionInput(X, X[:9] + 0.01, None, y), theta0)
    m_o, S_o, _ = orc.fast_predict(X, y, X[:9] + 0.01, theta0)
    assert np.allclose(dist.mean, m_o, rtol=1e-9, atol=1e-12)
    again = p.logLikelihoodWithDerivatives(inp, theta0, 10)
    assert again[0] == ref[0] and np.array_equal(again[1], ref[1])
    h.close()


def test_ep_sweep_replay_equals_eager():
    """EpParameterEstimator.estimateSiteParams (EpParameterEstimator.scala:29-69): sweep 1 eager, sweep 2 captured, later sweeps
    replayed (graph mode 2) -- same site parameters, factor and log Z bit for bit as on a handle that never captures, and the
    oracle's to 1e-9.  n = 4200 goes through the look-ahead factorisation inside the sweep."""
    for n, sweeps in ((300, 5), (4200, 4)):
        X, t, th = synthetic.make_c3(n=n, D=4)
        kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
        hg, he = _lib.Handle(0), _lib.Handle(0)
        hg.set_graph_mode(2)
        he.set_graph_mode(0)
        K = gp.MatrixUtils.buildKernelMatrix(kf, X, handle=he)
        sg, Lg = gp.EpParameterEstimator(K, t, gp.FixedSweeps(sweeps), hg).estimateSiteParams
        se, Le = gp.EpParameterEstimator(K, t, gp.FixedSweeps(sweeps), he).estimateSiteParams
        assert np.array_equal(sg.tauSiteParams, se.tauSiteParams) and np.array_equal(sg.niSiteParams, se.niSiteParams)
        assert sg.marginalLogLikelihood == se.marginalLogLikelihood and np.array_equal(Lg, Le)
        # a second run on the same handle starts from the cached graph
        sg2, _ = gp.EpParameterEstimator(K, t, gp.FixedSweeps(sweeps), hg).estimateSiteParams
        assert np.array_equal(sg2.tauSiteParams, se.tauSiteParams) and sg2.marginalLogLikelihood == se.marginalLogLikelihood
        if n <= 400:
            o = orc.lit_ep_estimate(K, t, fixed_sweeps=sweeps)
            assert np.all(np.abs(sg.tauSiteParams - o["tau"]) <= RTOL * np.abs(o["tau"]).max())
            assert abs(sg.marginalLogLikelihood - o["logZ"]) <= RTOL * abs(o["logZ"])
        hg.close(); he.close()

"""CPU tests for the Co2Kernel row (gp/regression/Co2Prediction.scala:16-186): the oracle restatement against finite
differences, a high-precision arbiter and the committed golden / soft fixtures, and the host mirror against the oracle."""
import os

import numpy as np
import pytest

import gp_algos_b200 as gp
from oracle import gp_oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HP = orc.CO2_SHIPPED_HP


def test_oracle_kernel_value_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    h = [mp.mpf(float(v)) for v in HP]
    for x1, x2 in ((1958.25, 1958.25), (1961.25, 1973.5833333333333), (2003.0, 1958.1666666666667), (1970.0, 1970.0833333333333)):
        d = mp.mpf(x1) - mp.mpf(x2)
        k = (h[0] ** 2 * mp.exp(-d * d / (2 * h[1] ** 2)) + h[2] ** 2 * mp.exp(-d * d / (2 * h[3] ** 2) - 2 * mp.sin(mp.pi * d) ** 2 / h[4] ** 2)
             + h[5] ** 2 * (1 + d * d / (2 * h[7] * h[6] ** 2)) ** (-h[7]) + h[8] ** 2 * mp.exp(-d * d / (2 * h[9] ** 2)))
        v = orc.co2_k(x1, x2, HP, False)
        assert abs(v - float(k)) <= 1e-13 * abs(float(k))     # sin(pi * d) at |d| ~ 45 carries ~1e-14 absolute argument rounding
        assert orc.co2_k(x1, x2, HP, True) == v + HP[10] * HP[10] or x1 != x2
    assert orc.co2_k(3.0, 3.0, HP, True) == ((((HP[0] ** 2 + HP[2] ** 2) + HP[5] ** 2) + HP[8] ** 2) + HP[10] ** 2)


def test_oracle_derivatives_against_finite_differences():
    rng = np.random.default_rng(0)
    hp = HP * 10 ** rng.uniform(-0.1, 0.1, size=11)
    for x1, x2 in ((1961.25, 1961.9), (1999.0, 1980.3)):
        for p in range(1, 12):
            for same in (False, True):
                xx2 = x1 if same else x2
                step = 1e-6 * abs(hp[p - 1])
                up, dn = hp.copy(), hp.copy()
                up[p - 1] += step
                dn[p - 1] -= step
                fd = (orc.co2_k(x1, xx2, up, same) - orc.co2_k(x1, xx2, dn, same)) / (2 * step)
                an = orc.co2_dk(p, x1, xx2, hp, same)
                assert abs(an - fd) <= 2e-5 * abs(an) + 2e-6, (p, same, an, fd)   # k ~ 4e3: the difference quotient carries ~4e-7 of rounding
    with pytest.raises(LookupError):        # scala.MatchError (Co2Prediction.scala:70-80)
        orc.co2_dk(12, 1.0, 2.0, hp, False)


def test_oracle_reproduces_the_golden_fixture_and_the_shipped_results():
    g = np.load(os.path.join(G, "co2_maunaloa.npz"))
    train, test, hp = g["train"], g["test"], g["theta"]
    assert train.shape == (424, 2) and test.shape == (183, 2)
    whole = np.vstack([train, test])
    # the time axis written by the reference (co2PredResults.txt, column 0) pins co2DataToYearWithValue exactly
    assert np.array_equal(whole[:, 0], g["ref_year"])
    assert train[0, 0] == 1958 + 2 / 12. and train[0, 1] == 315.70       # first valid month of maunaLoa.txt (March 1958)
    with orc.co2_kernel():
        ll, grad = orc.lit_loglik_with_derivs(train[:, :1], train[:, 1], hp, None)
        mean, sigma, ll_p = orc.lit_predict(train[:, :1], train[:, 1], whole[:, :1], hp, None)
    assert ll == float(g["ll"]) == ll_p and np.array_equal(grad, g["grad"])
    assert np.array_equal(mean, g["mean"]) and np.array_equal(np.diag(sigma), g["var"])
    # soft: the shipped file was produced after the reference's own L-BFGS run from these hyper-parameters
    nt = train.shape[0]
    assert np.abs(mean[:nt] - g["ref_mean"][:nt]).max() < 0.2           # ppm, on values of 313..357
    assert np.abs(np.sqrt(np.diag(sigma))[:nt] - g["ref_std"][:nt]).max() < 0.02
    assert np.abs(mean[nt:] - g["ref_mean"][nt:]).max() < 4 * g["ref_std"][nt:].max()   # extrapolation: inside the reference's own band
    # the family switch is scoped
    assert orc._L().orc_get_kernel() == orc.KERNEL_SE_ARD


def test_gradient_of_the_likelihood_against_finite_differences():
    X, y, hp = orc.make_co2_like(n=60, seed=3)
    with orc.co2_kernel():
        ll, g = orc.lit_loglik_with_derivs(X, y, hp, None)
        assert g.shape == (11,)
        for p in (0, 1, 4, 7, 10):
            step = 1e-6 * hp[p]
            up, dn = hp.copy(), hp.copy()
            up[p] += step
            dn[p] -= step
            fd = (orc.lit_loglik_with_derivs(X, y, up, None, 0)[0] - orc.lit_loglik_with_derivs(X, y, dn, None, 0)[0]) / (2 * step)
            assert abs(g[p] - fd) <= 1e-4 * max(abs(g[p]), 1.0), (p, g[p], fd)


def test_host_mirror_matches_the_oracle():
    kf = gp.Co2Kernel(gp.Co2HyperParams(HP))
    assert kf.hyperParametersNum == 11 and kf.family == 1
    assert kf.hyperParams.getAtPosition(1) == 60. and kf.hyperParams.getAtPosition(11) == 0.19     # 1-based (Co2Prediction.scala:23)
    with pytest.raises(IndexError):
        kf.hyperParams.getAtPosition(12)
    for x1, x2 in ((1958.25, 1958.25), (1961.25, 1973.5833333333333)):
        for same in (False, True):
            v, o = kf.apply([x1], [x2], same), orc.co2_k(x1, x2, HP, same)
            assert abs(v - o) <= 4e-16 * abs(o)
            for p in range(1, 12):
                d, do = kf.derAfterHyperParam(p)([x1], [x2], same), orc.co2_dk(p, x1, x2, HP, same)
                assert abs(d - do) <= 1e-15 * abs(do)
    with pytest.raises(LookupError):
        kf.derAfterHyperParam(12)
    with pytest.raises(ValueError):          # require(obj1.length == 1 && obj2.length == 1) Co2Prediction.scala:39
        kf.apply([1.0, 2.0], [1.0, 2.0], False)
    with pytest.raises(NotImplementedError):  # `???` at Co2Prediction.scala:62-64
        kf.gradient(True)
    k2 = kf.changeHyperParams(HP * 2)
    assert isinstance(k2, gp.Co2Kernel) and k2.hyperParams.getAtPosition(2) == 140.
    assert kf.hyperParams.fromDenseVector(HP[:10]).toDenseVector.shape == (10,)   # no length requirement in the reference


def test_co2_data_to_year_with_value():
    raw = np.array([[1958., -99.99, 315.7, 317.4, -99.99], [1959., 315.6, 316.4, 316.7, 316.0], [1960., 316.4, -99.99, 317.6, 316.9]])
    tr, te = gp.co2DataToYearWithValue(raw, 0.5)
    o_tr, o_te = orc.co2_data_to_year_with_value(raw, 0.5)
    assert np.array_equal(tr, o_tr) and np.array_equal(te, o_te)
    whole = np.vstack([tr, te])
    assert whole.shape == (7, 2) and tr.shape[0] == 3                     # the last column (annual mean) is never read
    assert whole[0, 0] == 1958 + 1 / 12. and whole[0, 1] == 315.7 and whole[-1, 0] == 1960 + 2 / 12.
    with pytest.raises(ValueError):
        gp.co2DataToYearWithValue(raw, 1.5)

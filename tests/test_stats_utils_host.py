"""CPU: host mirror of utils/StatsUtils.scala against the facts the reference's own src/test/scala/utils/StatsUtilsTest.scala
holds (precision 1e-4 there; exact here) and against the oracle's pnorm / dnorm."""
import math

import numpy as np
import pytest

from tests.host_callers import stats_utils as su
from gp_algos_b200.gp_predictor import GaussianDistribution
from oracle import gp_oracle as orc


def test_gaussian_density_at_the_mean():                       # StatsUtilsTest.scala:17-24 (prints the value)
    assert abs(su.gaussianDensity([0., 0.], [0., 0.], np.eye(2)) - 1 / (2 * math.pi)) < 1e-15
    assert abs(su.logGaussianDensity([1., -2.], [0., 0.], np.diag([4., 9.]))
               - (-0.5 * (0.25 + 4 / 9.) - math.log(6.0) - math.log(2 * math.pi))) < 1e-14


def test_mean_and_var_of_data():                               # StatsUtilsTest.scala:28-37
    mean, sigma = su.meanAndVarOfData(np.array([[1., 5.], [0.5, 3.], [0.6, 4.]]))
    assert np.allclose(mean, [0.7, 4.], atol=1e-4) and np.allclose(mean, [2.1 / 3, 4.0], rtol=1e-15)
    np.linalg.cholesky(sigma)                                  # `cholesky(sigma)` must not throw
    assert np.allclose(sigma, np.cov(np.array([[1., 5.], [0.5, 3.], [0.6, 4.]]).T, bias=True), rtol=1e-14)


def test_mean_squared_error():                                 # StatsUtilsTest.scala:40-47
    est = np.array([[1., 2.], [3., 4.], [5., 6.]])
    assert abs(su.mse(est, est + 0.1) - 0.02) < 1e-4
    assert abs(su.mse(est, est + 0.1, horSample=False) - 0.03) < 1e-12
    with pytest.raises(ValueError):
        su.mse(est, est[:2])


def test_pnorm_dnorm_match_the_oracle():
    lib = orc._L()
    for z in (-6.0, -1.3, 0.0, 0.4, 2.5, 7.0):
        assert su.pnorm(z) == lib.orc_pnorm(orc.C.c_double(z)) and su.dnorm(z) == lib.orc_dnorm(orc.C.c_double(z))


def test_sampler_and_nll():
    g = GaussianDistribution(np.array([1.0, -2.0]), np.array([[2.0, 0.3], [0.3, 0.5]]))
    s = su.NormalDistributionSampler(g, np.random.default_rng(3))
    draws = np.array([s.sample for _ in range(4000)])
    m, c = su.meanAndVarOfData(draws)
    assert np.allclose(m, g.mean, atol=0.08) and np.allclose(c, g.sigma, atol=0.12)
    with pytest.raises(ValueError):
        su.NormalDistributionSampler(GaussianDistribution(np.zeros(2), np.eye(3)))
    X = draws[:5].T                                            # d x T
    a = su.nllOfHiddenData(X, [g] * 5, None)
    b = su.nllOfHiddenData(X, np.tile(g.mean[:, None], (1, 5)), [g.sigma] * 5)
    assert abs(a - b) < 1e-12 and a == pytest.approx(-sum(su.logGaussianDensity(X[:, i], g.mean, g.sigma) for i in range(5)))
    assert su.standard().dim == 1

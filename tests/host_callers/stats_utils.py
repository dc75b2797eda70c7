"""TEST HARNESS (not product code; SURVEY.md 2 marks StatsUtils out of scope).  Host mirror of utils/StatsUtils.scala: the small statistics helpers the callers of the hot path use (GPOptimizer's restart
sampler, the GP-UKF scoring, Co2PredictionExecutor's error measure).  O(d^3) arithmetic on d x d matrices, d = state / input
dimension -- host code in the reference and here; nothing n x n passes through this module.

breeze.stats.distributions.Gaussian(0,1).cdf/.pdf and commons-math3's MultivariateNormalDistribution are un-vendored third-party
code: pnorm = erfc(-x/sqrt 2)/2, dnorm = exp(-x^2/2)/sqrt(2 pi), and the N(mean, covs) sampler through NumPy's SVD route
(commons-math uses an eigen-decomposition; the distribution is the same, the stream of samples is not)."""
from __future__ import annotations

import math

import numpy as np

from gp_algos_b200.gp_predictor import GaussianDistribution


def dnorm(x: float) -> float:                                  # StatsUtils.scala:15
    return math.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


def pnorm(x: float) -> float:                                  # StatsUtils.scala:17
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def standard() -> GaussianDistribution:                        # GaussianDistribution.standard, StatsUtils.scala:23-25
    return GaussianDistribution(np.array([0.0]), np.array([[1.0]]))


class NormalDistributionSampler:
    """StatsUtils.scala:27-46."""

    def __init__(self, normalDistr: GaussianDistribution, rng=None):
        self.mean = np.asarray(normalDistr.mean, dtype=np.float64)
        self.covs = np.asarray(normalDistr.sigma, dtype=np.float64)
        if not (self.mean.shape[0] == self.covs.shape[0] and self.covs.shape[0] == self.covs.shape[1]):
            raise ValueError("requirement failed")                                   # :33-34
        self.rng = rng if rng is not None else np.random.default_rng()

    @property
    def sample(self) -> np.ndarray:
        return self.rng.multivariate_normal(self.mean, self.covs, method="svd")

    @staticmethod
    def sampleFrom(gaussianDistribution: GaussianDistribution, rng=None) -> np.ndarray:   # object NormalDistributionSampler.sample, :129-131
        return NormalDistributionSampler(gaussianDistribution, rng).sample


def gaussianDensity(at, means, covs) -> float:                 # StatsUtils.scala:48-55 (MultivariateNormalDistribution.density)
    at, means, covs = (np.asarray(v, dtype=np.float64) for v in (at, means, covs))
    covs = np.atleast_2d(covs)
    d = at - means
    sign, logdet = np.linalg.slogdet(covs)
    quad = float(d @ np.linalg.solve(covs, d))
    return float((2 * math.pi) ** (-0.5 * len(d)) * (sign * math.exp(logdet)) ** -0.5 * math.exp(-0.5 * quad))


def logGaussianDensity(at, means, covs) -> float:              # StatsUtils.scala:57-59: log of the density (-inf once it underflows)
    dens = gaussianDensity(at, means, covs)
    return math.log(dens) if dens > 0 else -math.inf


def meanAndVarOfData(data):
    """StatsUtils.scala:61-73; data(i, ::) is the i-th sample; the covariance is divided by the number of rows (not rows - 1)."""
    data = np.asarray(data, dtype=np.float64)
    mean = data.sum(axis=0) / float(data.shape[0])
    diff = data - mean
    return mean, diff.T @ diff / float(data.shape[0])


def mse(estimate, trueValues, horSample: bool = True) -> float:
    """StatsUtils.scala:75-90: sum of squared differences over ALL entries divided by the number of samples (rows when
    horSample, else columns)."""
    e, t = np.asarray(estimate, dtype=np.float64), np.asarray(trueValues, dtype=np.float64)
    if e.shape != t.shape:
        raise ValueError("requirement failed: Both matrices must have identical dimensions")
    return float(((e - t) ** 2).sum() / (e.shape[0] if horSample else e.shape[1]))


def nllOfHiddenData(trueHiddenStates, hiddenMeans, hiddenCovs) -> float:
    """StatsUtils.scala:92-114 (both overloads: a list of GaussianDistribution, or means d x T with a list of covariances)."""
    X = np.asarray(trueHiddenStates, dtype=np.float64)
    if hiddenCovs is None:                                     # first overload: hiddenMeans is Array[GaussianDistribution]
        dists = list(hiddenMeans)
        if X.shape[1] != len(dists):
            raise ValueError("requirement failed: Hidden states number must be equal to inferred states number")
        return float(-sum(logGaussianDensity(X[:, i], dists[i].mean, dists[i].sigma) for i in range(X.shape[1])))
    M = np.asarray(hiddenMeans, dtype=np.float64)
    if M.shape[1] != len(hiddenCovs):
        raise ValueError("requirement failed: Number of hidden means should be equal to number of hidden covariances")
    return float(-sum(logGaussianDensity(X[:, i], M[:, i], hiddenCovs[i]) for i in range(X.shape[1])))

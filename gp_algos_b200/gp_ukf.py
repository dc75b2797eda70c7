"""Host mirror of dynamicalsystems/filtering/{UnscentedKalmanFilter, GPUnscentedKalmanFilter}.scala (GP-UKF).

Split of work: the filter recursion itself is d x d algebra (Cholesky of the state covariance,
sigma points, Kalman gain) -- host control flow, exactly as in the reference.  What touches n x n data is the GP part: one GP
per state / observation dimension fitted on the sampled trajectory (GPUnscentedKalmanFilter.scala:123-136) and, per time step,
(2d+1) (d_state + d_obs) posterior means plus (d_state + d_obs) posterior variances (:77-88,138-147).  Here every GP is a
device-resident model (`gpk_gp_model_fit`: X, L^-1, alpha stay in HBM) and each unscented transform evaluates all of its 2d+1
sigma points in ONE `gpk_gp_model_predict` call per output dimension, instead of the reference's one `computePosterior`
(an O(n^2) scalar triangular solve) per sigma point and dimension.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib
from .gp_predictor import FittedGp, GaussianDistribution, GpPredictor, models_mean


@dataclass(frozen=True)
class UnscentedTransformParams:
    """UnscentedKalmanFilter.scala:171-187 (defaults alpha = 1, beta = 0, kappa = 2)."""
    alpha: float = 1.0
    beta: float = 0.0
    kappa: float = 2.0

    @property
    def toVector(self):
        return np.array([self.alpha, self.beta, self.kappa])

    @staticmethod
    def fromVector(arr):
        return UnscentedTransformParams(float(arr[0]), float(arr[1]), float(arr[2]))


@dataclass
class UnscentedTransformOutput:
    """UnscentedKalmanFilter.scala:190-192; weights = (w_0_m, w_0_c, w_i_c)."""
    distribution: GaussianDistribution
    weights: tuple
    sigmaPoints: np.ndarray
    transformedSigmaPoints: np.ndarray


@dataclass
class UkfInferenceContext:
    """UnscentedKalmanFilter.scala:213-216."""
    iteration: int
    hiddenMeans: np.ndarray
    hiddenCovs: list
    firstTransformFromIteration: Optional[UnscentedTransformOutput]
    secondTransformFromIteration: Optional[UnscentedTransformOutput]


@dataclass
class SsmModel:
    """SsmModel.scala:12-17, with the two mapping functions taking a MATRIX of points (one per row) so that a whole set of
    sigma points is one device call: transitionFuncImpl(u_t, points, t) -> (m, d_state); observationFuncImpl(points, t)."""
    transitionFuncImpl: Callable
    observationFuncImpl: Callable
    latentNoise: Optional[np.ndarray] = None
    obsNoise: Optional[np.ndarray] = None


@dataclass
class UnscentedFilteringInput:
    """UnscentedKalmanFilter.scala:194-198."""
    ssmModel: SsmModel
    observations: np.ndarray            # d_obs x tMax
    u: Optional[np.ndarray]
    initMean: np.ndarray
    initCov: np.ndarray
    qNoise: Callable
    rNoise: Callable


@dataclass
class FilteringOutput:
    """KalmanFilter.scala:138-139."""
    hiddenMeans: np.ndarray
    hiddenCovs: list
    logLikelihood: Optional[float]


def logGaussianDensity(at, means, covs) -> float:
    """StatsUtils.scala:45-56 (commons-math MultivariateNormalDistribution.density, then log)."""
    at = np.asarray(at, dtype=np.float64); means = np.asarray(means, dtype=np.float64)
    d = len(means)
    diff = at - means
    sign, logdet = np.linalg.slogdet(covs)
    quad = float(diff @ np.linalg.solve(covs, diff))
    dens = (2 * math.pi) ** (-0.5 * d) * (sign * math.exp(logdet)) ** -0.5 * math.exp(-0.5 * quad)
    return math.log(dens) if dens > 0 else -math.inf


def nllOfHiddenData(trueHiddenStates, hiddenMeans, hiddenCovs) -> float:
    """StatsUtils.scala:99-109."""
    return -sum(logGaussianDensity(trueHiddenStates[:, i], hiddenMeans[:, i], hiddenCovs[i]) for i in range(trueHiddenStates.shape[1]))


class UnscentedKalmanFilter:
    """UnscentedKalmanFilter.scala:13 (the gpOptimizer argument is only used by the inferWithUkfOptim* drivers)."""

    def __init__(self, gpOptimizer=None):
        self.gpOptimizer = gpOptimizer

    def unscentedTransform(self, normalDistribution: GaussianDistribution, params: UnscentedTransformParams, func) -> UnscentedTransformOutput:
        """UnscentedKalmanFilter.scala:82-118; `func` maps the (2d+1, d) sigma-point matrix to (2d+1, d_out) in one call."""
        mean = np.asarray(normalDistribution.mean, dtype=np.float64)
        cov = np.asarray(normalDistribution.sigma, dtype=np.float64)
        d = len(mean)
        L = np.linalg.cholesky(cov)                                                   # :85 breeze cholesky (d x d, host)
        lam = params.alpha * params.alpha * (d + params.kappa) - d
        sp = np.zeros((2 * d + 1, d))
        sp[0] = mean
        for col in range(d):
            sqrtCoeff = L[:, col] * math.sqrt(d + lam)
            sp[col + 1] = mean + sqrtCoeff
            sp[col + 1 + d] = mean - sqrtCoeff
        w_0_m = lam / (d + lam)
        w_0_c = (lam / (d + lam)) + (1 - params.alpha * params.alpha + params.beta)
        w_i_c = 1 / (2 * (d + lam))
        tsp = np.asarray(func(sp), dtype=np.float64)
        finalMean = tsp[0] * w_0_m
        for i in range(1, 2 * d + 1):                                                 # :100-105 (w_i_c also weights the mean, 8(c)(7))
            finalMean = finalMean + tsp[i] * w_i_c
        diff = tsp[0] - finalMean
        finalCov = np.outer(diff, diff) * w_0_c
        for i in range(1, 2 * d + 1):
            diff = tsp[i] - finalMean
            finalCov = finalCov + np.outer(diff, diff) * w_i_c
        return UnscentedTransformOutput(GaussianDistribution(finalMean, finalCov), (w_0_m, w_0_c, w_i_c), sp, tsp)

    def inferHiddenState(self, input: UnscentedFilteringInput, params: Optional[UnscentedTransformParams] = None,
                         computeLL: bool = True) -> FilteringOutput:
        """UnscentedKalmanFilter.scala:24-80."""
        up = params or UnscentedTransformParams()
        y = np.asarray(input.observations, dtype=np.float64)
        tMax, hid = y.shape[1], len(input.initMean)
        ll = 0.0 if computeLL else None
        u = input.u if input.u is not None else np.zeros((1, tMax))
        hiddenMeans = np.zeros((hid, tMax))
        hiddenCovs: List[Optional[np.ndarray]] = [None] * tMax
        hiddenMeans[:, 0] = input.initMean
        hiddenCovs[0] = np.asarray(input.initCov, dtype=np.float64)
        ctx = UkfInferenceContext(0, hiddenMeans, hiddenCovs, None, None)
        for t in range(1, tMax):
            u_t = u[:, t]
            prev = GaussianDistribution(hiddenMeans[:, t - 1].copy(), hiddenCovs[t - 1])
            first = self.unscentedTransform(prev, up, lambda pts: input.ssmModel.transitionFuncImpl(u_t, pts, t))
            qNoise = input.qNoise(replace(ctx, iteration=t, firstTransformFromIteration=first))
            pz = GaussianDistribution(first.distribution.mean, first.distribution.sigma + qNoise)
            second = self.unscentedTransform(pz, up, lambda pts: input.ssmModel.observationFuncImpl(pts, t))
            rNoise = input.rNoise(replace(ctx, iteration=t, firstTransformFromIteration=first, secondTransformFromIteration=second))
            py = GaussianDistribution(second.distribution.mean, second.distribution.sigma + rNoise)
            zT, yT, w = first.transformedSigmaPoints, second.transformedSigmaPoints, first.weights
            zy = np.outer(zT[0] - pz.mean, yT[0] - py.mean) * w[1]                    # :52-60
            for i in range(1, 2 * hid + 1):
                zy = zy + np.outer(zT[i] - pz.mean, yT[i] - py.mean) * w[2]
            S = py.sigma
            K = zy @ np.linalg.inv(S)                                                 # :64-65
            hiddenMeans[:, t] = pz.mean + K @ (y[:, t] - py.mean)
            hiddenCovs[t] = pz.sigma - (K @ S) @ K.T
            if ll is not None:
                ll += logGaussianDensity(y[:, t], py.mean, S)
        return FilteringOutput(hiddenMeans, hiddenCovs, ll)


class GPUnscentedKalmanFilter(UnscentedKalmanFilter):
    """GPUnscentedKalmanFilter.scala:15: UKF whose transition / observation functions and noises are GP posteriors."""

    def __init__(self, gpOptimizer, gpPredictor: GpPredictor):
        super().__init__(gpOptimizer)
        self.gpPredictor = gpPredictor
        self.kernelFunc = gpPredictor.kernelFunc
        self._models: List[FittedGp] = []

    # ---- GPUnscentedKalmanFilter.scala:105-136 ----------------------------------------------------------
    def _learnInputOutput(self, X, output, optimizeGPL: bool) -> List[FittedGp]:
        models = []
        for dim in range(output.shape[0]):
            targets = np.ascontiguousarray(output[dim, :])
            hp = self.gpPredictor.obtainOptimalHyperParams(X, None, targets, True) if optimizeGPL else self.kernelFunc.hyperParams
            models.append(self.gpPredictor.fit(X, None, targets, hp))                 # preComputeComponents, resident on the device
        self._models += models
        return models

    def learnNewSsmModelWithNoises(self, observations, trueHiddenStates, optimizeGpLearning: bool = False):
        """GPUnscentedKalmanFilter.scala:63-103 -> (SsmModel, qNoiseFunc, rNoiseFunc)."""
        hidden = np.asarray(trueHiddenStates, dtype=np.float64)
        obs = np.asarray(observations, dtype=np.float64)
        Xall = np.ascontiguousarray(hidden.T)                                         # trainingDataForPredictor
        Xprev = np.ascontiguousarray(Xall[:-1])                                       # trainingDataWithoutLastObj
        diffs = hidden[:, 1:] - hidden[:, :-1]                                        # :107-111
        sysModels = self._learnInputOutput(Xprev, diffs, optimizeGpLearning)
        obsModels = self._learnInputOutput(Xall, obs, optimizeGpLearning)

        def means(models, pts):
            pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
            return models_mean(models, pts)                                            # one device call for all dimensions

        def noise(models, point):                                                    # :138-147: diag of sigma(0,0) per dimension
            pt = np.atleast_2d(np.asarray(point, dtype=np.float64))
            return np.diag([float(m.computePosterior(pt, full_cov=False, want_v=False)[0].sigma[0]) for m in models])

        model = SsmModel(transitionFuncImpl=lambda u_t, pts, t: np.atleast_2d(pts) + means(sysModels, pts),   # :77-83
                         observationFuncImpl=lambda pts, t: means(obsModels, pts))                            # :84-90
        qNoiseFunc = lambda ctx: noise(sysModels, ctx.hiddenMeans[:, ctx.iteration - 1])                      # :95-98
        rNoiseFunc = lambda ctx: noise(obsModels, ctx.firstTransformFromIteration.distribution.mean)          # :99-102
        return model, qNoiseFunc, rNoiseFunc

    def inferHiddenStateFromSamples(self, input: UnscentedFilteringInput, hiddenSamples, params=None, computeLL: bool = True,
                                    optimizeGpLearning: bool = False) -> FilteringOutput:
        """GPUnscentedKalmanFilter.scala:26-34 with the sampled trajectory passed in (the reference draws it from
        input.ssmModel.generateSeries with a time-seeded sampler)."""
        model, q, r = self.learnNewSsmModelWithNoises(input.observations, hiddenSamples, optimizeGpLearning)
        try:
            return self.inferHiddenState(replace(input, ssmModel=model, qNoise=q, rNoise=r), params, computeLL)
        finally:
            self.close()

    def close(self):
        for m in self._models:
            m.close()
        self._models = []

"""TEST HARNESS (not product code): the reference's host-side UKF recursion, restated so that the tests can drive the
device-resident GP evaluations the way dynamicalsystems/filtering/UnscentedKalmanFilter.scala:24-118 drives
GpPredictor.computePosterior -- arbitrary Python state-space functions, d x d algebra in NumPy.  The product path is
gp_algos_b200.gp_ukf.GPUnscentedKalmanFilter (one device call per filter run, csrc/gpk_ukf.cu); this module exists to test it
against a host recursion over the same device GPs and to restate the facts of the reference's UnscentedKalmanFilterTest."""
from __future__ import annotations

import math
from dataclasses import replace
from typing import List, Optional

import numpy as np

from gp_algos_b200.gp_predictor import FittedGp, GaussianDistribution, GpPredictor, models_mean
from gp_algos_b200.gp_ukf import (FilteringOutput, SsmModel, UkfInferenceContext, UnscentedFilteringInput, UnscentedTransformOutput,
                                  UnscentedTransformParams)
from .stats_utils import logGaussianDensity, nllOfHiddenData  # noqa: F401


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



class HostRecursionGpUkf(UnscentedKalmanFilter):
    """GPUnscentedKalmanFilter.scala:63-147 with the recursion on the host and every GP evaluation on the device
    (round-1 arrangement; kept as the cross-check of the device-resident filter)."""

    def __init__(self, gpPredictor: GpPredictor):
        super().__init__(None)
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

"""dynamicalsystems/filtering/GPUnscentedKalmanFilter.scala on the device.

One GP per state / observation dimension is fitted on the sampled trajectory and stays resident in HBM
(GPUnscentedKalmanFilter.scala:123-136 -> `gpk_gp_model_fit`).  The filter run itself -- per time step two unscented transforms
(UnscentedKalmanFilter.scala:82-118), (2d+1)(d+p) GP posterior means, d+p posterior variances for the noise matrices
(:95-102,138-147), gain and update (:38-74) -- is ONE call, `gpk_gpukf_filter` (csrc/gpk_ukf.cu): nothing crosses the host
between time steps, and B independent series (different observations / initial states, same learned model) run in the same
launches.  The reference makes (2d+1)(d+p) + (d+p) `computePosterior` calls per step, each an O(n^2) scalar solve.

The reference's generic host recursion over arbitrary Scala closures (UnscentedKalmanFilter.scala:24-80) is out of scope
(SURVEY.md 2, row 8); a restatement lives in tests/host_callers/ukf_host.py as the cross-check of this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib
from .gp_predictor import FittedGp, GaussianDistribution, GpPredictor


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


class GPUnscentedKalmanFilter:
    """GPUnscentedKalmanFilter.scala:15 on the device.  `gpOptimizer` is accepted for signature parity (it only serves the
    inferWithUkfOptim* drivers, host control flow around repeated filter runs)."""

    def __init__(self, gpOptimizer, gpPredictor: GpPredictor):
        self.gpOptimizer = gpOptimizer
        self.gpPredictor = gpPredictor
        self.kernelFunc = gpPredictor.kernelFunc
        self._models: List[FittedGp] = []
        self.sysModels: List[FittedGp] = []
        self.obsModels: List[FittedGp] = []

    # ---- GPUnscentedKalmanFilter.scala:105-136: one GP per output dimension, resident on the device ----------------------------
    def _learnInputOutput(self, X, output, optimizeGPL: bool) -> List[FittedGp]:
        models = []
        for dim in range(output.shape[0]):
            targets = np.ascontiguousarray(output[dim, :])
            hp = self.gpPredictor.obtainOptimalHyperParams(X, None, targets, True) if optimizeGPL else self.kernelFunc.hyperParams
            models.append(self.gpPredictor.fit(X, None, targets, hp))
        self._models += models
        return models

    def learn(self, observations, trueHiddenStates, optimizeGpLearning: bool = False):
        """GPUnscentedKalmanFilter.scala:63-76,105-121: transition GPs on (z_t -> z_{t+1} - z_t), observation GPs on (z_t -> y_t)."""
        hidden = np.asarray(trueHiddenStates, dtype=np.float64)
        obs = np.asarray(observations, dtype=np.float64)
        Xall = np.ascontiguousarray(hidden.T)
        self.close()
        self.sysModels = self._learnInputOutput(np.ascontiguousarray(Xall[:-1]), hidden[:, 1:] - hidden[:, :-1], optimizeGpLearning)
        self.obsModels = self._learnInputOutput(Xall, obs, optimizeGpLearning)
        return self

    def filter_many(self, observations: Sequence[np.ndarray], initMeans, initCovs, params: Optional[UnscentedTransformParams] = None,
                    computeLL: bool = True) -> List[FilteringOutput]:
        """B filter runs of UnscentedKalmanFilter.scala:24-80 over the learned GP model in ONE device call: observations[b] is
        d_obs x tMax, initMeans[b] / initCovs[b] the prior of series b."""
        if not self.sysModels or not self.obsModels:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: learn() the GP state-space model first")
        up = params or UnscentedTransformParams()
        h = self.sysModels[0].handle
        ys = [np.asarray(o, dtype=np.float64) for o in observations]
        B = len(ys)
        p, T = ys[0].shape
        d = len(self.sysModels)
        if len(self.obsModels) != p or any(o.shape != (p, T) for o in ys):
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: observation series disagree with the learned model")
        y = np.ascontiguousarray(np.stack([o.T for o in ys]))                        # [b][t][j]: Breeze p x T column-major per series
        m0 = np.ascontiguousarray(np.stack([np.asarray(m, dtype=np.float64).reshape(d) for m in initMeans]))
        c0 = np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.float64).T for c in initCovs]))   # column-major d x d
        means = np.empty((B, T, d)); covs = np.empty((B, T, d, d)); ll = np.zeros(B)
        import ctypes as C
        arr = lambda ms: (C.c_void_p * len(ms))(*[m._m for m in ms])
        h.check(h.lib.gpk_gpukf_filter(h.h, arr(self.sysModels), d, arr(self.obsModels), p, B, T, _lib.ptr(y), _lib.ptr(m0), _lib.ptr(c0),
                                       float(up.alpha), float(up.beta), float(up.kappa), int(computeLL), _lib.ptr(means), _lib.ptr(covs),
                                       _lib.ptr(ll)))
        return [FilteringOutput(means[b].T.copy(), [covs[b, t].T.copy() for t in range(T)], float(ll[b]) if computeLL else None)
                for b in range(B)]

    def inferHiddenStateFromSamples(self, input: UnscentedFilteringInput, hiddenSamples, params=None, computeLL: bool = True,
                                    optimizeGpLearning: bool = False) -> FilteringOutput:
        """GPUnscentedKalmanFilter.scala:26-34 with the sampled trajectory passed in (the reference draws it from
        input.ssmModel.generateSeries with a time-seeded sampler)."""
        self.learn(input.observations, hiddenSamples, optimizeGpLearning)
        try:
            return self.filter_many([input.observations], [input.initMean], [input.initCov], params, computeLL)[0]
        finally:
            self.close()

    def close(self):
        for m in self._models:
            m.close()
        self._models, self.sysModels, self.obsModels = [], [], []

"""Host mirror of gp/regression/GpPredictor.scala over libgpk's fused C-ABI entry points."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .co2_prediction import Co2HyperParams, Co2Kernel
from .kernel_requisites import GaussianRbfKernel, GaussianRbfParams


@dataclass
class GaussianDistribution:
    """utils/StatsUtils.scala:19-21."""
    mean: np.ndarray
    sigma: np.ndarray

    @property
    def dim(self) -> int:
        return len(self.mean)


@dataclass
class PredictionTrainingInput:
    """GpPredictor.scala:171-172."""
    trainingData: np.ndarray
    sigmaNoise: Optional[float]
    targets: np.ndarray


@dataclass
class PredictionInput:
    """GpPredictor.scala:162-169."""
    trainingData: np.ndarray
    testData: np.ndarray
    sigmaNoise: Optional[float]
    targets: np.ndarray

    @property
    def toPredictionTrainingInput(self) -> PredictionTrainingInput:
        return PredictionTrainingInput(self.trainingData, self.sigmaNoise, self.targets)


def _theta_of(hyperParams) -> np.ndarray:
    if isinstance(hyperParams, (GaussianRbfParams, Co2HyperParams)):
        return np.ascontiguousarray(hyperParams.toDenseVector)
    return np.ascontiguousarray(np.asarray(hyperParams, dtype=np.float64))


class GpPredictor:
    """GpPredictor.scala:15: constructed from a kernel function (same one-argument constructor the Spring
    beans use, spring-context.xml:33-51)."""

    def __init__(self, kernelFunc, handle: _lib.Handle | None = None):
        if not isinstance(kernelFunc, (GaussianRbfKernel, Co2Kernel)):
            raise TypeError("only GaussianRbfKernel and Co2Kernel are lowered to the GPU path")
        self.kernelFunc = kernelFunc
        self._handle = handle

    def _theta(self, hyperParams, D: int) -> np.ndarray:
        """theta as libgpk reads it for this predictor's kernel family, with the reference's own failure modes."""
        theta = _theta_of(hyperParams if hyperParams is not None else self.kernelFunc.hyperParams)
        if isinstance(self.kernelFunc, Co2Kernel):
            if D != 1:       # require(...) Co2Prediction.scala:39
                raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: This kernel is applicable only for 1D objects")
            if len(theta) < 11:   # getHyperParams reads positions 1..11 (Co2Prediction.scala:87-92)
                raise IndexError(f"java.lang.IndexOutOfBoundsException: {len(theta)} not in [0,{len(theta)})")
            return np.ascontiguousarray(theta[:11])
        if len(theta) != D + 2:   # require(...) KernelRequisites.scala:55
            raise ValueError(f"requirement failed: {len(theta)} does not equal to {D + 2}")
        return theta

    def _family(self):
        return self.handle.kernel_family(self.kernelFunc.family)

    @property
    def handle(self) -> _lib.Handle:
        if self._handle is None:
            self._handle = _lib.default_handle()
        return self._handle

    @staticmethod
    def _xy(trainingData, targets):
        X = _lib.fmat(trainingData)
        y = _lib.fmat(targets)
        if X.shape[0] != len(y):  # require(...) GpPredictor.scala:108
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: Number of objects in training data "
                                            "matrix should be equal to targets vector length")
        return X, y

    # ---- GpPredictor.scala:104-124 (+ overload :89-94) ------------------------------------------------
    def preComputeComponents(self, trainingData, sigmaNoise, targets, hyperParams=None):
        """-> (L, alphaVec, Option[noise * I]).  Like the reference, the third element is a dense n x n matrix
        when sigmaNoise is defined (GpPredictor.scala:116-117); it is built lazily on the host."""
        h = self.handle
        X, y = self._xy(trainingData, targets)
        n, D = X.shape
        theta = self._theta(hyperParams, D)
        L = np.empty((n, n), order="F")
        alpha = np.empty(n)
        ll = C.c_double()
        with self._family():
            h.check(h.lib.gpk_gp_fit(h.h, _lib.ptr(X), n, D, n, _lib.ptr(y), _lib.ptr(theta), int(sigmaNoise is not None),
                                     float(sigmaNoise or 0.0), _lib.ptr(L), n, _lib.ptr(alpha), C.addressof(ll)))
        noise = None if sigmaNoise is None else np.eye(n) * sigmaNoise
        return L, alpha, noise

    # ---- GpPredictor.scala:60-80 ----------------------------------------------------------------------
    def logLikelihoodWithDerivatives(self, input: PredictionTrainingInput, hyperParams, optimizedParamsNum: int):
        h = self.handle
        X, y = self._xy(input.trainingData, input.targets)
        n, D = X.shape
        theta = self._theta(hyperParams, D)
        ll = C.c_double()
        g = np.zeros(max(optimizedParamsNum, 1))
        s = input.sigmaNoise
        with self._family():
            h.check(h.lib.gpk_gp_nll_grad(h.h, _lib.ptr(X), n, D, n, _lib.ptr(y), _lib.ptr(theta), int(s is not None),
                                          float(s or 0.0), int(optimizedParamsNum), C.addressof(ll), _lib.ptr(g)))
        return ll.value, g[:optimizedParamsNum]

    # ---- GpPredictor.scala:24-43 ----------------------------------------------------------------------
    def predict(self, input: PredictionInput, hyperParams=None):
        h = self.handle
        X, y = self._xy(input.trainingData, input.targets)
        Xs = _lib.fmat(input.testData)
        n, D = X.shape
        theta = self._theta(hyperParams, D)
        m = Xs.shape[0]
        mean = np.empty(m)
        sigma = np.empty((m, m), order="F")
        ll = C.c_double()
        s = input.sigmaNoise
        with self._family():
            h.check(h.lib.gpk_gp_predict(h.h, _lib.ptr(X), n, D, n, _lib.ptr(y), _lib.ptr(Xs), m, m, _lib.ptr(theta),
                                         int(s is not None), float(s or 0.0), _lib.ptr(mean), _lib.ptr(sigma), m, C.addressof(ll)))
        return GaussianDistribution(mean, sigma), ll.value

    # ---- GpPredictor.scala:126-142 ---------------------------------------------------------------------
    def obtainOptimalHyperParams(self, trainingData, sigmaNoise, targets, optimizeNoise: bool, optimizer=None):
        """L-BFGS(m = 4, maxIter = 20) maximisation of logLikelihoodWithDerivatives in the natural parameters; returns the
        best point seen (optimization/Optimization.scala:37-61).  The optimiser is host control flow (Breeze's in the
        reference, `BreezeLbfgsOptimizer` here); every objective/gradient evaluation is one gpk_gp_nll_grad call.
        optimizeNoise = False reproduces the reference's defect: the start point drops the noise entry
        (`toDenseVector(0 to -2)`, :130-132) and `fromDenseVector` then fails its length `require`
        (KernelRequisites.scala:55) on the first evaluation."""
        from .gp_optimizer import BreezeLbfgsOptimizer
        opt = optimizer or BreezeLbfgsOptimizer(maxIter=20)
        hp = self.kernelFunc.hyperParams
        init = hp.toDenseVector if optimizeNoise else hp.toDenseVector[:-1]
        ptInput = PredictionTrainingInput(trainingData, sigmaNoise, targets)

        def llObjFunction(currentParams):
            hyperParams = hp.fromDenseVector(np.asarray(currentParams, dtype=np.float64))
            return self.logLikelihoodWithDerivatives(ptInput, hyperParams, len(currentParams))

        return hp.fromDenseVector(opt.maximize(llObjFunction, np.asarray(init, dtype=np.float64)))

    # ---- GpPredictor.scala:82-87 -----------------------------------------------------------------------
    def predictWithParamsOptimization(self, input: PredictionInput, optimizeNoise: bool, optimizer=None):
        optimal = self.obtainOptimalHyperParams(input.trainingData, input.sigmaNoise, input.targets, optimizeNoise, optimizer)
        dist, ll = self.predict(input, optimal)
        return dist, ll, optimal

    # ---- GpPredictor.scala:96-101 ----------------------------------------------------------------------
    def preComputeComponentsWithHpOptimization(self, trainingData, sigmaNoise, targets, optimizer=None):
        optimal = self.obtainOptimalHyperParams(trainingData, sigmaNoise, targets, True, optimizer)
        return self.preComputeComponents(trainingData, sigmaNoise, targets, optimal), optimal

    # ---- GpPredictor.scala:45-58 ----------------------------------------------------------------------
    def computePosterior(self, trainingData, testData, l, alphaVec, kernelFunc=None):
        """-> (GaussianDistribution(mean, sigma), vMatrix).  The factor `l` is adopted on the device
        (its inverse is formed once) -- use FittedGp for the fit-once / predict-many pattern."""
        kf = kernelFunc or self.kernelFunc
        with self.handle.kernel_family(kf.family):
            model = FittedGp.from_factor(self.handle, trainingData, l, alphaVec, kf)
        try:
            return model.computePosterior(testData)
        finally:
            model.close()

    def fit(self, trainingData, sigmaNoise, targets, hyperParams=None) -> "FittedGp":
        """Device-resident preComputeComponents: the GP-UKF / GP-UCB call pattern
        (GPUnscentedKalmanFilter.scala:77-88,123-147) without moving L across PCIe."""
        theta = self._theta(hyperParams, _lib.fmat(trainingData).shape[1])
        with self._family():
            return FittedGp.fit(self.handle, trainingData, targets, theta, sigmaNoise)


def models_mean(models, testData) -> np.ndarray:
    """Posterior means of several resident models at the same test rows in ONE device call -> (len(testData), len(models)).
    The GP-UKF pattern: one GP per state / observation dimension, all sigma points at once
    (GPUnscentedKalmanFilter.scala:77-90)."""
    Xs = _lib.fmat(np.atleast_2d(testData))
    m, k = Xs.shape[0], len(models)
    h = models[0].handle
    arr = (C.c_void_p * k)(*[mdl._m for mdl in models])
    out = np.empty((k, m))
    h.check(h.lib.gpk_gp_models_mean(h.h, arr, k, _lib.ptr(Xs), m, m, _lib.ptr(out)))
    return out.T.copy()


def models_mean_var(models, testData):
    """Posterior means AND variances (diagonal of computePosterior's sigma, including noiseVar^2) of several resident models at
    the same test rows in ONE device call -> two (len(testData), len(models)) arrays.  GPUnscentedKalmanFilter.scala:77-90
    (means at the sigma points) and :138-147 (variances for the Q / R noise matrices)."""
    Xs = _lib.fmat(np.atleast_2d(testData))
    m, k = Xs.shape[0], len(models)
    h = models[0].handle
    arr = (C.c_void_p * k)(*[mdl._m for mdl in models])
    mean = np.empty((k, m)); var = np.empty((k, m))
    h.check(h.lib.gpk_gp_models_mean_var(h.h, arr, k, _lib.ptr(Xs), m, m, _lib.ptr(mean), _lib.ptr(var)))
    return mean.T.copy(), var.T.copy()


class FittedGp:
    """Opaque device token for (X, L^-1, alpha, theta) -- SURVEY.md 8(b) 'ownership'."""

    def __init__(self, handle, token, n, D, ll=None, sigmaNoise=None):
        self.handle, self._m, self.n, self.D, self.logLikelihood, self.sigmaNoise = handle, token, n, D, ll, sigmaNoise

    @classmethod
    def fit(cls, handle, trainingData, targets, theta, sigmaNoise=None):
        X, y = GpPredictor._xy(trainingData, targets)
        n, D = X.shape
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        tok = C.c_void_p()
        ll = C.c_double()
        handle.check(handle.lib.gpk_gp_model_fit(handle.h, _lib.ptr(X), n, D, n, _lib.ptr(y), _lib.ptr(theta),
                                                 int(sigmaNoise is not None), float(sigmaNoise or 0.0), C.byref(tok),
                                                 C.addressof(ll)))
        return cls(handle, tok, n, D, ll.value, sigmaNoise)

    @classmethod
    def from_factor(cls, handle, trainingData, l, alphaVec, kernelFunc):
        X = _lib.fmat(trainingData)
        L = _lib.fmat(l)
        a = _lib.fmat(alphaVec)
        n, D = X.shape
        theta = np.ascontiguousarray(kernelFunc.theta)
        tok = C.c_void_p()
        handle.check(handle.lib.gpk_gp_model_from_factor(handle.h, _lib.ptr(X), n, D, n, _lib.ptr(L), n, _lib.ptr(a),
                                                         _lib.ptr(theta), C.byref(tok)))
        return cls(handle, tok, n, D)

    def computePosterior(self, testData, full_cov: bool = True, want_v: bool = True):
        Xs = _lib.fmat(testData)
        m = Xs.shape[0]
        mean = np.empty(m)
        sigma = np.empty((m, m), order="F") if full_cov else np.empty(m)
        V = np.empty((self.n, m), order="F") if want_v else None
        self.handle.check(self.handle.lib.gpk_gp_model_predict(
            self.handle.h, self._m, _lib.ptr(Xs), m, m, int(full_cov), _lib.ptr(mean), _lib.ptr(sigma), m,
            _lib.ptr(V) if want_v else None, self.n))
        return GaussianDistribution(mean, sigma), V

    def mean(self, testData) -> np.ndarray:
        """Posterior mean only: K* alpha (GpPredictor.scala:53-54), no triangular work -- what the GP-UKF transition and
        observation functions read from computePosterior (GPUnscentedKalmanFilter.scala:77-90)."""
        Xs = _lib.fmat(np.atleast_2d(testData))
        m = Xs.shape[0]
        mean = np.empty(m)
        self.handle.check(self.handle.lib.gpk_gp_model_predict(self.handle.h, self._m, _lib.ptr(Xs), m, m, 0, _lib.ptr(mean), None, m,
                                                               None, self.n))
        return mean

    def append(self, point, target: float) -> None:
        """One more training point with unchanged hyper-parameters: the GP-UCB outer loop (GPOptimizer.scala:48-71) refits with
        preComputeComponents every iteration; here the resident L^-1 and alpha get the bordered row in O(n^2).  Raises
        NotPositiveDefiniteError (minor = n+1, model unchanged) where the refit's `cholesky` would throw."""
        x = np.ascontiguousarray(point, dtype=np.float64).ravel()
        if x.size != self.D:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, f"requirement failed: point has {x.size} coordinates, model has {self.D}")
        d = C.c_double()
        s = self.sigmaNoise
        self.handle.check(self.handle.lib.gpk_gp_model_append(self.handle.h, self._m, _lib.ptr(x), float(target), int(s is not None),
                                                              float(s or 0.0), C.addressof(d)))
        self.n += 1
        if self.logLikelihood is not None:
            self.logLikelihood += d.value

    @property
    def alphaVec(self) -> np.ndarray:
        a = np.empty(self.n)
        self.handle.check(self.handle.lib.gpk_gp_model_get_alpha(self.handle.h, self._m, _lib.ptr(a)))
        return a

    def close(self):
        if self._m:
            self.handle.lib.gpk_gp_model_destroy(self.handle.h, self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

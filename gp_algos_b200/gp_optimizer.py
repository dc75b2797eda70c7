"""Host mirror of gp/optimization/GPOptimizer.scala (GP-UCB Bayesian optimisation) and optimization/Optimization.scala.

The O(n^2)/O(n^3) work of the inner loop -- fitting the surrogate (GPOptimizer.scala:51), and per candidate point the
posterior, the two n x D kernel-derivative matrices and `inversedL * trainTestDerMtx` (GPOptimizer.scala:85-103) -- runs in
libgpk against a device-resident model (`gpk_gp_model_fit` once, `gpk_gp_model_append` for the point every outer iteration
adds -- the hyper-parameters do not change inside the loop, so the refit of GPOptimizer.scala:51 is a bordered update --
and `gpk_gp_model_ucb` per evaluation, any number of candidate points per call).  What stays on the host is control flow only: the random grid, the restarts'
start points and the gradient optimiser's line search, exactly the parts the reference delegates to Breeze / scala.util.Random.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from .gp_predictor import FittedGp, GpPredictor


class BreezeLbfgsOptimizer:
    """optimization/Optimization.scala:30-61: L-BFGS(m = 4, maxIter) with the best-seen bookkeeping of :37-55.
    Breeze's LBFGS is un-vendored third-party code (breeze 0.8.1) whose trajectory no reference test pins; SciPy's
    L-BFGS-B with the same memory and iteration cap stands in, and the reference's own wrapper logic is reproduced."""

    def __init__(self, maxIter: int = 10):
        self.maxIter = maxIter

    def minimize(self, func: Callable, initPoint: Sequence[float]) -> np.ndarray:
        from scipy.optimize import minimize as sp_minimize
        best = {"x": np.array(initPoint, dtype=np.float64), "v": float("inf")}       # Double.MaxValue

        def calc(p):
            value, grad = func(np.array(p, dtype=np.float64))
            if value < best["v"]:                                                    # Optimization.scala:44-46
                best["x"], best["v"] = np.array(p, dtype=np.float64), float(value)
            return float(value), np.asarray(grad, dtype=np.float64)

        res = sp_minimize(calc, np.array(initPoint, dtype=np.float64), jac=True, method="L-BFGS-B",
                          options={"maxiter": self.maxIter, "maxcor": 4})
        optimalVal = func(res.x)[0]                                                  # Optimization.scala:52
        return res.x if optimalVal < best["v"] else best["x"]                        # :53-55

    def maximize(self, func: Callable, initPoint: Sequence[float]) -> np.ndarray:
        def minus(p):                                                                # Optimization.scala:58-60
            v, g = func(p)
            return -v, -np.asarray(g, dtype=np.float64)
        return self.minimize(minus, initPoint)


@dataclass
class GPOInput:
    """GPOptimizer.scala:158-159; ranges: sequence of (start, end)."""
    ranges: Sequence
    mParam: int
    cParam: int
    kParam: float
    optimizeHpOnInitGrid: bool = False


def ucb_with_gradient(model: FittedGp, points, kParam: float):
    """maximizeUCB's objective (GPOptimizer.scala:87-104) at the rows of `points`: (ucb[m], grad[m, D], mean[m], var[m])."""
    Xs = _lib.fmat(np.atleast_2d(points))
    m, D = Xs.shape
    if D != model.D:
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: test point dimension")
    ucb = np.empty(m); grad = np.empty((m, D), order="F"); mean = np.empty(m); var = np.empty(m)
    h = model.handle
    h.check(h.lib.gpk_gp_model_ucb(h.h, model._m, _lib.ptr(Xs), m, m, float(kParam), _lib.ptr(ucb), _lib.ptr(grad), m,
                                   _lib.ptr(mean), _lib.ptr(var)))
    return ucb, grad, mean, var


class GPOptimizer:
    """GPOptimizer.scala:19: (gpPredictor, noise: Option[Double], gradientOptimizer)."""

    def __init__(self, gpPredictor: GpPredictor, noise: Optional[float], gradientOptimizer, seed: Optional[int] = None):
        self.gpPredictor, self.noise, self.gradientOptimizer = gpPredictor, noise, gradientOptimizer
        self.hyperParams = gpPredictor.kernelFunc.hyperParams
        self.rng = np.random.default_rng(seed)     # the reference seeds scala.util.Random with System.nanoTime (:136)

    def minimize(self, objFunc, params: GPOInput):
        opt, val = self.maximize(lambda p: -objFunc(p), params)                       # GPOptimizer.scala:30-33
        return opt, -val

    def maximize(self, func, params: GPOInput):
        ranges, m, c, k = params.ranges, params.mParam, params.cParam, params.kParam
        if not (c >= 1 and m >= 1):
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: Params m and c needs to be greater or equal 1")
        pointSet = self.prepareGrid(ranges)
        evaluated = self.evaluateGridPoints(pointSet, func)
        hp = (self.gpPredictor.obtainOptimalHyperParams(pointSet, self.noise, evaluated, True)     # GPOptimizer.scala:42-46
              if params.optimizeHpOnInitGrid else self.hyperParams)
        # preComputeComponents (GPOptimizer.scala:51) once; every later iteration's refit differs from this model by the one
        # appended point (same hyper-parameters, same noise), which FittedGp.append applies in place
        model = self.gpPredictor.fit(pointSet, self.noise, evaluated, hp)
        try:
            for it in range(m):                                                       # GPOptimizer.scala:48-77
                mean = pointSet.mean(axis=0)                                          # StatsUtils.scala:61-72
                diff = pointSet - mean
                cov = diff.T @ diff / pointSet.shape[0]
                best_pt, best_ucb = None, -np.finfo(float).max
                for _r in range(c):                                                   # :54-61
                    init = self.rng.multivariate_normal(mean, cov, method="svd")
                    val, pt = self.maximizeUCB(model, init, k)
                    if val > best_ucb:
                        best_pt, best_ucb = pt, val
                if best_pt is None:
                    best_pt = self.rng.multivariate_normal(mean, cov, method="svd")
                try:                                                                  # :64-71
                    v = func(np.array(best_pt))
                except Exception:
                    continue                                                          # point set unchanged (:70)
                if it + 1 < m:                    # the reference refits (and may throw from `cholesky`, :51) only if another iteration follows
                    model.append(best_pt, v)
                pointSet = np.vstack([pointSet, best_pt[None, :]])
                evaluated = np.concatenate([evaluated, [v]])
        finally:
            model.close()
        i = int(np.argmax(evaluated))                                                 # :73-79 (first maximum)
        return pointSet[i].copy(), float(evaluated[i])

    def maximizeUCB(self, model: FittedGp, initPoint, kParam: float):
        """GPOptimizer.scala:82-109 against the resident model -> (ucb value, point)."""
        def f(p):
            ucb, grad, _, _ = ucb_with_gradient(model, p, kParam)
            return float(ucb[0]), grad[0]
        opt = self.gradientOptimizer.maximize(f, np.asarray(initPoint, dtype=np.float64))
        return f(opt)[0], np.asarray(opt, dtype=np.float64)

    def evaluateGridPoints(self, grid, func) -> np.ndarray:                          # GPOptimizer.scala:128-132
        return np.array([func(np.array(grid[i])) for i in range(grid.shape[0])], dtype=np.float64)

    def prepareGrid(self, ranges) -> np.ndarray:                                     # GPOptimizer.scala:134-147: 3*dim random points
        dim = len(ranges)
        grid = np.zeros((3 * dim, dim))
        for i in range(grid.shape[0]):
            for j, (lo, hi) in enumerate(ranges):
                if not lo < hi:
                    raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed")
                grid[i, j] = lo + (hi - lo) * self.rng.random()
        return grid

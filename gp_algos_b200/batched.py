"""Batched independent GPs (BASELINE.json config 4) and their partition over the GPUs of one node.

The reference has no batching: `GPUnscentedKalmanFilter.learnInputOutput` (GPUnscentedKalmanFilter.scala:123-136) fits one
GP per output dimension in a Scala loop, `GPOptimizer` (GPOptimizer.scala:54-61) restarts L-BFGS `c` times sequentially, and
every sigma point costs one `computePosterior` call (GPUnscentedKalmanFilter.scala:77-88).  Those are independent problems of
one shape, so libgpk evaluates B of them per launch sequence (`gpk_gp_nll_grad_batched`, `gpk_gp_predict_batched`) and this
module splits a batch over ranks: problems b with shard_bounds(B, rank, world) go to GPU `rank`; NO data-path collective is
needed -- only the (tiny) results are gathered on the host."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib


def shard_bounds(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of problems 0..B-1: the first B % world ranks get one extra problem."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _stack_problems(X, B):
    """Accepts one shared (n, D) matrix or a (B, n, D) stack; returns (column-major buffer, strideX, n, D)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 2:
        Xf = np.asfortranarray(X)
        return Xf, 0, X.shape[0], X.shape[1]
    if X.ndim != 3 or X.shape[0] != B:
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "X must be (n, D) or (B, n, D)")
    n, D = X.shape[1], X.shape[2]
    # (B, D, n) C-order == B column-major n x D blocks (Breeze DenseMatrix layout); no copy when the caller's stack already
    # is one, e.g. np.empty((B, D, n)).transpose(0, 2, 1) over pinned memory
    buf = np.ascontiguousarray(np.transpose(X, (0, 2, 1)))
    return buf, n * D, n, D


def log_likelihood_with_derivatives_batched(X, ys, thetas, sigmaNoise=None, nparams: Optional[int] = None, handle=None):
    """B evaluations of GpPredictor.logLikelihoodWithDerivatives (GpPredictor.scala:60-80) in one call.
    X: (n, D) shared or (B, n, D); ys: (B, n); thetas: (B, D+2).  Returns (ll[B], grad[B, nparams], info[B])."""
    h = handle or _lib.default_handle()
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    thetas = np.ascontiguousarray(thetas, dtype=np.float64)
    B = ys.shape[0]
    Xb, strideX, n, D = _stack_problems(X, B)
    if ys.shape[1] != n or thetas.shape != (B, D + 2):
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: shapes of X, targets and hyper-parameters disagree")
    nparams = D + 2 if nparams is None else nparams
    ll = np.empty(B); grad = np.zeros((B, max(nparams, 1))); info = np.zeros(B, dtype=np.int32)
    rc = h.lib.gpk_gp_nll_grad_batched(h.h, B, _lib.ptr(Xb), n, D, n, strideX, _lib.ptr(ys), _lib.ptr(thetas),
                                       int(sigmaNoise is not None), float(sigmaNoise or 0.0), int(nparams), _lib.ptr(ll),
                                       _lib.ptr(grad), info.ctypes.data_as(C.c_void_p))
    if rc != _lib.GPK_ENOTPD:
        h.check(rc)
    return ll, grad[:, :nparams], info


def predict_batched(X, ys, thetas, Xs, sigmaNoise=None, handle=None):
    """Fit B GPs and return the posterior mean / variance of each at its own test rows
    (GpPredictor.computePosterior semantics, GpPredictor.scala:45-58; variance includes noiseVar^2).
    X: (n, D) or (B, n, D); Xs: (m, D) shared or (B, m, D).  Returns (mean[B, m], var[B, m], ll[B], info[B])."""
    h = handle or _lib.default_handle()
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    thetas = np.ascontiguousarray(thetas, dtype=np.float64)
    B = ys.shape[0]
    Xb, strideX, n, D = _stack_problems(X, B)
    Xsb, strideXs, m, D2 = _stack_problems(Xs, B)
    if D2 != D or ys.shape[1] != n or thetas.shape != (B, D + 2):
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: shapes disagree")
    mean = np.empty((B, m)); var = np.empty((B, m)); ll = np.empty(B); info = np.zeros(B, dtype=np.int32)
    rc = h.lib.gpk_gp_predict_batched(h.h, B, _lib.ptr(Xb), n, D, n, strideX, _lib.ptr(ys), _lib.ptr(thetas), _lib.ptr(Xsb), m, m,
                                      strideXs, int(sigmaNoise is not None), float(sigmaNoise or 0.0), _lib.ptr(mean),
                                      _lib.ptr(var), _lib.ptr(ll), info.ctypes.data_as(C.c_void_p))
    if rc != _lib.GPK_ENOTPD:
        h.check(rc)
    return mean, var, ll, info


def sharded_map(B: int, local_eval: Callable[[int, int], Sequence[np.ndarray]], group=None):
    """Evaluate problems [lo, hi) of this rank with `local_eval(lo, hi)` (a tuple of arrays whose first axis is the
    problem index) and gather every rank's results on every rank, in problem order.  One process per GPU
    (torch.distributed); the only communication is this host-side gather of results -- the data path has no collective."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return tuple(np.asarray(a) for a in local_eval(0, B))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(B, rank, world)
    mine = tuple(np.asarray(a) for a in local_eval(lo, hi)) if hi > lo else None
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    parts = [g for g in gathered if g is not None]
    return tuple(np.concatenate([p[k] for p in parts], axis=0) for k in range(len(parts[0])))


# ---- MLE restarts in lockstep ----------------------------------------------------------------------------------------------
class LockstepLbfgs:
    """R independent L-BFGS(m = 4, maxIter) minimisations advanced in lockstep: in every round each live instance proposes ONE
    point and all proposals are evaluated by ONE call of `batched_func(points[R', P], which[R']) -> (values[R'], grads[R', P])`
    -- for MLE restarts that is one gpk_gp_nll_grad_batched launch sequence for all restarts instead of R sequential
    evaluations (GPOptimizer.scala:54-61 and the `obtainOptimalHyperParams` callers restart sequentially).

    What is and is not the reference's: the WRAPPER logic per instance is BreezeLbfgsOptimizer's (optimization/Optimization.scala:
    37-61: best-seen point, one extra evaluation at the end point).  The inner iteration is this module's own: two-loop L-BFGS
    with memory m, Armijo-only backtracking (sufficient decrease c1, halving; NO curvature / Wolfe condition), first step
    1/max(1, |g|), stop at |g| <= 1e-9 max(1, |f|) or maxIter accepted steps.  Breeze's LBFGS uses a strong-Wolfe line search
    and its own tolerances, so from the same start point this does NOT follow the trajectory of
    GpPredictor.obtainOptimalHyperParams (nor does the reference pin that trajectory anywhere); every returned point is an exact
    point of the same objective and never worse than its start.  A non-finite value (failed factorisation) is a rejected
    line-search step; a restart that never saw a finite value returns its start point with value +inf (-inf from maximize)."""

    def __init__(self, maxIter: int = 20, m: int = 4, c1: float = 1e-4, maxLineSearch: int = 20):
        self.maxIter, self.m, self.c1, self.maxLineSearch = maxIter, m, c1, maxLineSearch
        self.rounds = 0
        self.evaluations = 0

    def minimize(self, batched_func, initPoints):
        X0 = np.array(initPoints, dtype=np.float64, copy=True)
        R, P = X0.shape
        st = [dict(x=X0[r].copy(), f=None, g=None, hist=[], d=None, slope=None, step=None, ls=0, it=0, phase="init",
                   trial=X0[r].copy(), best_x=X0[r].copy(), best_v=np.inf) for r in range(R)]

        def direction(s):
            g, hist = s["g"], s["hist"]
            q = g.copy()
            alphas = []
            for sv, yv in reversed(hist):
                a = (sv @ q) / (sv @ yv)
                alphas.append(a)
                q -= a * yv
            if hist:
                sv, yv = hist[-1]
                q *= (sv @ yv) / (yv @ yv)
            for (sv, yv), a in zip(hist, reversed(alphas)):
                b = (yv @ q) / (sv @ yv)
                q += (a - b) * sv
            d = -q
            slope = float(g @ d)
            if not slope < 0:                                        # not a descent direction: restart from steepest descent
                s["hist"] = []
                d, slope = -g, -float(g @ g)
            s["d"], s["slope"] = d, slope
            s["step"] = 1.0 if s["hist"] else 1.0 / max(1.0, float(np.sqrt(g @ g)))
            s["ls"] = 0
            s["trial"] = s["x"] + s["step"] * d

        while True:
            live = [r for r in range(R) if st[r]["phase"] != "done"]
            if not live:
                break
            pts = np.stack([st[r]["trial"] for r in live])
            vals, grads = batched_func(pts, np.array(live))
            self.rounds += 1
            self.evaluations += len(live)
            for j, r in enumerate(live):
                s = st[r]
                v, g = float(vals[j]), np.asarray(grads[j], dtype=np.float64)
                if np.isfinite(v) and v < s["best_v"]:               # Optimization.scala:44-46
                    s["best_v"], s["best_x"] = v, s["trial"].copy()
                if s["phase"] == "final":                            # :52-55
                    s["result"] = s["trial"] if (np.isfinite(v) and v < s["best_v"]) else s["best_x"]
                    s["phase"] = "done"
                    continue
                if s["phase"] == "init":
                    if not np.isfinite(v):
                        s["result"], s["phase"] = s["x"], "done"
                        continue
                    s["f"], s["g"] = v, g
                    accepted = True
                else:
                    accepted = np.isfinite(v) and v <= s["f"] + self.c1 * s["step"] * s["slope"]
                    if accepted:
                        sv, yv = s["trial"] - s["x"], g - s["g"]
                        if sv @ yv > 1e-12 * np.sqrt((sv @ sv) * (yv @ yv)):
                            s["hist"].append((sv, yv))
                            s["hist"] = s["hist"][-self.m:]
                        s["x"], s["f"], s["g"] = s["trial"].copy(), v, g
                        s["it"] += 1
                    else:
                        s["ls"] += 1
                        s["step"] *= 0.5
                        if s["ls"] < self.maxLineSearch:
                            s["trial"] = s["x"] + s["step"] * s["d"]
                            continue
                if (not accepted or s["it"] >= self.maxIter
                        or np.sqrt(s["g"] @ s["g"]) <= 1e-9 * max(1.0, abs(s["f"]))):
                    s["phase"], s["trial"] = "final", s["x"].copy()  # the wrapper's extra evaluation at the end point
                    continue
                s["phase"] = "search"
                direction(s)
        return np.stack([s["result"] for s in st]), np.array([s["best_v"] for s in st])

    def maximize(self, batched_func, initPoints):
        def minus(p, which):
            v, g = batched_func(p, which)
            return -np.asarray(v, dtype=np.float64), -np.asarray(g, dtype=np.float64)
        pts, best = self.minimize(minus, initPoints)
        return pts, -best


def obtain_optimal_hyper_params_multistart(trainingData, targets, initThetas, sigmaNoise=None, maxIter: int = 20, handle=None):
    """GpPredictor.obtainOptimalHyperParams (GpPredictor.scala:126-142, optimizeNoise = true) from R start points at once: the
    restarts share the training set (strideX = 0) and every lockstep round is ONE batched objective+gradient call.
    Returns (thetas[R, D+2], logLikelihood[R]) -- the best-seen point and value of every restart; a restart whose kernel
    matrix was never positive definite comes back as its start point with logLikelihood = -inf."""
    X = np.asarray(trainingData, dtype=np.float64)
    y = np.ascontiguousarray(targets, dtype=np.float64)
    init = np.atleast_2d(np.asarray(initThetas, dtype=np.float64))
    if init.shape[1] != X.shape[1] + 2:
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, f"requirement failed: {init.shape[1]} does not equal to {X.shape[1] + 2}")

    def objective(points, which):
        ll, grad, info = log_likelihood_with_derivatives_batched(X, np.tile(y, (len(points), 1)), points, sigmaNoise, handle=handle)
        ll = np.where(info == 0, ll, -np.inf)                        # a restart whose K is not positive definite backs off
        return ll, grad

    return LockstepLbfgs(maxIter=maxIter).maximize(objective, init)

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
    buf = np.ascontiguousarray(np.transpose(X, (0, 2, 1)))  # (B, D, n) C-order == B column-major n x D blocks
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

"""Host mirror of utils/MatrixUtils.scala: same function names and argument meaning, NumPy arrays
(column-major copies are made as needed) instead of Breeze DenseMatrix/DenseVector, CUDA underneath.

Only the closed-form kernels `GaussianRbfKernel` and `Co2Kernel` are lowered to the GPU; any other kernel
object is rejected with TypeError -- a Scala shim would leave those on the original JVM path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .co2_prediction import Co2Kernel
from .kernel_requisites import GaussianRbfKernel


def _need_rbf(kernelFun):
    if not isinstance(kernelFun, (GaussianRbfKernel, Co2Kernel)):
        raise TypeError("only GaussianRbfKernel and Co2Kernel are lowered to the GPU path")
    return np.ascontiguousarray(kernelFun.theta)


def buildKernelMatrix(kernelFun, input1, input2=None, handle=None) -> np.ndarray:
    """MatrixUtils.scala:57-70 (one matrix: symmetric, sn^2 on i==j) / :44-55 (two: m x n, no noise)."""
    h = handle or _lib.default_handle()
    theta = _need_rbf(kernelFun)
    X1 = _lib.fmat(input1)
    n, D = X1.shape
    if input2 is None:
        K = np.empty((n, n), order="F")
        with h.kernel_family(kernelFun.family):
            h.check(h.lib.gpk_cov_se_ard(h.h, _lib.ptr(X1), n, D, n, _lib.ptr(theta), _lib.ptr(K), n))
        return K
    X2 = _lib.fmat(input2)
    m = X2.shape[0]
    if X2.shape[1] != D:
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "feature dimensions differ")
    K = np.empty((n, m), order="F")
    with h.kernel_family(kernelFun.family):
        h.check(h.lib.gpk_cov_cross_se_ard(h.h, _lib.ptr(X1), n, n, _lib.ptr(X2), m, m, D, _lib.ptr(theta), _lib.ptr(K), max(n, 1)))
    return K


def buildKernelDerMatrix(kernelFun, data, paramNum: int, handle=None) -> np.ndarray:
    """MatrixUtils.scala:72-84 buildMatrixWithFunc(data)(derAfterHyperParam(paramNum)); 1-based paramNum."""
    h = handle or _lib.default_handle()
    theta = _need_rbf(kernelFun)
    X = _lib.fmat(data)
    n, D = X.shape
    if not 1 <= paramNum <= kernelFun.hyperParametersNum:
        raise LookupError(f"scala.MatchError: {paramNum}")
    dK = np.empty((n, n), order="F")
    with h.kernel_family(kernelFun.family):
        h.check(h.lib.gpk_cov_deriv_se_ard(h.h, int(paramNum), _lib.ptr(X), n, D, n, _lib.ptr(theta), _lib.ptr(dK), n))
    return dK


def _solve(upper, T, b, transposed, handle):
    h = handle or _lib.default_handle()
    T = np.asarray(T, dtype=np.float64)
    if T.ndim != 2 or T.shape[0] != T.shape[1]:
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed")  # MatrixUtils.scala:125
    if T.flags.c_contiguous and not T.flags.f_contiguous:
        # a row-major operand IS the column-major storage of its transpose: hand the buffer over as it lies and flip the
        # view flag instead of re-striding n^2 doubles on the host (the effective operand, hence `upper`, is unchanged)
        T, transposed = T.T, not transposed
    T = _lib.fmat(T)
    n = T.shape[0]
    b = _lib.fmat(b)
    vec = b.ndim == 1
    B = b.reshape(n, 1, order="F") if vec else b
    if B.shape[0] != n:
        raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "right-hand side has the wrong number of rows")
    X = np.empty_like(B, order="F")
    h.check(h.lib.gpk_trsm(h.h, int(upper), int(transposed), _lib.ptr(T), n, n, _lib.ptr(B), B.shape[1], n, _lib.ptr(X), n))
    return X[:, 0].copy() if vec else X


def forwardSolve(L, b, transposed: bool = False, handle=None):
    """MatrixUtils.scala:17-21 / :29-31.  transposed=True passes the operand as a `.t` view of `L`."""
    return _solve(0, L, b, transposed, handle)


def backSolve(R, b, transposed: bool = False, handle=None):
    """MatrixUtils.scala:23-27 / :33-35.  backSolve(R = L.t, b) of the reference == backSolve(L, b, transposed=True)."""
    return _solve(1, R, b, transposed, handle)


def invTriangular(matrix, isUpper: bool = False, handle=None) -> np.ndarray:
    """MatrixUtils.scala:106-113."""
    h = handle or _lib.default_handle()
    T = _lib.fmat(matrix)
    n = T.shape[0]
    Ti = np.empty((n, n), order="F")
    h.check(h.lib.gpk_trtri(h.h, int(isUpper), _lib.ptr(T), n, n, _lib.ptr(Ti), n))
    return Ti


def cholesky(A, handle=None, check_symmetric: bool = True) -> np.ndarray:
    """breeze.linalg.cholesky as used at GpPredictor.scala:120 / EpParameterEstimator.scala:58."""
    h = handle or _lib.default_handle()
    A = _lib.fmat(A)
    n = A.shape[0]
    L = np.empty((n, n), order="F")
    h.check(h.lib.gpk_potrf_lower(h.h, _lib.ptr(A), n, n, _lib.ptr(L), n, int(check_symmetric)))
    return L


def cloneCols(vec, colNum: int) -> np.ndarray:
    """MatrixUtils.scala:37-42 (host helper, O(n*colNum) copy)."""
    v = np.asarray(vec, dtype=np.float64)
    return np.asfortranarray(np.repeat(v[:, None], colNum, axis=1))

"""Host mirror of utils/KernelRequisites.scala (hyper-parameter packing and the SE/ARD kernel object).

Only the *description* of the kernel lives on the host; evaluating it over matrices is done by the
CUDA kernels (matrix_utils.buildKernelMatrix, GpPredictor).  Scalar `apply` / `derAfterHyperParam` /
`gradient` are provided for API completeness on single point pairs (O(D) host arithmetic)."""
from __future__ import annotations

import math
from dataclasses import dataclass, replace

import numpy as np


@dataclass(frozen=True)
class GaussianRbfParams:
    """KernelRequisites.scala:39-60.  signalVar / noiseVar are std-dev-like (squared inside the kernel)."""
    signalVar: float
    lengthScales: np.ndarray
    noiseVar: float

    def __post_init__(self):
        object.__setattr__(self, "lengthScales", np.asarray(self.lengthScales, dtype=np.float64).copy())

    def getAtPosition(self, i: int) -> float:
        """1-based (KernelRequisites.scala:17,40-46); out of range is a scala.MatchError."""
        D = len(self.lengthScales)
        if i == 1:
            return float(self.signalVar)
        if 1 < i < D + 2:
            return float(self.lengthScales[i - 2])
        if i == D + 2:
            return float(self.noiseVar)
        raise LookupError(f"scala.MatchError: {i}")

    @property
    def toDenseVector(self) -> np.ndarray:
        D = len(self.lengthScales)
        return np.array([self.getAtPosition(k + 1) for k in range(D + 2)], dtype=np.float64)

    def fromDenseVector(self, dv) -> "GaussianRbfParams":
        dv = np.asarray(dv, dtype=np.float64)
        if len(dv) != len(self.lengthScales) + 2:  # require(...) KernelRequisites.scala:55
            raise ValueError(f"requirement failed: {len(dv)} does not equal to {len(self.lengthScales) + 2}")
        return replace(self, signalVar=float(dv[0]), lengthScales=dv[1:-1].copy(), noiseVar=float(dv[-1]))

    def __eq__(self, o):
        return (isinstance(o, GaussianRbfParams) and self.signalVar == o.signalVar and self.noiseVar == o.noiseVar
                and np.array_equal(self.lengthScales, o.lengthScales))


@dataclass(frozen=True, eq=False)
class GaussianRbfKernel:
    """KernelRequisites.scala:62-114: k(x,x') = sf^2 exp(-1/2 (x-x')^t diag(l^-2) (x-x')) + sn^2 [p==q]."""
    rbfParams: GaussianRbfParams

    family = 0   # GPK_KERNEL_SE_ARD (include/gpk.h)

    @property
    def hyperParametersNum(self) -> int:
        return len(self.rbfParams.lengthScales) + 2

    @property
    def hyperParams(self) -> GaussianRbfParams:
        return self.rbfParams

    def changeHyperParams(self, dv) -> "GaussianRbfKernel":
        return GaussianRbfKernel(self.rbfParams.fromDenseVector(dv))

    @property
    def theta(self) -> np.ndarray:
        return self.rbfParams.toDenseVector

    def _r(self, a, b):
        diff = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
        inv = 1.0 / (self.rbfParams.lengthScales * self.rbfParams.lengthScales)
        return float(np.dot(diff * inv, diff))

    def apply(self, obj1, obj2, sameIndex: bool) -> float:  # KernelRequisites.scala:66-72
        p = self.rbfParams
        v = p.signalVar * p.signalVar * math.exp(-0.5 * self._r(obj1, obj2))
        return v + p.noiseVar * p.noiseVar if sameIndex else v

    __call__ = apply

    def derAfterHyperParam(self, paramNum: int):  # KernelRequisites.scala:76-86 (1-based)
        p = self.rbfParams
        D = len(p.lengthScales)
        if not 1 <= paramNum <= D + 2:
            raise LookupError(f"scala.MatchError: {paramNum}")

        def f(v1, v2, sameIndex):
            if paramNum == 1:
                return 2 * p.signalVar * math.exp(-0.5 * self._r(v1, v2))
            if paramNum < D + 2:
                d = paramNum - 2
                diff = float(v1[d]) - float(v2[d])
                return p.signalVar ** 2 * math.exp(-0.5 * self._r(v1, v2)) * diff ** 2 * p.lengthScales[d] ** -3
            return 2 * p.noiseVar if sameIndex else 0.0
        return f

    def gradient(self, afterFirstArg: bool):  # KernelRequisites.scala:99-107
        def g(v1, v2):
            diff = np.asarray(v1, dtype=np.float64) - np.asarray(v2, dtype=np.float64)
            inv = 1.0 / (self.rbfParams.lengthScales * self.rbfParams.lengthScales)
            a1 = self.apply(v1, v2, False)
            return (diff * inv) * (-a1 if afterFirstArg else a1)
        return g

    def gradientAt(self, afterFirstArg: bool, points):
        return self.gradient(afterFirstArg)(points[0], points[1])

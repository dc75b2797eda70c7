"""Host mirror of gp/regression/Co2Prediction.scala: the reference's second closed-form KernelFunc (R&W's Mauna Loa kernel,
1-D inputs, 11 hyper-parameters) and the data reshaping of its driver.  Like GaussianRbfKernel, only the DESCRIPTION of the
kernel lives here; kernel matrices, derivatives, the factorisation and the likelihood gradient run in libgpk with the handle's
kernel family set to GPK_KERNEL_CO2 (include/gpk.h).  Scalar `apply` / `derAfterHyperParam` are O(1) host arithmetic for single
point pairs, kept for API completeness."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

GPK_KERNEL_SE_ARD, GPK_KERNEL_CO2 = 0, 1


@dataclass(frozen=True, eq=False)
class Co2HyperParams:
    """Co2Prediction.scala:18-27: a bare DenseVector; getAtPosition is 1-based (`dv(i - 1)`)."""
    dv: np.ndarray

    def __post_init__(self):
        object.__setattr__(self, "dv", np.asarray(self.dv, dtype=np.float64).copy())

    def getAtPosition(self, i: int) -> float:
        if not 1 <= i <= len(self.dv):
            raise IndexError(f"java.lang.IndexOutOfBoundsException: {i - 1} not in [0,{len(self.dv)})")
        return float(self.dv[i - 1])

    @property
    def toDenseVector(self) -> np.ndarray:
        return self.dv.copy()

    def fromDenseVector(self, dv) -> "Co2HyperParams":     # no length requirement in the reference (:20)
        return Co2HyperParams(dv)

    def __eq__(self, o):
        return isinstance(o, Co2HyperParams) and np.array_equal(self.dv, o.dv)

    def __str__(self):
        return str(self.dv)


@dataclass(frozen=True, eq=False)
class Co2Kernel:
    """Co2Prediction.scala:29-137:
       k1 = hp1^2 exp(-(x1-x2)^2/(2 hp2^2));  k2 = hp3^2 exp(-(x1-x2)^2/(2 hp4^2) - 2 sin(pi (x1-x2))^2/hp5^2)
       k3 = hp6^2 (1 + (x1-x2)^2/(2 hp8 hp7^2))^(-hp8);  k4 = hp9^2 exp(-(x1-x2)^2/(2 hp10^2)) + hp11^2 [sameIndex]."""
    co2HyperParams: Co2HyperParams

    family = GPK_KERNEL_CO2
    hyperParametersNum = 11                                    # :85

    @property
    def hyperParams(self) -> Co2HyperParams:
        return self.co2HyperParams

    def changeHyperParams(self, dv) -> "Co2Kernel":           # :60
        return Co2Kernel(Co2HyperParams(dv))

    @property
    def theta(self) -> np.ndarray:
        """The 11 values libgpk reads; a shorter vector fails where the reference's getHyperParams (:87-92) would."""
        return np.array([self.co2HyperParams.getAtPosition(i) for i in range(1, 12)], dtype=np.float64)

    def apply(self, obj1, obj2, sameIndex: bool) -> float:    # :38-56
        obj1, obj2 = np.atleast_1d(obj1), np.atleast_1d(obj2)
        if not (len(obj1) == 1 and len(obj2) == 1):
            raise ValueError("requirement failed: This kernel is applicable only for 1D objects")
        hp1, hp2, hp3, hp4, hp5, hp6, hp7, hp8, hp9, hp10, hp11 = self.theta
        xDiff = float(obj1[0]) - float(obj2[0])
        xDiffSq = xDiff * xDiff
        k1Val = hp1 * hp1 * math.exp(-xDiffSq / (2 * hp2 * hp2))
        sinVal = math.sin(math.pi * xDiff)
        k2Val = hp3 * hp3 * math.exp((-xDiffSq / (2 * hp4 * hp4)) - 2 * sinVal * sinVal / (hp5 * hp5))
        k3Pow1 = 1 + xDiffSq / (2 * hp8 * hp7 * hp7)
        k3Val = hp6 * hp6 * math.pow(k3Pow1, -hp8)
        k4Val = hp9 * hp9 * math.exp(-xDiffSq / (2 * hp10 * hp10))
        return k1Val + k2Val + k3Val + k4Val + (hp11 * hp11 if sameIndex else 0.)

    __call__ = apply

    def derAfterHyperParam(self, paramNum: int):               # :66-83 (1-based; anything else is a scala.MatchError)
        if not 1 <= paramNum <= 11:
            raise LookupError(f"scala.MatchError: {paramNum}")
        hp1, hp2, hp3, hp4, hp5, hp6, hp7, hp8, hp9, hp10, hp11 = self.theta

        def f(vec1, vec2, sameIndex):
            xDiff = float(np.atleast_1d(vec1)[0]) - float(np.atleast_1d(vec2)[0])
            sqDiff = xDiff * xDiff
            if paramNum < 3:                                   # :93-99
                e = math.exp(-sqDiff / (2 * hp2 * hp2))
                return 2 * hp1 * e if paramNum == 1 else hp1 * hp1 * e * sqDiff * math.pow(hp2, -3)
            if paramNum < 6:                                   # :101-110
                sinVal = math.sin(math.pi * xDiff)
                k2Val = hp3 * hp3 * math.exp(-sqDiff / (2 * hp4 * hp4) - 2 * sinVal * sinVal / (hp5 * hp5))
                return {3: 2 * k2Val / hp3, 4: k2Val * sqDiff * math.pow(hp4, -3),
                        5: k2Val * 4 * sinVal * sinVal * math.pow(hp5, -3)}[paramNum]
            if paramNum < 9:                                   # :112-123
                k3Pow1 = 1 + sqDiff / (2 * hp8 * hp7 * hp7)
                if paramNum == 6:
                    return 2 * hp6 * math.pow(k3Pow1, -hp8)
                if paramNum == 7:
                    return hp6 * hp6 * math.pow(k3Pow1, -hp8 - 1) * sqDiff * math.pow(hp7, -3)
                firstTerm = math.exp(-hp8 * math.log(k3Pow1))
                secondTerm = -math.log(k3Pow1) + (hp8 * sqDiff / (2 * hp7 * hp7 * hp8 * hp8 * k3Pow1))
                return hp6 * hp6 * firstTerm * secondTerm
            k4Val = hp9 * hp9 * math.exp(-sqDiff / (2 * hp10 * hp10))   # :125-135
            if paramNum == 9:
                return 2 * k4Val / hp9
            if paramNum == 10:
                return k4Val * sqDiff * math.pow(hp10, -3)
            return 2 * hp11 if sameIndex else 0.
        return f

    def gradient(self, afterFirstArg: bool):                   # `???` at :62-64
        raise NotImplementedError("scala.NotImplementedError: an implementation is missing")

    def gradientAt(self, afterFirstArg: bool, points):
        raise NotImplementedError("scala.NotImplementedError: an implementation is missing")


def co2DataToYearWithValue(matrix, trainTestRatio: float):
    """Co2Prediction.scala:159-186: rows (year, 12 monthly ppm values, annual mean) -> N x 2 (year + (month-1)/12, ppm) for
    ppm > 0, split into (train, test) by trainNum = (rows * ratio).toInt."""
    if not 0 <= trainTestRatio <= 1:
        raise ValueError("requirement failed: Division's ratio should be between 0 and 1")
    matrix = np.asarray(matrix, dtype=np.float64)
    rows = []
    for r in range(matrix.shape[0]):
        year = matrix[r, 0]
        for month in range(1, matrix.shape[1] - 1):
            if matrix[r, month] > 0:
                rows.append((year + (1 / 12.) * (month - 1), matrix[r, month]))
    whole = np.array(rows, dtype=np.float64).reshape(-1, 2)
    trainNum = int(whole.shape[0] * trainTestRatio)
    return whole[:trainNum], whole[trainNum:]

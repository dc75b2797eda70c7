"""One large GP on all GPUs of the node through the C ABI (`gpk_mg_create` / `gpk_mg_potrf_solve`, csrc/gpk_mg.cu): a thin
caller -- the schedule, the peer-to-peer panel puts and every kernel live in libgpk.so, so a JVM host binds the same two
symbols.  What it replaces: `GpPredictor.preComputeComponents` + `logLikelihood`
(gp/regression/GpPredictor.scala:104-124,144-149) at n = 65536 (BASELINE.json config 5).

The multi-PROCESS arrangement of the same factorisation (one process per GPU under torchrun, 2-D block-cyclic, NCCL panel
broadcasts) is gp_algos_b200/distributed.py."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _lib


class MultiGpuFit(NamedTuple):
    logLikelihood: float
    alphaVec: np.ndarray
    seconds: float          # device time of build + factor + solves (CUDA events)
    put_bytes: int          # bytes put into peers' panel / alpha buffers


class MultiGpuGp:
    def __init__(self, ndev: int = 0, devices: Optional[Sequence[int]] = None, nb: int = 0):
        self.lib = _lib.load()
        self._mg = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int * len(devices))(*devices)
            ndev = len(devices)
        rc = self.lib.gpk_mg_create(C.byref(self._mg), int(ndev), dev)
        if rc != _lib.GPK_OK:
            raise _lib.GpkError(rc, f"gpk_mg_create(ndev={ndev}) failed: not enough usable CUDA devices (libgpk has no CPU fallback)")
        self.ndev = int(self.lib.gpk_mg_device_count(self._mg))
        self._check(self.lib.gpk_mg_set_block(self._mg, int(nb)))

    def _check(self, rc, minor=0):
        if rc == _lib.GPK_OK:
            return
        msg = self.lib.gpk_mg_last_error(self._mg).decode()
        if rc == _lib.GPK_EINVAL:
            raise _lib.IllegalArgumentError(rc, msg)
        if rc == _lib.GPK_ENOTPD:
            raise _lib.NotPositiveDefiniteError(rc, msg, minor)
        raise _lib.GpkError(rc, msg)

    def fit(self, X, y, theta, sigmaNoise: Optional[float] = None) -> MultiGpuFit:
        Xf = _lib.fmat(X)
        yv = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        th = np.ascontiguousarray(theta, dtype=np.float64)
        n, D = Xf.shape
        if yv.shape[0] != n or th.shape[0] != D + 2:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: shapes of X, targets and hyper-parameters disagree")
        alpha = np.empty(n); ll = C.c_double(); info = C.c_int()
        rc = self.lib.gpk_mg_potrf_solve(self._mg, _lib.ptr(Xf), n, D, n, _lib.ptr(yv), _lib.ptr(th), int(sigmaNoise is not None),
                                         float(sigmaNoise or 0.0), _lib.ptr(alpha), C.addressof(ll), C.addressof(info))
        self._check(rc, info.value)
        return MultiGpuFit(ll.value, alpha, float(self.lib.gpk_mg_last_seconds(self._mg)), int(self.lib.gpk_mg_last_put_bytes(self._mg)))

    def close(self):
        if self._mg:
            self.lib.gpk_mg_destroy(self._mg)
            self._mg = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

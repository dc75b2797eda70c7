"""GPU: the C ABI with non-compact leading dimensions (Breeze views: majorStride > rows, GpPredictor.scala:122 `L.t`,
GPUnscentedKalmanFilter.scala:67 row slices).  The Python mirror always passes compact arrays, so these calls go through
ctypes directly with padded buffers and check that (i) results equal the compact call and (ii) the padding is never written."""
import ctypes as C

import numpy as np
import pytest

from gp_algos_b200 import _lib
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
SENT = -777.25


def padded(a, ld):
    """Column-major copy of `a` in a buffer with leading dimension ld > rows; padding rows hold a sentinel."""
    a = np.asarray(a, dtype=np.float64)
    buf = np.full((a.shape[1], ld), SENT)
    buf[:, :a.shape[0]] = a.T
    return buf                                   # C-order (cols, ld) == column-major (ld, cols)


def unpad(buf, rows):
    return buf[:, :rows].T.copy()


def untouched(buf, rows):
    return np.all(buf[:, rows:] == SENT)


def test_cov_chol_trsm_predict_with_padded_leading_dimensions():
    h = _lib.default_handle()
    lib = h.lib
    n, D, m = 200, 3, 7
    X, y, th = orc.make_c2(n=n, D=D, seed=3)
    Xs = X[:m] + 0.01
    ldx, ldk, ldxs = n + 5, n + 11, m + 3
    Xp, Xsp = padded(X, ldx), padded(Xs, ldxs)
    # covariance, symmetric and cross
    Kp = np.full((n, ldk), SENT)
    h.check(lib.gpk_cov_se_ard(h.h, _lib.ptr(Xp), n, D, ldx, _lib.ptr(th), _lib.ptr(Kp), ldk))
    K = unpad(Kp, n)
    assert untouched(Kp, n) and np.allclose(K, orc.fast_build_kernel_matrix(X, th), rtol=1e-13, atol=0)
    Kcp = np.full((n, m + 2), SENT)
    h.check(lib.gpk_cov_cross_se_ard(h.h, _lib.ptr(Xsp), m, ldxs, _lib.ptr(Xp), n, ldx, D, _lib.ptr(th), _lib.ptr(Kcp), m + 2))
    assert untouched(Kcp, m) and np.allclose(unpad(Kcp, m), orc.fast_build_kernel_matrix(Xs, th, X), rtol=1e-13, atol=0)
    # cholesky in / out with different strides
    Lp = np.full((n, n + 4), SENT)
    h.check(lib.gpk_potrf_lower(h.h, _lib.ptr(Kp), n, ldk, _lib.ptr(Lp), n + 4, 1))
    L = unpad(Lp, n)
    assert untouched(Lp, n) and np.allclose(L @ L.T, K, rtol=1e-12, atol=1e-13) and np.all(np.triu(L, 1) == 0)
    # forward solve with a matrix right-hand side, and the transposed-view back solve (L.t)
    Bm = np.random.default_rng(0).standard_normal((n, 4))
    Bp, Zp = padded(Bm, n + 9), np.full((4, n + 2), SENT)
    h.check(lib.gpk_trsm(h.h, 0, 0, _lib.ptr(Lp), n, n + 4, _lib.ptr(Bp), 4, n + 9, _lib.ptr(Zp), n + 2))
    Z = unpad(Zp, n)
    assert untouched(Zp, n) and np.allclose(L @ Z, Bm, rtol=1e-10, atol=1e-11)
    Wp = np.full((4, n + 6), SENT)
    h.check(lib.gpk_trsm(h.h, 1, 1, _lib.ptr(Lp), n, n + 4, _lib.ptr(Zp), 4, n + 2, _lib.ptr(Wp), n + 6))   # backSolve(L.t, Z)
    assert untouched(Wp, n) and np.allclose(L.T @ unpad(Wp, n), Z, rtol=1e-10, atol=1e-11)
    # triangular inverse
    Lip = np.full((n, n + 1), SENT)
    h.check(lib.gpk_trtri(h.h, 0, _lib.ptr(Lp), n, n + 4, _lib.ptr(Lip), n + 1))
    assert untouched(Lip, n) and np.allclose(unpad(Lip, n) @ L, np.eye(n), atol=1e-9)
    # fused entry points with a strided X
    ll, g = C.c_double(), np.zeros(D + 2)
    h.check(lib.gpk_gp_nll_grad(h.h, _lib.ptr(Xp), n, D, ldx, _lib.ptr(y), _lib.ptr(th), 0, 0.0, D + 2, C.addressof(ll), _lib.ptr(g)))
    llo, go = orc.fast_loglik_with_derivs(X, y, th)
    assert abs(ll.value - llo) <= 1e-9 * abs(llo) and np.allclose(g, go, rtol=1e-9, atol=1e-9 * np.abs(go).max())
    mean, Sp, llp = np.zeros(m), np.full((m, m + 5), SENT), C.c_double()
    h.check(lib.gpk_gp_predict(h.h, _lib.ptr(Xp), n, D, ldx, _lib.ptr(y), _lib.ptr(Xsp), m, ldxs, _lib.ptr(th), 0, 0.0, _lib.ptr(mean),
                               _lib.ptr(Sp), m + 5, C.addressof(llp)))
    mo, So, _ = orc.fast_predict(X, y, Xs, th)
    assert untouched(Sp, m) and np.allclose(mean, mo, rtol=1e-9, atol=1e-12) and np.allclose(unpad(Sp, m), So, rtol=1e-8, atol=1e-12)
    fitL, alpha, ll2 = np.full((n, n + 3), SENT), np.zeros(n), C.c_double()
    h.check(lib.gpk_gp_fit(h.h, _lib.ptr(Xp), n, D, ldx, _lib.ptr(y), _lib.ptr(th), 0, 0.0, _lib.ptr(fitL), n + 3, _lib.ptr(alpha),
                           C.addressof(ll2)))
    assert untouched(fitL, n) and np.allclose(unpad(fitL, n), L, rtol=1e-12, atol=1e-14) and abs(ll2.value - llo) <= 1e-9 * abs(llo)


def test_ep_with_padded_leading_dimensions():
    h = _lib.default_handle()
    lib = h.lib
    n, D = 150, 2
    X, t, th = orc.make_c3(n=n, D=D, seed=8)
    K = orc.fast_build_kernel_matrix(X, th)
    ldk, ldl = n + 7, n + 2
    Kp, Lp = padded(K, ldk), np.full((n, ldl), SENT)
    tau, nu, mu, ct, cn = (np.zeros(n) for _ in range(5))
    logz, sw = C.c_double(), C.c_int()
    h.check(lib.gpk_ep_fit(h.h, _lib.ptr(Kp), n, ldk, t.ctypes.data_as(C.c_void_p), 0.0, 3, 3, 1, _lib.ptr(tau), _lib.ptr(nu), _lib.ptr(mu),
                           _lib.ptr(Lp), ldl, _lib.ptr(ct), _lib.ptr(cn), C.addressof(logz), C.addressof(sw)))
    o = orc.fast_ep_estimate(K, t, fixed_sweeps=3)
    assert untouched(Lp, n) and sw.value == 3
    assert np.allclose(tau, o["tau"], rtol=1e-9) and np.allclose(nu, o["nu"], rtol=1e-9, atol=1e-12)
    assert abs(logz.value - o["logZ"]) <= 1e-9 * abs(o["logZ"]) and np.allclose(unpad(Lp, n), o["L"], rtol=1e-9, atol=1e-12)
    Ks = orc.fast_build_kernel_matrix(X[:9] + 0.05, th, X)
    Ksp = padded(Ks, 9 + 4)
    kss = np.full(9, th[0] ** 2 + th[-1] ** 2)
    prob, fm, fv = np.zeros(9), np.zeros(9), np.zeros(9)
    h.check(lib.gpk_ep_classify(h.h, _lib.ptr(Kp), n, ldk, _lib.ptr(Ksp), 9, 13, _lib.ptr(kss), _lib.ptr(tau), _lib.ptr(nu), _lib.ptr(Lp), ldl,
                                _lib.ptr(prob), _lib.ptr(fm), _lib.ptr(fv)))
    Kss = orc.fast_build_kernel_matrix(X[:9] + 0.05, th)
    po, _, _ = orc.fast_ep_classify(K, Ks, Kss, o["tau"], o["nu"], o["L"])
    assert np.allclose(prob, po, rtol=1e-9, atol=1e-12)

"""CPU oracle for the dense-GP hot path of astroHaoPeng/gp_algos -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; nothing under gp_algos_b200/ does (tests/test_no_oracle_in_product.py checks).

Two flavours of the same algorithm:

* ``literal``  -- ctypes binding of oracle/gp_oracle.c, which follows the Scala line by line
  (lower-loop-and-mirror K build, scalar row-oriented triangular solves with the reference's
  summation order, invTriangular against a dense identity, one materialised dK per parameter,
  trace of the full product, the EP site loop with full rank-1 downdates).  O(n^3 P): n <= ~600.
* ``fast``     -- the same mathematics through NumPy/SciPy (OpenBLAS dpotrf/dpotri/dtrsm, all
  host cores), fused O(n^2 P) gradient.  Used at the BASELINE.json sizes and as the timed CPU
  baseline ("port").  Cross-checked against ``literal`` and an mpmath arbiter in tests/.

PARITY STATUS: the reference cannot run here (no JVM); see the header of gp_oracle.c for what the
reference's own tests pin (checked in tests/test_oracle_pins.py) and what is "parity unpinned".

Reference citations are relative to /root/reference/src/main/scala/:
KR = utils/KernelRequisites.scala, MU = utils/MatrixUtils.scala, GPP = gp/regression/GpPredictor.scala,
EP = gp/classification/EpParameterEstimator.scala, GPC = gp/classification/GpClassifier.scala,
Co2 = gp/regression/Co2Prediction.scala (second kernel family, `with co2_kernel():`).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgporacle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/gp_oracle.c (gcc) into oracle/_build/libgporacle.so."""
    src = os.path.join(_HERE, "gp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libgporacle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def _L():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = C.CDLL(_SO)
        _lib.orc_rbf_k.restype = C.c_double
        _lib.orc_rbf_dk.restype = C.c_double
        _lib.orc_gp_loglik.restype = C.c_double
        _lib.orc_pnorm.restype = C.c_double
        _lib.orc_dnorm.restype = C.c_double
        _lib.orc_ep_avg_between.restype = C.c_double
        _lib.orc_ep_marginal_likelihood.restype = C.c_double
        _lib.orc_co2_k.restype = C.c_double
        _lib.orc_co2_k.argtypes = [C.c_double, C.c_double, _dp, C.c_int]
        _lib.orc_co2_dk.restype = C.c_double
        _lib.orc_co2_dk.argtypes = [C.c_int, C.c_double, C.c_double, _dp, C.c_int]
    return _lib


def _f(a):
    """Column-major float64 copy (Breeze DenseMatrix layout)."""
    return np.asfortranarray(np.array(a, dtype=np.float64, copy=True))


def _p(a):
    return a.ctypes.data_as(_dp)


class NotPositiveDefinite(Exception):
    """Breeze NotConvergedException from `cholesky` (dpotrf info > 0)."""

    def __init__(self, minor):
        super().__init__(f"matrix not positive definite: leading minor {minor}")
        self.minor = minor


class NotSymmetric(Exception):
    """Breeze MatrixNotSymmetricException from `cholesky`."""


def _check(info):
    if info == -2:
        raise NotSymmetric()
    if info > 0:
        raise NotPositiveDefinite(info)
    if info != 0:
        raise ValueError(f"oracle status {info}")


# --------------------------------------------------------------------------------------------
# hyper-parameter packing (KR:39-58).  theta = [signalVar, lengthScales..., noiseVar]
# --------------------------------------------------------------------------------------------
def pack_theta(signal_var, length_scales, noise_var):
    return np.concatenate([[signal_var], np.asarray(length_scales, dtype=np.float64), [noise_var]])


def get_at_position(theta, i):
    """1-based accessor (KR:40-46); i outside 1..D+2 is a scala.MatchError in the reference."""
    D = len(theta) - 2
    if i == 1:
        return theta[0]
    if 1 < i < D + 2:
        return theta[i - 1]
    if i == D + 2:
        return theta[D + 1]
    raise LookupError("scala.MatchError")


# --------------------------------------------------------------------------------------------
# Co2Kernel (gp/regression/Co2Prediction.scala:29-137): the reference's second KernelFunc.  theta = hp1..hp11, 1-D inputs.
# `with co2_kernel():` switches every lit_* routine below to it (the C builders dispatch on the family, like GpPredictor
# dispatches on its kernelFunc).
# --------------------------------------------------------------------------------------------
KERNEL_SE_ARD, KERNEL_CO2 = 0, 1
CO2_SHIPPED_HP = np.array([60., 70., 8., 50., 2., 0.34, 2.4, 0.88, 0.26, 0.2, 0.19])   # utils/TestingUtils.scala:17-20


class co2_kernel:
    def __enter__(self):
        self.prev = _L().orc_get_kernel()
        _L().orc_set_kernel(KERNEL_CO2)
        return self

    def __exit__(self, *exc):
        _L().orc_set_kernel(self.prev)
        return False


def co2_k(x1, x2, hp, same_index=False):
    """Co2Prediction.scala:38-56 apply."""
    hp = np.ascontiguousarray(hp, dtype=np.float64)
    assert hp.size == 11
    return _L().orc_co2_k(float(x1), float(x2), _p(hp), int(same_index))


def co2_dk(param_num, x1, x2, hp, same_index=False):
    """Co2Prediction.scala:66-137 derAfterHyperParam(param_num), 1-based."""
    hp = np.ascontiguousarray(hp, dtype=np.float64)
    assert hp.size == 11
    v = _L().orc_co2_dk(int(param_num), float(x1), float(x2), _p(hp), int(same_index))
    if math.isnan(v) and not 1 <= param_num <= 11:
        raise LookupError("scala.MatchError")
    return v


def co2_data_to_year_with_value(matrix, train_test_ratio):
    """Co2Prediction.scala:159-186 co2DataToYearWithValue: rows (year, 12 monthly ppm, annual mean) -> (year + (month-1)/12, ppm)
    for ppm > 0, split by ratio (trainNum = (rows * ratio).toInt)."""
    if not 0 <= train_test_ratio <= 1:
        raise ValueError("requirement failed: Division's ratio should be between 0 and 1")
    rows = []
    for r in range(matrix.shape[0]):
        year = matrix[r, 0]
        for month in range(1, matrix.shape[1] - 1):
            if matrix[r, month] > 0:
                rows.append((year + (1 / 12.) * (month - 1), matrix[r, month]))
    whole = np.array(rows, dtype=np.float64)
    train_num = int(whole.shape[0] * train_test_ratio)
    return whole[:train_num], whole[train_num:]


def make_co2_like(n=400, seed=9, years=(1958.0, 1992.0)):
    """Synthetic stand-in for the Mauna Loa series (the GPU box cannot read /root/reference): monthly-ish samples of a trend
    + annual cycle + noise on the same time axis and ppm scale, hyper-parameters = the shipped ones."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(years[0], years[1], size=n))
    y = 315.0 + 1.3 * (t - years[0]) + 0.012 * (t - years[0]) ** 2 + 3.0 * np.sin(2 * np.pi * t) + 0.2 * rng.standard_normal(n)
    return t.reshape(n, 1), y, CO2_SHIPPED_HP.copy()


# --------------------------------------------------------------------------------------------
# literal flavour (C)
# --------------------------------------------------------------------------------------------
def lit_build_kernel_matrix(X, theta, X2=None):
    """MU:57-70 (symmetric, noise on i==j) or MU:44-55 (cross, never noise)."""
    X = _f(X)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    if X2 is None:
        K = np.zeros((n, n), order="F")
        _L().orc_build_kernel_matrix(_p(X), n, D, C.c_long(n), _p(theta), _p(K), C.c_long(n))
        return K
    X2 = _f(X2)
    m = X2.shape[0]
    K = np.zeros((n, m), order="F")
    _L().orc_build_kernel_matrix_cross(_p(X), n, C.c_long(n), _p(X2), m, C.c_long(m), D, _p(theta),
                                       _p(K), C.c_long(n))
    return K


def lit_build_der_matrix(param_num, X, theta):
    """MU:72-84 with f = derAfterHyperParam(param_num) (KR:76-86); param_num is 1-based."""
    X = _f(X)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    dK = np.zeros((n, n), order="F")
    _L().orc_build_der_matrix(param_num, _p(X), n, D, C.c_long(n), _p(theta), _p(dK), C.c_long(n))
    return dK


def lit_forward_solve(L, b, trans=False):
    """MU:17-21 / MU:29-31."""
    L = _f(L)
    b = _f(b)
    n = L.shape[0]
    x = np.zeros_like(b, order="F")
    if b.ndim == 1:
        _L().orc_forward_solve_vec(_p(L), n, C.c_long(n), int(trans), _p(b), _p(x))
    else:
        _L().orc_forward_solve_mat(_p(L), n, C.c_long(n), int(trans), _p(b), b.shape[1], C.c_long(n),
                                   _p(x), C.c_long(n))
    return x


def lit_back_solve(R, b, trans=False):
    """MU:23-27 / MU:33-35.  trans=True: R is given as its transpose (the `L.t` view, GPP:122)."""
    R = _f(R)
    b = _f(b)
    n = R.shape[0]
    x = np.zeros_like(b, order="F")
    if b.ndim == 1:
        _L().orc_back_solve_vec(_p(R), n, C.c_long(n), int(trans), _p(b), _p(x))
    else:
        _L().orc_back_solve_mat(_p(R), n, C.c_long(n), int(trans), _p(b), b.shape[1], C.c_long(n),
                                _p(x), C.c_long(n))
    return x


def lit_inv_triangular(A, is_upper=False):
    """MU:106-113."""
    A = _f(A)
    n = A.shape[0]
    Ai = np.zeros((n, n), order="F")
    _L().orc_inv_triangular(_p(A), n, C.c_long(n), int(is_upper), _p(Ai), C.c_long(n))
    return Ai


def lit_cholesky(A):
    """Breeze `cholesky` (GPP:120, EP:58) -> dpotf2('L') recurrence."""
    A = _f(A)
    n = A.shape[0]
    Lo = np.zeros((n, n), order="F")
    _check(_L().orc_cholesky_lower(_p(A), n, C.c_long(n), _p(Lo), C.c_long(n)))
    return Lo


def lit_precompute(X, y, theta, sigma_noise=None):
    """GPP:104-124 -> (L, alpha)."""
    X = _f(X)
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    Lo = np.zeros((n, n), order="F")
    alpha = np.zeros(n)
    _check(_L().orc_gp_precompute(_p(X), n, D, C.c_long(n), _p(y), _p(theta),
                                  int(sigma_noise is not None), C.c_double(sigma_noise or 0.0),
                                  _p(Lo), _p(alpha)))
    return Lo, alpha


def lit_loglik(alpha, L, y):
    """GPP:144-149."""
    L = _f(L)
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = L.shape[0]
    return _L().orc_gp_loglik(_p(alpha), _p(L), n, C.c_long(n), _p(y))


def lit_loglik_with_derivs(X, y, theta, sigma_noise=None, nparams=None):
    """GPP:60-80 -> (ll, grad[nparams])."""
    X = _f(X)
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    nparams = len(theta) if nparams is None else nparams       # D + 2 (GaussianRbfKernel) or 11 (Co2Kernel)
    ll = C.c_double()
    g = np.zeros(nparams)
    _check(_L().orc_gp_loglik_with_derivs(_p(X), n, D, C.c_long(n), _p(y), _p(theta),
                                          int(sigma_noise is not None), C.c_double(sigma_noise or 0.0),
                                          nparams, C.byref(ll), _p(g)))
    return ll.value, g


def lit_compute_posterior(X, Xs, L, alpha, theta):
    """GPP:45-58 -> (mean[m], sigma[m,m], V[n,m])."""
    X = _f(X)
    Xs = _f(Xs)
    L = _f(L)
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    m = Xs.shape[0]
    mean = np.zeros(m)
    sigma = np.zeros((m, m), order="F")
    V = np.zeros((n, m), order="F")
    _L().orc_gp_compute_posterior(_p(X), n, D, C.c_long(n), _p(Xs), m, C.c_long(m), _p(L), C.c_long(n),
                                  _p(alpha), _p(theta), _p(mean), _p(sigma), _p(V))
    return mean, sigma, V


def lit_predict(X, y, Xs, theta, sigma_noise=None):
    """GPP:24-43 -> (mean[m], sigma[m,m], ll)."""
    X = _f(X)
    Xs = _f(Xs)
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n, D = X.shape
    m = Xs.shape[0]
    mean = np.zeros(m)
    sigma = np.zeros((m, m), order="F")
    ll = C.c_double()
    _check(_L().orc_gp_predict(_p(X), n, D, C.c_long(n), _p(y), _p(Xs), m, C.c_long(m), _p(theta),
                               int(sigma_noise is not None), C.c_double(sigma_noise or 0.0),
                               _p(mean), _p(sigma), C.byref(ll)))
    return mean, sigma, ll.value


def lit_kernel_gradient(theta, v1, v2, after_first_arg):
    """KR:99-107 GaussianRbfKernel.gradient(afterFirstArg)(vec1, vec2), same association."""
    theta = np.asarray(theta, dtype=np.float64)
    D = len(theta) - 2
    v1 = np.asarray(v1, dtype=np.float64); v2 = np.asarray(v2, dtype=np.float64)
    diff = v1 - v2
    inv = np.array([1.0 / (ls * ls) for ls in theta[1:D + 1]])
    a1 = theta[0] * theta[0] * math.exp(-0.5 * float(np.dot(diff * inv, diff)))    # apply(vec1, vec2, sameIndex = false)
    return (diff * inv) * (-a1) if after_first_arg else (diff * inv) * a1


def lit_ucb_with_grad(X, L, alpha, theta, x, k_param):
    """GPO:87-104: the objective of maximizeUCB at one test point -> (ucb, ucbDer[D], mean, sigma)."""
    X = _f(X); L = _f(L)
    x = np.asarray(x, dtype=np.float64)
    n, D = X.shape
    inversedL = lit_inv_triangular(L, is_upper=False)                                   # GPO:85
    mean, sigma, V = lit_compute_posterior(X, x.reshape(1, D), L, alpha, theta)         # GPO:89-90
    ucb = mean[0] + k_param * math.sqrt(sigma[0, 0])                                    # GPO:93
    testTrain = np.zeros((D, n)); trainTest = np.zeros((n, D))
    for i in range(n):                                                                  # GPO:111-127
        testTrain[:, i] = lit_kernel_gradient(theta, x, X[i], True)
        trainTest[i, :] = lit_kernel_gradient(theta, X[i], x, False)
    derAfterMean = testTrain @ alpha                                                    # GPO:96
    derAfterVarFirst = lit_kernel_gradient(theta, x, x, True)                           # GPO:97
    vAfterX = inversedL @ trainTest                                                     # GPO:98
    derAfterVar = derAfterVarFirst - (vAfterX.T @ V[:, 0]) * 2.0                        # GPO:100
    coeff = k_param / (2 * math.sqrt(sigma[0, 0]))                                      # GPO:101
    return ucb, derAfterMean + derAfterVar * coeff, mean[0], sigma[0, 0]                # GPO:102-103


def pnorm(z):
    """StatsUtils.scala:17 (Breeze Gaussian(0,1).cdf) restated as 0.5*erfc(-z/sqrt 2)."""
    return _L().orc_pnorm(C.c_double(z))


def dnorm(z):
    """StatsUtils.scala:15."""
    return _L().orc_dnorm(C.c_double(z))


def lit_ep_estimate(K, targets, eps=0.01, fixed_sweeps=0, max_sweeps=100, keep_quirk=True):
    """EP:29-69.  Returns dict(tau, nu, L, Sigma, mu, cav_tau, cav_nu, logZ, sweeps)."""
    K = _f(K)
    n = K.shape[0]
    t = np.ascontiguousarray(targets, dtype=np.int32)
    tau = np.zeros(n); nu = np.zeros(n); mu = np.zeros(n)
    ct = np.zeros(n); cn = np.zeros(n)
    Lo = np.zeros((n, n), order="F"); Sig = np.zeros((n, n), order="F")
    logz = C.c_double(); sw = C.c_int()
    _check(_L().orc_ep_estimate(_p(K), n, t.ctypes.data_as(_ip), C.c_double(eps), int(fixed_sweeps),
                                int(max_sweeps), int(keep_quirk), _p(tau), _p(nu), _p(Lo), _p(Sig),
                                _p(mu), _p(ct), _p(cn), C.byref(logz), C.byref(sw)))
    return dict(tau=tau, nu=nu, L=Lo, Sigma=Sig, mu=mu, cav_tau=ct, cav_nu=cn, logZ=logz.value,
                sweeps=sw.value)


def lit_ep_classify(K, Ks, Kss, tau, nu, L):
    """GPC:24-47 -> (prob[m], fmean[m], fvar_diag[m])."""
    K = _f(K); Ks = _f(Ks); Kss = _f(Kss); L = _f(L)
    n = K.shape[0]; m = Ks.shape[0]
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    prob = np.zeros(m); fm = np.zeros(m); fv = np.zeros(m)
    _L().orc_ep_classify(_p(K), n, _p(Ks), m, _p(Kss), _p(tau), _p(nu), _p(L), _p(prob), _p(fm), _p(fv))
    return prob, fm, fv


def lit_ep_loglik_derivs(X, theta, K, tau, nu, L, keep_quirk=True):
    """MLE2:47-66 logLikelihoodDerivativesAfterHyperParams, statement by statement, on the literal solves / builders.
    keep_quirk=True reproduces the reference AS COMPILED: `val rMatrix = (bVector * bVector.t)` ends at the newline
    (MLE2:58), the following `- backSolve(...)` line is a discarded unary-minus expression, so rMatrix = b b^t.
    keep_quirk=False subtracts that term (what the source text seems to intend).  The missing inner forward solve of
    R&W Alg. 5.2 (MLE2:53-56) is reproduced in both modes: it is ordinary code, not a parsing accident."""
    X = _f(X); K = _f(K); L = _f(L)
    tau = np.asarray(tau, dtype=np.float64); nu = np.asarray(nu, dtype=np.float64)
    n, D = X.shape
    st = np.sqrt(tau)                                   # MLE2:51
    S = np.diag(st)                                     # MLE2:52
    temp = lit_back_solve(np.asfortranarray(L.T), (S @ K) @ nu)              # MLE2:53-54  (R = lowerTriangular.t)
    b = nu - lit_forward_solve(np.asfortranarray(S @ L), temp)               # MLE2:55-56
    R = np.outer(b, b)                                                       # MLE2:58
    if not keep_quirk:
        temp1 = lit_forward_solve(L, np.asfortranarray(S))                   # MLE2:57
        R = R - lit_back_solve(np.asfortranarray(S @ L.T), temp1)            # MLE2:59
    g = np.zeros(D + 2)
    for p in range(D + 2):                                                   # MLE2:60-65
        Cm = lit_build_der_matrix(p + 1, X, theta)
        g[p] = 0.5 * np.trace(R @ Cm)
    return g


def lit_ep_loglik_with_derivs(X, targets, theta, eps=0.01, fixed_sweeps=0, max_sweeps=100, keep_quirk=True):
    """MLE2:33-45 logLikelihood -> (logZ, gradient, ep dict)."""
    K = lit_build_kernel_matrix(X, theta)
    o = lit_ep_estimate(K, targets, eps=eps, fixed_sweeps=fixed_sweeps, max_sweeps=max_sweeps, keep_quirk=keep_quirk)
    g = lit_ep_loglik_derivs(X, theta, K, o["tau"], o["nu"], o["L"], keep_quirk=keep_quirk)
    return o["logZ"], g, o


# --------------------------------------------------------------------------------------------
# fast flavour (NumPy / SciPy LAPACK) -- same maths, used at full BASELINE sizes and as the timed
# CPU "port" baseline.
# --------------------------------------------------------------------------------------------
def _scaled_sqdist(X1, X2, ls):
    """KR:109-113 association: sum_d ((x1_d - x2_d) * (1/(ls_d*ls_d))) * (x1_d - x2_d)."""
    r = np.zeros((X1.shape[0], X2.shape[0]))
    for d in range(X1.shape[1]):
        diff = X1[:, d][:, None] - X2[:, d][None, :]
        inv = 1.0 / (ls[d] * ls[d])
        r += (diff * inv) * diff
    return r


def fast_build_kernel_matrix(X, theta, X2=None):
    """MU:57-70 / MU:44-55 vectorised."""
    X = np.asarray(X, dtype=np.float64)
    D = X.shape[1]
    sf, ls, sn = theta[0], theta[1:D + 1], theta[D + 1]
    if X2 is None:
        K = sf * sf * np.exp(-0.5 * _scaled_sqdist(X, X, ls))
        K[np.diag_indices_from(K)] += sn * sn
        return K
    return sf * sf * np.exp(-0.5 * _scaled_sqdist(X, np.asarray(X2, dtype=np.float64), ls))


def fast_precompute(X, y, theta, sigma_noise=None):
    """GPP:104-124 via dpotrf + dtrsv."""
    import scipy.linalg as sla
    K = fast_build_kernel_matrix(X, theta)
    if sigma_noise is not None:
        K[np.diag_indices_from(K)] += sigma_noise
    try:
        L = sla.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    except np.linalg.LinAlgError as e:  # pragma: no cover
        raise NotPositiveDefinite(-1) from e
    t = sla.solve_triangular(L, y, lower=True, check_finite=False)
    alpha = sla.solve_triangular(L, t, lower=True, trans="T", check_finite=False)
    return L, alpha


def fast_loglik(alpha, L, y):
    """GPP:144-149."""
    n = L.shape[0]
    return -0.5 * float(np.dot(y, alpha)) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * math.log(2 * math.pi)


def fast_loglik_with_derivs(X, y, theta, sigma_noise=None, nparams=None, block=1024):
    """GPP:60-80 with K^-1 from dpotri and the gradient trace fused (dK never materialised):
    g_p = 0.5 * sum_ij (alpha_i alpha_j - Kinv_ij) * dk_p(x_i, x_j, i==j)   (KR:76-86)."""
    import scipy.linalg as sla
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n, D = X.shape
    nparams = D + 2 if nparams is None else nparams
    sf, ls, sn = theta[0], theta[1:D + 1], theta[D + 1]
    L, alpha = fast_precompute(X, y, theta, sigma_noise)
    ll = fast_loglik(alpha, L, y)
    Kinv, info = sla.lapack.dpotri(L, lower=1)
    assert info == 0
    # dpotri fills the lower triangle only; symmetrise by blocks while accumulating.  The block pairs are independent: they
    # run on a thread pool (NumPy releases the GIL inside its ufuncs) so that the timed CPU baseline uses every host core in
    # this leg too, and their partial gradients are added in block order (deterministic).
    pairs = [(i0, j0) for i0 in range(0, n, block) for j0 in range(0, min(n, i0 + block), block)]

    def one(pair):
        i0, j0 = pair
        i1, j1 = min(n, i0 + block), min(n, j0 + block)
        gb = np.zeros(D + 2)
        Wb = np.outer(alpha[i0:i1], alpha[j0:j1]) - Kinv[i0:i1, j0:j1]
        if i0 == j0:
            Wl = np.tril(Wb, -1)
            Wb = Wl + Wl.T + np.diag(np.diag(Wb))
            mult = 1.0
        else:
            mult = 2.0
        E = np.exp(-0.5 * _scaled_sqdist(X[i0:i1], X[j0:j1], ls))
        WE = Wb * E
        gb[0] = mult * (2 * sf) * WE.sum()
        for d in range(D):
            diff = X[i0:i1, d][:, None] - X[j0:j1, d][None, :]
            gb[1 + d] = mult * (sf * sf) * float((WE * (diff * diff)).sum()) * ls[d] ** -3
        if i0 == j0:
            gb[D + 1] = (2 * sn) * float(np.trace(Wb))
        return gb

    workers = min(len(pairs), os.cpu_count() or 1)
    if workers > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=workers) as ex:
            parts = list(ex.map(one, pairs))
    else:
        parts = [one(pq) for pq in pairs]
    g = np.zeros(D + 2)
    for gb in parts:
        g += gb
    return ll, 0.5 * g[:nparams]


def fast_compute_posterior(X, Xs, L, alpha, theta, full_cov=True):
    """GPP:45-58 via dtrsm; sigma diagonal includes sn^2 (MU:63)."""
    import scipy.linalg as sla
    Ks = fast_build_kernel_matrix(Xs, theta, X)           # m x n
    mean = Ks @ alpha
    V = sla.solve_triangular(L, Ks.T, lower=True, check_finite=False)
    if full_cov:
        sigma = fast_build_kernel_matrix(Xs, theta) - V.T @ V
    else:
        D = X.shape[1]
        sigma = (theta[0] ** 2 + theta[D + 1] ** 2) - np.einsum("ij,ij->j", V, V)
    return mean, sigma, V


def fast_predict(X, y, Xs, theta, sigma_noise=None):
    """GPP:24-43."""
    L, alpha = fast_precompute(X, y, theta, sigma_noise)
    mean, sigma, _ = fast_compute_posterior(X, Xs, L, alpha, theta)
    if sigma_noise is not None:
        sigma = sigma + sigma_noise * np.eye(sigma.shape[0])
    return mean, sigma, fast_loglik(alpha, L, y)


def _pnorm_vec(z):
    from scipy.special import erfc
    return 0.5 * erfc(-np.asarray(z) / math.sqrt(2.0))


def fast_ep_estimate(K, targets, eps=0.01, fixed_sweeps=0, max_sweeps=100, keep_quirk=True):
    """EP:29-69 with vectorised rank-1 downdates and LAPACK re-factorisation (same recurrences;
    only mu_i, which is all the next site reads, is refreshed inside the site loop)."""
    import scipy.linalg as sla
    from scipy.special import erfc
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    y = np.asarray(targets, dtype=np.int64)
    tau = np.zeros(n); nu = np.zeros(n); mu = np.zeros(n)
    ct = np.zeros(n); cn = np.zeros(n)
    Sigma = K.copy()
    L = None
    sweeps = 0
    tau_old = tau.copy(); nu_old = nu.copy()
    j = 0
    while True:
        if fixed_sweeps > 0:
            if j >= fixed_sweeps:
                break
        elif j > 0:
            avg = float(np.sum((nu - nu_old) + (tau - tau_old))) / 2 * n
            if abs(avg) < eps or j >= max_sweeps:
                break
        tau_old = tau.copy(); nu_old = nu.copy()
        for i in range(n):
            sii = Sigma[i, i]
            ct[i] = 1 / sii - tau[i]
            cn[i] = mu[i] / sii - nu[i]
            cmu, csig = cn[i] / ct[i], 1 / ct[i]
            temp = math.sqrt(1 + csig)
            z = (y[i] * cmu) / temp
            dn = math.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)
            pn = 0.5 * float(erfc(-z / math.sqrt(2.0)))
            mu_hat = cmu + (y[i] * csig * dn) / (pn * temp)
            sig_hat = csig - ((csig * csig * dn) * (z + dn / pn)) / ((1 + csig) * pn)
            dtau = 1 / sig_hat - ct[i] - tau[i]
            tau[i] = tau[i] + dtau
            nu[i] = mu_hat / sig_hat - cn[i]
            s = Sigma[:, i].copy()
            Sigma -= np.outer(s, s) * (1 / (1 / dtau + sii))
            mu = Sigma @ nu
        st = np.sqrt(tau)
        B = np.eye(n) + np.outer(st, st) * K
        L = sla.cholesky(B, lower=True, check_finite=False)
        V = sla.solve_triangular(L, st[:, None] * K, lower=True, check_finite=False)
        Sigma = K - V.T @ V
        mu = Sigma @ nu
        sweeps += 1
        j += 1
    cav_mu = cn / ct
    sum_inv = 1 / (tau + ct)
    first = float(nu @ (Sigma - np.diag(sum_inv)) @ nu)
    second = float(((cav_mu * ct) * sum_inv) @ ((tau * cav_mu) - nu * 2.0))
    third = float(np.sum(np.log(_pnorm_vec(y * cav_mu / np.sqrt(1 + 1 / ct)))))
    fourth = 0.0 if keep_quirk else float(np.sum(0.5 * np.log(1 + tau / ct) - np.log(np.diag(L))))
    logz = third + fourth + 0.5 * (first + second)
    return dict(tau=tau, nu=nu, L=L, Sigma=Sigma, mu=mu, cav_tau=ct, cav_nu=cn, logZ=logz, sweeps=sweeps)


def fast_ep_classify(K, Ks, Kss, tau, nu, L):
    """GPC:24-47."""
    import scipy.linalg as sla
    st = np.sqrt(tau)
    rhs = (K * st[:, None]) @ nu
    t1 = sla.solve_triangular(L, rhs, lower=True, check_finite=False)
    z = st * sla.solve_triangular(L, t1, lower=True, trans="T", check_finite=False)
    fmean = Ks @ (nu - z)
    V = sla.solve_triangular(L, Ks.T * st[:, None], lower=True, check_finite=False)
    fvar = np.diag(Kss) - np.einsum("ij,ij->j", V, V)
    return _pnorm_vec(fmean / np.sqrt(1 + fvar)), fmean, fvar


# --------------------------------------------------------------------------------------------
# synthetic workloads of SURVEY.md 8(d) (seeds fixed there)
# --------------------------------------------------------------------------------------------
def fast_ep_loglik_derivs(X, theta, K, tau, nu, L):
    """MLE2:47-66 as compiled (rMatrix = b b^t), LAPACK solves, dK/dtheta in closed form: g_p = 1/2 b^t C_p b."""
    import scipy.linalg as sla
    X = np.asarray(X, dtype=np.float64)
    n, D = X.shape
    sf, ls, sn = theta[0], np.asarray(theta[1:D + 1]), theta[D + 1]
    st = np.sqrt(tau)
    temp = sla.solve_triangular(L, st * (K @ nu), lower=True, trans="T", check_finite=False)
    b = nu - sla.solve_triangular(L, temp / st, lower=True, check_finite=False)
    E = np.exp(-0.5 * _scaled_sqdist(X, X, ls))
    g = np.zeros(D + 2)
    g[0] = 0.5 * 2.0 * sf * (b @ E @ b)
    W = E * np.outer(b, b)
    for d in range(D):
        diff = X[:, d][:, None] - X[:, d][None, :]
        g[1 + d] = 0.5 * sf * sf * np.sum(W * diff * diff) / ls[d] ** 3
    g[D + 1] = 0.5 * 2.0 * sn * (b @ b)
    return g


# --------------------------------------------------------------------------------------------
# GP-UKF (dynamicalsystems/filtering/UnscentedKalmanFilter.scala, GPUnscentedKalmanFilter.scala): restated per sigma
# point and per output dimension exactly like the reference drives GpPredictor.computePosterior (m = 1 every time).
# --------------------------------------------------------------------------------------------
def ukf_unscented_transform(mean, cov, alpha, beta, kappa, func):
    """UKF:82-118 -> (finalMean, finalCov, (w_0_m, w_0_c, w_i_c), sigmaPoints, transformedSigmaPoints); func: vector -> vector."""
    mean = np.asarray(mean, dtype=np.float64); cov = np.asarray(cov, dtype=np.float64)
    d = len(mean)
    L = np.linalg.cholesky(cov)
    lam = alpha * alpha * (d + kappa) - d
    sp = np.zeros((2 * d + 1, d)); sp[0] = mean
    for col in range(d):
        c = L[:, col] * math.sqrt(d + lam)
        sp[col + 1] = mean + c
        sp[col + 1 + d] = mean - c
    w0m = lam / (d + lam); w0c = (lam / (d + lam)) + (1 - alpha * alpha + beta); wic = 1 / (2 * (d + lam))
    first = np.asarray(func(sp[0]), dtype=np.float64)
    tsp = np.zeros((2 * d + 1, len(first))); tsp[0] = first
    fm = first * w0m
    for i in range(1, 2 * d + 1):
        tsp[i] = func(sp[i])
        fm = fm + tsp[i] * wic
    diff = tsp[0] - fm
    fc = np.outer(diff, diff) * w0c
    for i in range(1, 2 * d + 1):
        diff = tsp[i] - fm
        fc = fc + np.outer(diff, diff) * wic
    return fm, fc, (w0m, w0c, wic), sp, tsp


def _log_gaussian_density(at, means, covs):
    """StatsUtils.scala:45-56."""
    d = len(means)
    diff = np.asarray(at) - np.asarray(means)
    sign, logdet = np.linalg.slogdet(covs)
    dens = (2 * math.pi) ** (-0.5 * d) * (sign * math.exp(logdet)) ** -0.5 * math.exp(-0.5 * float(diff @ np.linalg.solve(covs, diff)))
    return math.log(dens) if dens > 0 else -math.inf


def gpukf_infer(hidden, observations, theta, init_mean, init_cov, alpha=1.0, beta=0.0, kappa=2.0):
    """GPUKF:63-147 (learn one GP per state / observation dimension with fixed hyper-parameters) + UKF:24-80.
    hidden: d_state x tMax sampled trajectory, observations: d_obs x tMax.  -> (hiddenMeans, hiddenCovs[list], ll)."""
    hidden = np.asarray(hidden, dtype=np.float64); obs = np.asarray(observations, dtype=np.float64)
    Xall = np.ascontiguousarray(hidden.T); Xprev = np.ascontiguousarray(Xall[:-1])
    diffs = hidden[:, 1:] - hidden[:, :-1]
    sys_lc = [fast_precompute(Xprev, diffs[dim], theta) for dim in range(hidden.shape[0])]      # GPUKF:105-121
    obs_lc = [fast_precompute(Xall, obs[dim], theta) for dim in range(obs.shape[0])]

    def post(X, lc, x):                                                                           # computePosterior, m = 1
        m, S, _ = fast_compute_posterior(X, np.asarray(x, dtype=np.float64).reshape(1, -1), lc[0], lc[1], theta)
        return m[0], S[0, 0]

    trans = lambda p: np.asarray(p) + np.array([post(Xprev, lc, p)[0] for lc in sys_lc])         # GPUKF:77-83
    obsf = lambda p: np.array([post(Xall, lc, p)[0] for lc in obs_lc])                           # GPUKF:84-90
    tMax, hid = obs.shape[1], len(init_mean)
    means = np.zeros((hid, tMax)); covs = [None] * tMax
    means[:, 0] = init_mean; covs[0] = np.asarray(init_cov, dtype=np.float64)
    ll = 0.0
    for t in range(1, tMax):                                                                       # UKF:38-78
        m1, c1, w, _, zT = ukf_unscented_transform(means[:, t - 1], covs[t - 1], alpha, beta, kappa, trans)
        Q = np.diag([post(Xprev, lc, means[:, t - 1])[1] for lc in sys_lc])                      # GPUKF:95-98
        mz, cz = m1, c1 + Q
        m2, c2, _, _, yT = ukf_unscented_transform(mz, cz, alpha, beta, kappa, obsf)
        R = np.diag([post(Xall, lc, m1)[1] for lc in obs_lc])                                    # GPUKF:99-102
        my, S = m2, c2 + R
        zy = np.outer(zT[0] - mz, yT[0] - my) * w[1]
        for i in range(1, 2 * hid + 1):
            zy = zy + np.outer(zT[i] - mz, yT[i] - my) * w[2]
        K = zy @ np.linalg.inv(S)
        means[:, t] = mz + K @ (obs[:, t] - my)
        covs[t] = cz - (K @ S) @ K.T
        ll += _log_gaussian_density(obs[:, t], my, S)
    return means, covs, ll


def make_ssm_series(tMax=80, seed=7):
    """A small nonlinear 2-state / 2-observation state-space series for the GP-UKF tests (synthetic, seeded)."""
    rng = np.random.default_rng(seed)
    h = np.zeros((2, tMax)); o = np.zeros((2, tMax))
    h[:, 0] = rng.standard_normal(2) * 0.5
    for t in range(tMax):
        o[:, t] = np.array([np.sin(h[0, t]) + 0.5 * h[1, t], 0.3 * h[0, t] * h[1, t] + h[1, t]]) + 0.05 * rng.standard_normal(2)
        if t + 1 < tMax:
            h[:, t + 1] = np.array([0.9 * h[0, t] + 0.2 * np.sin(h[1, t]), 0.8 * h[1, t] + 0.3 * np.cos(h[0, t])]) + 0.1 * rng.standard_normal(2)
    return h, o


def make_c1(n=1000, m=500, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 10, size=(n, 1))
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(n)
    xs = np.linspace(0, 10, m)[:, None]
    return x, y, xs, pack_theta(1.0, [1.0], 0.1)


def make_c2(n=8192, D=8, seed=2):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, size=(n, D))
    w = rng.standard_normal(D)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    return X, y, pack_theta(1.0, [0.7] * D, 0.1)


def make_c3(n=4096, D=4, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    w = rng.standard_normal(D)
    t = np.sign(X @ w + 0.3 * rng.standard_normal(n)).astype(np.int32)
    t[t == 0] = 1
    return X, t, pack_theta(1.0, [1.0] * D, 0.0)


def make_c4_problem(b, n=1024, D=8, m=17):
    rng = np.random.default_rng(1000 + b)
    X = rng.uniform(0, 1, size=(n, D))
    w = rng.standard_normal(D)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    sf = 10 ** rng.uniform(-0.3, 0.3)
    ls = 10 ** rng.uniform(-0.5, 0.2, size=D)
    Xs = rng.uniform(0, 1, size=(m, D))
    return X, y, Xs, pack_theta(sf, ls, 0.1)

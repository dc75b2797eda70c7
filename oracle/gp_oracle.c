/*
 * gp_oracle.c -- CPU restatement ("literal" flavour) of the dense-GP hot path of
 * astroHaoPeng/gp_algos.  TEST INFRASTRUCTURE ONLY: nothing under gp_algos_b200/ may
 * import, link or execute this file.  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the CUDA path.
 *
 * PARITY STATUS: the reference (Scala 2.10 + Breeze 0.8.1 + netlib-java 1.1.2) cannot be
 * run in this environment (no JVM), and its own tests pin only: the hyper-parameter packing
 * (src/test/scala/utils/KernelRequisitesTest.scala:18-49), K_ii == 1.0 for a 3x3 SE kernel
 * matrix and "cholesky does not throw" (src/test/scala/utils/MatrixUtilsTest.scala:90-102),
 * 3x3 triangular solves to 1e-3 (MatrixUtilsTest.scala:27-65) and (L^-1)^T L^-1 ~= inv(K)
 * to 1e-3 (MatrixUtilsTest.scala:104-114).  Those are checked in tests/test_oracle_pins.py.
 * Everything else (log-likelihood, gradient, predictive moments, the whole EP path) is
 * "PARITY UNPINNED" by the reference; it is defended instead by an mpmath 50-digit arbiter,
 * finite-difference gradient checks and a literal-vs-LAPACK cross-check (tests/).
 *
 * The reference's second closed-form kernel, Co2Kernel (gp/regression/Co2Prediction.scala:29-137), is restated further down;
 * its only pin from the reference is the shipped output src/main/resources/co2/co2PredResults.txt (soft, ~1e-4 relative on the
 * training range; tests/test_co2_oracle_and_host.py), everything else about it is "parity unpinned" as well.
 *
 * Third-party arithmetic that is NOT in /root/reference and is restated here:
 *   - Breeze 0.8.1 `cholesky`  -> LAPACK dpotrf('L') : unblocked dpotf2 recurrence below.
 *   - Breeze `*`, `trace`, `dot` -> plain ascending-index dot products.
 *   - breeze.stats.distributions.Gaussian(0,1).cdf/.pdf -> 0.5*erfc(-z/sqrt2), exp(-z^2/2)/sqrt(2pi).
 *
 * All matrices are column-major (Breeze DenseMatrix layout) with explicit leading dimension.
 * Every function cites the reference file:line it follows; paths are relative to
 * /root/reference/src/main/scala/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_EINVAL -1
#define ORC_ENOTSYM -2
/* not positive definite: returns +minor index (1-based), like LAPACK info */

/* ---- utils/KernelRequisites.scala:109-113  inputWithLsProduct ------------------------- */
/* (diff :* inversedSqLs) dot diff, inversedSqLs(d) = 1./(ls*ls); summed d = 0..D-1. */
static double orc_scaled_sqdist(const double *x1, long inc1, const double *x2, long inc2,
                                int D, const double *ls)
{
    double acc = 0.0;
    for (int d = 0; d < D; ++d) {
        double diff = x1[d * inc1] - x2[d * inc2];
        double inv = 1.0 / (ls[d] * ls[d]);
        acc += (diff * inv) * diff;
    }
    return acc;
}

/* ---- utils/KernelRequisites.scala:66-72  GaussianRbfKernel.apply ----------------------- */
double orc_rbf_k(const double *x1, long inc1, const double *x2, long inc2, int D,
                 double sf, const double *ls, double sn, int same_index)
{
    double v = sf * sf * exp(-0.5 * orc_scaled_sqdist(x1, inc1, x2, inc2, D, ls));
    return same_index ? v + sn * sn : v;
}

/* ---- utils/KernelRequisites.scala:76-86  derAfterHyperParam (param_num is 1-based) ------ */
double orc_rbf_dk(int param_num, const double *x1, long inc1, const double *x2, long inc2,
                  int D, double sf, const double *ls, double sn, int same_index)
{
    if (param_num == 1)
        return 2 * sf * exp(-0.5 * orc_scaled_sqdist(x1, inc1, x2, inc2, D, ls));
    if (param_num < D + 2) {
        int d = param_num - 2;
        double diff = x1[d * inc1] - x2[d * inc2];
        return pow(sf, 2) * exp(-0.5 * orc_scaled_sqdist(x1, inc1, x2, inc2, D, ls)) *
               pow(diff, 2) * pow(ls[d], -3);
    }
    if (param_num == D + 2)
        return same_index ? 2 * sn : 0.0;
    return NAN; /* scala.MatchError in the reference (KernelRequisitesTest.scala:32-34) */
}

/* ---- utils/KernelRequisites.scala:99-107  gradient(afterFirstArg) ----------------------- */
void orc_rbf_grad_x(int after_first_arg, const double *x1, long inc1, const double *x2, long inc2,
                    int D, double sf, const double *ls, double sn, double *out)
{
    double a1 = orc_rbf_k(x1, inc1, x2, inc2, D, sf, ls, sn, 0);
    for (int d = 0; d < D; ++d) {
        double diff = x1[d * inc1] - x2[d * inc2];
        double inv = 1.0 / (ls[d] * ls[d]);
        out[d] = after_first_arg ? (diff * inv) * (-a1) : (diff * inv) * a1;
    }
}


/* ---- gp/regression/Co2Prediction.scala:29-137  Co2Kernel (1-D inputs, 11 hyper-parameters) ----------
 * The reference's second KernelFunc (R&W's Mauna Loa kernel); it flows through the same GpPredictor code
 * (changeHyperParams / apply / derAfterHyperParam).  hp[0..10] = hp1..hp11 (getAtPosition is 1-based, :23).
 * Expressions keep the Scala evaluation order (left-associative products and sums).  breeze.numerics
 * exp/sin/pow are java.lang.Math on doubles -> libm here.
 * Kernel family switch: orc_set_kernel(0) = GaussianRbfKernel (theta = [sf, ls.., sn], D+2 entries),
 * orc_set_kernel(1) = Co2Kernel (theta = hp1..hp11, D must be 1).  The matrix builders below dispatch on it,
 * so every GpPredictor-level routine of this file serves both kernels, like the Scala class does. */
static int orc_kernel_kind = 0;
void orc_set_kernel(int kind) { orc_kernel_kind = kind; }
int orc_get_kernel(void) { return orc_kernel_kind; }

/* Co2Prediction.scala:38-56 apply */
double orc_co2_k(double x1, double x2, const double *hp, int same_index)
{
    double hp1 = hp[0], hp2 = hp[1], hp3 = hp[2], hp4 = hp[3], hp5 = hp[4], hp6 = hp[5], hp7 = hp[6],
           hp8 = hp[7], hp9 = hp[8], hp10 = hp[9], hp11 = hp[10];
    double xDiff = x1 - x2, xDiffSq = (x1 - x2) * (x1 - x2);
    double k1Val = hp1 * hp1 * exp(-xDiffSq / (2 * hp2 * hp2));
    double sinVal = sin(M_PI * xDiff);
    double k2Val = hp3 * hp3 * exp((-xDiffSq / (2 * hp4 * hp4)) - 2 * sinVal * sinVal / (hp5 * hp5));
    double k3Pow1 = 1 + xDiffSq / (2 * hp8 * hp7 * hp7);
    double k3Val = hp6 * hp6 * pow(k3Pow1, -hp8);
    double k4Val = hp9 * hp9 * exp(-xDiffSq / (2 * hp10 * hp10));
    double indNoise = same_index ? hp11 * hp11 : 0.;
    return k1Val + k2Val + k3Val + k4Val + indNoise;
}

/* Co2Prediction.scala:66-137 derAfterHyperParam (param_num 1-based; outside 1..11 is a scala.MatchError -> NAN) */
double orc_co2_dk(int param_num, double x1, double x2, const double *hp, int same_index)
{
    double hp1 = hp[0], hp2 = hp[1], hp3 = hp[2], hp4 = hp[3], hp5 = hp[4], hp6 = hp[5], hp7 = hp[6],
           hp8 = hp[7], hp9 = hp[8], hp10 = hp[9], hp11 = hp[10];
    double xDiff = x1 - x2, sqDiff = (x1 - x2) * (x1 - x2);
    if (param_num < 1 || param_num > 11) return NAN;
    if (param_num < 3) {                                             /* derAfterFirstKernel :93-99 */
        if (param_num == 1) return 2 * hp1 * exp(-sqDiff / (2 * hp2 * hp2));
        return hp1 * hp1 * exp(-sqDiff / (2 * hp2 * hp2)) * sqDiff * pow(hp2, -3);
    }
    if (param_num < 6) {                                             /* derAfterSecondKernel :101-110 */
        double sinVal = sin(M_PI * xDiff);
        double k2Val = hp3 * hp3 * exp(-sqDiff / (2 * hp4 * hp4) - 2 * sinVal * sinVal / (hp5 * hp5));
        if (param_num == 3) return 2 * k2Val / hp3;
        if (param_num == 4) return k2Val * sqDiff * pow(hp4, -3);
        return k2Val * 4 * sinVal * sinVal * pow(hp5, -3);
    }
    if (param_num < 9) {                                             /* derAfterThirdKernel :112-123 */
        double k3Pow1 = 1 + sqDiff / (2 * hp8 * hp7 * hp7);
        if (param_num == 6) return 2 * hp6 * pow(k3Pow1, -hp8);
        if (param_num == 7) return hp6 * hp6 * pow(k3Pow1, -hp8 - 1) * sqDiff * pow(hp7, -3);
        double firstTerm = exp(-hp8 * log(k3Pow1));
        double secondTerm = -log(k3Pow1) + (hp8 * sqDiff / (2 * hp7 * hp7 * hp8 * hp8 * k3Pow1));
        return hp6 * hp6 * firstTerm * secondTerm;
    }
    {                                                                /* derAfterFourthKernel :125-135 */
        double k4Val = hp9 * hp9 * exp(-sqDiff / (2 * hp10 * hp10));
        if (param_num == 9) return 2 * k4Val / hp9;
        if (param_num == 10) return k4Val * sqDiff * pow(hp10, -3);
        return same_index ? 2 * hp11 : 0.;
    }
}

/* ---- utils/MatrixUtils.scala:57-70  buildKernelMatrix(kernel,data) ---------------------- */
/* theta = [sf, ls_1..ls_D, sn] (KernelRequisites.scala:39-58). X is n x D column-major. */
void orc_build_kernel_matrix(const double *X, int n, int D, long ldx, const double *theta,
                             double *K, long ldk)
{
    double sf = theta[0], sn = orc_kernel_kind ? 0.0 : theta[D + 1];
    const double *ls = theta + 1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double v = orc_kernel_kind ? orc_co2_k(X[i], X[j], theta, i == j)
                                       : orc_rbf_k(X + i, ldx, X + j, ldx, D, sf, ls, sn, i == j);
            K[i + (long)j * ldk] = v;
            K[j + (long)i * ldk] = v;
        }
}

/* ---- utils/MatrixUtils.scala:44-55,86-97  buildKernelMatrix(kernel,in1,in2) ------------- */
void orc_build_kernel_matrix_cross(const double *X1, int m, long ldx1, const double *X2, int n,
                                   long ldx2, int D, const double *theta, double *K, long ldk)
{
    double sf = theta[0], sn = orc_kernel_kind ? 0.0 : theta[D + 1];
    const double *ls = theta + 1;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j)
            K[i + (long)j * ldk] = orc_kernel_kind ? orc_co2_k(X1[i], X2[j], theta, 0)
                                                   : orc_rbf_k(X1 + i, ldx1, X2 + j, ldx2, D, sf, ls, sn, 0);
}

/* ---- utils/MatrixUtils.scala:72-84 with f = derAfterHyperParam(p) (GpPredictor.scala:72-75) */
void orc_build_der_matrix(int param_num, const double *X, int n, int D, long ldx,
                          const double *theta, double *dK, long ldk)
{
    double sf = theta[0], sn = orc_kernel_kind ? 0.0 : theta[D + 1];
    const double *ls = theta + 1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double v = orc_kernel_kind ? orc_co2_dk(param_num, X[i], X[j], theta, i == j)
                                       : orc_rbf_dk(param_num, X + i, ldx, X + j, ldx, D, sf, ls, sn, i == j);
            dK[i + (long)j * ldk] = v;
            dK[j + (long)i * ldk] = v;
        }
}

/* ---- utils/MatrixUtils.scala:123-133 solveTriangular (vector) --------------------------- */
/* trans == 0: M(r,c) = A[r + c*lda]; trans != 0: M(r,c) = A[c + r*lda] (the `L.t` view used at
 * GpPredictor.scala:122, GpClassifier.scala:36). forward: rows 0..n-1, cols 0..r-1 ascending
 * (MatrixUtils.scala:17-21). backward: rows n-1..0, cols n-1..r+1 DESCENDING (MatrixUtils.scala:23-27). */
static void orc_solve_tri_vec(const double *A, int n, long lda, int trans, int backward,
                              const double *b, long incb, double *x, long incx)
{
#define M_(r, c) (trans ? A[(c) + (long)(r) * lda] : A[(r) + (long)(c) * lda])
    if (!backward) {
        for (int r = 0; r < n; ++r) {
            double acc = 0.0;
            for (int c = 0; c < r; ++c) acc = acc + M_(r, c) * x[c * incx];
            x[r * incx] = (b[r * incb] - acc) / M_(r, r);
        }
    } else {
        for (int r = n - 1; r >= 0; --r) {
            double acc = 0.0;
            for (int c = n - 1; c > r; --c) acc = acc + M_(r, c) * x[c * incx];
            x[r * incx] = (b[r * incb] - acc) / M_(r, r);
        }
    }
#undef M_
}

/* utils/MatrixUtils.scala:17-21 */
void orc_forward_solve_vec(const double *L, int n, long ldl, int trans, const double *b, double *x)
{
    orc_solve_tri_vec(L, n, ldl, trans, 0, b, 1, x, 1);
}
/* utils/MatrixUtils.scala:23-27 */
void orc_back_solve_vec(const double *R, int n, long ldr, int trans, const double *b, double *x)
{
    orc_solve_tri_vec(R, n, ldr, trans, 1, b, 1, x, 1);
}
/* utils/MatrixUtils.scala:29-31,115-121 : column-by-column */
void orc_forward_solve_mat(const double *L, int n, long ldl, int trans, const double *B, int m,
                           long ldb, double *Xo, long ldxo)
{
    for (int c = 0; c < m; ++c)
        orc_solve_tri_vec(L, n, ldl, trans, 0, B + (long)c * ldb, 1, Xo + (long)c * ldxo, 1);
}
/* utils/MatrixUtils.scala:33-35,115-121 */
void orc_back_solve_mat(const double *R, int n, long ldr, int trans, const double *B, int m,
                        long ldb, double *Xo, long ldxo)
{
    for (int c = 0; c < m; ++c)
        orc_solve_tri_vec(R, n, ldr, trans, 1, B + (long)c * ldb, 1, Xo + (long)c * ldxo, 1);
}

/* ---- utils/MatrixUtils.scala:106-113 invTriangular: solve against a DENSE identity ------- */
void orc_inv_triangular(const double *A, int n, long lda, int is_upper, double *Ainv, long ldi)
{
    double *e = (double *)calloc((size_t)n, sizeof(double));
    for (int c = 0; c < n; ++c) {
        e[c] = 1.0;
        orc_solve_tri_vec(A, n, lda, 0, is_upper ? 1 : 0, e, 1, Ainv + (long)c * ldi, 1);
        e[c] = 0.0;
    }
    free(e);
}

/* ---- Breeze 0.8.1 `cholesky` (GpPredictor.scala:120, EpParameterEstimator.scala:58) ------ */
/* Restated from the published LAPACK algorithm dpotf2('L') (netlib-java 1.1.2 dpotrf): checks
 * exact symmetry first (Breeze MatrixNotSymmetricException), left-looking column recurrence,
 * strict upper triangle zeroed.  Returns 0, ORC_ENOTSYM, or j+1 for a non-positive pivot. */
int orc_cholesky_lower(const double *A, int n, long lda, double *L, long ldl)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (A[i + (long)j * lda] != A[j + (long)i * lda]) return ORC_ENOTSYM;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) L[i + (long)j * ldl] = (i >= j) ? A[i + (long)j * lda] : 0.0;
    for (int j = 0; j < n; ++j) {
        double ajj = L[j + (long)j * ldl];
        for (int k = 0; k < j; ++k) ajj -= L[j + (long)k * ldl] * L[j + (long)k * ldl];
        if (!(ajj > 0.0)) return j + 1;
        ajj = sqrt(ajj);
        L[j + (long)j * ldl] = ajj;
        for (int i = j + 1; i < n; ++i) {
            double v = L[i + (long)j * ldl];
            for (int k = 0; k < j; ++k) v -= L[i + (long)k * ldl] * L[j + (long)k * ldl];
            L[i + (long)j * ldl] = v / ajj;
        }
    }
    return 0;
}

/* ---- gp/regression/GpPredictor.scala:104-124 preComputeComponents ------------------------ */
/* has_sigma_noise/sigma_noise: Option[Double]; note it is added UN-squared (GpPredictor.scala:116-117)
 * on top of the kernel's own sn^2 diagonal. Outputs L (n x n, ld n) and alpha (n). */
int orc_gp_precompute(const double *X, int n, int D, long ldx, const double *y, const double *theta,
                      int has_sigma_noise, double sigma_noise, double *L, double *alpha)
{
    double *K = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    orc_build_kernel_matrix(X, n, D, ldx, theta, K, n);
    if (has_sigma_noise)
        for (int i = 0; i < n; ++i) K[i + (long)i * n] += sigma_noise;
    int info = orc_cholesky_lower(K, n, n, L, n);
    if (info == 0) {
        orc_forward_solve_vec(L, n, n, 0, y, tmp);       /* GpPredictor.scala:121 */
        orc_back_solve_vec(L, n, n, 1, tmp, alpha);      /* GpPredictor.scala:122 (R = L.t) */
    }
    free(K);
    free(tmp);
    return info;
}

/* ---- gp/regression/GpPredictor.scala:144-149 logLikelihood ------------------------------- */
double orc_gp_loglik(const double *alpha, const double *L, int n, long ldl, const double *y)
{
    double dot = 0.0;
    for (int i = 0; i < n; ++i) dot += y[i] * alpha[i];
    double a1 = -0.5 * dot;
    double a2 = 0.0;
    for (int i = 0; i < n; ++i) a2 = a2 + log(L[i + (long)i * ldl]);
    return a1 - a2 - 0.5 * n * log(2 * M_PI);
}

/* ---- gp/regression/GpPredictor.scala:60-80 logLikelihoodWithDerivatives ------------------ */
/* Literal route: explicit L^-1 (dense-identity solves), K^-1 = (L^-1)^T L^-1, alpha alpha^T,
 * one materialised dK per parameter, g_p = 0.5*trace((aa^T - K^-1) * dK_p).  Only the DIAGONAL
 * of the product is formed (ascending-k dot products, the entries a reference dgemm+trace would
 * read); the off-diagonal entries of the product never influence the result. */
int orc_gp_loglik_with_derivs(const double *X, int n, int D, long ldx, const double *y,
                              const double *theta, int has_sigma_noise, double sigma_noise,
                              int nparams, double *ll_out, double *grad_out)
{
    size_t nn = (size_t)n * n;
    double *L = (double *)malloc(sizeof(double) * nn);
    double *alpha = (double *)malloc(sizeof(double) * (size_t)n);
    int info = orc_gp_precompute(X, n, D, ldx, y, theta, has_sigma_noise, sigma_noise, L, alpha);
    if (info != 0) { free(L); free(alpha); return info; }
    *ll_out = orc_gp_loglik(alpha, L, n, n, y);
    double *Li = (double *)malloc(sizeof(double) * nn);
    double *W = (double *)malloc(sizeof(double) * nn);
    double *dK = (double *)malloc(sizeof(double) * nn);
    orc_inv_triangular(L, n, n, 0, Li, n);                               /* GpPredictor.scala:66 */
    for (int j = 0; j < n; ++j)                                           /* GpPredictor.scala:67 */
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += Li[k + (long)i * n] * Li[k + (long)j * n];
            /* GpPredictor.scala:69,76: alphaSq - inversedK */
            W[i + (long)j * n] = alpha[i] * alpha[j] - acc;
        }
    for (int p = 0; p < nparams; ++p) {                                   /* GpPredictor.scala:70-78 */
        orc_build_der_matrix(p + 1, X, n, D, ldx, theta, dK, n);
        double tr = 0.0;
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += W[i + (long)k * n] * dK[k + (long)i * n];
            tr += acc;
        }
        grad_out[p] = 0.5 * tr;
    }
    free(L); free(alpha); free(Li); free(W); free(dK);
    return 0;
}

/* ---- gp/regression/GpPredictor.scala:45-58 computePosterior ------------------------------ */
/* mean (m), sigma (m x m, ld m; diagonal INCLUDES sn^2 via MatrixUtils.scala:63), V (n x m, ld n). */
void orc_gp_compute_posterior(const double *X, int n, int D, long ldx, const double *Xs, int m,
                              long ldxs, const double *L, long ldl, const double *alpha,
                              const double *theta, double *mean, double *sigma, double *V)
{
    double *Ks = (double *)malloc(sizeof(double) * (size_t)m * n);   /* m x n */
    double *KsT = (double *)malloc(sizeof(double) * (size_t)m * n);  /* n x m */
    orc_build_kernel_matrix_cross(Xs, m, ldxs, X, n, ldx, D, theta, Ks, m);
    for (int i = 0; i < m; ++i) {                                     /* GpPredictor.scala:54 */
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += Ks[i + (long)j * m] * alpha[j];
        mean[i] = acc;
    }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) KsT[j + (long)i * n] = Ks[i + (long)j * m];
    orc_forward_solve_mat(L, n, ldl, 0, KsT, m, n, V, n);             /* GpPredictor.scala:55 */
    orc_build_kernel_matrix(Xs, m, D, ldxs, theta, sigma, m);         /* GpPredictor.scala:56 */
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += V[k + (long)i * n] * V[k + (long)j * n];
            sigma[i + (long)j * m] -= acc;
        }
    free(Ks); free(KsT);
}

/* ---- gp/regression/GpPredictor.scala:24-43 predict --------------------------------------- */
int orc_gp_predict(const double *X, int n, int D, long ldx, const double *y, const double *Xs, int m,
                   long ldxs, const double *theta, int has_sigma_noise, double sigma_noise,
                   double *mean, double *sigma, double *ll_out)
{
    double *L = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *alpha = (double *)malloc(sizeof(double) * (size_t)n);
    double *V = (double *)malloc(sizeof(double) * (size_t)n * m);
    int info = orc_gp_precompute(X, n, D, ldx, y, theta, has_sigma_noise, sigma_noise, L, alpha);
    if (info == 0) {
        orc_gp_compute_posterior(X, n, D, ldx, Xs, m, ldxs, L, n, alpha, theta, mean, sigma, V);
        if (has_sigma_noise)                                          /* GpPredictor.scala:37-39 */
            for (int i = 0; i < m; ++i) sigma[i + (long)i * m] += sigma_noise;
        *ll_out = orc_gp_loglik(alpha, L, n, n, y);                   /* GpPredictor.scala:41 */
    }
    free(L); free(alpha); free(V);
    return info;
}

/* ---- utils/StatsUtils.scala:13-17 pnorm/dnorm (Breeze Gaussian(0,1).cdf/.pdf) ------------- */
double orc_pnorm(double z) { return 0.5 * erfc(-z / M_SQRT2); }
double orc_dnorm(double z) { return exp(-0.5 * z * z) / sqrt(2 * M_PI); }

/* ---- gp/classification/EpParameterEstimator.scala:98-109 marginalMoments ------------------ */
static void orc_marginal_moments(double cav_mu, double cav_sigma, int target, double *mu_hat,
                                 double *sigma_hat)
{
    double temp = sqrt(1 + cav_sigma);
    double z = (target * cav_mu) / temp;
    double dn = orc_dnorm(z), pn = orc_pnorm(z);
    *mu_hat = cav_mu + (target * cav_sigma * dn) / (pn * temp);
    *sigma_hat = cav_sigma - ((cav_sigma * cav_sigma * dn) * (z + dn / pn)) / ((1 + cav_sigma) * pn);
}

/* ---- gp/classification/EpParameterEstimator.scala:195-202 avgBetweenSiteParams ----------- */
double orc_ep_avg_between(const double *nu_old, const double *tau_old, const double *nu, const double *tau,
                          int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = s + (nu[i] - nu_old[i]) + (tau[i] - tau_old[i]);
    return (s / 2 * n); /* precedence as written at EpParameterEstimator.scala:201 */
}

/* ---- gp/classification/EpParameterEstimator.scala:71-96 epMarginalLikelihood ------------- */
/* keep_linebreak_quirk != 0 reproduces the reference as compiled: the statement at :91 ends at the
 * newline, so the "fourth and first" term stays 0.  With 0 the (R&W eq. 3.65) term is included. */
double orc_ep_marginal_likelihood(const double *tau, const double *nu, const double *cav_tau,
                                  const double *cav_nu, const int *targets, const double *L,
                                  const double *Sigma, int n, int keep_linebreak_quirk)
{
    double *cav_mu = (double *)malloc(sizeof(double) * (size_t)n);
    double *sum_inv = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        cav_mu[i] = cav_nu[i] / cav_tau[i];
        sum_inv[i] = 1 / (tau[i] + cav_tau[i]);
    }
    /* (nu^T * (Sigma - diag(1/(tau+cavTau)))) * nu  (EpParameterEstimator.scala:81-82) */
    double first = 0.0;
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            double t2 = Sigma[i + (long)j * n] - (i == j ? sum_inv[i] : 0.0);
            acc += nu[i] * t2;
        }
        first += acc * nu[j];
    }
    double second = 0.0; /* temp3 dot temp4 (EpParameterEstimator.scala:83-85) */
    for (int i = 0; i < n; ++i) {
        double t3 = (cav_mu[i] * cav_tau[i]) * sum_inv[i];
        double t4 = (tau[i] * cav_mu[i]) - (nu[i] * 2.);
        second += t3 * t4;
    }
    double third = 0.0, fourth_first = 0.0;
    for (int i = 0; i < n; ++i) {
        third = third + log(orc_pnorm(targets[i] * cav_mu[i] / sqrt(1 + 1 / cav_tau[i])));
        if (!keep_linebreak_quirk)
            fourth_first = fourth_first + 0.5 * log(1 + tau[i] / cav_tau[i]) - log(L[i + (long)i * n]);
    }
    free(cav_mu); free(sum_inv);
    return third + fourth_first + 0.5 * (first + second);
}

/* ---- gp/classification/EpParameterEstimator.scala:29-69 estimateSiteParams --------------- */
/* Literal: per site a full rank-1 downdate of Sigma and a full mu = Sigma*nu; per sweep the
 * re-factorisation B = I + S^1/2 K S^1/2, V = L \ (S^1/2 K), Sigma = K - V^T V.
 * stop rule: fixed_sweeps > 0 -> run exactly that many sweeps; otherwise AvgBasedStopCriterion(eps)
 * (EpParameterEstimator.scala:187-193).  Outputs tau, nu (n), L (n x n), Sigma (n x n, optional),
 * mu (n, optional), cavity params (optional), logZ and the sweep count. */
int orc_ep_estimate(const double *K, int n, const int *targets, double eps, int fixed_sweeps,
                    int max_sweeps, int keep_linebreak_quirk, double *tau, double *nu, double *L,
                    double *Sigma_out, double *mu_out, double *cav_tau_out, double *cav_nu_out,
                    double *logz_out, int *sweeps_out)
{
    size_t nn = (size_t)n * n;
    double *Sigma = (double *)malloc(sizeof(double) * nn);
    double *B = (double *)malloc(sizeof(double) * nn);
    double *V = (double *)malloc(sizeof(double) * nn);
    double *mu = (double *)calloc((size_t)n, sizeof(double));
    double *cav_tau = (double *)calloc((size_t)n, sizeof(double));
    double *cav_nu = (double *)calloc((size_t)n, sizeof(double));
    double *tau_old = (double *)calloc((size_t)n, sizeof(double));
    double *nu_old = (double *)calloc((size_t)n, sizeof(double));
    double *s = (double *)malloc(sizeof(double) * (size_t)n);
    double *st = (double *)malloc(sizeof(double) * (size_t)n);
    int info = 0, sweeps = 0;
    memset(tau, 0, sizeof(double) * (size_t)n);
    memset(nu, 0, sizeof(double) * (size_t)n);
    memcpy(Sigma, K, sizeof(double) * nn);
    for (int j = 0;; ++j) {
        if (fixed_sweeps > 0) { if (j >= fixed_sweeps) break; }
        else if (j > 0 && (fabs(orc_ep_avg_between(nu_old, tau_old, nu, tau, n)) < eps || j >= max_sweeps)) break;
        memcpy(tau_old, tau, sizeof(double) * (size_t)n);
        memcpy(nu_old, nu, sizeof(double) * (size_t)n);
        for (int i = 0; i < n; ++i) {
            double sii = Sigma[i + (long)i * n];
            cav_tau[i] = 1 / sii - tau[i];                              /* :45 */
            cav_nu[i] = mu[i] / sii - nu[i];                            /* :46 */
            double mu_hat, sigma_hat;
            orc_marginal_moments(cav_nu[i] / cav_tau[i], 1 / cav_tau[i], targets[i], &mu_hat, &sigma_hat);
            double dtau = 1 / sigma_hat - cav_tau[i] - tau[i];          /* :49 */
            tau[i] = tau[i] + dtau;                                     /* :50 */
            nu[i] = mu_hat / sigma_hat - cav_nu[i];                     /* :51 */
            for (int a = 0; a < n; ++a) s[a] = Sigma[a + (long)i * n];  /* :52 */
            double c = 1 / (1 / dtau + sii);                            /* :53 */
            for (int b = 0; b < n; ++b)
                for (int a = 0; a < n; ++a) Sigma[a + (long)b * n] -= (s[a] * s[b]) * c;
            for (int a = 0; a < n; ++a) {                               /* :54 */
                double acc = 0.0;
                for (int b = 0; b < n; ++b) acc += Sigma[a + (long)b * n] * nu[b];
                mu[a] = acc;
            }
        }
        for (int i = 0; i < n; ++i) st[i] = sqrt(tau[i]);               /* :56 */
        for (int b = 0; b < n; ++b)                                     /* :58 */
            for (int a = 0; a < n; ++a)
                B[a + (long)b * n] = (a == b ? 1.0 : 0.0) + (st[a] * st[b]) * K[a + (long)b * n];
        info = orc_cholesky_lower(B, n, n, L, n);
        if (info != 0) break;
        for (int b = 0; b < n; ++b)                                     /* :59 rhs = cloneCols(st,n) :* K */
            for (int a = 0; a < n; ++a) B[a + (long)b * n] = st[a] * K[a + (long)b * n];
        orc_forward_solve_mat(L, n, n, 0, B, n, n, V, n);
        for (int b = 0; b < n; ++b)                                     /* :60 */
            for (int a = 0; a < n; ++a) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc += V[k + (long)a * n] * V[k + (long)b * n];
                Sigma[a + (long)b * n] = K[a + (long)b * n] - acc;
            }
        for (int a = 0; a < n; ++a) {                                   /* :61 */
            double acc = 0.0;
            for (int b = 0; b < n; ++b) acc += Sigma[a + (long)b * n] * nu[b];
            mu[a] = acc;
        }
        ++sweeps;
    }
    if (info == 0)
        *logz_out = orc_ep_marginal_likelihood(tau, nu, cav_tau, cav_nu, targets, L, Sigma, n,
                                               keep_linebreak_quirk);  /* :67 */
    *sweeps_out = sweeps;
    if (Sigma_out) memcpy(Sigma_out, Sigma, sizeof(double) * nn);
    if (mu_out) memcpy(mu_out, mu, sizeof(double) * (size_t)n);
    if (cav_tau_out) memcpy(cav_tau_out, cav_tau, sizeof(double) * (size_t)n);
    if (cav_nu_out) memcpy(cav_nu_out, cav_nu, sizeof(double) * (size_t)n);
    free(Sigma); free(B); free(V); free(mu); free(cav_tau); free(cav_nu);
    free(tau_old); free(nu_old); free(s); free(st);
    return info;
}

/* ---- gp/classification/GpClassifier.scala:24-47 classify ---------------------------------- */
/* K n x n, Ks m x n (test-train), Kss m x m; tau, nu, L from estimateSiteParams. Output p (m);
 * optional fmean (m) and fvar_diag (m). */
void orc_ep_classify(const double *K, int n, const double *Ks, int m, const double *Kss,
                     const double *tau, const double *nu, const double *L, double *prob,
                     double *fmean_out, double *fvar_diag_out)
{
    double *st = (double *)malloc(sizeof(double) * (size_t)n);
    double *rhs = (double *)calloc((size_t)n, sizeof(double));
    double *t1 = (double *)malloc(sizeof(double) * (size_t)n);
    double *z = (double *)malloc(sizeof(double) * (size_t)n);
    double *rhs1 = (double *)malloc(sizeof(double) * (size_t)n * m);
    double *V = (double *)malloc(sizeof(double) * (size_t)n * m);
    for (int i = 0; i < n; ++i) st[i] = sqrt(tau[i]);
    for (int a = 0; a < n; ++a) {                                       /* :33-34 */
        double acc = 0.0;
        for (int b = 0; b < n; ++b) acc += (K[a + (long)b * n] * st[a]) * nu[b];
        rhs[a] = acc;
    }
    orc_forward_solve_vec(L, n, n, 0, rhs, t1);                         /* :35 */
    orc_back_solve_vec(L, n, n, 1, t1, z);                              /* :36 */
    for (int a = 0; a < n; ++a) z[a] = st[a] * z[a];
    for (int c = 0; c < m; ++c) {                                       /* :37 */
        double acc = 0.0;
        for (int b = 0; b < n; ++b) acc += Ks[c + (long)b * m] * (nu[b] - z[b]);
        if (fmean_out) fmean_out[c] = acc;
        prob[c] = acc; /* finished below */
    }
    for (int c = 0; c < m; ++c)                                         /* :39 */
        for (int a = 0; a < n; ++a) rhs1[a + (long)c * n] = Ks[c + (long)a * m] * st[a];
    orc_forward_solve_mat(L, n, n, 0, rhs1, m, n, V, n);                /* :40 */
    for (int c = 0; c < m; ++c) {                                       /* :41-44 (diagonal only is read) */
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += V[k + (long)c * n] * V[k + (long)c * n];
        double fvar = Kss[c + (long)c * m] - acc;
        if (fvar_diag_out) fvar_diag_out[c] = fvar;
        prob[c] = orc_pnorm(prob[c] / sqrt(1 + fvar));
    }
    free(st); free(rhs); free(t1); free(z); free(rhs1); free(V);
}

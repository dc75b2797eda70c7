// gpk_api.cu -- the C ABI (include/gpk.h): handle lifetime, host<->device staging and the fused
// GpPredictor pipelines.  Everything numeric runs in the CUDA kernels of the sibling files; there is no
// CPU fallback -- a missing device is an error.
#include "gpk_internal.cuh"

#include <stdarg.h>
#include <stdlib.h>

#include <new>

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
int gpk_set_error(gpk_handle h, int status, const char* fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return status;
}

void* gpk_arena(gpk_handle h, int which, size_t bytes) {
    if (bytes <= h->arena_bytes[which] && h->arena[which]) return h->arena[which];
    if (h->arena[which]) {
        cudaStreamSynchronize(h->stream);
        cudaFree(h->arena[which]);
        h->arena[which] = nullptr;
        h->arena_bytes[which] = 0;
    }
    bytes = (bytes + 255) & ~(size_t)255;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        gpk_set_error(h, GPK_ENOMEM, "cudaMalloc of %zu bytes failed (arena %d)", bytes, which);
        return nullptr;
    }
    h->arena[which] = p;
    h->arena_bytes[which] = bytes;
    return p;
}

#define ARENA_OR_FAIL(ptr, type, h, which, bytes)           \
    type ptr = (type)gpk_arena((h), (which), (bytes));      \
    if (!ptr) return GPK_ENOMEM

extern "C" {

const char* gpk_version(void) { return "gpk 0.1 (sm_100a, FP64 DMMA)"; }

int gpk_create(gpk_handle* out, int device, void* stream) {
    if (!out) return GPK_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return GPK_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return GPK_ECUDA;
    gpk_handle h = new (std::nothrow) gpk_handle_s();
    if (!h) return GPK_ENOMEM;
    memset(h, 0, sizeof(*h));
    h->device = device;
    if (stream) {
        h->stream = (cudaStream_t)stream;
        h->own_stream = false;
    } else {
        // own main stream at the highest priority: its (critical-path) CTAs are scheduled before the side streams'
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest) != cudaSuccess) { delete h; return GPK_ECUDA; }
        h->own_stream = true;
    }
    if (cudaMallocHost((void**)&h->h_pinned, 4096 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&h->d_info, 4 * sizeof(int)) != cudaSuccess) {
        delete h;
        return GPK_ENOMEM;
    }
    cudaMemset(h->d_info, 0, 4 * sizeof(int));
    for (int i = 0; i < GPK_NSIDE; ++i)
        if (cudaStreamCreateWithPriority(&h->side[i], cudaStreamNonBlocking, 0 /* lowest */) != cudaSuccess) { delete h; return GPK_ECUDA; }
    for (int i = 0; i < GPK_NEVENTS; ++i)
        if (cudaEventCreateWithFlags(&h->evpool[i], cudaEventDisableTiming) != cudaSuccess) { delete h; return GPK_ECUDA; }
    *out = h;
    return GPK_OK;
}

int gpk_destroy(gpk_handle h) {
    if (!h) return GPK_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 8; ++i)
        if (h->arena[i]) cudaFree(h->arena[i]);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->d_info) cudaFree(h->d_info);
    for (int i = 0; i < GPK_NSIDE; ++i)
        if (h->side[i]) { cudaStreamSynchronize(h->side[i]); cudaStreamDestroy(h->side[i]); }
    for (int i = 0; i < GPK_NEVENTS; ++i)
        if (h->evpool[i]) cudaEventDestroy(h->evpool[i]);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return GPK_OK;
}

const char* gpk_last_error(gpk_handle h) { return h ? h->err : "null handle"; }
int gpk_last_info(gpk_handle h) { return h ? h->last_info : 0; }
int64_t gpk_launch_count(gpk_handle h) { return h ? h->launches : 0; }

int gpk_synchronize(gpk_handle h) {
    GPK_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPK_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
namespace {

int upload_matrix(gpk_handle h, double* dst, const double* src, int rows, int cols, int64_t ld) {
    if (rows <= 0 || cols <= 0) return GPK_OK;
    GPK_CUDA(h, cudaMemcpy2DAsync(dst, (size_t)rows * sizeof(double), src, (size_t)ld * sizeof(double),
                                  (size_t)rows * sizeof(double), (size_t)cols, cudaMemcpyHostToDevice, h->stream));
    return GPK_OK;
}
int download_matrix(gpk_handle h, double* dst, int64_t ld, const double* src, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return GPK_OK;
    GPK_CUDA(h, cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), src, (size_t)rows * sizeof(double),
                                  (size_t)rows * sizeof(double), (size_t)cols, cudaMemcpyDeviceToHost, h->stream));
    return GPK_OK;
}

// after a factorisation: fetch info, map to status
int finish_info(gpk_handle h) {
    int info = 0;
    GPK_CUDA(h, cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GPK_CUDA(h, cudaStreamSynchronize(h->stream));
    h->last_info = info;
    if (info != 0) return gpk_set_error(h, GPK_ENOTPD, "matrix not positive definite: leading minor %d", info);
    return GPK_OK;
}

struct Theta {
    int D;
    double sf, sn;
    const double* ls;
};
Theta unpack(const double* theta, int D) { return Theta{D, theta[0], theta[D + 1], theta + 1}; }

// device-side fit: K -> (L if keep_L, in A), Li, alpha, ll.  X, y on device.  Workspace in the handle's arenas.
struct FitBuffers {
    double *A, *Li, *T, *ypad, *z, *alpha, *scratch;
    int N;
};
int fit_buffers(gpk_handle h, int n, int D, FitBuffers* fb) {
    const int N = gpk_pad(n);
    fb->N = N;
    const size_t nn = (size_t)N * N * sizeof(double);
    fb->A = (double*)gpk_arena(h, ARENA_A, nn);
    fb->Li = (double*)gpk_arena(h, ARENA_B, nn);
    fb->T = (double*)gpk_arena(h, ARENA_T, gpk_chol_scratch_doubles(N) * sizeof(double));
    const size_t trmv_scratch = (size_t)(N / 1024 + 1) * N;
    const size_t misc = (size_t)3 * N + trmv_scratch + gpk_grad_scratch_doubles(N, D) + 128;
    double* m = (double*)gpk_arena(h, ARENA_MISC, misc * sizeof(double));
    if (!fb->A || !fb->Li || !fb->T || !m) return GPK_ENOMEM;
    fb->ypad = m;
    fb->z = m + N;
    fb->alpha = m + 2 * (size_t)N;
    fb->scratch = m + 3 * (size_t)N;
    return GPK_OK;
}

int fit_dev(gpk_handle h, const double* dX, int n, int D, int64_t ldx, const double* dy, const double* theta,
            int has_s, double s, int keep_L, const FitBuffers& fb, CovParams* cp, double* ll_dev) {
    int rc = gpk_make_cov_params(h, theta, D, has_s, s, cp);
    if (rc) return rc;
    rc = gpk_cov_sym_lower_padded(h, dX, n, ldx, *cp, fb.A, fb.N);
    if (rc) return rc;
    rc = gpk_potrf_inv(h, fb.A, fb.Li, fb.T, fb.N, keep_L, 0);
    if (rc) return rc;
    rc = gpk_pad_vector(h, fb.ypad, fb.N, dy, n);
    if (rc) return rc;
    rc = gpk_trmv_lower(h, fb.Li, fb.N, fb.ypad, fb.z, fb.scratch);   // z = L^-1 y      (GpPredictor.scala:121)
    if (rc) return rc;
    rc = gpk_trmv_lower_t(h, fb.Li, fb.N, fb.z, fb.alpha);            // alpha = L^-t z  (GpPredictor.scala:122)
    if (rc) return rc;
    return gpk_loglik(h, fb.A, fb.N, n, fb.ypad, fb.alpha, ll_dev);    // GpPredictor.scala:144-149
}

}  // namespace

struct gpk_model_s {
    int n, N, D;
    double* X;      // n x D, ld n
    double* Li;     // N x N
    double* alpha;  // N
    double theta[GPK_MAX_D + 2];
    CovParams cp;
};

extern "C" {

// ------------------------------------------------------------------------------------------------
// covariance
// ------------------------------------------------------------------------------------------------
int gpk_cov_se_ard_dev(gpk_handle h, const double* dX, int n, int D, int64_t ldx, const double* theta, double* dK, int64_t ldk) {
    if (!h || n < 0 || ldx < n || ldk < n) return gpk_set_error(h, GPK_EINVAL, "gpk_cov_se_ard: bad dimensions");
    CovParams cp;
    int rc = gpk_make_cov_params(h, theta, D, 0, 0.0, &cp);
    if (rc) return rc;
    return gpk_cov_sym_full(h, dX, n, ldx, cp, dK, ldk);
}

int gpk_cov_se_ard(gpk_handle h, const double* X, int n, int D, int64_t ldx, const double* theta, double* K, int64_t ldk) {
    if (!h || n < 0 || ldx < n || ldk < n) return gpk_set_error(h, GPK_EINVAL, "gpk_cov_se_ard: bad dimensions");
    if (n == 0) return GPK_OK;
    GPK_CUDA(h, cudaSetDevice(h->device));
    ARENA_OR_FAIL(dX, double*, h, ARENA_X, (size_t)n * D * sizeof(double));
    ARENA_OR_FAIL(dK, double*, h, ARENA_IO, (size_t)n * n * sizeof(double));
    int rc = upload_matrix(h, dX, X, n, D, ldx);
    if (rc) return rc;
    rc = gpk_cov_se_ard_dev(h, dX, n, D, n, theta, dK, n);
    if (rc) return rc;
    rc = download_matrix(h, K, ldk, dK, n, n);
    if (rc) return rc;
    return gpk_synchronize(h);
}

int gpk_cov_cross_se_ard_dev(gpk_handle h, const double* dX1, int m, int64_t ldx1, const double* dX2, int n, int64_t ldx2,
                             int D, const double* theta, double* dK, int64_t ldk) {
    if (!h || m < 0 || n < 0 || ldx1 < m || ldx2 < n || ldk < m) return gpk_set_error(h, GPK_EINVAL, "gpk_cov_cross: bad dimensions");
    CovParams cp;
    int rc = gpk_make_cov_params(h, theta, D, 0, 0.0, &cp);
    if (rc) return rc;
    return gpk_cov_cross(h, dX1, m, ldx1, dX2, n, ldx2, cp, dK, ldk, m, n);
}

int gpk_cov_cross_se_ard(gpk_handle h, const double* X1, int m, int64_t ldx1, const double* X2, int n, int64_t ldx2, int D,
                         const double* theta, double* K, int64_t ldk) {
    if (!h || m < 0 || n < 0 || ldx1 < m || ldx2 < n || ldk < m) return gpk_set_error(h, GPK_EINVAL, "gpk_cov_cross: bad dimensions");
    if (m == 0 || n == 0) return GPK_OK;
    GPK_CUDA(h, cudaSetDevice(h->device));
    ARENA_OR_FAIL(dX, double*, h, ARENA_X, (size_t)(m + n) * D * sizeof(double));
    ARENA_OR_FAIL(dK, double*, h, ARENA_IO, (size_t)m * n * sizeof(double));
    double* dX2 = dX + (size_t)m * D;
    int rc = upload_matrix(h, dX, X1, m, D, ldx1);
    if (rc) return rc;
    rc = upload_matrix(h, dX2, X2, n, D, ldx2);
    if (rc) return rc;
    rc = gpk_cov_cross_se_ard_dev(h, dX, m, m, dX2, n, n, D, theta, dK, m);
    if (rc) return rc;
    rc = download_matrix(h, K, ldk, dK, m, n);
    if (rc) return rc;
    return gpk_synchronize(h);
}

int gpk_cov_deriv_se_ard(gpk_handle h, int param_num, const double* X, int n, int D, int64_t ldx, const double* theta,
                         double* dKout, int64_t ldk) {
    if (!h || n < 0 || ldx < n || ldk < n) return gpk_set_error(h, GPK_EINVAL, "gpk_cov_deriv: bad dimensions");
    if (param_num < 1 || param_num > D + 2)
        return gpk_set_error(h, GPK_EINVAL, "scala.MatchError: hyper-parameter %d outside 1..%d", param_num, D + 2);
    if (n == 0) return GPK_OK;
    GPK_CUDA(h, cudaSetDevice(h->device));
    ARENA_OR_FAIL(dX, double*, h, ARENA_X, (size_t)n * D * sizeof(double));
    ARENA_OR_FAIL(dK, double*, h, ARENA_IO, (size_t)n * n * sizeof(double));
    CovParams cp;
    int rc = gpk_make_cov_params(h, theta, D, 0, 0.0, &cp);
    if (rc) return rc;
    rc = upload_matrix(h, dX, X, n, D, ldx);
    if (rc) return rc;
    Theta t = unpack(theta, D);
    rc = gpk_cov_deriv(h, param_num, dX, n, n, cp, t.sf, t.sn, t.ls, dK, n);
    if (rc) return rc;
    rc = download_matrix(h, dKout, ldk, dK, n, n);
    if (rc) return rc;
    return gpk_synchronize(h);
}

// ------------------------------------------------------------------------------------------------
// cholesky / triangular solves / triangular inverse on caller-supplied matrices
// ------------------------------------------------------------------------------------------------
int gpk_potrf_lower(gpk_handle h, const double* A, int n, int64_t lda, double* L, int64_t ldl, int check_symmetric) {
    if (!h || n < 0 || lda < n || ldl < n) return gpk_set_error(h, GPK_EINVAL, "gpk_potrf_lower: bad dimensions");
    if (n == 0) return GPK_OK;
    GPK_CUDA(h, cudaSetDevice(h->device));
    const int N = gpk_pad(n);
    ARENA_OR_FAIL(dIn, double*, h, ARENA_IO, (size_t)n * n * sizeof(double));
    ARENA_OR_FAIL(dA, double*, h, ARENA_A, (size_t)N * N * sizeof(double));
    ARENA_OR_FAIL(dLi, double*, h, ARENA_B, (size_t)N * N * sizeof(double));
    ARENA_OR_FAIL(dT, double*, h, ARENA_T, gpk_chol_scratch_doubles(N) * sizeof(double));
    int rc = upload_matrix(h, dIn, A, n, n, lda);
    if (rc) return rc;
    int* d_notsym = h->d_info + 1;
    GPK_CUDA(h, cudaMemsetAsync(d_notsym, 0, sizeof(int), h->stream));
    rc = gpk_load_sym_padded(h, dA, N, dIn, n, n, check_symmetric ? d_notsym : nullptr);
    if (rc) return rc;
    if (check_symmetric) {
        int ns = 0;
        GPK_CUDA(h, cudaMemcpyAsync(&ns, d_notsym, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        GPK_CUDA(h, cudaStreamSynchronize(h->stream));
        if (ns) return gpk_set_error(h, GPK_ENOTSYM, "matrix is not symmetric");
    }
    rc = gpk_potrf_inv(h, dA, dLi, dT, N, 1, 0);
    if (rc) return rc;
    rc = gpk_store_lower(h, dIn, n, dA, N, n);
    if (rc) return rc;
    rc = finish_info(h);
    if (rc) return rc;
    rc = download_matrix(h, L, ldl, dIn, n, n);
    if (rc) return rc;
    return gpk_synchronize(h);
}

int gpk_trtri(gpk_handle h, int is_upper, const double* T, int n, int64_t ldt, double* Tinv, int64_t ldi) {
    if (!h || n < 0 || ldt < n || ldi < n) return gpk_set_error(h, GPK_EINVAL, "gpk_trtri: bad dimensions");
    if (n == 0) return GPK_OK;
    GPK_CUDA(h, cudaSetDevice(h->device));
    const int N = gpk_pad(n);
    ARENA_OR_FAIL(dIn, double*, h, ARENA_IO, (size_t)n * n * sizeof(double));
    ARENA_OR_FAIL(dA, double*, h, ARENA_A, (size_t)N * N * sizeof(double));
    ARENA_OR_FAIL(dLi, double*, h, ARENA_B, (size_t)N * N * sizeof(double));
    ARENA_OR_FAIL(dT, double*, h, ARENA_T, gpk_chol_scratch_doubles(N) * sizeof(double));
    int rc = upload_matrix(h, dIn, T, n, n, ldt);
    if (rc) return rc;
    // an upper-triangular U is handled through its transpose: (U^t)^-1 = (U^-1)^t
    rc = gpk_load_tri_padded(h, dA, N, dIn, n, n, is_upper ? 1 : 0);
    if (rc) return rc;
    rc = gpk_trtri_lower(h, dA, dLi, dT, N);
    if (rc) return rc;
    rc = gpk_store_tri(h, dIn, n, dLi, N, n, is_upper ? 1 : 0);
    if (rc) return rc;
    rc = download_matrix(h, Tinv, ldi, dIn, n, n);
    if (rc) return rc;
    return gpk_synchronize(h);
}

int gpk_trsm(gpk_handle h, int upper, int transposed, const double* T, int n, int64_t ldt, const double* B, int nrhs,
             int64_t ldb, double* Xout, int64_t ldx) {
    if (!h || n < 0 || nrhs < 0 || ldt < n || ldb < n || ldx < n) return gpk_set_error(h, GPK_EINVAL, "gpk_trsm: bad dimensions");
    if (n == 0 || nrhs == 0) return GPK_OK;
    GPK_CUDA(h, cudaSetDevice(h->device));
    // The stored matrix is lower triangular iff (upper XOR transposed) == 0.
    const int stored_upper = (upper != 0) != (transposed != 0);
    const int N = gpk_pad(n), M = gpk_pad(nrhs);
    ARENA_OR_FAIL(dIn, double*, h, ARENA_IO, (size_t)n * n * sizeof(double));
    ARENA_OR_FAIL(dA, double*, h, ARENA_A, (size_t)N * N * sizeof(double));
    ARENA_OR_FAIL(dLi, double*, h, ARENA_B, (size_t)N * N * sizeof(double));
    ARENA_OR_FAIL(dT, double*, h, ARENA_T, gpk_chol_scratch_doubles(N) * sizeof(double));
    ARENA_OR_FAIL(dB, double*, h, ARENA_IO2, (size_t)2 * N * M * sizeof(double));
    double* dX = dB + (size_t)N * M;
    int rc = upload_matrix(h, dIn, T, n, n, ldt);
    if (rc) return rc;
    rc = gpk_load_tri_padded(h, dA, N, dIn, n, n, stored_upper);   // dA = lower-triangular Lw (= stored or stored^t)
    if (rc) return rc;
    rc = gpk_trtri_lower(h, dA, dLi, dT, N);                        // dLi = Lw^-1
    if (rc) return rc;
    GPK_CUDA(h, cudaMemsetAsync(dB, 0, (size_t)N * M * sizeof(double), h->stream));
    GPK_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)N * sizeof(double), B, (size_t)ldb * sizeof(double), (size_t)n * sizeof(double),
                                  (size_t)nrhs, cudaMemcpyHostToDevice, h->stream));
    // effective operand E: lower (forwardSolve) -> E = Lw, X = Lw^-1 B;  upper (backSolve) -> E = Lw^t, X = Lw^-t B
    GemmDesc g = gemm_desc();
    g.P = dB; g.ldp = N; g.p_kcontig = 1;            // P(r,k) = B(k,r)
    g.Q = dLi; g.ldq = N;
    g.D = dX; g.ldd = N; g.R = M; g.S = N; g.K = N;
    if (!upper) { g.q_kcontig = 0; g.ke_s = 1; }     // Q(s,k) = Lw^-1(s,k), zero for k > s
    else        { g.q_kcontig = 1; g.kb_s = 1; }     // Q(s,k) = Lw^-1(k,s), zero for k < s
    rc = gpk_gemm(h, g);
    if (rc) return rc;
    GPK_CUDA(h, cudaMemcpy2DAsync(Xout, (size_t)ldx * sizeof(double), dX, (size_t)N * sizeof(double), (size_t)n * sizeof(double),
                                  (size_t)nrhs, cudaMemcpyDeviceToHost, h->stream));
    return gpk_synchronize(h);
}

int gpk_syrk_lower_dev(gpk_handle h, const double* dP, int64_t ldp, double* dC, int64_t ldc, int n, int k) {
    if (!h || n <= 0 || k <= 0) return gpk_set_error(h, GPK_EINVAL, "gpk_syrk_lower_dev: bad dimensions");
    GemmDesc g = gemm_desc();
    g.P = dP; g.ldp = ldp; g.Q = dP; g.ldq = ldp;
    g.D = dC; g.ldd = ldc; g.Cin = dC; g.ldc = ldc;
    g.R = n; g.S = n; g.K = k; g.alpha = -1.0; g.beta = 1.0; g.tri_out = 1;
    return gpk_gemm(h, g);
}

// ------------------------------------------------------------------------------------------------
// fused GpPredictor paths
// ------------------------------------------------------------------------------------------------
int gpk_gp_nll_grad_dev(gpk_handle h, const double* dX, int n, int D, int64_t ldx, const double* dy, const double* theta,
                        int has_s, double s, int nparams, double* out_dev, int* info_dev) {
    if (!h || n <= 0 || ldx < n) return gpk_set_error(h, GPK_EINVAL, "gpk_gp_nll_grad: bad dimensions");
    if (nparams < 0 || nparams > D + 2) return gpk_set_error(h, GPK_EINVAL, "nparams=%d outside 0..%d", nparams, D + 2);
    FitBuffers fb;
    int rc = fit_buffers(h, n, D, &fb);
    if (rc) return rc;
    CovParams cp;
    rc = fit_dev(h, dX, n, D, ldx, dy, theta, has_s, s, /*keep_L=*/0, fb, &cp, out_dev);
    if (rc) return rc;
    if (nparams > 0) {
        rc = gpk_lauum_lower(h, fb.Li, fb.A, fb.N);  // K^-1 = L^-t L^-1 (GpPredictor.scala:67), lower tiles, into A
        if (rc) return rc;
        Theta t = unpack(theta, D);
        rc = gpk_grad_trace(h, fb.A, fb.N, dX, n, ldx, fb.alpha, cp, t.sf, t.sn, t.ls, nparams, out_dev + 1, fb.scratch);
        if (rc) return rc;
    }
    if (info_dev) GPK_CUDA(h, cudaMemcpyAsync(info_dev, h->d_info, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    return GPK_OK;
}

int gpk_gp_nll_grad(gpk_handle h, const double* X, int n, int D, int64_t ldx, const double* y, const double* theta,
                    int has_s, double s, int nparams, double* ll, double* grad) {
    if (!h || n <= 0 || ldx < n) return gpk_set_error(h, GPK_EINVAL, "gpk_gp_nll_grad: bad dimensions (require rows == targets)");
    GPK_CUDA(h, cudaSetDevice(h->device));
    ARENA_OR_FAIL(dX, double*, h, ARENA_X, ((size_t)n * D + n + 256) * sizeof(double));
    double* dy = dX + (size_t)n * D;
    double* dout = dy + n;
    int rc = upload_matrix(h, dX, X, n, D, ldx);
    if (rc) return rc;
    GPK_CUDA(h, cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = gpk_gp_nll_grad_dev(h, dX, n, D, n, dy, theta, has_s, s, nparams, dout, nullptr);
    if (rc) return rc;
    GPK_CUDA(h, cudaMemcpyAsync(h->h_pinned, dout, (size_t)(nparams + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    rc = finish_info(h);
    if (rc) return rc;
    *ll = h->h_pinned[0];
    for (int p = 0; p < nparams; ++p) grad[p] = h->h_pinned[1 + p];
    return GPK_OK;
}

int gpk_gp_fit(gpk_handle h, const double* X, int n, int D, int64_t ldx, const double* y, const double* theta, int has_s,
               double s, double* L, int64_t ldl, double* alpha, double* ll) {
    if (!h || n <= 0 || ldx < n || (L && ldl < n)) return gpk_set_error(h, GPK_EINVAL, "gpk_gp_fit: bad dimensions");
    GPK_CUDA(h, cudaSetDevice(h->device));
    ARENA_OR_FAIL(dX, double*, h, ARENA_X, ((size_t)n * D + n + 256) * sizeof(double));
    double* dy = dX + (size_t)n * D;
    double* dout = dy + n;
    int rc = upload_matrix(h, dX, X, n, D, ldx);
    if (rc) return rc;
    GPK_CUDA(h, cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FitBuffers fb;
    rc = fit_buffers(h, n, D, &fb);
    if (rc) return rc;
    CovParams cp;
    rc = fit_dev(h, dX, n, D, n, dy, theta, has_s, s, L != nullptr, fb, &cp, dout);
    if (rc) return rc;
    GPK_CUDA(h, cudaMemcpyAsync(h->h_pinned, dout, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    rc = finish_info(h);
    if (rc) return rc;
    if (ll) *ll = h->h_pinned[0];
    if (alpha) GPK_CUDA(h, cudaMemcpyAsync(alpha, fb.alpha, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (L) {
        ARENA_OR_FAIL(dOut, double*, h, ARENA_IO, (size_t)n * n * sizeof(double));
        rc = gpk_store_lower(h, dOut, n, fb.A, fb.N, n);
        if (rc) return rc;
        rc = download_matrix(h, L, ldl, dOut, n, n);
        if (rc) return rc;
    }
    return gpk_synchronize(h);
}

// ------------------------------------------------------------------------------------------------
// resident model + prediction
// ------------------------------------------------------------------------------------------------
static int model_alloc(gpk_handle h, int n, int D, const double* theta, gpk_model* out) {
    gpk_model m = new (std::nothrow) gpk_model_s();
    if (!m) return GPK_ENOMEM;
    memset(m, 0, sizeof(*m));
    m->n = n; m->N = gpk_pad(n); m->D = D;
    memcpy(m->theta, theta, sizeof(double) * (D + 2));
    if (cudaMalloc((void**)&m->X, (size_t)n * D * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&m->Li, (size_t)m->N * m->N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&m->alpha, (size_t)m->N * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        if (m->X) cudaFree(m->X);
        if (m->Li) cudaFree(m->Li);
        if (m->alpha) cudaFree(m->alpha);
        delete m;
        return gpk_set_error(h, GPK_ENOMEM, "model allocation failed (n=%d)", n);
    }
    *out = m;
    return GPK_OK;
}

int gpk_gp_model_destroy(gpk_handle h, gpk_model m) {
    if (!m) return GPK_OK;
    if (h) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); }
    cudaFree(m->X); cudaFree(m->Li); cudaFree(m->alpha);
    delete m;
    return GPK_OK;
}

int gpk_gp_model_fit(gpk_handle h, const double* X, int n, int D, int64_t ldx, const double* y, const double* theta,
                     int has_s, double s, gpk_model* out, double* ll) {
    if (!h || !out || n <= 0 || ldx < n) return gpk_set_error(h, GPK_EINVAL, "gpk_gp_model_fit: bad dimensions");
    if (D < 1 || D > GPK_MAX_D) return gpk_set_error(h, GPK_EINVAL, "feature dimension D=%d outside 1..%d", D, GPK_MAX_D);
    GPK_CUDA(h, cudaSetDevice(h->device));
    gpk_model m = nullptr;
    int rc = model_alloc(h, n, D, theta, &m);
    if (rc) return rc;
    ARENA_OR_FAIL(dy, double*, h, ARENA_X, ((size_t)n + 256) * sizeof(double));
    double* dout = dy + n;
    rc = upload_matrix(h, m->X, X, n, D, ldx);
    if (!rc) rc = (cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream) == cudaSuccess) ? 0 : GPK_ECUDA;
    FitBuffers fb;
    if (!rc) rc = fit_buffers(h, n, D, &fb);
    if (!rc) {
        fb.Li = m->Li;      // factor straight into the model's resident buffers
        fb.alpha = m->alpha;
        rc = fit_dev(h, m->X, n, D, n, dy, theta, has_s, s, 0, fb, &m->cp, dout);
    }
    if (!rc) rc = (cudaMemcpyAsync(h->h_pinned, dout, sizeof(double), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess) ? 0 : GPK_ECUDA;
    if (!rc) rc = finish_info(h);
    if (rc) { gpk_gp_model_destroy(h, m); return rc; }
    // predictions never include the Option sigmaNoise in the kernel (GpPredictor.scala:31,36 use newKernelFunc only)
    m->cp.extra_diag = 0.0;
    if (ll) *ll = h->h_pinned[0];
    *out = m;
    return GPK_OK;
}

int gpk_gp_model_from_factor(gpk_handle h, const double* X, int n, int D, int64_t ldx, const double* L, int64_t ldl,
                             const double* alpha, const double* theta, gpk_model* out) {
    if (!h || !out || n <= 0 || ldx < n || ldl < n) return gpk_set_error(h, GPK_EINVAL, "gpk_gp_model_from_factor: bad dimensions");
    GPK_CUDA(h, cudaSetDevice(h->device));
    gpk_model m = nullptr;
    int rc = model_alloc(h, n, D, theta, &m);
    if (rc) return rc;
    const int N = m->N;
    rc = gpk_make_cov_params(h, theta, D, 0, 0.0, &m->cp);
    double* dIn = (double*)gpk_arena(h, ARENA_IO, (size_t)n * n * sizeof(double));
    double* dA = (double*)gpk_arena(h, ARENA_A, (size_t)N * N * sizeof(double));
    double* dT = (double*)gpk_arena(h, ARENA_T, gpk_chol_scratch_doubles(N) * sizeof(double));
    if (!dIn || !dA || !dT) rc = GPK_ENOMEM;
    if (!rc) rc = upload_matrix(h, m->X, X, n, D, ldx);
    if (!rc) rc = upload_matrix(h, dIn, L, n, n, ldl);
    if (!rc) rc = gpk_load_tri_padded(h, dA, N, dIn, n, n, 0);
    if (!rc) rc = gpk_trtri_lower(h, dA, m->Li, dT, N);
    if (!rc) rc = (cudaMemsetAsync(m->alpha, 0, (size_t)N * sizeof(double), h->stream) == cudaSuccess) ? 0 : GPK_ECUDA;
    if (!rc) rc = (cudaMemcpyAsync(m->alpha, alpha, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream) == cudaSuccess) ? 0 : GPK_ECUDA;
    if (!rc) rc = gpk_synchronize(h);
    if (rc) { gpk_gp_model_destroy(h, m); return rc; }
    *out = m;
    return GPK_OK;
}

int gpk_gp_model_predict(gpk_handle h, gpk_model m, const double* Xs, int ms, int64_t ldxs, int want_full_cov, double* mean,
                         double* sigma, int64_t lds, double* V, int64_t ldv) {
    if (!h || !m || ms <= 0 || ldxs < ms || (V && ldv < m->n) || (want_full_cov && sigma && lds < ms))
        return gpk_set_error(h, GPK_EINVAL, "gpk_gp_model_predict: bad dimensions");
    GPK_CUDA(h, cudaSetDevice(h->device));
    const int n = m->n, N = m->N, D = m->D, M = gpk_pad(ms);
    ARENA_OR_FAIL(dXs, double*, h, ARENA_X, (size_t)ms * D * sizeof(double));
    // KsT (N x M), V (N x M), Sig (M x M), mean (M), var (M)
    ARENA_OR_FAIL(buf, double*, h, ARENA_IO2, ((size_t)2 * N * M + (size_t)M * M + 2 * M) * sizeof(double));
    double* dKsT = buf;
    double* dV = dKsT + (size_t)N * M;
    double* dSig = dV + (size_t)N * M;
    double* dMean = dSig + (size_t)M * M;
    double* dVar = dMean + M;
    int rc = upload_matrix(h, dXs, Xs, ms, D, ldxs);
    if (rc) return rc;
    // K*^t = k(X, X*) : N x M with zero padding (GpPredictor.scala:53, no noise)
    rc = gpk_cov_cross(h, m->X, n, n, dXs, ms, ms, m->cp, dKsT, N, N, M);
    if (rc) return rc;
    // mean = K* alpha (GpPredictor.scala:54)
    rc = gpk_colwise_dot(h, dKsT, N, N, ms, m->alpha, dMean, 0);
    if (rc) return rc;
    // V = L^-1 K*^t (GpPredictor.scala:55): C(i,c) = sum_{k<=i} Li(i,k) KsT(k,c)
    GemmDesc g = gemm_desc();
    g.P = dKsT; g.ldp = N; g.p_kcontig = 1;
    g.Q = m->Li; g.ldq = N; g.q_kcontig = 0;
    g.D = dV; g.ldd = N; g.R = M; g.S = N; g.K = N; g.ke_s = 1; g.heavy_last = 1;
    rc = gpk_gemm(h, g);
    if (rc) return rc;
    if (mean) GPK_CUDA(h, cudaMemcpyAsync(mean, dMean, (size_t)ms * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (sigma) {
        if (want_full_cov) {
            // sigma = k(X*,X*) (+ sn^2 on the diagonal, MatrixUtils.scala:63) - V^t V   (GpPredictor.scala:56)
            GPK_CUDA(h, cudaMemsetAsync(dSig, 0, (size_t)M * M * sizeof(double), h->stream));
            rc = gpk_cov_sym_full(h, dXs, ms, ms, m->cp, dSig, M);
            if (rc) return rc;
            g = gemm_desc();
            g.P = dV; g.ldp = N; g.p_kcontig = 1;
            g.Q = dV; g.ldq = N; g.q_kcontig = 1;
            g.D = dSig; g.ldd = M; g.Cin = dSig; g.ldc = M; g.R = M; g.S = M; g.K = N; g.alpha = -1.0; g.beta = 1.0;
            rc = gpk_gemm(h, g);
            if (rc) return rc;
            GPK_CUDA(h, cudaMemcpy2DAsync(sigma, (size_t)lds * sizeof(double), dSig, (size_t)M * sizeof(double),
                                          (size_t)ms * sizeof(double), (size_t)ms, cudaMemcpyDeviceToHost, h->stream));
        } else {
            rc = gpk_colwise_dot(h, dV, N, N, ms, nullptr, dVar, 1);
            if (rc) return rc;
            GPK_CUDA(h, cudaMemcpyAsync(sigma, dVar, (size_t)ms * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
    }
    if (V) GPK_CUDA(h, cudaMemcpy2DAsync(V, (size_t)ldv * sizeof(double), dV, (size_t)N * sizeof(double), (size_t)n * sizeof(double),
                                         (size_t)ms, cudaMemcpyDeviceToHost, h->stream));
    rc = gpk_synchronize(h);
    if (rc) return rc;
    if (sigma && !want_full_cov) {
        const double kss = m->cp.sf2 + m->cp.sn2;  // k(x*,x*) incl. the i==j noise term
        for (int i = 0; i < ms; ++i) sigma[i] = kss - sigma[i];
    }
    return GPK_OK;
}

int gpk_gp_predict(gpk_handle h, const double* X, int n, int D, int64_t ldx, const double* y, const double* Xs, int ms,
                   int64_t ldxs, const double* theta, int has_s, double s, double* mean, double* sigma, int64_t lds, double* ll) {
    gpk_model m = nullptr;
    int rc = gpk_gp_model_fit(h, X, n, D, ldx, y, theta, has_s, s, &m, ll);
    if (rc) return rc;
    rc = gpk_gp_model_predict(h, m, Xs, ms, ldxs, 1, mean, sigma, lds, nullptr, 0);
    gpk_gp_model_destroy(h, m);
    if (rc) return rc;
    if (has_s && sigma)  // GpPredictor.scala:37-39: + sigmaNoise * I
        for (int i = 0; i < ms; ++i) sigma[i + (int64_t)i * lds] += s;
    return GPK_OK;
}

}  // extern "C"

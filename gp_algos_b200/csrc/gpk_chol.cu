// gpk_chol.cu -- FP64 Cholesky factor AND its triangular inverse, built together.
// Replaces breeze `cholesky` -> LAPACK dpotrf('L') (GpPredictor.scala:120, EpParameterEstimator.scala:58),
// utils/MatrixUtils.scala:106-113 invTriangular (n dense forward solves, n^3/2 scalar MACs in the
// reference) and the `lInversed.t * lInversed` dgemm (GpPredictor.scala:67).
//
// Algorithm (recursive, everything O(n^3) is a DMMA GEMM from gpk_gemm.cu):
//   potrf_inv(A, Li, n):                      A = [A11 . ; A21 A22],  Li = [Li11 . ; Li21 Li22]
//     potrf_inv(A11, Li11)                    L11, L11^-1
//     L21 = A21 * Li11^t                      (TRSM as a GEMM with the already-known inverse; k <= r0+127)
//     A22 -= L21 * L21^t                      (SYRK, lower tiles only)          <- the trailing update
//     potrf_inv(A22, Li22)
//     Li21 = -Li22 * (L21 * Li11)             (two triangular-aware GEMMs)
//   base case n = 128: one CTA factors the block in shared memory and inverts it in place (gpk_base.cu).
// Total n^3/3 (factor) + n^3/3 (inverse) flops; the K^-1 = Li^t Li product (another n^3/3, lower only)
// is a single GEMM launch (gpk_lauum_lower).  Because L^-1 is a by-product, every triangular solve of
// the path (alpha, V = L^-1 K*^t) becomes a matrix product as well.
#include "gpk_internal.cuh"

#include <stdlib.h>

namespace {

constexpr int NB = GPK_TILE;  // 128

int launch_base(gpk_handle h, double* A, int64_t lda, double* Li, int64_t ldi, int col_offset, int mode, int batch,
                int64_t strideA, int64_t strideLi) {
    return gpk_base_potrf_trtri(h, A, lda, Li, ldi, col_offset, mode, batch, strideA, strideLi);
}

struct StreamSwap {  // run the enclosed launches on another stream of the same handle
    gpk_handle h; cudaStream_t saved;
    StreamSwap(gpk_handle h_, cudaStream_t s) : h(h_), saved(h_->stream) { h->stream = s; }
    ~StreamSwap() { h->stream = saved; }
};

int side_min() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GPK_SIDE_MIN"); v = e ? atoi(e) : 256; }
    return v;
}

// T (n2 x n1, ld n2) = L21 * Li11 :  C(m,c) = sum_{k>=c} L21(m,k) Li11(k,c)
int gemm_T(gpk_handle h, const double* L21, int64_t ldl, const double* Li11, int64_t ldi, double* T, int n1, int n2) {
    GemmDesc g = gemm_desc();
    g.P = Li11; g.ldp = ldi; g.p_kcontig = 1;   // P(r,k) = Li11(k,r)
    g.Q = L21; g.ldq = ldl; g.q_kcontig = 0;    // Q(s,k) = L21(s,k)
    g.D = T; g.ldd = n2;
    g.R = n1; g.S = n2; g.K = n1; g.kb_r = 1;
    return gpk_gemm(h, g);
}
// Li21 (n2 x n1) = -Li22 * T :  C(m,c) = -sum_{k<=m} Li22(m,k) T(k,c)
int gemm_Li21(gpk_handle h, const double* T, const double* Li22, double* Li21, int64_t ldi, int n1, int n2) {
    GemmDesc g = gemm_desc();
    g.P = T; g.ldp = n2; g.p_kcontig = 1;       // P(r,k) = T(k,r)
    g.Q = Li22; g.ldq = ldi; g.q_kcontig = 0;   // Q(s,k) = Li22(s,k)
    g.D = Li21; g.ldd = ldi;
    g.R = n1; g.S = n2; g.K = n2; g.ke_s = 1; g.alpha = -1.0; g.heavy_last = 1;
    return gpk_gemm(h, g);
}

// Schedule: everything on the handle's stream except T = L21 * Li11, which is not needed until the very end of the
// node.  For nodes >= side_min() it is forked to a side stream (one per recursion depth) right after L21 exists, so it
// fills the SMs that the latency-bound sub-tree of A22 (small GEMMs, 128-block spine) leaves idle, and is joined before
// Li21 = -Li22 * T.  Each depth has its own T region because a node's T is live across its whole second sub-tree.
int potrf_inv_rec(gpk_handle h, double* A, int64_t lda, double* Li, int64_t ldi, double* T, int n, int keep_L,
                  int col_offset, int depth) {
    if (n == NB) return launch_base(h, A, lda, Li, ldi, col_offset, 0, 1, 0, 0);
    const int n1 = (n / NB / 2) * NB, n2 = n - n1;
    double* A21 = A + n1;
    double* A22 = A + n1 + (int64_t)n1 * lda;
    double* Li21 = Li + n1;
    double* Li22 = Li + n1 + (int64_t)n1 * ldi;
    double* Tchild = T + (size_t)n1 * n2;
    int rc = potrf_inv_rec(h, A, lda, Li, ldi, Tchild, n1, keep_L, col_offset, depth + 1);
    if (rc) return rc;
    // L21 = A21 * Li11^t, staged in Li21's (still free) slot: C(m,c) = sum_{k<=c} A21(m,k) Li11(c,k)
    GemmDesc g = gemm_desc();
    g.P = Li; g.ldp = ldi; g.p_kcontig = 0;     // P(r,k) = Li11(r,k)
    g.Q = A21; g.ldq = lda; g.q_kcontig = 0;    // Q(s,k) = A21(s,k)
    g.D = Li21; g.ldd = ldi;
    g.R = n1; g.S = n2; g.K = n1; g.ke_r = 1; g.heavy_last = 1;
    rc = gpk_gemm(h, g);
    if (rc) return rc;
    const bool fork = n >= side_min();
    cudaEvent_t ev_done = nullptr;
    if (fork) {
        cudaStream_t side = h->side[depth % GPK_NSIDE];
        cudaEvent_t ev_l21 = h->evpool[h->ev_next++ % GPK_NEVENTS];
        ev_done = h->evpool[h->ev_next++ % GPK_NEVENTS];
        GPK_CUDA(h, cudaEventRecord(ev_l21, h->stream));
        GPK_CUDA(h, cudaStreamWaitEvent(side, ev_l21, 0));
        {
            StreamSwap sw(h, side);
            rc = gemm_T(h, Li21, ldi, Li, ldi, T, n1, n2);
        }
        if (rc) return rc;
        GPK_CUDA(h, cudaEventRecord(ev_done, side));
    }
    // A22 -= L21 * L21^t (lower tiles)
    g = gemm_desc();
    g.P = Li21; g.ldp = ldi; g.Q = Li21; g.ldq = ldi;
    g.D = A22; g.ldd = lda; g.Cin = A22; g.ldc = lda;
    g.R = n2; g.S = n2; g.K = n1; g.alpha = -1.0; g.beta = 1.0; g.tri_out = 1;
    rc = gpk_gemm(h, g);
    if (rc) return rc;
    if (keep_L) {
        rc = gpk_copy2d(h, A21, lda, Li21, ldi, n2, n1);
        if (rc) return rc;
    }
    rc = potrf_inv_rec(h, A22, lda, Li22, ldi, Tchild, n2, keep_L, col_offset + n1, depth + 1);
    if (rc) return rc;
    if (fork) {
        GPK_CUDA(h, cudaStreamWaitEvent(h->stream, ev_done, 0));
    } else {
        rc = gemm_T(h, Li21, ldi, Li, ldi, T, n1, n2);  // L21 is still staged in Li21's slot
        if (rc) return rc;
    }
    // overwrites the staging slot; T has fully consumed it (stream order / ev_done)
    return gemm_Li21(h, T, Li22, Li21, ldi, n1, n2);
}

// Li21 = -Li22 * (L21 * Li11)   for the split [n1 | n2] of an (n1+n2) block
int inverse_offdiag(gpk_handle h, const double* L21, int64_t ldl, double* Li, int64_t ldi, double* T, int n1, int n2) {
    int rc = gemm_T(h, L21, ldl, Li, ldi, T, n1, n2);
    if (rc) return rc;
    return gemm_Li21(h, T, Li + n1 + (int64_t)n1 * ldi, Li + n1, ldi, n1, n2);
}

int trtri_rec(gpk_handle h, const double* L, int64_t ldl, double* Li, int64_t ldi, double* T, int n) {
    const int n1 = (n / NB / 2) * NB, n2 = n - n1;
    if (n == NB) return GPK_OK;  // diagonal blocks were inverted by one batched launch up front
    int rc = trtri_rec(h, L, ldl, Li, ldi, T, n1);
    if (rc) return rc;
    rc = trtri_rec(h, L + n1 + (int64_t)n1 * ldl, ldl, Li + n1 + (int64_t)n1 * ldi, ldi, T, n2);
    if (rc) return rc;
    return inverse_offdiag(h, L + n1, ldl, Li, ldi, T, n1, n2);
}

}  // namespace

size_t gpk_chol_scratch_doubles(int N) {
    // sum over recursion depths of n1*n2 <= (N/2+64)^2 * (1 + 1/4 + 1/16 + ...) plus slack for uneven splits
    const size_t half = (size_t)(N / 2 + NB);
    return half * half * 3 / 2 + (size_t)N * NB;
}

int gpk_potrf_inv(gpk_handle h, double* A, double* Li, double* T, int N, int keep_L, int col_offset) {
    GPK_CUDA(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
    return potrf_inv_rec(h, A, N, Li, N, T, N, keep_L, col_offset, 0);
}

int gpk_trtri_lower(gpk_handle h, const double* L, double* Li, double* T, int N) {
    // const_cast: mode 1 never writes A
    int rc = launch_base(h, const_cast<double*>(L), N, Li, N, 0, 1, N / NB, (int64_t)NB * (N + 1), (int64_t)NB * (N + 1));
    if (rc) return rc;
    return trtri_rec(h, L, N, Li, N, T, N);
}

int gpk_lauum_lower(gpk_handle h, const double* Li, double* Kinv, int N) {
    // Kinv(m,c) = sum_{k >= max(m,c)} Li(k,m) Li(k,c), lower tiles (m >= c)
    GemmDesc g = gemm_desc();
    g.P = Li; g.ldp = N; g.p_kcontig = 1;  // P(r,k) = Li(k,r)
    g.Q = Li; g.ldq = N; g.q_kcontig = 1;  // Q(s,k) = Li(k,s)
    g.D = Kinv; g.ldd = N;
    g.R = N; g.S = N; g.K = N; g.kb_r = 1; g.kb_s = 1; g.tri_out = 1;
    return gpk_gemm(h, g);
}

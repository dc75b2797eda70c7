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
//
// Batched mode (`batch` independent problems of the same N, matrices N*N apart): every launch of the recursion
// covers all problems through the GEMM's / base kernel's batch dimension, so 512 problems of n = 1024 cost the
// same 45 launches as one.
#include "gpk_internal.cuh"

#include <stdlib.h>

namespace {

constexpr int NB = GPK_TILE;  // 128

struct Ctx {
    gpk_handle h;
    int64_t lda, ldi;      // leading dimensions of A and Li (== N)
    int keep_L;
    int* info;             // device, one int per problem
    int batch;
    int64_t sM;            // matrix stride between problems (N*N)
    int64_t sT;            // scratch stride between problems
};

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

void set_batch(GemmDesc& g, const Ctx& c, int64_t sP, int64_t sQ, int64_t sD, int64_t sC) {
    g.batch = c.batch; g.strideP = sP; g.strideQ = sQ; g.strideD = sD; g.strideC = sC;
}

// T (n2 x n1, ld n2) = L21 * Li11 :  C(m,c) = sum_{k>=c} L21(m,k) Li11(k,c)
int gemm_T(const Ctx& c, const double* L21, int64_t ldl, int64_t sL21, const double* Li11, double* T, int n1, int n2) {
    GemmDesc g = gemm_desc();
    g.P = Li11; g.ldp = c.ldi; g.p_kcontig = 1;   // P(r,k) = Li11(k,r)
    g.Q = L21; g.ldq = ldl; g.q_kcontig = 0;      // Q(s,k) = L21(s,k)
    g.D = T; g.ldd = n2;
    g.R = n1; g.S = n2; g.K = n1; g.kb_r = 1;
    set_batch(g, c, c.sM, sL21, c.sT, 0);
    return gpk_gemm(c.h, g);
}
// Li21 (n2 x n1) = -Li22 * T :  C(m,c) = -sum_{k<=m} Li22(m,k) T(k,c)
int gemm_Li21(const Ctx& c, const double* T, const double* Li22, double* Li21, int n1, int n2) {
    GemmDesc g = gemm_desc();
    g.P = T; g.ldp = n2; g.p_kcontig = 1;         // P(r,k) = T(k,r)
    g.Q = Li22; g.ldq = c.ldi; g.q_kcontig = 0;   // Q(s,k) = Li22(s,k)
    g.D = Li21; g.ldd = c.ldi;
    g.R = n1; g.S = n2; g.K = n2; g.ke_s = 1; g.alpha = -1.0; g.heavy_last = 1;
    set_batch(g, c, c.sT, c.sM, c.sM, 0);
    return gpk_gemm(c.h, g);
}

// Schedule: everything on the handle's stream except T = L21 * Li11, which is not needed until the very end of the
// node.  For nodes >= side_min() it is forked to a (low-priority) side stream, one per recursion depth, right after
// L21 exists, and joined before Li21 = -Li22 * T.  Each depth has its own T region because a node's T is live across
// its whole second sub-tree.
int potrf_inv_rec(const Ctx& c, double* A, double* Li, double* T, int n, int col_offset, int depth) {
    gpk_handle h = c.h;
    if (n == NB)
        return gpk_base_potrf_trtri(h, A, c.lda, Li, c.ldi, c.info, col_offset, 0, c.batch, c.sM, c.sM, 1, 0);
    const int n1 = (n / NB / 2) * NB, n2 = n - n1;
    double* A21 = A + n1;
    double* A22 = A + n1 + (int64_t)n1 * c.lda;
    double* Li21 = Li + n1;
    double* Li22 = Li + n1 + (int64_t)n1 * c.ldi;
    double* Tchild = T + (size_t)n1 * n2;
    int rc = potrf_inv_rec(c, A, Li, Tchild, n1, col_offset, depth + 1);
    if (rc) return rc;
    // L21 = A21 * Li11^t, staged in Li21's (still free) slot: C(m,c) = sum_{k<=c} A21(m,k) Li11(c,k)
    GemmDesc g = gemm_desc();
    g.P = Li; g.ldp = c.ldi; g.p_kcontig = 0;     // P(r,k) = Li11(r,k)
    g.Q = A21; g.ldq = c.lda; g.q_kcontig = 0;    // Q(s,k) = A21(s,k)
    g.D = Li21; g.ldd = c.ldi;
    g.R = n1; g.S = n2; g.K = n1; g.ke_r = 1; g.heavy_last = 1;
    set_batch(g, c, c.sM, c.sM, c.sM, 0);
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
            rc = gemm_T(c, Li21, c.ldi, c.sM, Li, T, n1, n2);
        }
        if (rc) return rc;
        GPK_CUDA(h, cudaEventRecord(ev_done, side));
    }
    // A22 -= L21 * L21^t (lower tiles)
    g = gemm_desc();
    g.P = Li21; g.ldp = c.ldi; g.Q = Li21; g.ldq = c.ldi;
    g.D = A22; g.ldd = c.lda; g.Cin = A22; g.ldc = c.lda;
    g.R = n2; g.S = n2; g.K = n1; g.alpha = -1.0; g.beta = 1.0; g.tri_out = 1;
    set_batch(g, c, c.sM, c.sM, c.sM, c.sM);
    rc = gpk_gemm(h, g);
    if (rc) return rc;
    if (c.keep_L) {
        for (int b = 0; b < c.batch && !rc; ++b)
            rc = gpk_copy2d(h, A21 + b * c.sM, c.lda, Li21 + b * c.sM, c.ldi, n2, n1);
        if (rc) return rc;
    }
    rc = potrf_inv_rec(c, A22, Li22, Tchild, n2, col_offset + n1, depth + 1);
    if (rc) return rc;
    if (fork) {
        GPK_CUDA(h, cudaStreamWaitEvent(h->stream, ev_done, 0));
    } else {
        rc = gemm_T(c, Li21, c.ldi, c.sM, Li, T, n1, n2);  // L21 is still staged in Li21's slot
        if (rc) return rc;
    }
    // overwrites the staging slot; T has fully consumed it (stream order / ev_done)
    return gemm_Li21(c, T, Li22, Li21, n1, n2);
}

int trtri_rec(const Ctx& c, const double* L, double* Li, double* T, int n) {
    const int n1 = (n / NB / 2) * NB, n2 = n - n1;
    if (n == NB) return GPK_OK;  // diagonal blocks were inverted by one batched launch up front
    int rc = trtri_rec(c, L, Li, T, n1);
    if (rc) return rc;
    rc = trtri_rec(c, L + n1 + (int64_t)n1 * c.lda, Li + n1 + (int64_t)n1 * c.ldi, T, n2);
    if (rc) return rc;
    rc = gemm_T(c, L + n1, c.lda, c.sM, Li, T, n1, n2);
    if (rc) return rc;
    return gemm_Li21(c, T, Li + n1 + (int64_t)n1 * c.ldi, Li + n1, n1, n2);
}

}  // namespace

size_t gpk_chol_scratch_doubles(int N) {
    // sum over recursion depths of n1*n2 <= (N/2+64)^2 * (1 + 1/4 + 1/16 + ...) plus slack for uneven splits
    const size_t half = (size_t)(N / 2 + NB);
    return half * half * 3 / 2 + (size_t)N * NB;
}

int gpk_potrf_inv(gpk_handle h, double* A, double* Li, double* T, int N, int keep_L, int* info_dev, int batch) {
    GPK_CUDA(h, cudaMemsetAsync(info_dev, 0, sizeof(int) * (size_t)batch, h->stream));
    Ctx c{h, N, N, keep_L, info_dev, batch, (int64_t)N * N, (int64_t)gpk_chol_scratch_doubles(N)};
    return potrf_inv_rec(c, A, Li, T, N, 0, 0);
}

int gpk_trtri_lower(gpk_handle h, const double* L, double* Li, double* T, int N) {
    // const_cast: mode 1 never writes A.  One launch inverts all N/128 diagonal blocks.
    int rc = gpk_base_potrf_trtri(h, const_cast<double*>(L), N, Li, N, h->d_info, 0, 1, N / NB, (int64_t)NB * (N + 1),
                                  (int64_t)NB * (N + 1), 0, NB);
    if (rc) return rc;
    Ctx c{h, N, N, 0, h->d_info, 1, (int64_t)N * N, (int64_t)gpk_chol_scratch_doubles(N)};
    return trtri_rec(c, L, Li, T, N);
}

int gpk_lauum_lower(gpk_handle h, const double* Li, double* Kinv, int N, int batch) {
    // Kinv(m,c) = sum_{k >= max(m,c)} Li(k,m) Li(k,c), lower tiles (m >= c)
    GemmDesc g = gemm_desc();
    g.P = Li; g.ldp = N; g.p_kcontig = 1;  // P(r,k) = Li(k,r)
    g.Q = Li; g.ldq = N; g.q_kcontig = 1;  // Q(s,k) = Li(k,s)
    g.D = Kinv; g.ldd = N;
    g.R = N; g.S = N; g.K = N; g.kb_r = 1; g.kb_s = 1; g.tri_out = 1;
    g.batch = batch; g.strideP = g.strideQ = g.strideD = (int64_t)N * N;
    return gpk_gemm(h, g);
}

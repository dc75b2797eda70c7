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
//   base case n = 128: one CTA factors the block in shared memory and inverts it in place.
// Total n^3/3 (factor) + n^3/3 (inverse) flops; the K^-1 = Li^t Li product (another n^3/3, lower only)
// is a single GEMM launch (gpk_lauum_lower).  Because L^-1 is a by-product, every triangular solve of
// the path (alpha, V = L^-1 K*^t) becomes a matrix product as well.
#include "gpk_internal.cuh"

namespace {

constexpr int NB = GPK_TILE;  // 128
constexpr int SLD = NB + 1;   // shared-memory column stride (odd -> conflict-free row walks)

// One CTA of 128 threads; thread r owns row r.  mode 0: factor A (lower) -> L (written back to A,
// lower incl. diagonal) and Li = L^-1;  mode 1: A already holds a lower-triangular L, only invert.
// Li gets the full 128x128 block (zeros above the diagonal).  A non-positive pivot records
// col_offset + j + 1 into *info (first failure wins) and poisons the block with NaN.
__global__ void __launch_bounds__(NB) potf2_trti2_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ Li,
                                                         int64_t ldi, int* info, int col_offset, int mode,
                                                         int64_t strideA, int64_t strideLi) {
    extern __shared__ double S[];  // S[r + c*SLD]
    __shared__ double colbuf[NB];
    __shared__ double piv_s;
    A += blockIdx.x * strideA;
    Li += blockIdx.x * strideLi;
    col_offset += blockIdx.x * NB;
    const int r = threadIdx.x;

    for (int c = 0; c < NB; ++c) S[r + c * SLD] = (r >= c) ? A[r + (int64_t)c * lda] : 0.0;
    __syncthreads();

    if (mode == 0) {
        for (int p = 0; p < NB / 32; ++p) {
            const int j0 = 32 * p;
            double a[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) a[c] = S[r + (j0 + c) * SLD];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                const int j = j0 + jj;
                if (r == j) piv_s = a[jj];
                __syncthreads();
                const double piv = piv_s;
                if (!(piv > 0.0) && r == j) atomicCAS(info, 0, col_offset + j + 1);
                const double d = sqrt(piv);
                if (r >= j) {
                    a[jj] = (r == j) ? d : a[jj] / d;
                    colbuf[r] = a[jj];
                }
                __syncthreads();
                if (r > j) {
                    const double l = a[jj];
#pragma unroll
                    for (int cc = jj + 1; cc < 32; ++cc)
                        if (r >= j0 + cc) a[cc] -= l * colbuf[j0 + cc];
                }
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) S[r + (j0 + c) * SLD] = (r >= j0 + c) ? a[c] : 0.0;
            __syncthreads();
            // trailing update of columns >= j0+32 with this 32-wide panel
            if (r >= j0 + 32) {
                for (int c = j0 + 32; c <= r; ++c) {
                    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        acc0 += a[k] * S[c + (j0 + k) * SLD];
                        acc1 += a[k + 1] * S[c + (j0 + k + 1) * SLD];
                    }
                    S[r + c * SLD] -= (acc0 + acc1);
                }
            }
            __syncthreads();
        }
        for (int c = 0; c < NB; ++c)
            if (r >= c) A[r + (int64_t)c * lda] = S[r + c * SLD];
        __syncthreads();
    }

    // in-place inverse of the lower-triangular S (LAPACK dtrti2 'L' ordering: last column first)
    for (int j = NB - 1; j >= 0; --j) {
        const double ajj = 1.0 / S[j + j * SLD];
        double acc0 = 0.0, acc1 = 0.0;
        if (r > j) {
            int k = j + 1;
            for (; k + 1 <= r; k += 2) {
                acc0 += S[r + k * SLD] * S[k + j * SLD];
                acc1 += S[r + (k + 1) * SLD] * S[(k + 1) + j * SLD];
            }
            if (k <= r) acc0 += S[r + k * SLD] * S[k + j * SLD];
        }
        __syncthreads();
        if (r > j) S[r + j * SLD] = -(acc0 + acc1) * ajj;
        if (r == j) S[j + j * SLD] = ajj;
        __syncthreads();
    }
    for (int c = 0; c < NB; ++c) Li[r + (int64_t)c * ldi] = S[r + c * SLD];
}

int launch_base(gpk_handle h, double* A, int64_t lda, double* Li, int64_t ldi, int col_offset, int mode, int batch,
                int64_t strideA, int64_t strideLi) {
    const size_t smem = (size_t)NB * SLD * sizeof(double);
    if (!(h->func_cfg & (1u << 8))) {
        GPK_CUDA(h, cudaFuncSetAttribute(potf2_trti2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->func_cfg |= (1u << 8);
    }
    potf2_trti2_kernel<<<batch, NB, smem, h->stream>>>(A, lda, Li, ldi, h->d_info, col_offset, mode, strideA, strideLi);
    GPK_LAUNCH_CHECK(h);
    return GPK_OK;
}

// Li21 = -Li22 * (L21 * Li11)   for the split [n1 | n2] of an (n1+n2) block
int inverse_offdiag(gpk_handle h, const double* L21, int64_t ldl, double* Li, int64_t ldi, double* T, int n1, int n2) {
    const double* Li11 = Li;
    double* Li21 = Li + n1;
    const double* Li22 = Li + n1 + (int64_t)n1 * ldi;
    // T (n2 x n1, ld n2) = L21 * Li11 :  C(m,c) = sum_{k>=c} L21(m,k) Li11(k,c)
    GemmDesc g = gemm_desc();
    g.P = Li11; g.ldp = ldi; g.p_kcontig = 1;   // P(r,k) = Li11(k,r)
    g.Q = L21; g.ldq = ldl; g.q_kcontig = 0;    // Q(s,k) = L21(s,k)
    g.D = T; g.ldd = n2;
    g.R = n1; g.S = n2; g.K = n1; g.kb_r = 1;
    int rc = gpk_gemm(h, g);
    if (rc) return rc;
    // Li21 (n2 x n1) = -Li22 * T :  C(m,c) = -sum_{k<=m} Li22(m,k) T(k,c)
    g = gemm_desc();
    g.P = T; g.ldp = n2; g.p_kcontig = 1;       // P(r,k) = T(k,r)
    g.Q = Li22; g.ldq = ldi; g.q_kcontig = 0;   // Q(s,k) = Li22(s,k)
    g.D = Li21; g.ldd = ldi;
    g.R = n1; g.S = n2; g.K = n2; g.ke_s = 1; g.alpha = -1.0; g.heavy_last = 1;
    return gpk_gemm(h, g);
}

int potrf_inv_rec(gpk_handle h, double* A, int64_t lda, double* Li, int64_t ldi, double* T, int n, int keep_L,
                  int col_offset) {
    if (n == NB) return launch_base(h, A, lda, Li, ldi, col_offset, 0, 1, 0, 0);
    const int n1 = (n / NB / 2) * NB, n2 = n - n1;
    double* A21 = A + n1;
    double* A22 = A + n1 + (int64_t)n1 * lda;
    double* Li21 = Li + n1;
    double* Li22 = Li + n1 + (int64_t)n1 * ldi;
    int rc = potrf_inv_rec(h, A, lda, Li, ldi, T, n1, keep_L, col_offset);
    if (rc) return rc;
    // L21 = A21 * Li11^t, staged in Li21's (still free) slot: C(m,c) = sum_{k<=c} A21(m,k) Li11(c,k)
    GemmDesc g = gemm_desc();
    g.P = Li; g.ldp = ldi; g.p_kcontig = 0;     // P(r,k) = Li11(r,k)
    g.Q = A21; g.ldq = lda; g.q_kcontig = 0;    // Q(s,k) = A21(s,k)
    g.D = Li21; g.ldd = ldi;
    g.R = n1; g.S = n2; g.K = n1; g.ke_r = 1; g.heavy_last = 1;
    rc = gpk_gemm(h, g);
    if (rc) return rc;
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
    rc = potrf_inv_rec(h, A22, lda, Li22, ldi, T, n2, keep_L, col_offset + n1);
    if (rc) return rc;
    // the L21 operand of the inverse is read from the staging slot when A was not updated; the second GEMM
    // of inverse_offdiag overwrites that slot only after the first one has consumed it (stream order).
    return inverse_offdiag(h, keep_L ? A21 : Li21, keep_L ? lda : ldi, Li, ldi, T, n1, n2);
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
    const size_t half = (size_t)(N / 2 + NB);
    return half * half;
}

int gpk_potrf_inv(gpk_handle h, double* A, double* Li, double* T, int N, int keep_L, int col_offset) {
    GPK_CUDA(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
    return potrf_inv_rec(h, A, N, Li, N, T, N, keep_L, col_offset);
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

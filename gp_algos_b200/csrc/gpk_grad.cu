// gpk_grad.cu -- fused likelihood-gradient trace.
// Reference (GpPredictor.scala:69-78): for every hyper-parameter p it materialises dK/dtheta_p with one
// closure call per element (MatrixUtils.scala:72-84 + KernelRequisites.scala:76-86), runs a full n x n x n dgemm
// (alphaSq - inversedK) * dK_p and takes the trace: P * (2 n^3 flops + 8 n^2 bytes).
// Here: g_p = 1/2 sum_ij W_ij dk_p(x_i,x_j,[i==j]),  W = alpha alpha^t - K^-1, in ONE pass over the lower
// triangle of K^-1 (symmetry: off-diagonal terms counted twice).  dk_p is recomputed from X tiles staged in
// shared memory, so dK/dtheta is never materialised: 4 n^2 bytes read, O(n^2 (D + 30)) FP64 flops.
//   p = 1      : dk = 2 sf e                       e = exp(-r/2)
//   2..D+1     : dk = sf^2 e (x_id - x_jd)^2 / l_d^3
//   D+2        : dk = 2 sn [i == j]
// Batched over independent problems through blockIdx.y.
#include "gpk_internal.cuh"

namespace {

constexpr int GT = 64;  // tile edge
constexpr int GC = 8;   // parameters accumulated per register chunk

struct GradArgs {
    const double* Kinv; int N;
    const double* X; int64_t ldx; int n;
    const double* alpha;
    double* partial;  // [batch][numBlocks][D + 2]
    int64_t strideX;
    const ProblemParams* pp;  // device array or nullptr -> cp
    CovParams cp;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; ++i) s += red[i];
    return s;  // valid on thread 0
}

__global__ void __launch_bounds__(256) grad_trace_kernel(const GradArgs a) {
    extern __shared__ double sm[];  // xi[D][GT], xj[D][GT]
    __shared__ double red[8];
    const int64_t pb = blockIdx.y;
    const CovParams& cp = a.pp ? a.pp[pb].cp : a.cp;
    const double* Kinv = a.Kinv ? a.Kinv + pb * (int64_t)a.N * a.N : nullptr;   // nullptr: W = alpha alpha^t only
    const double* X = a.X + pb * a.strideX;
    const double* alpha = a.alpha + pb * a.N;
    const int D = cp.D;
    double* xi = sm;
    double* xj = sm + D * GT;
    // linear block id -> lower-triangular tile (bi >= bj)
    const int t = blockIdx.x;
    int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((int64_t)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while ((int64_t)bi * (bi + 1) / 2 > t) --bi;
    const int bj = t - (int)((int64_t)bi * (bi + 1) / 2);
    const int i0 = bi * GT, j0 = bj * GT;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;

    for (int e = tid; e < D * GT; e += 256) {
        const int d = e / GT, l = e % GT;
        const int gi = i0 + l, gj = j0 + l;
        xi[e] = (gi < a.n) ? X[gi + (int64_t)d * a.ldx] : 0.0;
        xj[e] = (gj < a.n) ? X[gj + (int64_t)d * a.ldx] : 0.0;
    }
    __syncthreads();

    // per-thread elements: rows gi0, gi0+1; columns j0 + ty + 8b
    const int gi0 = i0 + 2 * tx;
    double we[2][8];  // mult * W_ij * e_ij   (0 for masked elements)
    double wdiag = 0.0;
    {
        double r[2][8];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int b = 0; b < 8; ++b) r[q][b] = 0.0;
        for (int d = 0; d < D; ++d) {
            const double inv = cp.inv_ls2[d];
            const double x0 = xi[d * GT + 2 * tx], x1 = xi[d * GT + 2 * tx + 1];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const double xv = xj[d * GT + ty + 8 * b];
                const double f0 = x0 - xv, f1 = x1 - xv;
                r[0][b] += (f0 * inv) * f0;
                r[1][b] += (f1 * inv) * f1;
            }
        }
        const double al0 = (gi0 < a.n) ? alpha[gi0] : 0.0, al1 = (gi0 + 1 < a.n) ? alpha[gi0 + 1] : 0.0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int gj = j0 + ty + 8 * b;
            const double aj = (gj < a.n) ? alpha[gj] : 0.0;
            const double2 kv = Kinv ? *reinterpret_cast<const double2*>(Kinv + gi0 + (int64_t)gj * a.N) : make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gi = gi0 + q;
                const double w = (q ? al1 : al0) * aj - (q ? kv.y : kv.x);
                const bool valid = gi < a.n && gj < a.n && gi >= gj;
                const double mult = (gi == gj) ? 1.0 : 2.0;
                we[q][b] = valid ? mult * w * exp(-0.5 * r[q][b]) : 0.0;
                if (valid && gi == gj) wdiag += w;
            }
        }
    }

    double* out = a.partial + (pb * gridDim.x + blockIdx.x) * (int64_t)(D + 2);
    {   // p = 1 (signal) and p = D+2 (noise)
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int b = 0; b < 8; ++b) s += we[q][b];
        const double ts = block_sum(s, red);
        const double tn = block_sum(wdiag, red);
        if (tid == 0) { out[0] = ts; out[D + 1] = tn; }
    }
    for (int d0 = 0; d0 < D; d0 += GC) {
        double acc[GC];
#pragma unroll
        for (int c = 0; c < GC; ++c) acc[c] = 0.0;
#pragma unroll
        for (int c = 0; c < GC; ++c) {
            const int d = d0 + c;
            if (d < D) {
                const double x0 = xi[d * GT + 2 * tx], x1 = xi[d * GT + 2 * tx + 1];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const double xv = xj[d * GT + ty + 8 * b];
                    const double f0 = x0 - xv, f1 = x1 - xv;
                    acc[c] += we[0][b] * (f0 * f0) + we[1][b] * (f1 * f1);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < GC; ++c) {
            if (d0 + c < D) {
                const double tsum = block_sum(acc[c], red);
                if (tid == 0) out[1 + d0 + c] = tsum;
            }
        }
    }
}

// Co2Kernel (gp/regression/Co2Prediction.scala:66-137): same tiling and reduction, 1-D inputs, the 11 derivatives of one pair
// share their sub-expressions (co2_derivs); partial[..][p] = sum over the tile of mult * W_ij * dk_p(x_i, x_j, i == j).
__global__ void __launch_bounds__(256) grad_trace_co2_kernel(const GradArgs a) {
    __shared__ double red[8];
    __shared__ double xi[GT], xj[GT];
    const int64_t pb = blockIdx.y;
    const CovParams& cp = a.pp ? a.pp[pb].cp : a.cp;
    const double* Kinv = a.Kinv ? a.Kinv + pb * (int64_t)a.N * a.N : nullptr;
    const double* X = a.X + pb * a.strideX;
    const double* alpha = a.alpha + pb * a.N;
    const int t = blockIdx.x;
    int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((int64_t)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while ((int64_t)bi * (bi + 1) / 2 > t) --bi;
    const int bj = t - (int)((int64_t)bi * (bi + 1) / 2);
    const int i0 = bi * GT, j0 = bj * GT;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    if (tid < GT) {
        xi[tid] = (i0 + tid < a.n) ? X[i0 + tid] : 0.0;
        xj[tid] = (j0 + tid < a.n) ? X[j0 + tid] : 0.0;
    }
    __syncthreads();
    const int gi0 = i0 + 2 * tx;
    const double al0 = (gi0 < a.n) ? alpha[gi0] : 0.0, al1 = (gi0 + 1 < a.n) ? alpha[gi0 + 1] : 0.0;
    double acc[GPK_CO2_NPARAMS];
#pragma unroll
    for (int p = 0; p < GPK_CO2_NPARAMS; ++p) acc[p] = 0.0;
    for (int b = 0; b < 8; ++b) {
        const int gj = j0 + ty + 8 * b;
        const double aj = (gj < a.n) ? alpha[gj] : 0.0;
        const double2 kv = Kinv ? *reinterpret_cast<const double2*>(Kinv + gi0 + (int64_t)gj * a.N) : make_double2(0.0, 0.0);
        for (int q = 0; q < 2; ++q) {
            const int gi = gi0 + q;
            if (!(gi < a.n && gj < a.n && gi >= gj)) continue;
            const double w = ((q ? al1 : al0) * aj - (q ? kv.y : kv.x)) * ((gi == gj) ? 1.0 : 2.0);
            Co2Terms tm;
            double dk[GPK_CO2_NPARAMS];
            co2_value(cp.inv_ls2, xi[2 * tx + q] - xj[ty + 8 * b], tm);
            co2_derivs(cp.inv_ls2, tm, dk);
            dk[10] = (gi == gj) ? 2 * cp.inv_ls2[10] : 0.0;
#pragma unroll
            for (int p = 0; p < GPK_CO2_NPARAMS; ++p) acc[p] += w * dk[p];
        }
    }
    double* out = a.partial + (pb * gridDim.x + blockIdx.x) * (int64_t)GPK_CO2_NPARAMS;
#pragma unroll
    for (int p = 0; p < GPK_CO2_NPARAMS; ++p) {
        const double ts = block_sum(acc[p], red);
        if (tid == 0) out[p] = ts;
    }
}

// g[b][p] = gscale_b[p] * sum_blocks partial[b][block][p]   (fixed order -> deterministic)
struct GradScale { double s[GPK_MAX_D + 2]; };
__global__ void __launch_bounds__(256) grad_finish_kernel(const double* partial, int nblocks, int np_all, int nparams,
                                                          const GradScale scale, const ProblemParams* pp, double* g,
                                                          int64_t strideOut) {
    __shared__ double sd[256];
    const int p = blockIdx.x;
    const int64_t pb = blockIdx.y;
    if (p >= nparams) return;
    partial += pb * (int64_t)nblocks * np_all;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(int64_t)b * np_all + p];
    sd[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sd[threadIdx.x] += sd[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) g[pb * strideOut + p] = (pp ? pp[pb].gscale[p] : scale.s[p]) * sd[0];
}

}  // namespace

size_t gpk_grad_scratch_doubles(int N, int D) {
    const size_t tiles = (size_t)(N / GT) * (N / GT + 1) / 2;
    const size_t np = (size_t)(D + 2 > GPK_CO2_NPARAMS ? D + 2 : GPK_CO2_NPARAMS);   // either kernel family
    return tiles * np + np;
}

int gpk_grad_trace(gpk_handle h, const double* Kinv, int N, const double* dX, int n, int64_t ldx, const double* alpha,
                   const ProblemParams& pp, int nparams, double* g_out, double* scratch, int batch, int64_t strideX,
                   const ProblemParams* pp_dev, int64_t strideOut) {
    const int D = pp.cp.D;
    const int np_all = pp.cp.kind == GPK_KERNEL_CO2 ? GPK_CO2_NPARAMS : D + 2;
    if (nparams < 0 || nparams > np_all) return gpk_set_error(h, GPK_EINVAL, "nparams=%d outside 0..%d", nparams, np_all);
    if (nparams == 0) return GPK_OK;
    const int nt = (n + GT - 1) / GT;
    const int nblocks = nt * (nt + 1) / 2;
    GradArgs a;
    a.Kinv = Kinv; a.N = N; a.X = dX; a.ldx = ldx; a.n = n; a.alpha = alpha; a.partial = scratch; a.cp = pp.cp;
    a.strideX = strideX; a.pp = pp_dev;
    const size_t smem = (size_t)2 * D * GT * sizeof(double);
    if (smem > 48 * 1024 && !(h->func_cfg & (1u << 9))) {
        // opt in once for the LARGEST request any later call can make (D = GPK_MAX_D): a handle that sees D = 50 first and
        // D = 64 afterwards must not be left with the smaller limit
        GPK_CUDA(h, cudaFuncSetAttribute(grad_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)2 * GPK_MAX_D * GT * sizeof(double))));
        h->func_cfg |= (1u << 9);
    }
    if (pp.cp.kind == GPK_KERNEL_CO2) grad_trace_co2_kernel<<<dim3(nblocks, batch), 256, 0, h->stream>>>(a);
    else grad_trace_kernel<<<dim3(nblocks, batch), 256, smem, h->stream>>>(a);
    GPK_LAUNCH_CHECK(h);
    GradScale sc;
    memcpy(sc.s, pp.gscale, sizeof(sc.s));
    grad_finish_kernel<<<dim3(nparams, batch), 256, 0, h->stream>>>(scratch, nblocks, np_all, nparams, sc, pp_dev, g_out, strideOut);
    GPK_LAUNCH_CHECK(h);
    return GPK_OK;
}

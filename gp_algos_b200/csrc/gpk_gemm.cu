// gpk_gemm.cu -- FP64 tensor-core (DMMA.8x8x4) tile GEMM used by every O(n^3) step of the path:
// Cholesky trailing update (SYRK), L21 = A21 L11^-T, the blocked triangular inverse, K^-1 = L^-T L^-1
// (reference: GpPredictor.scala:66-67,120; EpParameterEstimator.scala:58-60) and the predictive
// V = L^-1 K*^t (GpPredictor.scala:34,55).
//
// sm_100a has no tcgen05 kind for f64 (ptxas rejects .kind::f64), so the FP64 tensor path is the
// warp-level mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4.  Measured on B200 (profiles/r01_fp64_microbench.txt):
// DMMA and DFMA both peak at 37.1 TFLOP/s; DMMA needs 4x fewer shared-memory operand bytes per flop, which is what lets
// a small CTA tile run pipe-bound.
//
// Default layout (SmallTile): 64 x 64 x 16 CTA tile, 128 threads = 4 warps as 2 (r) x 2 (s), each warp owning a 32 x 32
// sub-tile = 4 x 4 DMMA accumulators; 3 CTAs per SM (60 KB of shared memory each), which beat every larger tile shape
// (profiles/r01_tile_configs.txt).  Operands are staged global -> shared with 16-byte cp.async (LDGSTS) through a 3-stage
// ring; shared rows are padded by 4 doubles so that the 8-byte fragment reads of a half-warp (4 k x 4 rows) hit 16 distinct
// bank pairs.  The main loop carries running global pointers and cycling stage counters and issues the next refill behind
// the first DMMA group of each k-step.  Measured: 32.5 TFLOP/s on the SYRK of n = k = 4096 (91.7 % of cuBLAS Dgemm), DMMA
// pipe 95 % of active cycles (profiles/r01_ncu_summary_v3.txt).
#include "gpk_internal.cuh"

#include <stdlib.h>

namespace {

constexpr int TK = 16;
constexpr int LDK = TK + 4;  // row stride (doubles) of a [x][k] stage   (k-contiguous source)
constexpr int GROUP_S = 8;   // raster: blocks walk bands of 8 s-tiles so a wave shares operand panels in L2

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Tile configuration: WR x WS warps, each owning a (BI*8) x (BJ*8) sub-tile.
//   Big  : 2 x 4 warps, 8 x 4 blocks -> 128 x 128 CTA tile, 256 threads, 1 CTA/SM   (throughput)
//   Small: 2 x 2 warps, 4 x 4 blocks ->  64 x  64 CTA tile, 128 threads, 2 CTAs/SM  (few-tile problems: spreads a
//          128-tile over 4 SMs; one SM can only retire 0.25 TFLOP/s of FP64)
template <int WR_, int WS_, int BI_, int BJ_, int STAGES_, int MINB_>
struct TileCfg {
    static constexpr int WR = WR_, WS = WS_, BI = BI_, BJ = BJ_, STAGES = STAGES_;
    static constexpr int TR = WR * BI * 8, TS = WS * BJ * 8, NT = 32 * WR * WS;
    static constexpr int LDRP = TR + 4, LDRQ = TS + 4;  // row stride of a [k][x] stage (x-contiguous source), %16 == 4
    static constexpr int P_STAGE = (TR * LDK > TK * LDRP) ? TR * LDK : TK * LDRP;
    static constexpr int Q_STAGE = (TS * LDK > TK * LDRQ) ? TS * LDK : TK * LDRQ;
    static constexpr size_t SMEM = (size_t)STAGES * (P_STAGE + Q_STAGE) * sizeof(double);
    static constexpr int MIN_BLOCKS = MINB_;
};
using BigTile = TileCfg<2, 4, 8, 4, 4, 1>;
using WideTile = TileCfg<4, 4, 4, 4, 4, 1>;    // 128 x 128 tile on 16 warps of 32 x 32 (4 warps per scheduler)
using MidTile = TileCfg<2, 2, 8, 4, 3, 2>;     // 128 x 64 tile, 4 warps of 64 x 32, 2 CTAs per SM
using SmallTile = TileCfg<2, 2, 4, 4, 3, 3>;   // 3 stages x 20 KB = 60 KB -> 3 CTAs (12 warps) per SM

template <bool PK, bool QK, class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MIN_BLOCKS) gemm_f64_dmma_kernel(const GemmDesc g) {
    constexpr int TR = Cfg::TR, TS = Cfg::TS, NT = Cfg::NT, BI = Cfg::BI, BJ = Cfg::BJ;
    constexpr int LDRP = Cfg::LDRP, LDRQ = Cfg::LDRQ, STAGES = Cfg::STAGES;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int tilesS = g.S / TS, tilesR = g.R / TR;
    int bid = blockIdx.x;
    if (g.heavy_last) bid = gridDim.x - 1 - bid;
    const int group = bid / (GROUP_S * tilesR);
    const int first_s = group * GROUP_S;
    const int gs = min(GROUP_S, tilesS - first_s);
    const int within = bid - group * GROUP_S * tilesR;
    const int tr = within / gs, ts = first_s + within % gs;
    const int r0 = tr * TR, s0 = ts * TS;
    if (g.tri_out && s0 + TS <= r0) return;
    int kbeg = 0, kend = g.K;
    if (g.kb_r) kbeg = max(kbeg, r0);
    if (g.kb_s) kbeg = max(kbeg, s0);
    if (g.ke_r) kend = min(kend, r0 + TR);
    if (g.ke_s) kend = min(kend, s0 + TS);
    const int nk = (kend - kbeg) / TK;

    const int64_t b = blockIdx.y;
    const double* __restrict__ P = g.P + b * g.strideP;
    const double* __restrict__ Q = g.Q + b * g.strideQ;
    double* __restrict__ D = g.D + b * g.strideD;
    const double* Cin = g.Cin ? g.Cin + b * g.strideC : nullptr;

    double* Ps = smem;
    double* Qs = smem + STAGES * Cfg::P_STAGE;

    const int warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int wr0 = (warp / Cfg::WS) * (BI * 8), ws0 = (warp % Cfg::WS) * (BJ * 8);

    double acc[BI][BJ][2];
#pragma unroll
    for (int i = 0; i < BI; ++i)
#pragma unroll
        for (int j = 0; j < BJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // Per-thread copy descriptors: running global pointers (advanced by one k-chunk per issue) and fixed offsets inside a
    // stage, so the steady-state loop carries no 64-bit address arithmetic and no `% STAGES`.
    constexpr int PPT = TR * TK / 2 / NT, QPT = TS * TK / 2 / NT;
    const double* pg[PPT];
    const double* qg[QPT];
    int pso[PPT], qso[QPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int c = tid + i * NT;
        if (!PK) { const int k = c / (TR / 2), x = (c % (TR / 2)) * 2; pso[i] = k * LDRP + x; pg[i] = P + (int64_t)(kbeg + k) * g.ldp + (r0 + x); }
        else     { const int x = c >> 3, k = (c & 7) * 2;              pso[i] = x * LDK + k;  pg[i] = P + (int64_t)(r0 + x) * g.ldp + (kbeg + k); }
    }
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
        const int c = tid + i * NT;
        if (!QK) { const int k = c / (TS / 2), x = (c % (TS / 2)) * 2; qso[i] = k * LDRQ + x; qg[i] = Q + (int64_t)(kbeg + k) * g.ldq + (s0 + x); }
        else     { const int x = c >> 3, k = (c & 7) * 2;              qso[i] = x * LDK + k;  qg[i] = Q + (int64_t)(s0 + x) * g.ldq + (kbeg + k); }
    }
    const int64_t pstep = PK ? (int64_t)TK : (int64_t)TK * g.ldp;
    const int64_t qstep = QK ? (int64_t)TK : (int64_t)TK * g.ldq;
    auto issue = [&](int stage) {
        double* pst = Ps + stage * Cfg::P_STAGE;
        double* qst = Qs + stage * Cfg::Q_STAGE;
#pragma unroll
        for (int i = 0; i < PPT; ++i) { cp_async16(pst + pso[i], pg[i]); pg[i] += pstep; }
#pragma unroll
        for (int i = 0; i < QPT; ++i) { cp_async16(qst + qso[i], qg[i]); qg[i] += qstep; }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) issue(s);
        cp_async_commit();
    }

    int cstage = 0, lstage = STAGES - 1;   // stage computed on / stage loaded into, both cycle through 0..STAGES-1
    for (int kc = 0; kc < nk; ++kc) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const double* ps = Ps + cstage * Cfg::P_STAGE;
        const double* qs = Qs + cstage * Cfg::Q_STAGE;
        const double* pf = PK ? ps + (wr0 + gid) * LDK + tig : ps + tig * LDRP + wr0 + gid;
        const double* qf = QK ? qs + (ws0 + gid) * LDK + tig : qs + tig * LDRQ + ws0 + gid;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            double a[BI], bb[BJ];
#pragma unroll
            for (int i = 0; i < BI; ++i) a[i] = PK ? pf[i * 8 * LDK + kk] : pf[kk * LDRP + i * 8];
#pragma unroll
            for (int j = 0; j < BJ; ++j) bb[j] = QK ? qf[j * 8 * LDK + kk] : qf[kk * LDRQ + j * 8];
#pragma unroll
            for (int i = 0; i < BI; ++i)
#pragma unroll
                for (int j = 0; j < BJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
            if (kk == 0) {
                // refill the stage that was consumed in the previous iteration (every warp is past the barrier above, so
                // nobody reads it any more); issued behind the first DMMA group so the copies overlap the math
                if (kc + STAGES - 1 < nk) issue(lstage);
                cp_async_commit();
            }
        }
        cstage = (cstage + 1 == STAGES) ? 0 : cstage + 1;
        lstage = (lstage + 1 == STAGES) ? 0 : lstage + 1;
    }
    cp_async_wait<0>();

    const double alpha = g.alpha, beta = g.beta;
#pragma unroll
    for (int i = 0; i < BI; ++i) {
        const int64_t r = r0 + wr0 + 8 * i + gid;
#pragma unroll
        for (int j = 0; j < BJ; ++j) {
            const int s = s0 + ws0 + 8 * j + 2 * tig;
            double2 v = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
            if (Cin != nullptr) {
                const double2 c = *reinterpret_cast<const double2*>(Cin + r * g.ldc + s);
                v.x += beta * c.x;
                v.y += beta * c.y;
            }
            *reinterpret_cast<double2*>(D + r * g.ldd + s) = v;
        }
    }
}

template <bool PK, bool QK, class Cfg>
int launch(gpk_handle h, const GemmDesc& g, int cfg_id) {
    const unsigned bit = 1u << (12 + cfg_id * 4 + (PK ? 2 : 0) + (QK ? 1 : 0));
    if (!(h->func_cfg & bit)) {
        GPK_CUDA(h, cudaFuncSetAttribute(gemm_f64_dmma_kernel<PK, QK, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM));
        h->func_cfg |= bit;
    }
    dim3 grid((unsigned)((g.R / Cfg::TR) * (g.S / Cfg::TS)), (unsigned)g.batch);
    gemm_f64_dmma_kernel<PK, QK, Cfg><<<grid, Cfg::NT, Cfg::SMEM, h->stream>>>(g);
    GPK_LAUNCH_CHECK(h);
    return GPK_OK;
}

template <class Cfg>
int dispatch(gpk_handle h, const GemmDesc& g, int cfg_id) {
    if (g.p_kcontig) return g.q_kcontig ? launch<true, true, Cfg>(h, g, cfg_id) : launch<true, false, Cfg>(h, g, cfg_id);
    return g.q_kcontig ? launch<false, true, Cfg>(h, g, cfg_id) : launch<false, false, Cfg>(h, g, cfg_id);
}

int small_tile_threshold() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPK_SMALL_TILE_THRESHOLD");
        // Measured on B200 (profiles/r01_tile_configs.txt): 64x64 tiles at 3 CTAs/SM beat the 128x128 (1 CTA/SM) and
        // 128x64 (2 CTAs/SM) configurations at every size (SYRK n=k=4096: 31.2 vs 26.8 vs 29.1 TFLOP/s), so they are
        // the default everywhere; the threshold stays as a tuning knob.
        v = e ? atoi(e) : 0x7fffffff;
    }
    return v;
}

}  // namespace

int gpk_gemm(gpk_handle h, const GemmDesc& g) {
    if (g.R <= 0 || g.S <= 0 || g.batch <= 0) return GPK_OK;
    if (g.R % 128 || g.S % 128 || g.K % TK || (g.ldp & 1) || (g.ldq & 1) || (g.ldd & 1) ||
        ((uintptr_t)g.P & 15) || ((uintptr_t)g.Q & 15) || ((uintptr_t)g.D & 15) || (g.Cin && (((uintptr_t)g.Cin & 15) || (g.ldc & 1))))
        return gpk_set_error(h, GPK_EINVAL, "gpk_gemm: unaligned problem R=%d S=%d K=%d", g.R, g.S, g.K);
    int64_t tiles = (int64_t)(g.R / 128) * (g.S / 128) * g.batch;
    if (g.tri_out) tiles = (tiles + g.batch * (g.R / 128)) / 2;
    if (tiles < small_tile_threshold()) return dispatch<SmallTile>(h, g, 1);
    static int big_kind = -1;
    if (big_kind < 0) { const char* e = getenv("GPK_BIG_KIND"); big_kind = e ? atoi(e) : 0; }
    if (big_kind == 1) return dispatch<WideTile>(h, g, 2);
    if (big_kind == 2) return dispatch<MidTile>(h, g, 3);
    return dispatch<BigTile>(h, g, 0);
}

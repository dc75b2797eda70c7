// gpk_mg.cu -- one large GP on every GPU of a node behind the C ABI (BASELINE.json config 5; SURVEY.md 8(b) gpk_mg_create /
// gpk_mg_potrf_solve): K build, FP64 Cholesky, alpha = K^-1 y and the log marginal likelihood of
// GpPredictor.preComputeComponents + logLikelihood (gp/regression/GpPredictor.scala:104-124,144-149) for a training set too
// large for one device to factor quickly.  The reference has no multi-device code; a JVM caller binds these two entry points
// like every other gpk_* symbol (one process, no launcher, no Python).
//
// Design (single process, ONE host thread enqueueing on all devices; everything is asynchronous on streams and events):
//   * block columns of width nb are dealt round-robin to the G devices (column j on device j mod G); device g stores its
//     columns full height, N x nb each, and builds them itself from the replicated X (8 n D bytes) -- K never moves;
//   * right-looking factorisation with look-ahead 1.  Step k, on the owner's high-priority stream: factor + invert the diagonal
//     block (gpk_potrf_inv_block_dev), solve the whole panel L_ik = A_ik L_kk^-T as ONE DMMA GEMM with that inverse, then PUT
//     the panel (with L_kk^-1 riding in its top block) into every peer's panel buffer over NVLink: cudaMemcpy2DAsync on
//     peer-mapped memory (cudaDeviceEnablePeerAccess), issued on the owner's copy stream, next owner first.  There is no
//     collective and no rendezvous: a put waits for the peer's "buffer free" events, the peer's consumers wait for the put's
//     event.  (The multi-process path, gp_algos_b200/distributed.py, moves the same panels with NCCL broadcasts.)
//   * every device updates its own columns j > k with panel k, one DMMA GEMM per column: column k+1 on its owner's
//     high-priority stream (so that step k+1 starts at once), the others on two low-priority lanes; panel buffers are triple-
//     buffered;
//   * the forward solve rides along replicated (every device holds every panel once); the back solve walks the columns in
//     reverse on their owners and puts each alpha block (nb doubles) to all peers.
// Traffic per device: n^2/2 * 8 B received in total (17 GB at n = 65536), (G-1)/G of that sent; flops n^3/3 / G + the
// replicated O(n^2) solves.
#include "gpk_internal.cuh"

#include <math.h>
#include <stdlib.h>

#include <new>
#include <vector>

namespace {

constexpr int MG_MAXDEV = 16;
constexpr int MG_NLANE = 2;
constexpr int MG_NSLOT = 3;
constexpr int MG_NCOPY = 3;           // put streams per device: the next owner gets its own, the other peers alternate on the rest

struct MgDev {
    int dev;
    gpk_handle h;                    // main (high-priority) stream
    gpk_handle lane[MG_NLANE];       // low-priority bulk lanes
    cudaStream_t lane_stream[MG_NLANE];
    cudaStream_t copy[MG_NCOPY];     // outgoing puts
    double* buf;                     // one allocation, carved below
    size_t buf_bytes;
    double *A, *panel[MG_NSLOT], *Dbuf, *LiAll, *X, *y0, *ypad, *z, *alpha, *tmp, *scal;
    int* info;
    int ncols;
};

}  // namespace

struct gpk_mg_s {
    int G;
    int nb;
    MgDev d[MG_MAXDEV];
    double last_seconds;
    int last_info;
    int64_t put_bytes;               // bytes put to peers by the last solve (all devices)
    char err[512];
};

namespace {

int mg_error(gpk_mg mg, int status, const char* fmt, const char* detail) {
    if (mg) snprintf(mg->err, sizeof(mg->err), fmt, detail ? detail : "");
    return status;
}

#define MG_CUDA(mg, call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            snprintf((mg)->err, sizeof((mg)->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return e__ == cudaErrorMemoryAllocation ? GPK_ENOMEM : GPK_ECUDA;                     \
        }                                                                                         \
    } while (0)

#define MG_GPK(mg, dv, call)                                                                      \
    do {                                                                                          \
        int rc__ = (call);                                                                        \
        if (rc__ != GPK_OK) {                                                                     \
            snprintf((mg)->err, sizeof((mg)->err), "device %d: %s", (dv).dev, gpk_last_error((dv).h)); \
            return rc__;                                                                          \
        }                                                                                         \
    } while (0)

struct EventPool {                   // events of one solve, per device, destroyed at the end
    std::vector<cudaEvent_t> ev[MG_MAXDEV];
    cudaEvent_t get(gpk_mg mg, int g) {
        cudaSetDevice(mg->d[g].dev);
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ev[g].push_back(e);
        return e;
    }
    void destroy(gpk_mg mg) {
        for (int g = 0; g < mg->G; ++g) {
            cudaSetDevice(mg->d[g].dev);
            for (cudaEvent_t e : ev[g]) cudaEventDestroy(e);
            ev[g].clear();
        }
    }
};

int mg_alloc(gpk_mg mg, int g, int N, int nb, int D, int ncols) {
    MgDev& dv = mg->d[g];
    const size_t nA = (size_t)N * nb * (ncols > 0 ? ncols : 1);
    const size_t total = nA + (size_t)MG_NSLOT * N * nb + (size_t)nb * nb + (size_t)(ncols > 0 ? ncols : 1) * nb * nb + (size_t)N * D +
                         (size_t)5 * N + nb + 64;
    const size_t bytes = total * sizeof(double) + (size_t)(ncols + 1) * sizeof(int);
    MG_CUDA(mg, cudaSetDevice(dv.dev));
    if (dv.buf_bytes < bytes) {
        if (dv.buf) { cudaDeviceSynchronize(); cudaFree(dv.buf); dv.buf = nullptr; dv.buf_bytes = 0; }
        MG_CUDA(mg, cudaMalloc((void**)&dv.buf, bytes));
        dv.buf_bytes = bytes;
    }
    double* p = dv.buf;
    dv.A = p; p += nA;
    for (int s = 0; s < MG_NSLOT; ++s) { dv.panel[s] = p; p += (size_t)N * nb; }
    dv.Dbuf = p; p += (size_t)nb * nb;
    dv.LiAll = p; p += (size_t)(ncols > 0 ? ncols : 1) * nb * nb;
    dv.X = p; p += (size_t)N * D;
    dv.y0 = p; p += N; dv.ypad = p; p += N; dv.z = p; p += N; dv.alpha = p; p += N; dv.tmp = p; p += N + nb;
    dv.scal = p; p += 64;
    dv.info = (int*)p;
    dv.ncols = ncols;
    return GPK_OK;
}

// cudaMemcpy2DAsync of a (rows x cols) column-major block, any two devices of the process (UVA + peer access)
inline cudaError_t copy_block(double* dst, int64_t ldd, const double* src, int64_t lds, int rows, int cols, cudaStream_t st) {
    return cudaMemcpy2DAsync(dst, (size_t)ldd * sizeof(double), src, (size_t)lds * sizeof(double), (size_t)rows * sizeof(double),
                             (size_t)cols, cudaMemcpyDefault, st);
}

}  // namespace

extern "C" {

int gpk_mg_create(gpk_mg* out, int ndev, const int* devices) {
    if (!out) return GPK_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return GPK_ECUDA;
    if (ndev <= 0) ndev = count;
    if (ndev > count || ndev > MG_MAXDEV) return GPK_EINVAL;
    gpk_mg mg = new (std::nothrow) gpk_mg_s();
    if (!mg) return GPK_ENOMEM;
    memset(mg, 0, sizeof(*mg));
    mg->G = ndev;
    mg->nb = 0;      // automatic
    for (int g = 0; g < ndev; ++g) {
        MgDev& dv = mg->d[g];
        dv.dev = devices ? devices[g] : g;
        if (dv.dev < 0 || dv.dev >= count || cudaSetDevice(dv.dev) != cudaSuccess || gpk_create(&dv.h, dv.dev, nullptr) != GPK_OK) {
            gpk_mg_destroy(mg);
            return GPK_ECUDA;
        }
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        for (int l = 0; l < MG_NLANE; ++l) {
            if (cudaStreamCreateWithPriority(&dv.lane_stream[l], cudaStreamNonBlocking, least) != cudaSuccess ||
                gpk_create(&dv.lane[l], dv.dev, dv.lane_stream[l]) != GPK_OK) { gpk_mg_destroy(mg); return GPK_ECUDA; }
            gpk_set_graph_mode(dv.lane[l], 0);
        }
        gpk_set_graph_mode(dv.h, 0);
        for (int c = 0; c < MG_NCOPY; ++c)
            if (cudaStreamCreateWithPriority(&dv.copy[c], cudaStreamNonBlocking, greatest) != cudaSuccess) { gpk_mg_destroy(mg); return GPK_ECUDA; }
    }
    // peer mappings for the puts (NVLink / NVSwitch on a B200 node); without them cudaMemcpy2DAsync stages through the host
    for (int g = 0; g < ndev; ++g) {
        cudaSetDevice(mg->d[g].dev);
        for (int q = 0; q < ndev; ++q) {
            if (q == g) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, mg->d[g].dev, mg->d[q].dev);
            if (can && cudaDeviceEnablePeerAccess(mg->d[q].dev, 0) != cudaSuccess) cudaGetLastError();   // already enabled is fine
        }
    }
    *out = mg;
    return GPK_OK;
}

int gpk_mg_destroy(gpk_mg mg) {
    if (!mg) return GPK_OK;
    for (int g = 0; g < mg->G; ++g) {
        MgDev& dv = mg->d[g];
        cudaSetDevice(dv.dev);
        cudaDeviceSynchronize();
        for (int l = 0; l < MG_NLANE; ++l) {
            if (dv.lane[l]) gpk_destroy(dv.lane[l]);
            if (dv.lane_stream[l]) cudaStreamDestroy(dv.lane_stream[l]);
        }
        if (dv.h) gpk_destroy(dv.h);
        for (int c = 0; c < MG_NCOPY; ++c)
            if (dv.copy[c]) cudaStreamDestroy(dv.copy[c]);
        if (dv.buf) cudaFree(dv.buf);
    }
    delete mg;
    return GPK_OK;
}

const char* gpk_mg_last_error(gpk_mg mg) { return mg ? mg->err : "null multi-GPU handle"; }
int gpk_mg_device_count(gpk_mg mg) { return mg ? mg->G : 0; }
double gpk_mg_last_seconds(gpk_mg mg) { return mg ? mg->last_seconds : 0.0; }
int64_t gpk_mg_last_put_bytes(gpk_mg mg) { return mg ? mg->put_bytes : 0; }
int gpk_mg_set_block(gpk_mg mg, int nb) {
    if (!mg || nb < 0 || (nb > 0 && (nb < GPK_TILE || nb % GPK_TILE)))
        return mg_error(mg, GPK_EINVAL, "block width must be 0 (automatic) or a positive multiple of 128%s", nullptr);
    mg->nb = nb;
    return GPK_OK;
}

int gpk_mg_potrf_solve(gpk_mg mg, const double* X, int n, int D, int64_t ldx, const double* y, const double* theta,
                       int has_s, double s, double* alpha_out, double* ll_out, int* info_out) {
    if (!mg || !X || !y || !theta || n <= 0 || ldx < n) return mg_error(mg, GPK_EINVAL, "gpk_mg_potrf_solve: bad arguments%s", nullptr);
    if (D < 1 || D > GPK_MAX_D) return mg_error(mg, GPK_EINVAL, "gpk_mg_potrf_solve: feature dimension outside 1..64%s", nullptr);
    const int G = mg->G;
    // automatic block width: at least ~16 block columns per device (the serial chain factor -> panel -> put shrinks with the
    // width, the bulk GEMMs want it large): 1024 up to 4 devices at n = 65536, 512 on 8 (measured 0.431 vs 0.417 s)
    int nb = mg->nb;
    if (nb <= 0) {
        nb = (n / (16 * G)) / GPK_TILE * GPK_TILE;
        nb = nb < 256 ? 256 : (nb > 1024 ? 1024 : nb);
    }
    const int nt = (n + nb - 1) / nb, N = nt * nb;
    const double sn2 = theta[D + 1] * theta[D + 1] + (has_s ? s : 0.0);          // KernelRequisites.scala:69 + GpPredictor.scala:116
    mg->err[0] = 0; mg->last_info = 0; mg->put_bytes = 0;
    if (info_out) *info_out = 0;

    // ---- workspace, inputs --------------------------------------------------------------------------------------------------
    std::vector<double> Xp((size_t)N * D, 0.0), yp((size_t)N, 0.0);
    for (int c = 0; c < D; ++c) memcpy(&Xp[(size_t)c * N], X + (size_t)c * ldx, (size_t)n * sizeof(double));
    memcpy(yp.data(), y, (size_t)n * sizeof(double));
    for (int g = 0; g < G; ++g) {
        const int ncols = g < nt ? (nt - 1 - g) / G + 1 : 0;
        int rc = mg_alloc(mg, g, N, nb, D, ncols);
        if (rc) return rc;
        MgDev& dv = mg->d[g];
        MG_CUDA(mg, cudaMemcpyAsync(dv.X, Xp.data(), (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice, dv.h->stream));
        MG_CUDA(mg, cudaMemcpyAsync(dv.y0, yp.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, dv.h->stream));
        MG_CUDA(mg, cudaMemcpyAsync(dv.ypad, dv.y0, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, dv.h->stream));
        MG_CUDA(mg, cudaMemsetAsync(dv.scal, 0, 64 * sizeof(double), dv.h->stream));
        MG_CUDA(mg, cudaMemsetAsync(dv.z, 0, (size_t)2 * N * sizeof(double), dv.h->stream));      // z, alpha
        MG_CUDA(mg, cudaMemsetAsync(dv.info, 0, (size_t)(ncols + 1) * sizeof(int), dv.h->stream));
    }
    for (int g = 0; g < G; ++g) { MG_CUDA(mg, cudaSetDevice(mg->d[g].dev)); MG_CUDA(mg, cudaDeviceSynchronize()); }
    EventPool pool;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    MG_CUDA(mg, cudaSetDevice(mg->d[0].dev));
    MG_CUDA(mg, cudaEventCreate(&t0));
    MG_CUDA(mg, cudaEventCreate(&t1));
    MG_CUDA(mg, cudaEventRecord(t0, mg->d[0].h->stream));
    int rc = GPK_OK;
    auto body = [&]() -> int {
        // ---- K blocks from the replicated X (MatrixUtils.scala:57-70), each device its own columns ------------------------------
        for (int g = 0; g < G; ++g) {
            MgDev& dv = mg->d[g];
            MG_CUDA(mg, cudaSetDevice(dv.dev));
            for (int lj = 0; lj < dv.ncols; ++lj) {
                const int j = lj * G + g, bj = j * nb;
                double* col = dv.A + (size_t)lj * nb * N;
                MG_GPK(mg, dv, gpk_cov_cross_se_ard_dev(dv.h, dv.X + bj, N - bj, N, dv.X + bj, nb, N, D, theta, col + bj, N));
                MG_GPK(mg, dv, gpk_add_diag_dev(dv.h, col + bj, N, nb, sn2));
            }
            if (N > n) {                                           // padding rows / columns -> identity
                const int pad0 = n - (nt - 1) * nb;                // first padded row inside the last block
                for (int lj = 0; lj < dv.ncols; ++lj) {
                    double* col = dv.A + (size_t)lj * nb * N;
                    MG_CUDA(mg, cudaMemset2DAsync(col + n, (size_t)N * sizeof(double), 0, (size_t)(N - n) * sizeof(double), nb, dv.h->stream));
                }
                if ((nt - 1) % G == g) {
                    double* col = dv.A + (size_t)((nt - 1) / G) * nb * N;
                    MG_CUDA(mg, cudaMemset2DAsync(col + (size_t)pad0 * N, (size_t)N * sizeof(double), 0, (size_t)N * sizeof(double), nb - pad0,
                                                  dv.h->stream));
                    MG_GPK(mg, dv, gpk_add_diag_dev(dv.h, col + n + (size_t)pad0 * N, N, N - n, 1.0));
                }
            }
            cudaEvent_t built = pool.get(mg, g);
            MG_CUDA(mg, cudaEventRecord(built, dv.h->stream));
            for (int l = 0; l < MG_NLANE; ++l) MG_CUDA(mg, cudaStreamWaitEvent(dv.lane_stream[l], built, 0));
            for (int c = 0; c < MG_NCOPY; ++c) MG_CUDA(mg, cudaStreamWaitEvent(dv.copy[c], built, 0));
        }
        // events: arr[g][k] panel k is in device g's buffer; used[g][k][0..NLANE] device g no longer reads panel k
        std::vector<std::vector<cudaEvent_t>> arr(G, std::vector<cudaEvent_t>(nt, nullptr));
        std::vector<std::vector<std::vector<cudaEvent_t>>> used(G, std::vector<std::vector<cudaEvent_t>>(nt));
        std::vector<std::vector<cudaEvent_t>> colev(G);           // last lane update of each local column
        for (int g = 0; g < G; ++g) colev[g].assign(mg->d[g].ncols > 0 ? mg->d[g].ncols : 1, nullptr);

        auto factor_panel = [&](int k) -> int {                   // column k is fully updated on its owner's main stream
            const int o = k % G, lk = k / G, bk = k * nb, b1 = bk + nb, slot = k % MG_NSLOT;
            MgDev& dv = mg->d[o];
            MG_CUDA(mg, cudaSetDevice(dv.dev));
            cudaStream_t M = dv.h->stream;
            if (k >= MG_NSLOT) for (cudaEvent_t e : used[o][k - MG_NSLOT]) MG_CUDA(mg, cudaStreamWaitEvent(M, e, 0));
            double* col = dv.A + (size_t)lk * nb * N;
            double* Li = dv.LiAll + (size_t)lk * nb * nb;
            MG_CUDA(mg, copy_block(dv.Dbuf, nb, col + bk, N, nb, nb, M));
            MG_GPK(mg, dv, gpk_potrf_inv_block_dev(dv.h, dv.Dbuf, Li, nb, dv.info + lk));
            MG_CUDA(mg, copy_block(col + bk, N, dv.Dbuf, nb, nb, nb, M));
            MG_GPK(mg, dv, gpk_sum_log_diag_dev(dv.h, dv.Dbuf, nb, nb, dv.scal, 1));
            double* pan = dv.panel[slot];
            MG_CUDA(mg, copy_block(pan + bk, N, Li, nb, nb, nb, M));                       // L_kk^-1 rides in the panel's top block
            if (b1 < N) {
                MG_GPK(mg, dv, gpk_gemm_nt_dev(dv.h, N - b1, nb, nb, 1.0, col + b1, N, Li, nb, 0.0, pan + b1, N, 1));   // L_ik = A_ik L_kk^-T
                MG_CUDA(mg, copy_block(col + b1, N, pan + b1, N, N - b1, nb, M));             // keep L in place for the back solve
            }
            arr[o][k] = pool.get(mg, o);
            MG_CUDA(mg, cudaEventRecord(arr[o][k], M));
            if (G > 1) {
                for (int c = 0; c < MG_NCOPY; ++c) MG_CUDA(mg, cudaStreamWaitEvent(dv.copy[c], arr[o][k], 0));
                for (int q = 1; q < G; ++q) {                                                 // next owner first, on its own stream
                    const int g = (o + q) % G;
                    cudaStream_t cs = dv.copy[q == 1 ? 0 : 1 + (q % (MG_NCOPY - 1))];
                    if (k >= MG_NSLOT) for (cudaEvent_t e : used[g][k - MG_NSLOT]) MG_CUDA(mg, cudaStreamWaitEvent(cs, e, 0));
                    MG_CUDA(mg, copy_block(mg->d[g].panel[slot] + bk, N, pan + bk, N, N - bk, nb, cs));
                    arr[g][k] = pool.get(mg, o);
                    MG_CUDA(mg, cudaEventRecord(arr[g][k], cs));
                    mg->put_bytes += (int64_t)(N - bk) * nb * 8;
                }
            }
            return GPK_OK;
        };

        int r = factor_panel(0);
        if (r) return r;
        for (int k = 0; k < nt; ++k) {
            const int bk = k * nb, b1 = bk + nb, slot = k % MG_NSLOT;
            for (int q = 0; q < G; ++q) {
                const int g = (k + 1 + q) % G;                                                // the next owner's launches first
                MgDev& dv = mg->d[g];
                MG_CUDA(mg, cudaSetDevice(dv.dev));
                cudaStream_t M = dv.h->stream;
                const double* pan = dv.panel[slot];
                MG_CUDA(mg, cudaStreamWaitEvent(M, arr[g][k], 0));
                // forward substitution, replicated (MatrixUtils.scala:17-21): z_k = L_kk^-1 y_k ; y_i -= L_ik z_k
                MG_GPK(mg, dv, gpk_gemv_dev(dv.h, 0, nb, nb, 1.0, pan + bk, N, dv.ypad + bk, 0.0, dv.z + bk));
                if (b1 < N) MG_GPK(mg, dv, gpk_gemv_dev(dv.h, 0, N - b1, nb, -1.0, pan + b1, N, dv.z + bk, 1.0, dv.ypad + b1));
                bool lane_waited[MG_NLANE] = {false, false};
                for (int lj = 0; lj < dv.ncols; ++lj) {
                    const int j = lj * G + g, bj = j * nb;
                    if (j <= k) continue;
                    double* tgt = dv.A + (size_t)lj * nb * N + bj;
                    if (j == k + 1) {                                                         // look-ahead column: main stream
                        if (colev[g][lj]) MG_CUDA(mg, cudaStreamWaitEvent(M, colev[g][lj], 0));
                        MG_GPK(mg, dv, gpk_gemm_nt_dev(dv.h, N - bj, nb, nb, -1.0, pan + bj, N, pan + bj, N, 1.0, tgt, N, 0));
                    } else {
                        const int l = lj % MG_NLANE;
                        if (!lane_waited[l]) { MG_CUDA(mg, cudaStreamWaitEvent(dv.lane_stream[l], arr[g][k], 0)); lane_waited[l] = true; }
                        MG_GPK(mg, dv, gpk_gemm_nt_dev(dv.lane[l], N - bj, nb, nb, -1.0, pan + bj, N, pan + bj, N, 1.0, tgt, N, 0));
                        colev[g][lj] = pool.get(mg, g);
                        MG_CUDA(mg, cudaEventRecord(colev[g][lj], dv.lane_stream[l]));
                    }
                }
                cudaEvent_t em = pool.get(mg, g);
                MG_CUDA(mg, cudaEventRecord(em, M));
                used[g][k].push_back(em);
                for (int l = 0; l < MG_NLANE; ++l)
                    if (lane_waited[l]) {
                        cudaEvent_t el = pool.get(mg, g);
                        MG_CUDA(mg, cudaEventRecord(el, dv.lane_stream[l]));
                        used[g][k].push_back(el);
                    }
                if (q == 0 && k + 1 < nt) {                                                   // panel k+1 goes out before the others' bulk
                    r = factor_panel(k + 1);
                    if (r) return r;
                }
            }
        }
        // ---- back solve L^t alpha = z (MatrixUtils.scala:23-27), block columns in reverse on their owners ---------------------
        std::vector<cudaEvent_t> aev(nt, nullptr);
        for (int k = nt - 1; k >= 0; --k) {
            const int o = k % G, lk = k / G, bk = k * nb, b1 = bk + nb;
            MgDev& dv = mg->d[o];
            MG_CUDA(mg, cudaSetDevice(dv.dev));
            cudaStream_t M = dv.h->stream;
            for (int j = k + 1; j < nt && j <= k + G; ++j)                                    // alpha blocks put by the later owners
                if (aev[j] && j % G != o) MG_CUDA(mg, cudaStreamWaitEvent(M, aev[j], 0));
            MG_CUDA(mg, cudaMemcpyAsync(dv.tmp, dv.z + bk, (size_t)nb * sizeof(double), cudaMemcpyDeviceToDevice, M));
            const double* col = dv.A + (size_t)lk * nb * N;
            if (b1 < N) MG_GPK(mg, dv, gpk_gemv_dev(dv.h, 1, N - b1, nb, -1.0, col + b1, N, dv.alpha + b1, 1.0, dv.tmp));
            MG_GPK(mg, dv, gpk_gemv_dev(dv.h, 1, nb, nb, 1.0, dv.LiAll + (size_t)lk * nb * nb, nb, dv.tmp, 0.0, dv.alpha + bk));
            for (int q = 1; q < G; ++q) {
                const int g = (o + q) % G;
                MG_CUDA(mg, cudaMemcpyAsync(mg->d[g].alpha + bk, dv.alpha + bk, (size_t)nb * sizeof(double), cudaMemcpyDefault, M));
            }
            aev[k] = pool.get(mg, o);
            MG_CUDA(mg, cudaEventRecord(aev[k], M));
        }
        // ---- y . alpha on device 0; join every device's streams into device 0's main stream ----------------------------------
        MgDev& d0 = mg->d[0];
        MG_CUDA(mg, cudaSetDevice(d0.dev));
        for (int j = 0; j < nt && j < G; ++j) if (aev[j]) MG_CUDA(mg, cudaStreamWaitEvent(d0.h->stream, aev[j], 0));
        MG_GPK(mg, d0, gpk_gemv_dev(d0.h, 1, N, 1, 1.0, d0.y0, N, d0.alpha, 0.0, d0.scal + 1));
        for (int g = 1; g < G; ++g) {
            MgDev& dv = mg->d[g];
            MG_CUDA(mg, cudaSetDevice(dv.dev));
            for (int l = 0; l < MG_NLANE; ++l) {
                cudaEvent_t e = pool.get(mg, g);
                MG_CUDA(mg, cudaEventRecord(e, dv.lane_stream[l]));
                MG_CUDA(mg, cudaStreamWaitEvent(dv.h->stream, e, 0));
            }
            cudaEvent_t e = pool.get(mg, g);
            MG_CUDA(mg, cudaEventRecord(e, dv.h->stream));
            MG_CUDA(mg, cudaSetDevice(d0.dev));
            MG_CUDA(mg, cudaStreamWaitEvent(d0.h->stream, e, 0));
        }
        MG_CUDA(mg, cudaSetDevice(d0.dev));
        for (int l = 0; l < MG_NLANE; ++l) {
            cudaEvent_t e = pool.get(mg, 0);
            MG_CUDA(mg, cudaEventRecord(e, d0.lane_stream[l]));
            MG_CUDA(mg, cudaStreamWaitEvent(d0.h->stream, e, 0));
        }
        MG_CUDA(mg, cudaEventRecord(t1, d0.h->stream));
        return GPK_OK;
    };
    rc = body();
    for (int g = 0; g < G; ++g) { cudaSetDevice(mg->d[g].dev); cudaDeviceSynchronize(); }
    if (rc == GPK_OK) {
        float ms = 0.f;
        cudaSetDevice(mg->d[0].dev);
        if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) mg->last_seconds = ms * 1e-3;
    }
    cudaSetDevice(mg->d[0].dev);
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
    pool.destroy(mg);
    if (rc != GPK_OK) return rc;
    if (cudaGetLastError() != cudaSuccess) return mg_error(mg, GPK_ECUDA, "asynchronous CUDA error during the factorisation%s", nullptr);

    // ---- results --------------------------------------------------------------------------------------------------------------
    double logdet = 0.0, ya = 0.0;
    int minor = 0;
    for (int g = 0; g < G; ++g) {
        MgDev& dv = mg->d[g];
        MG_CUDA(mg, cudaSetDevice(dv.dev));
        double sc[2];
        MG_CUDA(mg, cudaMemcpy(sc, dv.scal, sizeof(sc), cudaMemcpyDeviceToHost));
        logdet += sc[0];
        if (g == 0) ya = sc[1];
        std::vector<int> inf((size_t)dv.ncols + 1, 0);
        if (dv.ncols) MG_CUDA(mg, cudaMemcpy(inf.data(), dv.info, (size_t)dv.ncols * sizeof(int), cudaMemcpyDeviceToHost));
        for (int lj = 0; lj < dv.ncols; ++lj)
            if (inf[lj]) {
                const int m = (lj * G + g) * nb + inf[lj];
                if (!minor || m < minor) minor = m;
            }
    }
    if (minor) {
        mg->last_info = minor;
        if (info_out) *info_out = minor;
        snprintf(mg->err, sizeof(mg->err), "matrix not positive definite: leading minor %d", minor);
        return GPK_ENOTPD;
    }
    if (alpha_out) {
        MG_CUDA(mg, cudaSetDevice(mg->d[0].dev));
        MG_CUDA(mg, cudaMemcpy(alpha_out, mg->d[0].alpha, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (ll_out) *ll_out = -0.5 * ya - logdet - 0.5 * n * log(2.0 * 3.14159265358979323846);   // GpPredictor.scala:144-149
    return GPK_OK;
}

}  // extern "C"

// gpk_part.cu -- SM partition for the spine of the look-ahead factorisation (CUDA green contexts).
//
// The look-ahead Cholesky drivers (gpk_chol.cu) keep a high-priority "spine" (diagonal block, next panel, next column update)
// ahead of low-priority bulk GEMMs.  Stream priorities order PENDING CTAs, they do not preempt: a spine kernel still waits for
// resident bulk CTAs to retire, and the 128 x 128 diagonal-block kernel needs 170 KB of shared memory, i.e. an SM on which all
// three resident bulk CTAs (60 KB each) have retired with no new one slipping in.  Measured on the EP re-factorisation
// (profiles/r02_ep_refactor_trace.log): one spine step F + Pc + Uc takes 0.35-0.5 ms beside the bulk work against 0.175 ms alone.
// Here the spine gets SMs of its own: the device's SMs are split into a small group (GPK_SPINE_SMS, default 16) and the rest,
// each wrapped in a green context; streams created in a green context run their kernels on that context's SMs only.  Memory,
// modules and events are shared with the primary context, so nothing else changes for the kernels.
//
// The driver entry points are looked up at run time (cudaGetDriverEntryPoint): libgpk.so keeps loading on machines without
// libcuda (the CPU-side ABI tests).  Any failure -- old driver, unsupported device, GPK_PARTITION=0 -- leaves the partition
// off and the drivers on the handle's ordinary streams.
#include "gpk_internal.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace {

struct DrvApi {
    CUresult (*DeviceGet)(CUdevice*, int);
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    CUresult (*GreenCtxDestroy)(CUgreenCtx);
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);
    CUresult (*StreamDestroy)(CUstream);
    bool ok;
};

template <class F>
bool entry(const char* name, F* fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

const DrvApi& drv() {
    static DrvApi d = [] {
        DrvApi a{};
        a.ok = entry("cuDeviceGet", &a.DeviceGet) && entry("cuDeviceGetDevResource", &a.DeviceGetDevResource) &&
               entry("cuDevSmResourceSplitByCount", &a.DevSmResourceSplitByCount) &&
               entry("cuDevResourceGenerateDesc", &a.DevResourceGenerateDesc) && entry("cuGreenCtxCreate", &a.GreenCtxCreate) &&
               entry("cuGreenCtxDestroy", &a.GreenCtxDestroy) && entry("cuGreenCtxStreamCreate", &a.GreenCtxStreamCreate) &&
               entry("cuStreamDestroy", &a.StreamDestroy);
        return a;
    }();
    return d;
}

}  // namespace

struct gpk_partition {
    CUgreenCtx spine_ctx, bulk_ctx;
    int spine_sms, bulk_sms;
    cudaStream_t spine;                    // the spine itself (highest priority)
    cudaStream_t spine_side[GPK_NSIDE];    // fork/join streams of the diagonal-block recursion, on the spine's SMs
    cudaStream_t bulk_near;                // near-critical bulk work (next-but-one panel / update), side priority
    cudaStream_t bulk[GPK_NPIPE];          // lowest priority
};

// 1: partition available (created on first use), 0: not available / switched off
int gpk_partition_get(gpk_handle h, gpk_partition** out) {
    *out = nullptr;
    if (h->part_state < 0) return 0;
    if (h->part_state > 0) { *out = h->part; return 1; }
    h->part_state = -1;
    const char* e = getenv("GPK_PARTITION");            // opt-in: measured gains are within a few per cent either way
    if (!e || atoi(e) == 0) return 0;
    const char* es = getenv("GPK_SPINE_SMS");
    const unsigned want = es ? (unsigned)atoi(es) : 16u;
    const DrvApi& d = drv();
    if (!d.ok || want == 0) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return 0;
    cudaFree(0);
    CUdevice dev;
    CUdevResource all, grp, rest;
    unsigned int ngroups = 1;
    if (d.DeviceGet(&dev, h->device) != CUDA_SUCCESS) return 0;
    if (d.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return 0;
    if (d.DevSmResourceSplitByCount(&grp, &ngroups, &all, &rest, 0, want) != CUDA_SUCCESS || ngroups < 1) return 0;
    if (grp.sm.smCount == 0 || rest.sm.smCount == 0) return 0;
    gpk_partition* p = new (std::nothrow) gpk_partition();
    if (!p) return 0;
    CUdevResourceDesc dsp, dbk;
    bool ok = d.DevResourceGenerateDesc(&dsp, &grp, 1) == CUDA_SUCCESS && d.DevResourceGenerateDesc(&dbk, &rest, 1) == CUDA_SUCCESS &&
              d.GreenCtxCreate(&p->spine_ctx, dsp, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS;
    if (ok && d.GreenCtxCreate(&p->bulk_ctx, dbk, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        d.GreenCtxDestroy(p->spine_ctx);
        ok = false;
    }
    if (!ok) { delete p; return 0; }
    p->spine_sms = (int)grp.sm.smCount; p->bulk_sms = (int)rest.sm.smCount;
    auto mk = [&](CUgreenCtx c, int prio, cudaStream_t* s) {
        CUstream cs = nullptr;
        if (d.GreenCtxStreamCreate(&cs, c, CU_STREAM_NON_BLOCKING, prio) != CUDA_SUCCESS) return false;
        *s = (cudaStream_t)cs;
        return true;
    };
    ok = mk(p->spine_ctx, h->prio_main, &p->spine) && mk(p->bulk_ctx, h->prio_side, &p->bulk_near);
    for (int i = 0; ok && i < GPK_NSIDE; ++i) ok = mk(p->spine_ctx, h->prio_side, &p->spine_side[i]);
    for (int i = 0; ok && i < GPK_NPIPE; ++i) ok = mk(p->bulk_ctx, h->prio_pipe, &p->bulk[i]);
    if (!ok) { gpk_partition_destroy(p); return 0; }
    h->part = p;
    h->part_state = 1;
    *out = p;
    return 1;
}

void gpk_partition_destroy(gpk_partition* p) {
    if (!p) return;
    const DrvApi& d = drv();
    auto rm = [&](cudaStream_t s) { if (s) d.StreamDestroy((CUstream)s); };
    rm(p->spine); rm(p->bulk_near);
    for (int i = 0; i < GPK_NSIDE; ++i) rm(p->spine_side[i]);
    for (int i = 0; i < GPK_NPIPE; ++i) rm(p->bulk[i]);
    if (p->spine_ctx) d.GreenCtxDestroy(p->spine_ctx);
    if (p->bulk_ctx) d.GreenCtxDestroy(p->bulk_ctx);
    delete p;
}

int gpk_partition_sms(const gpk_partition* p, int which) { return which == 0 ? p->spine_sms : p->bulk_sms; }
cudaStream_t gpk_partition_stream(const gpk_partition* p, int kind, int i) {
    switch (kind) {
        case 0: return p->spine;
        case 1: return p->spine_side[i % GPK_NSIDE];
        case 2: return p->bulk_near;
        default: return p->bulk[i % GPK_NPIPE];
    }
}

// development aid (include/gpk.h): is the partition up on this handle (creating it if GPK_PARTITION=1), and with how many SMs?
extern "C" int gpk_debug_partition(gpk_handle h, int* spine_sms, int* bulk_sms) {
    gpk_partition* p = nullptr;
    if (!h || !gpk_partition_get(h, &p)) return 0;
    if (spine_sms) *spine_sms = p->spine_sms;
    if (bulk_sms) *bulk_sms = p->bulk_sms;
    return 1;
}

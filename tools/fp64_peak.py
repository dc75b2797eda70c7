"""Measure the FP64 roofline denominators on the GPU box: cuBLAS Dgemm (torch.matmul, float64)
burst + sustained, cuSOLVER potrf/potri for context, and an HBM copy. Writes JSON to stdout.
Not part of the product path; used only to fix roofline denominators."""
import json, sys, time
import torch

def ev_time(fn, reps):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def main():
    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)
        for _ in range(3): torch.matmul(a, b, out=c)
        ms = ev_time(lambda: torch.matmul(a, b, out=c), 10)
        out[f"dgemm_{n}_burst_tflops"] = 2.0 * n**3 / ms * 1e-9
        if n == 8192:
            torch.cuda.synchronize(); t0 = time.time(); cnt = 0
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            while time.time() - t0 < 4.0:
                for _ in range(5): torch.matmul(a, b, out=c)
                cnt += 5
                torch.cuda.synchronize()
            e1.record(); e1.synchronize()
            out["dgemm_8192_sustained_tflops"] = 2.0 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
        # syrk-like A A^T
        ms = ev_time(lambda: torch.matmul(a, a.t(), out=c), 5)
        out[f"dgemm_nt_{n}_tflops"] = 2.0 * n**3 / ms * 1e-9
        k = a @ a.t() + n * torch.eye(n, dtype=torch.float64, device=dev)
        torch.linalg.cholesky(k)
        ms = ev_time(lambda: torch.linalg.cholesky(k), 3)
        out[f"cusolver_potrf_{n}_ms"] = ms
        out[f"cusolver_potrf_{n}_tflops"] = n**3 / 3.0 / ms * 1e-9
        L = torch.linalg.cholesky(k)
        ms = ev_time(lambda: torch.cholesky_inverse(L), 3)
        out[f"cusolver_potri_{n}_ms"] = ms
        del a, b, c, k, L
    x = torch.empty(1 << 29, dtype=torch.float64, device=dev); y = torch.empty_like(x)
    for _ in range(3): y.copy_(x)
    ms = ev_time(lambda: y.copy_(x), 10)
    out["hbm_copy_gbs"] = 2 * x.numel() * 8 / ms * 1e-6
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()

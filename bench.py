#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native dense-GP path (BASELINE.json metric):
"MLE objective+gradient evals/sec at n=8192 D=8 (FP64); Cholesky TFLOP/s vs peak".

One "step" = one evaluation of GpPredictor.logLikelihoodWithDerivatives (gp/regression/GpPredictor.scala:60-80)
with nParams = D+2 = 10 on the synthetic C2 workload of SURVEY.md 8(d) (seed 2, X ~ U(0,1)^{8192x8}).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* value  : evals/s with X, y resident in HBM (gpk_gp_nll_grad_dev), CUDA-event timed, max over ranks.
* e2e    : same metric through the public host API (GpPredictor.logLikelihoodWithDerivatives) with pinned
           HOST buffers: H2D of X,y and D2H of (ll, g) inside the timed region.
* N > 1  : MLE restarts are independent evaluations: each rank evaluates its own restart
           (different theta, same data), no data-path collective; value = N evals / max-over-ranks time.
* roofline: the Cholesky trailing update (SYRK on DMMA) timed alone against a live cuBLAS Dgemm peak.
* cpu_baseline / --impl reference: the oracle's LAPACK-backed port on the box's host cores (the Scala
  reference needs a JVM, absent here -- see DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAIN, DIM = 8192, 8
NPARAMS = DIM + 2
METRIC = "MLE objective+gradient evals/sec at n=8192 D=8 (FP64)"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpk", choices=["gpk", "reference"])
    ap.add_argument("--n", type=int, default=N_TRAIN, help="(debug only) problem size; the reported config is n=8192")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def config(n, n_gpus):
    return {"workload": f"C2: GP regression MLE objective+gradient, ARD-SE kernel, n={n}, D={DIM}, P={NPARAMS}, FP64, "
                        f"synthetic seed 2 (SURVEY.md 8(d))",
            "n": n, "D": DIM, "nparams": NPARAMS,
            "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} independent MLE restarts, one per GPU, no collective",
            "l2": "working set (K, L^-1: 2 x 512 MiB) exceeds the 126 MB L2; no explicit flush",
            "launches": ("eager launches (GPK_GRAPH=0)" if os.environ.get("GPK_GRAPH", "1") == "0" else
                         "CUDA-graph replay: per step one parameter kernel + one graph launch; gpu_launches counts the graph's "
                         "kernel nodes (398 per evaluation), captured during warm-up")}


# ------------------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ------------------------------------------------------------------------------------------------------
def cpu_eval_time(n, reps=1):
    from oracle import gp_oracle as orc
    X, y, theta = orc.make_c2(n=n, D=DIM)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.fast_loglik_with_derivs(X, y, theta, None, NPARAMS)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_literal_time(n=768):
    from oracle import gp_oracle as orc
    X, y, theta = orc.make_c2(n=n, D=DIM)
    t0 = time.perf_counter()
    orc.lit_loglik_with_derivs(X, y, theta, None, NPARAMS)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    n = args.n
    for _ in range(min(args.warmup, 1)):  # one warm-up eval is enough to page in BLAS (each takes seconds)
        cpu_eval_time(n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_eval_time(n)
    dt = (time.perf_counter() - t0) / args.steps
    v = 1.0 / dt
    sample = (f"each step = one full n={n} evaluation through the oracle's LAPACK-backed port (OpenBLAS dpotrf+dpotri, "
              f"fused O(n^2 P) gradient, {cores} threads); the Scala reference as written does ~23x more flops single-threaded")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config(n, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_gpk(args):
    import numpy as np
    import torch
    import ctypes as C
    import gp_algos_b200 as gp
    from gp_algos_b200 import _lib, synthetic   # oracle/ is imported by the cpu_baseline / --impl reference legs only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libgpk has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    X, y, theta0 = synthetic.make_c2(n=n, D=DIM)
    # each rank = one MLE restart: same data, its own hyper-parameter point
    rng = np.random.default_rng(100 + rank)
    theta = theta0.copy()
    if rank > 0:
        theta[0] *= 10 ** rng.uniform(-0.1, 0.1)
        theta[1:-1] *= 10 ** rng.uniform(-0.1, 0.1, size=DIM)
    # a real (non-default) torch stream shared with libgpk so torch.cuda.Event brackets the library's launches
    tstream = torch.cuda.Stream(priority=-1)  # above libgpk's low-priority side streams
    torch.cuda.set_stream(tstream)
    h = _lib.Handle(local, tstream.cuda_stream)
    dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()  # (D, n) row-major == n x D column-major, ld n
    dy = torch.from_numpy(y).cuda()
    dout = torch.zeros(NPARAMS + 1, dtype=torch.float64, device="cuda")
    dinfo = torch.zeros(1, dtype=torch.int32, device="cuda")
    th_c = np.ascontiguousarray(theta)

    def step_dev():
        h.check(h.lib.gpk_gp_nll_grad_dev(h.h, dX.data_ptr(), n, DIM, n, dy.data_ptr(), _lib.ptr(th_c), 0, 0.0, NPARAMS,
                                          dout.data_ptr(), dinfo.data_ptr()))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launch_count() - l0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    assert int(dinfo.item()) == 0, "factorisation failed"
    res_dev = dout.cpu().numpy().copy()

    # ---- e2e: public host API, pinned host buffers, H2D + D2H inside the timed region -----------------
    Xp = torch.from_numpy(np.asfortranarray(X).T.copy()).pin_memory()
    yp = torch.from_numpy(y.copy()).pin_memory()
    Xh = Xp.numpy().T  # n x D column-major view over pinned memory
    yh = yp.numpy()
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(theta[0], theta[1:-1], theta[-1]))
    pred = gp.GpPredictor(kf, handle=h)
    inp = gp.PredictionTrainingInput(Xh, None, yh)
    for _ in range(2):
        ll, g = pred.logLikelihoodWithDerivatives(inp, theta, NPARAMS)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll, g = pred.logLikelihoodWithDerivatives(inp, theta, NPARAMS)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert abs(ll - res_dev[0]) <= 1e-12 * abs(ll) and np.allclose(g, res_dev[1:], rtol=1e-12, atol=0)

    # ---- roofline: the Cholesky trailing update (SYRK, DMMA) alone; peak = live cuBLAS Dgemm ----------
    roofline = None
    if rank == 0:
        nn, kk = n // 2, n // 2
        P = torch.randn(kk, nn, dtype=torch.float64, device="cuda")  # n x k column-major
        Cm = torch.zeros(nn, nn, dtype=torch.float64, device="cuda")
        for _ in range(2):
            h.check(h.lib.gpk_syrk_lower_dev(h.h, P.data_ptr(), nn, Cm.data_ptr(), nn, nn, kk))
        reps = 5
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s0.record()
        for _ in range(reps):
            h.check(h.lib.gpk_syrk_lower_dev(h.h, P.data_ptr(), nn, Cm.data_ptr(), nn, nn, kk))
        s1.record(); torch.cuda.synchronize()
        syrk_ms = s0.elapsed_time(s1) / reps
        syrk_flops = float(nn) * nn * kk  # algorithmic: lower triangle only, 2 * (n^2/2) * k
        A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda")
        Co = torch.empty_like(A)
        for _ in range(2):
            torch.matmul(A, B, out=Co)
        best = 1e30
        for _ in range(5):
            s0.record(); torch.matmul(A, B, out=Co); s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        peak = 2.0 * n ** 3 / best * 1e-9
        ach = syrk_flops / syrk_ms * 1e-9
        roofline = {"kernel": "gemm_f64_dmma_kernel<false,false,SmallTile 64x64> as the Cholesky trailing update (SYRK lower, "
                              f"n={nn}, k={kk})", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel and shape from the ncu --set full capture in
                    # profiles/r01_ncu_summary_v3.txt (608.4 + 60.3 MB; algorithmic 268 MB), bytes per launch
                    "traffic": 668.71e6 if (nn, kk) == (4096, 4096) else None,
                    "peak_source": f"cuBLAS Dgemm {n}^3 measured live in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                   "DMMA pipe ceiling 37.1 TFLOP/s, profiles/r01_fp64_microbench.txt)",
                    "eval_flops": float(n) ** 3, "eval_tflops": float(n) ** 3 / (ms / args.steps) * 1e-9,
                    "eval_frac_of_peak": float(n) ** 3 / (ms / args.steps) * 1e-9 / peak}
        del A, B, Co, P, Cm

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count()
            t_cpu = cpu_eval_time(n)
            t_lit = cpu_literal_time(768)
            cpu = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"one full n={n} evaluation, oracle LAPACK-backed port (OpenBLAS, {cores} threads): {t_cpu:.2f} s",
                   "literal_port_n768_s": t_lit,
                   "literal_port_extrapolated_n8192_s": t_lit * (8192 / 768) ** 3,
                   "literal_note": "line-by-line C restatement of the Scala path, 1 thread, timed at n=768 and scaled by n^3"}
        out = {
            "metric": METRIC, "value": world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(n, world),
            "e2e": {"value": world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": (n * DIM + n) * 8,
                    "d2h_bytes_per_step": (NPARAMS + 1) * 8 + 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "result": {"ll": float(res_dev[0]), "grad_inf_norm": float(np.abs(res_dev[1:]).max())},
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpk(args)


if __name__ == "__main__":
    main()

"""BASELINE.json config 4: 512 independent GPs of n=1024, D=8 (MLE-restart flavour: objective+gradient; GP-UKF flavour:
fit + 17 sigma-point predictions), sharded over the ranks of one node with NO data-path collective.
  python tools/bench_c4.py                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4.py
Prints one JSON line (rank 0): problems/s for both flavours, device-timed, max over ranks."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_algos_b200 import _lib, batched
from gp_algos_b200 import synthetic

B, N, D, M = int(os.environ.get("C4_B", 512)), 1024, 8, 17
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = batched.shard_bounds(B, rank, world)
probs = [synthetic.make_c4_problem(b) for b in range(lo, hi)]
X = np.stack([p[0] for p in probs]); ys = np.stack([p[1] for p in probs]); Xs = np.stack([p[2] for p in probs]); th = np.stack([p[3] for p in probs])
ts = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(ts)
h = _lib.Handle(local, ts.cuda_stream)
nb = hi - lo
dX = torch.from_numpy(np.ascontiguousarray(np.transpose(X, (0, 2, 1)))).cuda(); dy = torch.from_numpy(ys).cuda()
out = torch.zeros(nb * 11, dtype=torch.float64, device="cuda"); info = torch.zeros(nb, dtype=torch.int32, device="cuda")
thc = np.ascontiguousarray(th)
def step():
    h.check(h.lib.gpk_gp_nll_grad_batched_dev(h.h, nb, dX.data_ptr(), N, D, N, N * D, dy.data_ptr(), _lib.ptr(thc), 0, 0.0, 10, out.data_ptr(), info.data_ptr()))
def barrier():
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
for _ in range(3): step()
barrier()
reps = 5
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
l0 = h.launch_count(); e0.record()
for _ in range(reps): step()
e1.record(); barrier()
ms = e0.elapsed_time(e1) / reps
launches = (h.launch_count() - l0) // reps
assert int(info.abs().sum().item()) == 0
# e2e: host buffers in, results out (both flavours)
batched.log_likelihood_with_derivatives_batched(X, ys, th, handle=h)
t0 = time.perf_counter(); ll, g, _ = batched.log_likelihood_with_derivatives_batched(X, ys, th, handle=h); t_e2e = time.perf_counter() - t0
batched.predict_batched(X, ys, th, Xs, handle=h)
t0 = time.perf_counter(); mean, var, _, _ = batched.predict_batched(X, ys, th, Xs, handle=h); t_pred = time.perf_counter() - t0
vals = torch.tensor([ms * 1e-3, t_e2e, t_pred], dtype=torch.float64, device="cuda")
if dist is not None: dist.all_reduce(vals, op=dist.ReduceOp.MAX)
if rank == 0:
    tm, te, tp = [float(v) for v in vals.tolist()]
    # cross-check: problem 0 through the single-problem entry point (parity with the oracle is tests/test_gpu_batched.py's job)
    import gp_algos_b200 as gp
    llo, go = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0][0], th[0][1:-1], th[0][-1])), h).logLikelihoodWithDerivatives(
        gp.PredictionTrainingInput(X[0], None, ys[0]), th[0], 10)
    print(json.dumps({"config": f"C4: {B} independent GPs, n={N}, D={D}, m={M}; {world} GPU(s), {hi - lo} problems on rank 0, no collective",
                      "nll_grad_problems_per_s": B / tm, "nll_grad_ms_per_batch": tm * 1e3, "nll_grad_eff_tflops": B * float(N) ** 3 / tm * 1e-12,
                      "nll_grad_e2e_problems_per_s": B / te, "fit_predict_e2e_problems_per_s": B / tp, "launches_per_batch": int(launches),
                      "n_gpus": world, "batched_vs_single_ll_rel": abs(ll[0] - llo) / abs(llo)}))
if dist is not None: dist.destroy_process_group()

"""Weak scaling of the headline evaluation (one MLE restart per GPU, no collective on the data path) with eager launches and with
graph replay, in one short run:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/eval_weak.py
Prints one JSON line on rank 0: ms per evaluation (CUDA events, max over ranks) and host enqueue ms per evaluation per mode."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gp_algos_b200 import _lib, synthetic

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, D, P, steps = int(os.environ.get("EVAL_N", 8192)), 8, 10, int(os.environ.get("EVAL_STEPS", 10))
X, y, theta = synthetic.make_c2(n=n, D=D)
theta = np.ascontiguousarray(theta * (1.0 + 0.01 * rank))
dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda(); dy = torch.from_numpy(y).cuda()
out = {}
for mode in (0, 1):
    st = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(st)
    h = _lib.Handle(local, st.cuda_stream); h.set_graph_mode(mode)
    dout = torch.zeros(P + 1, dtype=torch.float64, device="cuda"); dinfo = torch.zeros(1, dtype=torch.int32, device="cuda")
    def ev():
        h.check(h.lib.gpk_gp_nll_grad_dev(h.h, dX.data_ptr(), n, D, n, dy.data_ptr(), _lib.ptr(theta), 0, 0.0, P, dout.data_ptr(), dinfo.data_ptr()))
    for _ in range(3): ev()
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): ev()
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    v = torch.tensor([e0.elapsed_time(e1) / steps, 1e3 * t_host / steps], dtype=torch.float64, device="cuda")
    if dist is not None: dist.all_reduce(v, op=dist.ReduceOp.MAX)
    out["graph" if mode else "eager"] = {"ms_per_eval_max_over_ranks": float(v[0]), "host_enqueue_ms_per_eval_max": float(v[1]),
                                         "evals_per_s_total": world * 1e3 / float(v[0])}
    assert int(dinfo.item()) == 0
    h.close()
if rank == 0:
    print(json.dumps({"config": f"C2 weak scaling, n={n}, one restart per GPU", "n_gpus": world, **out}))
if dist is not None: dist.destroy_process_group()

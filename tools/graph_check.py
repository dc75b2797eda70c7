"""Graph replay vs eager launches of the MLE objective+gradient evaluation (gpk_gp_nll_grad_dev):
  * results bit-identical for a sequence of different hyper-parameter points,
  * device time per evaluation (CUDA events around back-to-back calls) and host time spent enqueueing them.
GRAPH_NS="700,8192" selects the sizes; run under `taskset -c 0` (+ a busy neighbour) to see the host-contended case."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gp_algos_b200 import _lib, synthetic

D, P = 8, 10
steps = int(os.environ.get("GRAPH_STEPS", 10))
for n in [int(v) for v in os.environ.get("GRAPH_NS", "700,8192").split(",")]:
    X, y, theta0 = synthetic.make_c2(n=n, D=D)
    dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()
    dy = torch.from_numpy(y).cuda()
    rng = np.random.default_rng(7)
    thetas = [theta0 * 10 ** rng.uniform(-0.1, 0.1, size=D + 2) for _ in range(5)]
    res = {}
    for mode in [int(v) for v in os.environ.get("GRAPH_MODES", "0,1").split(",")]:
        st = torch.cuda.Stream(priority=-1)
        torch.cuda.set_stream(st)
        h = _lib.Handle(0, st.cuda_stream)
        h.set_graph_mode(bool(mode))
        dout = torch.zeros(P + 1, dtype=torch.float64, device="cuda")
        dinfo = torch.zeros(1, dtype=torch.int32, device="cuda")

        def ev(th):
            th = np.ascontiguousarray(th)
            h.check(h.lib.gpk_gp_nll_grad_dev(h.h, dX.data_ptr(), n, D, n, dy.data_ptr(), _lib.ptr(th), 0, 0.0, P,
                                              dout.data_ptr(), dinfo.data_ptr()))
        outs = []
        for i, th in enumerate(thetas):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev(th)
            torch.cuda.synchronize()
            if i < 3:   # graph mode: eager, capture + instantiate + launch, replay
                print(f"n={n} graph={mode}: call {i + 1} took {1e3 * (time.perf_counter() - t0):8.2f} ms wall", flush=True)
            outs.append(dout.cpu().numpy().copy())
        res[mode] = np.array(outs)
        for _ in range(3):
            ev(theta0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = h.launch_count()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            ev(theta0)
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"n={n} graph={mode}: {e0.elapsed_time(e1) / steps:8.3f} ms/eval on the device, host enqueue "
              f"{1e3 * t_host / steps:7.3f} ms/eval, {(h.launch_count() - l0) // steps} kernels/eval, info={int(dinfo.item())}", flush=True)
        h.close()
    if len(res) < 2:
        continue
    same = np.array_equal(res[0], res[1])
    print(f"n={n}: graph == eager bit for bit over {len(thetas)} hyper-parameter points: {same}; ll[0]={res[1][0][0]:.12g}", flush=True)
    assert same

import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_algos_b200 import _lib
from gp_algos_b200 import synthetic
ts = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(ts)
h = _lib.Handle(0, ts.cuda_stream)
for n in (128, 256, 512, 1024, 2048, 4096, 8192):
    X, y, th = synthetic.make_c2(n=n)
    dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda(); dy = torch.from_numpy(y).cuda()
    out = torch.zeros(11, dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    thc = np.ascontiguousarray(th)
    def step():
        h.check(h.lib.gpk_gp_nll_grad_dev(h.h, dX.data_ptr(), n, 8, n, dy.data_ptr(), _lib.ptr(thc), 0, 0.0, 10, out.data_ptr(), info.data_ptr()))
    for _ in range(3): step()
    torch.cuda.synchronize()
    l0 = h.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): step()
    t_host = (time.perf_counter() - t0) / reps
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"n={n:5d}  {ms*1e3:9.1f} us/eval  host enqueue {t_host*1e6:8.1f} us  launches/eval {(h.launch_count()-l0)/reps:.0f}  eff {n**3/ms*1e-9:.2f} TFLOP/s", flush=True)

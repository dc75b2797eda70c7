import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gp_algos_b200 import _lib, batched, synthetic
B = int(os.environ.get("C4_B", 512)); N, D = 1024, 8
probs = [synthetic.make_c4_problem(b) for b in range(B)]
X = np.stack([p[0] for p in probs]); ys = np.stack([p[1] for p in probs]); th = np.ascontiguousarray(np.stack([p[3] for p in probs]))
ts = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(ts)
h = _lib.Handle(0, ts.cuda_stream)
dX = torch.from_numpy(np.ascontiguousarray(np.transpose(X, (0, 2, 1)))).cuda(); dy = torch.from_numpy(ys).cuda()
out = torch.zeros(B * 11, dtype=torch.float64, device="cuda"); info = torch.zeros(B, dtype=torch.int32, device="cuda")
res = []
for rep in range(int(os.environ.get("REPS", 4))):
    h.check(h.lib.gpk_gp_nll_grad_batched_dev(h.h, B, dX.data_ptr(), N, D, N, N * D, dy.data_ptr(), _lib.ptr(th), 0, 0.0, 10, out.data_ptr(), info.data_ptr()))
    torch.cuda.synchronize()
    res.append(out.cpu().numpy().reshape(B, 11).copy())
nbad = 0
for i, r in enumerate(res[1:]):
    if not np.array_equal(r, res[0]):
        nbad += 1
        d = np.abs(r - res[0]) / np.maximum(np.abs(res[0]), 1e-300)
        print("rep", i + 1, "differs: nan count", int(np.isnan(r).sum()), "max rel", np.nanmax(d), "first rows", np.argwhere(~(r == res[0]))[:4].tolist())
print("dev vs dev:", len(res) - 1, "repetitions,", nbad, "differ from the first; nan in first:", int(np.isnan(res[0]).sum()))
ll, g, inf = batched.log_likelihood_with_derivatives_batched(X, ys, th, handle=h)
e = np.concatenate([ll[:, None], g], axis=1)
d = np.abs(e - res[0]) / np.maximum(np.abs(res[0]), 1e-300)
print("host vs dev: max rel", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "n bad", int((d > 1e-12).sum()), "info", int(np.abs(inf).sum()))
bad = np.argwhere(d > 1e-12)[:8]
for b, p in bad: print("  problem", b, "param", p, "dev", res[0][b, p], "host", e[b, p])

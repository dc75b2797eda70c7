"""gpk_mg_potrf_solve (single process, all GPUs, C ABI) at n = 65536: block-width sweep.  python tools/bench_mg.py [ndev] [nb ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_algos_b200 import synthetic
from gp_algos_b200.multi_gpu import MultiGpuGp
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nbs = [int(v) for v in sys.argv[2:]] or [1024]
n = int(os.environ.get("C5_N", 65536))
X, y, theta = synthetic.make_c2(n=n, D=8, seed=5)
for nb in nbs:
    mg = MultiGpuGp(ndev, nb=nb)
    runs = [mg.fit(X, y, theta) for _ in range(2)]
    best = min(r.seconds for r in runs)
    print(json.dumps({"what": "gpk_mg_potrf_solve", "n": n, "ndev": mg.ndev, "nb": nb, "seconds": [r.seconds for r in runs],
                      "tflops_per_gpu": n ** 3 / 3 / best * 1e-12 / mg.ndev, "ll": runs[-1].logLikelihood, "put_bytes": runs[-1].put_bytes}), flush=True)
    mg.close()

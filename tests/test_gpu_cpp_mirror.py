"""GPU: the C++ host mirror of the reference's Scala interface (include/gpk.hpp) -- a compiled caller's view of the drop-in
boundary.  tests/cpp/host_mirror_test (built by __graft_entry__.build()) runs MatrixUtils / GpPredictor / EpParameterEstimator /
GpClassifier through the mirror on a problem written here; its numbers are compared with the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def test_cpp_mirror_matches_oracle(tmp_path):
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    n, D, m = 180, 3, 11
    X, y, th = orc.make_c2(n=n, D=D, seed=77)
    th = th.copy(); th[-1] = 0.2
    Xs = X[:m] * 0.9 + 0.03
    rng = np.random.default_rng(5)
    t = np.where(X @ rng.standard_normal(D) - 0.8 + 0.2 * rng.standard_normal(n) >= 0, 1, -1).astype(np.int32)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.json"
    with open(fin, "wb") as f:
        np.array([n, D, m, D + 2], dtype=np.int32).tofile(f)
        np.asfortranarray(X).T.copy().tofile(f)          # column-major payload
        np.asfortranarray(Xs).T.copy().tofile(f)
        y.tofile(f); th.tofile(f); t.tofile(f)
    r = subprocess.run([EXE, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    o = json.loads(fout.read_text())
    K = orc.fast_build_kernel_matrix(X, th)
    L, alpha = orc.fast_precompute(X, y, th)
    rel = lambda a, b, tol=1e-9: np.all(np.abs(np.asarray(a) - np.asarray(b)) <= tol * np.maximum(np.abs(b), np.abs(np.asarray(b)).max() * 1e-6))
    assert rel(o["K_diag"], np.diag(K), 1e-14)
    assert rel(o["alpha_via_solves"], alpha)
    assert rel(o["Li_last_row"], np.linalg.inv(L)[-1])
    llo, go = orc.fast_loglik_with_derivs(X, y, th)
    assert abs(o["ll"][0] - llo) <= 1e-9 * abs(llo) and rel(o["grad"], go)
    mo, So, ll2 = orc.fast_predict(X, y, Xs, th, 0.05)
    assert rel(o["pred_mean"], mo) and rel(o["pred_var"], np.diag(So)) and abs(o["pred_ll"][0] - ll2) <= 1e-9 * abs(ll2)
    mp, Sp, V = orc.fast_compute_posterior(X, Xs, L, alpha, th)
    assert rel(o["post_mean"], mp) and rel(o["V_col0"], V[:, 0])
    e = orc.fast_ep_estimate(K, t, fixed_sweeps=3)
    assert rel(o["ep_tau"], e["tau"]) and rel(o["ep_nu"], e["nu"]) and abs(o["ep_logZ"][0] - e["logZ"]) <= 1e-9 * abs(e["logZ"])
    Ks = orc.fast_build_kernel_matrix(Xs, th, X); Kss = orc.fast_build_kernel_matrix(Xs, th)
    po, _, _ = orc.fast_ep_classify(K, Ks, Kss, e["tau"], e["nu"], e["L"])
    assert rel(o["ep_prob"], po)
    assert int(o["errors_caught"][0]) == 15      # IllegalArgument, NotConverged(minor 4), MatchError, MatrixNotSymmetric

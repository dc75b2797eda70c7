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
    # IllegalArgument, NotConverged(minor 4), MatchError, MatrixNotSymmetric, optimizeNoise=false defect (SE: require, Co2: index), Co2 on 3-D data
    assert int(o["errors_caught"][0]) == 127
    # obtainOptimalHyperParams: the reported optimum is a point of the oracle's objective and not worse than the start
    th_opt = np.array(o["opt_theta"])
    assert abs(o["opt_ll"][0] - orc.fast_loglik_with_derivs(X, y, th_opt, None, 0)[0]) <= 1e-9 * abs(o["opt_ll"][0])
    assert o["opt_ll"][0] >= llo
    assert o["ll_again"][0] == o["ll"][0]
    # resident model after three appends == refit on all rows
    assert int(o["model_size"][0]) == n
    assert rel(o["model_alpha"], alpha) and abs(o["model_ll"][0] - llo) <= 1e-9 * abs(llo)
    assert rel(o["model_mean"], mp)
    Ll, al = orc.lit_precompute(X, y, th)
    ug = np.array(o["model_ucb_grad"]).reshape(D, m).T
    for i in (0, m - 1):
        u_o, g_o, _, _ = orc.lit_ucb_with_grad(X, Ll, al, th, Xs[i], 1.5)
        assert abs(o["model_ucb"][i] - u_o) <= 1e-9 * max(abs(u_o), 1e-3) and rel(ug[i], g_o)
    # Co2Kernel through the same template
    T, Ts, yc = 1958.0 + 30.0 * X[:, :1], 1958.0 + 30.0 * Xs[:, :1], 330.0 + 10.0 * y
    chp = np.array([60., 70., 8., 50., 2., 0.34, 2.4, 0.88, 0.26, 0.2, 1.5])
    with orc.co2_kernel():
        Kc = orc.lit_build_kernel_matrix(T, chp)
        tol = 1e-9 * max(1.0, np.linalg.cond(Kc) / 1e5)
        cll, cg = orc.lit_loglik_with_derivs(T, yc, chp, None)
        cm, _, _ = orc.lit_predict(T, yc, Ts, chp, None)
    assert abs(o["co2_ll"][0] - cll) <= tol * abs(cll) and rel(o["co2_grad"], cg, tol)
    assert np.all(np.abs(np.array(o["co2_mean"]) - cm) <= tol * np.abs(cm).max())
    assert np.abs(np.array(o["co2_K_row0"]) - Kc[0]).max() <= 8 * np.finfo(float).eps * np.abs(Kc).max()
    assert abs(o["co2_apply"][0] - orc.co2_k(T[0, 0], T[1, 0], chp, False)) <= 4e-16 * abs(o["co2_apply"][0])
    assert abs(o["co2_apply"][1] - orc.co2_k(T[0, 0], T[0, 0], chp, True)) <= 4e-16 * abs(o["co2_apply"][1])

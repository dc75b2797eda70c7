"""GPU: the large-GP solver (gp_algos_b200/distributed.py, BASELINE.json config 5) through libgpk's device-level blocks,
against the oracle.  1 x 1 grid in-process; 2 GPUs (when the box has them) through torchrun + NCCL."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,nb,sigma", [(1000, 256, None), (1536, 512, 0.05), (640, 128, None)])
def test_single_gpu_block_cholesky_matches_oracle(n, nb, sigma):
    from oracle import gp_oracle as orc
    from gp_algos_b200.distributed import DistributedGp
    X, y, theta = orc.make_c2(n=n, D=8, seed=21)
    solver = DistributedGp(nb=nb, device=0)
    fit = solver.fit(X, y, theta, sigmaNoise=sigma)
    L_o, alpha_o = orc.fast_precompute(X, y, theta, sigma)
    ll_o = orc.fast_loglik(alpha_o, L_o, y)
    assert abs(fit.logLikelihood - ll_o) <= 1e-9 * abs(ll_o)
    assert np.allclose(fit.alphaVec, alpha_o, rtol=1e-9, atol=1e-9 * np.abs(alpha_o).max())
    L = solver.gather_factor()
    K = orc.fast_build_kernel_matrix(X, theta)
    if sigma is not None:
        K[np.diag_indices_from(K)] += sigma
    assert np.linalg.norm(K - L @ L.T) <= 8 * n * np.finfo(float).eps * np.linalg.norm(K)      # backward-error bound
    assert np.linalg.norm(L - L_o) <= 1e-9 * np.linalg.norm(L_o)
    if sigma is None:
        assert solver.residual(y, fit.alphaVec) < 1e-10
    Xs = X[:13] * 0.95 + 0.02
    mean_o = orc.fast_build_kernel_matrix(Xs, theta, X) @ alpha_o                 # the cross-covariance never carries noise
    assert np.allclose(solver.predict_mean(Xs, fit.alphaVec), mean_o, rtol=1e-9, atol=1e-9 * np.abs(mean_o).max())
    if sigma is None:                                                                # computePosterior carries no sigmaNoise
        pm, ps = solver.predict(Xs, fit.alphaVec)
        _, S_o, _ = orc.fast_compute_posterior(X, Xs, L_o, alpha_o, theta)
        assert np.allclose(pm, mean_o, rtol=1e-9, atol=1e-9 * np.abs(mean_o).max())
        assert np.allclose(ps, S_o, rtol=1e-8, atol=1e-9 * np.abs(S_o).max())


def test_not_positive_definite_is_reported_with_its_minor():
    from oracle import gp_oracle as orc
    import gp_algos_b200 as gp
    from gp_algos_b200.distributed import DistributedGp
    X, y, theta = orc.make_c2(n=512, D=8, seed=22)
    import scipy.linalg.lapack as lp
    K = orc.fast_build_kernel_matrix(X, theta)
    K[np.diag_indices_from(K)] += -0.5            # Option sigmaNoise is added un-squared (GpPredictor.scala:116): K - 0.5 I is indefinite
    _, info = lp.dpotrf(K, lower=1)
    assert info > 0
    solver = DistributedGp(nb=128, device=0)
    with pytest.raises(gp.NotPositiveDefiniteError) as e:
        solver.fit(X, y, theta, sigmaNoise=-0.5)
    assert e.value.minor == info


def test_two_gpus_nccl_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = {}
    for nproc, grid in ((1, "1x1"), (2, "2x1"), (2, "1x2")):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", "29731", os.path.join(ROOT, "tools", "bench_c5.py"), "--size", "5000", "--nb", "256", "--grid", grid,
               "--reps", "1", "--predict", "21"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[grid] = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert outs[grid]["residual_Kalpha_minus_y_over_y"] < 1e-10
        assert max(outs[grid]["predict_rel_diff_vs_single_handle"]) < 1e-8      # mean and full covariance over NCCL
    for grid in ("2x1", "1x2"):
        assert abs(outs[grid]["ll"] - outs["1x1"]["ll"]) <= 1e-11 * abs(outs["1x1"]["ll"])


def test_single_gpu_medium_size_residual_and_loglik_against_resident_model():
    """n = 12288 (12 block columns of 1024): the block-cyclic solver and the single-handle path (gpk_gp_model_fit) must
    agree on alpha and the log-likelihood, and K alpha = y must hold -- no oracle needed at this size."""
    import gp_algos_b200 as gp
    from gp_algos_b200.distributed import DistributedGp
    X, y, theta = orc_make(12288)
    solver = DistributedGp(nb=1024, device=0)
    fit = solver.fit(X, y, theta)
    assert solver.residual(y, fit.alphaVec) < 1e-10
    pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(theta[0], theta[1:-1], theta[-1])))
    model = pred.fit(X, None, y, theta)
    assert abs(model.logLikelihood - fit.logLikelihood) <= 1e-11 * abs(fit.logLikelihood)
    assert np.allclose(model.alphaVec, fit.alphaVec, rtol=1e-7, atol=1e-9 * np.abs(fit.alphaVec).max())
    model.close()


def orc_make(n):
    from oracle import gp_oracle as orc
    return orc.make_c2(n=n, D=8, seed=5)

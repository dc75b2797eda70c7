"""GPU: the multi-GPU C-ABI entry points (gpk_mg_create / gpk_mg_potrf_solve, csrc/gpk_mg.cu) against the oracle.  On a
1-GPU box the whole schedule runs with one device (every path but the peer puts); with more devices the same call must give the
same answer on 1, 2, ... devices (the 2+ device cases skip themselves on a 1-GPU box; bench.py repeats the comparison in-run on
the multi-GPU boxes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nb,sigma", [(1000, 256, None), (1536, 512, 0.05), (640, 128, None), (300, 1024, None)])
def test_single_device_matches_oracle(n, nb, sigma):
    from oracle import gp_oracle as orc
    from gp_algos_b200.multi_gpu import MultiGpuGp
    X, y, theta = orc.make_c2(n=n, D=8, seed=21)
    mg = MultiGpuGp(1, nb=nb)
    fit = mg.fit(X, y, theta, sigmaNoise=sigma)
    L_o, alpha_o = orc.fast_precompute(X, y, theta, sigma)
    ll_o = orc.fast_loglik(alpha_o, L_o, y)
    assert abs(fit.logLikelihood - ll_o) <= 1e-9 * abs(ll_o)
    assert np.allclose(fit.alphaVec, alpha_o, rtol=1e-9, atol=1e-9 * np.abs(alpha_o).max())
    K = orc.fast_build_kernel_matrix(X, theta)
    if sigma is not None:
        K[np.diag_indices_from(K)] += sigma
    assert np.linalg.norm(K @ fit.alphaVec - y) <= 1e-10 * np.linalg.norm(y) * max(1.0, np.linalg.cond(K) / 1e5)
    assert fit.put_bytes == 0 and fit.seconds > 0
    fit2 = mg.fit(X, y, theta, sigmaNoise=sigma)                      # the workspace is reused: identical results
    assert fit2.logLikelihood == fit.logLikelihood and np.array_equal(fit2.alphaVec, fit.alphaVec)
    mg.close()


def test_not_positive_definite_reports_the_failing_minor():
    import scipy.linalg.lapack as lp
    from oracle import gp_oracle as orc
    import gp_algos_b200 as gp
    from gp_algos_b200.multi_gpu import MultiGpuGp
    X, y, theta = orc.make_c2(n=512, D=8, seed=22)
    K = orc.fast_build_kernel_matrix(X, theta)
    K[np.diag_indices_from(K)] += -0.5
    _, info = lp.dpotrf(K, lower=1)
    assert info > 0
    mg = MultiGpuGp(1, nb=128)
    with pytest.raises(gp.NotPositiveDefiniteError) as e:
        mg.fit(X, y, theta, sigmaNoise=-0.5)
    assert e.value.minor == info
    fit = mg.fit(X, y, theta)                                           # the handle stays usable
    assert np.isfinite(fit.logLikelihood)
    mg.close()


def test_bad_arguments():
    import gp_algos_b200 as gp
    from gp_algos_b200 import _lib
    from gp_algos_b200.multi_gpu import MultiGpuGp
    with pytest.raises(_lib.GpkError):
        MultiGpuGp(64)                                                  # more devices than the box has
    with pytest.raises(ValueError):
        MultiGpuGp(1, nb=100)
    mg = MultiGpuGp(1)
    with pytest.raises(ValueError):
        mg.fit(np.zeros((10, 3)), np.zeros(9), np.ones(5))
    mg.close()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_several_devices_agree_with_one(ndev):
    import torch
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    from oracle import gp_oracle as orc
    from gp_algos_b200.multi_gpu import MultiGpuGp
    X, y, theta = orc.make_c2(n=5000, D=8, seed=5)
    one = MultiGpuGp(1, nb=256); f1 = one.fit(X, y, theta); one.close()
    mg = MultiGpuGp(ndev, nb=256); fn = mg.fit(X, y, theta); mg.close()
    assert abs(fn.logLikelihood - f1.logLikelihood) <= 1e-11 * abs(f1.logLikelihood)
    assert np.allclose(fn.alphaVec, f1.alphaVec, rtol=1e-9, atol=1e-9 * np.abs(f1.alphaVec).max())
    assert fn.put_bytes > 0

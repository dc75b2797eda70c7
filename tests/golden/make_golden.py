"""Generates the committed golden fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted; the GPU box never reads /root/reference).

  python tests/golden/make_golden.py

* mu_3x3.npz        -- the 3x3 cases of the reference's src/test/scala/utils/MatrixUtilsTest.scala:24-114 with the
                       oracle's outputs (and numpy.linalg.solve/inv answers the Scala test compares against).
* c1_small.npz, c2_small.npz, c3_small.npz, c3_grad_small.npz, c4_small.npz
                    -- reduced-size instances of BASELINE.json configs 1-4 (SURVEY.md 8(d) generators and seeds)
                       evaluated by the LITERAL oracle (oracle/gp_oracle.c).
* co2_maunaloa.npz  -- the Co2Kernel path (gp/regression/Co2Prediction.scala) on the reference's src/main/resources/co2/maunaLoa.txt:
                       literal-oracle outputs at the shipped hyper-parameters + the reference's shipped co2/co2PredResults.txt
                       as a soft fixture (see co2_maunaloa below).
* boston_soft.npz   -- soft fixture: the reference's own shipped resources src/main/resources/boston.csv and
                       src/main/resources/boston/bostonPredResults.txt (506 rows: idx, mean, sqrt(var), target; written by
                       gp/tasks/MasterThesisRelatedTasks.scala:105-122) with the hyper-parameters of
                       utils/TestingUtils.scala:29-34.  The file holds results AFTER a further L-BFGS run, so it pins the
                       oracle only to ~1e-3, not to 1e-9.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gp_oracle as orc  # noqa: E402

REF = "/root/reference/src/main/resources"


def mu_3x3():
    lower = np.array([[0.3, 0., 0.], [0.2, 0.3, 0.], [0.1, 0.99, 0.11]])
    upper = np.array([[0.4, 0.1, 0.9], [0., 0.2, 0.89], [0., 0., .5]])
    inp = np.array([[2.4, 1.3, 1.9], [2.1, 0.99, 3.1], [1.89, 2.01, 4.]])
    rhs_l, rhs_u = np.array([3., 2., 1.]), np.array([7., 3., 4.])
    rhs_m = np.array([[0.4, 0.9], [0.8, 0.3], [0.7, 0.4]])
    th = orc.pack_theta(1., [1., 1., 1.], 0.)
    K = orc.lit_build_kernel_matrix(inp, th)
    L = orc.lit_cholesky(K)
    np.savez(os.path.join(HERE, "mu_3x3.npz"), lower=lower, upper=upper, input=inp, rhs_l=rhs_l, rhs_u=rhs_u, rhs_m=rhs_m,
             theta=th, fwd_vec=orc.lit_forward_solve(lower, rhs_l), back_vec=orc.lit_back_solve(upper, rhs_u),
             fwd_mat=orc.lit_forward_solve(lower, rhs_m), back_mat=orc.lit_back_solve(upper, rhs_m),
             np_fwd_vec=np.linalg.solve(lower, rhs_l), np_back_vec=np.linalg.solve(upper, rhs_u),
             K=K, L=L, Linv=orc.lit_inv_triangular(L), Kinv_np=np.linalg.inv(K))


def c1_small(n=200, m=50):
    X, y, Xs, th = orc.make_c1(n=n, m=m, seed=1)
    mean, sigma, ll = orc.lit_predict(X, y, Xs, th, None)
    mean_s, sigma_s, ll_s = orc.lit_predict(X, y, Xs, th, 0.05)
    np.savez(os.path.join(HERE, "c1_small.npz"), X=X, y=y, Xs=Xs, theta=th, mean=mean, sigma=sigma, ll=ll,
             mean_s=mean_s, sigma_s=sigma_s, ll_s=ll_s, sigma_noise=0.05)


def c2_small(n=320, D=8):
    X, y, th = orc.make_c2(n=n, D=D, seed=2)
    ll, g = orc.lit_loglik_with_derivs(X, y, th, None)
    ll_s, g_s = orc.lit_loglik_with_derivs(X, y, th, 0.02)
    L, alpha = orc.lit_precompute(X, y, th, None)
    np.savez(os.path.join(HERE, "c2_small.npz"), X=X, y=y, theta=th, ll=ll, grad=g, ll_s=ll_s, grad_s=g_s, sigma_noise=0.02,
             alpha=alpha, L_diag=np.diag(L).copy(), L_last_row=L[-1].copy())


def c3_small(n=160, D=4, m=23):
    X, t, th = orc.make_c3(n=n, D=D, seed=3)
    K = orc.lit_build_kernel_matrix(X, th)
    fixed = orc.lit_ep_estimate(K, t, fixed_sweeps=5)
    ship = orc.lit_ep_estimate(K, t, eps=0.01)
    rng = np.random.default_rng(33)
    Xs = rng.standard_normal((m, D))
    Ks = orc.lit_build_kernel_matrix(Xs, th, X)
    Kss = orc.lit_build_kernel_matrix(Xs, th)
    p, fm, fv = orc.lit_ep_classify(K, Ks, Kss, ship["tau"], ship["nu"], ship["L"])
    np.savez(os.path.join(HERE, "c3_small.npz"), X=X, targets=t, theta=th, Xs=Xs,
             tau5=fixed["tau"], nu5=fixed["nu"], logZ5=fixed["logZ"], L5_diag=np.diag(fixed["L"]).copy(),
             tau=ship["tau"], nu=ship["nu"], logZ=ship["logZ"], sweeps=ship["sweeps"], prob=p, fmean=fm, fvar=fv)


def c3_grad_small(n=140, D=3):
    """MarginalLikelihoodEvaluator.logLikelihood (MLE2:33-66) by the literal oracle, as compiled and with the dropped term."""
    X, t, th = orc.make_c3(n=n, D=D, seed=31)
    th = th.copy(); th[-1] = 0.2          # non-zero noise so that the d/d(noiseVar) component is exercised
    lz, g, o = orc.lit_ep_loglik_with_derivs(X, t, th, eps=0.01)
    K = orc.lit_build_kernel_matrix(X, th)
    g_noquirk = orc.lit_ep_loglik_derivs(X, th, K, o["tau"], o["nu"], o["L"], keep_quirk=False)
    np.savez(os.path.join(HERE, "c3_grad_small.npz"), X=X, targets=t, theta=th, logZ=lz, grad=g, grad_noquirk=g_noquirk,
             sweeps=o["sweeps"], tau=o["tau"], nu=o["nu"])


def c4_small(B=6, n=192, D=8, m=17):
    out = {}
    for b in range(B):
        X, y, Xs, th = orc.make_c4_problem(b, n=n, D=D, m=m)
        ll, g = orc.lit_loglik_with_derivs(X, y, th, None)
        L, alpha = orc.lit_precompute(X, y, th, None)
        mean, sigma, _ = orc.lit_compute_posterior(X, Xs, L, alpha, th)
        out.update({f"X{b}": X, f"y{b}": y, f"Xs{b}": Xs, f"theta{b}": th, f"ll{b}": ll, f"grad{b}": g,
                    f"mean{b}": mean, f"var{b}": np.diag(sigma).copy()})
    np.savez(os.path.join(HERE, "c4_small.npz"), B=B, **out)


def c3_full(n=4096, D=4, sweeps=3):
    """BASELINE.json config 3 at FULL size (n = 4096, D = 4, seed 3): three fixed EP sweeps by the LAPACK-backed
    oracle (fast_ep_estimate: the reference's site loop with full rank-1 downdates, ~2.7 TB of memory traffic per sweep -- about
    a quarter of an hour on the build container's cores, which is why it is generated once and committed: tau, nu, mu, cavity
    parameters, diag L and logZ, 230 KB).  Run with `python tests/golden/make_golden.py c3_full`."""
    X, t, th = orc.make_c3(n=n, D=D, seed=3)
    K = orc.fast_build_kernel_matrix(X, th)
    o = orc.fast_ep_estimate(K, t, fixed_sweeps=sweeps)
    o2 = dict(o)
    np.savez_compressed(os.path.join(HERE, "c3_full.npz"), n=n, D=D, sweeps=sweeps, theta=th, tau=o["tau"], nu=o["nu"], mu=o["mu"],
                        cav_tau=o["cav_tau"], cav_nu=o["cav_nu"], diagL=np.diag(o["L"]).copy(), logZ=o["logZ"],
                        x_checksum=float(X.sum()), t_checksum=int(t.sum()))


def boston_soft():
    data = np.loadtxt(os.path.join(REF, "boston.csv"))
    res = np.loadtxt(os.path.join(REF, "boston", "bostonPredResults.txt"))
    X, y = data[:, :13], data[:, 13]
    # utils/TestingUtils.scala:29-34 optimalBostonHp
    theta = orc.pack_theta(-1130.9947925594922,
                           [566.7442989546967, 735.3624053303566, 536.1791714384265, 610.4651246027757, 626.0353185058663,
                            5.528303239800252, 2853.7974583131668, 1006.5144425910395, 504.78702267976087, 1287.102910849582,
                            387.26286609421436, 2678.0145551405353, 1093.4657445540006], 2.1355086421066893)
    ntrain = 354  # first 70 % of the rows (MasterThesisRelatedTasks.scala:108, divRatio 0.7)
    mean, sigma, ll = orc.fast_predict(X[:ntrain], y[:ntrain], X, theta, None)
    std = np.sqrt(np.diag(sigma))
    print("boston: ll", ll, "max|mean - ref|", np.abs(mean - res[:, 1]).max(), "max|std - ref|", np.abs(std - res[:, 2]).max(),
          "train rows rel", np.abs(mean[:ntrain] - res[:ntrain, 1]).max() / np.abs(res[:ntrain, 1]).max())
    np.savez(os.path.join(HERE, "boston_soft.npz"), X=X, y=y, theta=theta, ntrain=ntrain, ref_mean=res[:, 1], ref_std=res[:, 2],
             ref_target=res[:, 3], oracle_mean=mean, oracle_std=std, oracle_ll=ll)


def co2_maunaloa():
    """The reference's Co2Kernel path on its own Mauna Loa data (gp/tasks/MasterThesisRelatedTasks.scala:59-71
    evaluateGpPredictionOnCo2Ds: first 70 % of the monthly series as training set, posterior over the whole series).
    Hard part: literal-oracle outputs at the SHIPPED hyper-parameters (utils/TestingUtils.scala:17-20).  Soft part: the
    reference's shipped output co2/co2PredResults.txt (year, mean, sqrt(var)), written AFTER its own 20-iteration L-BFGS run
    whose end point was only printed -- it pins the time axis exactly and the training-range posterior to ~0.15 ppm."""
    raw = np.loadtxt(os.path.join(REF, "co2", "maunaLoa.txt"))
    res = np.loadtxt(os.path.join(REF, "co2", "co2PredResults.txt"))
    train, test = orc.co2_data_to_year_with_value(raw, 0.7)
    whole = np.vstack([train, test])
    hp = orc.CO2_SHIPPED_HP.copy()
    X, y, Xs = train[:, :1], train[:, 1], whole[:, :1]
    with orc.co2_kernel():
        mean, sigma, ll = orc.lit_predict(X, y, Xs, hp, None)
        ll2, grad = orc.lit_loglik_with_derivs(X, y, hp, None)
        ll_s, grad_s = orc.lit_loglik_with_derivs(X, y, hp, 0.05, 7)
        K = orc.lit_build_kernel_matrix(X, hp)
        L, alpha = orc.lit_precompute(X, y, hp, None)
    std = np.sqrt(np.diag(sigma))
    nt = train.shape[0]
    print("co2: n", nt, "ll", ll, "cond(K)", np.linalg.cond(K), "train-range max|mean - shipped|", np.abs(mean[:nt] - res[:nt, 1]).max(),
          "max|std - shipped|", np.abs(std[:nt] - res[:nt, 2]).max(), "time axis", np.abs(whole[:, 0] - res[:, 0]).max())
    assert ll == ll2
    np.savez(os.path.join(HERE, "co2_maunaloa.npz"), raw_head=raw[:3], train=train, test=test, theta=hp, ll=ll, grad=grad,
             sigma_noise=0.05, ll_s=ll_s, grad_s=grad_s, mean=mean, var=np.diag(sigma).copy(), alpha=alpha, L_diag=np.diag(L).copy(),
             cond=np.linalg.cond(K), ref_year=res[:, 0], ref_mean=res[:, 1], ref_std=res[:, 2])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c3_full":
        c3_full()
        sys.exit(0)
    mu_3x3(); c1_small(); c2_small(); c3_small(); c3_grad_small(); c4_small()
    if os.path.isdir(REF):
        boston_soft()
        co2_maunaloa()
    print("golden fixtures written to", HERE)

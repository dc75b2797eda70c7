"""CPU: the C-ABI library loads and exports every symbol include/gpk.h declares; the host mirror of
KernelRequisites behaves like the reference's KernelRequisitesTest; the product never touches the oracle and
fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import gp_algos_b200 as gp
from gp_algos_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.lib_path()):
        import __graft_entry__ as ge
        ge.build()
    hdr = open(os.path.join(ROOT, "include", "gpk.h")).read()
    names = sorted(set(re.findall(r"\b(gpk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.lib_path())
    for nme in names:
        assert hasattr(lib, nme), f"{nme} declared in include/gpk.h but not exported by libgpk.so"
    assert b"sm_100a" in ctypes.c_char_p(ctypes.cast(lib.gpk_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value


def test_library_contains_sm100a_dmma_code():
    out = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert "DMMA.8x8x4" in out.stdout      # FP64 tensor path of the trailing update
    assert "LDGSTS" in out.stdout          # cp.async operand staging


def test_no_cpu_fallback_without_gpu():
    if _have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(gp.GpkError):
        _lib.Handle(0)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(1., np.ones(3), 0.))
    with pytest.raises(gp.GpkError):
        gp.MatrixUtils.buildKernelMatrix(kf, np.zeros((3, 3)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gp_algos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "gp_oracle" not in src, f


# ---- mirrors src/test/scala/utils/KernelRequisitesTest.scala:18-49 -------------------------------------
def test_hyperparams_to_dense_vector():
    hp = gp.GaussianRbfParams(signalVar=1., lengthScales=np.ones(5), noiseVar=0.)
    assert np.array_equal(hp.toDenseVector, np.array([1., 1., 1., 1., 1., 1., 0.]))


def test_hyperparams_get_at_position():
    hp = gp.GaussianRbfParams(signalVar=1., lengthScales=np.array([5., 2., 3.]), noiseVar=0.)
    assert [hp.getAtPosition(i) for i in range(1, 6)] == [1., 5., 2., 3., 0.]
    with pytest.raises(LookupError):  # intercept[MatchError]
        hp.getAtPosition(6)


def test_kernel_updates_its_params():
    ls = np.ones(5)
    k0 = gp.GaussianRbfKernel(gp.GaussianRbfParams(1., ls, 0.))
    assert k0.rbfParams == gp.GaussianRbfParams(1., ls, 0.)
    k1 = k0.changeHyperParams(np.array([2., 3., 3., 3., 3., 3., 0.]))
    assert k1.rbfParams == gp.GaussianRbfParams(2., ls * 3., 0.)
    with pytest.raises(ValueError):  # require(...) KernelRequisites.scala:55
        k0.changeHyperParams(np.ones(6))


def test_scalar_kernel_matches_oracle():
    from oracle import gp_oracle as orc
    rng = np.random.default_rng(0)
    th = orc.pack_theta(1.3, [0.5, 0.9, 2.0], 0.2)
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(th[0], th[1:-1], th[-1]))
    X = rng.standard_normal((4, 3))
    K = orc.lit_build_kernel_matrix(X, th)
    for i in range(4):
        for j in range(4):
            assert abs(kf(X[i], X[j], i == j) - K[i, j]) <= 1e-14 * abs(K[i, j])
    for p in range(1, 6):
        dK = orc.lit_build_der_matrix(p, X, th)
        f = kf.derAfterHyperParam(p)
        assert abs(f(X[1], X[2], False) - dK[1, 2]) <= 1e-14 * max(abs(dK[1, 2]), 1e-300)
    with pytest.raises(LookupError):
        kf.derAfterHyperParam(6)


def test_breeze_lbfgs_wrapper_best_seen_logic():   # optimization/Optimization.scala:37-61
    from gp_algos_b200.gp_optimizer import BreezeLbfgsOptimizer
    calls = []

    def f(p):   # concave quadratic with maximum at (1, -2)
        calls.append(np.array(p))
        return -((p[0] - 1.0) ** 2 + 2.0 * (p[1] + 2.0) ** 2), np.array([-2.0 * (p[0] - 1.0), -4.0 * (p[1] + 2.0)])

    x = BreezeLbfgsOptimizer(20).maximize(f, [5.0, 5.0])
    assert np.allclose(x, [1.0, -2.0], atol=1e-6)
    assert len(calls) >= 3                                # objective evaluated through the minus-wrapper, plus the final re-evaluation
    xm = BreezeLbfgsOptimizer(20).minimize(lambda p: (float(np.sum((np.asarray(p) - 3.0) ** 2)), 2.0 * (np.asarray(p) - 3.0)), [0.0, 0.0, 0.0])
    assert np.allclose(xm, 3.0, atol=1e-6)
    # maxIter = 0-like behaviour: with one iteration the best-seen point is still returned, never something worse than the start
    x1 = BreezeLbfgsOptimizer(1).maximize(f, [5.0, 5.0])
    assert f(x1)[0] >= f([5.0, 5.0])[0]


def test_obtain_optimal_hyper_params_without_noise_throws_like_the_reference():
    """GpPredictor.scala:130-132 drops the noise entry from the start point and KernelRequisites.scala:55 then rejects the
    shorter vector: `optimizeNoise = false` always throws in the reference; no GPU is reached."""
    import gp_algos_b200 as gp
    pred = gp.GpPredictor(gp.GaussianRbfKernel(gp.GaussianRbfParams(1.0, [1.0, 2.0], 0.1)))
    X = np.random.default_rng(0).uniform(size=(10, 2)); y = X[:, 0]
    with pytest.raises(ValueError, match="does not equal to 4"):
        pred.obtainOptimalHyperParams(X, None, y, optimizeNoise=False)


def test_synthetic_workloads_match_the_oracles_generators():
    """bench.py / tools build their inputs with gp_algos_b200.synthetic (so the measured paths never import oracle/); the tests
    use the oracle's generators.  Both must produce the same SURVEY.md 8(d) workloads bit for bit."""
    from gp_algos_b200 import synthetic
    from oracle import gp_oracle as orc
    for a, b in ((synthetic.make_c1(50, 7), orc.make_c1(50, 7)), (synthetic.make_c2(64, 8), orc.make_c2(64, 8)),
                 (synthetic.make_c2(40, 8, seed=5), orc.make_c2(40, 8, seed=5)), (synthetic.make_c3(33, 4), orc.make_c3(33, 4)),
                 (synthetic.make_c4_problem(3, n=20), orc.make_c4_problem(3, n=20))):
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_product_and_tools_never_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for sub in ("gp_algos_b200", "tools"):
        for fn in sorted(os.listdir(os.path.join(root, sub))):
            if fn.endswith(".py"):
                src = open(os.path.join(root, sub, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                    offenders.append(f"{sub}/{fn}")
    assert offenders == []
    # bench.py may only reach oracle/ from its CPU legs (cpu_baseline and --impl reference)
    bench = open(os.path.join(root, "bench.py")).read()
    gpu_arm = bench[bench.index("def run_gpk"):]
    head, tail = gpu_arm.split("t_cpu = cpu_eval_time", 1)
    assert "oracle" not in head.replace("oracle/ is imported by the cpu_baseline", "").replace("oracle LAPACK", "")


def test_ctypes_bindings_match_the_header_prototypes():
    """Every entry point the Python mirror calls is declared in include/gpk.h, and its argtypes list has exactly as many
    entries as the C prototype has parameters (a drifted binding would pass garbage through the ABI silently)."""
    hdr = open(os.path.join(ROOT, "include", "gpk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|double|const char\s*\*)\s+(gpk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    assert len(protos) >= 40
    lib = _lib.load()
    bound = 0
    for name, nargs in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        bound += 1
        assert len(fn.argtypes) == nargs, f"{name}: {len(fn.argtypes)} argtypes bound, {nargs} parameters in include/gpk.h"
    assert bound >= 35
    # and the other direction: nothing is bound that the header does not declare
    src = open(os.path.join(ROOT, "gp_algos_b200", "_lib.py")).read()
    for name in set(re.findall(r"lib\.(gpk_[a-z0-9_]+)\.", src)):
        assert name in protos, f"{name} bound in _lib.py but not declared in include/gpk.h"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/gpk.h is the drop-in boundary for JNA / JNI / cgo-style bindings: it must compile as C99 (no C++ types) and a
    plain C program must link against libgpk.so.  Without a GPU gpk_create fails cleanly (no CPU fallback), with one it succeeds."""
    src = tmp_path / "abi.c"
    src.write_text('#include "gpk.h"\n#include <stdio.h>\n'
                   "int main(void) { gpk_handle h = 0; int rc = gpk_create(&h, 0, 0);\n"
                   '  printf("%d %s\\n", rc, gpk_version()); if (rc == GPK_OK) gpk_destroy(h); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.lib_path())
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", libdir, "-lgpk", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    rc, version = out.stdout.split(" ", 1)
    assert "sm_100a" in version
    assert int(rc) == (0 if _have_gpu() else _lib.GPK_ECUDA)


def test_erfcx_table_matches_its_generator():
    """gp_algos_b200/csrc/gpk_erfcx_table.inc (the EP site kernel's phi/Phi table) is what tools/make_erfcx_table.py generates,
    and the table is as accurate as the kernel's header says (erfcx 4.4e-16, phi/Phi 1.8e-14 relative)."""
    mp = pytest.importorskip("mpmath")
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_erfcx_table", os.path.join(root, "tools", "make_erfcx_table.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    rows = [l for l in open(os.path.join(root, "gp_algos_b200", "csrc", "gpk_erfcx_table.inc")) if l.startswith("{")]
    assert len(rows) == gen.NI
    T = np.array([[float.fromhex(v) for v in r.strip().strip("{},").split(", ")] for r in rows])
    assert T.shape == (gen.NI, gen.DEG + 1)
    for i in (0, 7, 31, 63):
        assert np.array_equal(T[i], np.array(gen.fit(i)))
    xs = np.linspace(0, 8, 801)[:-1]
    ref = np.array([float(gen.erfcx(mp.mpf(float(x)))) for x in xs])
    assert np.max(np.abs(gen.table_erfcx(T, xs) / ref - 1)) < 6e-16
    zs = np.linspace(-11.3, 11.3, 453)
    rr = np.array([float(mp.npdf(mp.mpf(float(z))) / mp.ncdf(mp.mpf(float(z)))) for z in zs])
    assert np.max(np.abs(gen.ratio(T, zs) / rr - 1)) < 3e-14


def test_ep_fast_site_update_algebra_matches_the_reference_formulas():
    """The EP site kernel's scalar update (gpk_ep.cu: ep_site_fast -- cavity through 1 - tau_i Sigma_ii, phi/Phi through the erfcx
    table, c and g without 1/sig_hat on the chain) restated in NumPy against the formulas as written in
    EpParameterEstimator.scala:41-55, over the ranges an EP run visits."""
    mp = pytest.importorskip("mpmath")
    import importlib.util, os
    from scipy.stats import norm
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_erfcx_table", os.path.join(root, "tools", "make_erfcx_table.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    rows = [l for l in open(os.path.join(root, "gp_algos_b200", "csrc", "gpk_erfcx_table.inc")) if l.startswith("{")]
    T = np.array([[float.fromhex(v) for v in r.strip().strip("{},").split(", ")] for r in rows])
    rng = np.random.default_rng(7)
    m = 20000
    sii = rng.uniform(0.02, 3.0, m); t_old = rng.uniform(0.0, 0.3, m) / sii; mui = rng.normal(0, 2, m)
    n_old = rng.normal(0, 1, m) * t_old; y = rng.choice([-1.0, 1.0], m)
    # the reference, as written
    ct = 1 / sii - t_old; cn = mui / sii - n_old
    csig = 1 / ct; cmu = cn * csig
    z = y * cmu / np.sqrt(1 + csig)
    ok = np.abs(z) < 11.0                                    # (beyond: the kernel's libdevice fallback)
    ratio = norm.pdf(z) / norm.cdf(z)
    mu_hat = cmu + y * csig * ratio / np.sqrt(1 + csig)
    sig_hat = csig - csig ** 2 * ratio / (1 + csig) * (z + ratio)
    dtau = 1 / sig_hat - ct - t_old; n_new = mu_hat / sig_hat - cn
    c = dtau / (1 + dtau * sii); dnu = n_new - n_old; g = dnu - c * (mui + dnu * sii)
    # the kernel's formulation
    den = 1 - t_old * sii; num = mui - n_old * sii
    r = 1 / den; rs = 1 / sii; rt = np.sqrt(den) / np.sqrt(den + sii)
    csig2 = sii * r; cmu2 = num * r; ct2 = den * rs; cn2 = num * rs
    z2 = (y * cmu2) * rt
    ratio2 = gen.ratio(T, z2)
    mu_hat2 = cmu2 + (y * csig2 * rt) * ratio2
    sig_hat2 = csig2 - (csig2 * rt) ** 2 * (ratio2 * (z2 + ratio2))
    c2 = (sii - sig_hat2) * rs * rs; g2 = (mu_hat2 - mui) * rs
    dtau2 = 1 / sig_hat2 - rs; n_new2 = mu_hat2 / sig_hat2 - cn2

    def rel(a, b):
        return np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-3 * np.abs(b[ok]).max()))
    assert ok.sum() > 0.95 * m
    assert rel(ct2, ct) < 1e-12 and rel(cn2, cn) < 1e-12
    assert rel(dtau2, dtau) < 1e-10 and rel(n_new2, n_new) < 1e-10      # (1/sig_hat - 1/sii: a difference of close numbers)
    assert rel(c2, c) < 1e-10 and rel(g2, g) < 1e-10


def test_triangular_solve_host_logic_passes_row_major_operands_without_copy():
    """MatrixUtils.forwardSolve / backSolve (utils/MatrixUtils.scala:17-35): a row-major triangular operand IS the column-major
    storage of its transpose, so the host mirror hands the caller's buffer to gpk_trsm as it lies and flips the view flag;
    a column-major operand goes through unchanged; `upper` (the EFFECTIVE operand's shape) never changes."""
    from gp_algos_b200 import matrix_utils as MU
    calls = []

    class FakeLib:
        def gpk_trsm(self, h, upper, transposed, T, n, ldt, B, nrhs, ldb, X, ldx):
            calls.append((upper, transposed, T.value, n, ldt, nrhs))
            return 0

    class FakeHandle:
        h = None
        lib = FakeLib()

        def check(self, status):
            assert status == 0

    n = 7
    Lc = np.ascontiguousarray(np.tril(np.arange(1.0, n * n + 1).reshape(n, n)))      # row-major
    Lf = np.asfortranarray(Lc)                                                         # column-major
    b = np.ones(n)
    MU.forwardSolve(Lc, b, handle=FakeHandle())
    MU.forwardSolve(Lf, b, handle=FakeHandle())
    MU.backSolve(Lc, b, transposed=True, handle=FakeHandle())
    MU.backSolve(Lf, b, transposed=True, handle=FakeHandle())
    (u0, t0, p0, *_), (u1, t1, p1, *_), (u2, t2, p2, *_), (u3, t3, p3, *_) = calls
    assert (u0, t0) == (0, 1) and p0 == Lc.ctypes.data          # caller's row-major buffer, no copy, flag flipped
    assert (u1, t1) == (0, 0) and p1 == Lf.ctypes.data          # caller's column-major buffer, unchanged
    assert (u2, t2) == (1, 0) and p2 == Lc.ctypes.data
    assert (u3, t3) == (1, 1) and p3 == Lf.ctypes.data
    with pytest.raises(_lib.IllegalArgumentError):
        MU.forwardSolve(np.ones((3, 4)), np.ones(3), handle=FakeHandle())              # require(...) at MatrixUtils.scala:125

"""CPU: the lockstep multi-start L-BFGS driver of gp_algos_b200/batched.py on closed-form objectives (the GPU objective is
exercised by tests/test_gpu_batched.py)."""
import numpy as np

from gp_algos_b200.batched import LockstepLbfgs


def test_quadratics_converge_with_one_batched_call_per_round():
    rng = np.random.default_rng(0)
    R, P = 7, 5
    A = [(lambda M: M @ M.T + np.eye(P))(rng.standard_normal((P, P))) for _ in range(R)]
    b = rng.standard_normal((R, P))
    calls = []

    def f(points, which):
        calls.append(len(which))
        vals = np.array([0.5 * p @ A[r] @ p - b[r] @ p for p, r in zip(points, which)])
        grads = np.stack([A[r] @ p - b[r] for p, r in zip(points, which)])
        return vals, grads

    opt = LockstepLbfgs(maxIter=60)
    x, v = opt.minimize(f, rng.standard_normal((R, P)))
    for r in range(R):
        xs = np.linalg.solve(A[r], b[r])
        assert np.allclose(x[r], xs, rtol=1e-6, atol=1e-7)
        assert abs(v[r] - (0.5 * xs @ A[r] @ xs - b[r] @ xs)) < 1e-10
    assert opt.rounds == len(calls) and opt.evaluations == sum(calls)
    assert calls[0] == R and opt.rounds < 4 * 60            # lockstep: rounds ~ iterations of the slowest restart, not R times that


def test_rosenbrock_restarts_and_best_seen_logic():
    def f(points, which):
        x, y = points[:, 0], points[:, 1]
        v = (1 - x) ** 2 + 100 * (y - x * x) ** 2
        g = np.stack([-2 * (1 - x) - 400 * x * (y - x * x), 200 * (y - x * x)], axis=1)
        return v, g

    starts = np.array([[-1.2, 1.0], [0.0, 0.0], [2.0, 2.0], [1.0, 1.0]])
    x, v = LockstepLbfgs(maxIter=200).minimize(f, starts)
    assert np.allclose(x, 1.0, atol=1e-4) and np.all(v < 1e-8)
    # a capped run returns, per restart, a point that is never worse than its start (best-seen, Optimization.scala:44-55)
    x2, v2 = LockstepLbfgs(maxIter=3).minimize(f, starts)
    assert np.all(v2 <= f(starts, None)[0] + 1e-15) and np.allclose(f(x2, None)[0], v2)


def test_maximize_and_failed_evaluations_back_off():
    def f(points, which):                                    # concave with a forbidden region (a "not positive definite" restart)
        v = -np.sum((points - 2.0) ** 2, axis=1)
        v = np.where(points[:, 0] > 3.5, -np.inf, v)
        return v, -2 * (points - 2.0)

    x, v = LockstepLbfgs(maxIter=50).maximize(f, np.array([[0.0, 0.0], [3.4, 5.0], [-4.0, 1.0]]))
    assert np.allclose(x, 2.0, atol=1e-6) and np.allclose(v, 0.0, atol=1e-10)

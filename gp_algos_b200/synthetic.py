"""Synthetic workloads of BASELINE.json's configs (distributions, seeds and hyper-parameters), used by
bench.py and tools/ so that the measured paths never touch oracle/.  All inputs come from numpy.random.default_rng(seed) in
FP64.  (The test oracle carries the same generators; tests/test_abi_and_host.py checks that they agree bit for bit.)"""
from __future__ import annotations

import numpy as np


def pack_theta(signal_var, length_scales, noise_var) -> np.ndarray:
    """theta = [signalVar, lengthScale_1..D, noiseVar] (utils/KernelRequisites.scala:39-58)."""
    return np.concatenate([[float(signal_var)], np.asarray(length_scales, dtype=np.float64), [float(noise_var)]])


def make_c1(n=1000, m=500, seed=1):
    """C1: 1-D sin(x) + noise, 500 test points on a grid."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 10, size=(n, 1))
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(n)
    xs = np.linspace(0, 10, m)[:, None]
    return x, y, xs, pack_theta(1.0, [1.0], 0.1)


def make_c2(n=8192, D=8, seed=2):
    """C2 (and C5 with seed 5, n = 65536): X ~ U(0,1)^D, y = sin(X w) + 0.1 N(0,1), theta = (1, 0.7.., 0.1)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, size=(n, D))
    w = rng.standard_normal(D)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    return X, y, pack_theta(1.0, [0.7] * D, 0.1)


def make_c3(n=4096, D=4, seed=3):
    """C3: two-class data, labels = sign(X w + 0.3 N(0,1)) as int32 in {-1, +1}, theta = (1, 1.., 0)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    w = rng.standard_normal(D)
    t = np.sign(X @ w + 0.3 * rng.standard_normal(n)).astype(np.int32)
    t[t == 0] = 1
    return X, t, pack_theta(1.0, [1.0] * D, 0.0)


def make_c4_problem(b, n=1024, D=8, m=17):
    """C4: problem b of the batch (seed 1000 + b): as C2 with its own hyper-parameters and m = 2 D + 1 test rows."""
    rng = np.random.default_rng(1000 + b)
    X = rng.uniform(0, 1, size=(n, D))
    w = rng.standard_normal(D)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    sf = 10 ** rng.uniform(-0.3, 0.3)
    ls = 10 ** rng.uniform(-0.5, 0.2, size=D)
    Xs = rng.uniform(0, 1, size=(m, D))
    return X, y, Xs, pack_theta(sf, ls, 0.1)

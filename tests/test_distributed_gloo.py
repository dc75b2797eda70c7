"""CPU: the N>1 path of the large-GP solver (gp_algos_b200/distributed.py, BASELINE.json config 5) -- 2-D block-cyclic
ownership, panel-piece layout, look-ahead schedule, broadcasts / all-reduces -- with world_size 2 and 4 on gloo.

The product's block backend is the CUDA library (GpuBlockOps); here the test injects a NumPy stand-in for the per-block
numerics (built on the oracle's K builder) so that the host logic and the communication pattern run without a GPU.
Results are compared with the oracle's LAPACK flavour on the same inputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from gp_algos_b200.distributed import (BlockCyclicGrid, DistributedGp, Mat, Vec, choose_grid, count_from,
                                       first_at_least)


def test_grid_arithmetic():
    assert choose_grid(1) == (1, 1) and choose_grid(2) == (1, 2) and choose_grid(4) == (2, 2) and choose_grid(8) == (2, 4)
    for P in (1, 2, 3, 4):
        for q in range(P):
            for g in range(0, 9):
                f = first_at_least(g, q, P)
                assert f >= g and f % P == q and f - g < P
                for nt in range(0, 12):
                    assert count_from(g, q, P, nt) == len([i for i in range(g, nt) if i % P == q])


@pytest.mark.parametrize("Pr,Pc,nt", [(1, 1, 5), (2, 1, 7), (1, 2, 6), (2, 2, 9), (4, 2, 11), (2, 4, 8)])
def test_block_cyclic_cover_and_panel_layout(Pr, Pc, nt):
    grids = [BlockCyclicGrid(nt, Pr, Pc, r) for r in range(Pr * Pc)]
    owned = {}
    for g in grids:
        assert g.rank_of(g.pr, g.pc) == g.rank
        for i in g.row_blocks():
            for j in g.col_blocks():
                assert g.owner(i, j) == g.rank
                assert (i, j) not in owned
                owned[(i, j)] = g.rank
    assert len(owned) == nt * nt                                  # disjoint cover
    for k in range(nt):
        first, cnt, off = grids[0].panel_layout(k)
        rows = []
        for q in range(Pr):
            blk = [first[q] + t * Pr for t in range(cnt[q])]
            assert all(k < b < nt and b % Pr == q for b in blk)
            assert off[q] == len(rows)
            rows += blk
        assert sorted(rows) == list(range(k + 1, nt))             # every panel row in exactly one piece


class NumpyBlockOps:
    """Test-only stand-in for GpuBlockOps: same interface, NumPy arithmetic on CPU tensors (gloo can move them)."""

    def __init__(self):
        from oracle import gp_oracle as orc
        self.orc = orc

    def alloc(self, count, zero=False):
        return torch.zeros(max(int(count), 1), dtype=torch.float64)

    def alloc_int(self, count):
        return torch.zeros(max(int(count), 1), dtype=torch.int32)

    def upload(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1).copy())

    class _Null:
        def __enter__(self): return self
        def __exit__(self, *a): return False

    def run(self):
        return self._Null()

    def synchronize(self):
        pass

    @staticmethod
    def _m(m: Mat, rows, cols):
        return torch.as_strided(m.buf, (cols, rows), (m.ld, 1), m.off).numpy().T       # writable [r, c] view

    @staticmethod
    def _v(v: Vec, n):
        return v.buf.numpy()[v.off:v.off + n]

    def cov_cross(self, X1, m, X2, n, D, theta, K):
        K0 = np.array(theta, dtype=np.float64); K0[-1] = 0.0
        self._m(K, m, n)[...] = self.orc.fast_build_kernel_matrix(self._m(X1, m, D), K0, self._m(X2, n, D))

    def add_diag(self, A, n, v):
        a = self._m(A, n, n)
        a[np.arange(n), np.arange(n)] += v

    def potrf_inv(self, A, Li, N, info, info_off):
        a = self._m(A, N, N)
        try:
            L = np.linalg.cholesky(np.tril(a) + np.tril(a, -1).T)
        except np.linalg.LinAlgError:
            info[info_off] = 1
            return
        a[...] = L
        self._m(Li, N, N)[...] = np.linalg.inv(L)

    def sum_log_diag(self, A, n, out, accumulate):
        s = float(np.log(np.diag(self._m(A, n, n))).sum())
        out.buf[out.off] = (float(out.buf[out.off]) if accumulate else 0.0) + s

    def gemm_nt(self, m, p, k, alpha, P, Q, beta, Cm, q_lower_tri=False):
        q = self._m(Q, p, k)
        if q_lower_tri:
            q = np.tril(q)
        c = self._m(Cm, m, p)
        r = alpha * (self._m(P, m, k) @ q.T)
        c[...] = r + (beta * c if beta != 0.0 else 0.0)

    def gemv(self, trans, m, ncols, alpha, M, x, beta, y):
        mm = self._m(M, m, ncols)
        if trans:
            r = alpha * (mm.T @ self._v(x, m)); out = self._v(y, ncols)
        else:
            r = alpha * (mm @ self._v(x, ncols)); out = self._v(y, m)
        out[...] = r + (beta * out if beta != 0.0 else 0.0)


def _worker(rank, world, port, grid, n, nb, q):
    import torch.distributed as dist
    from oracle import gp_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, y, theta = orc.make_c2(n=n, D=3, seed=11)
    solver = DistributedGp(ops=NumpyBlockOps(), grid=grid, nb=nb)
    fit = solver.fit(X, y, theta)
    L = solver.gather_factor()
    res = solver.residual(y, fit.alphaVec)
    pm, ps = solver.predict(X[:9] + 0.01, fit.alphaVec)
    q.put((rank, fit.logLikelihood, fit.alphaVec, L, res, solver.launch_gemm, pm, ps))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,grid,n,nb", [(2, (2, 1), 700, 128), (2, (1, 2), 640, 128), (4, (2, 2), 900, 128)])
def test_distributed_fit_matches_oracle_on_gloo(world, grid, n, nb):
    from oracle import gp_oracle as orc
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, n, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, y, theta = orc.make_c2(n=n, D=3, seed=11)
    L_o, alpha_o = orc.fast_precompute(X, y, theta)
    ll_o = orc.fast_loglik(alpha_o, L_o, y)
    total_gemms = 0
    mean_o = orc.fast_build_kernel_matrix(X[:9] + 0.01, np.r_[theta[:-1], 0.0], X) @ alpha_o
    _, S_o, _ = orc.fast_compute_posterior(X, X[:9] + 0.01, L_o, alpha_o, theta)
    for rank, ll, alpha, L, res, ngemm, pm, ps in out:
        assert np.allclose(ps, S_o, rtol=1e-7, atol=1e-9 * np.abs(S_o).max())             # full predictive covariance, replicated
        assert np.allclose(pm, mean_o, rtol=1e-8, atol=1e-9 * np.abs(mean_o).max())       # posterior mean, replicated
        assert abs(ll - ll_o) <= 1e-9 * abs(ll_o)
        assert np.allclose(alpha, alpha_o, rtol=1e-8, atol=1e-9 * np.abs(alpha_o).max())   # replicated on every rank
        assert res < 1e-10
        if rank == 0:
            assert np.linalg.norm(L - L_o) <= 1e-10 * np.linalg.norm(L_o)
        else:
            assert L is None
        total_gemms += ngemm
    nt = (n + nb - 1) // nb
    # every strictly-lower block column pair (j > k) is updated exactly once per owning process row that has rows >= j
    assert total_gemms > 0 and total_gemms <= grid[0] * nt * (nt - 1) // 2


def test_single_process_grid_needs_no_process_group():
    from oracle import gp_oracle as orc
    X, y, theta = orc.make_c2(n=300, D=2, seed=4)
    solver = DistributedGp(ops=NumpyBlockOps(), nb=128)
    fit = solver.fit(X, y, theta, sigmaNoise=0.05)
    L_o, alpha_o = orc.fast_precompute(X, y, theta, 0.05)
    assert abs(fit.logLikelihood - orc.fast_loglik(alpha_o, L_o, y)) <= 1e-9 * abs(orc.fast_loglik(alpha_o, L_o, y))
    assert np.allclose(fit.alphaVec, alpha_o, rtol=1e-8, atol=1e-12)
    assert np.linalg.norm(solver.gather_factor() - L_o) <= 1e-10 * np.linalg.norm(L_o)

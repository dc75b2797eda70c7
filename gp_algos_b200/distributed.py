"""One large GP across the GPUs of a node (BASELINE.json config 5: n = 65536, D = 8): K build, FP64 Cholesky and
alpha = K^-1 y on a 2-D block-cyclic process grid, one process per GPU, NCCL panel broadcasts over NVLink.

What it replaces: `GpPredictor.preComputeComponents` + `logLikelihood` (gp/regression/GpPredictor.scala:104-124,144-149) --
`buildKernelMatrix` (utils/MatrixUtils.scala:57-70), breeze `cholesky` (LAPACK dpotrf 'L', GpPredictor.scala:120) and the
two triangular solves `backSolve(L.t, forwardSolve(L, targets))` (GpPredictor.scala:121-122) -- for a training set too large
for one device to factor quickly.  The reference has no distributed code; the decomposition below is this library's own.

Layout.  n is padded to nt blocks of nb rows (padding = identity block, so chol(K_pad) = diag(L, I)).  Block (i, j) of the lower
triangle lives on grid rank (i mod Pr, j mod Pc); each rank stores its blocks as one column-major local matrix (local row
block i // Pr, local column block j // Pc).  K is never communicated: every rank builds its own blocks from the replicated
X (8 n D bytes).

Right-looking factorisation with look-ahead 1.  Step k:
  * owner of (k, k): L_kk, L_kk^-1 = potrf+inverse of the diagonal block (gpk_potrf_inv_block_dev); L_kk^-1 is broadcast;
  * the Pr ranks of process column k mod Pc: L_ik = A_ik L_kk^-T for their rows (DMMA GEMM with the known inverse, no
    TRSM), written to their piece of the panel buffer; each piece is broadcast, so every rank ends up with the whole
    panel, stored piece by piece (piece q = rows i == q mod Pr, ascending);
  * every rank: A_ij -= L_ik L_jk^T for its blocks, one DMMA GEMM per local block column (left operand = own piece, right
    operand = block j of piece j mod Pr).  Column k+1 is updated first and its panel is factored and put on the wire
    before the rest of the trailing update is issued, so the broadcasts overlap the update (NCCL runs on its own stream).
The forward solve L z = y rides along (every rank has the panel, so z is replicated with no extra traffic); the back solve
L^T alpha = z walks the block columns backwards with one nb-vector all-reduce per step.

No CPU fallback: `GpuBlockOps` is the only block backend in this package and it needs libgpk.so + a CUDA device.  (The
gloo world_size-2 CPU test injects its own NumPy block backend to exercise the grid / schedule / communication logic.)
"""
from __future__ import annotations

import ctypes as C
import math
from typing import NamedTuple, Optional

import numpy as np

from . import _lib


class Mat(NamedTuple):
    """A column-major sub-matrix of a flat FP64 buffer: element (r, c) at buf[off + r + c*ld]."""
    buf: object
    off: int
    ld: int

    def at(self, r: int, c: int) -> "Mat":
        return Mat(self.buf, self.off + r + c * self.ld, self.ld)


class Vec(NamedTuple):
    buf: object
    off: int

    def at(self, i: int) -> "Vec":
        return Vec(self.buf, self.off + i)


# ----------------------------------------------------------------------------------------------------------------
# grid arithmetic (pure host logic, covered by the CPU tests)
# ----------------------------------------------------------------------------------------------------------------
def choose_grid(world: int) -> tuple[int, int]:
    """Pr x Pc with Pr * Pc == world, as square as possible, Pr <= Pc: fewer process rows keep each rank's trailing-update
    GEMMs tall (measured on 8 B200, n = 65536: 2x4 0.480 s, 4x2 0.496 s; on 2: 1x2 1.54 s, 2x1 1.61 s)."""
    pr = int(math.isqrt(world))
    while world % pr:
        pr -= 1
    return pr, world // pr


def first_at_least(g: int, q: int, P: int) -> int:
    """Smallest block index i >= g with i mod P == q."""
    return g + ((q - g) % P)


def count_from(g: int, q: int, P: int, nt: int) -> int:
    """Number of block indices i in [g, nt) with i mod P == q."""
    f = first_at_least(g, q, P)
    return 0 if f >= nt else (nt - 1 - f) // P + 1


class BlockCyclicGrid:
    """2-D block-cyclic map of an nt x nt block matrix onto a Pr x Pc grid (ranks numbered row-major)."""

    def __init__(self, nt: int, Pr: int, Pc: int, rank: int):
        if Pr <= 0 or Pc <= 0 or not 0 <= rank < Pr * Pc:
            raise ValueError("bad process grid")
        self.nt, self.Pr, self.Pc, self.rank = nt, Pr, Pc, rank
        self.pr, self.pc = divmod(rank, Pc)
        self.nrow_blocks = count_from(0, self.pr, Pr, nt)   # local block rows
        self.ncol_blocks = count_from(0, self.pc, Pc, nt)   # local block columns

    def rank_of(self, q: int, c: int) -> int:
        return q * self.Pc + c

    def owner(self, i: int, j: int) -> int:
        return self.rank_of(i % self.Pr, j % self.Pc)

    def row_blocks(self, q: Optional[int] = None):
        return range(self.pr if q is None else q, self.nt, self.Pr)

    def col_blocks(self, c: Optional[int] = None):
        return range(self.pc if c is None else c, self.nt, self.Pc)

    def panel_layout(self, k: int):
        """Pieces of panel k (block rows i > k): per process row q -> (first block, block count, offset in blocks)."""
        first, cnt, off, o = [], [], [], 0
        for q in range(self.Pr):
            first.append(first_at_least(k + 1, q, self.Pr))
            cnt.append(count_from(k + 1, q, self.Pr, self.nt))
            off.append(o)
            o += cnt[-1]
        return first, cnt, off


# ----------------------------------------------------------------------------------------------------------------
# block backend on the GPU: every numeric operation is a libgpk kernel
# ----------------------------------------------------------------------------------------------------------------
class GpuBlockOps:
    """libgpk device-level calls on torch-allocated HBM; torch is used for allocation, copies, streams and NCCL only."""

    def __init__(self, device: int):
        import torch
        if not torch.cuda.is_available():
            raise _lib.GpkError(_lib.GPK_ECUDA, "distributed GP needs a CUDA device (libgpk has no CPU fallback)")
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        # a non-default stream: the handle, torch copies and the NCCL hand-offs all order on it
        self.stream = torch.cuda.Stream(self.device, priority=-1)
        self.handle = _lib.Handle(device, stream=self.stream.cuda_stream)
        self.lib = self.handle.lib
        # auxiliary lanes: the bulk of a trailing update is spread over two more streams, so the partial last wave of one
        # block-column GEMM is filled by the next one (a lane has its own handle because a handle is bound to one stream)
        self.naux = 2
        self.aux_streams = [torch.cuda.Stream(self.device) for _ in range(self.naux)]
        self.aux_handles = [_lib.Handle(device, stream=st.cuda_stream) for st in self.aux_streams]

    # -- memory ---------------------------------------------------------------------------------------------
    def alloc(self, count: int, zero: bool = False):
        f = self.torch.zeros if zero else self.torch.empty
        return f(max(int(count), 1), dtype=self.torch.float64, device=self.device)

    def alloc_int(self, count: int):
        return self.torch.zeros(max(int(count), 1), dtype=self.torch.int32, device=self.device)

    def upload(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(self.device)

    def run(self):
        return self.torch.cuda.stream(self.stream)

    def synchronize(self):
        self.stream.synchronize()

    @staticmethod
    def _p(buf, off, itemsize=8):
        return C.c_void_p(buf.data_ptr() + itemsize * off)

    def _pm(self, m: Mat):
        return self._p(m.buf, m.off)

    # -- numeric blocks --------------------------------------------------------------------------------------
    def cov_cross(self, X1: Mat, m: int, X2: Mat, n: int, D: int, theta: np.ndarray, K: Mat):
        self.handle.check(self.lib.gpk_cov_cross_se_ard_dev(self.handle.h, self._pm(X1), m, X1.ld, self._pm(X2), n, X2.ld, D,
                                                            _lib.ptr(theta), self._pm(K), K.ld))

    def add_diag(self, A: Mat, n: int, v: float):
        self.handle.check(self.lib.gpk_add_diag_dev(self.handle.h, self._pm(A), A.ld, n, float(v)))

    def potrf_inv(self, A: Mat, Li: Mat, N: int, info, info_off: int):
        assert A.ld == N and Li.ld == N
        self.handle.check(self.lib.gpk_potrf_inv_block_dev(self.handle.h, self._pm(A), self._pm(Li), N, self._p(info, info_off, 4)))

    def sum_log_diag(self, A: Mat, n: int, out: Vec, accumulate: bool):
        self.handle.check(self.lib.gpk_sum_log_diag_dev(self.handle.h, self._pm(A), A.ld, n, self._p(out.buf, out.off), int(accumulate)))

    def gemm_nt(self, m: int, p: int, k: int, alpha: float, P: Mat, Q: Mat, beta: float, Cm: Mat, q_lower_tri: bool = False,
                lane: Optional[int] = None):
        hd = self.handle if lane is None else self.aux_handles[lane]
        hd.check(self.lib.gpk_gemm_nt_dev(hd.h, m, p, k, float(alpha), self._pm(P), P.ld, self._pm(Q), Q.ld,
                                          float(beta), self._pm(Cm), Cm.ld, int(q_lower_tri)))

    # -- stream plumbing for the auxiliary lanes (lane None = the main stream) ----------------------------------
    def record(self, lane: Optional[int] = None):
        ev = self.torch.cuda.Event()
        ev.record(self.stream if lane is None else self.aux_streams[lane])
        return ev

    def wait(self, ev, lane: Optional[int] = None):
        if ev is not None:
            (self.stream if lane is None else self.aux_streams[lane]).wait_event(ev)

    def gemv(self, trans: bool, m: int, ncols: int, alpha: float, M: Mat, x: Vec, beta: float, y: Vec):
        self.handle.check(self.lib.gpk_gemv_dev(self.handle.h, int(trans), m, ncols, float(alpha), self._pm(M), M.ld,
                                                self._p(x.buf, x.off), float(beta), self._p(y.buf, y.off)))


def _view(torch, m: Mat, rows: int, cols: int):
    """Strided torch view of a column-major sub-matrix, indexed [c, r] (copies / fills only)."""
    return torch.as_strided(m.buf, (cols, rows), (m.ld, 1), m.off)


class DistributedFit(NamedTuple):
    logLikelihood: float
    alphaVec: np.ndarray       # replicated on every rank
    info: int                  # 0, or the failing leading minor (1-based), like gpk_last_info
    seconds: float             # device time of build + factor + solves on this rank (CUDA events), nan on CPU backends


class DistributedGp:
    """`GpPredictor.preComputeComponents` + `logLikelihood` for one large training set on a Pr x Pc grid of GPUs.

    Call `fit` collectively from every rank of `group` (torch.distributed, NCCL; rank r drives cuda:LOCAL_RANK)."""

    def __init__(self, ops=None, grid: Optional[tuple[int, int]] = None, nb: int = 1024, group=None, device: Optional[int] = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.dist_on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.dist_on else 1
        self.rank = dist.get_rank(group) if self.dist_on else 0
        self.Pr, self.Pc = grid if grid else choose_grid(self.world)
        if self.Pr * self.Pc != self.world:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, f"process grid {self.Pr}x{self.Pc} does not match world size {self.world}")
        if nb <= 0 or nb % 128:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "block size nb must be a positive multiple of 128")
        self.nb = nb
        if ops is None:
            if device is None:
                import os
                device = int(os.environ.get("LOCAL_RANK", self.rank))
            ops = GpuBlockOps(device)
        self.ops = ops
        self.A = None
        self.launch_gemm = 0
        self.bcast_bytes = 0          # bytes this rank sent or received through broadcasts (bench.py reports it)

    # -- communication (no-ops on a 1 x 1 grid) -----------------------------------------------------------------
    def _global_rank(self, r: int) -> int:
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def _bcast(self, t, src: int):
        if self.world == 1:
            return None
        self.bcast_bytes += t.numel() * t.element_size()
        return self.dist.broadcast(t, src=self._global_rank(src), group=self.group, async_op=True)

    def _allreduce(self, t, op=None):
        if self.world > 1:
            self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM, group=self.group)

    @staticmethod
    def _wait(works):
        for w in works:
            if w is not None:
                w.wait()

    # -- the fit ---------------------------------------------------------------------------------------------
    def fit(self, X, y, theta, sigmaNoise: Optional[float] = None) -> DistributedFit:
        torch, ops, nb = self.torch, self.ops, self.nb
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n, D = X.shape
        if y.shape[0] != n or theta.shape[0] != D + 2:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: shapes of X, targets and hyper-parameters disagree")
        nt = (n + nb - 1) // nb
        npad = nt * nb
        g = self.grid = BlockCyclicGrid(nt, self.Pr, self.Pc, self.rank)
        self.n, self.npad, self.nt = n, npad, nt
        mloc, cloc = g.nrow_blocks * nb, g.ncol_blocks * nb
        sn2 = float(theta[-1]) * float(theta[-1]) + (float(sigmaNoise) if sigmaNoise is not None else 0.0)  # GpPredictor.scala:116

        def gather_rows(blocks):
            out = np.zeros((max(len(blocks), 1) * nb, D))
            for l, b in enumerate(blocks):
                lo, hi = b * nb, min((b + 1) * nb, n)
                if hi > lo:
                    out[l * nb:l * nb + hi - lo] = X[lo:hi]
            return np.asfortranarray(out)

        with ops.run():
            t0 = t1 = None
            if hasattr(ops, "stream"):
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            Xr_h, Xc_h = gather_rows(list(g.row_blocks())), gather_rows(list(g.col_blocks()))
            Xr = Mat(ops.upload(Xr_h.T), 0, Xr_h.shape[0])     # .T of an F-array is C-contiguous: column-major payload
            Xc = Mat(ops.upload(Xc_h.T), 0, Xc_h.shape[0])
            if self.A is None or self.A.buf.numel() < mloc * cloc:
                self.A = None
                self.A = Mat(ops.alloc(mloc * cloc), 0, max(mloc, 1))
            A = self.A = Mat(self.A.buf, 0, max(mloc, 1))
            self.Li = ops.alloc(nt * nb * nb, zero=True)          # every L_kk^-1, replicated (back solve)
            naux = getattr(ops, "naux", 0)
            NP = 3 if naux else 2                                 # panel buffers: with lanes a panel is read one step longer
            panel = [ops.alloc(max(nt - 1, 1) * nb * nb) for _ in range(NP)]
            Dbuf = Mat(ops.alloc(nb * nb), 0, nb)
            info = ops.alloc_int(nt)
            scal = ops.alloc(4, zero=True)                        # [sum log L_ii, y.alpha]
            ypad = np.zeros(npad); ypad[:n] = y
            # forward-solve work vector, piece-major: yq[q] = entries of y in the block rows == q mod Pr
            yq = [Vec(ops.upload(np.concatenate([ypad[b * nb:(b + 1) * nb] for b in g.row_blocks(q)] or [np.zeros(1)])), 0)
                  for q in range(g.Pr)]
            z = ops.alloc(npad, zero=True)
            if t0 is not None:
                if self.world > 1:
                    self.dist.barrier(group=self.group)
                t0.record()

            # ---- K blocks from the replicated X (MatrixUtils.scala:57-70; nothing n^2 is communicated) ----
            for lj, j in enumerate(g.col_blocks()):
                i0 = first_at_least(j, g.pr, g.Pr)
                if i0 >= nt:
                    continue
                li0 = i0 // g.Pr
                ops.cov_cross(Xr.at(li0 * nb, 0), mloc - li0 * nb, Xc.at(lj * nb, 0), nb, D, theta, A.at(li0 * nb, lj * nb))
                if i0 == j:
                    ops.add_diag(A.at(li0 * nb, lj * nb), nb, sn2)
            if npad > n:   # padding rows/columns -> identity
                pad0 = n - (nt - 1) * nb
                if (nt - 1) % g.Pr == g.pr and cloc:
                    _view(torch, A.at((g.nrow_blocks - 1) * nb + pad0, 0), nb - pad0, cloc).zero_()
                if (nt - 1) % g.Pc == g.pc and mloc:
                    _view(torch, A.at(0, (g.ncol_blocks - 1) * nb + pad0), mloc, nb - pad0).zero_()
                if g.owner(nt - 1, nt - 1) == self.rank:
                    _view(torch, A.at((g.nrow_blocks - 1) * nb + pad0, (g.ncol_blocks - 1) * nb + pad0), nb - pad0, nb - pad0
                          ).diagonal().fill_(1.0)

            # ---- factorisation + forward solve -------------------------------------------------------------
            def factor_panel(k):
                """Column k is fully updated: factor its diagonal block, solve the panel, start the broadcasts."""
                oq, oc = k % g.Pr, k % g.Pc
                Li_k = Mat(self.Li, k * nb * nb, nb)
                if (g.pr, g.pc) == (oq, oc):
                    blk = A.at((k // g.Pr) * nb, (k // g.Pc) * nb)
                    _view(torch, Dbuf, nb, nb).copy_(_view(torch, blk, nb, nb))
                    ops.potrf_inv(Dbuf, Li_k, nb, info, k)
                    _view(torch, blk, nb, nb).copy_(_view(torch, Dbuf, nb, nb))
                    ops.sum_log_diag(Dbuf, nb, Vec(scal, 0), True)
                works = [self._bcast(self.Li[k * nb * nb:(k + 1) * nb * nb], g.rank_of(oq, oc))]
                first, cnt, off = g.panel_layout(k)
                pbuf = panel[k % NP]
                if g.pc == oc and cnt[g.pr]:
                    self._wait(works)
                    rows = cnt[g.pr] * nb
                    src = A.at((first[g.pr] // g.Pr) * nb, (k // g.Pc) * nb)
                    piece = Mat(pbuf, off[g.pr] * nb * nb, rows)
                    ops.gemm_nt(rows, nb, nb, 1.0, src, Li_k, 0.0, piece, q_lower_tri=True)   # L_ik = A_ik L_kk^-T
                    _view(torch, src, rows, nb).copy_(_view(torch, piece, rows, nb))
                for q in range(g.Pr):
                    if cnt[q]:
                        works.append(self._bcast(pbuf[off[q] * nb * nb:(off[q] + cnt[q]) * nb * nb], g.rank_of(q, oc)))
                return works

            my_cols = list(g.col_blocks())
            pend = factor_panel(0)
            step_done = {}                 # k -> events closing the lanes' work of step k
            next_col_ready = None          # this rank's block column k+1 has received update k-1 (issued first on its lane)
            for k in range(nt):
                self._wait(pend)
                pend = []
                first, cnt, off = g.panel_layout(k)
                pbuf = panel[k % NP]
                Li_k = Mat(self.Li, k * nb * nb, nb)
                # forward substitution, replicated: z_k = L_kk^-1 y_k ; y_i -= L_ik z_k  (MatrixUtils.scala:17-21)
                zk = Vec(z, k * nb)
                ops.gemv(False, nb, nb, 1.0, Li_k, yq[k % g.Pr].at((k // g.Pr) * nb), 0.0, zk)
                for q in range(g.Pr):
                    if cnt[q]:
                        ops.gemv(False, cnt[q] * nb, nb, -1.0, Mat(pbuf, off[q] * nb * nb, cnt[q] * nb), zk, 1.0,
                                 yq[q].at((first[q] // g.Pr) * nb))
                if naux:
                    ev_panel = ops.record()                       # panel k has arrived (the main stream waited for NCCL)
                    for lane in range(naux):
                        ops.wait(ev_panel, lane)
                    for ev in step_done.pop(k - 2, []):           # panel buffer (k+1) % 3 was last read by the lanes at step k-2
                        ops.wait(ev)
                if k + 1 < nt and g.pc != (k + 1) % g.Pc:
                    pend = factor_panel(k + 1)        # not in the next panel's process column: just join its broadcasts
                col_ready, new_ready = next_col_ready, None
                for j in my_cols:
                    if j <= k:
                        continue
                    # column k+1 stays on the main stream (its panel is factored right behind it); the rest alternate lanes,
                    # a block column always on the same lane so that its successive updates stay ordered
                    lane = None if (j == k + 1 or not naux) else (j // g.Pc) % naux
                    i0 = first_at_least(j, g.pr, g.Pr)
                    if i0 < nt:
                        rows = (g.nrow_blocks - i0 // g.Pr) * nb
                        left = Mat(pbuf, off[g.pr] * nb * nb + ((i0 - first[g.pr]) // g.Pr) * nb, cnt[g.pr] * nb)
                        qj = j % g.Pr
                        right = Mat(pbuf, off[qj] * nb * nb + ((j - first[qj]) // g.Pr) * nb, cnt[qj] * nb)
                        if naux and j == k + 1:
                            ops.wait(col_ready)                   # update k-1 of this column ran on a lane
                        if naux:
                            ops.gemm_nt(rows, nb, nb, -1.0, left, right, 1.0, A.at((i0 // g.Pr) * nb, (j // g.Pc) * nb), lane=lane)
                        else:
                            ops.gemm_nt(rows, nb, nb, -1.0, left, right, 1.0, A.at((i0 // g.Pr) * nb, (j // g.Pc) * nb))
                        self.launch_gemm += 1
                        if naux and j == k + 2:
                            new_ready = ops.record(lane)          # first job of its lane in this step: next step's look-ahead waits on it
                    if j == k + 1:
                        pend = factor_panel(k + 1)    # look-ahead: panel k+1 goes on the wire before the rest of the update
                next_col_ready = new_ready
                if naux:
                    step_done[k] = [ops.record(lane) for lane in range(naux)]
            if naux:
                for evs in step_done.values():
                    for ev in evs:
                        ops.wait(ev)

            # ---- back solve L^T alpha = z (MatrixUtils.scala:23-27), block columns in reverse --------------------
            alpha = ops.alloc(npad, zero=True)
            aq = Vec(ops.alloc(mloc, zero=True), 0)               # alpha restricted to this rank's block rows
            s = ops.alloc(nb)
            for k in range(nt - 1, -1, -1):
                if self.rank == 0:
                    s.copy_(z[k * nb:(k + 1) * nb])
                else:
                    s.zero_()
                i0 = first_at_least(k + 1, g.pr, g.Pr)
                if g.pc == k % g.Pc and i0 < nt:
                    li0 = i0 // g.Pr
                    ops.gemv(True, mloc - li0 * nb, nb, -1.0, A.at(li0 * nb, (k // g.Pc) * nb), aq.at(li0 * nb), 1.0, Vec(s, 0))
                self._allreduce(s)
                ops.gemv(True, nb, nb, 1.0, Mat(self.Li, k * nb * nb, nb), Vec(s, 0), 0.0, Vec(alpha, k * nb))
                if k % g.Pr == g.pr:
                    aq.buf[(k // g.Pr) * nb:(k // g.Pr + 1) * nb].copy_(alpha[k * nb:(k + 1) * nb])

            # ---- log marginal likelihood (GpPredictor.scala:144-149) ------------------------------------------
            ops.gemv(True, npad, 1, 1.0, Mat(ops.upload(ypad), 0, npad), Vec(alpha, 0), 0.0, Vec(scal, 1))
            self._allreduce(scal[0:1])
            self._allreduce(info, op=self.dist.ReduceOp.MAX if self.world > 1 else None)
            if t1 is not None:
                t1.record()
            ops.synchronize()
            scal_h = scal.cpu().numpy()
            info_h = info.cpu().numpy()
            alpha_h = alpha.cpu().numpy()[:n].copy()
            seconds = t0.elapsed_time(t1) * 1e-3 if t0 is not None else float("nan")
        bad = np.nonzero(info_h)[0]
        minor = int(bad[0] * nb + info_h[bad[0]]) if bad.size else 0
        ll = -0.5 * float(scal_h[1]) - float(scal_h[0]) - 0.5 * n * math.log(2.0 * math.pi)
        self.X, self.theta, self.sn2 = X, theta, sn2
        if minor:
            raise _lib.NotPositiveDefiniteError(_lib.GPK_ENOTPD, f"matrix not positive definite: leading minor {minor}", minor)
        return DistributedFit(ll, alpha_h, minor, seconds)

    # -- checks used by tests / tools ------------------------------------------------------------------------
    def gather_factor(self) -> Optional[np.ndarray]:
        """The n x n lower factor on rank 0 (None elsewhere).  Small n only: test helper, not a data path."""
        g, nb, torch = self.grid, self.nb, self.torch
        loc = _view(torch, self.A, g.nrow_blocks * nb, g.ncol_blocks * nb).cpu().numpy().T   # [r, c]
        parts = [None] * self.world
        if self.world > 1:
            self.dist.all_gather_object(parts, loc, group=self.group)
        else:
            parts = [loc]
        if self.rank != 0:
            return None
        L = np.zeros((self.npad, self.npad))
        for r, part in enumerate(parts):
            gr = BlockCyclicGrid(self.nt, self.Pr, self.Pc, r)
            for li, i in enumerate(gr.row_blocks()):
                for lj, j in enumerate(gr.col_blocks()):
                    if i >= j:
                        L[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = part[li * nb:(li + 1) * nb, lj * nb:(lj + 1) * nb]
        return np.tril(L)[:self.n, :self.n]

    def predict_mean(self, Xs, alpha) -> np.ndarray:
        """Posterior mean K(X*, X) alpha (GpPredictor.scala:53-54) for the fitted training set; block column j of K(X*, X) is
        generated and contracted on rank j mod world, one all-reduce of m doubles at the end.  (The predictive variance needs a
        distributed multi-right-hand-side forward solve on the block-cyclic factor and is not built.)"""
        torch, ops, nb, n, npad = self.torch, self.ops, self.nb, self.n, self.npad
        Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
        m, D = Xs.shape
        if D != self.X.shape[1]:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: test point dimension")
        with ops.run():
            Xp = np.zeros((npad, D)); Xp[:n] = self.X
            Xd = Mat(ops.upload(np.asfortranarray(Xp).T), 0, npad)
            Xsd = Mat(ops.upload(np.asfortranarray(Xs).T), 0, m)
            ap = np.zeros(npad); ap[:n] = alpha
            a = ops.upload(ap)
            Kb = Mat(ops.alloc(m * nb), 0, m)
            out = ops.alloc(m, zero=True)
            for j in range(self.rank, self.nt, self.world):
                ops.cov_cross(Xsd, m, Xd.at(j * nb, 0), nb, D, self.theta, Kb)        # K(X*, X_j): m x nb, no noise
                ops.gemv(False, m, nb, 1.0, Kb, Vec(a, j * nb), 1.0, Vec(out, 0))
            self._allreduce(out)
            ops.synchronize()
            return out.cpu().numpy().copy()

    def predict(self, Xs, alpha):
        """computePosterior for the large GP (GpPredictor.scala:45-58): mean = K* alpha and the full m x m covariance
        sigma = K** - V^t V with V = L^-1 K*^t (its diagonal includes noiseVar^2, MatrixUtils.scala:63).  V^t is kept replicated
        (m x n, stored per process row like the panels) and swept block column by block column: the owners re-broadcast the
        panel of L, every rank applies V_k^t = B_k^t L_kk^-t and B_i^t -= V_k^t L_ik^t (DMMA GEMMs), then sigma -= V^t V."""
        torch, ops, nb, n, npad, nt, g = self.torch, self.ops, self.nb, self.n, self.npad, self.nt, self.grid
        Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
        m, D = Xs.shape
        if D != self.X.shape[1]:
            raise _lib.IllegalArgumentError(_lib.GPK_EINVAL, "requirement failed: test point dimension")
        M = (m + 127) // 128 * 128
        mean = self.predict_mean(Xs, alpha)
        A = self.A
        with ops.run():
            Xsp = np.zeros((M, D)); Xsp[:m] = Xs
            Xsd = Mat(ops.upload(np.asfortranarray(Xsp).T), 0, M)
            # B^t = K(X*, X) with the training rows in piece-major order (block rows == q mod Pr together); padded rows/cols zero
            Btq, ncol_q = [], []
            for q in range(g.Pr):
                blocks = list(g.row_blocks(q))
                Xq = np.zeros((max(len(blocks), 1) * nb, D))
                for l, b in enumerate(blocks):
                    lo, hi = b * nb, min((b + 1) * nb, n)
                    if hi > lo:
                        Xq[l * nb:l * nb + hi - lo] = self.X[lo:hi]
                cols = len(blocks) * nb
                buf = Mat(ops.alloc(M * max(cols, 1), zero=True), 0, M)
                if cols:
                    ops.cov_cross(Xsd, M, Mat(ops.upload(np.asfortranarray(Xq).T), 0, Xq.shape[0]), cols, D, self.theta, buf)
                    _view(torch, buf.at(m, 0), M - m, cols).zero_()                       # padded test rows
                    if blocks and blocks[-1] == nt - 1 and npad > n:                      # padded training rows
                        pad0 = n - (nt - 1) * nb
                        _view(torch, buf.at(0, (len(blocks) - 1) * nb + pad0), M, nb - pad0).zero_()
                Btq.append(buf); ncol_q.append(cols)
            sigma = Mat(ops.alloc(M * M, zero=True), 0, M)
            ops.cov_cross(Xsd, M, Xsd, M, D, self.theta, sigma)
            _view(torch, sigma, M, M)[m:, :].zero_(); _view(torch, sigma, M, M)[:, m:].zero_()
            ops.add_diag(sigma, m, float(self.theta[-1]) ** 2)                            # sameIndex noise of buildKernelMatrix(k, X*)
            piece_buf = ops.alloc(max(nt - 1, 1) * nb * nb)
            Tk = Mat(ops.alloc(M * nb), 0, M)
            for k in range(nt):
                oq, oc = k % g.Pr, k % g.Pc
                first, cnt, off = g.panel_layout(k)
                works = []
                if g.pc == oc and cnt[g.pr]:                                              # pack my piece of L's panel k
                    rows = cnt[g.pr] * nb
                    src = A.at((first[g.pr] // g.Pr) * nb, (k // g.Pc) * nb)
                    _view(torch, Mat(piece_buf, off[g.pr] * nb * nb, rows), rows, nb).copy_(_view(torch, src, rows, nb))
                for q in range(g.Pr):
                    if cnt[q]:
                        works.append(self._bcast(piece_buf[off[q] * nb * nb:(off[q] + cnt[q]) * nb * nb], g.rank_of(q, oc)))
                Bk = Btq[oq].at(0, (k // g.Pr) * nb)
                ops.gemm_nt(M, nb, nb, 1.0, Bk, Mat(self.Li, k * nb * nb, nb), 0.0, Tk, q_lower_tri=True)   # V_k^t = B_k^t L_kk^-t
                _view(torch, Bk, M, nb).copy_(_view(torch, Tk, M, nb))
                self._wait(works)
                for q in range(g.Pr):
                    if cnt[q]:
                        ops.gemm_nt(M, cnt[q] * nb, nb, -1.0, Tk, Mat(piece_buf, off[q] * nb * nb, cnt[q] * nb), 1.0,
                                    Btq[q].at(0, (first[q] // g.Pr) * nb))                                   # B_i^t -= V_k^t L_ik^t
                if self.world > 1:
                    ops.synchronize()      # piece_buf is reused by the next step's broadcasts
            for q in range(g.Pr):
                if ncol_q[q]:
                    ops.gemm_nt(M, M, ncol_q[q], -1.0, Btq[q], Btq[q], 1.0, sigma)                           # sigma -= V^t V
            ops.synchronize()
            S = _view(torch, sigma, M, M).cpu().numpy().T[:m, :m].copy()
        return mean, S

    def residual(self, y, alpha) -> float:
        """||K alpha - y||_2 / ||y||_2 with K regenerated block column by block column (rank r takes columns == r mod world)."""
        torch, ops, nb, n, npad = self.torch, self.ops, self.nb, self.n, self.npad
        D = self.X.shape[1]
        with ops.run():
            Xp = np.zeros((npad, D)); Xp[:n] = self.X
            Xp = np.asfortranarray(Xp)
            Xd = Mat(ops.upload(Xp.T), 0, npad)
            ap = np.zeros(npad); ap[:n] = alpha
            a = ops.upload(ap)
            Kb = Mat(ops.alloc(npad * nb), 0, npad)
            r = ops.alloc(npad, zero=True)
            for j in range(self.rank, self.nt, self.world):
                ops.cov_cross(Xd, npad, Xd.at(j * nb, 0), nb, D, self.theta, Kb)
                ops.gemv(False, npad, nb, 1.0, Kb, Vec(a, j * nb), 1.0, Vec(r, 0))
            self._allreduce(r)
            ops.synchronize()
            rh = r.cpu().numpy()[:n]
        rh = rh + self.sn2 * np.asarray(alpha) - np.asarray(y)
        return float(np.linalg.norm(rh) / np.linalg.norm(y))

"""CPU: the N>1 path (batch partition over ranks + host-side result gather) with world_size 2 on gloo.
The local evaluator is a stand-in closure injected by the test -- the product's evaluator is the CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from gp_algos_b200.batched import shard_bounds, sharded_map


def test_shard_bounds_partition_properties():
    for B in (0, 1, 7, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def local_eval(lo, hi):  # deterministic function of the problem index
        calls.append((lo, hi))
        idx = np.arange(lo, hi, dtype=np.float64)
        return idx ** 2, np.stack([idx, -idx], axis=1)

    a, b = sharded_map(B, local_eval)
    q.put((rank, calls, a, b))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_sharded_map_world_size_2(B):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(B, dtype=np.float64)
    seen = []
    for rank, calls, a, b in out:
        assert np.array_equal(a, idx ** 2) and np.array_equal(b, np.stack([idx, -idx], axis=1))  # same, ordered, on every rank
        assert len(calls) == 1 and calls[0] == shard_bounds(B, rank, 2)
        seen.append(calls[0])
    assert sorted(seen) == [shard_bounds(B, 0, 2), shard_bounds(B, 1, 2)]  # disjoint cover, no problem evaluated twice


def test_sharded_map_without_process_group_runs_everything_locally():
    a, = sharded_map(6, lambda lo, hi: (np.arange(lo, hi),))
    assert np.array_equal(a, np.arange(6))


def test_stack_problems_takes_a_breeze_layout_stack_without_copying():
    """A (B, n, D) view over a (B, D, n) C-order block IS a stack of column-major n x D matrices (Breeze layout): the batched
    entry points must hand it to the C ABI as it is (bench.py's end-to-end leg passes pinned memory this way)."""
    from gp_algos_b200.batched import _stack_problems
    base = np.arange(4 * 3 * 5, dtype=np.float64).reshape(4, 3, 5)          # (B, D, n)
    X = base.transpose(0, 2, 1)                                              # (B, n, D) view
    buf, strideX, n, D = _stack_problems(X, 4)
    assert (strideX, n, D) == (15, 5, 3) and np.shares_memory(buf, base) and buf.flags.c_contiguous
    Xc = np.ascontiguousarray(X)                                             # ordinary C-order stack: one transposing copy
    buf2, _, _, _ = _stack_problems(Xc, 4)
    assert np.array_equal(buf2, base) and not np.shares_memory(buf2, Xc)
    shared, s0, _, _ = _stack_problems(Xc[0], 4)                             # one shared (n, D) matrix: stride 0
    assert s0 == 0 and shared.flags.f_contiguous


def test_gp_ukf_host_side_requirements_without_a_device():
    import gp_algos_b200 as gp
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(1.0, [1.0, 1.0], 0.1))
    ukf = gp.GPUnscentedKalmanFilter(None, gp.GpPredictor(kf))
    with pytest.raises(ValueError):                                          # require: the GP state-space model is learned first
        ukf.filter_many([np.zeros((2, 5))], [np.zeros(2)], [np.eye(2)])
    p = gp.UnscentedTransformParams()
    assert (p.alpha, p.beta, p.kappa) == (1.0, 0.0, 2.0)                     # UnscentedKalmanFilter.scala:186 defaults
    assert gp.UnscentedTransformParams.fromVector([0.5, 2.0, 1.0]) == gp.UnscentedTransformParams(0.5, 2.0, 1.0)

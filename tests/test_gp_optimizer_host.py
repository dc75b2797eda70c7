"""CPU: host logic of the GP-UCB mirror that needs no device -- the facts of the reference's
src/test/scala/gp/optimization/GPOptimizerTest.scala:35-45 (initial grid) and the argument requirements of GPOptimizer.scala:38."""
import numpy as np
import pytest

import gp_algos_b200 as gp


def _optimizer(seed=0):
    kf = gp.GaussianRbfKernel(gp.GaussianRbfParams(1.0, [1.0], 0.1))
    return gp.GPOptimizer(gp.GpPredictor(kf), None, gp.BreezeLbfgsOptimizer(maxIter=5), seed=seed)


def test_initial_grid_is_bounded_and_has_three_points_per_dimension():
    opt = _optimizer()
    grid = opt.prepareGrid([(-6, 6)])                                   # GPOInput(ranges = IndexedSeq(-6 to 6), ...)
    assert grid.shape == (3, 1)                                         # :38 `grid.rows == 3`
    assert np.all((grid[:, 0] < 6) & (grid[:, 0] > -6))                 # :39-43
    grid2 = opt.prepareGrid([(-5, 5), (0, 2)])                          # 3 * dim rows (GPOptimizer.scala:136)
    assert grid2.shape == (6, 2) and np.all((grid2[:, 1] > 0) & (grid2[:, 1] < 2))
    with pytest.raises(ValueError):                                     # require(lower < upper), GPOptimizer.scala:150
        opt.prepareGrid([(2, 2)])
    vals = opt.evaluateGridPoints(grid2, lambda p: float(p[0] + 10 * p[1]))   # :128-132
    assert np.array_equal(vals, grid2[:, 0] + 10 * grid2[:, 1])


def test_requirements_of_maximize():
    opt = _optimizer()
    with pytest.raises(ValueError):                                     # require(c >= 1 && m >= 1), GPOptimizer.scala:38
        opt.maximize(lambda p: 0.0, gp.GPOInput([(-1, 1)], 0, 1, 2.0))
    with pytest.raises(ValueError):
        opt.maximize(lambda p: 0.0, gp.GPOInput([(-1, 1)], 1, 0, 2.0))

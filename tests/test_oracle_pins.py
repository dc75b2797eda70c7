"""Pin the oracle against every fact the reference's own tests assert for the hot path
(SURVEY.md 8(c)):  src/test/scala/utils/KernelRequisitesTest.scala:18-49 and
src/test/scala/utils/MatrixUtilsTest.scala:27-114."""
import numpy as np
import pytest

from oracle import gp_oracle as orc

LOWER = np.array([[0.3, 0., 0.], [0.2, 0.3, 0.], [0.1, 0.99, 0.11]])
UPPER = np.array([[0.4, 0.1, 0.9], [0., 0.2, 0.89], [0., 0., .5]])
INPUT = np.array([[2.4, 1.3, 1.9], [2.1, 0.99, 3.1], [1.89, 2.01, 4.]])
EPS = 0.001  # MatrixUtilsTest.scala:22


def test_packing_to_dense_vector():  # KernelRequisitesTest.scala:20-23
    th = orc.pack_theta(1., np.ones(5), 0.)
    assert np.array_equal(th, np.array([1., 1., 1., 1., 1., 1., 0.]))


def test_get_at_position_one_based():  # KernelRequisitesTest.scala:25-35
    th = orc.pack_theta(1., [5., 2., 3.], 0.)
    assert [orc.get_at_position(th, i) for i in (1, 2, 3, 4, 5)] == [1., 5., 2., 3., 0.]
    with pytest.raises(LookupError):
        orc.get_at_position(th, 6)


def test_forward_solve_lower_vec():  # MatrixUtilsTest.scala:29-36
    rhs = np.array([3., 2., 1.])
    sol = np.linalg.solve(LOWER, rhs)
    assert np.all(np.abs(orc.lit_forward_solve(LOWER, rhs) - sol) < EPS)


def test_back_solve_upper_vec():  # MatrixUtilsTest.scala:38-44
    rhs = np.array([7., 3., 4.])
    sol = np.linalg.solve(UPPER, rhs)
    assert np.all(np.abs(orc.lit_back_solve(UPPER, rhs) - sol) < EPS)


def test_solves_matrix_rhs():  # MatrixUtilsTest.scala:46-63
    rhs = np.array([[0.4, 0.9], [0.8, 0.3], [0.7, 0.4]])
    assert np.all(np.abs(orc.lit_forward_solve(LOWER, rhs) - np.linalg.solve(LOWER, rhs)) < EPS)
    assert np.all(np.abs(orc.lit_back_solve(UPPER, rhs) - np.linalg.solve(UPPER, rhs)) < EPS)


def test_kernel_matrix_unit_diagonal_and_pd():  # MatrixUtilsTest.scala:90-102
    th = orc.pack_theta(1., [1., 1., 1.], 0.)
    for K in (orc.lit_build_kernel_matrix(INPUT, th), orc.fast_build_kernel_matrix(INPUT, th)):
        assert K.shape == (3, 3)
        assert all(K[i, i] == 1.0 for i in range(3))  # exact, as the reference asserts
        orc.lit_cholesky(K)  # "does not throw"


def test_inv_triangular_identity():  # MatrixUtilsTest.scala:104-114
    th = orc.pack_theta(1., [1., 1., 1.], 0.)
    K = orc.lit_build_kernel_matrix(INPUT, th)
    L = orc.lit_cholesky(K)
    Li = orc.lit_inv_triangular(L, is_upper=False)
    assert np.all(np.abs(Li.T @ Li - np.linalg.inv(K)) < EPS)


def test_cholesky_error_behaviour():  # Breeze cholesky: not symmetric / not PD
    with pytest.raises(orc.NotSymmetric):
        orc.lit_cholesky(np.array([[1., 2.], [3., 4.]]))
    with pytest.raises(orc.NotPositiveDefinite) as e:
        orc.lit_cholesky(np.array([[1., 2.], [2., 1.]]))
    assert e.value.minor == 2

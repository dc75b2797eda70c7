// GpuMatrixUtils.scala -- drop-in for utils.MatrixUtils (utils/MatrixUtils.scala:17-133) plus Breeze `cholesky`
// (GpPredictor.scala:120, EpParameterEstimator.scala:58): same names, argument and result types; every O(n^2)/O(n^3) body is one
// libgpk call.  Kernels that libgpk does not lower (anything but GaussianRbfKernel / Co2Kernel, e.g. PMKKernel) fall through to
// the original Scala code.  Uncompiled here (no JVM in the build image) -- see GpkLib.scala.
package gpk

import breeze.linalg.{DenseMatrix, DenseVector}
import utils.KernelRequisites._

object GpuMatrixUtils {
  import Gpk.{check, handle, lib, withFamily}

  /** column-major copy with offset 0 and majorStride == rows (what the ABI's (pointer, ld) pair describes) */
  private def canon(m: DenseMatrix[Double]): DenseMatrix[Double] =
    if (m.offset == 0 && !m.isTranspose && m.majorStride == m.rows) m else m.copy

  /** (kernel family, theta in the ABI's order) of a lowered kernel; None = stay on the JVM */
  def lowered(k: KernelFunc): Option[(Int, Array[Double])] = k match {
    case g: GaussianRbfKernel                       => Some((0, g.rbfParams.toDenseVector.toArray))
    case c: gp.regression.Co2Prediction.Co2Kernel   => Some((1, c.hyperParams.toDenseVector.toArray))
    case _                                          => None
  }

  // utils/MatrixUtils.scala:57-70
  def buildKernelMatrix(kernelFun: KernelFunc, data: DenseMatrix[Double]): kernelMatrixType = lowered(kernelFun) match {
    case Some((family, theta)) =>
      val x = canon(data); val out = DenseMatrix.zeros[Double](x.rows, x.rows)
      withFamily(family) { check(lib.gpk_cov_se_ard(handle, x.data, x.rows, x.cols, x.rows, theta, out.data, x.rows)) }
      out
    case None => utils.MatrixUtils.buildKernelMatrix(kernelFun, data)
  }

  // utils/MatrixUtils.scala:44-55 (never adds the noise term)
  def buildKernelMatrix(kernelFun: KernelFunc, input1: DenseMatrix[Double], input2: DenseMatrix[Double]): kernelMatrixType =
    lowered(kernelFun) match {
      case Some((family, theta)) =>
        val (a, b) = (canon(input1), canon(input2)); val out = DenseMatrix.zeros[Double](a.rows, b.rows)
        withFamily(family) {
          check(lib.gpk_cov_cross_se_ard(handle, a.data, a.rows, a.rows, b.data, b.rows, b.rows, a.cols, theta, out.data, a.rows))
        }
        out
      case None => utils.MatrixUtils.buildKernelMatrix(kernelFun, input1, input2)
    }

  /** buildMatrixWithFunc(data)(derAfterHyperParam(paramNum)) of GpPredictor.scala:70-75 / MarginalLikelihoodEvaluator.scala:68-79
    * as one call (paramNum is 1-based like the reference). */
  def buildKernelDerMatrix(kernelFun: KernelFunc, data: DenseMatrix[Double], paramNum: Int): kernelMatrixType = lowered(kernelFun) match {
    case Some((family, theta)) =>
      val x = canon(data); val out = DenseMatrix.zeros[Double](x.rows, x.rows)
      withFamily(family) { check(lib.gpk_cov_deriv_se_ard(handle, paramNum, x.data, x.rows, x.cols, x.rows, theta, out.data, x.rows)) }
      out
    case None =>
      utils.MatrixUtils.buildMatrixWithFunc(data) { (v1, v2, same) => kernelFun.derAfterHyperParam(paramNum)(v1, v2, same) }
  }

  /** breeze.linalg.cholesky: lower factor, strict upper zeroed; MatrixNotSymmetricException / NotConvergedException as Breeze */
  def cholesky(a: DenseMatrix[Double]): DenseMatrix[Double] = {
    require(a.rows == a.cols)
    val x = canon(a); val out = DenseMatrix.zeros[Double](x.rows, x.rows)
    check(lib.gpk_potrf_lower(handle, x.data, x.rows, x.rows, out.data, x.rows, 1))
    out
  }

  // utils/MatrixUtils.scala:17-35: the `.t` view of GpPredictor.scala:122 / GpClassifier.scala:36 is passed as a flag, not copied
  private def solve(upper: Int, t: DenseMatrix[Double], b: DenseMatrix[Double]): DenseMatrix[Double] = {
    require(t.rows == t.cols)
    val stored = if (t.isTranspose) canon(t.t) else canon(t)
    val bb = canon(b); val x = DenseMatrix.zeros[Double](b.rows, b.cols)
    check(lib.gpk_trsm(handle, upper, if (t.isTranspose) 1 else 0, stored.data, t.rows, t.rows, bb.data, b.cols, b.rows, x.data, b.rows))
    x
  }
  def forwardSolve(L: DenseMatrix[Double], b: DenseMatrix[Double]): DenseMatrix[Double] = solve(0, L, b)
  def backSolve(R: DenseMatrix[Double], b: DenseMatrix[Double]): DenseMatrix[Double] = solve(1, R, b)
  def forwardSolve(L: DenseMatrix[Double], b: DenseVector[Double]): DenseVector[Double] = solve(0, L, b.toDenseMatrix.t)(::, 0)
  def backSolve(R: DenseMatrix[Double], b: DenseVector[Double]): DenseVector[Double] = solve(1, R, b.toDenseMatrix.t)(::, 0)

  // utils/MatrixUtils.scala:106-113: dense n x n inverse of a triangular matrix
  def invTriangular(matrix: DenseMatrix[Double], isUpper: Boolean): DenseMatrix[Double] = {
    require(matrix.rows == matrix.cols)
    val stored = if (matrix.isTranspose) canon(matrix.t) else canon(matrix)       // (T^t)^-1 = (T^-1)^t: invert what is stored
    val storedUpper = isUpper != matrix.isTranspose
    val out = DenseMatrix.zeros[Double](matrix.rows, matrix.rows)
    check(lib.gpk_trtri(handle, if (storedUpper) 1 else 0, stored.data, matrix.rows, matrix.rows, out.data, matrix.rows))
    if (matrix.isTranspose) out.t.copy else out
  }

  def cloneCols(vec: DenseVector[Double], colNum: Int): DenseMatrix[Double] = utils.MatrixUtils.cloneCols(vec, colNum)   // O(n m), host
}

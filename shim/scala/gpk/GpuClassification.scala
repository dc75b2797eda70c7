// GpuClassification.scala -- gp.classification on libgpk: EpParameterEstimator.estimateSiteParams (EpParameterEstimator.scala:29-69
// incl. epMarginalLikelihood :71-96 as compiled), GpClassifier.classify (GpClassifier.scala:24-47) and
// MarginalLikelihoodEvaluator.logLikelihood / logLikelihoodDerivativesAfterHyperParams (MarginalLikelihoodEvaluator.scala:33-66).
// Uncompiled here (no JVM in the build image) -- see GpkLib.scala.
package gpk

import breeze.linalg.{DenseMatrix, DenseVector}
import com.sun.jna.ptr.{DoubleByReference, IntByReference}
import gp.classification.EpParameterEstimator
import gp.classification.EpParameterEstimator.{AvgBasedStopCriterion, SiteParams, stopCriterionFunc}
import gp.classification.GpClassifier.{AfterEstimationClassifierInput, classifyOut, learnParams}
import utils.KernelRequisites.{KernelFunc, KernelFuncHyperParams}

/** How the shim hands a stop criterion to the device: the shipped AvgBasedStopCriterion(eps) (spring-context.xml:53-55) or a fixed
  * number of sweeps.  Any other stopCriterionFunc is an arbitrary JVM closure over site parameters and keeps the original class. */
sealed trait DeviceStop
case class AvgEps(eps: Double) extends DeviceStop
case class FixedSweeps(sweeps: Int) extends DeviceStop

class GpuEpParameterEstimator(kernelMatrix: DenseMatrix[Double], targets: DenseVector[Int], stop: DeviceStop,
                              keepLinebreakQuirk: Boolean = true, maxSweeps: Int = 100) {
  import Gpk.{check, handle, lib}
  require(kernelMatrix.rows == targets.length)                                          // EpParameterEstimator.scala:20

  var sweeps: Int = 0
  /** (SiteParams(tau, ni, Some(logZ)), L) exactly like EpParameterEstimator.estimateSiteParams */
  def estimateSiteParams: (SiteParams, DenseMatrix[Double]) = {
    val n = kernelMatrix.rows
    val k = if (kernelMatrix.offset == 0 && !kernelMatrix.isTranspose && kernelMatrix.majorStride == n) kernelMatrix else kernelMatrix.copy
    val (eps, fixed) = stop match { case AvgEps(e) => (e, 0); case FixedSweeps(s) => (0.0, s) }
    val tau = new Array[Double](n); val nu = new Array[Double](n); val l = DenseMatrix.zeros[Double](n, n)
    val logZ = new DoubleByReference(); val nsweeps = new IntByReference()
    check(lib.gpk_ep_fit(handle, k.data, n, n, targets.toArray, eps, fixed, maxSweeps, if (keepLinebreakQuirk) 1 else 0,
                         tau, nu, null, l.data, n, null, null, logZ, nsweeps))
    sweeps = nsweeps.getValue
    (SiteParams(tauSiteParams = DenseVector(tau), niSiteParams = DenseVector(nu), marginalLogLikelihood = Some(logZ.getValue)), l)
  }
}

class GpuGpClassifier(stop: DeviceStop) {
  import Gpk.{check, handle, lib}

  def trainClassifier(kernelMatrix: DenseMatrix[Double], targets: DenseVector[Int]): learnParams =
    new GpuEpParameterEstimator(kernelMatrix, targets, stop).estimateSiteParams

  // GpClassifier.scala:24-47; only the diagonal of testKernelMatrix is read (:44)
  def classify(input: AfterEstimationClassifierInput): classifyOut = {
    val (siteParams, lowerTriangular) = input.learnParams.getOrElse(trainClassifier(input.trainKernelMatrix, input.targets))
    val (k, ks, l) = (input.trainKernelMatrix.copy, input.testTrainKernelMatrix.copy, lowerTriangular.copy)
    val (n, m) = (k.rows, ks.rows)
    val kss = Array.tabulate(m)(i => input.testKernelMatrix(i, i))
    val prob = new Array[Double](m)
    check(lib.gpk_ep_classify(handle, k.data, n, n, ks.data, m, m, kss, siteParams.tauSiteParams.toArray, siteParams.niSiteParams.toArray,
                              l.data, n, prob, null, null))
    DenseVector(prob)
  }
}

/** MarginalLikelihoodEvaluator(stopCriterion, kernelFunc): logLikelihood (:33-45) is ONE device call -- K is built from X on the
  * device, EP runs to the stop rule and the gradient of :47-66 (as compiled: rMatrix = b b^t) follows; nothing n x n crosses PCIe. */
class GpuMarginalLikelihoodEvaluator(stop: DeviceStop, kernelFunc: KernelFunc, keepLinebreakQuirk: Boolean = true, maxSweeps: Int = 100) {
  import Gpk.{check, handle, lib, withFamily}

  def logLikelihood(trainInput: DenseMatrix[Double], targets: DenseVector[Int], hyperParams: KernelFuncHyperParams): (Double, DenseVector[Double]) = {
    val (family, _) = GpuMatrixUtils.lowered(kernelFunc).getOrElse(throw new IllegalArgumentException("kernel is not lowered to the device"))
    val x = trainInput.copy; val theta = hyperParams.toDenseVector.toArray
    val (eps, fixed) = stop match { case AvgEps(e) => (e, 0); case FixedSweeps(s) => (0.0, s) }
    val logZ = new DoubleByReference(); val grad = new Array[Double](theta.length); val nsweeps = new IntByReference()
    withFamily(family) {
      check(lib.gpk_ep_nll_grad(handle, x.data, x.rows, x.cols, x.rows, theta, targets.toArray, eps, fixed, maxSweeps,
                                if (keepLinebreakQuirk) 1 else 0, theta.length, logZ, grad, null, null, nsweeps))
    }
    (logZ.getValue, DenseVector(grad))
  }

  def logLikelihoodWithoutGrad(trainInput: DenseMatrix[Double], targets: DenseVector[Int], hyperParams: KernelFuncHyperParams): Double = {
    val k = GpuMatrixUtils.buildKernelMatrix(kernelFunc.changeHyperParams(hyperParams.toDenseVector), trainInput)
    new GpuEpParameterEstimator(k, targets, stop, keepLinebreakQuirk, maxSweeps).estimateSiteParams._1.marginalLogLikelihood.get
  }
}

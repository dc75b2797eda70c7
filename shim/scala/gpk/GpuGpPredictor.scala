// GpuGpPredictor.scala -- gp.regression.GpPredictor (GpPredictor.scala:15-149) with its numeric bodies on libgpk.  Same
// constructor and method signatures, so the Spring beans (spring-context.xml:25-51) and the callers listed in SURVEY.md 8(b)
// (GPOptimizer, GPUnscentedKalmanFilter, Co2PredictionExecutor, MasterThesisRelatedTasks, the tests) only swap the class name.
// obtainOptimalHyperParams / predictWithParamsOptimization / preComputeComponentsWithHpOptimization are inherited unchanged:
// Breeze's L-BFGS keeps running on the JVM and calls the overridden logLikelihoodWithDerivatives.
// Uncompiled here (no JVM in the build image) -- see GpkLib.scala.
package gpk

import breeze.linalg.{DenseMatrix, DenseVector}
import com.sun.jna.Pointer
import com.sun.jna.ptr.{DoubleByReference, PointerByReference}
import gp.regression.GpPredictor
import gp.regression.GpPredictor.{PredictionInput, PredictionTrainingInput, afterLearningComponents}
import utils.KernelRequisites.{KernelFunc, KernelFuncHyperParams}
import utils.StatsUtils.GaussianDistribution

class GpuGpPredictor(kernelFunc: KernelFunc) extends GpPredictor(kernelFunc) {
  import Gpk.{check, handle, lib, withFamily}
  import GpuMatrixUtils.lowered

  private def canon(m: DenseMatrix[Double]) = if (m.offset == 0 && !m.isTranspose && m.majorStride == m.rows) m else m.copy
  private def opt(s: Option[Double]) = (if (s.isDefined) 1 else 0, s.getOrElse(0.0))
  /** family of this predictor's kernel and theta of `hp` in the ABI order; None = kernel not lowered */
  private def familyAndTheta(hp: KernelFuncHyperParams) = lowered(kernelFunc).map { case (family, _) => (family, hp.toDenseVector.toArray) }

  // GpPredictor.scala:24-43
  override def predict(input: PredictionInput, hyperParams: KernelFuncHyperParams = kernelFunc.hyperParams): (GaussianDistribution, Double) =
    familyAndTheta(hyperParams) match {
      case Some((family, theta)) =>
        require(input.trainingData.rows == input.targets.length,
          "Number of objects in training data matrix should be equal to targets vector length")
        val (x, xs) = (canon(input.trainingData), canon(input.testData))
        val (hasS, s) = opt(input.sigmaNoise)
        val mean = new Array[Double](xs.rows); val sigma = DenseMatrix.zeros[Double](xs.rows, xs.rows); val ll = new DoubleByReference()
        withFamily(family) {
          check(lib.gpk_gp_predict(handle, x.data, x.rows, x.cols, x.rows, input.targets.toArray, xs.data, xs.rows, xs.rows, theta,
                                   hasS, s, mean, sigma.data, xs.rows, ll))
        }
        (GaussianDistribution(mean = DenseVector(mean), sigma = sigma), ll.getValue)
      case None => super.predict(input, hyperParams)
    }

  // GpPredictor.scala:50-58 (the caller hands over L and alphaVec; V = L^-1 K*^t is returned like the reference does)
  override def computePosterior(trainingData: DenseMatrix[Double], testData: DenseMatrix[Double], l: DenseMatrix[Double],
                                alphaVec: DenseVector[Double], kernelFunc: KernelFunc): (GaussianDistribution, DenseMatrix[Double]) =
    lowered(kernelFunc) match {
      case Some((family, theta)) =>
        val (x, xs, lc) = (canon(trainingData), canon(testData), canon(l))
        val model = new PointerByReference()
        withFamily(family) {
          check(lib.gpk_gp_model_from_factor(handle, x.data, x.rows, x.cols, x.rows, lc.data, lc.rows, alphaVec.toArray, theta, model))
        }
        try {
          val mean = new Array[Double](xs.rows); val sigma = DenseMatrix.zeros[Double](xs.rows, xs.rows)
          val v = DenseMatrix.zeros[Double](x.rows, xs.rows)
          check(lib.gpk_gp_model_predict(handle, model.getValue, xs.data, xs.rows, xs.rows, 1, mean, sigma.data, xs.rows, v.data, x.rows))
          (GaussianDistribution(mean = DenseVector(mean), sigma = sigma), v)
        } finally lib.gpk_gp_model_destroy(handle, model.getValue)
      case None => super.computePosterior(trainingData, testData, l, alphaVec, kernelFunc)
    }

  // GpPredictor.scala:60-80 -- the metric's call: K, L, L^-1, K^-1 never leave the device, dK/dtheta is never materialised
  override def logLikelihoodWithDerivatives(input: PredictionTrainingInput, hyperParams: KernelFuncHyperParams,
                                            optimizedParamsNum: Int): (Double, DenseVector[Double]) =
    familyAndTheta(hyperParams) match {
      case Some((family, theta)) =>
        require(input.trainingData.rows == input.targets.length,
          "Number of objects in training data matrix should be equal to targets vector length")
        val x = canon(input.trainingData); val (hasS, s) = opt(input.sigmaNoise)
        val ll = new DoubleByReference(); val g = new Array[Double](optimizedParamsNum)
        withFamily(family) {
          check(lib.gpk_gp_nll_grad(handle, x.data, x.rows, x.cols, x.rows, input.targets.toArray, theta, hasS, s, optimizedParamsNum, ll, g))
        }
        (ll.getValue, DenseVector(g))
      case None => super.logLikelihoodWithDerivatives(input, hyperParams, optimizedParamsNum)
    }

  // GpPredictor.scala:104-124: (L, alphaVec, Option[noise * I]); sigmaNoise is added UN-squared (:116-117)
  override def preComputeComponents(trainingData: DenseMatrix[Double], hyperParams: KernelFuncHyperParams, sigmaNoise: Option[Double],
                                    targets: DenseVector[Double]): afterLearningComponents =
    familyAndTheta(hyperParams) match {
      case Some((family, theta)) =>
        require(trainingData.rows == targets.length,
          "Number of objects in training data matrix should be equal to targets vector length")
        val x = canon(trainingData); val n = x.rows; val (hasS, s) = opt(sigmaNoise)
        val l = DenseMatrix.zeros[Double](n, n); val alpha = new Array[Double](n); val ll = new DoubleByReference()
        withFamily(family) {
          check(lib.gpk_gp_fit(handle, x.data, n, x.cols, n, targets.toArray, theta, hasS, s, l.data, n, alpha, ll))
        }
        (l, DenseVector(alpha), sigmaNoise.map(sn => DenseMatrix.eye[Double](n) :* sn))
      case None => super.preComputeComponents(trainingData, hyperParams, sigmaNoise, targets)
    }

  // ---- beyond the reference's surface: the resident model behind the GP-UKF / GP-UCB call pattern -----------------------------
  /** Fit once, keep X, L^-1 and alpha on the device (GPUnscentedKalmanFilter.scala:123-136 fits one GP per output dimension and
    * then asks each for (2d+1) m = 1 posteriors per time step, :77-88,138-147). */
  def fitResident(trainingData: DenseMatrix[Double], sigmaNoise: Option[Double], targets: DenseVector[Double],
                  hyperParams: KernelFuncHyperParams = kernelFunc.hyperParams): GpuFittedGp = {
    val (family, theta) = familyAndTheta(hyperParams).getOrElse(throw new IllegalArgumentException("kernel is not lowered to the device"))
    val x = canon(trainingData); val (hasS, s) = opt(sigmaNoise)
    val model = new PointerByReference(); val ll = new DoubleByReference()
    withFamily(family) { check(lib.gpk_gp_model_fit(handle, x.data, x.rows, x.cols, x.rows, targets.toArray, theta, hasS, s, model, ll)) }
    new GpuFittedGp(model.getValue, x.cols, ll.getValue, sigmaNoise)
  }
}

/** A device-resident fitted GP (opaque gpk_model token).  close() frees the device memory. */
class GpuFittedGp(private val model: Pointer, val dim: Int, val logLikelihood: Double, sigmaNoise: Option[Double]) {
  import Gpk.{check, handle, lib}
  def size: Int = lib.gpk_gp_model_size(handle, model)
  def alphaVec: DenseVector[Double] = { val a = new Array[Double](size); check(lib.gpk_gp_model_get_alpha(handle, model, a)); DenseVector(a) }
  /** computePosterior(trainingData, testData, l, alphaVec) of GpPredictor.scala:45-58 without re-sending X, L, alpha */
  def computePosterior(testData: DenseMatrix[Double]): GaussianDistribution = {
    val xs = testData.copy; val mean = new Array[Double](xs.rows); val sigma = DenseMatrix.zeros[Double](xs.rows, xs.rows)
    check(lib.gpk_gp_model_predict(handle, model, xs.data, xs.rows, xs.rows, 1, mean, sigma.data, xs.rows, null, 0))
    GaussianDistribution(mean = DenseVector(mean), sigma = sigma)
  }
  /** GPOptimizer.scala:82-109: (ucb, d ucb / d x) of maximizeUCB's objective at the rows of `points` */
  def ucbWithGradient(points: DenseMatrix[Double], kParam: Double): (DenseVector[Double], DenseMatrix[Double]) = {
    val xs = points.copy; val ucb = new Array[Double](xs.rows); val grad = DenseMatrix.zeros[Double](xs.rows, xs.cols)
    check(lib.gpk_gp_model_ucb(handle, model, xs.data, xs.rows, xs.rows, kParam, ucb, grad.data, xs.rows, null, null))
    (DenseVector(ucb), grad)
  }
  /** GPOptimizer.scala:64-71: the point each outer iteration adds -- a bordered O(n^2) update instead of the O(n^3) refit of :51 */
  def append(x: DenseVector[Double], y: Double): Double = {
    val d = new DoubleByReference()
    check(lib.gpk_gp_model_append(handle, model, x.toArray, y, if (sigmaNoise.isDefined) 1 else 0, sigmaNoise.getOrElse(0.0), d))
    d.getValue
  }
  def close(): Unit = lib.gpk_gp_model_destroy(handle, model)
  private[gpk] def token: Pointer = model
}

object GpuFittedGp {
  /** GPUnscentedKalmanFilter.scala:77-90,138-147: posterior mean AND variance of every model (one GP per state / observation
    * dimension) at the same sigma points, ONE call: mean(j, i), variance(j, i) of model j at row i of `points`. */
  def meansAndVariances(models: Seq[GpuFittedGp], points: DenseMatrix[Double]): (DenseMatrix[Double], DenseMatrix[Double]) = {
    import Gpk.{check, handle, lib}
    val xs = points.copy; val ms = xs.rows
    val mean = new Array[Double](models.length * ms); val variance = new Array[Double](models.length * ms)
    check(lib.gpk_gp_models_mean_var(handle, models.map(_.token).toArray, models.length, xs.data, ms, ms, mean, variance))
    (new DenseMatrix(ms, models.length, mean).t, new DenseMatrix(ms, models.length, variance).t)
  }

  /** dynamicalsystems/filtering/GPUnscentedKalmanFilter.inferHiddenState (GPUnscentedKalmanFilter.scala:26-34 over
    * UnscentedKalmanFilter.scala:24-80) for B observation series in ONE device call: sysModels = one resident GP per state dimension
    * (trained on z_t -> z_{t+1} - z_t, :105-113), obsModels = one per observation dimension (:115-121); observations(b) is
    * p x T, priors (initMeans(b), initCovs(b)).  Returns per series (hiddenMeans d x T, hiddenCovs, logLikelihood). */
  def filterMany(sysModels: Seq[GpuFittedGp], obsModels: Seq[GpuFittedGp], observations: Seq[DenseMatrix[Double]],
                 initMeans: Seq[DenseVector[Double]], initCovs: Seq[DenseMatrix[Double]], alpha: Double = 1.0, beta: Double = 0.0,
                 kappa: Double = 2.0): Seq[(DenseMatrix[Double], Array[DenseMatrix[Double]], Double)] = {
    import Gpk.{check, handle, lib}
    val (d, p, b, t) = (sysModels.length, obsModels.length, observations.length, observations.head.cols)
    val y = observations.flatMap(_.copy.data).toArray                                  // p x T column-major per series
    val m0 = initMeans.flatMap(_.toArray).toArray
    val c0 = initCovs.flatMap(_.copy.data).toArray
    val means = new Array[Double](b * t * d); val covs = new Array[Double](b * t * d * d); val ll = new Array[Double](b)
    check(lib.gpk_gpukf_filter(handle, sysModels.map(_.token).toArray, d, obsModels.map(_.token).toArray, p, b, t, y, m0, c0,
                               alpha, beta, kappa, 1, means, covs, ll))
    (0 until b).map { s =>
      val hm = new DenseMatrix(d, t, means.slice(s * t * d, (s + 1) * t * d))
      val hc = Array.tabulate(t)(i => new DenseMatrix(d, d, covs.slice((s * t + i) * d * d, (s * t + i + 1) * d * d)))
      (hm, hc, ll(s))
    }
  }
}

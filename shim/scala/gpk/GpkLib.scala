// GpkLib.scala -- JNA binding of libgpk.so (include/gpk.h), one declaration per C symbol the Scala shim uses.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no JVM (no java / scalac / mvn, no jars, no network).  The files under
// shim/scala are the source a maintainer of astroHaoPeng/gp_algos adds next to src/main/scala (Scala 2.10, Breeze 0.8.1, plus
// the dependency net.java.dev.jna:jna:4.x in pom.xml).  The same C entry points are exercised by tests/ through the Python
// ctypes mirror (gp_algos_b200/) and by tests/cpp/host_mirror_test.cpp through include/gpk.hpp.
//
// Breeze DenseMatrix[Double] is column-major `data` + (offset, rows, cols, majorStride, isTranspose): exactly the
// (pointer, rows, cols, ld) convention of the ABI.  The shim canonicalises views with `.copy` (offset 0, majorStride = rows) before a call.
package gpk

import com.sun.jna.{Library, Native, Pointer}
import com.sun.jna.ptr.{DoubleByReference, IntByReference, PointerByReference}

trait GpkLib extends Library {
  // ---- lifetime (gpk.h "lifetime") ----
  def gpk_create(out: PointerByReference, device: Int, stream: Pointer): Int
  def gpk_destroy(h: Pointer): Int
  def gpk_last_error(h: Pointer): String
  def gpk_last_info(h: Pointer): Int
  def gpk_set_kernel_family(h: Pointer, family: Int): Int
  def gpk_get_kernel_family(h: Pointer): Int
  def gpk_theta_length(h: Pointer, d: Int): Int
  def gpk_set_graph_mode(h: Pointer, on: Int): Int
  // ---- utils/MatrixUtils.scala:17-35,44-70,106-113 and breeze cholesky (GpPredictor.scala:120, EpParameterEstimator.scala:58) ----
  def gpk_cov_se_ard(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, theta: Array[Double], k: Array[Double], ldk: Long): Int
  def gpk_cov_cross_se_ard(h: Pointer, x1: Array[Double], m: Int, ldx1: Long, x2: Array[Double], n: Int, ldx2: Long, d: Int,
                           theta: Array[Double], k: Array[Double], ldk: Long): Int
  def gpk_cov_deriv_se_ard(h: Pointer, paramNum: Int, x: Array[Double], n: Int, d: Int, ldx: Long, theta: Array[Double],
                           dk: Array[Double], ldk: Long): Int
  def gpk_potrf_lower(h: Pointer, a: Array[Double], n: Int, lda: Long, l: Array[Double], ldl: Long, checkSymmetric: Int): Int
  def gpk_trsm(h: Pointer, upper: Int, transposed: Int, t: Array[Double], n: Int, ldt: Long, b: Array[Double], nrhs: Int,
               ldb: Long, x: Array[Double], ldx: Long): Int
  def gpk_trtri(h: Pointer, isUpper: Int, t: Array[Double], n: Int, ldt: Long, tinv: Array[Double], ldi: Long): Int
  // ---- gp/regression/GpPredictor.scala:24-149 ----
  def gpk_gp_fit(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, y: Array[Double], theta: Array[Double],
                 hasSigmaNoise: Int, sigmaNoise: Double, l: Array[Double], ldl: Long, alpha: Array[Double], ll: DoubleByReference): Int
  def gpk_gp_nll_grad(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, y: Array[Double], theta: Array[Double],
                      hasSigmaNoise: Int, sigmaNoise: Double, nparams: Int, ll: DoubleByReference, grad: Array[Double]): Int
  def gpk_gp_predict(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, y: Array[Double], xs: Array[Double], ms: Int,
                     ldxs: Long, theta: Array[Double], hasSigmaNoise: Int, sigmaNoise: Double, mean: Array[Double],
                     sigma: Array[Double], lds: Long, ll: DoubleByReference): Int
  def gpk_gp_model_fit(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, y: Array[Double], theta: Array[Double],
                       hasSigmaNoise: Int, sigmaNoise: Double, out: PointerByReference, ll: DoubleByReference): Int
  def gpk_gp_model_from_factor(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, l: Array[Double], ldl: Long,
                               alpha: Array[Double], theta: Array[Double], out: PointerByReference): Int
  def gpk_gp_model_predict(h: Pointer, m: Pointer, xs: Array[Double], ms: Int, ldxs: Long, wantFullCov: Int,
                           mean: Array[Double], sigma: Array[Double], lds: Long, v: Array[Double], ldv: Long): Int
  def gpk_gp_model_get_alpha(h: Pointer, m: Pointer, alpha: Array[Double]): Int
  def gpk_gp_model_append(h: Pointer, m: Pointer, xNew: Array[Double], yNew: Double, hasSigmaNoise: Int, sigmaNoise: Double,
                          llDelta: DoubleByReference): Int
  def gpk_gp_model_size(h: Pointer, m: Pointer): Int
  def gpk_gp_model_destroy(h: Pointer, m: Pointer): Int
  def gpk_gp_model_ucb(h: Pointer, m: Pointer, xs: Array[Double], ms: Int, ldxs: Long, kParam: Double, ucb: Array[Double],
                       grad: Array[Double], ldg: Long, mean: Array[Double], variance: Array[Double]): Int
  def gpk_gp_models_mean(h: Pointer, models: Array[Pointer], nmodels: Int, xs: Array[Double], ms: Int, ldxs: Long, mean: Array[Double]): Int
  def gpk_gp_models_mean_var(h: Pointer, models: Array[Pointer], nmodels: Int, xs: Array[Double], ms: Int, ldxs: Long,
                             mean: Array[Double], variance: Array[Double]): Int
  // GP-UKF filter run, device-resident, B series at once (UnscentedKalmanFilter.scala:24-118 over GPUnscentedKalmanFilter.scala:63-103)
  def gpk_gpukf_filter(h: Pointer, sysModels: Array[Pointer], d: Int, obsModels: Array[Pointer], p: Int, b: Int, t: Int,
                       y: Array[Double], initMean: Array[Double], initCov: Array[Double], alpha: Double, beta: Double, kappa: Double,
                       computeLl: Int, hiddenMeans: Array[Double], hiddenCovs: Array[Double], ll: Array[Double]): Int
  // ---- batched independent GPs (GPUnscentedKalmanFilter.scala:123-136, GPOptimizer.scala:54-61) ----
  def gpk_gp_nll_grad_batched(h: Pointer, b: Int, x: Array[Double], n: Int, d: Int, ldx: Long, strideX: Long, y: Array[Double],
                              thetas: Array[Double], hasSigmaNoise: Int, sigmaNoise: Double, nparams: Int, ll: Array[Double],
                              grad: Array[Double], info: Array[Int]): Int
  def gpk_gp_predict_batched(h: Pointer, b: Int, x: Array[Double], n: Int, d: Int, ldx: Long, strideX: Long, y: Array[Double],
                             thetas: Array[Double], xs: Array[Double], ms: Int, ldxs: Long, strideXs: Long, hasSigmaNoise: Int,
                             sigmaNoise: Double, mean: Array[Double], variance: Array[Double], ll: Array[Double], info: Array[Int]): Int
  // ---- gp/classification (EpParameterEstimator.scala:29-109, GpClassifier.scala:24-47, MarginalLikelihoodEvaluator.scala:18-66) ----
  def gpk_ep_fit(h: Pointer, k: Array[Double], n: Int, ldk: Long, targets: Array[Int], eps: Double, fixedSweeps: Int,
                 maxSweeps: Int, keepLinebreakQuirk: Int, tau: Array[Double], nu: Array[Double], mu: Array[Double],
                 l: Array[Double], ldl: Long, cavTau: Array[Double], cavNu: Array[Double], logZ: DoubleByReference,
                 sweeps: IntByReference): Int
  def gpk_ep_classify(h: Pointer, k: Array[Double], n: Int, ldk: Long, ks: Array[Double], m: Int, ldks: Long, kssDiag: Array[Double],
                      tau: Array[Double], nu: Array[Double], l: Array[Double], ldl: Long, prob: Array[Double],
                      fmean: Array[Double], fvar: Array[Double]): Int
  def gpk_ep_nll_grad(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, theta: Array[Double], targets: Array[Int],
                      eps: Double, fixedSweeps: Int, maxSweeps: Int, keepLinebreakQuirk: Int, nparams: Int, logZ: DoubleByReference,
                      grad: Array[Double], tau: Array[Double], nu: Array[Double], sweeps: IntByReference): Int
  def gpk_ep_grad_from_factor(h: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, theta: Array[Double], k: Array[Double],
                              ldk: Long, tau: Array[Double], nu: Array[Double], l: Array[Double], ldl: Long, nparams: Int,
                              grad: Array[Double]): Int
  // ---- one large GP on every GPU of the node (include/gpk.h "one large GP on all GPUs"; same libgpk.so) ----
  def gpk_mg_create(out: PointerByReference, ndev: Int, devices: Array[Int]): Int
  def gpk_mg_destroy(mg: Pointer): Int
  def gpk_mg_last_error(mg: Pointer): String
  def gpk_mg_potrf_solve(mg: Pointer, x: Array[Double], n: Int, d: Int, ldx: Long, y: Array[Double], theta: Array[Double],
                         hasSigmaNoise: Int, sigmaNoise: Double, alpha: Array[Double], ll: DoubleByReference, info: IntByReference): Int
}

object Gpk {
  val lib: GpkLib = Native.loadLibrary(System.getProperty("gpk.lib", "gpk"), classOf[GpkLib]).asInstanceOf[GpkLib]

  /** One handle per JVM, like the reference's Spring singletons (spring-context.xml:33-51).  A handle is not re-entrant:
    * the reference's callers are single-threaded; a multi-threaded host creates one handle per thread with `newHandle`. */
  lazy val handle: Pointer = newHandle(Integer.getInteger("gpk.device", 0))

  def newHandle(device: Int): Pointer = {
    val p = new PointerByReference()
    val rc = lib.gpk_create(p, device, null)
    if (rc != 0) throw new RuntimeException("gpk_create(device=" + device + ") failed with status " + rc +
                                            ": libgpk needs a CUDA device, there is no CPU fallback")
    p.getValue
  }

  /** status -> the exception class the reference would have thrown at the same place (SURVEY.md 8(b) "errors") */
  def check(rc: Int): Unit = check(handle, rc)
  def check(h: Pointer, rc: Int): Unit = rc match {
    case 0  =>
    case -1 =>                                                                                   // require(...) / MatchError
      val msg = lib.gpk_last_error(h)
      if (msg.startsWith("scala.MatchError")) throw new MatchError(msg) else throw new IllegalArgumentException(msg)
    case -2 => throw new breeze.linalg.MatrixNotSymmetricException
    case -3 => throw new breeze.linalg.NotConvergedException(breeze.linalg.NotConvergedException.Iterations,
                                                             "cholesky: leading minor " + lib.gpk_last_info(h) + " is not positive")
    case _  => throw new RuntimeException(lib.gpk_last_error(h))                                  // CUDA error / out of memory
  }

  /** run `body` with the handle reading (D, theta) as kernel family `family` (1 = Co2Kernel), then put the SE family back */
  def withFamily[T](family: Int)(body: => T): T = {
    check(lib.gpk_set_kernel_family(handle, family))
    try body finally lib.gpk_set_kernel_family(handle, 0)
  }
}

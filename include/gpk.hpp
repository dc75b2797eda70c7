// gpk.hpp -- C++ host-side mirror of the reference's Scala interface for the dense-GP hot path, over the C ABI of gpk.h.
//
// The reference (astroHaoPeng/gp_algos) is Scala on the JVM; no JVM exists in the build image, so the host side above the
// C ABI exists twice: the Python mirror (gp_algos_b200/*.py, used by tests/ and bench.py) and this header-only C++17 mirror
// for compiled callers.  Same names, argument meaning and error behaviour as the Scala objects; every method is a thin
// marshalling layer around one gpk_* call -- no arithmetic beyond O(D) scalar kernel evaluations happens here, and there is
// no CPU fallback (construction of the Handle throws when no CUDA device is usable).
//
// Reference paths are relative to /root/reference/src/main/scala/.
#pragma once
#include <cmath>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "gpk.h"

namespace gpk {

// ---- Breeze-shaped containers: DenseMatrix is column-major (element (r,c) at data[r + c*rows]) --------------------------
using DenseVector = std::vector<double>;
struct DenseMatrix {
    int rows = 0, cols = 0;
    std::vector<double> data;
    DenseMatrix() = default;
    DenseMatrix(int r, int c, double v = 0.0) : rows(r), cols(c), data((size_t)r * c, v) {}
    double& operator()(int r, int c) { return data[(size_t)r + (size_t)c * rows]; }
    double operator()(int r, int c) const { return data[(size_t)r + (size_t)c * rows]; }
    static DenseMatrix eye(int n) { DenseMatrix m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
    DenseMatrix t() const { DenseMatrix o(cols, rows); for (int c = 0; c < cols; ++c) for (int r = 0; r < rows; ++r) o(c, r) = (*this)(r, c); return o; }
};

// ---- errors: the exception classes the reference would have thrown (SURVEY.md 8(b) "errors") ---------------------------
struct IllegalArgumentException : std::invalid_argument { using std::invalid_argument::invalid_argument; };   // require(...)
struct MatrixNotSymmetricException : std::runtime_error { using std::runtime_error::runtime_error; };         // breeze cholesky
struct NotConvergedException : std::runtime_error {                                                            // breeze cholesky
    int minor;
    NotConvergedException(const std::string& m, int minor_) : std::runtime_error(m), minor(minor_) {}
};
struct MatchError : std::out_of_range { using std::out_of_range::out_of_range; };                             // scala.MatchError
struct GpkError : std::runtime_error { using std::runtime_error::runtime_error; };                            // CUDA / memory

class Handle {
public:
    explicit Handle(int device = 0) {
        if (gpk_create(&h_, device, nullptr) != GPK_OK)
            throw GpkError("gpk_create failed: no usable CUDA device (libgpk has no CPU fallback)");
    }
    ~Handle() { if (h_) gpk_destroy(h_); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    gpk_handle get() const { return h_; }
    void check(int rc) const {
        if (rc == GPK_OK) return;
        const std::string msg = gpk_last_error(h_);
        switch (rc) {
            case GPK_EINVAL: throw IllegalArgumentException("requirement failed: " + msg);
            case GPK_ENOTSYM: throw MatrixNotSymmetricException(msg);
            case GPK_ENOTPD: throw NotConvergedException(msg, gpk_last_info(h_));
            default: throw GpkError(msg);
        }
    }
    static Handle& instance() { static Handle h(0); return h; }   // the Scala objects are process-wide singletons too
private:
    gpk_handle h_ = nullptr;
};

// ---- utils/KernelRequisites.scala ------------------------------------------------------------------------------------
struct GaussianRbfParams {   // KernelRequisites.scala:39-60; signalVar / noiseVar are std-dev-like (squared inside the kernel)
    double signalVar;
    DenseVector lengthScales;
    double noiseVar;
    double getAtPosition(int i) const {   // 1-based (:40-46); anything else is a scala.MatchError
        const int D = (int)lengthScales.size();
        if (i == 1) return signalVar;
        if (i > 1 && i < D + 2) return lengthScales[i - 2];
        if (i == D + 2) return noiseVar;
        throw MatchError("scala.MatchError: " + std::to_string(i));
    }
    DenseVector toDenseVector() const {
        DenseVector v(lengthScales.size() + 2);
        for (size_t k = 0; k < v.size(); ++k) v[k] = getAtPosition((int)k + 1);
        return v;
    }
    GaussianRbfParams fromDenseVector(const DenseVector& dv) const {   // :54-58
        if (dv.size() != lengthScales.size() + 2)
            throw IllegalArgumentException("requirement failed: " + std::to_string(dv.size()) + " does not equal to " +
                                           std::to_string(lengthScales.size() + 2));
        return GaussianRbfParams{dv.front(), DenseVector(dv.begin() + 1, dv.end() - 1), dv.back()};
    }
};

class GaussianRbfKernel {   // KernelRequisites.scala:62-114
public:
    explicit GaussianRbfKernel(GaussianRbfParams p) : rbfParams(std::move(p)) {}
    GaussianRbfParams rbfParams;
    int hyperParametersNum() const { return (int)rbfParams.lengthScales.size() + 2; }
    const GaussianRbfParams& hyperParams() const { return rbfParams; }
    GaussianRbfKernel changeHyperParams(const DenseVector& dv) const { return GaussianRbfKernel(rbfParams.fromDenseVector(dv)); }
    DenseVector theta() const { return rbfParams.toDenseVector(); }
    double apply(const DenseVector& a, const DenseVector& b, bool sameIndex) const {   // :66-72, association of :109-113
        double r = 0.0;
        for (size_t d = 0; d < a.size(); ++d) {
            const double diff = a[d] - b[d], inv = 1.0 / (rbfParams.lengthScales[d] * rbfParams.lengthScales[d]);
            r += (diff * inv) * diff;
        }
        const double k = rbfParams.signalVar * rbfParams.signalVar * std::exp(-0.5 * r);
        return sameIndex ? k + rbfParams.noiseVar * rbfParams.noiseVar : k;
    }
};

// ---- utils/MatrixUtils.scala (+ breeze cholesky) -----------------------------------------------------------------------
namespace MatrixUtils {
inline DenseMatrix buildKernelMatrix(const GaussianRbfKernel& k, const DenseMatrix& data, Handle& h = Handle::instance()) {   // :57-70
    DenseMatrix K(data.rows, data.rows);
    const DenseVector th = k.theta();
    h.check(gpk_cov_se_ard(h.get(), data.data.data(), data.rows, data.cols, data.rows, th.data(), K.data.data(), data.rows));
    return K;
}
inline DenseMatrix buildKernelMatrix(const GaussianRbfKernel& k, const DenseMatrix& in1, const DenseMatrix& in2,
                                     Handle& h = Handle::instance()) {                                                      // :44-55
    if (in1.cols != in2.cols) throw IllegalArgumentException("requirement failed: feature dimensions differ");
    DenseMatrix K(in1.rows, in2.rows);
    const DenseVector th = k.theta();
    h.check(gpk_cov_cross_se_ard(h.get(), in1.data.data(), in1.rows, in1.rows, in2.data.data(), in2.rows, in2.rows, in1.cols, th.data(),
                                 K.data.data(), in1.rows));
    return K;
}
inline DenseMatrix cholesky(const DenseMatrix& A, Handle& h = Handle::instance()) {   // breeze cholesky (GpPredictor.scala:120)
    if (A.rows != A.cols) throw IllegalArgumentException("requirement failed: matrix must be square");
    DenseMatrix L(A.rows, A.rows);
    h.check(gpk_potrf_lower(h.get(), A.data.data(), A.rows, A.rows, L.data.data(), A.rows, 1));
    return L;
}
// `transposed`: the operand is the stored matrix's transpose (the `L.t` view of GpPredictor.scala:122)
inline DenseMatrix solve(bool upper, const DenseMatrix& T, const DenseMatrix& B, bool transposed, Handle& h) {
    if (T.rows != T.cols || T.rows != B.rows) throw IllegalArgumentException("requirement failed");   // MatrixUtils.scala:125
    DenseMatrix X(B.rows, B.cols);
    h.check(gpk_trsm(h.get(), upper, transposed, T.data.data(), T.rows, T.rows, B.data.data(), B.cols, B.rows, X.data.data(), B.rows));
    return X;
}
inline DenseMatrix forwardSolve(const DenseMatrix& L, const DenseMatrix& b, bool transposed = false, Handle& h = Handle::instance()) {   // :29-31
    return solve(false, L, b, transposed, h);
}
inline DenseMatrix backSolve(const DenseMatrix& R, const DenseMatrix& b, bool transposed = false, Handle& h = Handle::instance()) {      // :33-35
    return solve(true, R, b, transposed, h);
}
inline DenseVector forwardSolve(const DenseMatrix& L, const DenseVector& b, bool transposed = false, Handle& h = Handle::instance()) {   // :17-21
    DenseMatrix B((int)b.size(), 1); B.data = b;
    return solve(false, L, B, transposed, h).data;
}
inline DenseVector backSolve(const DenseMatrix& R, const DenseVector& b, bool transposed = false, Handle& h = Handle::instance()) {      // :23-27
    DenseMatrix B((int)b.size(), 1); B.data = b;
    return solve(true, R, B, transposed, h).data;
}
inline DenseMatrix invTriangular(const DenseMatrix& m, bool isUpper = false, Handle& h = Handle::instance()) {                           // :106-113
    DenseMatrix out(m.rows, m.rows);
    h.check(gpk_trtri(h.get(), isUpper, m.data.data(), m.rows, m.rows, out.data.data(), m.rows));
    return out;
}
}  // namespace MatrixUtils

// ---- gp/regression/GpPredictor.scala -----------------------------------------------------------------------------------
struct GaussianDistribution { DenseVector mean; DenseMatrix sigma; int dim() const { return (int)mean.size(); } };   // StatsUtils.scala:19-21
struct PredictionTrainingInput { DenseMatrix trainingData; std::optional<double> sigmaNoise; DenseVector targets; }; // GpPredictor.scala:171-172
struct PredictionInput {                                                                                             // :162-169
    DenseMatrix trainingData, testData; std::optional<double> sigmaNoise; DenseVector targets;
    PredictionTrainingInput toPredictionTrainingInput() const { return {trainingData, sigmaNoise, targets}; }
};

class GpPredictor {
public:
    explicit GpPredictor(GaussianRbfKernel k, Handle& h = Handle::instance()) : kernelFunc(std::move(k)), h_(h) {}
    GaussianRbfKernel kernelFunc;

    // :104-124 -> (L, alphaVec, Option[sigmaNoise * I])
    std::tuple<DenseMatrix, DenseVector, std::optional<DenseMatrix>> preComputeComponents(
        const DenseMatrix& X, const GaussianRbfParams& hp, std::optional<double> sigmaNoise, const DenseVector& targets) const {
        require_rows(X, targets);
        const DenseVector th = theta_for(hp, X.cols);
        DenseMatrix L(X.rows, X.rows); DenseVector alpha(X.rows); double ll = 0.0;
        h_.check(gpk_gp_fit(h_.get(), X.data.data(), X.rows, X.cols, X.rows, targets.data(), th.data(), sigmaNoise.has_value(),
                            sigmaNoise.value_or(0.0), L.data.data(), X.rows, alpha.data(), &ll));
        std::optional<DenseMatrix> noise;
        if (sigmaNoise) { noise = DenseMatrix::eye(X.rows); for (auto& v : noise->data) v *= *sigmaNoise; }   // :116
        return {std::move(L), std::move(alpha), std::move(noise)};
    }
    // :60-80 -> (logLikelihood, gradient[optimizedParamsNum])
    std::pair<double, DenseVector> logLikelihoodWithDerivatives(const PredictionTrainingInput& in, const GaussianRbfParams& hp,
                                                                int optimizedParamsNum) const {
        require_rows(in.trainingData, in.targets);
        const DenseMatrix& X = in.trainingData;
        const DenseVector th = theta_for(hp, X.cols);
        double ll = 0.0; DenseVector g(std::max(optimizedParamsNum, 1), 0.0);
        h_.check(gpk_gp_nll_grad(h_.get(), X.data.data(), X.rows, X.cols, X.rows, in.targets.data(), th.data(), in.sigmaNoise.has_value(),
                                 in.sigmaNoise.value_or(0.0), optimizedParamsNum, &ll, g.data()));
        g.resize(optimizedParamsNum);
        return {ll, std::move(g)};
    }
    // :24-43 -> (GaussianDistribution(mean, sigma), logLikelihood); the sigma diagonal includes noiseVar^2 (+ sigmaNoise)
    std::pair<GaussianDistribution, double> predict(const PredictionInput& in, const GaussianRbfParams& hp) const {
        require_rows(in.trainingData, in.targets);
        const DenseMatrix &X = in.trainingData, &Xs = in.testData;
        const DenseVector th = theta_for(hp, X.cols);
        GaussianDistribution d{DenseVector(Xs.rows), DenseMatrix(Xs.rows, Xs.rows)}; double ll = 0.0;
        h_.check(gpk_gp_predict(h_.get(), X.data.data(), X.rows, X.cols, X.rows, in.targets.data(), Xs.data.data(), Xs.rows, Xs.rows,
                                th.data(), in.sigmaNoise.has_value(), in.sigmaNoise.value_or(0.0), d.mean.data(), d.sigma.data.data(),
                                Xs.rows, &ll));
        return {std::move(d), ll};
    }
    std::pair<GaussianDistribution, double> predict(const PredictionInput& in) const { return predict(in, kernelFunc.hyperParams()); }
    // :45-58 computePosterior(trainingData, testData, l, alphaVec, kernelFunc) -> (GaussianDistribution, vMatrix)
    std::pair<GaussianDistribution, DenseMatrix> computePosterior(const DenseMatrix& X, const DenseMatrix& Xs, const DenseMatrix& l,
                                                                  const DenseVector& alphaVec) const {
        const DenseVector th = kernelFunc.theta();
        gpk_model m = nullptr;
        h_.check(gpk_gp_model_from_factor(h_.get(), X.data.data(), X.rows, X.cols, X.rows, l.data.data(), l.rows, alphaVec.data(),
                                          th.data(), &m));
        GaussianDistribution d{DenseVector(Xs.rows), DenseMatrix(Xs.rows, Xs.rows)}; DenseMatrix V(X.rows, Xs.rows);
        const int rc = gpk_gp_model_predict(h_.get(), m, Xs.data.data(), Xs.rows, Xs.rows, 1, d.mean.data(), d.sigma.data.data(), Xs.rows,
                                            V.data.data(), X.rows);
        gpk_gp_model_destroy(h_.get(), m);
        h_.check(rc);
        return {std::move(d), std::move(V)};
    }

private:
    Handle& h_;
    static void require_rows(const DenseMatrix& X, const DenseVector& y) {   // :108
        if (X.rows != (int)y.size())
            throw IllegalArgumentException("requirement failed: Number of objects in training data matrix should be equal to targets vector length");
    }
    static DenseVector theta_for(const GaussianRbfParams& hp, int D) {
        DenseVector th = hp.toDenseVector();
        if ((int)th.size() != D + 2)
            throw IllegalArgumentException("requirement failed: " + std::to_string(th.size()) + " does not equal to " + std::to_string(D + 2));
        return th;
    }
};

// ---- gp/classification/{EpParameterEstimator, GpClassifier}.scala ------------------------------------------------------
struct SiteParams { DenseVector tauSiteParams, niSiteParams; std::optional<double> marginalLogLikelihood; };   // EpParameterEstimator.scala:181-182
struct AvgBasedStopCriterion { double eps = 0.01; int maxSweeps = 100; };                                      // :187-193
struct FixedSweeps { int sweeps = 5; };

class EpParameterEstimator {   // :11-12 (kernelMatrix, targets, stopCriterion)
public:
    EpParameterEstimator(DenseMatrix K, std::vector<int> targets, AvgBasedStopCriterion s, Handle& h = Handle::instance())
        : K_(std::move(K)), t_(std::move(targets)), eps_(s.eps), fixed_(0), max_(s.maxSweeps), h_(h) { require(); }
    EpParameterEstimator(DenseMatrix K, std::vector<int> targets, FixedSweeps s, Handle& h = Handle::instance())
        : K_(std::move(K)), t_(std::move(targets)), eps_(0.0), fixed_(s.sweeps), max_(s.sweeps), h_(h) { require(); }
    int sweeps = 0;
    std::pair<SiteParams, DenseMatrix> estimateSiteParams() {   // :29-69
        const int n = K_.rows;
        SiteParams sp{DenseVector(n), DenseVector(n), 0.0}; DenseMatrix L(n, n); double logZ = 0.0;
        h_.check(gpk_ep_fit(h_.get(), K_.data.data(), n, n, t_.data(), eps_, fixed_, max_, /*as compiled*/ 1, sp.tauSiteParams.data(),
                            sp.niSiteParams.data(), nullptr, L.data.data(), n, nullptr, nullptr, &logZ, &sweeps));
        sp.marginalLogLikelihood = logZ;
        return {std::move(sp), std::move(L)};
    }
private:
    void require() const { if (K_.rows != (int)t_.size()) throw IllegalArgumentException("requirement failed"); }   // :20
    DenseMatrix K_; std::vector<int> t_; double eps_; int fixed_, max_; Handle& h_;
};

class GpClassifier {   // GpClassifier.scala:11, classify :24-47 with learnParams = (siteParams, L)
public:
    explicit GpClassifier(Handle& h = Handle::instance()) : h_(h) {}
    DenseVector classify(const DenseMatrix& trainKernelMatrix, const DenseMatrix& testTrainKernelMatrix, const DenseMatrix& testKernelMatrix,
                         const SiteParams& site, const DenseMatrix& L) const {
        const int n = trainKernelMatrix.rows, m = testTrainKernelMatrix.rows;
        DenseVector kss(m), prob(m);
        for (int i = 0; i < m; ++i) kss[i] = testKernelMatrix(i, i);   // only the diagonal is read (:44)
        h_.check(gpk_ep_classify(h_.get(), trainKernelMatrix.data.data(), n, n, testTrainKernelMatrix.data.data(), m, m, kss.data(),
                                 site.tauSiteParams.data(), site.niSiteParams.data(), L.data.data(), n, prob.data(), nullptr, nullptr));
        return prob;
    }
private:
    Handle& h_;
};

}  // namespace gpk

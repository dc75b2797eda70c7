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
#include <algorithm>
#include <cmath>
#include <deque>
#include <functional>
#include <limits>
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
    void setGraphMode(bool on) const { check(gpk_set_graph_mode(h_, on ? 1 : 0)); }
private:
    gpk_handle h_ = nullptr;
};

// how the (D, theta) arguments of the enclosed gpk_* calls are read (gpk_kernel_family); restored on scope exit
class FamilyScope {
public:
    FamilyScope(Handle& h, int family) : h_(h), prev_(gpk_get_kernel_family(h.get())) { h_.check(gpk_set_kernel_family(h_.get(), family)); }
    ~FamilyScope() { gpk_set_kernel_family(h_.get(), prev_); }
    FamilyScope(const FamilyScope&) = delete;
    FamilyScope& operator=(const FamilyScope&) = delete;
private:
    Handle& h_; int prev_;
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
    using Params = GaussianRbfParams;
    static constexpr int family = GPK_KERNEL_SE_ARD;
    explicit GaussianRbfKernel(GaussianRbfParams p) : rbfParams(std::move(p)) {}
    GaussianRbfParams rbfParams;
    // theta as libgpk reads it for D-dimensional inputs, with the reference's own requirement (:55)
    static DenseVector thetaFor(const GaussianRbfParams& hp, int D) {
        DenseVector th = hp.toDenseVector();
        if ((int)th.size() != D + 2)
            throw IllegalArgumentException("requirement failed: " + std::to_string(th.size()) + " does not equal to " + std::to_string(D + 2));
        return th;
    }
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


// ---- gp/regression/Co2Prediction.scala:16-137: the second closed-form KernelFunc (1-D inputs, 11 hyper-parameters) ----------
struct Co2HyperParams {   // :18-27, a bare DenseVector; getAtPosition is 1-based (dv(i - 1))
    DenseVector dv;
    double getAtPosition(int i) const {
        if (i < 1 || i > (int)dv.size()) throw std::out_of_range("java.lang.IndexOutOfBoundsException: " + std::to_string(i - 1));
        return dv[i - 1];
    }
    DenseVector toDenseVector() const { return dv; }
    Co2HyperParams fromDenseVector(const DenseVector& v) const { return Co2HyperParams{v}; }   // no length requirement (:20)
};

class Co2Kernel {
public:
    using Params = Co2HyperParams;
    static constexpr int family = GPK_KERNEL_CO2;
    explicit Co2Kernel(Co2HyperParams p) : co2HyperParams(std::move(p)) {}
    Co2HyperParams co2HyperParams;
    int hyperParametersNum() const { return 11; }   // :85
    const Co2HyperParams& hyperParams() const { return co2HyperParams; }
    Co2Kernel changeHyperParams(const DenseVector& dv) const { return Co2Kernel(Co2HyperParams{dv}); }
    DenseVector theta() const { return thetaFor(co2HyperParams, 1); }
    static DenseVector thetaFor(const Co2HyperParams& hp, int D) {
        if (D != 1) throw IllegalArgumentException("requirement failed: This kernel is applicable only for 1D objects");   // :39
        DenseVector th(11);
        for (int i = 1; i <= 11; ++i) th[i - 1] = hp.getAtPosition(i);   // getHyperParams :87-92
        return th;
    }
    double apply(const DenseVector& a, const DenseVector& b, bool sameIndex) const {   // :38-56
        if (a.size() != 1 || b.size() != 1) throw IllegalArgumentException("requirement failed: This kernel is applicable only for 1D objects");
        const DenseVector h = theta();
        const double xDiff = a[0] - b[0], xDiffSq = xDiff * xDiff;
        const double k1 = h[0] * h[0] * std::exp(-xDiffSq / (2 * h[1] * h[1]));
        const double sinVal = std::sin(3.14159265358979323846 * xDiff);
        const double k2 = h[2] * h[2] * std::exp((-xDiffSq / (2 * h[3] * h[3])) - 2 * sinVal * sinVal / (h[4] * h[4]));
        const double k3 = h[5] * h[5] * std::pow(1 + xDiffSq / (2 * h[7] * h[6] * h[6]), -h[7]);
        const double k4 = h[8] * h[8] * std::exp(-xDiffSq / (2 * h[9] * h[9]));
        return k1 + k2 + k3 + k4 + (sameIndex ? h[10] * h[10] : 0.);
    }
};

// ---- optimization/Optimization.scala:30-61 -----------------------------------------------------------------------------------
// BreezeLbfgsOptimizer: L-BFGS(m = 4, maxIter) with the reference's best-seen bookkeeping (:37-55).  Breeze's LBFGS is un-vendored
// third-party code whose trajectory no reference test pins; a textbook two-loop L-BFGS with an Armijo backtracking line search
// stands in (host control flow: each evaluation of `func` is one device call).
using objectiveFunctionWithGradient = std::function<std::pair<double, DenseVector>(const DenseVector&)>;
class BreezeLbfgsOptimizer {
public:
    explicit BreezeLbfgsOptimizer(int maxIter = 10, int m = 4) : maxIter_(maxIter), m_(m) {}
    int evaluations = 0;
    DenseVector minimize(const objectiveFunctionWithGradient& func, const DenseVector& initPoint) {
        DenseVector bestX = initPoint; double bestV = std::numeric_limits<double>::max();     // :38-39
        auto calc = [&](const DenseVector& x) {
            auto r = func(x); ++evaluations;
            if (r.first < bestV) { bestV = r.first; bestX = x; }                               // :44-46
            return r;
        };
        const size_t n = initPoint.size();
        DenseVector x = initPoint;
        auto [f, g] = calc(x);
        std::deque<std::pair<DenseVector, DenseVector>> hist;   // (s, y)
        auto dot = [n](const DenseVector& a, const DenseVector& b) { double s = 0; for (size_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; };
        for (int it = 0; it < maxIter_; ++it) {
            if (std::sqrt(dot(g, g)) <= 1e-9 * std::max(1.0, std::fabs(f))) break;
            DenseVector d = g; std::vector<double> a(hist.size());
            for (int k = (int)hist.size() - 1; k >= 0; --k) {
                a[k] = dot(hist[k].first, d) / dot(hist[k].first, hist[k].second);
                for (size_t i = 0; i < n; ++i) d[i] -= a[k] * hist[k].second[i];
            }
            if (!hist.empty()) {
                const double gam = dot(hist.back().first, hist.back().second) / dot(hist.back().second, hist.back().second);
                for (auto& v : d) v *= gam;
            }
            for (size_t k = 0; k < hist.size(); ++k) {
                const double b = dot(hist[k].second, d) / dot(hist[k].first, hist[k].second);
                for (size_t i = 0; i < n; ++i) d[i] += (a[k] - b) * hist[k].first[i];
            }
            for (auto& v : d) v = -v;
            double slope = dot(g, d);
            if (!(slope < 0)) { d = g; for (auto& v : d) v = -v; slope = -dot(g, g); hist.clear(); }
            double step = hist.empty() ? 1.0 / std::max(1.0, std::sqrt(dot(g, g))) : 1.0;
            DenseVector xn(n), gn; double fn = f; bool ok = false;
            for (int ls = 0; ls < 20; ++ls, step *= 0.5) {
                for (size_t i = 0; i < n; ++i) xn[i] = x[i] + step * d[i];
                auto r = calc(xn); fn = r.first; gn = std::move(r.second);
                if (std::isfinite(fn) && fn <= f + 1e-4 * step * slope) { ok = true; break; }
            }
            if (!ok) break;
            DenseVector sv(n), yv(n);
            for (size_t i = 0; i < n; ++i) { sv[i] = xn[i] - x[i]; yv[i] = gn[i] - g[i]; }
            if (dot(sv, yv) > 1e-12 * std::sqrt(dot(sv, sv) * dot(yv, yv))) {
                hist.emplace_back(std::move(sv), std::move(yv));
                if ((int)hist.size() > m_) hist.pop_front();
            }
            x = xn; f = fn; g = gn;
        }
        const double optimalVal = func(x).first;                  // :52
        return optimalVal < bestV ? x : bestX;                    // :53-55
    }
    DenseVector maximize(const objectiveFunctionWithGradient& func, const DenseVector& initPoint) {   // :58-60
        return minimize([&](const DenseVector& p) { auto r = func(p); for (auto& v : r.second) v = -v; return std::make_pair(-r.first, std::move(r.second)); },
                        initPoint);
    }
private:
    int maxIter_, m_;
};

// ---- utils/MatrixUtils.scala (+ breeze cholesky) -----------------------------------------------------------------------
namespace MatrixUtils {
template <class Kernel>
inline DenseMatrix buildKernelMatrix(const Kernel& k, const DenseMatrix& data, Handle& h = Handle::instance()) {   // :57-70
    DenseMatrix K(data.rows, data.rows);
    const DenseVector th = Kernel::thetaFor(k.hyperParams(), data.cols);
    FamilyScope fam(h, Kernel::family);
    h.check(gpk_cov_se_ard(h.get(), data.data.data(), data.rows, data.cols, data.rows, th.data(), K.data.data(), data.rows));
    return K;
}
template <class Kernel>
inline DenseMatrix buildKernelMatrix(const Kernel& k, const DenseMatrix& in1, const DenseMatrix& in2,
                                     Handle& h = Handle::instance()) {                                                      // :44-55
    if (in1.cols != in2.cols) throw IllegalArgumentException("requirement failed: feature dimensions differ");
    DenseMatrix K(in1.rows, in2.rows);
    const DenseVector th = Kernel::thetaFor(k.hyperParams(), in1.cols);
    FamilyScope fam(h, Kernel::family);
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

// Device-resident (X, L^-1, alpha, theta): the fit-once / predict-many pattern of GP-UKF and GP-UCB
// (GPUnscentedKalmanFilter.scala:77-88,123-147; GPOptimizer.scala:48-109) without moving L across PCIe.
class FittedGp {
public:
    FittedGp(Handle& h, gpk_model m, int n, int D, double ll, std::optional<double> sigmaNoise)
        : logLikelihood(ll), h_(h), m_(m), n_(n), D_(D), sigmaNoise_(sigmaNoise) {}
    FittedGp(FittedGp&& o) noexcept : logLikelihood(o.logLikelihood), h_(o.h_), m_(o.m_), n_(o.n_), D_(o.D_), sigmaNoise_(o.sigmaNoise_) { o.m_ = nullptr; }
    FittedGp(const FittedGp&) = delete;
    FittedGp& operator=(const FittedGp&) = delete;
    ~FittedGp() { if (m_) gpk_gp_model_destroy(h_.get(), m_); }
    double logLikelihood;
    int size() const { return n_; }
    // GpPredictor.scala:45-58 computePosterior -> (GaussianDistribution, vMatrix)
    std::pair<GaussianDistribution, DenseMatrix> computePosterior(const DenseMatrix& Xs) const;
    DenseVector mean(const DenseMatrix& Xs) const {   // K* alpha only (GPUnscentedKalmanFilter.scala:77-90 reads nothing else)
        DenseVector mu(Xs.rows);
        h_.check(gpk_gp_model_predict(h_.get(), m_, Xs.data.data(), Xs.rows, Xs.rows, 0, mu.data(), nullptr, Xs.rows, nullptr, n_));
        return mu;
    }
    DenseVector alphaVec() const { DenseVector a(n_); h_.check(gpk_gp_model_get_alpha(h_.get(), m_, a.data())); return a; }
    // GPOptimizer.scala:64-71 + :51: one more evaluated point, hyper-parameters unchanged -> bordered O(n^2) update
    void append(const DenseVector& point, double target) {
        if ((int)point.size() != D_) throw IllegalArgumentException("requirement failed: point dimension");
        double d = 0.0;
        h_.check(gpk_gp_model_append(h_.get(), m_, point.data(), target, sigmaNoise_.has_value(), sigmaNoise_.value_or(0.0), &d));
        ++n_; logLikelihood += d;
    }
    // GPOptimizer.scala:87-104: maximizeUCB's objective and gradient at the rows of Xs -> (ucb[m], grad m x D)
    std::pair<DenseVector, DenseMatrix> ucbWithGradient(const DenseMatrix& Xs, double kParam) const {
        DenseVector ucb(Xs.rows); DenseMatrix grad(Xs.rows, D_);
        h_.check(gpk_gp_model_ucb(h_.get(), m_, Xs.data.data(), Xs.rows, Xs.rows, kParam, ucb.data(), grad.data.data(), Xs.rows, nullptr, nullptr));
        return {std::move(ucb), std::move(grad)};
    }
private:
    Handle& h_; gpk_model m_; int n_, D_; std::optional<double> sigmaNoise_;
};
inline std::pair<GaussianDistribution, DenseMatrix> FittedGp::computePosterior(const DenseMatrix& Xs) const {
    GaussianDistribution d{DenseVector(Xs.rows), DenseMatrix(Xs.rows, Xs.rows)}; DenseMatrix V(n_, Xs.rows);
    h_.check(gpk_gp_model_predict(h_.get(), m_, Xs.data.data(), Xs.rows, Xs.rows, 1, d.mean.data(), d.sigma.data.data(), Xs.rows,
                                  V.data.data(), n_));
    return {std::move(d), std::move(V)};
}

template <class Kernel>
class GpPredictorT {
public:
    using Params = typename Kernel::Params;
    explicit GpPredictorT(Kernel k, Handle& h = Handle::instance()) : kernelFunc(std::move(k)), h_(h) {}
    Kernel kernelFunc;

    // :104-124 -> (L, alphaVec, Option[sigmaNoise * I])
    std::tuple<DenseMatrix, DenseVector, std::optional<DenseMatrix>> preComputeComponents(
        const DenseMatrix& X, const Params& hp, std::optional<double> sigmaNoise, const DenseVector& targets) const {
        require_rows(X, targets);
        const DenseVector th = Kernel::thetaFor(hp, X.cols);
        DenseMatrix L(X.rows, X.rows); DenseVector alpha(X.rows); double ll = 0.0;
        FamilyScope fam(h_, Kernel::family);
        h_.check(gpk_gp_fit(h_.get(), X.data.data(), X.rows, X.cols, X.rows, targets.data(), th.data(), sigmaNoise.has_value(),
                            sigmaNoise.value_or(0.0), L.data.data(), X.rows, alpha.data(), &ll));
        std::optional<DenseMatrix> noise;
        if (sigmaNoise) { noise = DenseMatrix::eye(X.rows); for (auto& v : noise->data) v *= *sigmaNoise; }   // :116
        return {std::move(L), std::move(alpha), std::move(noise)};
    }
    // :60-80 -> (logLikelihood, gradient[optimizedParamsNum])
    std::pair<double, DenseVector> logLikelihoodWithDerivatives(const PredictionTrainingInput& in, const Params& hp,
                                                                int optimizedParamsNum) const {
        require_rows(in.trainingData, in.targets);
        const DenseMatrix& X = in.trainingData;
        const DenseVector th = Kernel::thetaFor(hp, X.cols);
        double ll = 0.0; DenseVector g(std::max(optimizedParamsNum, 1), 0.0);
        FamilyScope fam(h_, Kernel::family);
        h_.check(gpk_gp_nll_grad(h_.get(), X.data.data(), X.rows, X.cols, X.rows, in.targets.data(), th.data(), in.sigmaNoise.has_value(),
                                 in.sigmaNoise.value_or(0.0), optimizedParamsNum, &ll, g.data()));
        g.resize(optimizedParamsNum);
        return {ll, std::move(g)};
    }
    // :24-43 -> (GaussianDistribution(mean, sigma), logLikelihood); the sigma diagonal includes noiseVar^2 (+ sigmaNoise)
    std::pair<GaussianDistribution, double> predict(const PredictionInput& in, const Params& hp) const {
        require_rows(in.trainingData, in.targets);
        const DenseMatrix &X = in.trainingData, &Xs = in.testData;
        const DenseVector th = Kernel::thetaFor(hp, X.cols);
        GaussianDistribution d{DenseVector(Xs.rows), DenseMatrix(Xs.rows, Xs.rows)}; double ll = 0.0;
        FamilyScope fam(h_, Kernel::family);
        h_.check(gpk_gp_predict(h_.get(), X.data.data(), X.rows, X.cols, X.rows, in.targets.data(), Xs.data.data(), Xs.rows, Xs.rows,
                                th.data(), in.sigmaNoise.has_value(), in.sigmaNoise.value_or(0.0), d.mean.data(), d.sigma.data.data(),
                                Xs.rows, &ll));
        return {std::move(d), ll};
    }
    std::pair<GaussianDistribution, double> predict(const PredictionInput& in) const { return predict(in, kernelFunc.hyperParams()); }
    // :45-58 computePosterior(trainingData, testData, l, alphaVec, kernelFunc) -> (GaussianDistribution, vMatrix)
    std::pair<GaussianDistribution, DenseMatrix> computePosterior(const DenseMatrix& X, const DenseMatrix& Xs, const DenseMatrix& l,
                                                                  const DenseVector& alphaVec) const {
        const DenseVector th = Kernel::thetaFor(kernelFunc.hyperParams(), X.cols);
        gpk_model m = nullptr;
        {
            FamilyScope fam(h_, Kernel::family);
            h_.check(gpk_gp_model_from_factor(h_.get(), X.data.data(), X.rows, X.cols, X.rows, l.data.data(), l.rows, alphaVec.data(),
                                              th.data(), &m));
        }
        return FittedGp(h_, m, X.rows, X.cols, 0.0, std::nullopt).computePosterior(Xs);
    }
    // resident preComputeComponents
    FittedGp fit(const DenseMatrix& X, std::optional<double> sigmaNoise, const DenseVector& targets, const Params& hp) const {
        require_rows(X, targets);
        const DenseVector th = Kernel::thetaFor(hp, X.cols);
        gpk_model m = nullptr; double ll = 0.0;
        FamilyScope fam(h_, Kernel::family);
        h_.check(gpk_gp_model_fit(h_.get(), X.data.data(), X.rows, X.cols, X.rows, targets.data(), th.data(), sigmaNoise.has_value(),
                                  sigmaNoise.value_or(0.0), &m, &ll));
        return FittedGp(h_, m, X.rows, X.cols, ll, sigmaNoise);
    }
    // :126-142: L-BFGS(m = 4, maxIter = 20) maximisation of logLikelihoodWithDerivatives in the natural parameters.  optimizeNoise =
    // false reproduces the reference's defect: the start point drops the last entry (`toDenseVector(0 to -2)`) and the first
    // evaluation fails in fromDenseVector's require (KernelRequisites.scala:55) / getAtPosition(11) (Co2Prediction.scala:23).
    Params obtainOptimalHyperParams(const DenseMatrix& X, std::optional<double> sigmaNoise, const DenseVector& targets, bool optimizeNoise,
                                    int maxIter = 20) const {
        BreezeLbfgsOptimizer opt(maxIter);
        const Params hp0 = kernelFunc.hyperParams();
        DenseVector init = hp0.toDenseVector();
        if (!optimizeNoise) init.pop_back();
        const PredictionTrainingInput ptInput{X, sigmaNoise, targets};
        auto llObjFunction = [&](const DenseVector& currentParams) {
            const Params hp = hp0.fromDenseVector(currentParams);
            return logLikelihoodWithDerivatives(ptInput, hp, (int)currentParams.size());
        };
        return hp0.fromDenseVector(opt.maximize(llObjFunction, init));
    }
    // :82-87 -> (posterior, logLikelihood, optimal hyper-parameters)
    std::tuple<GaussianDistribution, double, Params> predictWithParamsOptimization(const PredictionInput& in, bool optimizeNoise) const {
        Params optimal = obtainOptimalHyperParams(in.trainingData, in.sigmaNoise, in.targets, optimizeNoise);
        auto r = predict(in, optimal);
        return {std::move(r.first), r.second, std::move(optimal)};
    }

private:
    Handle& h_;
    static void require_rows(const DenseMatrix& X, const DenseVector& y) {   // :108
        if (X.rows != (int)y.size())
            throw IllegalArgumentException("requirement failed: Number of objects in training data matrix should be equal to targets vector length");
    }
};
using GpPredictor = GpPredictorT<GaussianRbfKernel>;   // spring-context.xml:33-47
using Co2GpPredictor = GpPredictorT<Co2Kernel>;        // spring-context.xml:49-51 "co2GpPredictor"

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

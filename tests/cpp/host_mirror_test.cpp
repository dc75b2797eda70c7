// Exercises the C++ host mirror (include/gpk.hpp) end to end: reads a problem written by tests/test_gpu_cpp_mirror.py,
// calls the mirror exactly like a compiled caller of the reference's Scala API would, and writes the results as JSON for the
// Python side to compare with the oracle.   usage: host_mirror_test <in.bin> <out.json>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../include/gpk.hpp"

using namespace gpk;

static void put(std::ostream& o, const char* name, const DenseVector& v, bool last = false) {
    o << "\"" << name << "\": [";
    o.precision(17);
    for (size_t i = 0; i < v.size(); ++i) o << (i ? ", " : "") << v[i];
    o << "]" << (last ? "\n" : ",\n");
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.bin out.json\n", argv[0]); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    int hdr[4];   // n, D, m, P
    in.read((char*)hdr, sizeof(hdr));
    const int n = hdr[0], D = hdr[1], m = hdr[2];
    DenseMatrix X(n, D), Xs(m, D);
    DenseVector y(n), theta(D + 2);
    std::vector<int> targets(n);
    in.read((char*)X.data.data(), sizeof(double) * n * D);
    in.read((char*)Xs.data.data(), sizeof(double) * m * D);
    in.read((char*)y.data(), sizeof(double) * n);
    in.read((char*)theta.data(), sizeof(double) * (D + 2));
    in.read((char*)targets.data(), sizeof(int) * n);
    if (!in) { std::fprintf(stderr, "short input file\n"); return 2; }

    std::ofstream out(argv[2]);
    out << "{\n";
    try {
        GaussianRbfParams hp{theta[0], DenseVector(theta.begin() + 1, theta.end() - 1), theta.back()};
        GaussianRbfKernel kf(hp);
        GpPredictor pred(kf);
        // utils.MatrixUtils
        DenseMatrix K = MatrixUtils::buildKernelMatrix(kf, X);
        DenseMatrix L = MatrixUtils::cholesky(K);
        DenseVector z = MatrixUtils::forwardSolve(L, y);
        DenseVector alpha = MatrixUtils::backSolve(L, z, /*transposed = L.t view*/ true);
        DenseMatrix Li = MatrixUtils::invTriangular(L, false);
        put(out, "K_diag", [&] { DenseVector d(n); for (int i = 0; i < n; ++i) d[i] = K(i, i); return d; }());
        put(out, "alpha_via_solves", alpha);
        put(out, "Li_last_row", [&] { DenseVector d(n); for (int i = 0; i < n; ++i) d[i] = Li(n - 1, i); return d; }());
        // gp.regression.GpPredictor
        auto [ll, grad] = pred.logLikelihoodWithDerivatives(PredictionTrainingInput{X, std::nullopt, y}, hp, D + 2);
        put(out, "ll", DenseVector{ll});
        put(out, "grad", grad);
        auto [dist, ll2] = pred.predict(PredictionInput{X, Xs, 0.05, y}, hp);
        put(out, "pred_mean", dist.mean);
        put(out, "pred_var", [&] { DenseVector d(m); for (int i = 0; i < m; ++i) d[i] = dist.sigma(i, i); return d; }());
        put(out, "pred_ll", DenseVector{ll2});
        auto [Lf, alphaF, noise] = pred.preComputeComponents(X, hp, std::nullopt, y);
        auto [post, V] = pred.computePosterior(X, Xs, Lf, alphaF);
        put(out, "post_mean", post.mean);
        put(out, "V_col0", DenseVector(V.data.begin(), V.data.begin() + n));
        // gp.classification
        EpParameterEstimator ep(K, targets, FixedSweeps{3});
        auto [site, Lep] = ep.estimateSiteParams();
        put(out, "ep_tau", site.tauSiteParams);
        put(out, "ep_nu", site.niSiteParams);
        put(out, "ep_logZ", DenseVector{*site.marginalLogLikelihood});
        DenseMatrix Ks = MatrixUtils::buildKernelMatrix(kf, Xs, X), Kss = MatrixUtils::buildKernelMatrix(kf, Xs);
        put(out, "ep_prob", GpClassifier().classify(K, Ks, Kss, site, Lep));
        // error behaviour: require(...) -> IllegalArgumentException, non-PD -> NotConvergedException(minor), MatchError
        int errors = 0;
        try { pred.logLikelihoodWithDerivatives(PredictionTrainingInput{X, std::nullopt, DenseVector(n - 1)}, hp, 1); } catch (const IllegalArgumentException&) { errors |= 1; }
        try { DenseMatrix B = DenseMatrix::eye(5); B(3, 3) = -1.0; MatrixUtils::cholesky(B); } catch (const NotConvergedException& e) { if (e.minor == 4) errors |= 2; }
        try { hp.getAtPosition(D + 3); } catch (const MatchError&) { errors |= 4; }
        try { DenseMatrix B(2, 2); B(0, 0) = B(1, 1) = 1.0; B(0, 1) = 2.0; B(1, 0) = 3.0; MatrixUtils::cholesky(B); } catch (const MatrixNotSymmetricException&) { errors |= 8; }
        // GpPredictor.obtainOptimalHyperParams (GpPredictor.scala:126-142) through the C++ L-BFGS wrapper: every evaluation is a device call
        GaussianRbfParams opt = pred.obtainOptimalHyperParams(X, std::nullopt, y, true);
        put(out, "opt_theta", opt.toDenseVector());
        put(out, "opt_ll", DenseVector{pred.logLikelihoodWithDerivatives(PredictionTrainingInput{X, std::nullopt, y}, opt, 0).first});
        try { pred.obtainOptimalHyperParams(X, std::nullopt, y, false); } catch (const IllegalArgumentException&) { errors |= 16; }
        // resident model: fit on all but the last 3 rows, append them one by one (GPOptimizer.scala:48-71), posterior + UCB
        {
            DenseMatrix Xh(n - 3, D); DenseVector yh(y.begin(), y.end() - 3);
            for (int c = 0; c < D; ++c) for (int r = 0; r < n - 3; ++r) Xh(r, c) = X(r, c);
            FittedGp model = pred.fit(Xh, std::nullopt, yh, hp);
            for (int r = n - 3; r < n; ++r) { DenseVector pt(D); for (int c = 0; c < D; ++c) pt[c] = X(r, c); model.append(pt, y[r]); }
            put(out, "model_size", DenseVector{(double)model.size()});
            put(out, "model_ll", DenseVector{model.logLikelihood});
            put(out, "model_alpha", model.alphaVec());
            put(out, "model_mean", model.mean(Xs));
            auto [ucb, ugrad] = model.ucbWithGradient(Xs, 1.5);
            put(out, "model_ucb", ucb);
            put(out, "model_ucb_grad", ugrad.data);
        }
        // Co2Kernel (Co2Prediction.scala:29-137) through the same predictor template: 1-D inputs on a year axis
        {
            DenseMatrix T(n, 1), Ts(m, 1); DenseVector yc(n);
            for (int r = 0; r < n; ++r) { T(r, 0) = 1958.0 + 30.0 * X(r, 0); yc[r] = 330.0 + 10.0 * y[r]; }
            for (int r = 0; r < m; ++r) Ts(r, 0) = 1958.0 + 30.0 * Xs(r, 0);
            Co2HyperParams chp{DenseVector{60., 70., 8., 50., 2., 0.34, 2.4, 0.88, 0.26, 0.2, 1.5}};
            Co2Kernel ck(chp);
            Co2GpPredictor cpred(ck);
            auto [cll, cgrad] = cpred.logLikelihoodWithDerivatives(PredictionTrainingInput{T, std::nullopt, yc}, chp, 11);
            put(out, "co2_ll", DenseVector{cll});
            put(out, "co2_grad", cgrad);
            auto [cdist, cll2] = cpred.predict(PredictionInput{T, Ts, std::nullopt, yc}, chp);
            put(out, "co2_mean", cdist.mean);
            DenseMatrix Kc = MatrixUtils::buildKernelMatrix(ck, T);
            put(out, "co2_K_row0", [&] { DenseVector d(n); for (int i = 0; i < n; ++i) d[i] = Kc(0, i); return d; }());
            put(out, "co2_apply", DenseVector{ck.apply({T(0, 0)}, {T(1, 0)}, false), ck.apply({T(0, 0)}, {T(0, 0)}, true)});
            try { cpred.logLikelihoodWithDerivatives(PredictionTrainingInput{X, std::nullopt, y}, chp, 11); } catch (const IllegalArgumentException&) { errors |= 32; }
            try { cpred.obtainOptimalHyperParams(T, std::nullopt, yc, false); } catch (const std::out_of_range&) { errors |= 64; }
        }
        // the SE family is back in force afterwards
        put(out, "ll_again", DenseVector{pred.logLikelihoodWithDerivatives(PredictionTrainingInput{X, std::nullopt, y}, hp, 0).first});
        put(out, "errors_caught", DenseVector{(double)errors}, true);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_mirror_test failed: %s\n", e.what());
        return 1;
    }
    out << "}\n";
    return 0;
}

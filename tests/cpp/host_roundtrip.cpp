// Drives the C++ host mirror (include/ifb200_host.hpp) the way the reference's Scala tests drive its classes
// (isolation-forest/src/test/scala/.../IsolationForestModelWriteReadTest.scala:41-110): fit, transform,
// save, load, transform again, compare.  Built and run by tests/test_host_gpu.py::test_cpp_host_program.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "ifb200_host.hpp"

using namespace ifb200;

#define CHECK(cond)                                                       \
    do {                                                                  \
        if (!(cond)) {                                                    \
            std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); \
            return 1;                                                     \
        }                                                                 \
    } while (0)

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp/ifb200_cpp_model";
    const int64_t n = 20000;
    const int d = 8;
    std::mt19937_64 gen(7);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<double> X((size_t)n * d);
    for (auto &v : X) v = nd(gen);
    for (int64_t r = 0; r < n; r += 50)  // 2 % outliers
        for (int c = 0; c < d; c++) X[(size_t)r * d + c] *= 4.0;
    FeatureMatrix data{n, d, X.data(), nullptr};

    IsolationForest est;
    est.setNumEstimators(64).setMaxSamples(256).setContamination(0.02).setRandomSeed(3);
    auto model = est.fit(data);
    CHECK(model->numTrees() == 64 && model->getNumSamples() == 256 && model->getTotalNumFeatures() == d);
    ScoredData a = model->transform(data);
    double frac = 0;
    for (double l : a.predictedLabel) frac += l;
    frac /= (double)n;
    CHECK(std::fabs(frac - 0.02) <= 0.02 * 0.01 + 1.0 / n);   // exact-quantile threshold
    double mo = 0, mi = 0;
    for (int64_t r = 0; r < n; r++) (r % 50 == 0 ? mo : mi) += a.outlierScore[r];
    CHECK(mo / (n / 50) > mi / (n - n / 50) + 0.1);            // planted outliers score higher

    model->save(dir, /*overwrite=*/true);
    auto loaded = IsolationForestModel::load(dir);
    CHECK(loaded->uid() == model->uid());
    CHECK(loaded->getOutlierScoreThreshold() == model->getOutlierScoreThreshold());
    CHECK(loaded->paramMapJson(false) == model->paramMapJson(false));
    ScoredData b = loaded->transform(data);
    CHECK(a.outlierScore == b.outlierScore && a.predictedLabel == b.predictedLabel);
    CHECK(loaded->treeToString(5) == model->treeToString(5));

    ExtendedIsolationForest eest;
    eest.setNumEstimators(32).setRandomSeed(5);
    auto em = eest.fit(data);
    CHECK(em->getExtensionLevel() == d - 1 && !eest.isSetExtensionLevel());
    bool threw = false;
    try {
        ExtendedIsolationForest bad;
        bad.setExtensionLevel(d);
        bad.fit(data);
    } catch (const IllegalArgumentException &) {
        threw = true;
    }
    CHECK(threw);
    std::printf("host_roundtrip ok: threshold %.17g, outlier mean %.3f, inlier mean %.3f\n",
                model->getOutlierScoreThreshold(), mo / (n / 50), mi / (n - n / 50));
    return 0;
}

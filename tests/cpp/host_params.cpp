// CPU-only exercise of the C++ host mirror (include/ifb200_host.hpp): everything here runs before any device work,
// so it is part of the `-m "not gpu"` suite.  Mirrors the reference's parameter / persistence tests:
//   IsolationForestTest.isolationForestEstimatorWriteReadTest          (IFT/IsolationForestTest.scala:16-45)
//   IsolationForestModelWriteReadTest.emptyIsolationForestModelWriteReadTest (IFT/...WriteReadTest.scala:252-296)
//   Params.validate messages                                          (IF/core/IsolationForestParamsBase.scala)
//   validateAndResolveParams messages                                 (IF/core/SharedTrainLogic.scala:27-78)
#include <cstdio>
#include <string>

#include "ifb200_host.hpp"

using namespace ifb200;

#define CHECK(cond)                                                                \
    do {                                                                           \
        if (!(cond)) {                                                             \
            std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

template <typename F>
static std::string message_of(F &&f) {
    try {
        f();
    } catch (const IllegalArgumentException &e) {
        return std::string("IAE:") + e.what();
    } catch (const std::exception &e) {
        return std::string("EXC:") + e.what();
    }
    return "";
}

int main(int argc, char **argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp/ifb200_cpp_params";

    // ---- estimator params round trip --------------------------------------------------------------------
    IsolationForest est("isolation-forest_0123456789ab");
    est.setNumEstimators(200).setBootstrap(true).setMaxSamples(10000).setMaxFeatures(0.7).setContamination(0.02)
        .setContaminationError(0.0002).setRandomSeed(1).setFeaturesCol("featuresTestColumn")
        .setPredictionCol("predictedLabelTestColumn").setScoreCol("outlierScoreTestColumn");
    est.save(dir + "/est", /*overwrite=*/true);
    CHECK(message_of([&] { est.save(dir + "/est"); }).find("already exists") != std::string::npos);
    auto est2 = IsolationForest::load(dir + "/est");
    CHECK(est2->uid() == est.uid());
    CHECK(est2->paramMapJson(false) == est.paramMapJson(false));
    CHECK(est2->isSet("maxFeatures") && est2->isSet("bootstrap") && est2->getMaxSamples() == 10000.0);
    IsolationForest fresh;
    fresh.save(dir + "/fresh", true);
    auto fresh2 = IsolationForest::load(dir + "/fresh");
    CHECK(!fresh2->isSet("numEstimators") && fresh2->getNumEstimators() == 100 && fresh2->getMaxSamples() == 256.0);
    CHECK(fresh2->uid().rfind("isolation-forest_", 0) == 0 && fresh2->uid().size() == std::string("isolation-forest_").size() + 12);

    ExtendedIsolationForest ext;
    CHECK(!ext.isSetExtensionLevel());
    ext.setExtensionLevel(4).setNumEstimators(10);
    ext.save(dir + "/ext", true);
    auto ext2 = ExtendedIsolationForest::load(dir + "/ext");
    CHECK(ext2->isSetExtensionLevel() && ext2->getExtensionLevel() == 4 && ext2->getNumEstimators() == 10);
    CHECK(message_of([&] { IsolationForest::load(dir + "/ext"); }).find("Expected class") != std::string::npos);

    // ---- Params.validate messages ------------------------------------------------------------------------
    CHECK(message_of([&] { est.setNumEstimators(0); }).find("parameter numEstimators given invalid value 0.") != std::string::npos);
    CHECK(message_of([&] { est.setContamination(0.5); }).find("parameter contamination given invalid value 0.5.") != std::string::npos);
    CHECK(message_of([&] { est.setMaxSamples(-1.0); }).find("parameter maxSamples given invalid value -1.0.") != std::string::npos);
    CHECK(message_of([&] { est.setRandomSeed(0); }).find("parameter randomSeed given invalid value 0.") != std::string::npos);
    CHECK(message_of([&] { ext.setExtensionLevel(-1); }).find("parameter extensionLevel given invalid value -1.") != std::string::npos);

    // ---- validateAndResolveParams ------------------------------------------------------------------------
    ResolvedParams rp = validateAndResolveParams(11183, 6, 1.0, 256.0);
    CHECK(rp.numFeatures == 6 && rp.numSamples == 256 && rp.totalNumSamples == 11183 && rp.totalNumFeatures == 6);
    rp = validateAndResolveParams(11183, 6, 0.5, 0.01);
    CHECK(rp.numFeatures == 3 && rp.numSamples == 111);
    CHECK(message_of([&] { validateAndResolveParams(100, 6, 1.0, 256.0); })
              .find("specifying the use of 256 samples, but only 100 samples are in the input dataset.") != std::string::npos);
    CHECK(message_of([&] { validateAndResolveParams(100, 6, 7.0, 10.0); })
              .find("specifying the use of 7 features, but only 6 features are available.") != std::string::npos);
    CHECK(message_of([&] { validateAndResolveParams(100, 6, 1.0, 0.01); })
              .find("specifying the use of 1 samples, but >=2 samples are required.") != std::string::npos);

    // ---- an empty model survives write / read; constructor requires -------------------------------------
    ForestTables none;
    IsolationForestModel empty("testUid", none, /*numSamples=*/1, /*numFeatures=*/1, /*totalNumFeatures=*/1);
    empty.setOutlierScoreThreshold(0.6);
    empty.save(dir + "/empty", true);
    auto empty2 = IsolationForestModel::load(dir + "/empty");
    CHECK(empty2->numTrees() == 0 && empty2->uid() == "testUid" && empty2->getOutlierScoreThreshold() == 0.6);
    CHECK(message_of([&] { IsolationForestModel bad("u", none, 0, 1); }).find("parameter numSamples must be >0") != std::string::npos);
    CHECK(message_of([&] { ExtendedIsolationForestModel bad("u", none, 4, 5, 3); })
              .find("parameter numFeatures must be <= totalNumFeatures") != std::string::npos);

    // ---- a reference-written model, when one is given (tests pass the path of the committed Spark output) -----
    if (argc > 2) {
        auto m = IsolationForestModel::load(argv[2]);
        CHECK(m->numTrees() > 0 && m->getNumSamples() > 0);
        const std::string text = m->treeToString(0);
        CHECK(text.rfind("InternalNode(", 0) == 0 || text.rfind("ExternalNode(", 0) == 0);
        m->save(dir + "/resaved", true);
        auto again = IsolationForestModel::load(dir + "/resaved");
        CHECK(again->tables().threshold == m->tables().threshold && again->tables().left == m->tables().left);
        CHECK(again->treeToString(0) == text && again->paramMapJson(false) == m->paramMapJson(false));
        std::printf("reference model: %d trees, %zu nodes\n", m->numTrees(), m->tables().left.size());
    }
    std::printf("host_params ok\n");
    return 0;
}

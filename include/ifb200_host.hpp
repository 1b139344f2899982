// ifb200_host.hpp -- host-side mirror of the reference's Spark-ML surface, in C++ (libifb200_host.so).
//
// There is no JVM/Scala/Spark in the build image, so the host code above the C ABI (include/ifb200.h) is
// written in C++ with the reference's names, argument meaning and error behaviour:
//
//   reference (IF/ = isolation-forest/src/main/scala/com/linkedin/relevance/isolationforest/)   here
//   IsolationForest                 IF/IsolationForest.scala:25-105                 ifb200::IsolationForest
//   IsolationForestModel            IF/IsolationForestModel.scala:37-191            ifb200::IsolationForestModel
//   ExtendedIsolationForest         IF/extended/ExtendedIsolationForest.scala       ifb200::ExtendedIsolationForest
//   ExtendedIsolationForestModel    IF/extended/ExtendedIsolationForestModel.scala  ifb200::ExtendedIsolationForestModel
//   IsolationForestParamsBase       IF/core/IsolationForestParamsBase.scala:10-109  ifb200::ForestParams
//   model save / load               IF/IsolationForestModelReadWrite.scala, IF/extended/...ReadWrite.scala,
//                                   IF/core/IsolationForestModelReadWriteUtils.scala (metadata JSON + Avro rows)
//
// A "Dataset" here is a dense matrix of feature vectors (what the featuresCol of the DataFrame holds); the
// column plumbing of Spark (schemas, column names) is outside the hot path and not reproduced beyond the
// param names.  `require` failures throw IllegalArgumentException with the reference's message text.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define IFBH_CLASS __attribute__((visibility("default")))

namespace ifb200 {

struct IFBH_CLASS IllegalArgumentException : std::invalid_argument {
    using std::invalid_argument::invalid_argument;
};
struct IFBH_CLASS IllegalStateException : std::logic_error {
    using std::logic_error::logic_error;
};

// Feature vectors of a Dataset: `rows` vectors of size `cols`, row-major.  Exactly one of f64 / f32 is set.
// f64 is what Spark's ml.linalg.Vector holds; the engine casts to f32 (`.toFloat`) like the reference
// (IF/IsolationForest.scala:54, IF/IsolationForestModel.scala:136).
// A column of SparseVectors is passed as CSR (indptr[rows + 1], ascending indices per row, f64 values as Spark
// stores them); absent entries are 0.0, exactly what Vector.apply / toArray give the reference.
struct FeatureMatrix {
    int64_t rows = 0;
    int32_t cols = 0;
    const double *f64 = nullptr;          // dense, row-major rows x cols
    const float *f32 = nullptr;           // dense, row-major rows x cols
    const int64_t *csr_indptr = nullptr;  // sparse: exactly one of f64 / f32 / csr_indptr is set
    const int32_t *csr_indices = nullptr;
    const double *csr_values = nullptr;
};

struct ScoredData {                      // the two columns transform appends
    std::vector<double> outlierScore;    // $(scoreCol)
    std::vector<double> predictedLabel;  // $(predictionCol)
};

// Persisted node tables of a forest (pre-order rows, the reference's NodeData / ExtendedNodeData).
struct ForestTables {
    bool extended = false;
    std::vector<int32_t> node_off, left, right, feature;
    std::vector<int64_t> num_instances;
    std::vector<double> threshold, offset;
    std::vector<int64_t> hp_off;
    std::vector<int32_t> hp_idx;
    std::vector<float> hp_w;
    int32_t num_trees() const { return node_off.empty() ? 0 : (int32_t)node_off.size() - 1; }
};

// IsolationForestParamsBase (+ extensionLevel of ExtendedIsolationForestParams).
class IFBH_CLASS ForestParams {
   public:
    ForestParams &setNumEstimators(int v);
    ForestParams &setMaxSamples(double v);
    ForestParams &setContamination(double v);
    ForestParams &setContaminationError(double v);
    ForestParams &setMaxFeatures(double v);
    ForestParams &setBootstrap(bool v);
    ForestParams &setRandomSeed(int64_t v);
    ForestParams &setFeaturesCol(const std::string &v);
    ForestParams &setPredictionCol(const std::string &v);
    ForestParams &setScoreCol(const std::string &v);
    ForestParams &setExtensionLevel(int v);  // extended estimators/models only
    int getNumEstimators() const { return numEstimators; }
    double getMaxSamples() const { return maxSamples; }
    double getContamination() const { return contamination; }
    double getContaminationError() const { return contaminationError; }
    double getMaxFeatures() const { return maxFeatures; }
    bool getBootstrap() const { return bootstrap; }
    int64_t getRandomSeed() const { return randomSeed; }
    const std::string &getFeaturesCol() const { return featuresCol; }
    const std::string &getPredictionCol() const { return predictionCol; }
    const std::string &getScoreCol() const { return scoreCol; }
    bool isSetExtensionLevel() const { return extensionLevelSet; }
    int getExtensionLevel() const;  // throws NoSuchElement-like IllegalStateException when unset
    // engine parameters (no reference counterpart): device ordinal and the P of the tree-seed formula
    ForestParams &setDevice(int v) { device = v; return *this; }
    ForestParams &setNumPartitions(int v) { numPartitions = v; return *this; }
    int getDevice() const { return device; }
    int getNumPartitions() const { return numPartitions; }

    // generic access by Spark param name (used by the flat C API and by persistence)
    void setByName(const std::string &name, const std::string &json_value);
    std::string paramMapJson(bool extended) const;
    // Params.isSet(param): explicitly set through a setter (as opposed to carrying its default)
    bool isSet(const std::string &name) const;

   protected:
    std::string owner = "isolation-forest";
    int numEstimators = 100;
    double maxSamples = 256.0;
    double contamination = 0.0;
    double contaminationError = 0.0;
    double maxFeatures = 1.0;
    bool bootstrap = false;
    int64_t randomSeed = 1;
    std::string featuresCol = "features", predictionCol = "predictedLabel", scoreCol = "outlierScore";
    int extensionLevel = 0;
    bool extensionLevelSet = false;
    int device = 0;
    int numPartitions = 1;
    unsigned explicitlySet = 0;   // bit per Spark param, see paramBit() in host/model.cpp
    friend class ForestModelBase;
    friend class ForestEstimatorBase;
};

class IFBH_CLASS ForestModelBase : public ForestParams {
   public:
    virtual ~ForestModelBase();
    ForestModelBase(const ForestModelBase &) = delete;
    ForestModelBase &operator=(const ForestModelBase &) = delete;

    const std::string &uid() const { return uid_; }
    int getNumSamples() const { return numSamples_; }
    int getNumFeatures() const { return numFeatures_; }
    int getTotalNumFeatures() const { return totalNumFeatures_; }
    bool hasKnownTotalNumFeatures() const { return totalNumFeatures_ != -1; }
    double getOutlierScoreThreshold() const { return outlierScoreThreshold_; }
    void setOutlierScoreThreshold(double v);
    int numTrees() const { return tables_.num_trees(); }
    const ForestTables &tables() const { return tables_; }
    bool extended() const { return tables_.extended; }

    // Model.transform: appends scoreCol then predictionCol (IF/IsolationForestModel.scala:116-151)
    ScoredData transform(const FeatureMatrix &data) const;
    // MLWritable.write.save(path) / MLWriter.overwrite semantics: fails if the path exists unless overwrite
    void save(const std::string &path, bool overwrite = false) const;
    // tree `t` rendered like the reference's Node.toString (IF/Nodes.scala:32,63-65)
    std::string treeToString(int t) const;

   protected:
    ForestModelBase(bool extended, std::string uid, ForestTables tables, int numSamples, int numFeatures,
                    int totalNumFeatures, int device);
    void *native() const;  // ifb_forest*, created lazily
    std::string uid_;
    ForestTables tables_;
    int numSamples_, numFeatures_, totalNumFeatures_;
    double outlierScoreThreshold_ = -1.0;
    mutable void *handle_ = nullptr;
    friend class ForestEstimatorBase;
};

class IFBH_CLASS IsolationForestModel final : public ForestModelBase {
   public:
    static constexpr int UnknownTotalNumFeatures = -1;
    // new IsolationForestModel(uid, trees, numSamples, numFeatures[, totalNumFeatures])
    IsolationForestModel(std::string uid, ForestTables trees, int numSamples, int numFeatures,
                         int totalNumFeatures = UnknownTotalNumFeatures, int device = 0);
    static std::unique_ptr<IsolationForestModel> load(const std::string &path, int device = 0);
};

class IFBH_CLASS ExtendedIsolationForestModel final : public ForestModelBase {
   public:
    ExtendedIsolationForestModel(std::string uid, ForestTables trees, int numSamples, int numFeatures,
                                 int totalNumFeatures, int device = 0);
    static std::unique_ptr<ExtendedIsolationForestModel> load(const std::string &path, int device = 0);
};

class IFBH_CLASS ForestEstimatorBase : public ForestParams {
   public:
    const std::string &uid() const { return uid_; }
    // DefaultParamsWritable: estimator.write[.overwrite()].save(path) -> path/metadata/part-00000 holding
    // {class, timestamp, sparkVersion, uid, paramMap (explicitly set params), defaultParamMap}
    void save(const std::string &path, bool overwrite = false) const;

   protected:
    ForestEstimatorBase(bool extended, std::string uid);
    std::unique_ptr<ForestModelBase> fitImpl(const FeatureMatrix &data) const;
    bool extended_;
    std::string uid_;
};

class IFBH_CLASS IsolationForest final : public ForestEstimatorBase {
   public:
    IsolationForest();                           // Identifiable.randomUID("isolation-forest")
    explicit IsolationForest(std::string uid);
    std::unique_ptr<IsolationForestModel> fit(const FeatureMatrix &data) const;   // IF/IsolationForest.scala:46
    static std::unique_ptr<IsolationForest> load(const std::string &path);       // DefaultParamsReadable, :114
};

class IFBH_CLASS ExtendedIsolationForest final : public ForestEstimatorBase {
   public:
    ExtendedIsolationForest();                   // randomUID("extended-isolation-forest")
    explicit ExtendedIsolationForest(std::string uid);
    std::unique_ptr<ExtendedIsolationForestModel> fit(const FeatureMatrix &data) const;  // extended/...:40
    static std::unique_ptr<ExtendedIsolationForest> load(const std::string &path);       // extended/...:125
};

// resolved numFeatures / numSamples (validateAndResolveParams, IF/core/SharedTrainLogic.scala:27-78)
struct ResolvedParams {
    int numFeatures, totalNumFeatures, numSamples;
    int64_t totalNumSamples;
};
IFBH_CLASS ResolvedParams validateAndResolveParams(int64_t totalNumSamples, int totalNumFeatures, double maxFeatures,
                                        double maxSamples);

}  // namespace ifb200

// ---- flat C API over the classes above (what the Python mirror binds with ctypes) ----------------------
extern "C" {
#define IFBH_API __attribute__((visibility("default")))
IFBH_API const char *ifbh_last_error(void);
IFBH_API int ifbh_last_error_kind(void);  // 1 IllegalArgumentException, 2 IllegalStateException, 3 other
IFBH_API int ifbh_estimator_create(int extended, const char *uid_or_null, void **out);
IFBH_API int ifbh_estimator_destroy(void *est);
IFBH_API int ifbh_estimator_save(void *est, const char *path, int overwrite);
IFBH_API int ifbh_estimator_load(int extended, const char *path, void **est_out);
// JSON {uid, paramMap (every param with a value), set: [names explicitly set]} ; returns the length needed
IFBH_API int64_t ifbh_estimator_describe(void *est, char *buf, int64_t cap);
IFBH_API int ifbh_estimator_set(void *est, const char *param, const char *json_value);
IFBH_API int ifbh_estimator_fit(void *est, const double *x_f64, const float *x_f32, int64_t rows, int32_t cols,
                                void **model_out);
IFBH_API int ifbh_model_create(int extended, const char *uid, int32_t num_trees, const int32_t *node_off,
                               const int32_t *left, const int32_t *right, const int32_t *feature,
                               const double *threshold, const int64_t *num_instances, const double *offset,
                               const int64_t *hp_off, const int32_t *hp_idx, const float *hp_w, int32_t num_samples,
                               int32_t num_features, int32_t total_num_features, void **model_out);
IFBH_API int ifbh_model_destroy(void *model);
IFBH_API int ifbh_model_set(void *model, const char *param, const char *json_value);
IFBH_API int ifbh_model_transform(void *model, const double *x_f64, const float *x_f32, int64_t rows, int32_t cols,
                                  double *scores, double *predictions);
// the same two calls for a SparseVector column (CSR: indptr[rows + 1], indices, f64 values)
IFBH_API int ifbh_estimator_fit_csr(void *est, const int64_t *indptr, const int32_t *indices, const double *values,
                                    int64_t rows, int32_t cols, void **model_out);
IFBH_API int ifbh_model_transform_csr(void *model, const int64_t *indptr, const int32_t *indices, const double *values,
                                      int64_t rows, int32_t cols, double *scores, double *predictions);
IFBH_API int ifbh_model_save(void *model, const char *path, int overwrite);
IFBH_API int ifbh_model_load(int extended, const char *path, void **model_out);
// JSON description: uid, class, paramMap, numSamples, numFeatures, totalNumFeatures, outlierScoreThreshold,
// numTrees, numNodes, numHpEntries.  Returns the needed size; copies at most `cap` bytes.
IFBH_API int64_t ifbh_model_describe(void *model, char *buf, int64_t cap);
IFBH_API int ifbh_model_tables(void *model, int32_t *node_off, int32_t *left, int32_t *right, int32_t *feature,
                               double *threshold, int64_t *num_instances, double *offset, int64_t *hp_off,
                               int32_t *hp_idx, float *hp_w);
IFBH_API int64_t ifbh_model_tree_string(void *model, int32_t tree, char *buf, int64_t cap);
}

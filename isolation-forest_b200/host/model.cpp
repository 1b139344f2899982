// Host mirror of the reference's Estimator / Model classes (see include/ifb200_host.hpp for the mapping).
// All numeric work goes through the C ABI of libifb200.so; this file is orchestration only.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <functional>
#include <mutex>
#include <fstream>
#include <random>
#include <sstream>
#include <thread>

#include "host_internal.hpp"
#include "ifb200.h"

namespace ifb200 {

namespace {

const char *kStdModelClass = "com.linkedin.relevance.isolationforest.IsolationForestModel";
const char *kExtModelClass = "com.linkedin.relevance.isolationforest.extended.ExtendedIsolationForestModel";

void require(bool cond, const std::string &msg) {
    if (!cond) throw IllegalArgumentException(msg);
}

// status of the C ABI -> the exception the JVM glue would raise (include/ifb200.h error contract)
void check(int rc) {
    if (rc == IFB_OK) return;
    const std::string msg = ifb_last_error();
    if (rc == IFB_EINVAL) throw IllegalArgumentException(msg);
    if (rc == IFB_ESTATE) throw IllegalStateException(msg);
    throw std::runtime_error(msg);
}

void logWarning(const std::string &m) { fprintf(stderr, "WARN ifb200: %s\n", m.c_str()); }

// Identifiable.randomUID(prefix): prefix + "_" + 12 hex digits
std::string randomUID(const std::string &prefix) {
    std::random_device rd;
    std::mt19937_64 gen(((uint64_t)rd() << 32) ^ rd() ^ (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count());
    char buf[32];
    snprintf(buf, sizeof buf, "%012llx", (unsigned long long)(gen() & 0xffffffffffffULL));
    return prefix + "_" + buf;
}

std::string fmtDouble(double v) { return json::javaDouble(v); }

// Pinned staging buffers are expensive to create (cudaHostAlloc pins pages): keep a small process-wide pool.
class PinnedPool {
   public:
    static PinnedPool &get() {
        static PinnedPool p;
        return p;
    }
    void *acquire(size_t bytes, size_t *cap) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].second >= bytes && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
            if (best != free_.size()) {
                auto e = free_[best];
                free_.erase(free_.begin() + (long)best);
                cached_ -= e.second;
                *cap = e.second;
                return e.first;
            }
        }
        void *p = nullptr;
        const size_t want = std::max<size_t>(bytes, 1 << 20);
        check(ifb_host_alloc(want, &p));
        *cap = want;
        return p;
    }
    void release(void *p, size_t cap) {
        std::lock_guard<std::mutex> lk(mu_);
        if (cached_ + cap > kMaxCached) {
            ifb_host_free(p);
            return;
        }
        free_.emplace_back(p, cap);
        cached_ += cap;
    }

   private:
    static constexpr size_t kMaxCached = (size_t)3 << 30;
    std::mutex mu_;
    std::vector<std::pair<void *, size_t>> free_;
    size_t cached_ = 0;
};

struct Pinned {
    void *p = nullptr;
    size_t cap = 0;
    explicit Pinned(size_t bytes) { p = PinnedPool::get().acquire(bytes, &cap); }
    ~Pinned() { PinnedPool::get().release(p, cap); }
    Pinned(const Pinned &) = delete;
};

// f64 -> f32 (`.toFloat`, round to nearest even) of rows [r0, r1) into a staging buffer, split over threads
void castRows(const FeatureMatrix &m, int64_t r0, int64_t r1, float *dst) {
    const int64_t off = r0 * (int64_t)m.cols, total = (r1 - r0) * (int64_t)m.cols;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nt = (int)std::min<int64_t>(std::max<int64_t>(1, total / (1 << 20)), std::min(32u, hw));
    std::vector<std::thread> th;
    if (m.csr_indptr) {   // SparseVector rows: zero-fill, then scatter the stored entries (`.toFloat` each)
        for (int t = 0; t < nt; t++) {
            const int64_t ra = r0 + (r1 - r0) * t / nt, rb = r0 + (r1 - r0) * (t + 1) / nt;
            th.emplace_back([=]() {
                std::memset(dst + (ra - r0) * m.cols, 0, (size_t)(rb - ra) * m.cols * 4);
                for (int64_t r = ra; r < rb; r++) {
                    float *row = dst + (r - r0) * m.cols;
                    for (int64_t e = m.csr_indptr[r]; e < m.csr_indptr[r + 1]; e++) row[m.csr_indices[e]] = (float)m.csr_values[e];
                }
            });
        }
        for (auto &x : th) x.join();
        return;
    }
    for (int t = 0; t < nt; t++) {
        const int64_t a = total * t / nt, b = total * (t + 1) / nt;
        if (m.f32) {
            th.emplace_back([=]() { std::memcpy(dst + a, m.f32 + off + a, (size_t)(b - a) * 4); });
        } else {
            th.emplace_back([=]() {
                const double *src = m.f64 + off;
                for (int64_t i = a; i < b; i++) dst[i] = (float)src[i];
            });
        }
    }
    for (auto &x : th) x.join();
}

void checkMatrix(const FeatureMatrix &m) {
    require(m.rows >= 0 && m.cols >= 1, "feature matrix must have at least one column");
    const int given = (m.f64 != nullptr) + (m.f32 != nullptr) + (m.csr_indptr != nullptr);
    require(given == 1 || (m.rows == 0 && given == 0), "exactly one of f64 / f32 / csr must be given");
    if (m.csr_indptr) {
        require(m.csr_indptr[0] == 0, "csr_indptr[0] must be 0");
        for (int64_t r = 0; r < m.rows; r++) {
            const int64_t a = m.csr_indptr[r], b = m.csr_indptr[r + 1];
            require(b >= a, "csr_indptr must be non-decreasing");
            require(b == a || (m.csr_indices && m.csr_values), "null csr arrays");
            for (int64_t e = a; e < b; e++) {
                // SparseVector invariants: indices in [0, size), strictly increasing
                require(m.csr_indices[e] >= 0 && m.csr_indices[e] < m.cols,
                        "sparse index " + std::to_string(m.csr_indices[e]) + " outside [0, " + std::to_string(m.cols) + ")");
                require(e == a || m.csr_indices[e] > m.csr_indices[e - 1], "sparse indices must be strictly increasing");
            }
        }
    }
}

}  // namespace

// ---- params (IF/core/IsolationForestParamsBase.scala:10-109) ---------------------------------------
namespace {
std::string invalid(const std::string &owner, const char *name, const std::string &v) {
    // Params.validate: s"$parent parameter $name given invalid value $value."
    return owner + " parameter " + name + " given invalid value " + v + ".";
}
}  // namespace

namespace {
const char *const kParamNames[] = {"numEstimators", "maxSamples", "contamination", "contaminationError", "maxFeatures", "bootstrap",
                                   "randomSeed", "featuresCol", "predictionCol", "scoreCol", "extensionLevel"};
unsigned paramBit(const std::string &name) {
    for (unsigned i = 0; i < sizeof kParamNames / sizeof *kParamNames; i++)
        if (name == kParamNames[i]) return 1u << i;
    return 0;
}
}  // namespace
bool ForestParams::isSet(const std::string &name) const { return (explicitlySet & paramBit(name)) != 0; }

ForestParams &ForestParams::setNumEstimators(int v) {
    require(v > 0, invalid(owner, "numEstimators", std::to_string(v)));
    numEstimators = v;
    explicitlySet |= paramBit("numEstimators");
    return *this;
}
ForestParams &ForestParams::setMaxSamples(double v) {
    require(v > 0.0, invalid(owner, "maxSamples", fmtDouble(v)));
    maxSamples = v;
    explicitlySet |= paramBit("maxSamples");
    return *this;
}
ForestParams &ForestParams::setContamination(double v) {
    require(v >= 0.0 && v < 0.5, invalid(owner, "contamination", fmtDouble(v)));
    contamination = v;
    explicitlySet |= paramBit("contamination");
    return *this;
}
ForestParams &ForestParams::setContaminationError(double v) {
    require(v >= 0.0 && v <= 1.0, invalid(owner, "contaminationError", fmtDouble(v)));
    contaminationError = v;
    explicitlySet |= paramBit("contaminationError");
    return *this;
}
ForestParams &ForestParams::setMaxFeatures(double v) {
    require(v > 0.0, invalid(owner, "maxFeatures", fmtDouble(v)));
    maxFeatures = v;
    explicitlySet |= paramBit("maxFeatures");
    return *this;
}
ForestParams &ForestParams::setBootstrap(bool v) {
    bootstrap = v;
    explicitlySet |= paramBit("bootstrap");
    return *this;
}
ForestParams &ForestParams::setRandomSeed(int64_t v) {
    require(v > 0, invalid(owner, "randomSeed", std::to_string(v)));
    randomSeed = v;
    explicitlySet |= paramBit("randomSeed");
    return *this;
}
ForestParams &ForestParams::setFeaturesCol(const std::string &v) { featuresCol = v; explicitlySet |= paramBit("featuresCol"); return *this; }
ForestParams &ForestParams::setPredictionCol(const std::string &v) { predictionCol = v; explicitlySet |= paramBit("predictionCol"); return *this; }
ForestParams &ForestParams::setScoreCol(const std::string &v) { scoreCol = v; explicitlySet |= paramBit("scoreCol"); return *this; }
ForestParams &ForestParams::setExtensionLevel(int v) {
    require(v >= 0, invalid(owner, "extensionLevel", std::to_string(v)));
    extensionLevel = v;
    extensionLevelSet = true;
    explicitlySet |= paramBit("extensionLevel");
    return *this;
}
int ForestParams::getExtensionLevel() const {
    if (!extensionLevelSet) throw IllegalStateException("Failed to find a default value for extensionLevel");
    return extensionLevel;
}

void ForestParams::setByName(const std::string &name, const std::string &json_value) {
    json::Value v = json::parse(json_value);
    auto num = [&]() {
        require(v.isNumber(), "parameter " + name + " expects a number");
        return v.asDouble();
    };
    if (name == "numEstimators") setNumEstimators((int)num());
    else if (name == "maxSamples") setMaxSamples(num());
    else if (name == "contamination") setContamination(num());
    else if (name == "contaminationError") setContaminationError(num());
    else if (name == "maxFeatures") setMaxFeatures(num());
    else if (name == "bootstrap") { require(v.kind == json::Value::Bool, "parameter bootstrap expects a boolean"); setBootstrap(v.b); }
    else if (name == "randomSeed") setRandomSeed((int64_t)(v.kind == json::Value::Int ? v.i : (long long)num()));
    else if (name == "featuresCol") setFeaturesCol(v.s);
    else if (name == "predictionCol") setPredictionCol(v.s);
    else if (name == "scoreCol") setScoreCol(v.s);
    else if (name == "extensionLevel") setExtensionLevel((int)num());
    else if (name == "device") setDevice((int)num());
    else if (name == "numPartitions") setNumPartitions((int)num());
    else throw IllegalArgumentException("Param " + name + " does not exist.");
}

std::string ForestParams::paramMapJson(bool ext) const {
    json::Value m = json::mkObject();
    m.obj.emplace_back("randomSeed", json::mkInt(randomSeed));
    m.obj.emplace_back("scoreCol", json::mkString(scoreCol));
    m.obj.emplace_back("contamination", json::mkDouble(contamination));
    m.obj.emplace_back("maxFeatures", json::mkDouble(maxFeatures));
    m.obj.emplace_back("contaminationError", json::mkDouble(contaminationError));
    m.obj.emplace_back("featuresCol", json::mkString(featuresCol));
    m.obj.emplace_back("bootstrap", json::mkBool(bootstrap));
    if (ext && extensionLevelSet) m.obj.emplace_back("extensionLevel", json::mkInt(extensionLevel));
    m.obj.emplace_back("predictionCol", json::mkString(predictionCol));
    m.obj.emplace_back("numEstimators", json::mkInt(numEstimators));
    m.obj.emplace_back("maxSamples", json::mkDouble(maxSamples));
    return json::render(m);
}

// ---- validateAndResolveParams (IF/core/SharedTrainLogic.scala:27-78) -------------------------------
ResolvedParams validateAndResolveParams(int64_t totalNumSamples, int totalNumFeatures, double maxFeatures,
                                        double maxSamples) {
    const int numFeatures = maxFeatures > 1.0 ? (int)std::floor(maxFeatures) : (int)std::floor(maxFeatures * totalNumFeatures);
    require(numFeatures > 0, "parameter maxFeatures given invalid value " + fmtDouble(maxFeatures) + " specifying the use of " +
                                 std::to_string(numFeatures) + " features, but >0 features are required.");
    require(numFeatures <= totalNumFeatures,
            "parameter maxFeatures given invalid value " + fmtDouble(maxFeatures) + " specifying the use of " +
                std::to_string(numFeatures) + " features, but only " + std::to_string(totalNumFeatures) + " features are available.");
    const double ns = maxSamples > 1.0 ? std::floor(maxSamples) : std::floor(maxSamples * (double)totalNumSamples);
    const int numSamples = (int)std::min<double>(ns, 2147483647.0);
    require(numSamples >= 2, "parameter maxSamples given invalid value " + fmtDouble(maxSamples) + " specifying the use of " +
                                 std::to_string(numSamples) + " samples, but >=2 samples are required.");
    require((int64_t)numSamples <= totalNumSamples,
            "parameter maxSamples given invalid value " + fmtDouble(maxSamples) + " specifying the use of " +
                std::to_string(numSamples) + " samples, but only " + std::to_string(totalNumSamples) +
                " samples are in the input dataset.");
    return ResolvedParams{numFeatures, totalNumFeatures, numSamples, totalNumSamples};
}

// ---- models ---------------------------------------------------------------------------------------
ForestModelBase::ForestModelBase(bool extended, std::string uid, ForestTables tables, int numSamples, int numFeatures,
                                 int totalNumFeatures, int dev)
    : uid_(std::move(uid)), tables_(std::move(tables)), numSamples_(numSamples), numFeatures_(numFeatures),
      totalNumFeatures_(totalNumFeatures) {
    owner = uid_;
    device = dev;
    tables_.extended = extended;
    if (tables_.node_off.empty()) tables_.node_off.push_back(0);
    if (extended && tables_.hp_off.empty()) tables_.hp_off.push_back(0);
    // constructor requires: IF/IsolationForestModel.scala:61-78, IF/extended/ExtendedIsolationForestModel.scala:35-52
    require(numSamples > 0, "parameter numSamples must be >0, but given invalid value " + std::to_string(numSamples));
    require(numFeatures > 0, "parameter numFeatures must be >0, but given invalid value " + std::to_string(numFeatures));
    if (extended) {
        require(totalNumFeatures > 0,
                "parameter totalNumFeatures must be >0, but given invalid value " + std::to_string(totalNumFeatures));
        require(numFeatures <= totalNumFeatures, "parameter numFeatures must be <= totalNumFeatures, but given invalid values numFeatures=" +
                                                     std::to_string(numFeatures) + ", totalNumFeatures=" + std::to_string(totalNumFeatures));
    } else {
        require(totalNumFeatures == -1 || totalNumFeatures > 0,
                "parameter totalNumFeatures must be >0 or UnknownTotalNumFeatures, but given invalid value " +
                    std::to_string(totalNumFeatures));
        require(totalNumFeatures == -1 || numFeatures <= totalNumFeatures,
                "parameter numFeatures must be <= totalNumFeatures, but given invalid values numFeatures=" +
                    std::to_string(numFeatures) + ", totalNumFeatures=" + std::to_string(totalNumFeatures));
    }
}

ForestModelBase::~ForestModelBase() {
    if (handle_) ifb_forest_destroy((ifb_forest *)handle_);
}

void ForestModelBase::setOutlierScoreThreshold(double value) {
    require(value == -1 || (value >= 0 && value <= 1),
            "parameter outlierScoreThreshold must be equal to -1 (no threshold) or be in the range [0, 1], but given invalid value " +
                fmtDouble(value));
    outlierScoreThreshold_ = value;
}

void *ForestModelBase::native() const {
    if (handle_) return handle_;
    const ForestTables &t = tables_;
    ifb_forest *f = nullptr;
    if (t.extended)
        check(ifb_forest_create_extended(device, t.num_trees(), t.node_off.data(), t.left.data(), t.right.data(),
                                         t.num_instances.data(), t.offset.data(), t.hp_off.data(), t.hp_idx.data(),
                                         t.hp_w.data(), numSamples_, totalNumFeatures_, &f));
    else
        check(ifb_forest_create_standard(device, t.num_trees(), t.node_off.data(), t.left.data(), t.right.data(),
                                         t.feature.data(), t.threshold.data(), t.num_instances.data(), numSamples_,
                                         totalNumFeatures_, &f));
    handle_ = f;
    return handle_;
}

ScoredData ForestModelBase::transform(const FeatureMatrix &data) const {
    // IF/IsolationForestModel.scala:118-125 / IF/extended/ExtendedIsolationForestModel.scala:100-107
    require(numSamples_ >= 2, "Cannot score with numSamples=" + std::to_string(numSamples_) + "; expected numSamples >= 2.");
    require(numTrees() > 0, extended() ? "Cannot score with an empty ExtendedIsolationForestModel."
                                       : "Cannot score with an empty IsolationForestModel.");
    checkMatrix(data);
    ScoredData out;
    out.outlierScore.resize((size_t)data.rows);
    out.predictedLabel.assign((size_t)data.rows, 0.0);
    if (data.rows == 0) return out;
    ifb_forest *f = (ifb_forest *)native();
    // batches of rows: while batch k is on the GPU (ifb_score_host pipelines its own H2D / kernel / D2H sub-chunks),
    // a helper thread casts batch k+1 (`.toFloat`) into the other pinned staging buffer
    const int64_t batch = std::max<int64_t>(1 << 16, std::min<int64_t>(data.rows, ((int64_t)256 << 20) / (4LL * data.cols)));
    const int64_t n_batches = (data.rows + batch - 1) / batch;
    Pinned stage0((size_t)std::min(batch, data.rows) * data.cols * 4);
    std::unique_ptr<Pinned> stage1;
    if (n_batches > 1) stage1 = std::make_unique<Pinned>((size_t)batch * data.cols * 4);
    Pinned sc((size_t)std::min(batch, data.rows) * 8);
    auto buf = [&](int64_t k) { return (float *)((k & 1) ? stage1->p : stage0.p); };
    castRows(data, 0, std::min(batch, data.rows), buf(0));
    for (int64_t k = 0; k < n_batches; k++) {
        const int64_t r0 = k * batch, r1 = std::min(data.rows, r0 + batch);
        std::thread next;
        if (k + 1 < n_batches)
            next = std::thread([&, k]() { castRows(data, (k + 1) * batch, std::min(data.rows, (k + 2) * batch), buf(k + 1)); });
        int rc = ifb_score_host(f, buf(k), r1 - r0, data.cols, data.cols, IFB_ROW_MAJOR, (double *)sc.p, nullptr, nullptr);
        if (next.joinable()) next.join();
        check(rc);
        std::memcpy(out.outlierScore.data() + r0, sc.p, (size_t)(r1 - r0) * 8);
    }
    if (outlierScoreThreshold_ > 0)  // :143-148
        for (int64_t i = 0; i < data.rows; i++) out.predictedLabel[i] = out.outlierScore[i] >= outlierScoreThreshold_ ? 1.0 : 0.0;
    return out;
}

std::string ForestModelBase::treeToString(int t) const {
    require(t >= 0 && t < numTrees(), "tree index out of range");
    const ForestTables &tb = tables_;
    const int32_t base = tb.node_off[t];
    std::string out;
    // iterative pre-order rendering (ids are pre-order, so a recursive descent on child ids suffices)
    std::function<void(int32_t)> rec = [&](int32_t i) {
        const int64_t g = (int64_t)base + i;
        if (tb.left[g] == -1) {
            out += (tb.extended ? "ExtendedExternalNode(numInstances = " : "ExternalNode(numInstances = ") +
                   std::to_string(tb.num_instances[g]) + ")";
            return;
        }
        if (tb.extended) {
            out += "ExtendedInternalNode(splitHyperplane = SplitHyperplane(indices = (";
            for (int64_t q = tb.hp_off[g]; q < tb.hp_off[g + 1]; q++) out += (q > tb.hp_off[g] ? ", " : "") + std::to_string(tb.hp_idx[q]);
            out += "), weights = (";
            for (int64_t q = tb.hp_off[g]; q < tb.hp_off[g + 1]; q++) out += (q > tb.hp_off[g] ? ", " : "") + json::javaFloat(tb.hp_w[q]);
            out += "), offset = " + json::javaDouble(tb.offset[g]) + "), leftChild = (";
        } else {
            out += "InternalNode(splitAttribute = " + std::to_string(tb.feature[g]) + ", splitValue = " +
                   json::javaDouble(tb.threshold[g]) + ", leftChild = (";
        }
        rec(tb.left[g]);
        out += "), rightChild = (";
        rec(tb.right[g]);
        out += "))";
    };
    rec(0);
    return out;
}

// ---- persistence (IF/IsolationForestModelReadWrite.scala:210-324, IF/core/...ReadWriteUtils.scala:97-187) ----
void ForestModelBase::save(const std::string &path, bool overwrite) const {
    namespace fs = std::filesystem;
    if (fs::exists(path)) {
        if (!overwrite)  // MLWriter.save: "Path ... already exists. To overwrite it, please use write.overwrite().save(path)"
            throw std::runtime_error("Path " + path + " already exists. To overwrite it, please use write.overwrite().save(path) for Scala and use write().overwrite().save(path) for Java and Python.");
        fs::remove_all(path);
    }
    fs::create_directories(fs::path(path) / "metadata");
    json::Value meta = json::mkObject();
    meta.obj.emplace_back("class", json::mkString(extended() ? kExtModelClass : kStdModelClass));
    meta.obj.emplace_back("timestamp", json::mkInt((long long)std::chrono::duration_cast<std::chrono::milliseconds>(
                                                        std::chrono::system_clock::now().time_since_epoch()).count()));
    meta.obj.emplace_back("sparkVersion", json::mkString("3.5.5"));  // format level written (spark-avro 3.5 layout)
    meta.obj.emplace_back("uid", json::mkString(uid_));
    meta.obj.emplace_back("paramMap", json::parse(paramMapJson(extended())));
    meta.obj.emplace_back("outlierScoreThreshold", json::mkDouble(outlierScoreThreshold_));
    meta.obj.emplace_back("numSamples", json::mkInt(numSamples_));
    meta.obj.emplace_back("numFeatures", json::mkInt(numFeatures_));
    meta.obj.emplace_back("totalNumFeatures", json::mkInt(totalNumFeatures_));
    {
        std::ofstream o(fs::path(path) / "metadata" / "part-00000");
        o << json::render(meta) << "\n";
    }
    std::ofstream(fs::path(path) / "metadata" / "_SUCCESS").close();
    avro::write_tables((fs::path(path) / "data").string(), tables_, "deflate");
}

namespace {
struct LoadedMeta {
    json::Value js;
    std::string uid;
};
LoadedMeta loadMetadata(const std::string &path, const std::string &expectedClass) {
    namespace fs = std::filesystem;
    const fs::path mdir = fs::path(path) / "metadata";
    std::vector<fs::path> parts;
    if (!fs::is_directory(mdir)) throw std::runtime_error("Input path does not exist: " + mdir.string());
    for (auto &e : fs::directory_iterator(mdir))
        if (e.path().filename().string().rfind("part-", 0) == 0) parts.push_back(e.path());
    std::sort(parts.begin(), parts.end());
    if (parts.empty()) throw std::runtime_error("no metadata part file under " + mdir.string());
    std::ifstream in(parts[0]);
    std::string line;
    std::getline(in, line);
    LoadedMeta m;
    m.js = json::parse(line);
    const json::Value *cls = m.js.get("class");
    require(cls && cls->kind == json::Value::String, "metadata has no class");
    require(cls->s == expectedClass, "Expected class " + expectedClass + ", but found " + cls->s);  // parseMetadata
    const json::Value *uid = m.js.get("uid");
    require(uid && uid->kind == json::Value::String, "metadata has no uid");
    m.uid = uid->s;
    return m;
}
}  // namespace

namespace {
template <typename M>
std::unique_ptr<M> loadModel(const std::string &path, bool extended, int device) {
    LoadedMeta m = loadMetadata(path, extended ? kExtModelClass : kStdModelClass);
    const json::Value &js = m.js;
    auto need = [&](const char *k) -> const json::Value & {
        const json::Value *v = js.get(k);
        require(v != nullptr, std::string("metadata field ") + k + " is missing");
        return *v;
    };
    const int numSamples = (int)need("numSamples").asInt();
    const int numFeatures = (int)need("numFeatures").asInt();
    int totalNumFeatures;
    if (extended) {
        totalNumFeatures = (int)need("totalNumFeatures").asInt();
    } else {
        const json::Value *tnf = js.get("totalNumFeatures");
        if (tnf) {
            totalNumFeatures = (int)tnf->asInt();
        } else {  // legacy layout (IF/IsolationForestModelReadWrite.scala:298-306)
            logWarning("Loading legacy IsolationForestModel from " + path +
                       " without totalNumFeatures metadata; feature-dimension validation will be unavailable for this model.");
            totalNumFeatures = -1;
        }
    }
    const double threshold = need("outlierScoreThreshold").asDouble();
    ForestTables t = avro::read_tables((std::filesystem::path(path) / "data").string(), extended);
    auto model = std::make_unique<M>(m.uid, std::move(t), numSamples, numFeatures, totalNumFeatures, device);
    if (const json::Value *pm = js.get("paramMap"))  // metadata.setParams(model)
        for (auto &kv : pm->obj) model->setByName(kv.first, json::render(kv.second));
    model->setOutlierScoreThreshold(threshold);
    return model;
}
}  // namespace

IsolationForestModel::IsolationForestModel(std::string uid, ForestTables trees, int numSamples, int numFeatures,
                                           int totalNumFeatures, int device)
    : ForestModelBase(false, std::move(uid), std::move(trees), numSamples, numFeatures, totalNumFeatures, device) {}
std::unique_ptr<IsolationForestModel> IsolationForestModel::load(const std::string &path, int device) {
    return loadModel<IsolationForestModel>(path, false, device);
}
ExtendedIsolationForestModel::ExtendedIsolationForestModel(std::string uid, ForestTables trees, int numSamples,
                                                           int numFeatures, int totalNumFeatures, int device)
    : ForestModelBase(true, std::move(uid), std::move(trees), numSamples, numFeatures, totalNumFeatures, device) {}
std::unique_ptr<ExtendedIsolationForestModel> ExtendedIsolationForestModel::load(const std::string &path, int device) {
    return loadModel<ExtendedIsolationForestModel>(path, true, device);
}

// ---- estimators (IF/IsolationForest.scala:46-105, IF/extended/ExtendedIsolationForest.scala:40-115) -------
ForestEstimatorBase::ForestEstimatorBase(bool extended, std::string uid) : extended_(extended), uid_(std::move(uid)) {
    owner = uid_;
}

namespace {
struct DeviceBuf {
    int dev;
    void *p = nullptr;
    DeviceBuf(int d, size_t bytes) : dev(d) { check(ifb_device_alloc(d, bytes, &p)); }
    ~DeviceBuf() { ifb_device_free(dev, p); }
    DeviceBuf(const DeviceBuf &) = delete;
};

ForestTables exportTables(ifb_forest *f) {
    ifb_forest_info info;
    check(ifb_forest_get_info(f, &info));
    ForestTables t;
    t.extended = info.extended != 0;
    const size_t n = (size_t)info.num_nodes;
    t.node_off.resize((size_t)info.num_trees + 1);
    t.left.resize(n);
    t.right.resize(n);
    t.num_instances.resize(n);
    if (t.extended) {
        t.offset.resize(n);
        t.hp_off.resize(n + 1);
        t.hp_idx.resize((size_t)info.num_hp_entries);
        t.hp_w.resize((size_t)info.num_hp_entries);
    } else {
        t.feature.resize(n);
        t.threshold.resize(n);
    }
    check(ifb_forest_export(f, t.node_off.data(), t.left.data(), t.right.data(), t.feature.data(), t.threshold.data(),
                            t.num_instances.data(), t.offset.data(), t.hp_off.data(), t.hp_idx.data(), t.hp_w.data()));
    return t;
}
}  // namespace

std::unique_ptr<ForestModelBase> ForestEstimatorBase::fitImpl(const FeatureMatrix &data) const {
    checkMatrix(data);
    require(data.rows >= 1, "The input dataset is empty.");
    // validateAndResolveParams (2 Spark jobs in the reference: head() and count())
    const ResolvedParams rp = validateAndResolveParams(data.rows, data.cols, maxFeatures, maxSamples);
    int resolvedExt = -1;
    if (extended_) {  // IF/extended/ExtendedIsolationForest.scala:57-69
        const int maxExt = rp.numFeatures - 1;
        if (extensionLevelSet) {
            require(extensionLevel <= maxExt, "parameter extensionLevel given invalid value " + std::to_string(extensionLevel) +
                                                  ", but must be in [0, " + std::to_string(maxExt) + "] for a subspace of " +
                                                  std::to_string(rp.numFeatures) + " features.");
            resolvedExt = extensionLevel;
        } else {
            resolvedExt = maxExt;
        }
    }
    // stage the training matrix on the device once: fit and (if contamination > 0) the threshold pass reuse it
    const size_t elems = (size_t)data.rows * data.cols;
    DeviceBuf dX(device, elems * 4);
    {
        const int64_t batch = std::max<int64_t>(1 << 16, std::min<int64_t>(data.rows, ((int64_t)256 << 20) / (4LL * data.cols)));
        Pinned stage((size_t)std::min(batch, data.rows) * data.cols * 4);
        for (int64_t r0 = 0; r0 < data.rows; r0 += batch) {
            const int64_t r1 = std::min(data.rows, r0 + batch);
            castRows(data, r0, r1, (float *)stage.p);
            check(ifb_copy_to_device(device, (float *)dX.p + r0 * data.cols, stage.p, (size_t)(r1 - r0) * data.cols * 4));
        }
    }
    ifb_fit_params fp;
    fp.num_estimators = numEstimators;
    fp.num_samples = rp.numSamples;
    fp.num_features = rp.numFeatures;
    fp.bootstrap = bootstrap ? 1 : 0;
    fp.random_seed = randomSeed;
    fp.num_partitions = numPartitions;
    fp.extension_level = resolvedExt;
    fp.tree_begin = 0;
    fp.tree_end = 0;
    ifb_forest *f = nullptr;
    check(ifb_fit_device(device, (const float *)dX.p, data.rows, data.cols, data.cols, IFB_ROW_MAJOR, &fp, &f, nullptr));
    ForestTables tables;
    try {
        tables = exportTables(f);
    } catch (...) {
        ifb_forest_destroy(f);
        throw;
    }
    std::unique_ptr<ForestModelBase> model;
    if (extended_) {
        auto m = std::make_unique<ExtendedIsolationForestModel>(uid_, std::move(tables), rp.numSamples, rp.numFeatures,
                                                                rp.totalNumFeatures, device);
        model = std::move(m);
    } else {
        auto m = std::make_unique<IsolationForestModel>(uid_, std::move(tables), rp.numSamples, rp.numFeatures,
                                                        rp.totalNumFeatures, device);
        model = std::move(m);
    }
    model->handle_ = f;  // reuse the forest the builder already uploaded
    // copyValues(model): the estimator's params carry over to the model
    model->numEstimators = numEstimators; model->maxSamples = maxSamples; model->contamination = contamination;
    model->contaminationError = contaminationError; model->maxFeatures = maxFeatures; model->bootstrap = bootstrap;
    model->randomSeed = randomSeed; model->featuresCol = featuresCol; model->predictionCol = predictionCol;
    model->scoreCol = scoreCol; model->numPartitions = numPartitions;
    if (extended_) model->setExtensionLevel(resolvedExt);  // :102 (the estimator itself is left untouched)

    // computeAndSetModelThreshold (IF/core/SharedTrainLogic.scala:175-242)
    if (contamination > 0.0) {
        DeviceBuf dS(device, (size_t)data.rows * 8);
        check(ifb_score_device(f, (const float *)dX.p, data.rows, data.cols, data.cols, IFB_ROW_MAJOR, (double *)dS.p, nullptr,
                               nullptr, nullptr));
        double thr = 0, observed = 0;
        // approxQuantile(scoreCol, [1 - contamination], contaminationError): the exact order statistic satisfies
        // every relative error, and is what the reference returns when contaminationError = 0
        check(ifb_quantile_device(device, (const double *)dS.p, data.rows, 1.0 - contamination, &thr, &observed, nullptr));
        model->setOutlierScoreThreshold(thr);
        const double verificationError = contaminationError == 0.0 ? contamination * 0.01 : contaminationError;
        if (std::fabs(observed - contamination) > verificationError)
            logWarning("Observed contamination is " + fmtDouble(observed) + ", which is outside the expected range of " +
                       fmtDouble(contamination) + " +/- " + fmtDouble(verificationError) +
                       ". If this is acceptable to you, then it is OK to proceed. If there is a very large discrepancy between"
                       " observed and expected values, then please try retraining the model with an exact threshold calculation"
                       " (set the contaminationError parameter value to 0.0).");
    }
    return model;
}

IsolationForest::IsolationForest() : ForestEstimatorBase(false, randomUID("isolation-forest")) {}
IsolationForest::IsolationForest(std::string uid) : ForestEstimatorBase(false, std::move(uid)) {}
std::unique_ptr<IsolationForestModel> IsolationForest::fit(const FeatureMatrix &data) const {
    std::unique_ptr<ForestModelBase> m = fitImpl(data);
    return std::unique_ptr<IsolationForestModel>(static_cast<IsolationForestModel *>(m.release()));
}
ExtendedIsolationForest::ExtendedIsolationForest() : ForestEstimatorBase(true, randomUID("extended-isolation-forest")) {}
ExtendedIsolationForest::ExtendedIsolationForest(std::string uid) : ForestEstimatorBase(true, std::move(uid)) {}

// ---- estimator persistence: Spark's DefaultParamsWriter / DefaultParamsReader layout ------------------
// (IF/IsolationForest.scala:25-28,114; IF/extended/ExtendedIsolationForest.scala:23-26,125)
namespace {
const char *kStdEstimatorClass = "com.linkedin.relevance.isolationforest.IsolationForest";
const char *kExtEstimatorClass = "com.linkedin.relevance.isolationforest.extended.ExtendedIsolationForest";
}  // namespace

void ForestEstimatorBase::save(const std::string &path, bool overwrite) const {
    namespace fs = std::filesystem;
    if (fs::exists(path)) {
        if (!overwrite)
            throw std::runtime_error("Path " + path + " already exists. To overwrite it, please use write.overwrite().save(path) for Scala and use write().overwrite().save(path) for Java and Python.");
        fs::remove_all(path);
    }
    fs::create_directories(fs::path(path) / "metadata");
    // paramMap holds the explicitly set params, defaultParamMap the defaults (a fresh instance of the same class)
    const json::Value all = json::parse(paramMapJson(extended_));
    const ForestEstimatorBase fresh(extended_, uid_);
    const json::Value defaults = json::parse(fresh.paramMapJson(extended_));
    json::Value setMap = json::mkObject();
    for (auto &kv : all.obj)
        if (isSet(kv.first)) setMap.obj.push_back(kv);
    json::Value meta = json::mkObject();
    meta.obj.emplace_back("class", json::mkString(extended_ ? kExtEstimatorClass : kStdEstimatorClass));
    meta.obj.emplace_back("timestamp", json::mkInt((long long)std::chrono::duration_cast<std::chrono::milliseconds>(
                                                        std::chrono::system_clock::now().time_since_epoch()).count()));
    meta.obj.emplace_back("sparkVersion", json::mkString("3.5.5"));
    meta.obj.emplace_back("uid", json::mkString(uid_));
    meta.obj.emplace_back("paramMap", setMap);
    meta.obj.emplace_back("defaultParamMap", defaults);
    {
        std::ofstream o(fs::path(path) / "metadata" / "part-00000");
        o << json::render(meta) << "\n";
    }
    std::ofstream(fs::path(path) / "metadata" / "_SUCCESS").close();
}

namespace {
template <typename E>
std::unique_ptr<E> loadEstimator(const std::string &path, bool extended) {
    LoadedMeta m = loadMetadata(path, extended ? kExtEstimatorClass : kStdEstimatorClass);
    auto est = std::make_unique<E>(m.uid);
    // DefaultParamsReader.getAndSetParams: paramMap entries are `set`; defaultParamMap entries only have to exist
    // as params of the class (their values are the class's own defaults)
    if (const json::Value *dm = m.js.get("defaultParamMap"))
        for (auto &kv : dm->obj)
            if (!paramBit(kv.first) || (kv.first == "extensionLevel" && !extended))
                throw IllegalArgumentException("Param " + kv.first + " does not exist.");
    if (const json::Value *pm = m.js.get("paramMap"))
        for (auto &kv : pm->obj) {
            if (kv.first == "extensionLevel" && !extended) throw IllegalArgumentException("Param extensionLevel does not exist.");
            est->setByName(kv.first, json::render(kv.second));
        }
    return est;
}
}  // namespace
std::unique_ptr<IsolationForest> IsolationForest::load(const std::string &path) {
    return loadEstimator<IsolationForest>(path, false);
}
std::unique_ptr<ExtendedIsolationForest> ExtendedIsolationForest::load(const std::string &path) {
    return loadEstimator<ExtendedIsolationForest>(path, true);
}
std::unique_ptr<ExtendedIsolationForestModel> ExtendedIsolationForest::fit(const FeatureMatrix &data) const {
    std::unique_ptr<ForestModelBase> m = fitImpl(data);
    return std::unique_ptr<ExtendedIsolationForestModel>(static_cast<ExtendedIsolationForestModel *>(m.release()));
}

}  // namespace ifb200

// ====================================== flat C API ==================================================
using namespace ifb200;

namespace {
thread_local std::string g_err;
thread_local int g_kind = 0;

template <typename F>
int guarded(F &&fn) {
    try {
        fn();
        return 0;
    } catch (const IllegalArgumentException &e) {
        g_err = e.what();
        g_kind = 1;
    } catch (const IllegalStateException &e) {
        g_err = e.what();
        g_kind = 2;
    } catch (const std::exception &e) {
        g_err = e.what();
        g_kind = 3;
    }
    return g_kind;
}

struct EstBox {
    bool extended;
    std::unique_ptr<IsolationForest> std_;
    std::unique_ptr<ExtendedIsolationForest> ext_;
    ForestParams &params() { return extended ? (ForestParams &)*ext_ : (ForestParams &)*std_; }
};

int64_t copyOut(const std::string &s, char *buf, int64_t cap) {
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size() + 1;
}
}  // namespace

extern "C" {

const char *ifbh_last_error(void) { return g_err.c_str(); }
int ifbh_last_error_kind(void) { return g_kind; }

int ifbh_estimator_create(int extended, const char *uid, void **out) {
    return guarded([&] {
        auto *b = new EstBox();
        b->extended = extended != 0;
        if (extended) b->ext_ = uid ? std::make_unique<ExtendedIsolationForest>(uid) : std::make_unique<ExtendedIsolationForest>();
        else b->std_ = uid ? std::make_unique<IsolationForest>(uid) : std::make_unique<IsolationForest>();
        *out = b;
    });
}
int ifbh_estimator_destroy(void *est) {
    delete (EstBox *)est;
    return 0;
}
int ifbh_estimator_set(void *est, const char *param, const char *json_value) {
    return guarded([&] {
        EstBox *b = (EstBox *)est;
        if (!b->extended && std::string(param) == "extensionLevel") throw IllegalArgumentException("Param extensionLevel does not exist.");
        b->params().setByName(param, json_value);
    });
}
int ifbh_estimator_save(void *est, const char *path, int overwrite) {
    return guarded([&] {
        EstBox *b = (EstBox *)est;
        if (b->extended) b->ext_->save(path, overwrite != 0);
        else b->std_->save(path, overwrite != 0);
    });
}
int ifbh_estimator_load(int extended, const char *path, void **est_out) {
    return guarded([&] {
        auto b = std::make_unique<EstBox>();
        b->extended = extended != 0;
        if (b->extended) b->ext_ = ExtendedIsolationForest::load(path);
        else b->std_ = IsolationForest::load(path);
        *est_out = b.release();
    });
}
int64_t ifbh_estimator_describe(void *est, char *buf, int64_t cap) {
    int64_t n = -1;
    guarded([&] {
        EstBox *b = (EstBox *)est;
        ForestParams &p = b->params();
        json::Value d = json::mkObject();
        d.obj.emplace_back("uid", json::mkString(b->extended ? b->ext_->uid() : b->std_->uid()));
        json::Value pm = json::parse(p.paramMapJson(b->extended));
        json::Value set = json::mkArray();
        for (auto &kv : pm.obj)
            if (p.isSet(kv.first)) set.arr.push_back(json::mkString(kv.first));
        d.obj.emplace_back("paramMap", pm);
        d.obj.emplace_back("set", set);
        n = copyOut(json::render(d), buf, cap);
    });
    return n;
}
int ifbh_estimator_fit(void *est, const double *x64, const float *x32, int64_t rows, int32_t cols, void **model_out) {
    return guarded([&] {
        EstBox *b = (EstBox *)est;
        FeatureMatrix m{rows, cols, x64, x32};
        ForestModelBase *model = b->extended ? (ForestModelBase *)b->ext_->fit(m).release() : (ForestModelBase *)b->std_->fit(m).release();
        *model_out = model;
    });
}
int ifbh_estimator_fit_csr(void *est, const int64_t *indptr, const int32_t *indices, const double *values, int64_t rows,
                           int32_t cols, void **model_out) {
    return guarded([&] {
        EstBox *b = (EstBox *)est;
        FeatureMatrix m{rows, cols, nullptr, nullptr, indptr, indices, values};
        ForestModelBase *model = b->extended ? (ForestModelBase *)b->ext_->fit(m).release() : (ForestModelBase *)b->std_->fit(m).release();
        *model_out = model;
    });
}
int ifbh_model_create(int extended, const char *uid, int32_t T, const int32_t *node_off, const int32_t *left,
                      const int32_t *right, const int32_t *feature, const double *threshold, const int64_t *ninst,
                      const double *offset, const int64_t *hp_off, const int32_t *hp_idx, const float *hp_w,
                      int32_t num_samples, int32_t num_features, int32_t total_num_features, void **model_out) {
    return guarded([&] {
        ForestTables t;
        t.extended = extended != 0;
        t.node_off.assign(node_off, node_off + T + 1);
        const size_t n = (size_t)node_off[T];
        t.left.assign(left, left + n);
        t.right.assign(right, right + n);
        t.num_instances.assign(ninst, ninst + n);
        if (extended) {
            t.offset.assign(offset, offset + n);
            t.hp_off.assign(hp_off, hp_off + n + 1);
            const size_t h = n ? (size_t)hp_off[n] : 0;
            t.hp_idx.assign(hp_idx, hp_idx + h);
            t.hp_w.assign(hp_w, hp_w + h);
        } else {
            t.feature.assign(feature, feature + n);
            t.threshold.assign(threshold, threshold + n);
        }
        ForestModelBase *m;
        if (extended) m = new ExtendedIsolationForestModel(uid, std::move(t), num_samples, num_features, total_num_features);
        else m = new IsolationForestModel(uid, std::move(t), num_samples, num_features, total_num_features);
        *model_out = m;
    });
}
int ifbh_model_destroy(void *model) {
    delete (ForestModelBase *)model;
    return 0;
}
int ifbh_model_set(void *model, const char *param, const char *json_value) {
    return guarded([&] {
        ForestModelBase *m = (ForestModelBase *)model;
        if (std::string(param) == "outlierScoreThreshold") m->setOutlierScoreThreshold(json::parse(json_value).asDouble());
        else m->setByName(param, json_value);
    });
}
int ifbh_model_transform(void *model, const double *x64, const float *x32, int64_t rows, int32_t cols, double *scores,
                         double *predictions) {
    return guarded([&] {
        ForestModelBase *m = (ForestModelBase *)model;
        ScoredData s = m->transform(FeatureMatrix{rows, cols, x64, x32});
        if (rows > 0) {
            std::memcpy(scores, s.outlierScore.data(), (size_t)rows * 8);
            if (predictions) std::memcpy(predictions, s.predictedLabel.data(), (size_t)rows * 8);
        }
    });
}
int ifbh_model_transform_csr(void *model, const int64_t *indptr, const int32_t *indices, const double *values,
                             int64_t rows, int32_t cols, double *scores, double *predictions) {
    return guarded([&] {
        ForestModelBase *m = (ForestModelBase *)model;
        ScoredData s = m->transform(FeatureMatrix{rows, cols, nullptr, nullptr, indptr, indices, values});
        if (rows > 0) {
            std::memcpy(scores, s.outlierScore.data(), (size_t)rows * 8);
            if (predictions) std::memcpy(predictions, s.predictedLabel.data(), (size_t)rows * 8);
        }
    });
}
int ifbh_model_save(void *model, const char *path, int overwrite) {
    return guarded([&] { ((ForestModelBase *)model)->save(path, overwrite != 0); });
}
int ifbh_model_load(int extended, const char *path, void **model_out) {
    return guarded([&] {
        *model_out = extended ? (ForestModelBase *)ExtendedIsolationForestModel::load(path).release()
                              : (ForestModelBase *)IsolationForestModel::load(path).release();
    });
}
int64_t ifbh_model_describe(void *model, char *buf, int64_t cap) {
    ForestModelBase *m = (ForestModelBase *)model;
    json::Value d = json::mkObject();
    d.obj.emplace_back("uid", json::mkString(m->uid()));
    d.obj.emplace_back("class", json::mkString(m->extended() ? kExtModelClass : kStdModelClass));
    d.obj.emplace_back("paramMap", json::parse(m->paramMapJson(m->extended())));
    d.obj.emplace_back("outlierScoreThreshold", json::mkDouble(m->getOutlierScoreThreshold()));
    d.obj.emplace_back("numSamples", json::mkInt(m->getNumSamples()));
    d.obj.emplace_back("numFeatures", json::mkInt(m->getNumFeatures()));
    d.obj.emplace_back("totalNumFeatures", json::mkInt(m->getTotalNumFeatures()));
    d.obj.emplace_back("numTrees", json::mkInt(m->numTrees()));
    d.obj.emplace_back("numNodes", json::mkInt((long long)m->tables().left.size()));
    d.obj.emplace_back("numHpEntries", json::mkInt((long long)m->tables().hp_idx.size()));
    return copyOut(json::render(d), buf, cap);
}
int ifbh_model_tables(void *model, int32_t *node_off, int32_t *left, int32_t *right, int32_t *feature, double *threshold,
                      int64_t *ninst, double *offset, int64_t *hp_off, int32_t *hp_idx, float *hp_w) {
    return guarded([&] {
        const ForestTables &t = ((ForestModelBase *)model)->tables();
        auto cp = [](auto *dst, const auto &v) {
            if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
        };
        cp(node_off, t.node_off); cp(left, t.left); cp(right, t.right); cp(ninst, t.num_instances);
        cp(feature, t.feature); cp(threshold, t.threshold); cp(offset, t.offset); cp(hp_off, t.hp_off);
        cp(hp_idx, t.hp_idx); cp(hp_w, t.hp_w);
    });
}
int64_t ifbh_model_tree_string(void *model, int32_t tree, char *buf, int64_t cap) {
    std::string s;
    if (guarded([&] { s = ((ForestModelBase *)model)->treeToString(tree); })) return -1;
    return copyOut(s, buf, cap);
}

}  // extern "C"

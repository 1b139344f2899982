// Avro object-container I/O for the saved-model node tables.
//
// On-disk contract (kept so that models written here load in the stock reference and vice versa):
//   data/*.avro rows {treeID:int, nodeData:{id,leftChild,rightChild,splitAttribute:int, splitValue:double,
//   numInstances:long}} (IF/IsolationForestModelReadWrite.scala:60-67,147) or
//   {treeID, extendedNodeData:{id,leftChild,rightChild:int, indices:int[], weights:float[], offset:double,
//   numInstances:long}} (IF/extended/ExtendedIsolationForestModelReadWrite.scala:59-67,147-150); pre-order ids,
//   -1 / 0.0 sentinels at leaves, numInstances -1 at internal nodes.  Codecs read: null, deflate, snappy (the
//   reference's fixtures use snappy and deflate); written: deflate (default) or null.
// Implements the public Avro 1.x container spec; the decoder is schema-driven (record/union/array/primitive).
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <random>
#include <stdexcept>

#include "host_internal.hpp"

namespace ifb200 {
namespace avro {

namespace {

struct Reader {
    const uint8_t *p, *end;
    uint8_t byte() {
        if (p >= end) throw std::runtime_error("avro: truncated data");
        return *p++;
    }
    int64_t zz() {
        uint64_t acc = 0;
        int shift = 0;
        while (true) {
            uint8_t c = byte();
            acc |= (uint64_t)(c & 0x7F) << shift;
            if (!(c & 0x80)) break;
            shift += 7;
            if (shift > 63) throw std::runtime_error("avro: varint too long");
        }
        return (int64_t)(acc >> 1) ^ -(int64_t)(acc & 1);
    }
    const uint8_t *take(size_t n) {
        if ((size_t)(end - p) < n) throw std::runtime_error("avro: truncated data");
        const uint8_t *q = p;
        p += n;
        return q;
    }
    std::string str() {
        int64_t n = zz();
        if (n < 0) throw std::runtime_error("avro: negative length");
        const uint8_t *q = take((size_t)n);
        return std::string((const char *)q, (size_t)n);
    }
};

// raw snappy block (public format): varint length, then literal / copy elements
std::vector<uint8_t> snappy_decompress(const uint8_t *src, size_t n) {
    size_t i = 0;
    uint64_t len = 0;
    int shift = 0;
    while (true) {
        if (i >= n) throw std::runtime_error("snappy: truncated header");
        uint8_t c = src[i++];
        len |= (uint64_t)(c & 0x7F) << shift;
        if (!(c & 0x80)) break;
        shift += 7;
    }
    std::vector<uint8_t> out;
    out.reserve(len);
    while (i < n) {
        uint8_t tag = src[i++];
        int t = tag & 3;
        if (t == 0) {
            size_t ln = tag >> 2;
            if (ln >= 60) {
                int nb = (int)ln - 59;
                if (i + nb > n) throw std::runtime_error("snappy: truncated literal length");
                ln = 0;
                for (int b = 0; b < nb; b++) ln |= (size_t)src[i + b] << (8 * b);
                i += nb;
            }
            ln += 1;
            if (i + ln > n) throw std::runtime_error("snappy: truncated literal");
            out.insert(out.end(), src + i, src + i + ln);
            i += ln;
            continue;
        }
        size_t ln, off;
        if (t == 1) {
            if (i + 1 > n) throw std::runtime_error("snappy: truncated copy");
            ln = ((tag >> 2) & 7) + 4;
            off = ((size_t)(tag >> 5) << 8) | src[i];
            i += 1;
        } else if (t == 2) {
            if (i + 2 > n) throw std::runtime_error("snappy: truncated copy");
            ln = (tag >> 2) + 1;
            off = src[i] | ((size_t)src[i + 1] << 8);
            i += 2;
        } else {
            if (i + 4 > n) throw std::runtime_error("snappy: truncated copy");
            ln = (tag >> 2) + 1;
            off = src[i] | ((size_t)src[i + 1] << 8) | ((size_t)src[i + 2] << 16) | ((size_t)src[i + 3] << 24);
            i += 4;
        }
        if (off == 0 || off > out.size()) throw std::runtime_error("snappy: bad copy offset");
        size_t start = out.size() - off;
        for (size_t k = 0; k < ln; k++) out.push_back(out[start + k]);
    }
    if (out.size() != len) throw std::runtime_error("snappy: length mismatch");
    return out;
}

std::vector<uint8_t> inflate_raw(const uint8_t *src, size_t n) {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("zlib: inflateInit2 failed");
    std::vector<uint8_t> out(std::max<size_t>(n * 4, 1 << 16));
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = (uInt)n;
    size_t have = 0;
    int rc;
    do {
        if (have == out.size()) out.resize(out.size() * 2);
        zs.next_out = out.data() + have;
        zs.avail_out = (uInt)(out.size() - have);
        rc = inflate(&zs, Z_NO_FLUSH);
        have = out.size() - zs.avail_out;
        if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) {
            inflateEnd(&zs);
            throw std::runtime_error("zlib: inflate failed");
        }
        if (rc == Z_BUF_ERROR && zs.avail_in == 0) break;
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
    out.resize(have);
    return out;
}

std::vector<uint8_t> deflate_raw(const std::vector<uint8_t> &in) {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK)
        throw std::runtime_error("zlib: deflateInit2 failed");
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)in.size()));
    zs.next_in = const_cast<Bytef *>(in.data());
    zs.avail_in = (uInt)in.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) {
        deflateEnd(&zs);
        throw std::runtime_error("zlib: deflate failed");
    }
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

// ---- schema-driven decode into a tiny dynamic value ----------------------------------------------
struct Dyn {
    int64_t i = 0;
    double d = 0;
    bool null = true;
    std::vector<Dyn> items;                          // array
    std::vector<std::pair<std::string, Dyn>> fields;  // record
    const Dyn *field(const char *name) const {
        for (auto &f : fields)
            if (f.first == name) return &f.second;
        return nullptr;
    }
};

Dyn decode(const json::Value &schema, Reader &r) {
    Dyn out;
    if (schema.kind == json::Value::Array) {  // union
        int64_t b = r.zz();
        if (b < 0 || (size_t)b >= schema.arr.size()) throw std::runtime_error("avro: bad union branch");
        return decode(schema.arr[(size_t)b], r);
    }
    std::string t;
    const json::Value *sch = &schema;
    if (schema.kind == json::Value::Object) {
        const json::Value *ty = schema.get("type");
        if (!ty) throw std::runtime_error("avro: schema without type");
        if (ty->kind != json::Value::String) return decode(*ty, r);
        t = ty->s;
    } else if (schema.kind == json::Value::String) {
        t = schema.s;
    } else {
        throw std::runtime_error("avro: unsupported schema node");
    }
    if (t == "record") {
        const json::Value *fs = sch->get("fields");
        if (!fs) throw std::runtime_error("avro: record without fields");
        out.null = false;
        for (auto &f : fs->arr) {
            const json::Value *fname = f.get("name"), *ftype = f.get("type");
            if (!fname || fname->kind != json::Value::String || !ftype) throw std::runtime_error("avro: malformed record field");
            out.fields.emplace_back(fname->s, decode(*ftype, r));
        }
    } else if (t == "array") {
        out.null = false;
        const json::Value *it = sch->get("items");
        if (!it) throw std::runtime_error("avro: array without items");
        while (true) {
            int64_t cnt = r.zz();
            if (cnt == 0) break;
            if (cnt < 0) {
                cnt = -cnt;
                r.zz();
            }
            for (int64_t k = 0; k < cnt; k++) out.items.push_back(decode(*it, r));
        }
    } else if (t == "int" || t == "long") {
        out.null = false;
        out.i = r.zz();
    } else if (t == "double") {
        out.null = false;
        std::memcpy(&out.d, r.take(8), 8);
    } else if (t == "float") {
        out.null = false;
        float f;
        std::memcpy(&f, r.take(4), 4);
        out.d = f;
    } else if (t == "null") {
    } else if (t == "boolean") {
        out.null = false;
        out.i = r.byte() != 0;
    } else if (t == "string" || t == "bytes") {
        out.null = false;
        r.str();
    } else {
        throw std::runtime_error("avro: unsupported type " + t);
    }
    return out;
}

struct Writer {
    std::vector<uint8_t> b;
    void zz(int64_t v) {
        uint64_t u = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
        while (u >= 0x80) {
            b.push_back((uint8_t)(u | 0x80));
            u >>= 7;
        }
        b.push_back((uint8_t)u);
    }
    void raw(const void *p, size_t n) { b.insert(b.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
    void str(const std::string &s) {
        zz((int64_t)s.size());
        raw(s.data(), s.size());
    }
};

const char *kStdSchema =
    "{\"type\":\"record\",\"name\":\"topLevelRecord\",\"fields\":[{\"name\":\"treeID\",\"type\":\"int\"},"
    "{\"name\":\"nodeData\",\"type\":[{\"type\":\"record\",\"name\":\"nodeData\",\"namespace\":\"topLevelRecord\","
    "\"fields\":[{\"name\":\"id\",\"type\":\"int\"},{\"name\":\"leftChild\",\"type\":\"int\"},"
    "{\"name\":\"rightChild\",\"type\":\"int\"},{\"name\":\"splitAttribute\",\"type\":\"int\"},"
    "{\"name\":\"splitValue\",\"type\":\"double\"},{\"name\":\"numInstances\",\"type\":\"long\"}]},\"null\"]}]}";
const char *kExtSchema =
    "{\"type\":\"record\",\"name\":\"topLevelRecord\",\"fields\":[{\"name\":\"treeID\",\"type\":\"int\"},"
    "{\"name\":\"extendedNodeData\",\"type\":[{\"type\":\"record\",\"name\":\"extendedNodeData\","
    "\"namespace\":\"topLevelRecord\",\"fields\":[{\"name\":\"id\",\"type\":\"int\"},"
    "{\"name\":\"leftChild\",\"type\":\"int\"},{\"name\":\"rightChild\",\"type\":\"int\"},"
    "{\"name\":\"indices\",\"type\":[{\"type\":\"array\",\"items\":\"int\"},\"null\"]},"
    "{\"name\":\"weights\",\"type\":[{\"type\":\"array\",\"items\":\"float\"},\"null\"]},"
    "{\"name\":\"offset\",\"type\":\"double\"},{\"name\":\"numInstances\",\"type\":\"long\"}]},\"null\"]}]}";

}  // namespace

// Read every *.avro file of `dir` and assemble the forest tables (rows may arrive in any order).
ForestTables read_tables(const std::string &dir, bool extended) {
    namespace fs = std::filesystem;
    struct Node {
        int32_t id, left, right, feature;
        double value;
        int64_t ninst;
        std::vector<int32_t> idx;
        std::vector<float> w;
    };
    std::map<int32_t, std::vector<Node>> trees;
    std::vector<fs::path> files;
    if (!fs::is_directory(dir)) throw std::runtime_error("model data directory not found: " + dir);
    for (auto &e : fs::directory_iterator(dir))
        if (e.path().extension() == ".avro") files.push_back(e.path());
    std::sort(files.begin(), files.end());
    const char *key = extended ? "extendedNodeData" : "nodeData";
    for (auto &path : files) {
        std::ifstream in(path, std::ios::binary);
        std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        Reader r{buf.data(), buf.data() + buf.size()};
        if (buf.size() < 4 || std::memcmp(r.take(4), "Obj\x01", 4) != 0)
            throw std::runtime_error("not an Avro container: " + path.string());
        std::map<std::string, std::string> meta;
        while (true) {
            int64_t cnt = r.zz();
            if (cnt == 0) break;
            if (cnt < 0) {
                cnt = -cnt;
                r.zz();
            }
            for (int64_t k = 0; k < cnt; k++) {
                std::string kname = r.str();
                meta[kname] = r.str();
            }
        }
        uint8_t sync[16];
        std::memcpy(sync, r.take(16), 16);
        if (!meta.count("avro.schema")) throw std::runtime_error("avro: missing schema in " + path.string());
        json::Value schema = json::parse(meta["avro.schema"]);
        std::string codec = meta.count("avro.codec") ? meta["avro.codec"] : "null";
        while (r.p < r.end) {
            int64_t cnt = r.zz();
            int64_t size = r.zz();
            if (size < 0) throw std::runtime_error("avro: negative block size");
            const uint8_t *payload = r.take((size_t)size);
            if (std::memcmp(r.take(16), sync, 16) != 0) throw std::runtime_error("avro: sync marker mismatch");
            std::vector<uint8_t> plain;
            if (codec == "deflate") {
                plain = inflate_raw(payload, (size_t)size);
            } else if (codec == "snappy") {
                if (size < 4) throw std::runtime_error("avro: snappy block too short");
                plain = snappy_decompress(payload, (size_t)size - 4);
                uint32_t want = ((uint32_t)payload[size - 4] << 24) | ((uint32_t)payload[size - 3] << 16) |
                                ((uint32_t)payload[size - 2] << 8) | payload[size - 1];
                if ((uint32_t)crc32(0L, plain.data(), (uInt)plain.size()) != want)
                    throw std::runtime_error("avro: snappy CRC mismatch");
            } else if (codec == "null" || codec == "uncompressed") {
                plain.assign(payload, payload + size);
            } else {
                throw std::runtime_error("avro: unsupported codec " + codec);
            }
            Reader br{plain.data(), plain.data() + plain.size()};
            for (int64_t k = 0; k < cnt; k++) {
                Dyn rec = decode(schema, br);
                const Dyn *tid = rec.field("treeID");
                const Dyn *nd = rec.field(key);
                if (!tid || !nd || nd->null)
                    throw std::runtime_error(std::string("avro: row without ") + key + " (wrong model class?)");
                Node n;
                n.id = (int32_t)nd->field("id")->i;
                n.left = (int32_t)nd->field("leftChild")->i;
                n.right = (int32_t)nd->field("rightChild")->i;
                n.ninst = nd->field("numInstances")->i;
                if (extended) {
                    n.feature = -1;
                    n.value = nd->field("offset")->d;
                    const Dyn *ix = nd->field("indices"), *ws = nd->field("weights");
                    if (ix)
                        for (auto &v : ix->items) n.idx.push_back((int32_t)v.i);
                    if (ws)
                        for (auto &v : ws->items) n.w.push_back((float)v.d);
                } else {
                    n.feature = (int32_t)nd->field("splitAttribute")->i;
                    n.value = nd->field("splitValue")->d;
                }
                trees[(int32_t)tid->i].push_back(std::move(n));
            }
        }
    }
    ForestTables t;
    t.extended = extended;
    t.node_off.push_back(0);
    if (extended) t.hp_off.push_back(0);
    int32_t expect = 0;
    for (auto &kv : trees) {
        // treeIDs are the ensemble positions: sortByKey in the reference's loader
        if (kv.first != expect++) throw IllegalArgumentException("Isolation forest load failed: tree IDs are not 0..T-1.");
        auto &nodes = kv.second;
        std::sort(nodes.begin(), nodes.end(), [](const Node &a, const Node &b) { return a.id < b.id; });
        for (size_t i = 0; i < nodes.size(); i++)
            if (nodes[i].id != (int32_t)i)
                throw IllegalArgumentException(  // IF/IsolationForestModelReadWrite.scala:183-188
                    "Isolation tree load failed. Expected the " + std::to_string(nodes.size()) +
                    " node IDs to be monotonically increasing from 0 to " + std::to_string(nodes.size() - 1) + ".");
        for (auto &n : nodes) {
            t.left.push_back(n.left);
            t.right.push_back(n.right);
            t.num_instances.push_back(n.ninst);
            if (extended) {
                t.offset.push_back(n.value);
                t.hp_idx.insert(t.hp_idx.end(), n.idx.begin(), n.idx.end());
                t.hp_w.insert(t.hp_w.end(), n.w.begin(), n.w.end());
                if (n.idx.size() != n.w.size())
                    throw IllegalArgumentException("indices and weights must have the same length.");
                t.hp_off.push_back((int64_t)t.hp_idx.size());
            } else {
                t.feature.push_back(n.feature);
                t.threshold.push_back(n.value);
            }
        }
        t.node_off.push_back((int32_t)t.left.size());
    }
    return t;
}

void write_tables(const std::string &dir, const ForestTables &t, const std::string &codec) {
    namespace fs = std::filesystem;
    fs::create_directories(dir);
    std::random_device rd;
    std::mt19937_64 gen(((uint64_t)rd() << 32) ^ rd());
    char name[96];
    snprintf(name, sizeof name, "part-00000-%08x-%04x-%04x-%04x-%012llx-c000.avro", (unsigned)gen(),
             (unsigned)gen() & 0xffff, (unsigned)gen() & 0xffff, (unsigned)gen() & 0xffff,
             (unsigned long long)(gen() & 0xffffffffffffULL));
    Writer hdr;
    hdr.raw("Obj\x01", 4);
    hdr.zz(2);
    hdr.str("avro.schema");
    hdr.str(t.extended ? kExtSchema : kStdSchema);
    hdr.str("avro.codec");
    hdr.str(codec == "uncompressed" ? "null" : codec);
    hdr.zz(0);
    uint8_t sync[16];
    for (int i = 0; i < 16; i++) sync[i] = (uint8_t)gen();
    hdr.raw(sync, 16);
    std::ofstream out(fs::path(dir) / name, std::ios::binary);
    out.write((const char *)hdr.b.data(), (std::streamsize)hdr.b.size());

    const int T = t.num_trees();
    Writer blk;
    int64_t count = 0;
    auto flush = [&]() {
        if (count == 0) return;
        std::vector<uint8_t> payload = (codec == "deflate") ? deflate_raw(blk.b) : blk.b;
        Writer h;
        h.zz(count);
        h.zz((int64_t)payload.size());
        out.write((const char *)h.b.data(), (std::streamsize)h.b.size());
        out.write((const char *)payload.data(), (std::streamsize)payload.size());
        out.write((const char *)sync, 16);
        blk.b.clear();
        count = 0;
    };
    for (int tr = 0; tr < T; tr++) {
        const int32_t base = t.node_off[tr], n = t.node_off[tr + 1] - base;
        for (int32_t i = 0; i < n; i++) {
            const int64_t g = (int64_t)base + i;
            blk.zz(tr);  // treeID
            blk.zz(0);   // union branch 0 = record
            blk.zz(i);
            blk.zz(t.left[g]);
            blk.zz(t.right[g]);
            if (t.extended) {
                const int64_t b = t.hp_off[g], e = t.hp_off[g + 1];
                blk.zz(0);  // indices: union branch 0 = array
                if (e > b) {
                    blk.zz(e - b);
                    for (int64_t q = b; q < e; q++) blk.zz(t.hp_idx[q]);
                }
                blk.zz(0);
                blk.zz(0);  // weights: union branch 0 = array
                if (e > b) {
                    blk.zz(e - b);
                    for (int64_t q = b; q < e; q++) blk.raw(&t.hp_w[q], 4);
                }
                blk.zz(0);
                blk.raw(&t.offset[g], 8);
            } else {
                blk.zz(t.feature[g]);
                blk.raw(&t.threshold[g], 8);
            }
            blk.zz(t.num_instances[g]);
            if (++count >= 4096) flush();
        }
    }
    flush();
    out.close();
    std::ofstream(fs::path(dir) / "_SUCCESS").close();
}

}  // namespace avro
}  // namespace ifb200

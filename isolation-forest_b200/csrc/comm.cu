// Tree-sharded scoring with the collective INSIDE the library (ifb_comm_init / ifb_score_sharded), so that a
// JVM executor -- which cannot call torch.distributed -- can run BASELINE.json's multi-GPU layout: numEstimators split over
// the GPUs (the reference's tree-parallel fit, IF/core/SharedTrainLogic.scala:140-149,276-317), every rank scores all rows
// against its slice of the ensemble, ONE NCCL all-reduce / reduce-scatter of the per-row f32 path-length sums over
// NVLink, then the 2^(-E/c) epilogue with the full ensemble size (IF/IsolationForestModel.scala:131-139).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, which resolves to the copy a host framework already loaded, if any):
// libifb200.so carries no link-time dependency on it, and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "ifb_internal.h"

namespace ifb {
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *n) { return dlsym(api.handle, n); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.ReduceScatter = reinterpret_cast<decltype(api.ReduceScatter)>(sym("ncclReduceScatter"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.ReduceScatter && api.GetErrorString;
    });
    return api;
}

int need_nccl() {
    if (!nccl().ok) {
        set_error("NCCL is not available: dlopen(libnccl.so.2) failed (%s)", dlerror() ? dlerror() : "symbols missing");
        return IFB_ENCCL;
    }
    return IFB_OK;
}

#define IFB_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            ::ifb::set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl().GetErrorString(_r)); \
            return IFB_ENCCL;                                                                       \
        }                                                                                           \
    } while (0)

}  // namespace
}  // namespace ifb

struct ifb_comm {
    int32_t device = 0, world = 1, rank = 0;
    ncclComm_t comm = nullptr;
};

using namespace ifb;

extern "C" {

int ifb_comm_unique_id(void *id128) {
    IFB_REQUIRE(id128, "null argument");
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    IFB_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
    return IFB_OK;
}

int ifb_comm_init(int32_t device, int32_t world, int32_t rank, const void *id128, ifb_comm **out) {
    IFB_REQUIRE(out && id128, "null argument");
    *out = nullptr;
    IFB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad world/rank %d/%d", world, rank);
    int rc = need_nccl();
    if (rc) return rc;
    DeviceGuard dg(device);
    IFB_REQUIRE(dg.ok, "cannot select device %d", device);
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    auto *c = new ifb_comm();
    c->device = device;
    c->world = world;
    c->rank = rank;
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
        delete c;
        return IFB_ENCCL;
    }
    *out = c;
    return IFB_OK;
}

int ifb_comm_destroy(ifb_comm *comm) {
    if (!comm) return IFB_OK;
    if (comm->comm && nccl().ok) {
        DeviceGuard dg(comm->device);
        nccl().CommDestroy(comm->comm);
    }
    delete comm;
    return IFB_OK;
}

int ifb_score_sharded(const ifb_forest *forest, ifb_comm *comm, const float *X, int64_t n_rows, int32_t d, int64_t ld,
                      int32_t layout, int32_t total_num_trees, int32_t mode, double *scores, int64_t *slice_begin,
                      int64_t *slice_end, void *stream_) {
    IFB_REQUIRE(forest && comm && comm->comm, "null forest / communicator");
    IFB_REQUIRE(mode == IFB_SHARD_ALLREDUCE || mode == IFB_SHARD_REDUCE_SCATTER, "unknown mode %d", mode);
    IFB_REQUIRE(total_num_trees >= forest->num_trees, "total_num_trees %d smaller than this shard (%d trees)", total_num_trees,
                forest->num_trees);
    IFB_REQUIRE(forest->device == comm->device, "forest lives on device %d, communicator on %d", forest->device, comm->device);
    IFB_REQUIRE(n_rows >= 0 && (n_rows == 0 || scores), "null scores");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t per = (n_rows + comm->world - 1) / comm->world;
    const int64_t r0 = std::min<int64_t>(n_rows, (int64_t)comm->rank * per), r1 = std::min<int64_t>(n_rows, r0 + per);
    if (slice_begin) *slice_begin = mode == IFB_SHARD_ALLREDUCE ? 0 : r0;
    if (slice_end) *slice_end = mode == IFB_SHARD_ALLREDUCE ? n_rows : r1;
    if (n_rows == 0) return IFB_OK;
    DeviceGuard dg(forest->device);
    struct Scratch {
        cudaStream_t s;
        float *p = nullptr;
        ~Scratch() {
            if (p) cudaFreeAsync(p, s);
        }
    } psum{stream}, part{stream};
    const size_t padded = (size_t)per * comm->world;
    IFB_CUDA(cudaMallocAsync((void **)&psum.p, padded * 4, stream));
    IFB_CUDA(cudaMemsetAsync(psum.p, 0, padded * 4, stream));
    int rc = ifb_score_partial_device(forest, X, n_rows, d, ld, layout, psum.p, nullptr, stream);
    if (rc) return rc;
    if (mode == IFB_SHARD_ALLREDUCE) {
        IFB_NCCL(nccl().AllReduce(psum.p, psum.p, padded, ncclFloat32, ncclSum, comm->comm, stream));
        return ifb_finalize_scores_device(forest->device, psum.p, n_rows, total_num_trees, forest->num_samples, scores, stream);
    }
    IFB_CUDA(cudaMallocAsync((void **)&part.p, (size_t)per * 4, stream));
    IFB_NCCL(nccl().ReduceScatter(psum.p, part.p, (size_t)per, ncclFloat32, ncclSum, comm->comm, stream));
    if (r1 > r0)
        return ifb_finalize_scores_device(forest->device, part.p, r1 - r0, total_num_trees, forest->num_samples, scores, stream);
    return IFB_OK;
}

}  // extern "C"

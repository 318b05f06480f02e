// zuko_b200 — the ONE collective of the path behind the C ABI: all-reduce(sum) of the
// per-device {sum log p, count} doubles for the scalar mean NLL, over NCCL / NVLink.  The Python mirror
// (zuko_b200/dist.py) issues the same collective through torch.distributed with one process per GPU; this
// is the form a single-process host (a C++ serving binary, one thread or process driving several GPUs)
// binds.  NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already loaded by the process if
// there is one), so the library has no link-time dependency on it.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;  // ncclSuccess = 0
constexpr int kNcclFloat64 = 8, kNcclSum = 0;  // ncclDataType_t ncclFloat64 / ncclRedOp_t ncclSum (nccl.h)

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char* n) { return dlsym(api.handle, n); };
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd;
    });
    return api;
}

zk_status nccl_fail(const char* what, ncclResult_t r) {
    const char* msg = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
    return zk::fail(ZK_ECUDA, "%s: NCCL error %d (%s)", what, (int)r, msg);
}

}  // namespace

struct zk_comm {
    std::vector<ncclComm_t> comms;  // one per device 0 .. ndev-1
};

extern "C" {

zk_status zk_comm_init_all(int ndev, zk_comm** out) {
    ZK_REQUIRE(out && ndev >= 1, "comm_init_all: bad arguments");
    *out = nullptr;
    int have = 0;
    ZK_CUDA(cudaGetDeviceCount(&have));
    ZK_REQUIRE(ndev <= have, "comm_init_all: %d devices requested, %d visible", ndev, have);
    if (!nccl().ok) return zk::fail(ZK_EUNSUPPORTED, "comm_init_all: libnccl.so.2 could not be loaded");
    zk_comm* c = new zk_comm();
    c->comms.assign(ndev, nullptr);
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; ++i) devs[i] = i;
    ncclResult_t r = nccl().CommInitAll(c->comms.data(), ndev, devs.data());
    if (r != 0) {
        delete c;
        return nccl_fail("ncclCommInitAll", r);
    }
    *out = c;
    return ZK_OK;
}

zk_status zk_comm_destroy(zk_comm* c) {
    if (!c) return ZK_OK;
    for (ncclComm_t q : c->comms)
        if (q) nccl().CommDestroy(q);
    delete c;
    return ZK_OK;
}

int zk_comm_size(const zk_comm* c) { return c ? (int)c->comms.size() : 0; }

zk_status zk_comm_group_begin(zk_comm* c) {
    ZK_REQUIRE(c, "comm_group_begin: null communicator");
    ncclResult_t r = nccl().GroupStart();
    return r == 0 ? ZK_OK : nccl_fail("ncclGroupStart", r);
}

zk_status zk_comm_group_end(zk_comm* c) {
    ZK_REQUIRE(c, "comm_group_end: null communicator");
    ncclResult_t r = nccl().GroupEnd();
    return r == 0 ? ZK_OK : nccl_fail("ncclGroupEnd", r);
}

zk_status zk_allreduce_sum(zk_comm* c, int dev, double* buf, int n, zk_stream stream) {
    ZK_REQUIRE(c && buf && n >= 1, "allreduce_sum: bad arguments");
    ZK_REQUIRE(dev >= 0 && dev < (int)c->comms.size(), "allreduce_sum: device %d outside the communicator", dev);
    int prev = 0;
    ZK_CUDA(cudaGetDevice(&prev));
    ZK_CUDA(cudaSetDevice(dev));
    ncclResult_t r = nccl().AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c->comms[dev], (cudaStream_t)stream);
    cudaSetDevice(prev);
    return r == 0 ? ZK_OK : nccl_fail("ncclAllReduce", r);
}

}  // extern "C"

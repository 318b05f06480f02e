// zuko_b200 — tcgen05 conditioner path (placeholder until the kernel lands in this round).
#include "mlp_tcgen05.cuh"

namespace zk {

zk_status tc_pack(zk_mlp* m, int requested_mode) {
    if (requested_mode == ZK_GEMM_AUTO || requested_mode == ZK_GEMM_FP32) {
        m->gemm_mode = ZK_GEMM_FP32;
        return ZK_OK;
    }
    return fail(ZK_EUNSUPPORTED, "tcgen05 conditioner path not built in this library");
}
void tc_destroy(zk_mlp*) {}
size_t tc_workspace_bytes(const zk_mlp*, int64_t) { return 0; }
zk_status tc_forward(const zk_mlp*, const float*, int64_t, int, const float*, int64_t, int, int64_t,
                     float*, int64_t, void*, size_t, cudaStream_t) {
    return fail(ZK_EUNSUPPORTED, "tcgen05 conditioner path not built in this library");
}

}  // namespace zk

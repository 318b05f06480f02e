// zuko_b200 — tcgen05 (5th-gen tensor core) conditioner path: interface.
#pragma once

#include "mlp.cuh"

namespace zk {

// Resolves m->gemm_mode from the requested mode and packs bf16 hi/lo weights when the
// tensor-core path is selected.  Runs on stream 0.
zk_status tc_pack(zk_mlp* m, int requested_mode);
void tc_destroy(zk_mlp* m);
size_t tc_workspace_bytes(const zk_mlp* m, int64_t B);
zk_status tc_forward(const zk_mlp* m, const float* x, int64_t ldx, int dx, const float* c,
                     int64_t ldc, int dc, int64_t B, float* out, int64_t ldo, void* ws,
                     size_t ws_bytes, cudaStream_t stream);

}  // namespace zk

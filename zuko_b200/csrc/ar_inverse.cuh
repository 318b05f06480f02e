// zuko_b200 — dimension-sequential inverse of a masked autoregressive layer: interface.
#pragma once

#include "mlp.cuh"

namespace zk {

struct ArInvPack;  // step-ordered weight stream of one layer (ar_inverse.cu)

// Builds the pack when the layer has a pure order-class structure (an `order` vector, RQS with 8 /
// 16 bins or affine); leaves *out null (and returns ZK_OK) when the layer must use the generic
// sweep-based inverse instead.  `mask_dev` are the DEVICE mask pointers of zk_mlp_desc (may be null).
zk_status ar_inverse_pack(const zk_mlp* m, const uint8_t* const* mask_dev, const int64_t* order, int D, int C,
                          int univariate, int bins, int passes, ArInvPack** out);
void ar_inverse_free(ArInvPack* pk);
bool ar_inverse_threads(const ArInvPack* pk, int* threads, size_t* smem);
zk_status launch_ar_inverse(const ArInvPack* pk, const float* y, int64_t ldy, const float* c, int64_t ldc, int64_t B,
                            float* x, int64_t ldx, float bound, float slope, bool fast, bool circular, cudaStream_t stream);

}  // namespace zk

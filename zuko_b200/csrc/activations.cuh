// zuko_b200 — activation functions of the conditioner (zuko/nn.py:160-192, 258-318: any
// `activation()` module between the linear layers; the engine implements torch's defaults of the
// element-wise ones below).  Internal codes: 0 = none, 1 = ReLU, >= 2 = ZK_ACT_* of the C ABI.
#pragma once

#include <math.h>

#include "../../include/zuko_b200.h"

#if defined(__CUDACC__)
#define ZK_ACT_HD __host__ __device__ __forceinline__
#else
#define ZK_ACT_HD inline
#endif

namespace zk {

ZK_ACT_HD float act_apply(float v, int act) {
    switch (act) {
        case 1: return fmaxf(v, 0.f);                                  // torch.nn.ReLU
        case ZK_ACT_ELU: return v > 0.f ? v : expm1f(v);               // ELU(alpha=1)
        case ZK_ACT_TANH: return tanhf(v);                             // Tanh
        case ZK_ACT_SILU: return v / (1.f + expf(-v));                 // SiLU
        case ZK_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // GELU(approximate='none')
        case ZK_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;        // LeakyReLU(0.01)
        case ZK_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));   // Softplus(beta=1, threshold=20)
        case ZK_ACT_SIGMOID: return 1.f / (1.f + expf(-v));            // Sigmoid
        default: return v;
    }
}

// d act(v) / dv at the PRE-activation v
ZK_ACT_HD float act_deriv(float v, int act) {
    switch (act) {
        case 1: return v > 0.f ? 1.f : 0.f;
        case ZK_ACT_ELU: return v > 0.f ? 1.f : expf(v);
        case ZK_ACT_TANH: { const float t = tanhf(v); return 1.f - t * t; }
        case ZK_ACT_SILU: { const float s = 1.f / (1.f + expf(-v)); return s * (1.f + v * (1.f - s)); }
        case ZK_ACT_GELU:
            return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * expf(-0.5f * v * v) * 0.39894228040143268f;
        case ZK_ACT_LEAKY_RELU: return v > 0.f ? 1.f : 0.01f;
        case ZK_ACT_SOFTPLUS: return v > 20.f ? 1.f : 1.f / (1.f + expf(-v));
        case ZK_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-v)); return s * (1.f - s); }
        default: return 1.f;
    }
}

}  // namespace zk

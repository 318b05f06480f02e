// CPU harness around zuko_b200/csrc/bijector_grad.cuh (the per-pair reverse-mode math that
// uni_bwd_kernel runs on the GPU), compiled by g++ from tests/test_bijector_grad_host.py so that
// the derivation is checked against the gradient oracle / reference autograd without a GPU.
#include <stdint.h>

#include "../../zuko_b200/csrc/bijector_grad.cuh"

extern "C" {

// x, gy, gl, gx: (n) pairs; phi, gphi: (n, 3K-1)
void rqs_backward_pairs(const float* x, const float* phi, const float* gy, const float* gl,
                        int64_t n, int K, float bound, float slope, float* gx, float* gphi) {
    const float absL = fabsf(logf(slope));
    const int P = 3 * K - 1;
    for (int64_t i = 0; i < n; ++i)
        zk::bijgrad::rqs_backward_pair<0>(phi + i * P, K, x[i], gy[i], gl[i], bound, 2.f / absL,
                                          1.f / absL, gx[i], gphi + i * P);
}

// same, through the K = 8 compile-time instance and IN PLACE (gphi aliases phi)
void rqs_backward_pairs_k8_inplace(const float* x, float* phi, const float* gy, const float* gl,
                                   int64_t n, float bound, float slope, float* gx) {
    const float absL = fabsf(logf(slope));
    for (int64_t i = 0; i < n; ++i)
        zk::bijgrad::rqs_backward_pair<8>(phi + i * 23, 8, x[i], gy[i], gl[i], bound, 2.f / absL,
                                          1.f / absL, gx[i], phi + i * 23);
}

void affine_backward_pairs(const float* x, const float* phi, const float* gy, const float* gl,
                           int64_t n, float slope, float* gx, float* gphi) {
    const float absL = fabsf(logf(slope));
    for (int64_t i = 0; i < n; ++i)
        zk::bijgrad::affine_backward_pair(phi + i * 2, x[i], gy[i], gl[i], 1.f / absL, gx[i],
                                          gphi + i * 2);
}
}

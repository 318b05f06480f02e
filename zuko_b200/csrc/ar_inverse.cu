// zuko_b200 — dimension-sequential inverse of a masked autoregressive layer.
//
// The reference inverts  y = f(x | x_<, c)  with `passes` fixed-point sweeps, each running the
// FULL conditioner and the full bijector over all D dims (zuko/transforms.py:994-1000): for
// BASELINE config 4 that is 64 sweeps x 8 layers of a 64-64-64-3008 MLP, although sweep p only
// finalises the dims of order class p.  The masks make the dependency structure explicit
// (zuko/nn.py:270-293, flows/autoregressive.py:121-124): a hidden unit / output row depends on the
// x dims of a bounded order class only, so everything can be evaluated exactly once, in order:
//
//   step p = 0 .. passes-1:
//     every hidden unit whose inputs became final at step p is computed (pull style: a full dot
//     product over its layer input — the not-yet-final inputs carry masked-out = 0 weights);
//     the P parameter rows of every dim of class p are computed from the last hidden layer;
//     the inverse bijector (transforms.py:534-548 / 443-444) yields x_d, which is stored and
//     becomes an input of the later steps.
//
// This visits every weight once (+ tile padding) instead of `passes` times and reproduces the
// reference's fixed point up to summation order.
//
// Mapping: thread = sample (128 per CTA), state vectors (x | c, hidden activations, y) live in
// shared memory as [k][thread]; the weights of a step are a pre-packed stream (8-row tiles,
// k-major) that the CTA stages into shared memory and every thread reads by broadcast.  This is
// FMA-pipe work by design: the per-step GEMMs are (samples x <=64 rows), far below a tcgen05 tile.

#include <string.h>

#include <algorithm>
#include <vector>

#include "ar_inverse.cuh"
#include "bijector_math.cuh"

namespace zk {

struct ArInvPack {
    int D = 0, C = 0, P = 0, uni = 0, bins = 0, passes = 0;
    int n_linear = 0;
    std::vector<int> dims;         // n_linear + 1
    std::vector<int> sec_off;      // state section offsets (floats): IN, H1 .. H_{L-1}, end
    int state_floats = 0;
    int max_step_words = 0;
    float* stream = nullptr;       // device: all step blocks
    int* step_off = nullptr;       // device (passes + 1): word offsets of the step blocks
    std::vector<int> h_step_off;
};

namespace {

using namespace bij;

constexpr int TILE = 8;

struct InvParams {
    const float* stream;
    const int* step_off;
    int passes, n_linear, D, C, P;
    int dims[8];       // layer widths
    int sec_off[9];    // state sections: IN, H1 .. H_{L-1}, end
    const float* y; int64_t ldy;
    const float* c; int64_t ldc;
    float* x; int64_t ldx;
    int64_t B;
    float bound, aw, ad;
    int circ;  // circular RQS (NCSF): CircularShiftTransform(bound) applied to the solved dimension
};

// acc[j] += sum_k state[k][tid] * w[k][j]  over one 8-row tile (weights broadcast from smem)
__device__ __forceinline__ void tile_dot(const float* __restrict__ w, const float* __restrict__ st, int K, int T,
                                         float (&acc)[TILE]) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float s = st[k * T];
        const float4 w0 = *reinterpret_cast<const float4*>(w + k * TILE);
        const float4 w1 = *reinterpret_cast<const float4*>(w + k * TILE + 4);
        acc[0] = fmaf(s, w0.x, acc[0]); acc[1] = fmaf(s, w0.y, acc[1]);
        acc[2] = fmaf(s, w0.z, acc[2]); acc[3] = fmaf(s, w0.w, acc[3]);
        acc[4] = fmaf(s, w1.x, acc[4]); acc[5] = fmaf(s, w1.y, acc[5]);
        acc[6] = fmaf(s, w1.z, acc[6]); acc[7] = fmaf(s, w1.w, acc[7]);
    }
}

template <int UNI, int KT, bool FAST>
__global__ void ar_inverse_kernel(const InvParams p) {
    constexpr int P = (UNI == ZK_UNI_RQS) ? 3 * KT - 1 : 2;
    constexpr int PT = (P + TILE - 1) / TILE;  // tiles per dim
    extern __shared__ __align__(16) float smem_f[];
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    float* state = smem_f;                                   // [state_floats][T]
    float* wbuf = smem_f + (size_t)p.sec_off[p.n_linear] * T;  // one step block
    const int L = p.n_linear;
    const int64_t row = (int64_t)blockIdx.x * T + tid;
    const bool row_ok = row < p.B;
    float* S_in = state + (size_t)p.sec_off[0] * T + tid;

    // x = zeros_like(y) (transforms.py:995); context and y are constant inputs
    for (int k = 0; k < p.D; ++k) S_in[k * T] = 0.f;
    for (int k = 0; k < p.C; ++k) S_in[(p.D + k) * T] = row_ok ? p.c[row * p.ldc + k] : 0.f;
    for (int l = 1; l < L; ++l) {
        float* Sh = state + (size_t)p.sec_off[l] * T + tid;
        for (int k = 0; k < p.dims[l]; ++k) Sh[k * T] = 0.f;
    }

    for (int step = 0; step < p.passes; ++step) {
        const int off = p.step_off[step], len = p.step_off[step + 1] - off;
        __syncthreads();  // previous step's weights are no longer read
        {
            const float4* src = reinterpret_cast<const float4*>(p.stream + off);
            float4* dst = reinterpret_cast<float4*>(wbuf);
            for (int i = tid; i < len / 4; i += T) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        // header: [n_tiles_0 .. n_tiles_{L-2}] [n_dims] [pad to 4] then per hidden tile 8 dest ids,
        // then the dim ids (padded to 4), then the tile data
        const int* hdr = reinterpret_cast<const int*>(wbuf);
        int hpos = (L + 3) & ~3;  // L header ints (L-1 tile counts + n_dims), padded
        const int n_dims = hdr[L - 1];
        int total_hidden_tiles = 0;
        for (int l = 0; l < L - 1; ++l) total_hidden_tiles += hdr[l];
        const int* dest = hdr + hpos;
        const int* dim_ids = dest + total_hidden_tiles * TILE;
        const float* wp = wbuf + hpos + total_hidden_tiles * TILE + ((n_dims + 3) & ~3);
        // ---- hidden units that become final at this step ----
        int tile_idx = 0;
        for (int l = 0; l < L - 1; ++l) {
            const int K = p.dims[l];
            const float* Sl = state + (size_t)p.sec_off[l] * T + tid;
            float* So = state + (size_t)p.sec_off[l + 1] * T + tid;
            for (int t = 0; t < hdr[l]; ++t, ++tile_idx) {
                float acc[TILE];
#pragma unroll
                for (int j = 0; j < TILE; ++j) acc[j] = wp[K * TILE + j];  // bias follows the tile
                tile_dot(wp, Sl, K, T, acc);
                wp += (K + 1) * TILE;
#pragma unroll
                for (int j = 0; j < TILE; ++j) {
                    const int u = dest[tile_idx * TILE + j];
                    if (u >= 0) So[u * T] = fmaxf(acc[j], 0.f);
                }
            }
            // a later layer of this step reads what this layer just wrote — same thread, same
            // column of the state: no barrier needed
        }
        // ---- dims of this order class: parameters -> inverse bijector ----
        {
            const int K = p.dims[L - 1];
            const float* Sl = state + (size_t)p.sec_off[L - 1] * T + tid;
            for (int i = 0; i < n_dims; ++i) {
                const int d = dim_ids[i];
                const float yv = row_ok ? __ldg(p.y + row * p.ldy + d) : 0.f;  // issued early: overlaps the dot products
                float phi[PT * TILE];
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    float acc[TILE];
#pragma unroll
                    for (int j = 0; j < TILE; ++j) acc[j] = wp[K * TILE + j];
                    tile_dot(wp, Sl, K, T, acc);
                    wp += (K + 1) * TILE;
#pragma unroll
                    for (int j = 0; j < TILE; ++j) phi[t * TILE + j] = acc[j];
                }
                float xv;
                if constexpr (UNI == ZK_UNI_RQS) {
                    Bin b = rqs_select<KT, FAST, true>(phi, KT, yv, p.bound, p.aw, p.ad);
                    xv = rqs_inverse_eval<FAST>(b, yv);
                    if (p.circ) xv = circ_shift(xv, p.bound);  // inverse of flows/spline.py:68-71
                } else {
                    const float ls = softclip<FAST>(phi[1], p.ad);
                    xv = zdiv<FAST>(yv - phi[0], zexp<FAST>(ls));
                }
                S_in[d * T] = xv;
                if (row_ok) p.x[row * p.ldx + d] = xv;
            }
        }
    }
}

template <int UNI, int KT>
zk_status launch_inv_t(const InvParams& p, bool fast, int grid, int T, size_t smem, cudaStream_t st) {
    auto go = [&](auto kern) -> zk_status {
        ZK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, T, smem, st>>>(p);
        return check_launch("ar_inverse_kernel");
    };
    if (fast) return go(ar_inverse_kernel<UNI, KT, true>);
    return go(ar_inverse_kernel<UNI, KT, false>);
}

}  // namespace

void ar_inverse_free(ArInvPack* pk) {
    if (!pk) return;
    cudaFree(pk->stream);
    cudaFree(pk->step_off);
    delete pk;
}

zk_status ar_inverse_pack(const zk_mlp* m, const uint8_t* const* mask_dev, const int64_t* order, int D, int C,
                          int uni, int bins, int passes, ArInvPack** out) {
    *out = nullptr;
    const int L = m->n_linear;
    if (!order || L < 1 || L > 7) return ZK_OK;
    if (m->act != 1 || !m->plain) return ZK_OK;  // ReLU MLPs only: other activations / residual blocks use the sweeps
    if (uni == ZK_UNI_RQS && bins != 8 && bins != 16) return ZK_OK;
    if (m->dims[0] != D + C) return ZK_OK;
    const int P = (uni == ZK_UNI_RQS) ? 3 * bins - 1 : 2;
    if (m->dims[L] != D * P) return ZK_OK;
    for (int d = 0; d < D; ++d)
        if (order[d] < 0 || order[d] >= passes) return ZK_OK;

    // host copies of the masked weights, biases and masks
    std::vector<std::vector<float>> W(L), Bv(L);
    std::vector<std::vector<uint8_t>> Mk(L);
    for (int l = 0; l < L; ++l) {
        const size_t n = (size_t)m->dims[l + 1] * m->dims[l];
        W[l].resize(n);
        Bv[l].resize(m->dims[l + 1]);
        Mk[l].assign(n, 1);
        ZK_CUDA(cudaMemcpy(W[l].data(), m->w[l], n * 4, cudaMemcpyDeviceToHost));
        ZK_CUDA(cudaMemcpy(Bv[l].data(), m->b[l], (size_t)m->dims[l + 1] * 4, cudaMemcpyDeviceToHost));
        if (mask_dev && mask_dev[l]) ZK_CUDA(cudaMemcpy(Mk[l].data(), mask_dev[l], n, cudaMemcpyDeviceToHost));
    }
    // step at which every unit's inputs are final: 1 + max order class of the x dims it depends on
    std::vector<std::vector<int>> ready(L);
    for (int l = 0; l < L; ++l) {
        const int K = m->dims[l], N = m->dims[l + 1];
        ready[l].assign(N, 0);
        for (int n = 0; n < N; ++n) {
            int r = 0;
            for (int k = 0; k < K; ++k) {
                if (!Mk[l][(size_t)n * K + k]) continue;
                const int rk = (l == 0) ? (k < D ? (int)order[k] + 1 : 0) : ready[l - 1][k];
                r = std::max(r, rk);
            }
            ready[l][n] = r;
        }
    }
    // every parameter row of dim d must be computable at step order[d] (autoregressive property)
    for (int d = 0; d < D; ++d)
        for (int q = 0; q < P; ++q)
            if (ready[L - 1][d * P + q] > order[d]) return ZK_OK;  // not a pure order-class structure

    ArInvPack* pk = new ArInvPack();
    pk->D = D; pk->C = C; pk->P = P; pk->uni = uni; pk->bins = bins; pk->passes = passes; pk->n_linear = L;
    pk->dims = m->dims;
    int off = 0;
    pk->sec_off.push_back(off);
    off += D + C;
    for (int l = 1; l < L; ++l) { pk->sec_off.push_back(off); off += m->dims[l]; }
    pk->sec_off.push_back(off);  // end
    pk->state_floats = off;

    std::vector<float> stream;
    pk->h_step_off.assign(passes + 1, 0);
    const int PT = (P + TILE - 1) / TILE;
    auto as_float = [](int v) { float f; memcpy(&f, &v, 4); return f; };
    for (int step = 0; step < passes; ++step) {
        pk->h_step_off[step] = (int)stream.size();
        std::vector<std::vector<int>> rows(L - 1 > 0 ? L - 1 : 0);
        for (int l = 0; l < L - 1; ++l)
            for (int n = 0; n < m->dims[l + 1]; ++n)
                if (ready[l][n] == step) rows[l].push_back(n);
        std::vector<int> dims_p;
        for (int d = 0; d < D; ++d)
            if (order[d] == step) dims_p.push_back(d);
        // header
        std::vector<int> hdr((L + 3) & ~3, 0);
        int total_tiles = 0;
        for (int l = 0; l < L - 1; ++l) { hdr[l] = ((int)rows[l].size() + TILE - 1) / TILE; total_tiles += hdr[l]; }
        hdr[L - 1] = (int)dims_p.size();
        for (int v : hdr) stream.push_back(as_float(v));
        for (int l = 0; l < L - 1; ++l)
            for (int t = 0; t < hdr[l]; ++t)
                for (int j = 0; j < TILE; ++j) {
                    const size_t i = (size_t)t * TILE + j;
                    stream.push_back(as_float(i < rows[l].size() ? rows[l][i] : -1));
                }
        for (size_t i = 0; i < ((dims_p.size() + 3) & ~(size_t)3); ++i) stream.push_back(as_float(i < dims_p.size() ? dims_p[i] : -1));
        // tile data: K x 8 weights (k-major) followed by 8 biases
        auto emit_tile = [&](int l, const int* row_ids) {
            const int K = m->dims[l];
            for (int k = 0; k < K; ++k)
                for (int j = 0; j < TILE; ++j) stream.push_back(row_ids[j] >= 0 ? W[l][(size_t)row_ids[j] * K + k] : 0.f);
            for (int j = 0; j < TILE; ++j) stream.push_back(row_ids[j] >= 0 ? Bv[l][row_ids[j]] : 0.f);
        };
        for (int l = 0; l < L - 1; ++l)
            for (int t = 0; t < hdr[l]; ++t) {
                int ids[TILE];
                for (int j = 0; j < TILE; ++j) {
                    const size_t i = (size_t)t * TILE + j;
                    ids[j] = i < rows[l].size() ? rows[l][i] : -1;
                }
                emit_tile(l, ids);
            }
        for (int d : dims_p)
            for (int t = 0; t < PT; ++t) {
                int ids[TILE];
                for (int j = 0; j < TILE; ++j) {
                    const int q = t * TILE + j;
                    ids[j] = q < P ? d * P + q : -1;
                }
                emit_tile(L - 1, ids);
            }
        pk->max_step_words = std::max(pk->max_step_words, (int)stream.size() - pk->h_step_off[step]);
    }
    pk->h_step_off[passes] = (int)stream.size();
    zk_status st = ZK_OK;
    if (cudaMalloc((void**)&pk->stream, std::max<size_t>(stream.size(), 4) * 4) != cudaSuccess ||
        cudaMalloc((void**)&pk->step_off, (size_t)(passes + 1) * 4) != cudaSuccess)
        st = fail(ZK_ENOMEM, "ar_inverse_pack: cudaMalloc failed");
    if (st == ZK_OK && (cudaMemcpy(pk->stream, stream.data(), stream.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                        cudaMemcpy(pk->step_off, pk->h_step_off.data(), (size_t)(passes + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess))
        st = fail(ZK_ECUDA, "ar_inverse_pack: upload failed");
    if (st != ZK_OK) {
        ar_inverse_free(pk);
        return st;
    }
    *out = pk;
    return ZK_OK;
}

bool ar_inverse_threads(const ArInvPack* pk, int* threads, size_t* smem) {
    const size_t wbytes = (size_t)pk->max_step_words * 4 + 16;
    const size_t per_thread = (size_t)pk->state_floats * 4;
    // two CTAs of 128 samples per SM when they fit (8 warps hide the FMA / LDS latencies of the
    // dot products far better than 4), else one CTA with as many samples as 200 KB hold
    if (per_thread * 128 + wbytes <= 112 * 1024) {
        *threads = 128;
        *smem = per_thread * 128 + wbytes;
        return true;
    }
    const size_t budget = 200 * 1024;
    if (wbytes >= budget) return false;
    int T = (int)((budget - wbytes) / per_thread);
    T = std::min(128, T / 32 * 32);
    if (T < 32) return false;
    *threads = T;
    *smem = per_thread * T + wbytes;
    return true;
}

zk_status launch_ar_inverse(const ArInvPack* pk, const float* y, int64_t ldy, const float* c, int64_t ldc, int64_t B,
                            float* x, int64_t ldx, float bound, float slope, bool fast, bool circular, cudaStream_t st) {
    if (B == 0) return ZK_OK;
    int T = 0;
    size_t smem = 0;
    ZK_REQUIRE(ar_inverse_threads(pk, &T, &smem), "ar_inverse: state does not fit shared memory");
    InvParams p;
    p.stream = pk->stream; p.step_off = pk->step_off; p.passes = pk->passes; p.n_linear = pk->n_linear;
    p.D = pk->D; p.C = pk->C; p.P = pk->P;
    for (int i = 0; i <= pk->n_linear; ++i) p.dims[i] = pk->dims[i];
    for (size_t i = 0; i < pk->sec_off.size(); ++i) p.sec_off[i] = pk->sec_off[i];
    p.y = y; p.ldy = ldy; p.c = c; p.ldc = ldc; p.x = x; p.ldx = ldx; p.B = B;
    p.bound = bound;
    p.circ = circular ? 1 : 0;
    const float absL = fabsf(logf(slope));
    p.aw = 2.f / absL;
    p.ad = 1.f / absL;
    const int64_t grid = ceil_div(B, T);
    ZK_REQUIRE(grid <= 0x7fffffff, "ar_inverse: batch too large for one launch");
    if (pk->uni == ZK_UNI_RQS && pk->bins == 8) return launch_inv_t<ZK_UNI_RQS, 8>(p, fast, (int)grid, T, smem, st);
    if (pk->uni == ZK_UNI_RQS && pk->bins == 16) return launch_inv_t<ZK_UNI_RQS, 16>(p, fast, (int)grid, T, smem, st);
    return launch_inv_t<ZK_UNI_AFFINE, 0>(p, fast, (int)grid, T, smem, st);
}

}  // namespace zk

// zuko_b200 — ONE kernel per flow layer: conditioner (all linear layers) + bijector + ladj.
//
//   x, c  --(bf16 hi/lo split, in-kernel)-->  A operand in shared memory
//   for every linear layer:  tcgen05.mma (A: smem, W: TMA-streamed from L2, D: TMEM)
//        hidden layers : epilogue (TMEM -> bias, ReLU -> hi/lo split) writes the next A operand
//                        straight back into shared memory in the canonical 128B-swizzled
//                        K-major layout — activations never touch HBM
//        last layer    : epilogue thread = sample row reads its D*P raw parameters from TMEM
//                        (tcgen05.ld 32x32b) and evaluates the spline / affine bijector, the
//                        log-derivative and the per-sample sum in registers — phi never
//                        touches HBM either ("lane = sample row")
//   HBM traffic per sample per layer: 4 (D + C) read + 4 (D + 1) written  (cfg2: 164 B instead of
//   the 9.3 kB of the unfused path) — the kernel is bound by the tensor pipe / L2 weight stream.
//
// Restates for one MaskedAutoregressiveTransform layer:  flows/autoregressive.py:207-215 (meta),
// nn.py:217-218 (masked linears, mask folded into W at pack time), transforms.py:469-490,
// 554-567 (RQS) or 426-446 (affine), transforms.py:210-214 (sum over D), and on the last
// layer of a flow distributions.py:115-119 + torch normal.py:87-102 (DiagNormal log-prob).
//
// Warp roles (384 threads, 1 CTA / SM, persistent over 128-row tiles):
//   warp 0      W producer   (TMA, 3-stage ring of 128 x 64 bf16 hi/lo tiles)
//   warp 1      MMA issuer   (one lane; M128 x N<=128 x K16, 3 MMAs per k-step: hh, hl, lh)
//   warp 2      TMEM allocator (256 columns = 2 accumulator buffers of 128)
//   warps 4-11  epilogue: two warp sets, set s handles half of the columns / dims of a chunk

#include "bijector_math.cuh"
#include "fused_layer.cuh"
#include "tc_common.cuh"

namespace zk {

namespace {

using namespace bij;

constexpr int FM = 128;             // rows per tile
constexpr int FK = 64;              // bf16 per K block (128-byte swizzle row)
constexpr int F_WSTAGES = 3;
constexpr int F_MAXKB = 4;          // A operand: up to 4 K blocks = 256 columns
constexpr int F_THREADS = 384;
constexpr int F_EPI_WARP0 = 4;
constexpr uint32_t F_PLANE = FM * FK * 2;       // 16 KB: one plane of one K block
constexpr uint32_t F_KBLOCK = 2 * F_PLANE;      // hi + lo
constexpr uint32_t F_A_BYTES = F_MAXKB * F_KBLOCK;          // 128 KB
constexpr uint32_t F_W_BYTES = F_WSTAGES * F_KBLOCK;        // 96 KB
constexpr uint32_t F_AUX_BYTES = 2048;                      // barriers + ladj scratch
constexpr size_t F_SMEM = (size_t)F_A_BYTES + F_W_BYTES + F_AUX_BYTES + 1024 /*alignment slack*/;

struct FusedParams {
    CUtensorMap mapW[ZK_FUSED_MAX_LINEAR];
    const float* bias[ZK_FUSED_MAX_LINEAR];
    int n_linear;
    int K0, KB0;        // real input width (D + C) and its number of 64-wide K blocks
    int H, CW;          // hidden width (multiple of 64, <= 256), hidden chunk width (128 or 64)
    int D, C;
    int n_last_chunks;  // ceil(D / DPC)
    int n_terms;        // 3 (split bf16) or 1
    int M;
    const float* x; int64_t ldx;
    const float* c; int64_t ldc;
    float* y; int64_t ldy;
    float* ladj; int accumulate;
    float* log_prob; const float* base_loc; const float* base_scale;
    float bound, aw, ad;
    long long* dbg;  // optional timeline buffer (clock64 stamps of CTA 0, third tile), see zk_debug_timeline
};

// 16 bytes into the canonical K-major SWIZZLE_128B tile: row r, 16-byte chunk c (8 bf16)
__device__ __forceinline__ void st_swizzled(uint8_t* tile, int r, int c, uint32_t a, uint32_t b, uint32_t cc, uint32_t d) {
    const uint32_t addr = smem_u32(tile) + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(cc), "r"(d) : "memory");
}

#define ZK_STAMP(slot)                                                        \
    do {                                                                      \
        if (p.dbg != nullptr && blockIdx.x == 0 && stamp_on) p.dbg[(slot)] = clock64(); \
    } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// per-(UNI, K) chunking of the last layer: DPC dims per 128-column accumulator chunk
template <int UNI, int KT>
struct LastCfg;
template <>
struct LastCfg<ZK_UNI_RQS, 8> { static constexpr int P = 23, DPC = 4; };
template <>
struct LastCfg<ZK_UNI_RQS, 16> { static constexpr int P = 47, DPC = 2; };
template <>
struct LastCfg<ZK_UNI_AFFINE, 0> { static constexpr int P = 2, DPC = 64; };

template <int UNI, int KT, bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1) fused_layer_kernel(const __grid_constant__ FusedParams p) {
    using Cfg = LastCfg<UNI, KT>;
    constexpr int P = Cfg::P, DPC = Cfg::DPC;
    constexpr int N_LAST = (DPC * P + 15) & ~15;       // MMA N of a last-layer chunk
    constexpr int DIMS_A = (DPC + 1) / 2;              // dims of a chunk handled by warp set 0
    constexpr int BASE_B = (DIMS_A * P) & ~31;         // first TMEM column loaded by warp set 1
    static_assert(DIMS_A * P <= 64 && DPC * P - BASE_B <= 64 && N_LAST <= 128, "chunk windows must fit 64 columns");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW = smem + F_A_BYTES;
    uint64_t* bars = (uint64_t*)(smem + F_A_BYTES + F_W_BYTES);
    uint64_t* w_full = bars;                  // [3]
    uint64_t* w_empty = bars + 3;             // [3]
    uint64_t* d_full = bars + 6;              // [2]
    uint64_t* d_empty = bars + 8;             // [2]
    uint64_t* a_ready = bars + 10;            // [4]  K block kb of the A operand written
    uint64_t* layer_done = bars + 14;         // [1]  all MMAs issued so far have completed
    uint32_t* tmem_slot = (uint32_t*)(bars + 15);
    float* s_part = (float*)(bars + 16);      // [2][128] ladj partials of warp set 1
    volatile int* s_tile = (volatile int*)(s_part + 2 * FM);  // tiles started by the epilogue (paces the prefetcher)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int L = p.n_linear;
    const int m_tiles = (p.M + FM - 1) / FM;
    // CTA pairs (clusters of 2) walk the tiles together: pair `cid` takes tiles 2 (cid + i ncl) + rank.
    // Both CTAs always run the same number of iterations (the W stream is shared through TMA
    // multicast); a CTA whose tile index falls off the end processes an all-masked dummy tile.
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int n_iter = ((m_tiles + 1) / 2 - cid + ncl - 1) / ncl;  // iterations of this pair (>= 0)
    const int KBH = p.H / FK;
    const int nch_hidden = p.H / p.CW;

    if (threadIdx.x == 0) {
        for (int s = 0; s < F_WSTAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 2); }  // empty: both CTAs of the pair
        for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], 256); }
        for (int k = 0; k < F_MAXKB; ++k) mbar_init(&a_ready[k], 256);
        mbar_init(layer_done, 1);
        *s_tile = 0;
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= W producer =======================
        if (lane == 0) {
            const uint32_t bytes = (p.n_terms == 3) ? F_KBLOCK : F_PLANE;
            int ws = 0;
            uint32_t wph = 0;
            for (int it = 0; it < n_iter; ++it) {
                for (int l = 0; l < L; ++l) {
                    const bool last = (l == L - 1);
                    const int nch = last ? p.n_last_chunks : nch_hidden;
                    const int KB = (l == 0) ? p.KB0 : KBH;
                    for (int ch = 0; ch < nch; ++ch) {
                        const int n0 = last ? ch * DPC * P : ch * p.CW;
                        for (int kb = 0; kb < KB; ++kb) {
                            // the slot is written in BOTH CTAs: wait until both MMA issuers released it
                            mbar_wait(&w_empty[ws], wph ^ 1);
                            // this CTA fetches rows [64 rank, 64 rank + 64) of the 128-row W tile and
                            // multicasts them to the pair; the peer delivers the other half
                            uint8_t* st = sW + (size_t)ws * F_KBLOCK + rank * (F_PLANE / 2);
                            mbar_arrive_expect_tx(&w_full[ws], bytes);
                            tma_load_3d_mc(st, &p.mapW[l], &w_full[ws], kb * FK, n0 + 64 * (int)rank, 0, (uint16_t)3);
                            if (p.n_terms == 3)
                                tma_load_3d_mc(st + F_PLANE, &p.mapW[l], &w_full[ws], kb * FK, n0 + 64 * (int)rank, 1, (uint16_t)3);
                            if (++ws == F_WSTAGES) { ws = 0; wph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0, chunk = 0, a_par = 0;  // a_par: bit kb = parity of a_ready[kb]
            for (int mma_tile = 0; mma_tile < n_iter; ++mma_tile) {
                const bool stamp_on = (mma_tile == 2);
                for (int l = 0; l < L; ++l) {
                    const bool last = (l == L - 1);
                    const int nch = last ? p.n_last_chunks : nch_hidden;
                    const int KB = (l == 0) ? p.KB0 : KBH;
                    const uint32_t idesc = umma_idesc_bf16(FM, last ? N_LAST : p.CW);
                    for (int ch = 0; ch < nch; ++ch, ++chunk) {
                        const uint32_t buf = chunk & 1u;
                        mbar_wait(&d_empty[buf], ((chunk >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + buf * 128u;
                        for (int kb = 0; kb < KB; ++kb) {
                            if (ch == 0) {  // first use of this K block in this layer
                                mbar_wait(&a_ready[kb], (a_par >> kb) & 1u);
                                a_par ^= (1u << kb);
                                if (kb == 0) ZK_STAMP(8 * l + 0);
                                if (kb == KB - 1) ZK_STAMP(8 * l + 1);
                            }
                            mbar_wait(&w_full[ws], wph);
                            if (ch == 0 && kb == 0) ZK_STAMP(8 * l + 2);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(sA + (size_t)kb * F_KBLOCK), a_lo = a_hi + F_PLANE;
                            const uint32_t w_hi = smem_u32(sW + (size_t)ws * F_KBLOCK), w_lo = w_hi + F_PLANE;
#pragma unroll
                            for (int k = 0; k < FK / 16; ++k) {
                                const uint32_t off = (uint32_t)k * 32u;
                                umma_bf16(d_tmem, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(w_hi + off), idesc,
                                          (kb > 0 || k > 0) ? 1u : 0u);
                                if (p.n_terms == 3) {
                                    umma_bf16(d_tmem, umma_desc_k_sw128(a_hi + off), umma_desc_k_sw128(w_lo + off), idesc, 1u);
                                    umma_bf16(d_tmem, umma_desc_k_sw128(a_lo + off), umma_desc_k_sw128(w_hi + off), idesc, 1u);
                                }
                            }
                            umma_commit_mc(&w_empty[ws], (uint16_t)3);  // releases the slot in both CTAs
                            if (++ws == F_WSTAGES) { ws = 0; wph ^= 1; }
                        }
                        umma_commit(&d_full[buf]);
                        if (ch == nch - 1) {
                            umma_commit(layer_done);
                            ZK_STAMP(8 * l + 3);
                        }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ======================= input prefetcher =======================
        // pulls the NEXT tile's x / c rows into L2 while the current tile computes
        for (int it = 0; it + 1 < n_iter; ++it) {
            const int tn = 2 * (cid + (it + 1) * ncl) + (int)rank;
            if (tn >= m_tiles) break;
            while (*s_tile < it + 1) __nanosleep(500);  // stay exactly one tile ahead of the epilogue
            const int64_t r0 = (int64_t)tn * FM;
            const int rows = (int)min((int64_t)FM, (int64_t)p.M - r0);
            if (p.ldx == p.D) {
                const char* base = (const char*)(p.x + r0 * p.ldx);
                for (int off = lane * 128; off < rows * p.D * 4; off += 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
            if (p.C > 0 && p.ldc == p.C) {
                const char* base = (const char*)(p.c + r0 * p.ldc);
                for (int off = lane * 128; off < rows * p.C * 4; off += 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
        }
    } else if (warp >= F_EPI_WARP0) {
        // ======================= epilogue =======================
        const int s = (warp - F_EPI_WARP0) >> 2;  // warp set 0 / 1
        const int q = warp & 3;                   // TMEM lane quadrant
        const int r = q * 32 + lane;              // row inside the tile
        uint32_t chunk = 0, ld_par = 0;
        for (int tile_iter = 0; tile_iter < n_iter; ++tile_iter) {
            const int t = 2 * (cid + tile_iter * ncl) + (int)rank;  // may be >= m_tiles: dummy tile, all rows masked
            const bool stamp_on = (tile_iter == 2) && (threadIdx.x == F_EPI_WARP0 * 32);
            ZK_STAMP(48);
            if (threadIdx.x == F_EPI_WARP0 * 32) *s_tile = tile_iter + 1;
            const int64_t row = (int64_t)t * FM + r;
            const bool row_ok = row < p.M;
            // ---- stage the layer-0 operand: cat(x, c) -> bf16 hi/lo, K block kb by set (kb & 1) ----
            for (int kb = s; kb < p.KB0; kb += 2) {
                uint8_t* hi = sA + (size_t)kb * F_KBLOCK;
                uint8_t* lo = hi + F_PLANE;
                // issue every load of the K block before the first conversion / store so that
                // their latencies overlap (the rows were prefetched into L2 by warp 3)
                float vals[FK];
#pragma unroll
                for (int j = 0; j < FK; ++j) {
                    const int k = kb * FK + j;
                    float val = 0.f;
                    if (row_ok && k < p.K0) val = (k < p.D) ? __ldg(p.x + row * p.ldx + k) : __ldg(p.c + row * p.ldc + (k - p.D));
                    vals[j] = val;
                }
#pragma unroll
                for (int cidx = 0; cidx < 8; ++cidx) {
                    uint32_t ph[4], pl[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split2_bf16(vals[cidx * 8 + 2 * j], vals[cidx * 8 + 2 * j + 1], ph[j], pl[j]);
                    st_swizzled(hi, r, cidx, ph[0], ph[1], ph[2], ph[3]);
                    st_swizzled(lo, r, cidx, pl[0], pl[1], pl[2], pl[3]);
                }
            }
            fence_proxy_async();
            for (int kb = 0; kb < p.KB0; ++kb) mbar_arrive(&a_ready[kb]);
            ZK_STAMP(49);

            // ---- hidden layers: D -> bias, ReLU -> hi/lo -> next A operand (shared memory) ----
            for (int l = 0; l < L - 1; ++l) {
                const float* bias = p.bias[l];
                for (int ch = 0; ch < nch_hidden; ++ch, ++chunk) {
                    const uint32_t buf = chunk & 1u;
                    mbar_wait(&d_full[buf], (chunk >> 1) & 1u);
                    tc_fence_after();
                    ZK_STAMP(64 + 16 * l + 4 * ch + 0);
                    const int ncols = p.CW >> 1;             // columns handled by this thread: 64 or 32
                    const int col0 = s * ncols;              // first column inside the chunk
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128u + (uint32_t)col0;
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(taddr, ra);
                    if (ncols == 64) tmem_ld_32x32b_x32(taddr + 32u, rb);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(&d_empty[buf]);  // accumulator buffer is free again
                    uint32_t ph[32], pl[32];
                    const int nbase = ch * p.CW + col0;
                    const float4* b4 = reinterpret_cast<const float4*>(bias + nbase);  // 128-byte aligned
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bb = __ldg(b4 + (j >> 2));  // one broadcast 16-byte load per 4 columns
                        split2_bf16(fmaxf(__uint_as_float(ra[j]) + bb.x, 0.f), fmaxf(__uint_as_float(ra[j + 1]) + bb.y, 0.f),
                                    ph[j >> 1], pl[j >> 1]);
                        split2_bf16(fmaxf(__uint_as_float(ra[j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(ra[j + 3]) + bb.w, 0.f),
                                    ph[(j >> 1) + 1], pl[(j >> 1) + 1]);
                    }
                    if (ncols == 64) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = __ldg(b4 + 8 + (j >> 2));
                            split2_bf16(fmaxf(__uint_as_float(rb[j]) + bb.x, 0.f), fmaxf(__uint_as_float(rb[j + 1]) + bb.y, 0.f),
                                        ph[16 + (j >> 1)], pl[16 + (j >> 1)]);
                            split2_bf16(fmaxf(__uint_as_float(rb[j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(rb[j + 3]) + bb.w, 0.f),
                                        ph[17 + (j >> 1)], pl[17 + (j >> 1)]);
                        }
                    }
                    ZK_STAMP(64 + 16 * l + 4 * ch + 1);
                    if (ch == 0) {  // the A operand may be overwritten once ALL MMAs of this layer are done
                        mbar_wait(layer_done, ld_par);
                        ld_par ^= 1u;
                    }
                    ZK_STAMP(64 + 16 * l + 4 * ch + 2);
                    if (ncols == 64) {  // CW = 128: this set owns K block 2 ch + s entirely
                        const int kb = ch * 2 + s;
                        uint8_t* hi = sA + (size_t)kb * F_KBLOCK;
                        uint8_t* lo = hi + F_PLANE;
#pragma unroll
                        for (int cidx = 0; cidx < 8; ++cidx) {
                            st_swizzled(hi, r, cidx, ph[cidx * 4], ph[cidx * 4 + 1], ph[cidx * 4 + 2], ph[cidx * 4 + 3]);
                            st_swizzled(lo, r, cidx, pl[cidx * 4], pl[cidx * 4 + 1], pl[cidx * 4 + 2], pl[cidx * 4 + 3]);
                        }
                        fence_proxy_async();
                        mbar_arrive(&a_ready[ch * 2]);
                        mbar_arrive(&a_ready[ch * 2 + 1]);
                        ZK_STAMP(64 + 16 * l + 4 * ch + 3);
                    } else {  // CW = 64: the two sets share K block ch (32 columns = 4 chunks each)
                        uint8_t* hi = sA + (size_t)ch * F_KBLOCK;
                        uint8_t* lo = hi + F_PLANE;
#pragma unroll
                        for (int cidx = 0; cidx < 4; ++cidx) {
                            st_swizzled(hi, r, s * 4 + cidx, ph[cidx * 4], ph[cidx * 4 + 1], ph[cidx * 4 + 2], ph[cidx * 4 + 3]);
                            st_swizzled(lo, r, s * 4 + cidx, pl[cidx * 4], pl[cidx * 4 + 1], pl[cidx * 4 + 2], pl[cidx * 4 + 3]);
                        }
                        fence_proxy_async();
                        mbar_arrive(&a_ready[ch]);
                    }
                }
            }

            // ---- last layer: raw parameters stay in TMEM -> bijector + ladj in registers ----
            float lsum = 0.f;
            const float* bias = p.bias[L - 1];
            for (int ch = 0; ch < p.n_last_chunks; ++ch, ++chunk) {
                const uint32_t buf = chunk & 1u;
                mbar_wait(&d_full[buf], (chunk >> 1) & 1u);
                tc_fence_after();
                if (ch < 8) ZK_STAMP(160 + 2 * ch);
                const int base = s ? BASE_B : 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128u + (uint32_t)base;
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(taddr, ra);
                tmem_ld_32x32b_x32(taddr + 32u, rb);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&d_empty[buf]);
                float v[64];
                const int nbase = ch * DPC * P + base;  // global output column of v[0]
                const int n_total = p.D * P;
                const float2* b2 = reinterpret_cast<const float2*>(bias + nbase);  // nbase is even
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    // float2 pairs are entirely inside or outside [0, n_total): both bounds are even
                    const float2 ba = (nbase + j < n_total) ? __ldg(b2 + (j >> 1)) : make_float2(0.f, 0.f);
                    const float2 bb = (nbase + 32 + j < n_total) ? __ldg(b2 + 16 + (j >> 1)) : make_float2(0.f, 0.f);
                    v[j] = __uint_as_float(ra[j]) + ba.x;
                    v[j + 1] = __uint_as_float(ra[j + 1]) + ba.y;
                    v[32 + j] = __uint_as_float(rb[j]) + bb.x;
                    v[33 + j] = __uint_as_float(rb[j + 1]) + bb.y;
                }
                auto do_dim = [&](const float* pp, int dloc) {
                    const int d = ch * DPC + dloc;
                    if (d >= p.D || !row_ok) return;
                    const float xv = p.x[row * p.ldx + d];
                    float yv, lj;
                    if constexpr (UNI == ZK_UNI_RQS) {
                        Bin b = rqs_select<KT, FAST, false>(pp, KT, xv, p.bound, p.aw, p.ad);
                        rqs_forward_eval<FAST>(b, xv, yv, lj);
                    } else {
                        const float ls = softclip<FAST>(pp[1], p.ad);
                        yv = fmaf(xv, zexp<FAST>(ls), pp[0]);
                        lj = ls;
                    }
                    if (p.y) p.y[row * p.ldy + d] = yv;
                    if (p.log_prob) {
                        const float mu = p.base_loc ? p.base_loc[d] : 0.f;
                        const float sg = p.base_scale ? p.base_scale[d] : 1.f;
                        const float u = (yv - mu) / sg;
                        lj += -0.5f * u * u - logf(sg) - kHalfLog2Pi;
                    }
                    lsum += lj;
                };
                if (s == 0) {
#pragma unroll
                    for (int j = 0; j < DIMS_A; ++j) do_dim(&v[j * P], j);
                } else {
#pragma unroll
                    for (int j = DIMS_A; j < DPC; ++j) do_dim(&v[j * P - BASE_B], j);
                }
                if (ch < 8) ZK_STAMP(161 + 2 * ch);
            }
            // all MMAs of this tile are complete once the last layer_done fires: A may be restaged
            mbar_wait(layer_done, ld_par);
            ld_par ^= 1u;
            // ---- per-sample sum: set 1 hands its partial to set 0 ----
            float* part = s_part + (tile_iter & 1) * FM;
            if (s == 1) part[r] = lsum;
            epi_bar_sync();
            if (s == 0 && row_ok) {
                const float tot = lsum + part[r] + (p.accumulate ? p.ladj[row] : 0.f);
                if (p.log_prob) p.log_prob[row] = tot;
                else if (p.ladj) p.ladj[row] = tot;
            }
            ZK_STAMP(50);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

template <int UNI, int KT>
zk_status launch_fused_t(const FusedParams& p, bool fast, int grid, cudaStream_t st) {
    auto go = [&](auto kern) -> zk_status {
        ZK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
        kern<<<grid, F_THREADS, F_SMEM, st>>>(p);
        return check_launch("fused_layer_kernel");
    };
    if (fast) return go(fused_layer_kernel<UNI, KT, true>);
    return go(fused_layer_kernel<UNI, KT, false>);
}

}  // namespace

long long* g_timeline = nullptr;

bool fused_layer_supported(const zk_mlp* m, int univariate, int bins, int D, int C) {
    const TcPack* pk = (const TcPack*)m->tc;
    if (!pk || m->gemm_mode == ZK_GEMM_FP32) return false;
    if (m->n_linear < 2 || m->n_linear > ZK_FUSED_MAX_LINEAR) return false;
    const int H = m->dims[1];
    if (H % 64 != 0 || H < 64 || H > 256) return false;
    for (int i = 1; i < m->n_linear; ++i)
        if (m->dims[i] != H) return false;
    if (D + C > 256 || m->dims[0] != D + C) return false;
    if (univariate == ZK_UNI_RQS) return bins == 8 || bins == 16;
    return univariate == ZK_UNI_AFFINE;
}

zk_status launch_fused_layer(const zk_mlp* m, const FusedLayerArgs& a, cudaStream_t st) {
    const TcPack* pk = (const TcPack*)m->tc;
    ZK_REQUIRE(pk && fused_layer_supported(m, a.univariate, a.bins, a.D, a.C), "fused layer: unsupported shape");
    ZK_REQUIRE(a.B < ((int64_t)1 << 31) - FM, "fused layer: batch too large for one launch");
    if (a.B == 0) return ZK_OK;
    FusedParams p;
    for (int i = 0; i < m->n_linear; ++i) {
        p.mapW[i] = pk->layers[i].mapW64;
        p.bias[i] = m->b[i];
    }
    p.n_linear = m->n_linear;
    p.K0 = a.D + a.C;
    p.KB0 = pk->layers[0].Kp / FK;
    p.H = m->dims[1];
    p.CW = (p.H % 128 == 0) ? 128 : 64;
    p.D = a.D; p.C = a.C;
    p.n_terms = pk->n_terms;
    p.M = (int)a.B;
    p.x = a.x; p.ldx = a.ldx; p.c = a.c; p.ldc = a.ldc;
    p.y = a.y; p.ldy = a.ldy; p.ladj = a.ladj; p.accumulate = a.accumulate;
    p.log_prob = a.log_prob; p.base_loc = a.base_loc; p.base_scale = a.base_scale;
    p.bound = a.bound;
    const float absL = fabsf(logf(a.slope));
    p.aw = 2.f / absL;
    p.ad = 1.f / absL;
    p.dbg = g_timeline;
    // clusters of 2 CTAs: even grid, at most one CTA per SM
    const int64_t pairs = ceil_div(ceil_div(a.B, FM), 2);
    const int grid = 2 * (int)std::min<int64_t>(pairs, sm_count() / 2);
    if (a.univariate == ZK_UNI_RQS && a.bins == 8) {
        p.n_last_chunks = (a.D + 3) / 4;
        return launch_fused_t<ZK_UNI_RQS, 8>(p, a.fast_math, grid, st);
    }
    if (a.univariate == ZK_UNI_RQS && a.bins == 16) {
        p.n_last_chunks = (a.D + 1) / 2;
        return launch_fused_t<ZK_UNI_RQS, 16>(p, a.fast_math, grid, st);
    }
    p.n_last_chunks = (a.D + 63) / 64;
    return launch_fused_t<ZK_UNI_AFFINE, 0>(p, a.fast_math, grid, st);
}

}  // namespace zk

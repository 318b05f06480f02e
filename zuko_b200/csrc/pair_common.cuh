// zuko_b200 — what the CTA-pair fused kernels (fused_wide.cu, fused_dual.cu) share: the cta_group::2
// forms of the tcgen05 / TMA / mbarrier instructions, the bounded waits with their watchdog report,
// and the store of an A-operand lo plane fragment into its swizzled shared-memory layout.
#pragma once

#include "fused_common.cuh"

namespace zk {

constexpr uint32_t W_APLANE = 128 * 64 * 2;       // 16 KB: one K block (128 rows x 64 bf16) of the A lo plane
constexpr long long W_WD_CYCLES = 4000000000ll;   // watchdog: ~2 s of SM clock

// ---------------------------------------------------------------------------
// PTX: the cta_group::2 forms
// ---------------------------------------------------------------------------
__device__ __forceinline__ void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once all MMAs issued so far are complete) on the mbarrier at this offset in both CTAs
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// TMA load of this CTA's half of a pair's operand tile; the bytes are counted on the barrier at
// `bar_addr` (a shared::cluster address: the LEADER's barrier, cute::Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_addr,
                                                int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
    return r;
}
// Arrive on a barrier of the LEADER CTA.  Default semantics (.release.cta), as CUTLASS's
// ClusterBarrier::arrive(cta_id): what the arrival publishes lives in THIS CTA's tensor / shared
// memory and was completed before it (tcgen05.wait, fence.proxy.async); `.release.cluster` compiles to
// MEMBAR.ALL.GPU in front of every arrive (~2000 cycles each: profiles/r02_wide_timeline_*_v1.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// ---------------------------------------------------------------------------
// watchdog: a wait that does not end within ~2 s reports (role, what it waits for, where it is in
// the schedule) into mapped host memory and traps — a protocol bug then costs one GPU run and
// names the barrier, instead of hanging the device until the job is killed.
// ---------------------------------------------------------------------------
static __device__ __noinline__ void wd_report(uint32_t* watch, uint32_t code, uint32_t a, uint32_t b) {
    if (watch != nullptr && blockIdx.x < 2) {
        // slot per (CTA of cluster 0, warp): 8 words
        uint32_t* s = watch + 8 + ((blockIdx.x & 1) * 32 + (threadIdx.x >> 5)) * 8;
        s[0] = 0xDEAD0000u | code; s[1] = a; s[2] = b; s[3] = threadIdx.x;
        atomicAdd(watch, 1u);
        __threadfence_system();
    }
    const long long t0 = clock64();
    while (clock64() - t0 < 400000000ll) {}  // let the other roles report too
    __trap();
}
#define WD_SPIN(cond, code, a, b)                                                   \
    do {                                                                            \
        uint32_t _n = 0;                                                            \
        long long _t0 = 0;                                                          \
        while (!(cond)) {                                                           \
            if ((++_n & 1023u) == 0u) {                                             \
                const long long _t = clock64();                                     \
                if (_t0 == 0) _t0 = _t;                                             \
                else if (_t - _t0 > W_WD_CYCLES) wd_report(p.watch, (code), (a), (b)); \
            }                                                                       \
        }                                                                           \
    } while (0)

#define W_STAMP(slot)                                                                         \
    do {                                                                                      \
        if constexpr (DBG) {                                                                  \
            if (p.dbg != nullptr && blockIdx.x == 0 && stamp_on) p.dbg[(slot)] = clock64();   \
        }                                                                                     \
    } while (0)

// lo plane of 32 consecutive K elements of row r -> shared memory, K-major, 128-byte swizzle
// (the layout a TMA SWIZZLE_128B box of 64 bf16 x 128 rows has: 16-byte unit j of row r sits at
// unit j ^ (r & 7) of the row's 128 bytes; 8-row groups are 1024 bytes apart)
__device__ __forceinline__ void st_alo32(uint8_t* sAlo, int kb, int r, int half, const uint32_t* pl) {
    uint8_t* row = sAlo + (size_t)kb * W_APLANE + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = half * 4 + u;
        *reinterpret_cast<uint4*>(row + (((j ^ (r & 7)) & 7) << 4)) = make_uint4(pl[4 * u], pl[4 * u + 1], pl[4 * u + 2], pl[4 * u + 3]);
    }
}


}  // namespace zk

// Micro-benchmark: how fast can ONE warp feed the tensor pipe when its scheduler (SM sub-partition)
// is shared with busy ALU warps?  Warp 1 issues "entries" of 12 (or 24) tcgen05.mma M128 N128 K16
// (A in TMEM) followed by a tcgen05.commit and a handful of bookkeeping instructions, like the MMA
// issuer of the fused layer kernel; `busy` other warps per sub-partition spin on dependent FFMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issuer_contention issuer_contention.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

// mode 0: whole warp walks the loop, elected lane issues (as in the fused kernel)
// mode 1: a single thread runs the whole loop
template <int MODE, int MMAS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1) k(long long* out, const uint32_t* sched, int entries, int busy_warps, int ilp, int flags) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[8], done_bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) {
        if (MODE == 0 || lane == 0) {
            const uint32_t id = idesc(128, 128);
            long long t0 = clock64();
            int ws = 0;
            uint32_t c = 0;
            uint32_t cur = __ldg(sched);
            for (int i = 0; i < entries; ++i) {
                const uint32_t nxt = __ldg(sched + ((i + 1) & 63));
                const uint32_t kb = cur & 3u, buf = c & 1u;
                const uint32_t d_tmem = 256u + buf * 128u, a_hi = kb * 32u;
                const uint64_t db = desc_sw128(smem_u32(base) + (uint32_t)ws * 16384u);
                // the barrier was never armed: parity 1 reads as complete -> one probe, like a ready stage
                if (flags & 1) while (!test_wait(&bar[ws], 1u)) {}
                if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (MODE == 1 || elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < MMAS; ++kk)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_hi + (uint32_t)((kk / 3) & 3) * 8u), "l"(db + 2 * ((kk / 3) & 3)), "r"(id), "r"(1u) : "memory");
                    // commit to a barrier nobody waits for, as w_empty / d_full in the real kernel
                    if (flags & 32) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar[4 + (ws & 3)])), "h"((uint16_t)3) : "memory");
                    else if (flags & 4) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[4 + (ws & 3)])) : "memory");
                    if ((flags & 8) && (cur & 8u)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[4 + ((ws + 1) & 3)])) : "memory");
                }
                if (MODE == 0 && (flags & 16)) __syncwarp();
                c += (nxt >> 2) & 1u;
                cur = nxt;
                if (++ws == 4) ws = 0;
            }
            if (MODE == 1 || elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
            if (MODE == 0) __syncwarp();
            while (!test_wait(&done_bar, 0u)) {}
            const long long t1 = clock64();
            if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
        if (lane == 0) stop = 1;
    } else if (warp >= 4 && ((warp - 4) >> 2) < busy_warps) {
        // `busy_warps` warps on EVERY sub-partition (warp % 4), each with `ilp` independent FFMA chains
        float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
        const float m = 1.0001f, b = 0.5f;
        while (!stop) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                a0 = fmaf(a0, m, b);
                if (ilp > 1) a1 = fmaf(a1, m, b);
                if (ilp > 2) { a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b); }
            }
        }
        if (a0 + a1 + a2 + a3 == 12345.f) out[3] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

template <int MODE, int MMAS>
void run(long long* d, const uint32_t* sched, int busy, int ilp, int flags = 31) {
    const int entries = 2000;
    cudaFuncSetAttribute(k<MODE, MMAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaMemset(d, 0, 32);
    k<MODE, MMAS><<<148, 640, 96 * 1024>>>(d, sched, entries, busy, ilp, flags);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s, %2d MMAs/entry, %d busy warps per sub-partition (ilp %d), flags wait=%d fence=%d commit=%d commit2=%d syncwarp=%d multicast=%d: %.0f cycles / entry (ideal %d) %s\n", MODE ? "single thread " : "warp + elect   ", MMAS, busy, ilp, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, (flags >> 3) & 1, (flags >> 4) & 1, (flags >> 5) & 1,
           (double)h / entries, MMAS * 64, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    uint32_t* sched;
    uint32_t h[64];
    for (int i = 0; i < 64; ++i) h[i] = (uint32_t)(i & 3) | ((i & 3) == 0 ? 4u : 0u) | ((i & 3) == 3 ? 8u : 0u);
    cudaMalloc(&d, 32);
    cudaMalloc(&sched, sizeof(h));
    cudaMemcpy(sched, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int busy : {0, 4}) {
        for (int f : {0, 4, 32, 31, 63, 30, 62}) run<0, 12>(d, sched, busy, 4, f);
    }
    return 0;
}

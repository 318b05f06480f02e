// Micro-benchmark: does TMA (cp.async.bulk) write traffic into shared memory slow down
// tcgen05.mma whose B operand streams from shared memory (A in tensor memory)?  One CTA per SM:
// warp 0 issues M128 N128 K16 MMAs back to back, lane 0 of warp 1 keeps `batch` bytes of bulk
// copies global -> shared in flight every `period` cycles (0 = as fast as they complete).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_contention smem_contention.cu
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
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

// out[0] = MMA stream cycles, out[1] = bytes copied, out[2] = copy cycles
__global__ void __launch_bounds__(128, 1) k(long long* out, const uint8_t* src, int mma_iters, int batch, int period, int b_rows_step) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, cbar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        const uint32_t id = idesc(128, 128);
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < mma_iters; ++it) {
                // walk the B tile over 4 x 16 KB stages like a weight ring would
                const uint64_t db = desc_sw128(smem_u32(base + (size_t)(it & 3) * 16384 * b_rows_step));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(256u), "r"((uint32_t)(kk * 8)), "l"(db + 2 * kk), "r"(id), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        while (!try_wait(&bar, 0)) {}
        const long long t1 = clock64();
        const long long tt0 = __shfl_sync(0xffffffffu, t0, 0);
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - tt0;
        stop = 1;
    } else if (warp == 1 && (threadIdx.x & 31) == 0 && batch > 0) {
        uint8_t* dst = base + 65536;  // 64 KB landing zone after the B stages
        const uint8_t* s = src + (size_t)blockIdx.x * 65536;
        unsigned long long bytes = 0;
        uint32_t par = 0;
        const long long t0 = clock64();
        long long next = t0;
        while (!stop) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar)), "r"((uint32_t)batch) : "memory");
            for (int o = 0; o < batch; o += 16384)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + o)), "l"(s + o), "r"(16384u), "r"(smem_u32(&cbar)) : "memory");
            while (!try_wait(&cbar, par)) {}
            par ^= 1;
            bytes += batch;
            next += period;
            while (period > 0 && clock64() < next && !stop) {}
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[1] = (long long)bytes; out[2] = t1 - t0; }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

int main() {
    long long* d;
    uint8_t* src;
    cudaMalloc(&d, 32);
    cudaMalloc(&src, (size_t)148 * 65536);
    cudaMemset(src, 0, (size_t)148 * 65536);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int step : {0, 1}) {
        for (int cfg = 0; cfg < 5; ++cfg) {
            const int batch = (cfg == 0) ? 0 : (cfg < 4 ? 32768 : 65536);
            const int period = (cfg == 1) ? 1024 : (cfg == 2 ? 512 : 0);
            cudaMemset(d, 0, 32);
            k<<<148, 128, 160 * 1024>>>(d, src, 4000, batch, period, step);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[3];
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("B %s, bulk copies %5d B every %4d cycles: %.1f cycles / MMA (ideal 64); copy stream %.1f B/cycle/SM %s\n", step ? "walks 4 stages" : "fixed stage  ", batch, period,
                   (double)h[0] / 16000.0, h[2] ? (double)h[1] / (double)h[2] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}

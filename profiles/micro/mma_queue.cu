// Micro-benchmark: how far can the issuing thread run ahead of the tensor pipe?  Issues n
// back-to-back tcgen05.mma (M128 N128 K16, A in TMEM) and records when the issue loop returns
// and when the commit fires.  issue_time ~ max(0, n - depth) * 64 reveals the queue depth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_queue mma_queue.cu
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

__global__ void __launch_bounds__(128, 1) k(long long* out, int n) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint64_t db = desc_sw128(smem_u32(base));
        const uint32_t id = idesc(128, 128);
        long long t0 = 0, ti = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < n; ++it)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(256u), "r"((uint32_t)((it & 3) * 8)), "l"(db + 2 * (it & 3)), "r"(id), "r"(1u) : "memory");
            ti = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        const long long tt0 = __shfl_sync(0xffffffffu, t0, 0), tti = __shfl_sync(0xffffffffu, ti, 0);
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = tti - tt0; out[1] = t1 - tt0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 128}) {
        k<<<1, 128, 64 * 1024>>>(d, n);
        cudaDeviceSynchronize();
        k<<<1, 128, 64 * 1024>>>(d, n);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("n=%3d MMAs: issue loop returned after %5lld cycles, commit fired after %5lld cycles (n*64 = %d) %s\n", n, h[0], h[1], n * 64, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}

// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of
// the A-operand source (shared memory "SS" vs tensor memory "TS"), one CTA per SM, operands resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
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

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
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
        const uint32_t a_s = smem_u32(base), b_s = smem_u32(base + 16384);
        const uint64_t da = desc_sw128(a_s), db = desc_sw128(b_s);
        const uint32_t id = idesc(128, N);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (TS)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(256u), "r"((uint32_t)(kk * 8)), "l"(db + 2 * kk), "r"(id), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(256u), "l"(da + 2 * kk), "l"(db + 2 * kk), "r"(id), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
        long long tt0 = __shfl_sync(0xffffffffu, t0, 0);  // leader lane is lane 0 in practice
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - tt0;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

template <int N, bool TS>
void run(long long* d, int grid) {
    const int iters = 2000;
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<N, TS><<<grid, 128, 64 * 1024>>>(d, iters);
    cudaDeviceSynchronize();
    k<N, TS><<<grid, 128, 64 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("M=128 N=%3d K=16 %s grid=%3d: %.1f cycles / MMA (ideal %d)%s\n", N, TS ? "A in TMEM" : "A in smem", grid, (double)h / (iters * 4), N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    for (int grid : {1, 148}) {
        run<64, false>(d, grid); run<96, false>(d, grid); run<128, false>(d, grid); run<192, false>(d, grid); run<256, false>(d, grid);
        run<64, true>(d, grid); run<96, true>(d, grid); run<128, true>(d, grid); run<192, true>(d, grid); run<256, true>(d, grid);
    }
    return 0;
}

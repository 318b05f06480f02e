// Micro-benchmark: does tcgen05.ld / tcgen05.st traffic of epilogue warps slow down tcgen05.mma
// whose A operand lives in tensor memory (TS mode), and what is the TMEM load / store bandwidth
// of an SM?  One CTA per SM: warp 0 issues M128 N128 K16 MMAs back to back (A in TMEM, B in
// shared memory), E other warps stream tcgen05.ld.32x32b.x32 (or .st) on their lane quadrant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_contention tmem_contention.cu
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
__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
          "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// out[0] = cycles of the MMA stream, out[1] = x32 transfers done by all stream warps, out[2] = their cycles
__global__ void __launch_bounds__(640, 1) k(long long* out, int mma_iters, int stream_warps, int do_store, int fixed_loads, int n_mma) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    __shared__ unsigned long long n_done, cyc_max;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0; n_done = 0; cyc_max = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        if (mma_iters > 0) {
            const uint64_t db = desc_sw128(smem_u32(base));
            const uint32_t id = idesc(128, n_mma);
            long long t0 = 0;
            if (elect_one()) {
                t0 = clock64();
                for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(256u), "r"((uint32_t)(kk * 8)), "l"(db + 2 * kk), "r"(id), "r"(1u) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            uint32_t done = 0;
            while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            const long long tt0 = __shfl_sync(0xffffffffu, t0, 0);
            if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - tt0;
            stop = 1;
        }
    } else if (warp >= 4 && warp < 4 + stream_warps) {
        // the accumulator columns [384, 512) are not touched by the MMAs (D at 256, N <= 128)
        const uint32_t taddr = ((uint32_t)((warp & 3) * 32) << 16) + 384u + (uint32_t)(((warp - 4) >> 2) & 3) * 32u;
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
        unsigned long long n = 0;
        const long long t0 = clock64();
        while (mma_iters > 0 ? !stop : (n < (unsigned long long)fixed_loads)) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (do_store) st_x32(taddr, r); else ld_x32(taddr, r);
            }
            if (do_store) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            n += 4;
        }
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&n_done, n);
            atomicMax(&cyc_max, (unsigned long long)(t1 - t0));
        }
        if (r[5] == 0xdeadbeef) out[3] = r[7];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[1] = (long long)n_done; out[2] = (long long)cyc_max; }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int grid = 148;
    for (int n_mma : {128, 96}) {
        for (int st = 0; st < 2; ++st) {
            for (int sw : {0, 4, 8, 16}) {
                if (st && sw == 0) continue;
                cudaMemset(d, 0, 32);
                k<<<grid, 640, 64 * 1024>>>(d, 4000, sw, st, 0, n_mma);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[3];
                cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
                printf("MMA N=%3d TS + %2d warps streaming tcgen05.%s: %.1f cycles / MMA (ideal %d); stream %.1f B/cycle/SM %s\n", n_mma, sw, st ? "st" : "ld",
                       (double)h[0] / 16000.0, n_mma / 2, h[2] ? (double)h[1] * 4096.0 / (double)h[2] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    for (int st = 0; st < 2; ++st)
        for (int sw : {1, 4, 8, 16}) {
            cudaMemset(d, 0, 32);
            k<<<grid, 640, 64 * 1024>>>(d, 0, sw, st, 20000, 128);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[3];
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("no MMA, %2d warps streaming tcgen05.%s.32x32b.x32: %.1f B/cycle/SM %s\n", sw, st ? "st" : "ld", (double)h[1] * 4096.0 / (double)h[2],
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}

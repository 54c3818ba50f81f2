// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128) issued back-to-back by one thread,
// for kind::tf32 / kind::f16(bf16), N in {32,64,128,256}, SWIZZLE_128B K-major operands in smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int same_addr, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) ((uint32_t*)smem)[i] = 0;   // 192 KB of zeros
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        long long t0 = clock64();
        if (same_addr == 3) {
            // A in the NO-SWIZZLE K-major layout (16-byte granule planes 2608 B apart, rows 16 B apart), B SW128
            const uint64_t ad = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)(2608 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint64_t bd = make_desc(b0);
            for (int i = 0; i < iters; i += 36) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint64_t at = ad + (uint64_t)((tap / 3) * 16 + (tap % 3));
                    const uint64_t bt = bd + (uint64_t)((tap * 8192) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) mma<KIND>(tmem, at + (uint64_t)(kk * 2 * (2608 >> 4)), bt + 2 * kk, idesc, (i | tap | kk) ? 1u : 0u);
                }
            }
        } else if (same_addr == 2) {
            // cheapest possible issue loop: descriptors precomputed, +2 per k-step (32 B >> 4), 9 taps unrolled by offset
            const uint64_t ad = make_desc(a0), bd = make_desc(b0);
            for (int i = 0; i < iters; i += 36) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint64_t at = ad + (uint64_t)(((tap % 3) * 20480 + (tap / 3) * 2048) >> 4);
                    const uint64_t bt = bd + (uint64_t)(((tap % 3) * 8192) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) mma<KIND>(tmem, at + 2 * kk, bt + 2 * kk, idesc, (i | tap | kk) ? 1u : 0u);
                }
            }
        } else
        for (int i = 0; i < iters; ++i) {
            // walk through 9 "taps" x 4 k-steps of distinct smem like the conv kernel does
            const int tap = same_addr ? 0 : (i >> 2) % 9;
            const uint32_t aa = a0 + (tap % 3) * 20480 + (tap / 3) * 2048 + (i & 3) * 32;
            const uint32_t bb = b0 + (same_addr ? 0 : ((i >> 2) % 3) * (uint32_t)N * 128) + (i & 3) * 32;
            mma<KIND>(tmem, make_desc(aa), make_desc(bb), idesc, i ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 2 * 148 * sizeof(long long));
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2052;
    for (int grid : {148})
        for (int kind = 0; kind < 2; ++kind)
            for (int N : {64, 128})
                for (int same : {2, 3}) {
                    if (kind == 0) k<0><<<grid, 128, smem>>>(N, iters, same, out); else k<1><<<grid, 128, smem>>>(N, iters, same, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("grid %3d %s N=%3d %s: issue %.1f clk/mma, complete %.1f clk/mma (ideal %d)\n", grid, kind ? "bf16" : "tf32", N,
                           same == 2 ? "sw128 A  " : "noswz A  ", (double)out[0] / iters, (double)out[1] / iters, 128 * N / 256);
                }
    return 0;
}

// Microbenchmark: what slows tcgen05.mma inside the convolution kernel?  One thread issues the conv's MMA
// stream (no-swizzle A taps, SW128 resident B) on random data while other warps optionally generate the
// kernel's side traffic:  1 = cp.async 16-byte writers (paced, one 20 KB stage per `pace` clk),
// 2 = tcgen05.ld readers, 4 = readers also store to global, 8 = mbarrier pollers, 16 = zero operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_interf mma_interf.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

#define A_OFF 0
#define B_OFF (24 * 1024)
#define W_OFF (176 * 1024)
#define PLANE 2608

template <int KIND>
__global__ void __launch_bounds__(288, 1) k(int N, int iters, int flags, int pace, const uint4* gsrc, float* gdst, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, never;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    uint32_t seed = threadIdx.x * 2654435761u + blockIdx.x;
    for (int i = threadIdx.x; i < 50 * 1024; i += 288) {
        seed = seed * 1664525u + 1013904223u;
        // random bf16 pairs / tf32 values in [-1, 1)
        uint32_t v;
        if (flags & 16) v = 0;
        else if (KIND == 0) v = __float_as_uint(((int)(seed >> 8) - (1 << 23)) * (1.f / (1 << 23)));
        else v = ((0x3f00u | ((seed >> 9) & 0x7f) | ((seed >> 1) & 0x8000u)) << 16) | (0x3f00u | ((seed >> 17) & 0x7f) | ((seed >> 3) & 0x8000u));
        ((uint32_t*)smem)[i] = v;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        done = 0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
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
        const uint32_t a0 = smem_u32(smem + A_OFF), b0 = smem_u32(smem + B_OFF);
        const uint32_t wstep = (uint32_t)N * 128 >> 4;
        unsigned long long ns0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
        long long t0 = clock64();
        const uint64_t ad = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)(PLANE >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint64_t bd = make_desc(b0);
        for (int i = 0; i < iters; i += 36) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint64_t at = ad + (uint64_t)((tap / 3) * 16 + (tap % 3));
                const uint64_t bt = bd + (uint64_t)(tap * wstep);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma<KIND>(tmem, at + (uint64_t)(kk * 2 * (PLANE >> 4)), bt + 2 * kk, idesc, (tap | kk) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        unsigned long long ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        done = 1;
        if (blockIdx.x == 0) out[2 * 148] = (long long)(ns1 - ns0);
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    } else if (warp >= 1 && warp <= 4) {
        const int i = threadIdx.x - 32;
        const int col = i >> 3, plane = i & 7;
        if (flags & 1) {
            const uint32_t dst0 = smem_u32(smem + W_OFF) + plane * PLANE + col * 16;
            const uint4* src0 = gsrc + (size_t)blockIdx.x * 4096 + col * 8 + plane;
            long long t0 = clock64();
            int n = 0;
            while (!done) {
                for (int j = 0; j < 10; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(dst0 + j * 256), "l"(src0 + ((n * 10 + j) & 31) * 128) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 2;" ::: "memory");
                ++n;
                while (clock64() - t0 < (long long)n * pace) {}
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else if ((flags & 8) && lane == 0) {
            uint32_t ok = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)) : "memory");
        }
    } else if (warp >= 5) {
        const int q = warp & 3;
        if (flags & 2) {
            uint32_t r[16];
            float acc = 0.f;
            long long t0 = clock64();
            int n = 0;
            float* dst = gdst + ((size_t)blockIdx.x * 128 + q * 32 + lane) * 64;
            while (!done) {
                for (int c = 0; c < N; c += 16) {
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 256 + c, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (flags & 4) {
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(dst + (c & 63) + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    } else {
                        for (int j = 0; j < 16; ++j) acc += __uint_as_float(r[j]);
                    }
                }
                ++n;
                while (clock64() - t0 < (long long)n * pace) {}
            }
            if (acc == 123.456f) gdst[0] = acc;
        } else if ((flags & 8) && lane == 0) {
            uint32_t ok = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, (2 * 148 + 2) * sizeof(long long));
    uint4* gsrc; float* gdst;
    cudaMalloc(&gsrc, 148 * 4096 * sizeof(uint4));
    cudaMemset(gsrc, 0x3c, 148 * 4096 * sizeof(uint4));
    cudaMalloc(&gdst, 148 * 128 * 64 * sizeof(float));
    const int smem = 202 * 1024;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int iters : {36 * 200, 36 * 2000, 36 * 20000, 36 * 100000})
        for (int kind = 1; kind < 2; ++kind)
            for (int N : {64, 128})
                for (int flags : {0, 15}) {
                    const int grid = 148;
                    const int pace = 36 * (N == 64 ? 48 : 64);
                    k<1><<<grid, 288, smem>>>(N, iters, flags, pace, gsrc, gdst, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("bf16 N=%3d flags=%2d iters %8d: %.1f clk64/mma, %.2f ns/mma -> clock64 %.0f MHz, %.0f TFLOP/s chip-wide\n", N, flags, iters,
                           (double)out[1] / iters, (double)out[2 * 148] / iters, 1e3 * (double)out[1] / (double)out[2 * 148],
                           148.0 * 2 * 128 * N * 16 / ((double)out[2 * 148] / iters) / 1e3);
                    fflush(stdout);
                }
    return 0;
}

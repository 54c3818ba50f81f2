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


__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
}
#define A_OFF 0
#define B_OFF (24 * 1024)
#define PLANE 2608

// rmode 0: x16 + wait per 16 columns; 1: 4 x x16 then one wait; 2: x64 + wait.  `mma_on`: the MMA stream runs meanwhile.
template <int KIND>
__global__ void __launch_bounds__(288, 1) k(int N, int mma_on, int rmode, int nread, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    for (int i = threadIdx.x; i < 50 * 1024; i += 288) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
        if (mma_on) {
            const uint32_t fmt = KIND == 0 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
            const uint32_t a0 = smem_u32(smem + A_OFF), b0 = smem_u32(smem + B_OFF);
            const uint32_t wstep = (uint32_t)N * 128 >> 4;
            const uint64_t ad = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)(PLANE >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint64_t bd = make_desc(b0);
            long long t0 = clock64();
            int n = 0;
            while (!done) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint64_t at = ad + (uint64_t)((tap / 3) * 16 + (tap % 3));
                    const uint64_t bt = bd + (uint64_t)(tap * wstep);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) mma<KIND>(tmem, at + (uint64_t)(kk * 2 * (PLANE >> 4)), bt + 2 * kk, idesc, (tap | kk) ? 1u : 0u);
                }
                n += 36;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
            long long t2 = clock64();
            out[2] = (t2 - t0) / (n ? n : 1);
        }
    } else if (warp >= 5) {
        const int q = warp & 3;
        uint32_t r[64];
        float acc = 0.f;
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + 256;
        __nanosleep(2000);
        long long t0 = clock64();
        for (int n = 0; n < nread; ++n) {
            if (rmode == 0) {
                for (int c = 0; c < 64; c += 16) {
                    tmem_ld16(ta + c, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    for (int j = 0; j < 16; ++j) acc += __uint_as_float(r[j]);
                }
            } else if (rmode == 1) {
                tmem_ld16(ta, r); tmem_ld16(ta + 16, r + 16); tmem_ld16(ta + 32, r + 32); tmem_ld16(ta + 48, r + 48);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 64; ++j) acc += __uint_as_float(r[j]);
            } else {
                tmem_ld64(ta, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 64; ++j) acc += __uint_as_float(r[j]);
            }
        }
        long long t1 = clock64();
        if (acc == 123.456f) out[3] = 1;
        if (warp == 5 && lane == 0) out[0] = (t1 - t0) / nread;
        __syncwarp();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 5 && lane == 0) done = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 8 * sizeof(long long));
    const int smem = 202 * 1024;
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {64, 128, 256})
        for (int mma_on : {0, 1})
            for (int rmode : {0, 1, 2}) {
                out[0] = out[2] = 0;
                k<1><<<1, 288, smem>>>(N, mma_on, rmode, 200, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("bf16 N=%3d mma %s rmode %d: %lld clk per 64 columns x 4 warps, mma %lld clk each\n", N, mma_on ? "on " : "off", rmode, out[0], out[2]);
                fflush(stdout);
            }
    return 0;
}

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


// What can an epilogue warp do while the tensor core saturates the shared-memory pipe?
// smode 0: st.global.v8 (32 lanes x 32 B, 128-byte pixel stride)   1: st.global.v4 x2 (same bytes)
//       2: st.shared 32 B per lane (conflict-free)                  3: st.shared + one cp.async.bulk smem->global per warp (4 KB)
//       4: st.global.v8 with 32 lanes contiguous (1 KB contiguous per instruction)
template <int KIND>
__global__ void __launch_bounds__(288, 1) k(int N, int mma_on, int smode, int nrep, float* gdst, long long* out) {
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
        const int q = warp - 5;
        float v[8];
        for (int j = 0; j < 8; ++j) v[j] = (float)(lane + j);
        float* gbase = gdst + (size_t)blockIdx.x * (1 << 20) + q * (1 << 18);
        uint8_t* sbase = smem + 176 * 1024 + q * 4096;
        __nanosleep(2000);
        long long t0 = clock64();
        for (int n = 0; n < nrep; ++n) {
            float* g = gbase + (size_t)(n & 63) * 1024 * 4;
            if (smode == 0) {
                float* d = g + lane * 32 + (n & 3) * 8;
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
            } else if (smode == 1) {
                float4* d = reinterpret_cast<float4*>(g + lane * 32 + (n & 3) * 8);
                d[0] = make_float4(v[0], v[1], v[2], v[3]);
                d[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else if (smode == 2 || smode == 3) {
                // 32 B per lane, lanes contiguous: lane l writes [l*32, l*32+32) as two 16-byte halves, the halves
                // issued as separate conflict-free instructions (lane*32 + h*16 covers all banks per 8 lanes)
                const uint32_t sa = smem_u32(sbase) + lane * 32;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sa), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sa + 16), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
                if (smode == 3 && (n & 3) == 3) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1024;" ::"l"(g), "r"(smem_u32(sbase)) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                }
            } else if (smode == 5) {        // lane pairs write 64 contiguous bytes: 16 pixels x 64 B per instruction
                float* d = g + (lane >> 1) * 64 + (lane & 1) * 8 + (n & 1) * 16;
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
            } else if (smode == 6) {        // lane quads write one full 128-byte line: 8 pixels x 128 B per instruction
                float* d = g + (lane >> 2) * 128 + (lane & 3) * 8;
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
            } else {
                float* d = g + lane * 8;
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
            }
        }
        long long t1 = clock64();
        if (warp == 5 && lane == 0) out[0] = (t1 - t0) / nrep;
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
    float* gdst;
    cudaMalloc(&gdst, (size_t)148 * (1 << 20) * sizeof(float));
    const int smem = 202 * 1024;
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {148})
        for (int N : {64})
            for (int mma_on : {0, 1})
                for (int smode : {0, 5, 6, 4}) {
                    out[0] = out[2] = 0;
                    k<1><<<grid, 288, smem>>>(N, mma_on, smode, 400, gdst, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("grid %3d N=%3d mma %s smode %d: %lld clk per warp store step (4 warps), mma %lld clk each\n", grid, N, mma_on ? "on " : "off", smode, out[0], out[2]);
                    fflush(stdout);
                }
    return 0;
}

// PTX wrappers shared by the tcgen05 convolution kernels (mbarrier, TMA, tcgen05.mma/commit/ld, descriptors).
#pragma once
#include "common.cuh"
#include <cuda.h>

#define TC_SPIN_LIMIT (1u << 26)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  Unlike `lane == 0`, ptxas knows that code predicated on elect.sync runs in exactly one
// lane, so vector-register operands of the tensor-core instructions (TMEM address, descriptors) move to the uniform
// datapath with a plain R2UR instead of an ELECT / R2UR.BROADCAST / BRA waterfall around every UTCHMMA (measured: ~40 clk
// per MMA, which made the issuing thread, not the tensor pipe, the bottleneck of the convolution).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking poll.  Under MMA load every shared-memory round trip (mbarrier polls included) takes 250-400 clk
// because the tensor core's operand fetches saturate the pipe, so the MMA issuer polls the NEXT step's barriers
// before it issues the last MMAs of the current step and only reads the answer afterwards.
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// Four MMAs (the k-steps of one 128-byte K chunk: descriptors advance by 2 x 16 bytes, or by `a_kstep` for A) with three
// barrier polls issued BEFORE them and read AFTER them, in one asm block: a separate poll statement ends in a selp
// that stalls the in-order issuing thread for the whole shared-memory round trip (250-400 clk under MMA load) before
// the next MMA can be queued.  Returns bit i = poll i complete.
template <bool BF16>
__device__ __forceinline__ uint32_t umma_x4_poll3(uint32_t d_tmem, uint64_t a0, uint64_t a_kstep, uint64_t b0, uint32_t idesc,
                                                  uint32_t accum0, uint64_t* bar0, uint32_t par0, uint64_t* bar1, uint32_t par1,
                                                  uint64_t* bar2, uint32_t par2) {
    uint32_t ok;
    const uint64_t a1 = a0 + a_kstep, a2 = a1 + a_kstep, a3 = a2 + a_kstep;
    const uint64_t b1 = b0 + 2, b2 = b0 + 4, b3 = b0 + 6;
    if (BF16)
        asm volatile(
            "{\n\t.reg .pred p, pt, q0, q1, q2;\n\t.reg .b32 t0, t1, t2;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 q0, [%12], %13;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 q1, [%14], %15;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 q2, [%16], %17;\n\t"
            "setp.ne.b32 p, %11, 0;\n\tsetp.eq.b32 pt, %11, %11;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %6, %10, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %7, %10, pt;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%1], %4, %8, %10, pt;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%1], %5, %9, %10, pt;\n\t"
            "selp.u32 t0, 1, 0, q0;\n\tselp.u32 t1, 2, 0, q1;\n\tselp.u32 t2, 4, 0, q2;\n\t"
            "or.b32 t0, t0, t1;\n\tor.b32 %0, t0, t2;\n\t}"
            : "=r"(ok)
            : "r"(d_tmem), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accum0),
              "r"(smem_u32(bar0)), "r"(par0), "r"(smem_u32(bar1)), "r"(par1), "r"(smem_u32(bar2)), "r"(par2)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, pt, q0, q1, q2;\n\t.reg .b32 t0, t1, t2;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 q0, [%12], %13;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 q1, [%14], %15;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 q2, [%16], %17;\n\t"
            "setp.ne.b32 p, %11, 0;\n\tsetp.eq.b32 pt, %11, %11;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %2, %6, %10, p;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %7, %10, pt;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %4, %8, %10, pt;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%1], %5, %9, %10, pt;\n\t"
            "selp.u32 t0, 1, 0, q0;\n\tselp.u32 t1, 2, 0, q1;\n\tselp.u32 t2, 4, 0, q2;\n\t"
            "or.b32 t0, t0, t1;\n\tor.b32 %0, t0, t2;\n\t}"
            : "=r"(ok)
            : "r"(d_tmem), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(accum0),
              "r"(smem_u32(bar0)), "r"(par0), "r"(smem_u32(bar1)), "r"(par1), "r"(smem_u32(bar2)), "r"(par2)
            : "memory");
    return ok;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > TC_SPIN_LIMIT) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
// Relaxed wait for roles whose wake-up latency is not critical (epilogue / loader / issuer of a producer-bound kernel): a hot
// try_wait loop issues an instruction every few cycles and takes a large share of its scheduler's issue slots from the
// warps that do the work (dcn_tc: 44 % of all executed instructions were barrier polls).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int* err, int code, unsigned ns = 64) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++spins > TC_SPIN_LIMIT) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait_warp_relaxed(uint64_t* bar, uint32_t parity, int* err, int code, unsigned ns = 64) {
    if ((threadIdx.x & 31) == 0) mbar_wait_relaxed(bar, parity, err, code, ns);
    __syncwarp();
}
// Warp-collective wait: one lane polls the barrier (every poll is a shared-memory transaction that competes
// with the tensor core's operand fetches), the rest of the warp parks at __syncwarp.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int* err, int code) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, err, code);
    __syncwarp();
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// kind::f16 (bf16 / fp16 operands, K = 16 per instruction, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// 16 fp32 -> 16 bf16 (round to nearest even), stored as two 16-byte vectors
__device__ __forceinline__ void store_bf16x16(void* dst, const float* v) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
    }
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}
// Same instruction with the descriptors given as (lo, hi) words: the start-address field lives in the low
// word, so walking K / taps is one 32-bit add per operand instead of a 64-bit add chain on the uniform datapath.
__device__ __forceinline__ void umma_tf32_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
        : "memory");
}
// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// Epilogue store helper.  After tcgen05.ld each lane holds 16 consecutive output channels of ITS pixel; a
// direct store makes one instruction touch 32 different lines with 16 bytes each (measured: 5,400 clk per
// 128-pixel tile, more than the tile's MMAs).  Staging the 32 x 16 chunk through a per-warp shared buffer
// (pitch 17 floats) lets 4 lanes cover one pixel's 64 contiguous bytes: 8 lines per instruction, full sectors.
//   stg   : per-warp buffer of 32*17 floats
//   v     : this lane's 16 values;  rowptr(l) must return the destination of lane l's pixel (or nullptr)
#define EPI_PITCH 17
__device__ __forceinline__ void epi_store16(float* stg, const float* v, float* my_dst, int lane) {
#pragma unroll
    for (int j = 0; j < 16; ++j) stg[lane * EPI_PITCH + j] = v[j];
    __syncwarp();
    const int cg = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pl = i * 8 + (lane >> 2);                       // pixel (lane index within the warp) served now
        const unsigned long long dptr = __shfl_sync(0xffffffffu, (unsigned long long)my_dst, pl);
        const float* s = stg + pl * EPI_PITCH + cg * 4;
        const float4 o = make_float4(s[0], s[1], s[2], s[3]);
        if (dptr) *reinterpret_cast<float4*>(reinterpret_cast<float*>(dptr) + cg * 4) = o;
    }
    __syncwarp();
}

// 256-bit global accesses (sm_100: LDG/STG.256).  A conv epilogue lane owns one pixel, so every warp-level store
// touches 32 different lines; the LSU retires roughly one line per 3 clk, so bytes-per-line-visit is what counts.
__device__ __forceinline__ void st_global_v8(float* dst, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* src, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(src));
}
// 16 fp32 -> 16 fp16 in one 32-byte store (10-bit mantissa = TF32 precision at half the bytes; used for the
// per-pixel filter taps, the largest tensor of the forward)
__device__ __forceinline__ void store_f16x16_v8(void* dst, const float* v) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
// 16 fp32 -> 16 bf16 in one 32-byte store
__device__ __forceinline__ void store_bf16x16_v8(void* dst, const float* v) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// 16 consecutive floats from shared memory, same address in every lane (broadcast: one wavefront per instruction)
__device__ __forceinline__ void lds_bias16(uint32_t saddr, float* b) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b[4 * j]), "=f"(b[4 * j + 1]), "=f"(b[4 * j + 2]), "=f"(b[4 * j + 3]) : "r"(saddr + 16 * j));
}

// ---- shared epilogue for one 16-column chunk of one output pixel -----------------------------------------
//   v = act(acc + bias + pre) + res - res2;  y2 <- operand-typed copy;  y <- fp32 | TF32-rounded | bf16 | fp16,
//   optionally through pixel_shuffle(2) (GEMM columns ordered (i,j,c)).
struct EpiArgs {
    const float* bias; uint32_t bias_sa;      // bias_sa: shared-memory copy of bias (zeros when bias == nullptr), set by the kernel
    const float* pre; int ldpre; const float* res; int ldres; const float* res2; int ldres2;
    float* y; int ldy; float* y2; int ldy2;
    int round_out;      // 0 fp32, 1 operand-typed (TF32-rounded fp32 / bf16), 2 fp16
    int act; float slope; int ps; int wide; int c4; int H, W;
};

template <bool BF16>
__device__ __forceinline__ void epi_chunk16(const EpiArgs& e, const uint32_t* r, size_t pix, int n0, int b, int y, int x) {
    float v[16];
    lds_bias16(e.bias_sa + n0 * 4, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r[j]);
    if (e.pre) {
        float rv[16];
        ld_global_v8(e.pre + pix * e.ldpre + n0, rv);
        ld_global_v8(e.pre + pix * e.ldpre + n0 + 8, rv + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += rv[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fcvsr_act(v[j], e.act, e.slope);
    if (e.res) {
        float rv[16];
        if (e.wide) {
            ld_global_v8(e.res + pix * e.ldres + n0, rv);
            ld_global_v8(e.res + pix * e.ldres + n0 + 8, rv + 8);
        } else {
            const float4* rp = reinterpret_cast<const float4*>(e.res + pix * e.ldres + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float4 t4 = rp[j]; rv[4 * j] = t4.x; rv[4 * j + 1] = t4.y; rv[4 * j + 2] = t4.z; rv[4 * j + 3] = t4.w; }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += rv[j];
    }
    if (e.res2) {
        const float4* rp = reinterpret_cast<const float4*>(e.res2 + pix * e.ldres2 + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 t4 = rp[j]; v[4 * j] -= t4.x; v[4 * j + 1] -= t4.y; v[4 * j + 2] -= t4.z; v[4 * j + 3] -= t4.w; }
    }
    if (e.y2) {
        if (BF16) {
            if (e.wide) store_bf16x16_v8(reinterpret_cast<unsigned short*>(e.y2) + pix * e.ldy2 + n0, v);
            else store_bf16x16(reinterpret_cast<unsigned short*>(e.y2) + pix * e.ldy2 + n0, v);
        } else {
            float vr[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) vr[j] = round_tf32(v[j]);
            if (e.wide) {
                st_global_v8(e.y2 + pix * e.ldy2 + n0, vr);
                st_global_v8(e.y2 + pix * e.ldy2 + n0 + 8, vr + 8);
            } else {
                float4* d2 = reinterpret_cast<float4*>(e.y2 + pix * e.ldy2 + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j) d2[j] = make_float4(vr[4 * j], vr[4 * j + 1], vr[4 * j + 2], vr[4 * j + 3]);
            }
        }
    }
    size_t off;
    if (e.ps) {
        const int ij = n0 / e.c4, c = n0 - ij * e.c4;
        const size_t opix = ((size_t)b * 2 * e.H + 2 * y + (ij >> 1)) * (2 * (size_t)e.W) + 2 * x + (ij & 1);
        off = opix * e.ldy + c;
    } else {
        off = pix * e.ldy + n0;
    }
    if (e.round_out == 2) {
        store_f16x16_v8(reinterpret_cast<unsigned short*>(e.y) + off, v);
    } else if (BF16 && e.round_out) {
        if (e.wide) store_bf16x16_v8(reinterpret_cast<unsigned short*>(e.y) + off, v);
        else store_bf16x16(reinterpret_cast<unsigned short*>(e.y) + off, v);
    } else {
        if (e.round_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
        }
        if (e.wide) {
            st_global_v8(e.y + off, v);
            st_global_v8(e.y + off + 8, v + 8);
        } else {
            float4* dp = reinterpret_cast<float4*>(e.y + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) dp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
}

// Shared helpers for the fcvsr_b200 sm_100a kernels.
// Layout convention of the whole library: activations are NHWC fp32 ("pixel-major"): element
// (b, y, x, c) of a tensor with pixel stride `ld` (in elements, ld >= C) lives at
// base[((b*H + y)*W + x)*ld + c].  A channel slice of a wider tensor is just base+offset with the
// same ld, so torch.cat / channel splits of the reference never move data.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FCVSR_OK 0
#define FCVSR_ERR_ARG (-1)
#define FCVSR_ERR_CUDA (-2)
#define FCVSR_ERR_UNSUPPORTED (-3)

// activation codes shared by the conv epilogues (host mirror: fcvsr_b200/_capi.py)
#define FCVSR_ACT_NONE 0
#define FCVSR_ACT_RELU 1
#define FCVSR_ACT_LEAKY 2   // negative slope passed by value
#define FCVSR_ACT_PRELU 3   // negative slope read from a device scalar (nn.PReLU weight)

static inline int fcvsr_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FCVSR_OK : FCVSR_ERR_CUDA;
}

__device__ __forceinline__ float fcvsr_act(float v, int act, float slope) {
    if (act == FCVSR_ACT_NONE) return v;
    if (act == FCVSR_ACT_RELU) return fmaxf(v, 0.f);
    return v >= 0.f ? v : v * slope;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Round an fp32 value to the nearest TF32 (10-bit mantissa, ties away), kept in fp32 storage.  The
// tensor cores truncate kind::tf32 operands, so tensors that feed the tcgen05 convolutions are stored
// pre-rounded: truncation of an already-rounded value is exact and the truncation bias disappears.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// Operand-typed store of 4 consecutive channels: TF32-rounded fp32 (op16 = 0) or bf16 (op16 = 1).
// `base` is the tensor base in its own element type, `idx` the element index.
__device__ __forceinline__ void store_operand4(void* base, size_t idx, float4 v, int op16) {
    if (op16) {
        uint32_t lo, hi;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
        *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base) + idx) = make_uint2(lo, hi);
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) =
            make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    }
}
__device__ __forceinline__ void store_operand2(void* base, size_t idx, float2 v, int op16) {
    if (op16) {
        uint32_t w;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(v.y), "f"(v.x));
        *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned short*>(base) + idx) = w;
    } else {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx) = make_float2(round_tf32(v.x), round_tf32(v.y));
    }
}
__device__ __forceinline__ void store_operand1(void* base, size_t idx, float v, int op16) {
    if (op16) {
        unsigned short h;
        asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(v));
        reinterpret_cast<unsigned short*>(base)[idx] = h;
    } else {
        reinterpret_cast<float*>(base)[idx] = round_tf32(v);
    }
}

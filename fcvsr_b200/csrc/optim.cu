// Adam step of the reference's training loop (CVSR_train/train_LD_freqCVSR_22.py:204,251: torch.optim.Adam(lr = 5e-6,
// weight_decay = 1e-5), defaults betas = (0.9, 0.999), eps = 1e-8): L2 weight decay folded into the gradient, bias-corrected
// moments, fp32 state.  Multi-tensor: one launch updates up to 64 parameter tensors (their pointers travel in the kernel
// parameter space), so the 678 tensors of FCVSR take 11 launches instead of 678 x 4 elementwise ATen kernels.
//   g' = g + wd * p;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: 16 B read + 12 B written per element.
#include <math.h>

#include "common.cuh"

#define ADAM_MAX_TENSORS 64
#define ADAM_CHUNK 4096                 // elements per block (256 threads x 4 x float4)

struct AdamBatch {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
};

struct AdamHyper { float b1, omb1, b2, omb2, eps, wd, step_size, inv_sqrt_bc2; };   // 1 - beta formed in double on the host

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h) {
    g = fmaf(h.wd, p, g);
    m = fmaf(h.omb1, g - m, m);                  // exp_avg.lerp_(grad, 1 - beta1), as torch's single-tensor Adam
    v = fmaf(h.b2, v, h.omb2 * g * g);
    p -= h.step_size * m / (sqrtf(v) * h.inv_sqrt_bc2 + h.eps);
}

__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ AdamBatch t, const AdamHyper h) {
    const int ti = blockIdx.y;
    const long long n = t.n[ti];
    const long long base = (long long)blockIdx.x * ADAM_CHUNK;
    if (base >= n) return;
    float* p = t.p[ti];
    const float* g = t.g[ti];
    float* m = t.m[ti];
    float* v = t.v[ti];
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    const long long end = base + ADAM_CHUNK < n ? base + ADAM_CHUNK : n;
    if (vec && end - base == ADAM_CHUNK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = base + (long long)(u * 256 + threadIdx.x) * 4;
            float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            adam_update(pp.x, gg.x, mm.x, vv.x, h);
            adam_update(pp.y, gg.y, mm.y, vv.y, h);
            adam_update(pp.z, gg.z, mm.z, vv.z, h);
            adam_update(pp.w, gg.w, mm.w, vv.w, h);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        }
    } else {
        for (long long i = base + threadIdx.x; i < end; i += 256) {
            float pp = p[i], mm = m[i], vv = v[i];
            adam_update(pp, g[i], mm, vv, h);
            p[i] = pp; m[i] = mm; v[i] = vv;
        }
    }
}

// params / grads / exp_avg / exp_avg_sq: HOST arrays of `count` device pointers; numels: HOST array of element counts.
// step: 1-based step number shared by all tensors (the reference steps every parameter every iteration).
extern "C" int fcvsr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                               float* const* exp_avg_sq, const long long* numels, int count, double lr, double beta1,
                               double beta2, double eps, double weight_decay, int step, cudaStream_t st) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numels || count < 0 || step < 1) return FCVSR_ERR_ARG;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamHyper h;
    h.b1 = (float)beta1; h.omb1 = (float)(1.0 - beta1); h.b2 = (float)beta2; h.omb2 = (float)(1.0 - beta2);
    h.eps = (float)eps; h.wd = (float)weight_decay; h.step_size = (float)(lr / bc1); h.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    for (int i0 = 0; i0 < count; i0 += ADAM_MAX_TENSORS) {
        AdamBatch t;
        const int nb = count - i0 < ADAM_MAX_TENSORS ? count - i0 : ADAM_MAX_TENSORS;
        long long nmax = 0;
        for (int i = 0; i < ADAM_MAX_TENSORS; ++i) {
            const int k = i < nb ? i0 + i : i0;
            if (i < nb && (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || numels[k] < 0)) return FCVSR_ERR_ARG;
            t.p[i] = params[k]; t.g[i] = grads[k]; t.m[i] = exp_avg[k]; t.v[i] = exp_avg_sq[k];
            t.n[i] = i < nb ? numels[k] : 0;
            if (t.n[i] > nmax) nmax = t.n[i];
        }
        if (nmax == 0) continue;
        dim3 grid((unsigned)((nmax + ADAM_CHUNK - 1) / ADAM_CHUNK), nb);
        adam_step_kernel<<<grid, 256, 0, st>>>(t, h);
    }
    return fcvsr_launch_status();
}

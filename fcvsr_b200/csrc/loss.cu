// Charbonnier loss forward and backward (CVSR_train/opt/loss.py:20-31): sum over all elements of sqrt((x - y)^2 + eps), eps = 1e-4,
// with the optional mean_res variant (the per-sample mean of the difference first, :27-29).
// Two deterministic stages: a fixed grid of blocks accumulates grid-strided partial sums in double precision (the
// reference's fp32 torch.sum uses pairwise summation; a double accumulator is at least as accurate and order-independent
// results would need atomics), the last stage sums the block partials in index order.  HBM-bound: 8 bytes per element.
#include "common.cuh"

#define CH_BLOCKS 592          // 4 x 148 SMs
#define CH_THREADS 256

__global__ void __launch_bounds__(CH_THREADS) charbonnier_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                       size_t n, float eps, double* __restrict__ partial) {
    __shared__ double red[CH_THREADS / 32];
    double acc = 0.0;
    const size_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (size_t i = (size_t)blockIdx.x * CH_THREADS + threadIdx.x; i < n4; i += (size_t)CH_BLOCKS * CH_THREADS) {
        const float4 a = x4[i], b = y4[i];
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        acc += (double)(sqrtf(d0 * d0 + eps) + sqrtf(d1 * d1 + eps)) + (double)(sqrtf(d2 * d2 + eps) + sqrtf(d3 * d3 + eps));
    }
    if (blockIdx.x == 0)
        for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += CH_THREADS) {
            const float d = x[i] - y[i];
            acc += (double)sqrtf(d * d + eps);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < CH_THREADS / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void charbonnier_final_kernel(const double* __restrict__ partial, float* __restrict__ out) {
    // one warp, fixed order: lane l sums partial[l], partial[l + 32], ... and the lanes are combined by a shuffle tree
    double s = 0.0;
    for (int i = threadIdx.x; i < CH_BLOCKS; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = (float)s;
}

// mean_res: one block per sample reduces mean(x - y), then the loss is sum_b sqrt(mean_b^2 + eps)
__global__ void __launch_bounds__(CH_THREADS) charbonnier_meanres_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                       size_t per_sample, double* __restrict__ partial) {
    __shared__ double red[CH_THREADS / 32];
    const float* xb = x + (size_t)blockIdx.x * per_sample;
    const float* yb = y + (size_t)blockIdx.x * per_sample;
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < per_sample; i += CH_THREADS) acc += (double)(xb[i] - yb[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < CH_THREADS / 32; ++w) s += red[w];
        partial[blockIdx.x] = s / (double)per_sample;
    }
}

__global__ void charbonnier_meanres_final_kernel(const double* __restrict__ partial, int B, float eps, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < B; ++i) {
            const float m = (float)partial[i];
            s += (double)sqrtf(m * m + eps);
        }
        out[0] = (float)s;
    }
}

// scratch: at least max(CH_BLOCKS, B) doubles.  out: one float on the device.
extern "C" int fcvsr_charbonnier_loss(const float* x, const float* y, long long numel, int batch, int mean_res, float eps,
                                      double* scratch, float* out, cudaStream_t st) {
    if (!x || !y || !scratch || !out || numel <= 0 || batch <= 0 || numel % batch) return FCVSR_ERR_ARG;
    if (((uintptr_t)x | (uintptr_t)y) & 15) return FCVSR_ERR_ARG;
    if (mean_res) {
        charbonnier_meanres_kernel<<<batch, CH_THREADS, 0, st>>>(x, y, (size_t)(numel / batch), scratch);
        charbonnier_meanres_final_kernel<<<1, 32, 0, st>>>(scratch, batch, eps, out);
    } else {
        charbonnier_partial_kernel<<<CH_BLOCKS, CH_THREADS, 0, st>>>(x, y, (size_t)numel, eps, scratch);
        charbonnier_final_kernel<<<1, 32, 0, st>>>(scratch, out);
    }
    return fcvsr_launch_status();
}

// ---- backward: d loss / d x = g * d / sqrt(d^2 + eps) (d = x - y), d loss / d y = -that; with mean_res the per-sample mean
// m_b (kept in `scratch` by the forward call) gives g * m_b / sqrt(m_b^2 + eps) / per_sample for every element of sample b.
// `grad_out` is the upstream gradient of the scalar loss, read from the device (no host synchronisation).
__global__ void __launch_bounds__(CH_THREADS) charbonnier_backward_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                        size_t n, size_t per_sample, int mean_res, float eps,
                                                                        const float* __restrict__ grad_out,
                                                                        const double* __restrict__ means, float* __restrict__ gx,
                                                                        float* __restrict__ gy) {
    const float g = grad_out[0];
    for (size_t i = (size_t)blockIdx.x * CH_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * CH_THREADS) {
        float v;
        if (mean_res) {
            const float m = (float)means[i / per_sample];
            v = g * m / sqrtf(m * m + eps) / (float)per_sample;
        } else {
            const float d = x[i] - y[i];
            v = g * d / sqrtf(d * d + eps);
        }
        if (gx) gx[i] = v;
        if (gy) gy[i] = -v;
    }
}

// scratch: the buffer the forward call filled (only read when mean_res); grad_x / grad_y may each be NULL.
extern "C" int fcvsr_charbonnier_loss_backward(const float* x, const float* y, long long numel, int batch, int mean_res, float eps,
                                               const float* grad_out, const double* scratch, float* grad_x, float* grad_y,
                                               cudaStream_t st) {
    if (!x || !y || !grad_out || numel <= 0 || batch <= 0 || numel % batch || (mean_res && !scratch)) return FCVSR_ERR_ARG;
    if (!grad_x && !grad_y) return FCVSR_OK;
    const long long want = (numel + CH_THREADS - 1) / CH_THREADS;
    const int blocks = (int)(want < 8 * CH_BLOCKS ? want : 8 * CH_BLOCKS);
    charbonnier_backward_kernel<<<blocks, CH_THREADS, 0, st>>>(x, y, (size_t)numel, (size_t)(numel / batch), mean_res, eps, grad_out,
                                                               scratch, grad_x, grad_y);
    return fcvsr_launch_status();
}

// ---- mmedit pixel losses (mmedit_train/mmedit/models/losses/pixelwise_loss.py: l1_loss :13-24, mse_loss :27-38,
// charbonnier_loss :41-51 with eps = 1e-12; reduction 'mean' | 'sum' and loss_weight folded into `scale`): the REDS
// configuration of FCVSR trains with MSELoss(mean) (configs/restorers/fcvsr/fcvsr_redsLD_QP22.py:7).
//   kind 0: sqrt(d^2 + eps)    kind 1: d^2    kind 2: |d|          out = scale * sum
template <int KIND>
__device__ __forceinline__ float pl_value(float d, float eps) {
    return KIND == 0 ? sqrtf(d * d + eps) : (KIND == 1 ? d * d : fabsf(d));
}
template <int KIND>
__device__ __forceinline__ float pl_deriv(float d, float eps) {
    return KIND == 0 ? d / sqrtf(d * d + eps) : (KIND == 1 ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)));
}

template <int KIND>
__global__ void __launch_bounds__(CH_THREADS) pixel_loss_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                      size_t n, float eps, double* __restrict__ partial) {
    __shared__ double red[CH_THREADS / 32];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * CH_THREADS + threadIdx.x; i < n; i += (size_t)CH_BLOCKS * CH_THREADS)
        acc += (double)pl_value<KIND>(x[i] - y[i], eps);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < CH_THREADS / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void pixel_loss_final_kernel(const double* __restrict__ partial, double scale, float* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < CH_BLOCKS; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = (float)(s * scale);
}
template <int KIND>
__global__ void __launch_bounds__(CH_THREADS) pixel_loss_backward_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                       size_t n, float eps, float scale,
                                                                       const float* __restrict__ grad_out, float* __restrict__ gx,
                                                                       float* __restrict__ gy) {
    const float g = grad_out[0] * scale;
    for (size_t i = (size_t)blockIdx.x * CH_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * CH_THREADS) {
        const float v = g * pl_deriv<KIND>(x[i] - y[i], eps);
        if (gx) gx[i] = v;
        if (gy) gy[i] = -v;
    }
}

// scratch: CH_BLOCKS (592) doubles; out: one float on the device = scale * sum_i f(x_i - y_i)
extern "C" int fcvsr_pixel_loss(const float* x, const float* y, long long numel, int kind, float eps, double scale, double* scratch,
                                float* out, cudaStream_t st) {
    if (!x || !y || !scratch || !out || numel <= 0 || kind < 0 || kind > 2) return FCVSR_ERR_ARG;
    if (kind == 0) pixel_loss_partial_kernel<0><<<CH_BLOCKS, CH_THREADS, 0, st>>>(x, y, (size_t)numel, eps, scratch);
    else if (kind == 1) pixel_loss_partial_kernel<1><<<CH_BLOCKS, CH_THREADS, 0, st>>>(x, y, (size_t)numel, eps, scratch);
    else pixel_loss_partial_kernel<2><<<CH_BLOCKS, CH_THREADS, 0, st>>>(x, y, (size_t)numel, eps, scratch);
    pixel_loss_final_kernel<<<1, 32, 0, st>>>(scratch, scale, out);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_pixel_loss_backward(const float* x, const float* y, long long numel, int kind, float eps, double scale,
                                         const float* grad_out, float* grad_x, float* grad_y, cudaStream_t st) {
    if (!x || !y || !grad_out || numel <= 0 || kind < 0 || kind > 2) return FCVSR_ERR_ARG;
    if (!grad_x && !grad_y) return FCVSR_OK;
    const long long want = (numel + CH_THREADS - 1) / CH_THREADS;
    const int blocks = (int)(want < 8 * CH_BLOCKS ? want : 8 * CH_BLOCKS);
    const size_t n = (size_t)numel;
    if (kind == 0) pixel_loss_backward_kernel<0><<<blocks, CH_THREADS, 0, st>>>(x, y, n, eps, (float)scale, grad_out, grad_x, grad_y);
    else if (kind == 1) pixel_loss_backward_kernel<1><<<blocks, CH_THREADS, 0, st>>>(x, y, n, eps, (float)scale, grad_out, grad_x, grad_y);
    else pixel_loss_backward_kernel<2><<<blocks, CH_THREADS, 0, st>>>(x, y, n, eps, (float)scale, grad_out, grad_x, grad_y);
    return fcvsr_launch_status();
}

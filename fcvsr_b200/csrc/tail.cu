// Up-sampling tail helpers (GShiftNet.forward, CVSR_freq.py:2739-2751).
//   pixel_shuffle_nhwc   F.pixel_shuffle(x, 2) between channel slices of NHWC buffers (:2740-2741)
//   bilinear_up4         F.interpolate(center LR frame, x4, 'bilinear', align_corners=False) (:2750)
// (The pixel shuffles that follow a convolution are fused into that convolution's epilogue.)
#include "common.cuh"

// in [B,H,W,ldi] with 4*Co channels (channel co*4 + i*2 + j) -> out [B,2H,2W,ldo] with Co channels
__global__ void pixel_shuffle_kernel(const void* __restrict__ in, int ldi, void* __restrict__ out, int ldo, int H, int W,
                                     int Co, size_t total, int half) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ci = (int)(idx % (4 * Co));
    const size_t pix = idx / (4 * Co);
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
    const int co = ci >> 2, i = (ci >> 1) & 1, j = ci & 1;
    const size_t o = (((size_t)b * 2 * H + 2 * y + i) * (2 * W) + 2 * x + j) * ldo + co;
    if (half) reinterpret_cast<unsigned short*>(out)[o] = reinterpret_cast<const unsigned short*>(in)[pix * ldi + ci];
    else reinterpret_cast<float*>(out)[o] = reinterpret_cast<const float*>(in)[pix * ldi + ci];
}

extern "C" int fcvsr_pixel_shuffle(const void* in, int ldi, void* out, int ldo, int B, int H, int W, int Co,
                                   int half, cudaStream_t st) {
    if (!in || !out || Co <= 0) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * H * W * 4 * Co;
    pixel_shuffle_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, ldi, out, ldo, H, W, Co, total, half);
    return fcvsr_launch_status();
}

// in: plane [B][H][W] with batch stride `bstride` (elements); out [B,4H,4W] contiguous
__global__ void bilinear_up4_kernel(const float* __restrict__ in, size_t bstride, float* __restrict__ out, int H, int W,
                                    size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Wo = 4 * W, Ho = 4 * H;
    const int x = (int)(idx % Wo), y = (int)((idx / Wo) % Ho), b = (int)(idx / ((size_t)Wo * Ho));
    const float sy = fmaxf(0.25f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.25f * (x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float* p = in + (size_t)b * bstride;
    const float v = (1.f - ly) * ((1.f - lx) * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1]) +
                    ly * ((1.f - lx) * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1]);
    out[idx] = v;
}

extern "C" int fcvsr_bilinear_up4(const float* in, long long bstride, float* out, int B, int H, int W, cudaStream_t st) {
    if (!in || !out) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * 16 * H * W;
    bilinear_up4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, (size_t)bstride, out, H, W, total);
    return fcvsr_launch_status();
}

// zero a channel range of an NHWC buffer (padding channels of the 80->96 / 84->96 concat buffers)
__global__ void fill_channels_kernel(float* __restrict__ x, int ld, int c0, int nc, float v, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    x[(idx / nc) * ld + c0 + (idx % nc)] = v;
}

extern "C" int fcvsr_fill_channels(float* x, int ld, int c0, int nc, float v, long long npix, cudaStream_t st) {
    if (!x || nc <= 0) return FCVSR_ERR_ARG;
    const size_t total = (size_t)npix * nc;
    fill_channels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ld, c0, nc, v, total);
    return fcvsr_launch_status();
}

// NCHW clip [B,T,H,W] (T = 7 frames, C = 1) -> NHWC [B,H,W,32] with channels >= T zeroed, TF32-rounded:
// the tensor-core operand of feat_extract (CVSR_freq.py:2663), whose Cin = 7 is padded to one 32-channel chunk.
__global__ void pack_clip_kernel(const float* __restrict__ x, void* __restrict__ y, int T, int P, int cpad, size_t total,
                                 int op16) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cpad);
    const size_t pix = idx / cpad;
    const size_t b = pix / P, p = pix - b * P;
    store_operand1(y, idx, c < T ? x[(b * T + c) * P + p] : 0.f, op16);
}

extern "C" int fcvsr_pack_clip(const float* x, void* y, int B, int T, int H, int W, int op16, cudaStream_t st) {
    if (!x || !y || T < 1 || T > 32) return FCVSR_ERR_ARG;
    const int cpad = op16 ? 64 : 32;                  // one K chunk of the tensor-core conv
    const size_t total = (size_t)B * H * W * cpad;
    pack_clip_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, y, T, H * W, cpad, total, op16);
    return fcvsr_launch_status();
}

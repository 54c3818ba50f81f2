// Deformable convolution (DCNv1 / DCNv2 "modulated") forward, NCHW fp32, fused gather + GEMM.
//
// Replaces the reference operator's im2col kernel + cuBLAS addmm pair
//   modulated_deformable_im2col_gpu_kernel   ops/dcn/src/deform_conv_cuda_kernel.cu:570-632
//   deformable_im2col_gpu_kernel             ops/dcn/src/deform_conv_cuda_kernel.cu:190-242
//   output[b][g].addmm_(weight[g], columns[g]) + bias   ops/dcn/src/deform_conv_cuda.cpp:229-234, :545-563
// with one kernel: the bilinear-sampled, mask-modulated column tile is produced directly in shared
// memory (it never exists in HBM) and consumed by a register-tiled GEMM against the weight tile.
// Sampling semantics follow dmcn_im2col_bilinear (.cu:84-114): a tap contributes only when
// -1 < h < H and -1 < w < W, and corners outside the image contribute zero.
#include "common.cuh"

#define DC_TP 64
#define DC_TN 64
#define DC_TK 16

struct DcnArgs {
    const float* x; const float* w; const float* bias; const float* offset; const float* mask; float* y;
    int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, Ho, Wo;
    long long off_bs, mask_bs;   // batch strides (elements) of offset / mask
    int mask_sigmoid;
};

__device__ __forceinline__ float dcn_sample(const float* __restrict__ img, int H, int W, float h, float w) {
    if (!(h > -1.f && w > -1.f && h < (float)H && w < (float)W)) return 0.f;
    const float fh = floorf(h), fw = floorf(w);
    const int h0 = (int)fh, w0 = (int)fw, h1 = h0 + 1, w1 = w0 + 1;
    const float lh = h - fh, lw = w - fw, hh = 1.f - lh, hw = 1.f - lw;
    float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
    if (h0 >= 0 && w0 >= 0) v1 = img[h0 * W + w0];
    if (h0 >= 0 && w1 <= W - 1) v2 = img[h0 * W + w1];
    if (h1 <= H - 1 && w0 >= 0) v3 = img[h1 * W + w0];
    if (h1 <= H - 1 && w1 <= W - 1) v4 = img[h1 * W + w1];
    return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

__global__ void __launch_bounds__(256) dcn_forward_kernel(DcnArgs a) {
    __shared__ float As[DC_TK][DC_TP + 4];
    __shared__ float Ws[DC_TK][DC_TN + 4];
    const int tid = threadIdx.x;
    const int P = a.Ho * a.Wo;
    const int kk2 = a.kh * a.kw;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int K = cin_g * kk2;
    const int ntile_n = (cout_g + DC_TN - 1) / DC_TN;
    const int grp = blockIdx.y / ntile_n, n0 = (blockIdx.y % ntile_n) * DC_TN;
    const int p0 = blockIdx.x * DC_TP;
    const int b = blockIdx.z;
    const int ch_per_dg = a.Cin / a.dg;
    const int pg = tid >> 4, ng = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lp = tid & 63, lk = (tid >> 6) * 4;      // sampler mapping: pixel, 4 consecutive k
    const int p = p0 + lp;
    const int ho = p / a.Wo, wo = p - ho * a.Wo;
    const int wr = tid >> 4, wc = (tid & 15) * 4;
    for (int k0 = 0; k0 < K; k0 += DC_TK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + lk + u;
            float v = 0.f;
            if (p < P && k < K) {
                const int c = grp * cin_g + k / kk2, tap = k % kk2;
                const int i = tap / a.kw, j = tap - i * a.kw;
                const int g = c / ch_per_dg;
                const size_t obase = (size_t)b * a.off_bs + ((size_t)g * 2 * kk2 + 2 * tap) * P + p;
                const float oh = a.offset[obase], ow = a.offset[obase + P];
                const float hs = (float)(ho * a.sh - a.ph + i * a.dh) + oh;
                const float ws_ = (float)(wo * a.sw - a.pw + j * a.dw) + ow;
                v = dcn_sample(a.x + ((size_t)b * a.Cin + c) * a.H * a.W, a.H, a.W, hs, ws_);
                if (a.mask) {
                    float mk = a.mask[(size_t)b * a.mask_bs + ((size_t)g * kk2 + tap) * P + p];
                    if (a.mask_sigmoid) mk = 1.f / (1.f + __expf(-mk));
                    v *= mk;
                }
            }
            As[lk + u][lp] = v;
        }
        {
            const int k = k0 + wr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n = n0 + wc + u;
                Ws[wr][wc + u] = (k < K && n < cout_g) ? a.w[((size_t)(grp * cout_g + n)) * K + k] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DC_TK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][pg * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Ws[kk][ng * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pp = p0 + pg * 4 + i;
        if (pp >= P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + ng * 4 + j;
            if (n >= cout_g) continue;
            const int co = grp * cout_g + n;
            a.y[((size_t)b * a.Cout + co) * P + pp] = acc[i][j] + (a.bias ? a.bias[co] : 0.f);
        }
    }
}

extern "C" int fcvsr_modulated_deform_conv_forward(const float* input, const float* weight, const float* bias,
                                                   const float* offset, const float* mask, float* output, int B,
                                                   int Cin, int H, int W, int Cout, int kh, int kw, int stride_h,
                                                   int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                                   int deformable_groups, long long offset_batch_stride,
                                                   long long mask_batch_stride, int mask_sigmoid, cudaStream_t st) {
    if (!input || !weight || !offset || !output) return FCVSR_ERR_ARG;
    if (B <= 0 || groups <= 0 || deformable_groups <= 0 || Cin % groups || Cout % groups || Cin % deformable_groups)
        return FCVSR_ERR_ARG;
    DcnArgs a;
    a.x = input; a.w = weight; a.bias = bias; a.offset = offset; a.mask = mask; a.y = output;
    a.B = B; a.Cin = Cin; a.H = H; a.W = W; a.Cout = Cout; a.kh = kh; a.kw = kw; a.sh = stride_h; a.sw = stride_w;
    a.ph = pad_h; a.pw = pad_w; a.dh = dil_h; a.dw = dil_w; a.groups = groups; a.dg = deformable_groups;
    a.Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
    a.Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
    if (a.Ho <= 0 || a.Wo <= 0) return FCVSR_ERR_ARG;
    a.off_bs = offset_batch_stride > 0 ? offset_batch_stride : (long long)deformable_groups * 2 * kh * kw * a.Ho * a.Wo;
    a.mask_bs = mask_batch_stride > 0 ? mask_batch_stride : (long long)deformable_groups * kh * kw * a.Ho * a.Wo;
    a.mask_sigmoid = mask_sigmoid;
    const int cout_g = Cout / groups;
    dim3 grid((a.Ho * a.Wo + DC_TP - 1) / DC_TP, groups * ((cout_g + DC_TN - 1) / DC_TN), B);
    dcn_forward_kernel<<<grid, 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

extern "C" const char* fcvsr_version(void) { return "fcvsr_b200 0.1 sm_100a"; }

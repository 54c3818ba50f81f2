// Generic NHWC fp32 convolution on the CUDA cores (register-tiled FFMA).
//
// This is the any-shape kernel of the library: it covers the convolutions whose shapes do not fit
// the tcgen05 implicit-GEMM kernel (conv_tc.cu) -- Cin = 7 (feat_extract, CVSR_freq.py:2663),
// stride 2 (rconcat1/2 :2671-2672), Cout in {1, 4} (conv_last0 :2684, the per-bin MLP heads
// :1384,1395), Cin = 209 (convcorr[0] :1380) -- and serves as the fp32 cross-check of the
// tensor-core kernel in the GPU tests.  Same fused epilogue as conv_tc.cu:
//     v = act(acc + bias);  v += res;  v -= res2;  store (optionally through a pixel-shuffle(2)).
// Weights are pre-packed by the host as [kh*kw][Cin][Cout].
#include "common.cuh"

#define CD_TP 64   // pixels per tile (8x8)
#define CD_TN 64   // output channels per tile
#define CD_TK 16   // input channels per step

struct ConvDirectArgs {
    const float* x; int ldx; int x_nchw;
    const float* w; const float* bias;
    const float* res; int ldres;
    const float* res2; int ldres2;
    float* y; int ldy;
    int B, H, W, Cin, Cout, ks, stride, Ho, Wo;
    int act; float slope; const float* slope_ptr;
    int ps; int y_nchw;
    void* y2; int ldy2; int round_out; int op16;
    int transposed;      // data-gradient mode: x is dy [B,H,W] of a stride-`stride` convolution, the tile domain (Ho, Wo) is dx
};

__global__ void __launch_bounds__(256) conv_direct_kernel(ConvDirectArgs a) {
    __shared__ float As[CD_TK][CD_TP + 4];
    __shared__ float Ws[CD_TK][CD_TN + 4];
    const int tid = threadIdx.x;
    const int tiles_x = (a.Wo + 7) / 8;
    const int ty0 = (blockIdx.x / tiles_x) * 8, tx0 = (blockIdx.x % tiles_x) * 8;
    const int n0 = blockIdx.y * CD_TN;
    const int b = blockIdx.z;
    const int pad = a.ks / 2;
    const int pg = tid >> 4, ng = tid & 15;       // pixel group (4 px), cout group (4 ch)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // A loader mapping: pixel lp = tid/4, channel sub-block (tid%4)*4
    const int lp = tid >> 2, lc = (tid & 3) * 4;
    const int loy = ty0 + (lp >> 3), lox = tx0 + (lp & 7);
    // W loader mapping: row tid/16, 4 couts at (tid%16)*4
    const int wr = tid >> 4, wc = (tid & 15) * 4;

    const int ntaps = a.ks * a.ks;
    for (int tap = 0; tap < ntaps; ++tap) {
        const int ky = tap / a.ks, kx = tap - ky * a.ks;
        int iy = loy * a.stride + ky - pad, ix = lox * a.stride + kx - pad;
        bool inb = (loy < a.Ho) && (lox < a.Wo);
        if (a.transposed) {
            // dx[loy, lox] += dy[(loy + pad - ky) / s, (lox + pad - kx) / s] * w[tap] where the divisions are exact
            const int ty = loy + pad - ky, tx = lox + pad - kx;
            inb = inb && ty >= 0 && tx >= 0 && (ty % a.stride) == 0 && (tx % a.stride) == 0;
            iy = ty / a.stride; ix = tx / a.stride;
        }
        inb = inb && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
        for (int c0 = 0; c0 < a.Cin; c0 += CD_TK) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + lc + u;
                float v = 0.f;
                if (inb && c < a.Cin) {
                    v = a.x_nchw ? a.x[(((size_t)b * a.Cin + c) * a.H + iy) * a.W + ix]
                                 : a.x[(((size_t)b * a.H + iy) * a.W + ix) * a.ldx + c];
                }
                As[lc + u][lp] = v;
            }
            {
                const int c = c0 + wr;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n0 + wc + u;
                    Ws[wr][wc + u] = (c < a.Cin && n < a.Cout) ? a.w[((size_t)tap * a.Cin + c) * a.Cout + n] : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < CD_TK; ++kk) {
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
    }
    const float slope = a.act == FCVSR_ACT_PRELU ? a.slope_ptr[0] : a.slope;
    const int c4 = a.Cout >> 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = pg * 4 + i;
        const int oy = ty0 + (p >> 3), ox = tx0 + (p & 7);
        if (oy >= a.Ho || ox >= a.Wo) continue;
        const size_t pix = ((size_t)b * a.Ho + oy) * a.Wo + ox;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + ng * 4 + j;
            if (n >= a.Cout) continue;
            float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
            v = fcvsr_act(v, a.act, slope);
            if (a.res) v += a.res[pix * a.ldres + n];
            if (a.res2) v -= a.res2[pix * a.ldres2 + n];
            if (a.y2) store_operand1(a.y2, pix * a.ldy2 + n, v, a.op16);
            if (a.round_out) v = round_tf32(v);
            if (a.ps) {
                const int ij = n / c4, c = n - ij * c4;
                const size_t opix = ((size_t)b * 2 * a.Ho + 2 * oy + (ij >> 1)) * (2 * a.Wo) + 2 * ox + (ij & 1);
                a.y[opix * a.ldy + c] = v;
            } else if (a.y_nchw) {
                a.y[(((size_t)b * a.Cout + n) * a.Ho + oy) * a.Wo + ox] = v;
            } else {
                a.y[pix * a.ldy + n] = v;
            }
        }
    }
}

extern "C" int fcvsr_conv2d_direct(const float* x, int ldx, int x_nchw, const float* w, const float* bias,
                                   const float* res, int ldres, const float* res2, int ldres2, float* y, int ldy,
                                   int B, int H, int W, int Cin, int Cout, int ksize, int stride, int act, float slope,
                                   const float* slope_ptr, int pixel_shuffle, int y_nchw, void* y2, int ldy2,
                                   int round_out, int op16, cudaStream_t st) {
    if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || !(ksize & 1) || stride < 1)
        return FCVSR_ERR_ARG;
    if (act == FCVSR_ACT_PRELU && !slope_ptr) return FCVSR_ERR_ARG;
    if (pixel_shuffle && (Cout & 3)) return FCVSR_ERR_ARG;
    ConvDirectArgs a;
    a.x = x; a.ldx = ldx; a.x_nchw = x_nchw; a.w = w; a.bias = bias;
    a.res = res; a.ldres = ldres; a.res2 = res2; a.ldres2 = ldres2; a.y = y; a.ldy = ldy;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ks = ksize; a.stride = stride;
    const int pad = ksize / 2;
    a.Ho = (H + 2 * pad - ksize) / stride + 1;
    a.Wo = (W + 2 * pad - ksize) / stride + 1;
    a.act = act; a.slope = slope; a.slope_ptr = slope_ptr; a.ps = pixel_shuffle; a.y_nchw = y_nchw; a.y2 = y2; a.ldy2 = ldy2; a.round_out = round_out; a.op16 = op16;
    a.transposed = 0;
    dim3 grid(((a.Ho + 7) / 8) * ((a.Wo + 7) / 8), (Cout + CD_TN - 1) / CD_TN, B);
    conv_direct_kernel<<<grid, 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

// Data gradient of y = conv(x, w) (k x k, stride s, padding k/2) for any shape, on the CUDA cores:
//     dx[b, iy, ix, ci] = sum_{ky, kx, co} dy[b, (iy + pad - ky) / s, (ix + pad - kx) / s, co] * w[co][ci][ky][kx]
// over the taps for which both divisions are exact and the result lies inside dy.  wt is packed [k*k][Cout][Cin]
// (w.permute(2, 3, 0, 1)); dy [B,Ho,Wo,lddy] with Ho = (H + 2 pad - k) / s + 1; dx [B,H,W,lddx] is written.  What autograd's
// conv backward computes for the reference's nn.Conv2d layers (cuDNN dgrad there); stride-1 layers whose shapes fit use
// fcvsr_conv2d_tc on flipped, transposed weights instead.
extern "C" int fcvsr_conv2d_dgrad_direct(const float* dy, int lddy, const float* wt, float* dx, int lddx, int B, int H, int W,
                                         int Cin, int Cout, int ksize, int stride, cudaStream_t st) {
    if (!dy || !wt || !dx || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || !(ksize & 1) || stride < 1) return FCVSR_ERR_ARG;
    const int pad = ksize / 2;
    ConvDirectArgs a;
    a.x = dy; a.ldx = lddy; a.x_nchw = 0; a.w = wt; a.bias = nullptr;
    a.res = nullptr; a.ldres = 0; a.res2 = nullptr; a.ldres2 = 0; a.y = dx; a.ldy = lddx;
    a.B = B;
    a.H = (H + 2 * pad - ksize) / stride + 1;          // source = dy
    a.W = (W + 2 * pad - ksize) / stride + 1;
    a.Cin = Cout; a.Cout = Cin; a.ks = ksize; a.stride = stride;
    a.Ho = H; a.Wo = W;                                // tile domain = dx
    a.act = FCVSR_ACT_NONE; a.slope = 0.f; a.slope_ptr = nullptr; a.ps = 0; a.y_nchw = 0; a.y2 = nullptr; a.ldy2 = 0;
    a.round_out = 0; a.op16 = 0; a.transposed = 1;
    dim3 grid(((H + 7) / 8) * ((W + 7) / 8), (Cin + CD_TN - 1) / CD_TN, B);
    conv_direct_kernel<<<grid, 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

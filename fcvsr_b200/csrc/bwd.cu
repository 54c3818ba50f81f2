// Adjoint kernels of the FCVSR forward: what autograd derives for the reference's training step
// (CVSR_train/train_LD_freqCVSR_22.py:243-251: sr = model(frames); loss.backward()), written for the NHWC layout of
// this library.  The reference has no backward code of its own for these stages -- it relies on ATen's autograd
// kernels (cuDNN dgrad / wgrad, grid_sampler_2d_backward, elementwise) -- so each kernel cites the forward lines whose
// adjoint it is.
//
//   conv2d_wgrad            d/dW of y = conv(x, W): dW[tap][ci][co] += sum_pix x[pix + tap][ci] * dy[pix][co]
//   colsum                  d/dbias = sum over pixels of dy (deterministic two-stage)
//   flow_warp / _backward   flow_warp (CVSR_freq.py:1188-1227): bilinear gather and its adjoint (scatter + d/d offset)
//   sac / sac_backward      SAC (:1253-1276): vertical then horizontal per-pixel 3-tap filter with the same taps
//   corr_gather_backward    CorrBlock lookup (:1279-1337): the lookup is injective, so its adjoint is a gather too
//
// (The data gradient of a convolution is a convolution: fcvsr_conv2d_tc / fcvsr_conv2d_direct on flipped, transposed
// weights, or fcvsr_conv2d_direct in `transposed` mode for strided layers.  The FFT passes are their own adjoints up to
// conjugation and per-column weights, see fcvsr_b200/autograd.py.)
#include "common.cuh"

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient.  x [B,H,W,ldx] (Cin channels), dy [B,Ho,Wo,lddy] (Cout channels), dw [k*k][Cin][Cout] fp32 (the
// "direct" weight layout), ACCUMULATED with fp32 atomics over the pixel slices (the caller zero-fills it).
// Block = 64 ci x 64 co tile of one tap over one slice of the output pixels; 16 pixels per shared-memory step.
#define WG_TK 16
struct WgradArgs {
    const float* x; int ldx; const float* dy; int lddy; float* dw;
    int B, H, W, Cin, Cout, ks, stride, Ho, Wo;
    long long npix; int pix_per_slice;
};

__global__ void __launch_bounds__(256) conv_wgrad_kernel(WgradArgs a) {
    __shared__ __align__(16) float Xs[WG_TK][64 + 4];
    __shared__ __align__(16) float Gs[WG_TK][64 + 4];
    const int tid = threadIdx.x;
    const int tap = blockIdx.y, ky = tap / a.ks, kx = tap - ky * a.ks, pad = a.ks / 2;
    const int co_tiles = (a.Cout + 63) >> 6;
    const int ci0 = (blockIdx.z / co_tiles) * 64, co0 = (blockIdx.z % co_tiles) * 64;
    const long long p_begin = (long long)blockIdx.x * a.pix_per_slice;
    const long long p_end = min(a.npix, p_begin + a.pix_per_slice);
    const int ig = tid >> 4, jg = tid & 15;              // 4 ci x 4 co per thread
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lp = tid >> 4, lc = (tid & 15) * 4;        // loader: pixel lp of the step, channels lc..lc+3
    for (long long p0 = p_begin; p0 < p_end; p0 += WG_TK) {
        const long long p = p0 + lp;
        float xv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p < p_end) {
            const int ox = (int)(p % a.Wo);
            const long long r = p / a.Wo;
            const int oy = (int)(r % a.Ho), b = (int)(r / a.Ho);
            const int iy = oy * a.stride + ky - pad, ix = ox * a.stride + kx - pad;
            if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
                const float* xp = a.x + (((size_t)b * a.H + iy) * a.W + ix) * a.ldx + ci0 + lc;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ci0 + lc + u < a.Cin) xv[u] = xp[u];
            }
            const float* gp = a.dy + (size_t)p * a.lddy + co0 + lc;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (co0 + lc + u < a.Cout) gv[u] = gp[u];
        }
        *reinterpret_cast<float4*>(&Xs[lp][lc]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
        *reinterpret_cast<float4*>(&Gs[lp][lc]) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WG_TK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&Xs[kk][ig * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Gs[kk][jg * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ig * 4 + i;
        if (ci >= a.Cin) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + jg * 4 + j;
            if (co < a.Cout) atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, acc[i][j]);
        }
    }
}

// Thin heads (Cout <= 4: conv_last0 64 -> 1, the 64 -> 4 offset / similarity heads): the 64 x 64 slab of the kernel above is
// 1/16 .. 1/64 full.  Here a block owns WT_ROWS image rows and ONE filter tap; thread = (4 input channels, pixel slot): it
// walks its pixels (the 4-channel groups of a pixel are one coalesced line), 4 x Cout FMAs each, the slots are summed through
// shared memory and the block adds its 4 x Cin x Cout partial sums with atomics.
#define WT_ROWS 16
__global__ void __launch_bounds__(256) conv_wgrad_thin_kernel(WgradArgs a) {
    __shared__ float red[256 * 16];
    const int ncg = a.Cin >> 2, nslot = 256 / ncg;              // Cin % 4 == 0, Cin <= 256 (host)
    const int cg = threadIdx.x % ncg, slot = threadIdx.x / ncg;
    const int tap = blockIdx.y, ky = tap / a.ks, kx = tap - ky * a.ks, pad = a.ks >> 1;
    const int rows_total = a.B * a.H;
    const int r0 = blockIdx.x * WT_ROWS, r1 = min(r0 + WT_ROWS, rows_total);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[i][o] = 0.f;
    if (slot < nslot) {
        for (int r = r0; r < r1; ++r) {
            const int b = r / a.H, y = r - b * a.H, yy = y + ky - pad;
            if (yy < 0 || yy >= a.H) continue;
            const float* xrow = a.x + ((size_t)(b * a.H + yy) * a.W) * a.ldx + cg * 4;
            const float* grow = a.dy + ((size_t)r * a.W) * a.lddy;
            for (int xo = slot; xo < a.W; xo += nslot) {
                const int xx = xo + kx - pad;
                if (xx < 0 || xx >= a.W) continue;
                const float4 xv = __ldg(reinterpret_cast<const float4*>(xrow + (size_t)xx * a.ldx));
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                for (int o = 0; o < a.Cout; ++o) g[o] = __ldg(grow + (size_t)xo * a.lddy + o);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    acc[0][o] = fmaf(xv.x, g[o], acc[0][o]); acc[1][o] = fmaf(xv.y, g[o], acc[1][o]);
                    acc[2][o] = fmaf(xv.z, g[o], acc[2][o]); acc[3][o] = fmaf(xv.w, g[o], acc[3][o]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 0; o < 4; ++o) red[(i * 4 + o) * 256 + threadIdx.x] = acc[i][o];
    __syncthreads();
    // fixed-order sum over the slots, then one atomic per (ci, co) of the block
    for (int e = threadIdx.x; e < ncg * 16; e += 256) {
        const int c = e % ncg, io = e / ncg, i = io >> 2, o = io & 3;
        if (o >= a.Cout) continue;
        float sum = 0.f;
        for (int sl = 0; sl < nslot; ++sl) sum += red[io * 256 + sl * ncg + c];
        atomicAdd(a.dw + ((size_t)tap * a.Cin + c * 4 + i) * a.Cout + o, sum);
    }
}

extern "C" int fcvsr_conv2d_wgrad(const float* x, int ldx, const float* dy, int lddy, float* dw, int B, int H, int W,
                                  int Cin, int Cout, int ksize, int stride, cudaStream_t st) {
    if (!x || !dy || !dw || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || !(ksize & 1) || stride < 1) return FCVSR_ERR_ARG;
    if (Cout <= 4 && stride == 1 && !(Cin & 3) && Cin <= 256 && !(ldx & 3) && !((uintptr_t)x & 15) && (long long)B * H <= 0x7fffffffLL) {
        WgradArgs t;
        t.x = x; t.ldx = ldx; t.dy = dy; t.lddy = lddy; t.dw = dw;
        t.B = B; t.H = H; t.W = W; t.Cin = Cin; t.Cout = Cout; t.ks = ksize; t.stride = 1;
        t.Ho = H; t.Wo = W; t.npix = (long long)B * H * W; t.pix_per_slice = 0;
        dim3 grid((unsigned)((B * H + WT_ROWS - 1) / WT_ROWS), ksize * ksize);
        conv_wgrad_thin_kernel<<<grid, 256, 0, st>>>(t);
        return fcvsr_launch_status();
    }
    WgradArgs a;
    a.x = x; a.ldx = ldx; a.dy = dy; a.lddy = lddy; a.dw = dw;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ks = ksize; a.stride = stride;
    const int pad = ksize / 2;
    a.Ho = (H + 2 * pad - ksize) / stride + 1;
    a.Wo = (W + 2 * pad - ksize) / stride + 1;
    a.npix = (long long)B * a.Ho * a.Wo;
    const int taps = ksize * ksize;
    const int tiles = ((Cin + 63) / 64) * ((Cout + 63) / 64);
    // enough pixel slices for ~6 blocks per SM, at least 256 pixels each (one atomic pass per slice)
    long long want = (148LL * 6 + (long long)taps * tiles - 1) / ((long long)taps * tiles);
    if (want < 1) want = 1;
    long long pps = (a.npix + want - 1) / want;
    if (pps < 256) pps = 256;
    pps = (pps + WG_TK - 1) / WG_TK * WG_TK;
    a.pix_per_slice = (int)pps;
    const unsigned slices = (unsigned)((a.npix + pps - 1) / pps);
    dim3 grid(slices, taps, tiles);
    conv_wgrad_kernel<<<grid, 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------------------------
// Column sums: out[c] = sum over npix rows of x[row*ldx + c] (bias gradient).  Two deterministic stages: per-block partial
// sums in `scratch` (nblk x C floats, nblk = fcvsr_colsum_blocks), then one block sums them in fixed order.
#define CS_ROWS 64
// 256 threads = G row groups x Cw channels (Cw = min(C, 256)): group g sums rows r0 + g, r0 + g + G, ... of the block's row
// range for its channel, the groups are combined through shared memory in fixed order.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int ldx, int C, long long npix,
                                                             float* __restrict__ scratch) {
    __shared__ float red[256];
    const long long r0 = (long long)blockIdx.x * CS_ROWS, r1 = min(npix, r0 + CS_ROWS);
    const int Cw = C < 256 ? C : 256, G = 256 / Cw;
    const int g = threadIdx.x / Cw, cl = threadIdx.x - g * Cw;
    for (int cb = 0; cb < C; cb += Cw) {
        const int c = cb + cl;
        float s = 0.f;
        if (g < G && c < C)
            for (long long r = r0 + g; r < r1; r += G) s += x[(size_t)r * ldx + c];
        red[threadIdx.x] = s;
        __syncthreads();
        if (g == 0 && c < C) {
            float t = 0.f;
            for (int k = 0; k < G; ++k) t += red[k * Cw + cl];
            scratch[(size_t)blockIdx.x * C + c] = t;
        }
        __syncthreads();
    }
}
// 256 threads = 8 walkers x 32 channels: walker g sums partials g, g + 8, ... of its channel, the walkers are combined in fixed order
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ scratch, int nblk, int C, float* __restrict__ out,
                                                           int accumulate) {
    __shared__ float red[8][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float s = 0.f;
    if (c < C)
        for (int k = g; k < nblk; k += 8) s += scratch[(size_t)k * C + c];
    red[g][cl] = s;
    __syncthreads();
    if (g == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][cl];
        out[c] = accumulate ? out[c] + t : t;
    }
}
extern "C" int fcvsr_colsum(const float* x, int ldx, int C, long long npix, float* scratch, float* out, int accumulate,
                            cudaStream_t st) {
    if (!x || !scratch || !out || C <= 0 || npix <= 0) return FCVSR_ERR_ARG;
    const int nblk = (int)((npix + CS_ROWS - 1) / CS_ROWS);
    colsum_partial_kernel<<<nblk, 256, 0, st>>>(x, ldx, C, npix, scratch);
    colsum_final_kernel<<<(C + 31) / 32, 256, 0, st>>>(scratch, nblk, C, out, accumulate);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------------------------
// flow_warp (CVSR_freq.py:1188-1227): y[b,py,px,c] = bilinear(x[b,:,:,c], px + off[...,0], py + off[...,1]), zeros outside,
// align_corners=True (so the normalisation cancels).  One thread per (pixel, 4 channels).
struct WarpGeo { int i00, i01, i10, i11; float w00, w01, w10, w11; float lx, ly; bool v00, v01, v10, v11; bool inside; };
__device__ __forceinline__ WarpGeo warp_geo(float sx, float sy, int H, int W) {
    WarpGeo g;
    g.inside = sx > -1.f && sx < (float)W && sy > -1.f && sy < (float)H;
    g.i00 = g.i01 = g.i10 = g.i11 = 0;
    g.w00 = g.w01 = g.w10 = g.w11 = 0.f;
    g.lx = g.ly = 0.f;
    g.v00 = g.v01 = g.v10 = g.v11 = false;
    if (!g.inside) return g;
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    g.lx = sx - fx0; g.ly = sy - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const bool vx0 = x0 >= 0, vx1 = x1 < W, vy0 = y0 >= 0, vy1 = y1 < H;
    const int r0 = (vy0 ? y0 : 0) * W, r1 = (vy1 ? y1 : H - 1) * W, q0 = vx0 ? x0 : 0, q1 = vx1 ? x1 : W - 1;
    g.i00 = r0 + q0; g.i01 = r0 + q1; g.i10 = r1 + q0; g.i11 = r1 + q1;
    g.v00 = vy0 && vx0; g.v01 = vy0 && vx1; g.v10 = vy1 && vx0; g.v11 = vy1 && vx1;
    g.w00 = g.v00 ? (1.f - g.ly) * (1.f - g.lx) : 0.f;
    g.w01 = g.v01 ? (1.f - g.ly) * g.lx : 0.f;
    g.w10 = g.v10 ? g.ly * (1.f - g.lx) : 0.f;
    g.w11 = g.v11 ? g.ly * g.lx : 0.f;
    return g;
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}

__global__ void flow_warp_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ off, int ldoff,
                                 float* __restrict__ y, int ldy, int H, int W, int c4n, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % c4n) * 4;
    const size_t pix = idx / c4n;
    const int px = (int)(pix % W), py = (int)((pix / W) % H);
    const size_t img = pix - ((size_t)py * W + px);
    const float2 d = *reinterpret_cast<const float2*>(off + pix * ldoff);
    const WarpGeo g = warp_geo((float)px + d.x, (float)py + d.y, H, W);
    const float* xb = x + img * ldx + c;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.inside) {
        s = f4_scale(*reinterpret_cast<const float4*>(xb + (size_t)g.i00 * ldx), g.w00);
        s = f4_fma(g.w01, *reinterpret_cast<const float4*>(xb + (size_t)g.i01 * ldx), s);
        s = f4_fma(g.w10, *reinterpret_cast<const float4*>(xb + (size_t)g.i10 * ldx), s);
        s = f4_fma(g.w11, *reinterpret_cast<const float4*>(xb + (size_t)g.i11 * ldx), s);
    }
    *reinterpret_cast<float4*>(y + pix * ldy + c) = s;
}

extern "C" int fcvsr_flow_warp(const float* x, int ldx, const float* off, int ldoff, float* y, int ldy, int B, int H, int W,
                               int C, cudaStream_t st) {
    if (!x || !off || !y || (C & 3) || (ldx & 3) || (ldy & 3) || (ldoff & 1) || B <= 0) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * H * W * (C / 4);
    flow_warp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ldx, off, ldoff, y, ldy, H, W, C / 4, total);
    return fcvsr_launch_status();
}

// Adjoint: dx (ACCUMULATED with atomics; the caller zero-fills) and doff [B,H,W,2] (written).  A group of C/4 consecutive
// threads owns a pixel; d/d(offset) is reduced over the channels with shuffles when the group is a power of two <= 32
// lanes, else with atomics into a zero-filled doff.
__global__ void flow_warp_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ off, int ldoff,
                                     const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx,
                                     float* __restrict__ doff, int H, int W, int c4n, size_t total, int shuffle_ok) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < total;
    const size_t id2 = live ? idx : total - 1;
    const int c = (int)(id2 % c4n) * 4;
    const size_t pix = id2 / c4n;
    const int px = (int)(pix % W), py = (int)((pix / W) % H);
    const size_t img = pix - ((size_t)py * W + px);
    const float2 d = *reinterpret_cast<const float2*>(off + pix * ldoff);
    const WarpGeo g = warp_geo((float)px + d.x, (float)py + d.y, H, W);
    float gx = 0.f, gy = 0.f;
    if (live && g.inside) {
        const float4 gv = *reinterpret_cast<const float4*>(dy + pix * lddy + c);
        const float* xb = x + img * ldx + c;
        float* db = dx ? dx + img * lddx + c : nullptr;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 x00 = g.v00 ? *reinterpret_cast<const float4*>(xb + (size_t)g.i00 * ldx) : z4;
        const float4 x01 = g.v01 ? *reinterpret_cast<const float4*>(xb + (size_t)g.i01 * ldx) : z4;
        const float4 x10 = g.v10 ? *reinterpret_cast<const float4*>(xb + (size_t)g.i10 * ldx) : z4;
        const float4 x11 = g.v11 ? *reinterpret_cast<const float4*>(xb + (size_t)g.i11 * ldx) : z4;
        if (db) {
            const float gr[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (g.v00) atomicAdd(db + (size_t)g.i00 * lddx + u, g.w00 * gr[u]);
                if (g.v01) atomicAdd(db + (size_t)g.i01 * lddx + u, g.w01 * gr[u]);
                if (g.v10) atomicAdd(db + (size_t)g.i10 * lddx + u, g.w10 * gr[u]);
                if (g.v11) atomicAdd(db + (size_t)g.i11 * lddx + u, g.w11 * gr[u]);
            }
        }
        // d sample / d sx = (1-ly)(x01 - x00) + ly (x11 - x10);  d sample / d sy = (1-lx)(x10 - x00) + lx (x11 - x01)
        const float ax[4] = {x00.x, x00.y, x00.z, x00.w}, bx[4] = {x01.x, x01.y, x01.z, x01.w};
        const float cx[4] = {x10.x, x10.y, x10.z, x10.w}, ex[4] = {x11.x, x11.y, x11.z, x11.w};
        const float gr[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            gx += gr[u] * ((1.f - g.ly) * (bx[u] - ax[u]) + g.ly * (ex[u] - cx[u]));
            gy += gr[u] * ((1.f - g.lx) * (cx[u] - ax[u]) + g.lx * (ex[u] - bx[u]));
        }
    }
    if (!doff) return;
    if (shuffle_ok) {            // c4n is a power of two <= 32 and divides the block: the pixel's lanes are contiguous in one warp
        for (int o = c4n >> 1; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
        }
        if (live && (idx % c4n) == 0) *reinterpret_cast<float2*>(doff + pix * 2) = make_float2(gx, gy);
    } else if (live) {
        atomicAdd(doff + pix * 2, gx);
        atomicAdd(doff + pix * 2 + 1, gy);
    }
}

extern "C" int fcvsr_flow_warp_backward(const float* x, int ldx, const float* off, int ldoff, const float* dy, int lddy,
                                        float* dx, int lddx, float* doff, int B, int H, int W, int C, cudaStream_t st) {
    if (!x || !off || !dy || (C & 3) || (ldx & 3) || (lddy & 3) || (dx && (lddx & 3)) || (ldoff & 1) || B <= 0) return FCVSR_ERR_ARG;
    const int c4n = C / 4;
    const size_t total = (size_t)B * H * W * c4n;
    const int shuffle_ok = c4n <= 32 && (c4n & (c4n - 1)) == 0;
    if (doff && !shuffle_ok) {
        if (cudaMemsetAsync(doff, 0, (size_t)B * H * W * 2 * sizeof(float), st) != cudaSuccess) return FCVSR_ERR_CUDA;
    }
    flow_warp_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ldx, off, ldoff, dy, lddy, dx, lddx, doff, H, W, c4n,
                                                                       total, shuffle_ok);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------------------------
// SAC (CVSR_freq.py:1253-1276), taps K [B,H,W,ldk] with channel t*C + c (this library's [t][c] order):
//   v[y,x,c]   = sum_t wp[clamp(y+t-1), x, c] * K[y,x,t,c]
//   out[y,x,c] = sum_t v[y, clamp(x+t-1), c]  * K[y,x,t,c]          (the reference applies kernel1 in both passes)
__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 sac_v(const float* __restrict__ wp, int ldw, const float* __restrict__ K, int ldk, size_t img,
                                        int y, int x, int H, int W, int C, int c) {
    const float* kp = K + (img + (size_t)y * W + x) * ldk + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int yy = min(max(y + t - 1, 0), H - 1);
        v = f4_fma4(ldf4(wp + (img + (size_t)yy * W + x) * ldw + c), ldf4(kp + t * C), v);
    }
    return v;
}

// mode 0: out = SAC(wp, K); mode 1 (backward stage A): v -> out, and dv -> out2 from g
__global__ void sac_kernel(const float* __restrict__ wp, int ldw, const float* __restrict__ K, int ldk, float* __restrict__ out,
                           int ldo, int H, int W, int C, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4n = C / 4;
    const int c = (int)(idx % c4n) * 4;
    const size_t pix = idx / c4n;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const size_t img = pix - ((size_t)y * W + x);
    const float* kp = K + pix * ldk + c;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int xx = min(max(x + t - 1, 0), W - 1);
        o = f4_fma4(sac_v(wp, ldw, K, ldk, img, y, xx, H, W, C, c), ldf4(kp + t * C), o);
    }
    *reinterpret_cast<float4*>(out + pix * ldo + c) = o;
}

extern "C" int fcvsr_sac(const float* wp, int ldw, const float* taps, int ldk, float* out, int ldo, int B, int H, int W, int C,
                         cudaStream_t st) {
    if (!wp || !taps || !out || (C & 3) || (ldw & 3) || (ldk & 3) || (ldo & 3) || B <= 0) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * H * W * (C / 4);
    sac_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(wp, ldw, taps, ldk, out, ldo, H, W, C, total);
    return fcvsr_launch_status();
}

// Backward stage A: dv[y,x'] = sum over (x,t) with clamp(x+t-1) == x' of g[y,x] K[y,x,t]  (gather form, no atomics)
__global__ void sac_bwd_a_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ K, int ldk,
                                 float* __restrict__ dv, int H, int W, int C, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4n = C / 4;
    const int c = (int)(idx % c4n) * 4;
    const size_t pix = idx / c4n;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const size_t row = pix - x;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int xs = x - t + 1;                                    // source pixel whose tap t lands on x without clamping
        if (xs >= 0 && xs < W) s = f4_fma4(ldf4(g + (row + xs) * ldg + c), ldf4(K + (row + xs) * ldk + t * C + c), s);
    }
    if (x == 0) s = f4_fma4(ldf4(g + row * ldg + c), ldf4(K + row * ldk + c), s);                         // x = 0, t = 0 clamps to 0
    if (x == W - 1) s = f4_fma4(ldf4(g + (row + x) * ldg + c), ldf4(K + (row + x) * ldk + 2 * C + c), s);   // x = W-1, t = 2
    *reinterpret_cast<float4*>(dv + pix * C + c) = s;
}

// Backward stage B: dK[y,x,t] = g[y,x] v[y,clamp(x+t-1)] + dv[y,x] wp[clamp(y+t-1),x];
//                   dwp[y',x] = sum over (y,t) with clamp(y+t-1) == y' of dv[y,x] K[y,x,t]
__global__ void sac_bwd_b_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ K, int ldk,
                                 const float* __restrict__ wp, int ldw, const float* __restrict__ dv, float* __restrict__ dK,
                                 int lddk, float* __restrict__ dwp, int lddw, int H, int W, int C, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4n = C / 4;
    const int c = (int)(idx % c4n) * 4;
    const size_t pix = idx / c4n;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const size_t img = pix - ((size_t)y * W + x);
    if (dK) {
        const float4 gv = ldf4(g + pix * ldg + c), dvv = ldf4(dv + pix * C + c);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int xx = min(max(x + t - 1, 0), W - 1), yy = min(max(y + t - 1, 0), H - 1);
            const float4 vv = sac_v(wp, ldw, K, ldk, img, y, xx, H, W, C, c);
            const float4 wv = ldf4(wp + (img + (size_t)yy * W + x) * ldw + c);
            *reinterpret_cast<float4*>(dK + pix * lddk + t * C + c) = f4_fma4(gv, vv, f4_mul(dvv, wv));
        }
    }
    if (dwp) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int ys = y - t + 1;
            if (ys >= 0 && ys < H) {
                const size_t q = img + (size_t)ys * W + x;
                s = f4_fma4(ldf4(dv + q * C + c), ldf4(K + q * ldk + t * C + c), s);
            }
        }
        if (y == 0) s = f4_fma4(ldf4(dv + pix * C + c), ldf4(K + pix * ldk + c), s);
        if (y == H - 1) s = f4_fma4(ldf4(dv + pix * C + c), ldf4(K + pix * ldk + 2 * C + c), s);
        *reinterpret_cast<float4*>(dwp + pix * lddw + c) = s;
    }
}

// scratch: B*H*W*C floats (dv).  dtaps / dwp may be NULL.
extern "C" int fcvsr_sac_backward(const float* wp, int ldw, const float* taps, int ldk, const float* g, int ldg, float* scratch,
                                  float* dtaps, int lddk, float* dwp, int lddw, int B, int H, int W, int C, cudaStream_t st) {
    if (!wp || !taps || !g || !scratch || (C & 3) || (ldw & 3) || (ldk & 3) || (ldg & 3) || (dtaps && (lddk & 3)) ||
        (dwp && (lddw & 3)) || B <= 0)
        return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * H * W * (C / 4);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    sac_bwd_a_kernel<<<blocks, 256, 0, st>>>(g, ldg, taps, ldk, scratch, H, W, C, total);
    sac_bwd_b_kernel<<<blocks, 256, 0, st>>>(g, ldg, taps, ldk, wp, ldw, scratch, dtaps, lddk, dwp, lddw, H, W, C, total);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------------------------
// CorrBlock lookup adjoint.  Forward (mgaa.cu: corr_gather_kernel): out[b,p,i*9+j] = A[b,p2,m] * Bv[b,p2,m] / sqrt(C2) with
// (p2, ch) = divmod-decomposition of F = p*C2 + (y0+j-4)*2 + (x0+i-4) and m the interleaved index of reference channel ch.
// Every (p2, ch) is read by at most one (p, i, j): dprod[b,p2,ch] = dout[b,p,i*9+j] (or 0), dA = dprod * Bv / sqrt(C2),
// dB = dprod * A / sqrt(C2).  dS is WRITTEN at the two groups (a_off / b_off), layout like S.
__global__ void corr_gather_bwd_kernel(const float* __restrict__ S, int ldS, int a_off, int b_off, const float* __restrict__ dout,
                                       int ldo, float* __restrict__ dS, int lddS, int H, int Wf, int C2, float inv_sqrt_c,
                                       size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int P = H * Wf;
    const int ch = (int)(idx % C2);
    const size_t bp = idx / C2;
    const int p2 = (int)(bp % P), b = (int)(bp / P);
    const long long F = (long long)ch * P + p2;
    const int p = (int)(F / C2), rem = (int)(F - (long long)p * C2);
    const int row = rem >> 1, col = rem & 1;
    const int y0 = p / Wf, x0 = p - y0 * Wf;
    const int i = col - x0 + 4, j = row - y0 + 4;
    float dp = 0.f;
    if (i >= 0 && i < 9 && j >= 0 && j < 9) dp = dout[((size_t)b * P + p) * ldo + i * 9 + j] * inv_sqrt_c;
    const int half = C2 / 2;
    const int mi = ch < half ? 2 * ch + 1 : 2 * (ch - half);
    const float* s = S + ((size_t)b * P + p2) * ldS;
    float* d = dS + ((size_t)b * P + p2) * lddS;
    d[a_off + mi] = dp * s[b_off + mi];
    d[b_off + mi] = dp * s[a_off + mi];
}

extern "C" int fcvsr_corr_gather_backward(const float* S, int ldS, int a_off, int b_off, const float* dout, int ldo, float* dS,
                                          int lddS, int B, int H, int Wf, int C2, cudaStream_t st) {
    if (!S || !dout || !dS || C2 <= 0 || (C2 & 1) || a_off == b_off) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * H * Wf * C2;
    corr_gather_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S, ldS, a_off, b_off, dout, ldo, dS, lddS, H, Wf, C2,
                                                                          rsqrtf((float)C2), total);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------------------------
// Weight packing for the tensor-core convolution in ONE launch (the training path re-packs every weight twice per step:
// forward and data gradient).  w is the reference's [Cout][Cin][k][k]; out is K-major [rows][k*k*C] with TF32 rounding:
//   transposed = 0:  out[co][tap][ci] = w[co][ci][tap]                      rows = max(Cout, rows_pad), C = Cin   (forward)
//   transposed = 1:  out[ci][tap][co] = w[co][ci][k*k-1-tap]                rows = max(Cin, rows_pad),  C = Cout  (dgrad: the
//                    data gradient of a stride-1 'same' convolution is the convolution with flipped, transposed weights)
// rows beyond the real count are zero (thin heads are padded to 16 GEMM columns).
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int kk, int transposed,
                                        int rows, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int C = transposed ? Cout : Cin, R = transposed ? Cin : Cout;
    const int c = (int)(idx % C);
    const int tap = (int)((idx / C) % kk);
    const int r = (int)(idx / ((size_t)C * kk));
    float v = 0.f;
    if (r < R) {
        const int co = transposed ? c : r, ci = transposed ? r : c, t = transposed ? kk - 1 - tap : tap;
        v = round_tf32(w[((size_t)co * Cin + ci) * kk + t]);
    }
    out[idx] = v;
    (void)rows;
}

extern "C" int fcvsr_pack_conv_weight(const float* w, float* out, int Cout, int Cin, int ksize, int transposed, int rows_pad,
                                      cudaStream_t st) {
    if (!w || !out || Cout <= 0 || Cin <= 0 || ksize <= 0) return FCVSR_ERR_ARG;
    const int R = transposed ? Cin : Cout, C = transposed ? Cout : Cin;
    const int rows = R > rows_pad ? R : rows_pad;
    const size_t total = (size_t)rows * ksize * ksize * C;
    pack_conv_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, out, Cout, Cin, ksize * ksize, transposed, rows, total);
    return fcvsr_launch_status();
}

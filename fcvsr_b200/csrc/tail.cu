// Up-sampling tail helpers (GShiftNet.forward, CVSR_freq.py:2739-2751).
//   pixel_shuffle_nhwc   F.pixel_shuffle(x, 2) between channel slices of NHWC buffers (:2740-2741)
//   bilinear_up4         F.interpolate(center LR frame, x4, 'bilinear', align_corners=False) (:2750)
// (The pixel shuffles that follow a convolution are fused into that convolution's epilogue.)
#include "common.cuh"
#include <string.h>

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

// ---- conv_last0: 3x3, 64 -> 1 channel, bf16 NHWC input, + bias + residual plane (CVSR_freq.py:2749-2751) -------------
// On tcgen05 this is an N = 16 MMA with one useful column and a nine-fold re-read of the input: A-operand bound (369 us at
// 720x1280, batch 4, for 73 us of HBM time).  The first CUDA-core version (a thread = 4 outputs, weights in the constant bank)
// needed ~875 instructions per output pixel (bf16 unpacking + 576 FFMA) and ran at 254 us, issue bound.  The convolution is
// linear in the filter taps, so it splits into
//     d[p][k] = sum_c x[p][c] * w[k][c]          a [pixels x 64] x [64 x 9] GEMM: every input pixel is read ONCE
//     y[p]    = bias + res[p] + sum_k d[p + tap_k][k]      nine shifted adds of single values
// The GEMM runs on the tensor cores through warp-level mma.sync (m16n8k16, bf16, fp32 accumulate; N = 9 padded to 16 is far too
// thin for a tcgen05 tile and the kernel is bound by the one read of the input anyway): a block stages a haloed 18 x 34 pixel
// tile (128 B per pixel, 16-byte chunks XOR-swizzled by the pixel so that ldmatrix is conflict-free) with zero-filling cp.async,
// its 8 warps walk the 16-pixel row groups (A by ldmatrix.x4, the 9 x 64 weights as B fragments in 16 registers), d goes to
// shared memory as nine planes, and every thread sums the nine shifted taps of two outputs.  Weights are rounded to bf16 like
// every other convolution weight of the bf16 mode.
#define L1_TH 16
#define L1_TW 32
#define L1_THREADS 256
#define L1_HW (L1_TW + 2)
#define L1_NPX ((L1_TH + 2) * L1_HW)               // 612 haloed pixels
#define L1_MT ((L1_NPX + 15) / 16)                 // 39 row groups of 16 pixels
#define L1_DPITCH 644                              // floats per d plane: >= 16 * L1_MT and == 4 (mod 32): conflict-free fragment stores
struct ToOneParams {
    const unsigned short* x; int ldx;
    const float* res; float* y;
    int B, H, W;
    float bias;
    unsigned w16[16 * 32];     // bf16 pairs: w16[n * 32 + k / 2] = (w[n][k], w[n][k + 1]), n = tap (9..15 zero), k = channel
};

__global__ void __launch_bounds__(L1_THREADS, 2) conv3x3_c64_to1_kernel(const __grid_constant__ ToOneParams p) {
    extern __shared__ __align__(16) unsigned char l1_smem[];
    float* dsm = reinterpret_cast<float*>(l1_smem + L1_MT * 16 * 128);       // [9][L1_DPITCH]
    const int tx0 = blockIdx.x * L1_TW, ty0 = blockIdx.y * L1_TH, b = blockIdx.z;
    const unsigned short* xb = p.x + (size_t)b * p.H * p.W * p.ldx;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(l1_smem);
    for (int e = threadIdx.x; e < L1_MT * 16 * 8; e += L1_THREADS) {
        const int px = e >> 3, c8 = e & 7;
        const int r = px / L1_HW, c = px - r * L1_HW;
        const int yy = ty0 - 1 + r, xx = tx0 - 1 + c;
        const bool ok = px < L1_NPX && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        const void* src = ok ? (const void*)(xb + ((size_t)yy * p.W + xx) * p.ldx + c8 * 8) : (const void*)p.x;
        const unsigned dst = sbase + px * 128 + ((c8 ^ (px & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    // the weights go through shared memory (row pitch 36 words: conflict-free fragment reads): a per-lane index into the kernel
    // parameters is a constant-bank load that is replayed once per distinct address (measured: half of the kernel's stalls)
    unsigned* wsm = reinterpret_cast<unsigned*>(dsm + 9 * L1_DPITCH);         // [16][36]
    for (int i = threadIdx.x; i < 16 * 32; i += L1_THREADS) wsm[(i >> 5) * 36 + (i & 31)] = p.w16[i];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // B fragments (col-major K x N == w[n][k]): b0 = k pair tig * 2, b1 = k pair tig * 2 + 8 of column n = gid (+ 8 for the second n tile)
    unsigned bf[4][2][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            bf[ks][nt][0] = wsm[(nt * 8 + gid) * 36 + ks * 8 + tig];
            bf[ks][nt][1] = wsm[(nt * 8 + gid) * 36 + ks * 8 + tig + 4];
        }
    for (int mt = warp; mt < L1_MT; mt += L1_THREADS / 32) {
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
        // ldmatrix.x4: lanes 0-15 address rows 0-15 of the low 8 k (16-byte chunk 2 ks), lanes 16-31 the same rows of the high 8 k
        const int row = mt * 16 + (lane & 15);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int chunk = 2 * ks + (lane >> 4);
            const unsigned addr = sbase + row * 128 + ((chunk ^ (row & 7)) << 4);
            unsigned a0, a1, a2, a3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bf[ks][nt][0]), "r"(bf[ks][nt][1]));
        }
        // accumulator fragment: c0, c1 = row gid, columns 2 tig, 2 tig + 1; c2, c3 = row gid + 8.  Columns 0..8 are taps.
        const int r0 = mt * 16 + gid;
        dsm[(2 * tig) * L1_DPITCH + r0] = acc[0][0];
        dsm[(2 * tig + 1) * L1_DPITCH + r0] = acc[0][1];
        dsm[(2 * tig) * L1_DPITCH + r0 + 8] = acc[0][2];
        dsm[(2 * tig + 1) * L1_DPITCH + r0 + 8] = acc[0][3];
        if (tig == 0) {                                     // column 8 lives in the second n tile
            dsm[8 * L1_DPITCH + r0] = acc[1][0];
            dsm[8 * L1_DPITCH + r0 + 8] = acc[1][2];
        }
    }
    __syncthreads();
    // out(y, x) = sum_{ky, kx} d[(y + ky) * L1_HW + x + kx][ky * 3 + kx]: two horizontally adjacent outputs per thread
    const int ty = threadIdx.x >> 4, x0 = (threadIdx.x & 15) * 2;
    float o0 = p.bias, o1 = p.bias;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float* d = dsm + (ky * 3 + kx) * L1_DPITCH + (ty + ky) * L1_HW + x0 + kx;
            o0 += d[0];
            o1 += d[1];
        }
    const int y = ty0 + ty, x = tx0 + x0;
    if (y < p.H && x < p.W) {
        const size_t o = ((size_t)b * p.H + y) * p.W + x;
        if (x + 1 < p.W && !(o & 1)) {
            const float2 r = p.res ? *reinterpret_cast<const float2*>(p.res + o) : make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(p.y + o) = make_float2(o0 + r.x, o1 + r.y);
        } else {
            p.y[o] = o0 + (p.res ? p.res[o] : 0.f);
            if (x + 1 < p.W) p.y[o + 1] = o1 + (p.res ? p.res[o + 1] : 0.f);
        }
    }
}

static unsigned short l1_bf16_rn(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (unsigned short)(u >> 16);     // inf / nan
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

// x: bf16 NHWC [B,H,W,ldx] (64 channels used); w_host: HOST pointer to 9*64 floats [ky][kx][c] (rounded to bf16 here; they
// travel as kernel parameters); res / y: fp32 planes [B,H,W].  Replaces conv_last0 + the bilinear-skip add (CVSR_freq.py:2749-2751).
extern "C" int fcvsr_conv3x3_c64_to1(const void* x, int ldx, const float* w_host, float bias, const float* res, float* y, int B,
                                     int H, int W, cudaStream_t st) {
    if (!x || !w_host || !y || B <= 0 || H <= 0 || W <= 0 || (ldx & 7) || ((uintptr_t)x & 15)) return FCVSR_ERR_ARG;
    ToOneParams p;
    p.x = (const unsigned short*)x; p.ldx = ldx; p.res = res; p.y = y; p.B = B; p.H = H; p.W = W; p.bias = bias;
    for (int n = 0; n < 16; ++n)
        for (int k2 = 0; k2 < 32; ++k2) {
            const unsigned lo = n < 9 ? l1_bf16_rn(w_host[n * 64 + 2 * k2]) : 0u, hi = n < 9 ? l1_bf16_rn(w_host[n * 64 + 2 * k2 + 1]) : 0u;
            p.w16[n * 32 + k2] = lo | (hi << 16);
        }
    const size_t smem = (size_t)L1_MT * 16 * 128 + (size_t)9 * L1_DPITCH * 4 + 16 * 36 * 4;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(conv3x3_c64_to1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return FCVSR_ERR_CUDA;
        attr = true;
    }
    dim3 grid((W + L1_TW - 1) / L1_TW, (H + L1_TH - 1) / L1_TH, B);
    conv3x3_c64_to1_kernel<<<grid, L1_THREADS, smem, st>>>(p);
    return fcvsr_launch_status();
}

// ---- stride-2 3x3 convolutions of the pyramid (rconcat1/2, CVSR_freq.py:2671-2672, :2735-2736) on the tensor cores -------
// A stride-2, pad-1 3x3 convolution is the stride-1 convolution sampled at the even pixels: the stride-1 tcgen05 kernel
// produces the full-resolution map (4x the useful FLOPs at ~40x the CUDA-core kernel's rate) and this pass keeps
// y[b, i, j, :] = x[b, 2i, 2j, :], writing the fp32 stream and its operand-typed copy.  One thread per 4 channels.
__global__ void subsample2_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, void* __restrict__ y2,
                                  int ldy2, int H, int W, int c4, size_t total, int op16) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c = (int)(e % c4) * 4;
    size_t px = e / c4;
    const int Wo = W >> 1, Ho = H >> 1;
    const int j = (int)(px % Wo);
    px /= Wo;
    const int i = (int)(px % Ho), b = (int)(px / Ho);
    const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * H + 2 * i) * W + 2 * j) * ldx + c);
    const size_t o = ((size_t)b * Ho + i) * Wo + j;
    if (y) *reinterpret_cast<float4*>(y + o * ldy + c) = v;
    if (y2) store_operand4(y2, o * ldy2 + c, v, op16);
}

// x: fp32 NHWC [B,H,W,ldx] (H, W even, C % 4 == 0 channels used); y: fp32 [B,H/2,W/2,ldy] or NULL; y2: operand-typed copy
// (TF32-rounded fp32 or bf16) [B,H/2,W/2,ldy2] or NULL.
extern "C" int fcvsr_subsample2(const float* x, int ldx, float* y, int ldy, void* y2, int ldy2, int B, int H, int W, int C,
                                int op16, cudaStream_t st) {
    if (!x || (!y && !y2) || B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || (C & 3) || (ldx & 3) || (ldy & 3) || (ldy2 & 3))
        return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 4);
    subsample2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ldx, y, ldy, y2, ldy2, H, W, C / 4, total, op16);
    return fcvsr_launch_status();
}

// ---- 8-bit output of the evaluation drivers (CVSR_train/test_LD_freqCVSR.py:85-93) ----------------------------------------
// out[b, y, x] = (uint8) trunc(clamp(v[b, y, x], 0, 1) * 255) for the top-left Ho x Wo crop of a [B, H, W] fp32 plane: the
// reference crops the padded rows, clamps, scales and converts with numpy's astype(np.uint8) (truncation toward zero).
__global__ void quantize_u8_kernel(const float* __restrict__ v, unsigned char* __restrict__ out, int H, int W, int Ho, int Wo,
                                   size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % Wo), y = (int)((idx / Wo) % Ho);
    const size_t b = idx / ((size_t)Wo * Ho);
    const float f = fminf(fmaxf(v[(b * H + y) * W + x], 0.f), 1.f) * 255.0f;     // NaN -> 0 through fmaxf
    out[idx] = (unsigned char)(int)f;
}

extern "C" int fcvsr_quantize_u8(const float* v, unsigned char* out, int B, int H, int W, int Ho, int Wo, cudaStream_t st) {
    if (!v || !out || B <= 0 || Ho <= 0 || Wo <= 0 || Ho > H || Wo > W) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * Ho * Wo;
    quantize_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v, out, H, W, Ho, Wo, total);
    return fcvsr_launch_status();
}

// ---- multi-channel output of the mmedit variants (FCVSRNet / FCVSR_SNet, 21 -> 3 channels:
// mmedit_train/mmedit/models/backbones/sr_backbones/fcvsr.py:133-136) ---------------------------------------------------------
// out[b,c,y,x] (NCHW, the module's output layout) = t[b,y,x,c] (NHWC conv_last0 result, pixel stride ldt) +
// bilinear_x4(center[b,c])[y,x] with align_corners=False, center = LR frame [B,C,H,W] with batch stride `bstride`.
__global__ void rgb_tail_kernel(const float* __restrict__ t, int ldt, const float* __restrict__ center, size_t bstride,
                                float* __restrict__ out, int C, int H, int W, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Wo = 4 * W, Ho = 4 * H;
    const int x = (int)(idx % Wo), y = (int)((idx / Wo) % Ho);
    const int c = (int)((idx / ((size_t)Wo * Ho)) % C);
    const size_t b = idx / ((size_t)Wo * Ho * C);
    const float sy = fmaxf(0.25f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.25f * (x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float* p = center + b * bstride + (size_t)c * H * W;
    const float base = (1.f - ly) * ((1.f - lx) * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1]) +
                       ly * ((1.f - lx) * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1]);
    out[idx] = t[((b * Ho + y) * Wo + x) * ldt + c] + base;
}

extern "C" int fcvsr_rgb_tail(const float* t, int ldt, const float* center, long long bstride, float* out, int B, int C, int H,
                              int W, cudaStream_t st) {
    if (!t || !center || !out || B <= 0 || C <= 0 || C > ldt || H <= 0 || W <= 0) return FCVSR_ERR_ARG;
    const size_t total = (size_t)B * C * 16 * H * W;
    rgb_tail_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t, ldt, center, (size_t)bstride, out, C, H, W, total);
    return fcvsr_launch_status();
}

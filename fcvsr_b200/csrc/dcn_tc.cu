// Modulated deformable convolution forward on the tensor cores: fused bilinear gather + tcgen05 implicit GEMM.
//
// Same operator, tensors and semantics as dcn.cu (NCHW fp32 exactly like the reference extension,
// ops/dcn/src/deform_conv_cuda.cpp:486-564 / deform_conv_cuda_kernel.cu:570-632), for groups == 1, Cin % 32 == 0,
// (Cin / deformable_groups) % 4 == 0, Cout % 16 == 0, Cout <= 256.  The input is first transposed to NHWC in a
// caller-provided scratch buffer (one read + one write of x): with channels contiguous a bilinear corner of 32 channels is
// ONE 128-byte line and a thread fetches it as float4 -- the NCHW gather needed 4x the load instructions, each touching
// 2-3 lines, and measured slower than the CUDA-core kernel.  The sampled, mask-modulated column matrix never exists in HBM and not even as a
// whole tile: eight producer warps compute it 32 channels of one filter tap at a time straight into the K-major
// SWIZZLE_128B operand layout in shared memory (M = 128 consecutive output pixels), together with the matching
// [Cout x 32] slice of the weights (gathered from the reference's [Cout][Cin][kh][kw] layout, no host re-pack), one
// elected thread issues the kind::tf32 MMAs into a double-buffered TMEM accumulator, and four epilogue warps add the
// bias and store NCHW (lane = pixel, so every store instruction writes 128 contiguous bytes).
//
// Arithmetic: operands rounded to TF32 (cvt.rna), fp32 accumulate -- the same contract as the "tf32" mode of the
// convolutions (|error| <~ 1e-3 relative to the output scale); dcn.cu remains the exact-fp32 path.
// Roofline: 2*Cin*Cout*kh*kw FLOP per output pixel on the tensor pipe, 4*(Cin + Cout + 3*dg*kh*kw) bytes per pixel of
// HBM traffic.  Measured (64->64 3x3, dg 16, 180x320, sigma = 2 px random offsets): 183 us at batch 1 / 588 us at batch 4
// (dcn.cu: 520 / 2000 us).  ncu: the producers are instruction-issue bound (sample geometry is ~60 instructions per
// (pixel, tap, deformable group) and with dg = 16 it serves only 4 channels), DRAM at 6 %, L2 hit rate 88 %; the tensor
// pipe is idle most of the time.  A geometry pre-pass shared by the channels of a group is the next step.
#include "tc_common.cuh"

#define DT_M 128
#define DT_TH 8                         // tile = 8 x 16 output pixels: a compact gather footprint for the (small) L1
#define DT_TW 16
#define DT_GROUPS 1                     // producer groups: group i fills stages i, i + DT_GROUPS, ...
#define DT_PROD 512                     // threads per producer group: (8 granules) x (64 pixel rows), 2 pixels each
#define DT_PROD_WARPS (DT_GROUPS * DT_PROD / 32)
#define DT_MMA_WARP DT_PROD_WARPS
#define DT_EPI0 (DT_PROD_WARPS + 1)
#define DT_THREADS (32 * (DT_PROD_WARPS + 1 + 4))
#define DT_MAXSTAGE 8
#define DT_SMEM_MAX (224 * 1024)

struct DcnTcArgs {
    const float* xt; const float* w; const float* bias; const float* offset; const float* mask; float* y;
    int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, Ho, Wo;
    long long off_bs, mask_bs;
    int off_cs, off_ps;            // offset / mask addressing: element (channel c, output pixel p) at c * off_cs + p * off_ps
                                   // (NCHW planes: cs = Ho*Wo, ps = 1; NHWC rows of a conv_offset_mask output: cs = 1, ps = ld)
    int mask_sigmoid;
    int tiles_x, tiles_per_img, tiles, nstage, stage_bytes;
    int* err;
};

__global__ void __launch_bounds__(DT_THREADS, 1) dcn_tc_kernel(const DcnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + (size_t)a.nstage * a.stage_bytes);
    uint64_t* full = bars;                         // [MAXSTAGE] DT_PROD arrivals
    uint64_t* empty = bars + DT_MAXSTAGE;          // [MAXSTAGE] 1 (tcgen05.commit)
    uint64_t* tm_full = bars + 2 * DT_MAXSTAGE;    // [2]
    uint64_t* tm_empty = tm_full + 2;              // [2] 4 epilogue warps
    uint32_t* tmem_slot = (uint32_t*)(tm_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.Cout;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(2 * N)) tmem_cols <<= 1;
    if (threadIdx.x == 0) {
        for (int i = 0; i < DT_MAXSTAGE; ++i) { mbar_init(&full[i], DT_PROD); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == DT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int P = a.Ho * a.Wo, kk2 = a.kh * a.kw, kcc = a.Cin >> 5, nchunk = kk2 * kcc;
    const size_t HW = (size_t)a.H * a.W;

    if (warp < DT_PROD_WARPS) {
        // ===== producers: thread = (4-channel granule gc of the chunk, tile pixels mrow and mrow + 64).  Branch-free (clamped
        // indices, zero weights).  The loop is instruction-issue bound (ncu: issue slots 50 % busy with div/mod and 64-bit
        // address arithmetic in it), so everything that can be is a counter or a 32-bit offset from a per-tile base pointer.
        // The offsets / mask of the NEXT chunk (streamed from HBM, the longest latency) and this chunk's weight elements are
        // requested before the gathers are consumed. =====
        const int t = threadIdx.x;
        const int gc = t & 7, mrow = t >> 3;
        const int ch_per_dg = a.Cin / a.dg;
        const int och = a.off_cs, opx = a.off_ps;
        int wbase[4];                                     // ((n * Cin + j) * kk2) of this thread's first four weight elements
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = t + u * DT_PROD;
            wbase[u] = ((e >> 5) * a.Cin + (e & 31)) * kk2;
        }
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
            const int b = tile / a.tiles_per_img, tr = tile - b * a.tiles_per_img;
            const int ty = tr / a.tiles_x, ty0 = ty * DT_TH, tx0 = (tr - ty * a.tiles_x) * DT_TW;
            const float* xb = a.xt + (size_t)b * HW * a.Cin + gc * 4;
            const float* offb = a.offset + (size_t)b * a.off_bs;
            const float* mskb = a.mask ? a.mask + (size_t)b * a.mask_bs : a.offset;      // never read when there is no mask
            int pi[2], hob[2], wob[2];
            bool pok[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int m = mrow + 64 * r;
                const int ho = ty0 + (m >> 4), wo = tx0 + (m & 15);
                pok[r] = ho < a.Ho && wo < a.Wo;
                pi[r] = pok[r] ? ho * a.Wo + wo : 0;
                hob[r] = ho * a.sh - a.ph;
                wob[r] = wo * a.sw - a.pw;
            }
            // geometry of chunk (tap 0, cc 0)
            float oh[2], ow[2], mk[2];
            {
                const int g = (gc * 4) / ch_per_dg;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    oh[r] = __ldg(offb + g * 2 * kk2 * och + pi[r] * opx);
                    ow[r] = __ldg(offb + (g * 2 * kk2 + 1) * och + pi[r] * opx);
                    mk[r] = a.mask ? __ldg(mskb + g * kk2 * och + pi[r] * opx) : 1.f;
                }
            }
            int ki = 0, kj = 0;
            for (int tap = 0; tap < kk2; ++tap) {
                for (int cc = 0; cc < kcc; ++cc) {
                    int idx[2][4];
                    float wt[2][4];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float h = (float)(hob[r] + ki * a.dh) + oh[r], w = (float)(wob[r] + kj * a.dw) + ow[r];
                        float m_ = mk[r];
                        if (a.mask_sigmoid && a.mask) m_ = 1.f / (1.f + __expf(-m_));
                        const bool in = pok[r] && h > -1.f && w > -1.f && h < (float)a.H && w < (float)a.W;
                        if (!in) m_ = 0.f;
                        const float hcl = in ? h : 0.f, wcl = in ? w : 0.f;
                        const float fh = floorf(hcl), fw = floorf(wcl);
                        const int h0 = (int)fh, w0 = (int)fw, h1 = h0 + 1, w1 = w0 + 1;
                        const float lh = hcl - fh, lw = wcl - fw, hh = 1.f - lh, hw = 1.f - lw;
                        const bool vh0 = h0 >= 0, vh1 = h1 <= a.H - 1, vw0 = w0 >= 0, vw1 = w1 <= a.W - 1;
                        const int r0 = (vh0 ? h0 : 0) * a.W, r1 = (vh1 ? h1 : a.H - 1) * a.W;
                        const int q0 = vw0 ? w0 : 0, q1 = vw1 ? w1 : a.W - 1;
                        idx[r][0] = (r0 + q0) * a.Cin; idx[r][1] = (r0 + q1) * a.Cin;
                        idx[r][2] = (r1 + q0) * a.Cin; idx[r][3] = (r1 + q1) * a.Cin;
                        const float hm = hh * m_, lm = lh * m_;
                        wt[r][0] = (vh0 && vw0) ? hm * hw : 0.f;
                        wt[r][1] = (vh0 && vw1) ? hm * lw : 0.f;
                        wt[r][2] = (vh1 && vw0) ? lm * hw : 0.f;
                        wt[r][3] = (vh1 && vw1) ? lm * lw : 0.f;
                    }
                    float4 xv[2][4];
                    const float* xg = xb + cc * 32;
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) xv[r][k] = __ldg(reinterpret_cast<const float4*>(xg + idx[r][k]));
                    // this thread's share of the [N x 32] weight slice (first batch of 4 requested now, stored after the barrier)
                    const int woff = cc * 32 * kk2 + tap;
                    float wv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (t + u * DT_PROD < N * 32) wv[u] = __ldg(a.w + wbase[u] + woff);
                    // next chunk's offsets / mask
                    {
                        int ncc = cc + 1, ntap = tap;
                        if (ncc == kcc) { ncc = 0; ++ntap; }
                        if (ntap < kk2) {
                            const int g = (ncc * 32 + gc * 4) / ch_per_dg;
                            const int o = (g * 2 * kk2 + 2 * ntap) * och;
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                oh[r] = __ldg(offb + o + pi[r] * opx);
                                ow[r] = __ldg(offb + o + och + pi[r] * opx);
                                mk[r] = a.mask ? __ldg(mskb + (g * kk2 + ntap) * och + pi[r] * opx) : 1.f;
                            }
                        }
                    }
                    mbar_wait_warp(&empty[stage], phase ^ 1, a.err, 21);
                    uint8_t* sA = smem + stage * a.stage_bytes;
                    uint8_t* sB = sA + DT_M * 128;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = t + u * DT_PROD;
                        if (e < N * 32) {
                            const int n = e >> 5, j = e & 31;
                            *reinterpret_cast<float*>(sB + n * 128 + (((j >> 2) ^ (n & 7)) << 4) + ((j & 3) << 2)) = round_tf32(wv[u]);
                        }
                    }
                    for (int e0 = t + 4 * DT_PROD; e0 < N * 32; e0 += 4 * DT_PROD) {       // Cout > 64: further batches of 4
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int e = e0 + u * DT_PROD;
                            if (e < N * 32) wv[u] = __ldg(a.w + ((e >> 5) * a.Cin + (e & 31)) * kk2 + woff);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int e = e0 + u * DT_PROD;
                            if (e < N * 32) {
                                const int n = e >> 5, j = e & 31;
                                *reinterpret_cast<float*>(sB + n * 128 + (((j >> 2) ^ (n & 7)) << 4) + ((j & 3) << 2)) = round_tf32(wv[u]);
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        float4 o;
                        o.x = wt[r][0] * xv[r][0].x + wt[r][1] * xv[r][1].x + wt[r][2] * xv[r][2].x + wt[r][3] * xv[r][3].x;
                        o.y = wt[r][0] * xv[r][0].y + wt[r][1] * xv[r][1].y + wt[r][2] * xv[r][2].y + wt[r][3] * xv[r][3].y;
                        o.z = wt[r][0] * xv[r][0].z + wt[r][1] * xv[r][1].z + wt[r][2] * xv[r][2].z + wt[r][3] * xv[r][3].z;
                        o.w = wt[r][0] * xv[r][0].w + wt[r][1] * xv[r][1].w + wt[r][2] * xv[r][2].w + wt[r][3] * xv[r][3].w;
                        const int m = mrow + 64 * r;
                        *reinterpret_cast<float4*>(sA + m * 128 + ((gc ^ (m & 7)) << 4)) =
                            make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> tensor-core (async proxy) reads
                    mbar_arrive(&full[stage]);
                    if (++stage == a.nstage) { stage = 0; phase ^= 1; }
                }
                if (++kj == a.kw) { kj = 0; ++ki; }
            }
        }
    } else if (warp == DT_MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t pacc = 0;
            for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
                mbar_wait(&tm_empty[acc], pacc ^ 1, a.err, 22);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
                for (int s = 0; s < nchunk; ++s) {
                    mbar_wait(&full[stage], phase, a.err, 23);
                    tc_fence_after();
                    const uint64_t a_d = make_desc(smem_u32(smem + (size_t)stage * a.stage_bytes));
                    const uint64_t b_d = make_desc(smem_u32(smem + (size_t)stage * a.stage_bytes + DT_M * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, a_d + 2 * k, b_d + 2 * k, idesc, (s | k) ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (++stage == a.nstage) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tm_full[acc]);
                if (++acc == 2) { acc = 0; pacc ^= 1; }
            }
        }
    } else {
        // ===== epilogue: TMEM lane quarter = warp % 4, lane = pixel, NCHW stores coalesced across the warp =====
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int acc = 0; uint32_t pacc = 0;
        for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
            const int b = tile / a.tiles_per_img, tr = tile - b * a.tiles_per_img;
            const int ho = (tr / a.tiles_x) * DT_TH + m / DT_TW, wo = (tr % a.tiles_x) * DT_TW + m % DT_TW;
            const int p = (ho < a.Ho && wo < a.Wo) ? ho * a.Wo + wo : P;
            mbar_wait_warp(&tm_full[acc], pacc, a.err, 24);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N);
            float* yb = a.y + (size_t)b * a.Cout * P + p;
            for (int n0 = 0; n0 < N; n0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + n0, r);
                tmem_ld_wait();
                if (p < P) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        yb[(size_t)(n0 + j) * P] = __uint_as_float(r[j]) + (a.bias ? __ldg(a.bias + n0 + j) : 0.f);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tm_empty[acc]);
            if (++acc == 2) { acc = 0; pacc ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == DT_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// NCHW [B][C][P] -> NHWC [B][P][C] through a 32x33 shared tile (coalesced on both sides); rnd: store TF32-rounded values
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int P, int rnd) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* xb = x + (size_t)b * C * P;
    float* yb = y + (size_t)b * C * P;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < P) ? xb[(size_t)c * P + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < P && c < C) yb[(size_t)p * C + c] = rnd ? round_tf32(tile[threadIdx.x][i]) : tile[threadIdx.x][i];
    }
}

// NCHW fp32 [B,C,H,W] -> NHWC fp32 [B,H,W,C] (optionally TF32-rounded: the operand layout of the tcgen05 kernels).  The DCN
// modules keep the reference's NCHW tensor contract (ops/dcn/deform_conv.py), the kernels gather from pixel-major copies.
extern "C" int fcvsr_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, int round_tf32, cudaStream_t st) {
    if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0) return FCVSR_ERR_ARG;
    nchw_to_nhwc_kernel<<<dim3((H * W + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, st>>>(x, y, C, H * W, round_tf32);
    return fcvsr_launch_status();
}

// Returns FCVSR_ERR_UNSUPPORTED for shapes outside the tensor-core kernel's class (the caller then uses dcn.cu).
extern "C" int fcvsr_modulated_deform_conv_forward_tc(const float* input, const float* weight, const float* bias,
                                                      const float* offset, const float* mask, float* output, int B, int Cin,
                                                      int H, int W, int Cout, int kh, int kw, int stride_h, int stride_w,
                                                      int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                                      int deformable_groups, long long offset_batch_stride,
                                                      long long mask_batch_stride, int mask_sigmoid, float* scratch_nhwc,
                                                      int offset_pixel_stride, cudaStream_t st) {
    if (!input || !weight || !offset || !output || !scratch_nhwc) return FCVSR_ERR_ARG;
    if (B <= 0 || groups <= 0 || deformable_groups <= 0 || Cin % groups || Cout % groups || Cin % deformable_groups)
        return FCVSR_ERR_ARG;
    if (groups != 1 || (Cin & 31) || ((Cin / deformable_groups) & 3) || (Cout & 15) || Cout > 256 || Cout < 16)
        return FCVSR_ERR_UNSUPPORTED;
    if ((uintptr_t)scratch_nhwc & 15) return FCVSR_ERR_ARG;
    DcnTcArgs a;
    a.xt = scratch_nhwc; a.w = weight; a.bias = bias; a.offset = offset; a.mask = mask; a.y = output;
    a.B = B; a.Cin = Cin; a.H = H; a.W = W; a.Cout = Cout; a.kh = kh; a.kw = kw; a.sh = stride_h; a.sw = stride_w;
    a.ph = pad_h; a.pw = pad_w; a.dh = dil_h; a.dw = dil_w; a.dg = deformable_groups;
    a.Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
    a.Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
    if (a.Ho <= 0 || a.Wo <= 0) return FCVSR_ERR_ARG;
    // 32-bit element offsets inside one image of x / offset / mask and inside the weight tensor
    if ((long long)H * W * Cin > 0x7fffffffLL || (long long)deformable_groups * 2 * kh * kw * a.Ho * a.Wo > 0x7fffffffLL ||
        (long long)Cout * Cin * kh * kw > 0x7fffffffLL)
        return FCVSR_ERR_UNSUPPORTED;
    a.off_bs = offset_batch_stride > 0 ? offset_batch_stride : (long long)deformable_groups * 2 * kh * kw * a.Ho * a.Wo;
    a.mask_bs = mask_batch_stride > 0 ? mask_batch_stride : (long long)deformable_groups * kh * kw * a.Ho * a.Wo;
    a.mask_sigmoid = mask_sigmoid;
    if (offset_pixel_stride > 0) {          // NHWC offsets / mask: channel c of output pixel p at p * stride + c
        if ((long long)a.Ho * a.Wo * offset_pixel_stride > 0x7fffffffLL) return FCVSR_ERR_UNSUPPORTED;
        a.off_cs = 1; a.off_ps = offset_pixel_stride;
        if (offset_batch_stride <= 0) a.off_bs = (long long)a.Ho * a.Wo * offset_pixel_stride;
        if (mask_batch_stride <= 0) a.mask_bs = a.off_bs;
    } else {
        a.off_cs = a.Ho * a.Wo; a.off_ps = 1;
    }
    a.tiles_x = (a.Wo + DT_TW - 1) / DT_TW;
    a.tiles_per_img = a.tiles_x * ((a.Ho + DT_TH - 1) / DT_TH);
    a.tiles = a.tiles_per_img * B;
    // four stages (two per producer group): the rest of the SM's 256 KB stays L1 for the gathers
    a.stage_bytes = DT_M * 128 + ((Cout * 128 + 1023) & ~1023);
    a.nstage = 4;
    if ((size_t)a.nstage * a.stage_bytes > DT_SMEM_MAX - 2048) return FCVSR_ERR_UNSUPPORTED;
    static int* err = nullptr;
    static int num_sms = 0;
    if (!err) {
        if (cudaMalloc(&err, sizeof(int)) != cudaSuccess) return FCVSR_ERR_CUDA;
        cudaMemset(err, 0, sizeof(int));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(dcn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_MAX) != cudaSuccess)
            return FCVSR_ERR_CUDA;
    }
    a.err = err;
    const size_t smem = 1024 + (size_t)a.nstage * a.stage_bytes + 512;
    const int grid = a.tiles < num_sms ? a.tiles : num_sms;
    // input == scratch_nhwc: the caller already holds the pixel-major copy (ModulatedDeformConvPack shares it with its
    // conv_offset_mask convolution) and the transposition pre-pass is skipped
    if (input != scratch_nhwc)
        nchw_to_nhwc_kernel<<<dim3((H * W + 31) / 32, Cin / 32, B), dim3(32, 8), 0, st>>>(input, scratch_nhwc, Cin, H * W, 0);
    dcn_tc_kernel<<<grid, DT_THREADS, smem, st>>>(a);
    return fcvsr_launch_status();
}

// Deformable convolution (DCNv1 / DCNv2 "modulated") backward, NCHW fp32, fused: no column buffer in HBM.
//
// Replaces the reference operator's backward chain
//   modulated_deform_conv_cuda_backward            ops/dcn/src/deform_conv_cuda.cpp:566-700
//     columns = weight^T x grad_output              (addmm per group)
//     modulated_deformable_col2im_coord_gpu_kernel  ops/dcn/src/deform_conv_cuda_kernel.cu:695-766  -> grad_offset, grad_mask
//     modulated_deformable_col2im_gpu_kernel        .cu:635-692                                     -> grad_input (atomics)
//     modulated_deformable_im2col + addmm           -> grad_weight;  grad_bias = sum of grad_output
//   deform_conv_backward_input_cuda / deform_conv_backward_parameters_cuda (DCNv1, .cpp:260-484; kernels .cu:279-435)
// with two kernels:
//   dcn_backward_data_kernel    a register-tiled GEMM (reduction over the output channels) produces a 64-pixel x 64-k tile
//                               of d(loss)/d(columns) in registers; its epilogue applies the bilinear-sampling adjoint
//                               directly: 4 corner atomics into grad_input, and the offset / mask gradients summed over
//                               the channels a thread owns before one atomic per (tap, deformable group, pixel).
//   dcn_backward_weight_kernel  re-samples the (modulated) column tile into shared memory and contracts it with
//                               grad_output over a slice of the B*Ho*Wo pixels; slices are combined with atomics.
// The K index is walked tap-major (k' = tap * Cin_g + c) so that the consecutive k' of one thread share the tap and, for
// (Cin / deformable_groups) % 4 == 0, the deformable group: the sample geometry is computed once per thread and pixel.
// Derivatives follow dmcn_get_gradient_weight (.cu:499-523) and dmcn_get_coordinate_weight (.cu:526-567): a sample
// outside (-1, H) x (-1, W) has zero gradient everywhere, out-of-image corners contribute nothing.
// NHWC fast path (caller passes `scratch`, Cin_g % 4 == 0 and (Cin / deformable_groups) % 4 == 0): the input is transposed to
// NHWC and grad_input is accumulated in an NHWC buffer, so the four channels a thread owns are ONE 16-byte corner: four
// float4 loads instead of 16 scalar ones and four `red.global.add.v4.f32` (sm_90+ vector reductions) instead of 16 scalar
// atomics per (pixel, tap, 4 channels); a transpose-add pass returns grad_input to NCHW.
// All five outputs ACCUMULATE (the caller zero-fills them, as deform_conv.py:155-159 does with zeros_like).
#include "common.cuh"

struct DcnBwdArgs {
    const float* x; const float* w; const float* offset; const float* mask; const float* gy;
    float* gx; float* gw; float* goff; float* gmask;
    const float* xt; float* gxt;     // NHWC copy of x / NHWC accumulator of grad_input (fast path), else NULL
    int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, Ho, Wo;
    long long chunk;             // pixels (of the flattened B*Ho*Wo range) per weight-kernel slice
};

struct DcnGeo { int h0, w0; float lh, lw, m; bool valid; };

__device__ __forceinline__ void dcn_geo(const DcnBwdArgs& a, int b, int g, int tap, int p, int ho, int wo, DcnGeo& q) {
    const int kk2 = a.kh * a.kw, P = a.Ho * a.Wo;
    const int i = tap / a.kw, j = tap - i * a.kw;
    const size_t ob = ((size_t)(b * a.dg + g) * 2 * kk2 + 2 * tap) * P + p;
    // offset == NULL: zero offsets, i.e. the plain convolution (backward of the *Pack modules' conv_offset layers)
    const float h = (float)(ho * a.sh - a.ph + i * a.dh) + (a.offset ? a.offset[ob] : 0.f);
    const float w = (float)(wo * a.sw - a.pw + j * a.dw) + (a.offset ? a.offset[ob + P] : 0.f);
    q.valid = (h > -1.f && w > -1.f && h < (float)a.H && w < (float)a.W);
    if (q.valid) {
        const float fh = floorf(h), fw = floorf(w);
        q.h0 = (int)fh; q.w0 = (int)fw; q.lh = h - fh; q.lw = w - fw;
    } else {
        q.h0 = q.w0 = 0; q.lh = q.lw = 0.f;
    }
    q.m = a.mask ? a.mask[((size_t)(b * a.dg + g) * kk2 + tap) * P + p] : 1.f;
}

__device__ __forceinline__ void dcn_corners(const float* __restrict__ img, int H, int W, const DcnGeo& q, float v[4]) {
    const int h1 = q.h0 + 1, w1 = q.w0 + 1;
    v[0] = (q.h0 >= 0 && q.w0 >= 0) ? img[q.h0 * W + q.w0] : 0.f;
    v[1] = (q.h0 >= 0 && w1 <= W - 1) ? img[q.h0 * W + w1] : 0.f;
    v[2] = (h1 <= H - 1 && q.w0 >= 0) ? img[h1 * W + q.w0] : 0.f;
    v[3] = (h1 <= H - 1 && w1 <= W - 1) ? img[h1 * W + w1] : 0.f;
}


__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// four consecutive channels of the four bilinear corners from the NHWC image `img` (already offset to the first channel)
__device__ __forceinline__ void dcn_corners4(const float* __restrict__ img, int H, int W, int Cin, const DcnGeo& q, float4 v[4]) {
    const int h1 = q.h0 + 1, w1 = q.w0 + 1;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = (q.h0 >= 0 && q.w0 >= 0) ? __ldg(reinterpret_cast<const float4*>(img + (size_t)(q.h0 * W + q.w0) * Cin)) : z;
    v[1] = (q.h0 >= 0 && w1 <= W - 1) ? __ldg(reinterpret_cast<const float4*>(img + (size_t)(q.h0 * W + w1) * Cin)) : z;
    v[2] = (h1 <= H - 1 && q.w0 >= 0) ? __ldg(reinterpret_cast<const float4*>(img + (size_t)(h1 * W + q.w0) * Cin)) : z;
    v[3] = (h1 <= H - 1 && w1 <= W - 1) ? __ldg(reinterpret_cast<const float4*>(img + (size_t)(h1 * W + w1) * Cin)) : z;
}

#define DB_TP 64
#define DB_TK 64
#define DB_TO 16

template <bool NHWC>
__global__ void __launch_bounds__(256) dcn_backward_data_kernel(DcnBwdArgs a) {
    __shared__ __align__(16) float Gs[DB_TO][DB_TP + 4];
    __shared__ __align__(16) float Ws[DB_TO][DB_TK + 4];
    const int tid = threadIdx.x;
    const int P = a.Ho * a.Wo, kk2 = a.kh * a.kw;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups, K = cin_g * kk2;
    const int ntile_k = (K + DB_TK - 1) / DB_TK;
    const int grp = blockIdx.y / ntile_k, k0 = (blockIdx.y % ntile_k) * DB_TK;
    const int p0 = blockIdx.x * DB_TP, b = blockIdx.z;
    const int ch_per_dg = a.Cin / a.dg;
    const int pg = tid >> 4, ng = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lo = (tid >> 6) * 4, lp = tid & 63;        // grad_output loader: 4 rows, one pixel
    const int wr = tid >> 4, wk = (tid & 15) * 4;        // weight loader: one row, 4 consecutive k'
    // the four k' of the weight loader (tap-major walk of the reference's [c][tap] order)
    int wsrc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int kq = k0 + wk + u;
        if (kq < K) { const int tap = kq / cin_g, cl = kq - tap * cin_g; wsrc[u] = cl * kk2 + tap; } else wsrc[u] = -1;
    }
    for (int o0 = 0; o0 < cout_g; o0 += DB_TO) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int o = o0 + lo + u, p = p0 + lp;
            Gs[lo + u][lp] = (o < cout_g && p < P) ? a.gy[((size_t)b * a.Cout + grp * cout_g + o) * P + p] : 0.f;
        }
        {
            const int o = o0 + wr;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                Ws[wr][wk + u] = (o < cout_g && wsrc[u] >= 0) ? a.w[(size_t)(grp * cout_g + o) * K + wsrc[u]] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int oo = 0; oo < DB_TO; ++oo) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = Gs[oo][pg * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Ws[oo][ng * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (NHWC) {
        // fast path: the thread's four k' are four consecutive channels of one tap and one deformable group
        const int kq = k0 + ng * 4;
        if (kq >= K) return;
        const int tap = kq / cin_g, cl = kq - tap * cin_g;
        const int c = grp * cin_g + cl, g = c / ch_per_dg;
        const size_t img_off = (size_t)b * a.H * a.W * a.Cin + c;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = p0 + pg * 4 + i;
            if (p >= P) continue;
            const int ho = p / a.Wo, wo = p - ho * a.Wo;
            DcnGeo q;
            dcn_geo(a, b, g, tap, p, ho, wo, q);
            if (!q.valid) continue;
            const float hh = 1.f - q.lh, hw = 1.f - q.lw;
            const int h1 = q.h0 + 1, w1 = q.w0 + 1;
            const float t0 = acc[i][0] * q.m, t1 = acc[i][1] * q.m, t2 = acc[i][2] * q.m, t3 = acc[i][3] * q.m;
            if (a.gxt) {
                float* gimg = a.gxt + img_off;
                float wgt;
                if (q.h0 >= 0 && q.w0 >= 0) { wgt = hh * hw; red_add_v4(gimg + (size_t)(q.h0 * a.W + q.w0) * a.Cin, t0 * wgt, t1 * wgt, t2 * wgt, t3 * wgt); }
                if (q.h0 >= 0 && w1 <= a.W - 1) { wgt = hh * q.lw; red_add_v4(gimg + (size_t)(q.h0 * a.W + w1) * a.Cin, t0 * wgt, t1 * wgt, t2 * wgt, t3 * wgt); }
                if (h1 <= a.H - 1 && q.w0 >= 0) { wgt = q.lh * hw; red_add_v4(gimg + (size_t)(h1 * a.W + q.w0) * a.Cin, t0 * wgt, t1 * wgt, t2 * wgt, t3 * wgt); }
                if (h1 <= a.H - 1 && w1 <= a.W - 1) { wgt = q.lh * q.lw; red_add_v4(gimg + (size_t)(h1 * a.W + w1) * a.Cin, t0 * wgt, t1 * wgt, t2 * wgt, t3 * wgt); }
            }
            if (a.goff || a.gmask) {
                float4 v[4];
                dcn_corners4(a.xt + img_off, a.H, a.W, a.Cin, q, v);
                const float dhx = hw * (v[2].x - v[0].x) + q.lw * (v[3].x - v[1].x), dhy = hw * (v[2].y - v[0].y) + q.lw * (v[3].y - v[1].y);
                const float dhz = hw * (v[2].z - v[0].z) + q.lw * (v[3].z - v[1].z), dhw = hw * (v[2].w - v[0].w) + q.lw * (v[3].w - v[1].w);
                const float dwx = hh * (v[1].x - v[0].x) + q.lh * (v[3].x - v[2].x), dwy = hh * (v[1].y - v[0].y) + q.lh * (v[3].y - v[2].y);
                const float dwz = hh * (v[1].z - v[0].z) + q.lh * (v[3].z - v[2].z), dww = hh * (v[1].w - v[0].w) + q.lh * (v[3].w - v[2].w);
                const float w00 = hh * hw, w01 = hh * q.lw, w10 = q.lh * hw, w11 = q.lh * q.lw;
                const float sx = w00 * v[0].x + w01 * v[1].x + w10 * v[2].x + w11 * v[3].x, sy = w00 * v[0].y + w01 * v[1].y + w10 * v[2].y + w11 * v[3].y;
                const float sz = w00 * v[0].z + w01 * v[1].z + w10 * v[2].z + w11 * v[3].z, sw = w00 * v[0].w + w01 * v[1].w + w10 * v[2].w + w11 * v[3].w;
                const size_t ob = ((size_t)(b * a.dg + g) * 2 * kk2 + 2 * tap) * P + p;
                if (a.goff) {
                    atomicAdd(a.goff + ob, t0 * dhx + t1 * dhy + t2 * dhz + t3 * dhw);
                    atomicAdd(a.goff + ob + P, t0 * dwx + t1 * dwy + t2 * dwz + t3 * dww);
                }
                if (a.gmask)
                    atomicAdd(a.gmask + ((size_t)(b * a.dg + g) * kk2 + tap) * P + p,
                              acc[i][0] * sx + acc[i][1] * sy + acc[i][2] * sz + acc[i][3] * sw);
            }
        }
        return;
    }
    // epilogue: adjoint of the modulated bilinear sampling for the thread's 4 pixels x 4 k'
    const size_t HW = (size_t)a.H * a.W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + pg * 4 + i;
        if (p >= P) continue;
        const int ho = p / a.Wo, wo = p - ho * a.Wo;
        int cur_tap = -1, cur_g = -1;
        DcnGeo q;
        q.valid = false; q.h0 = q.w0 = 0; q.lh = q.lw = 0.f; q.m = 1.f;
        float goh = 0.f, gow = 0.f, gm = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kq = k0 + ng * 4 + j;
            if (kq >= K) break;
            const int tap = kq / cin_g, cl = kq - tap * cin_g;
            const int c = grp * cin_g + cl, g = c / ch_per_dg;
            if (tap != cur_tap || g != cur_g) {
                if (cur_tap >= 0 && q.valid) {
                    const size_t ob = ((size_t)(b * a.dg + cur_g) * 2 * kk2 + 2 * cur_tap) * P + p;
                    if (a.goff) { atomicAdd(a.goff + ob, goh); atomicAdd(a.goff + ob + P, gow); }
                    if (a.gmask) atomicAdd(a.gmask + ((size_t)(b * a.dg + cur_g) * kk2 + cur_tap) * P + p, gm);
                }
                dcn_geo(a, b, g, tap, p, ho, wo, q);
                cur_tap = tap; cur_g = g; goh = gow = gm = 0.f;
            }
            if (!q.valid) continue;
            const float gc = acc[i][j], t = gc * q.m;
            const float hh = 1.f - q.lh, hw = 1.f - q.lw;
            const int h1 = q.h0 + 1, w1 = q.w0 + 1;
            if (a.gx) {
                float* gimg = a.gx + ((size_t)b * a.Cin + c) * HW;
                if (q.h0 >= 0 && q.w0 >= 0) atomicAdd(gimg + q.h0 * a.W + q.w0, t * hh * hw);
                if (q.h0 >= 0 && w1 <= a.W - 1) atomicAdd(gimg + q.h0 * a.W + w1, t * hh * q.lw);
                if (h1 <= a.H - 1 && q.w0 >= 0) atomicAdd(gimg + h1 * a.W + q.w0, t * q.lh * hw);
                if (h1 <= a.H - 1 && w1 <= a.W - 1) atomicAdd(gimg + h1 * a.W + w1, t * q.lh * q.lw);
            }
            if (a.goff || a.gmask) {
                float v[4];
                dcn_corners(a.x + ((size_t)b * a.Cin + c) * HW, a.H, a.W, q, v);
                goh += t * (hw * (v[2] - v[0]) + q.lw * (v[3] - v[1]));
                gow += t * (hh * (v[1] - v[0]) + q.lh * (v[3] - v[2]));
                gm += gc * (hh * hw * v[0] + hh * q.lw * v[1] + q.lh * hw * v[2] + q.lh * q.lw * v[3]);
            }
        }
        if (cur_tap >= 0 && q.valid) {
            const size_t ob = ((size_t)(b * a.dg + cur_g) * 2 * kk2 + 2 * cur_tap) * P + p;
            if (a.goff) { atomicAdd(a.goff + ob, goh); atomicAdd(a.goff + ob + P, gow); }
            if (a.gmask) atomicAdd(a.gmask + ((size_t)(b * a.dg + cur_g) * kk2 + cur_tap) * P + p, gm);
        }
    }
}

#define DW_TO 64
#define DW_TK 64
#define DW_TP 16

template <bool NHWC>
__global__ void __launch_bounds__(256) dcn_backward_weight_kernel(DcnBwdArgs a) {
    __shared__ __align__(16) float Gs[DW_TP][DW_TO + 4];
    __shared__ __align__(16) float As[DW_TP][DW_TK + 4];
    const int tid = threadIdx.x;
    const int P = a.Ho * a.Wo, kk2 = a.kh * a.kw;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups, K = cin_g * kk2;
    const int ntile_o = (cout_g + DW_TO - 1) / DW_TO;
    const int k0 = blockIdx.x * DW_TK;
    const int grp = blockIdx.y / ntile_o, o0 = (blockIdx.y % ntile_o) * DW_TO;
    const int ch_per_dg = a.Cin / a.dg;
    const long long total = (long long)a.B * P;
    const long long q_begin = (long long)blockIdx.z * a.chunk;
    const long long q_end = q_begin + a.chunk < total ? q_begin + a.chunk : total;
    const int og = tid >> 4, kg = tid & 15;
    const int lp = tid & 15, lq = (tid >> 4) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // the sampler's four k' (fixed for the whole block)
    int s_tap[4], s_c[4], s_g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int kq = k0 + lq + u;
        if (kq < K) {
            s_tap[u] = kq / cin_g; s_c[u] = grp * cin_g + (kq - s_tap[u] * cin_g); s_g[u] = s_c[u] / ch_per_dg;
        } else { s_tap[u] = -1; s_c[u] = 0; s_g[u] = 0; }
    }
    const size_t HW = (size_t)a.H * a.W;
    for (long long q0 = q_begin; q0 < q_end; q0 += DW_TP) {
        const long long qq = q0 + lp;
        const bool ok = qq < q_end;
        const int b = ok ? (int)(qq / P) : 0;
        const int p = ok ? (int)(qq - (long long)b * P) : 0;
        const int ho = p / a.Wo, wo = p - ho * a.Wo;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int o = o0 + lq + u;
            Gs[lp][lq + u] = (ok && o < cout_g) ? a.gy[((size_t)b * a.Cout + grp * cout_g + o) * P + p] : 0.f;
        }
        if (NHWC) {
            // the sampler's four k' are four consecutive channels of one tap and one deformable group: one geometry, 4 x float4
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && s_tap[0] >= 0) {
                DcnGeo q;
                dcn_geo(a, b, s_g[0], s_tap[0], p, ho, wo, q);
                if (q.valid) {
                    float4 cv[4];
                    dcn_corners4(a.xt + (size_t)b * HW * a.Cin + s_c[0], a.H, a.W, a.Cin, q, cv);
                    const float w00 = (1.f - q.lh) * (1.f - q.lw) * q.m, w01 = (1.f - q.lh) * q.lw * q.m;
                    const float w10 = q.lh * (1.f - q.lw) * q.m, w11 = q.lh * q.lw * q.m;
                    o.x = w00 * cv[0].x + w01 * cv[1].x + w10 * cv[2].x + w11 * cv[3].x;
                    o.y = w00 * cv[0].y + w01 * cv[1].y + w10 * cv[2].y + w11 * cv[3].y;
                    o.z = w00 * cv[0].z + w01 * cv[1].z + w10 * cv[2].z + w11 * cv[3].z;
                    o.w = w00 * cv[0].w + w01 * cv[1].w + w10 * cv[2].w + w11 * cv[3].w;
                }
            }
            *reinterpret_cast<float4*>(&As[lp][lq]) = o;
        } else {
        int cur_tap = -1, cur_g = -1;
        DcnGeo q;
        q.valid = false; q.h0 = q.w0 = 0; q.lh = q.lw = 0.f; q.m = 1.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v = 0.f;
            if (ok && s_tap[u] >= 0) {
                if (s_tap[u] != cur_tap || s_g[u] != cur_g) {
                    dcn_geo(a, b, s_g[u], s_tap[u], p, ho, wo, q);
                    cur_tap = s_tap[u]; cur_g = s_g[u];
                }
                if (q.valid) {
                    float cv[4];
                    dcn_corners(a.x + ((size_t)b * a.Cin + s_c[u]) * HW, a.H, a.W, q, cv);
                    const float hh = 1.f - q.lh, hw = 1.f - q.lw;
                    v = (hh * hw * cv[0] + hh * q.lw * cv[1] + q.lh * hw * cv[2] + q.lh * q.lw * cv[3]) * q.m;
                }
            }
            As[lp][lq + u] = v;
        }
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < DW_TP; ++pp) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = Gs[pp][og * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = As[pp][kg * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = o0 + og * 4 + i;
        if (o >= cout_g) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kq = k0 + kg * 4 + j;
            if (kq >= K) continue;
            const int tap = kq / cin_g, cl = kq - tap * cin_g;
            atomicAdd(a.gw + (size_t)(grp * cout_g + o) * K + cl * kk2 + tap, acc[i][j]);
        }
    }
}

// grad_bias[co] += sum over batch and pixels of grad_output (deform_conv_cuda.cpp:688-692): one block per channel
__global__ void __launch_bounds__(256) dcn_backward_bias_kernel(const float* __restrict__ gy, float* gb, int B, int Cout, int P) {
    __shared__ float part[8];
    const int co = blockIdx.x;
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* src = gy + ((size_t)b * Cout + co) * P;
        for (int p = threadIdx.x; p < P; p += 256) s += src[p];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i];
        gb[co] += t;
    }
}

// NCHW [B][C][P] -> NHWC [B][P][C] through a 32x33 shared tile (coalesced on both sides)
__global__ void dcn_bwd_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int P) {
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
        if (p < P && c < C) yb[(size_t)p * C + c] = tile[threadIdx.x][i];
    }
}

// grad_input[b][c][p] += gxt[b][p][c]
__global__ void dcn_bwd_from_nhwc_add_kernel(const float* __restrict__ gxt, float* __restrict__ gx, int C, int P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* sb = gxt + (size_t)b * C * P;
    float* db = gx + (size_t)b * C * P;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < P && c < C) ? sb[(size_t)p * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < P) db[(size_t)c * P + p] += tile[threadIdx.x][i];
    }
}

extern "C" int fcvsr_modulated_deform_conv_backward(const float* input, const float* weight, const float* offset,
                                                    const float* mask, const float* grad_output, float* grad_input,
                                                    float* grad_weight, float* grad_bias, float* grad_offset,
                                                    float* grad_mask, int B, int Cin, int H, int W, int Cout, int kh, int kw,
                                                    int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                                    int groups, int deformable_groups, float* scratch, cudaStream_t st) {
    if (!input || !weight || !grad_output) return FCVSR_ERR_ARG;
    if (B <= 0 || groups <= 0 || deformable_groups <= 0 || Cin % groups || Cout % groups || Cin % deformable_groups)
        return FCVSR_ERR_ARG;
    if ((grad_mask && !mask) || (grad_offset && !offset)) return FCVSR_ERR_ARG;
    DcnBwdArgs a;
    a.x = input; a.w = weight; a.offset = offset; a.mask = mask; a.gy = grad_output;
    a.gx = grad_input; a.gw = grad_weight; a.goff = grad_offset; a.gmask = grad_mask;
    a.B = B; a.Cin = Cin; a.H = H; a.W = W; a.Cout = Cout; a.kh = kh; a.kw = kw; a.sh = stride_h; a.sw = stride_w;
    a.ph = pad_h; a.pw = pad_w; a.dh = dil_h; a.dw = dil_w; a.groups = groups; a.dg = deformable_groups;
    a.Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
    a.Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
    if (a.Ho <= 0 || a.Wo <= 0) return FCVSR_ERR_ARG;
    a.chunk = 0;
    a.xt = nullptr; a.gxt = nullptr;
    const int P = a.Ho * a.Wo, cin_g = Cin / groups, cout_g = Cout / groups, K = cin_g * kh * kw;
    const size_t n_in = (size_t)B * Cin * H * W;
    const bool nhwc = scratch && !((uintptr_t)scratch & 15) && (cin_g & 3) == 0 && ((Cin / deformable_groups) & 3) == 0;
    const dim3 tgrid((H * W + 31) / 32, (Cin + 31) / 32, B), tblock(32, 8);
    if (nhwc) {
        dcn_bwd_to_nhwc_kernel<<<tgrid, tblock, 0, st>>>(input, scratch, Cin, H * W);
        a.xt = scratch;
        if (grad_input) {
            a.gxt = scratch + n_in;
            if (cudaMemsetAsync(a.gxt, 0, n_in * sizeof(float), st) != cudaSuccess) return FCVSR_ERR_CUDA;
        }
    }
    if (grad_input || grad_offset || grad_mask) {
        dim3 grid((P + DB_TP - 1) / DB_TP, groups * ((K + DB_TK - 1) / DB_TK), B);
        if (nhwc) {
            dcn_backward_data_kernel<true><<<grid, 256, 0, st>>>(a);
            if (grad_input) dcn_bwd_from_nhwc_add_kernel<<<tgrid, tblock, 0, st>>>(a.gxt, grad_input, Cin, H * W);
        } else {
            dcn_backward_data_kernel<false><<<grid, 256, 0, st>>>(a);
        }
    }
    if (grad_weight) {
        const int tiles = ((K + DW_TK - 1) / DW_TK) * groups * ((cout_g + DW_TO - 1) / DW_TO);
        const long long total = (long long)B * P;
        long long nsplit = (2 * 148 + tiles - 1) / tiles;                  // about two blocks per SM
        const long long max_split = (total + 4 * DW_TP - 1) / (4 * DW_TP);  // at least 64 pixels per slice
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
        long long chunk = (total + nsplit - 1) / nsplit;
        chunk = (chunk + DW_TP - 1) / DW_TP * DW_TP;
        nsplit = (total + chunk - 1) / chunk;
        a.chunk = chunk;
        dim3 grid((K + DW_TK - 1) / DW_TK, groups * ((cout_g + DW_TO - 1) / DW_TO), (unsigned)nsplit);
        if (nhwc) dcn_backward_weight_kernel<true><<<grid, 256, 0, st>>>(a);
        else dcn_backward_weight_kernel<false><<<grid, 256, 0, st>>>(a);
    }
    if (grad_bias) dcn_backward_bias_kernel<<<Cout, 256, 0, st>>>(grad_output, grad_bias, B, Cout, P);
    return fcvsr_launch_status();
}

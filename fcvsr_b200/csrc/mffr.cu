// Multi-frequency refinement kernels (MultiFreq_Refinment, CVSR_freq.py:2183-2254).
//
// The band split itself (Split_freq :2075-2101) is done by the FFT kernels with the symmetrised
// band mask fused into the inverse H pass (fft.cu).  This file holds the sequential DivEnh chain
// (:2104-2133) and the CALayer gates (:1812-1828):
//
//   step i:   o  = x_i - Sb + 0.2 So          (i = 0:  o = x_0 - mean_hw(x_0))
//             t1 = 0.2 a o x_i + b x_i,   t2 = 0.2 a So x_i + b x_i   (t2 only for i > 0)
//             out_i = t1 * g(mean t1) + t2 * g(mean t2),   g = sigmoid(W2 relu(W1 .))
//             Sb += x_i,  So += out_i
//   final:    y = So * g(mean So) + x
//
// Each global mean is a grid-wide dependency, so one `divenh_step` launch applies step i-1 (now that
// its gates are known) and reduces step i in the same pass over the pixels; reductions are
// deterministic (per-block partials summed in fixed order by `reduce_finalize`).
#include "common.cuh"

#define MF_C 64
#define MF_PIX_PER_BLOCK 256

struct DivEnhArgs {
    // apply part (previous step), enabled when x_prev != nullptr
    const float* x_prev; const float* a_prev; const float* b_prev;
    const float* mean_prev;            // [B,64], only for prev step 0
    const float* gate_prev;            // [B,2,64]
    int prev_is_first;
    // reduce part (current step), enabled when mode != 0: 1 = DivEnh step, 2 = plain sum of So
    int mode; int cur_is_first;
    const float* x_cur; const float* a_cur; const float* b_cur; const float* mean_cur;
    float* sb; float* so;              // running sums [B,P,64] (read/write)
    float* partial;                    // [B][nblk][128]
    int P;
};

__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }

// A lane owns 4 channels (float4), a half-warp one pixel; a warp walks 32 consecutive pixels, 8 per iteration, with every
// load of the iteration issued before the first store (sb / so are read and written through the same pointers, so the
// compiler cannot overlap iterations by itself: one pixel per iteration ran at 3.4 TB/s).
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__global__ void __launch_bounds__(256) divenh_step_kernel(DivEnhArgs g) {
    __shared__ float red[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int c = 4 * (lane & 15), sub = lane >> 4;
    const size_t base = (size_t)b * g.P;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ap = z4, bp = z4, mp = z4, g1 = z4, g2 = z4, ac = z4, bc = z4, mc = z4;
    if (g.x_prev) {
        ap = ld4(g.a_prev + c);
        bp = ld4(g.b_prev + c);
        if (g.prev_is_first) mp = ld4(g.mean_prev + b * 128 + c);
        g1 = ld4(g.gate_prev + (b * 2 + 0) * MF_C + c);
        g2 = ld4(g.gate_prev + (b * 2 + 1) * MF_C + c);
    }
    if (g.mode == 1) {
        ac = ld4(g.a_cur + c);
        bc = ld4(g.b_cur + c);
        if (g.cur_is_first) mc = ld4(g.mean_cur + b * 128 + c);
    }
    const float apv[4] = {ap.x, ap.y, ap.z, ap.w}, bpv[4] = {bp.x, bp.y, bp.z, bp.w}, mpv[4] = {mp.x, mp.y, mp.z, mp.w};
    const float g1v[4] = {g1.x, g1.y, g1.z, g1.w}, g2v[4] = {g2.x, g2.y, g2.z, g2.w};
    const float acv[4] = {ac.x, ac.y, ac.z, ac.w}, bcv[4] = {bc.x, bc.y, bc.z, bc.w}, mcv[4] = {mc.x, mc.y, mc.z, mc.w};
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const int p_begin = blockIdx.x * MF_PIX_PER_BLOCK + warp * 32;
    const int p_stop = min(g.P, p_begin + 32);
    const bool load_sums = g.x_prev && !g.prev_is_first;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        float4 xp[4], sbv[4], sov[4], xc[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p_begin + it * 8 + u * 2 + sub;
            ok[u] = p < p_stop;
            const size_t o = (base + p) * MF_C + c;
            xp[u] = (ok[u] && g.x_prev) ? ld4(g.x_prev + o) : z4;
            sbv[u] = (ok[u] && load_sums) ? ld4(g.sb + o) : z4;
            sov[u] = (ok[u] && load_sums) ? ld4(g.so + o) : z4;
            xc[u] = (ok[u] && g.mode == 1) ? ld4(g.x_cur + o) : z4;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            const int p = p_begin + it * 8 + u * 2 + sub;
            const size_t o = (base + p) * MF_C + c;
            float x[4] = {xp[u].x, xp[u].y, xp[u].z, xp[u].w};
            float sb[4] = {sbv[u].x, sbv[u].y, sbv[u].z, sbv[u].w}, so[4] = {sov[u].x, sov[u].y, sov[u].z, sov[u].w};
            const float xq[4] = {xc[u].x, xc[u].y, xc[u].z, xc[u].w};
            if (g.x_prev) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (g.prev_is_first) {
                        const float t1 = 0.2f * apv[k] * (x[k] - mpv[k]) * x[k] + bpv[k] * x[k];
                        sb[k] = x[k];
                        so[k] = t1 * g1v[k];
                    } else {
                        const float oo = x[k] - sb[k] + 0.2f * so[k];
                        const float t1 = 0.2f * apv[k] * oo * x[k] + bpv[k] * x[k];
                        const float t2 = 0.2f * apv[k] * so[k] * x[k] + bpv[k] * x[k];
                        const float out = t1 * g1v[k] + t2 * g2v[k];
                        sb[k] += x[k];
                        so[k] += out;
                    }
                }
                *reinterpret_cast<float4*>(g.sb + o) = make_float4(sb[0], sb[1], sb[2], sb[3]);
                *reinterpret_cast<float4*>(g.so + o) = make_float4(so[0], so[1], so[2], so[3]);
            }
            if (g.mode == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (g.cur_is_first) {
                        s1[k] += 0.2f * acv[k] * (xq[k] - mcv[k]) * xq[k] + bcv[k] * xq[k];
                    } else {
                        const float oo = xq[k] - sb[k] + 0.2f * so[k];
                        s1[k] += 0.2f * acv[k] * oo * xq[k] + bcv[k] * xq[k];
                        s2[k] += 0.2f * acv[k] * so[k] * xq[k] + bcv[k] * xq[k];
                    }
                }
            } else if (g.mode == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s1[k] += so[k];
            }
        }
    }
    if (g.mode) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {       // the two pixels of the warp (half-warps), fixed order
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 16);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 16);
        }
        if (sub == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { red[warp][c + k] = s1[k]; red[warp][64 + c + k] = s2[k]; }
        }
        __syncthreads();
        if (threadIdx.x < 128) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
            g.partial[((size_t)b * gridDim.x + blockIdx.x) * 128 + threadIdx.x] = s;
        }
    }
}

extern "C" int fcvsr_divenh_step(const float* x_prev, const float* a_prev, const float* b_prev, const float* mean_prev,
                                 const float* gate_prev, int prev_is_first, int mode, int cur_is_first,
                                 const float* x_cur, const float* a_cur, const float* b_cur, const float* mean_cur,
                                 float* sb, float* so, float* partial, int B, int P, cudaStream_t st) {
    if (!sb || !so || (mode && !partial) || (mode == 1 && (!x_cur || !a_cur || !b_cur))) return FCVSR_ERR_ARG;
    DivEnhArgs g;
    g.x_prev = x_prev; g.a_prev = a_prev; g.b_prev = b_prev; g.mean_prev = mean_prev; g.gate_prev = gate_prev;
    g.prev_is_first = prev_is_first; g.mode = mode; g.cur_is_first = cur_is_first;
    g.x_cur = x_cur; g.a_cur = a_cur; g.b_cur = b_cur; g.mean_cur = mean_cur;
    g.sb = sb; g.so = so; g.partial = partial; g.P = P;
    dim3 grid((P + MF_PIX_PER_BLOCK - 1) / MF_PIX_PER_BLOCK, B);
    divenh_step_kernel<<<grid, 256, 0, st>>>(g);
    return fcvsr_launch_status();
}

// Plain per-channel sums of a 64-channel NHWC tensor -> partial[B][nblk][128] (first 64 used).
__global__ void __launch_bounds__(256) chansum64_kernel(const float* __restrict__ x, int ldx, float* __restrict__ partial, int P) {
    __shared__ float red[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y, c = 2 * lane;
    float2 s = make_float2(0, 0);
    const int p_end = min(P, (int)(blockIdx.x + 1) * MF_PIX_PER_BLOCK);
    for (int p = blockIdx.x * MF_PIX_PER_BLOCK + warp; p < p_end; p += 8) {
        const float2 v = *reinterpret_cast<const float2*>(x + ((size_t)b * P + p) * ldx + c);
        s.x += v.x; s.y += v.y;
    }
    red[warp][c] = s.x; red[warp][c + 1] = s.y;
    __syncthreads();
    if (threadIdx.x < 128) {
        float t = 0.f;
        if (threadIdx.x < 64)
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        partial[((size_t)b * gridDim.x + blockIdx.x) * 128 + threadIdx.x] = t;
    }
}

extern "C" int fcvsr_chansum64(const float* x, int ldx, float* partial, int B, int P, cudaStream_t st) {
    if (!x || !partial || (ldx & 1)) return FCVSR_ERR_ARG;
    dim3 grid((P + MF_PIX_PER_BLOCK - 1) / MF_PIX_PER_BLOCK, B);
    chansum64_kernel<<<grid, 256, 0, st>>>(x, ldx, partial, P);
    return fcvsr_launch_status();
}

// Sum the per-block partials in fixed order and turn them into means or CALayer gates.
//   mode 0: out[b][v][c] = mean            mode 1: out[b][v][c] = sigmoid(W2 relu(W1 mean))
// partial [B][nblk][128] holds nvec (1 or 2) vectors of 64; W1 [4][64], W2 [64][4] (reduction 16).
__global__ void __launch_bounds__(512) reduce_finalize_kernel(const float* __restrict__ partial, int nblk, int nvec,
                                                              float inv_count, int mode, const float* __restrict__ w1,
                                                              const float* __restrict__ w2, float* __restrict__ out) {
    __shared__ float part[4][128];
    __shared__ float mean[128];
    __shared__ float hid[2][4];
    const int b = blockIdx.x, t = threadIdx.x & 127, grp = threadIdx.x >> 7;
    // pure latency (B CTAs on the whole GPU): four thread groups walk interleaved quarters of the partial list with
    // independent loads, combined in fixed order
    float s = 0.f;
    if (t < nvec * 64)
        for (int k = grp; k < nblk; k += 4) s += partial[((size_t)b * nblk + k) * 128 + t];
    part[grp][t] = s;
    __syncthreads();
    if (grp) return;
    mean[t] = (part[0][t] + part[1][t] + part[2][t] + part[3][t]) * inv_count;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (mode == 0) {
        if (t < nvec * 64) out[(size_t)b * 128 + t] = mean[t];
        return;
    }
    if (t < nvec * 4) {
        const int v = t >> 2, h = t & 3;
        float a = 0.f;
        for (int c = 0; c < 64; ++c) a += w1[h * 64 + c] * mean[v * 64 + c];
        hid[v][h] = fmaxf(a, 0.f);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t < nvec * 64) {
        const int v = t >> 6, c = t & 63;
        float a = 0.f;
#pragma unroll
        for (int h = 0; h < 4; ++h) a += w2[c * 4 + h] * hid[v][h];
        out[(size_t)b * 128 + t] = 1.f / (1.f + __expf(-a));
    }
}

extern "C" int fcvsr_reduce_finalize(const float* partial, int nblk, int nvec, float inv_count, int mode,
                                     const float* w1, const float* w2, float* out, int B, cudaStream_t st) {
    if (!partial || !out || nvec < 1 || nvec > 2 || (mode == 1 && (!w1 || !w2))) return FCVSR_ERR_ARG;
    reduce_finalize_kernel<<<B, 512, 0, st>>>(partial, nblk, nvec, inv_count, mode, w1, w2, out);
    return fcvsr_launch_status();
}

// y = So * gate + x   (MultiFreq_Refinment :2229-2230), gate [B,128] (first 64 used)
__global__ void mffr_final_kernel(const float* __restrict__ so, const float* __restrict__ gate, const float* __restrict__ x,
                                  int ldx, float* __restrict__ y, int ldy, int P, size_t total4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i & 15) * 4;
    const size_t pix = i >> 4;
    const int b = (int)(pix / P);
    const float4 s = *reinterpret_cast<const float4*>(so + pix * 64 + c);
    const float4 g = *reinterpret_cast<const float4*>(gate + (size_t)b * 128 + c);
    const float4 xv = *reinterpret_cast<const float4*>(x + pix * ldx + c);
    *reinterpret_cast<float4*>(y + pix * ldy + c) =
        make_float4(fmaf(s.x, g.x, xv.x), fmaf(s.y, g.y, xv.y), fmaf(s.z, g.z, xv.z), fmaf(s.w, g.w, xv.w));
}

extern "C" int fcvsr_mffr_final(const float* so, const float* gate, const float* x, int ldx, float* y, int ldy, int B,
                                int P, cudaStream_t st) {
    if (!so || !gate || !x || !y || (ldx & 3) || (ldy & 3)) return FCVSR_ERR_ARG;
    const size_t total4 = (size_t)B * P * 16;
    mffr_final_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(so, gate, x, ldx, y, ldy, P, total4);
    return fcvsr_launch_status();
}

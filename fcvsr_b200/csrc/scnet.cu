// Non-GEMM kernels of the SCNetbk trunk (CVSR_freq.py:657-822).
//
//   ctx_block                    ContextBlock (:657-701): softmax-over-HW attention pooling done as an
//                                online-softmax reduction (running max / sum / weighted channel sums
//                                per block, merged in fixed order by the last block to finish) followed by
//                                the 64->64->64 MLP, in one launch.
//   rcb_finish                   RCB tail (:720-724):  r = lrelu_0.2(res + add_term) + r0
//   level_mix                    BlockRCB cross-level sum (:766-777):
//                                x += coef*r + avgpool2(td) + bilinear_x2(tu)
//                                (Interpolate(0.5) of an even-sized map == 2x2 mean; Interpolate(2.0) is
//                                bilinear, align_corners=False; td/tu are the 1x1 down/up conv outputs)
//   level_mix8                   the same sum for the bf16 mode's BlockRCB launch (every side tensor bf16, x carried as the
//                                bf16 operand copy, eight channels per thread)
#include "common.cuh"

// 4 consecutive channels at element index idx of an fp32 tensor, or of a bf16 tensor (b16) with the same element indexing
__device__ __forceinline__ float4 load4_any(const void* base, size_t idx, int b16) {
    if (b16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(base) + idx);
        return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                           __uint_as_float(u.y & 0xffff0000u));
    }
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
}

#define CTX_PIX_PER_BLOCK 128   // granularity of the caller's `partial` buffer (upper bound of the block count)
#define CTX_PPB 256             // pixels per block: 8 warps x 32 pixels per iteration
#define CTX_STRIDE 66           // m, z, acc[64]
#define CTX_LOG2E 1.4426950408889634f

#define SC_MAX_LEV 3
// Every helper below can run the three pyramid levels of a BlockRCB (CVSR_freq.py:766-777) in ONE launch: the levels hold 1, 1/4
// and 1/16 of the pixels, so per-level launches of the small ones are pure latency (8 us for a 6-CTA kernel) and used to run on
// side streams.  A launch gets up to SC_MAX_LEV level descriptors; blocks / threads find their level by index range.
struct CtxLevel { const void* x; int P; int nblk; int blk_begin; long long part_off; };    // part_off: floats into `partial`
struct CtxArgs { CtxLevel lv[SC_MAX_LEV]; int nlev; int ldx; int ppb; int B; const float* wmask; float* partial; };

#define CTX_SEL(field) (l == 0 ? a.lv[0].field : (l == 1 ? a.lv[1].field : a.lv[2].field))

__device__ __forceinline__ float ctx_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool X16>
__device__ __forceinline__ void ctx_unpack(const uint4& r, const float (&v)[8], float (&f)[8]) {
    if (X16) {
        f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
        f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
        f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
        f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = v[j];
    }
}

#define CTX_MAX_NBLK 4096
// ContextBlock in ONE kernel.  A block reduces ppb pixels (a multiple of 256) of one (level, image) to an online-softmax
// partial (running max m, sum z, weighted channel sums acc[64]); the LAST block of a (level, image) to finish -- found with a
// counter that it resets, so the launch can be replayed from a CUDA graph -- merges the partials in index order (deterministic)
// and applies the 64 -> 64 -> 64 MLP.  No second launch: the finalize of one image overlaps the pooling of the others.
//
// Pooling pass (the first version of this kernel was issue-bound at 1.7 TB/s: a lane owned 2 channels, so every pixel cost a
// 5-step shuffle reduction and an online-softmax update with two exponentials): a lane owns 8 channels (one 16-byte load of a
// bf16 tensor), so 8 lanes cover a pixel, a warp load covers 4 pixels and 8 loads are in flight per lane; the logit needs 3
// shuffle steps per FOUR pixels; the soft-max state (m, z, acc) is kept per 8-lane group and rescaled once per 8 pixels (the
// groups are merged once, at the end); logits are pre-scaled by log2(e) so that an exponential is one ex2.approx.
template <bool X16>
__global__ void __launch_bounds__(256, X16 ? 3 : 2) ctx_block_kernel(const CtxArgs a, const float* __restrict__ w1, const float* __restrict__ w2,
                                                           float* __restrict__ add, unsigned* __restrict__ counters,
                                                           float* __restrict__ pool_out) {
    __shared__ float sm_m[8], sm_z[8];
    __shared__ float sm_acc[8][64];
    __shared__ float esc[CTX_MAX_NBLK];
    __shared__ float part[8][64];
    __shared__ float ctx[64], hid[64];
    __shared__ float red[8];
    __shared__ int is_last;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, b = blockIdx.y;
    const int sub = lane >> 3, oct = lane & 7;
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const int P = CTX_SEL(P), ppb = a.ppb, ldx = a.ldx, lbx = bx - CTX_SEL(blk_begin), nblk = CTX_SEL(nblk);
    const long long part_off = CTX_SEL(part_off);
    const void* xv = CTX_SEL(x);
    // conv_mask weights x log2(e) in shared memory (8 registers that the packed-load variant does not have)
    __shared__ __align__(16) float wsm[64];
    if (t < 64) wsm[t] = a.wmask[t] * CTX_LOG2E;
    __syncthreads();
    float m = -INFINITY, z = 0.f;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int per_warp = ppb >> 3;                          // multiple of 32
    const int p0 = lbx * ppb + warp * per_warp + sub;
    const size_t img = (size_t)b * P * ldx + oct * 8;
    for (int it = 0; it < per_warp; it += 32) {
        // X16: the eight 16-byte loads stay packed (32 registers) and are unpacked where they are used, once for the logit and once
        // for the weighted sums -- 16 more ALU instructions per pixel, but three blocks per SM instead of two (the kernel is bound
        // by bytes in flight, not by issue slots)
        uint4 raw[X16 ? 8 : 1];
        float v[X16 ? 1 : 8][8];
        float lg[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = p0 + it + 4 * u;
            if (X16) {
                raw[u] = make_uint4(0u, 0u, 0u, 0u);
                if (p < P) raw[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(xv) + img + (size_t)p * ldx));
            } else {
                float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
                if (p < P) {
                    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xv) + img + (size_t)p * ldx);
                    r0 = __ldg(src); r1 = __ldg(src + 1);
                }
                v[u][0] = r0.x; v[u][1] = r0.y; v[u][2] = r0.z; v[u][3] = r0.w;
                v[u][4] = r1.x; v[u][5] = r1.y; v[u][6] = r1.z; v[u][7] = r1.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float f[8];
            ctx_unpack<X16>(raw[X16 ? u : 0], v[X16 ? 0 : u], f);
            const float4 wa = *reinterpret_cast<const float4*>(wsm + oct * 8), wb = *reinterpret_cast<const float4*>(wsm + oct * 8 + 4);
            lg[u] = fmaf(f[7], wb.w, fmaf(f[6], wb.z, fmaf(f[5], wb.y, fmaf(f[4], wb.x,
                    fmaf(f[3], wa.w, fmaf(f[2], wa.z, fmaf(f[1], wa.y, f[0] * wa.x)))))));
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) lg[u] += __shfl_xor_sync(0xffffffffu, lg[u], o);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (p0 + it + 4 * u >= P) lg[u] = -INFINITY;
            mx = fmaxf(mx, lg[u]);
        }
        const float mn = fmaxf(m, mx);
        const float ms = mn == -INFINITY ? 0.f : mn;        // no valid pixel yet: every exponential below is ex2(-inf) = 0
        const float sc = ctx_ex2(m - ms);
        z *= sc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= sc;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float e = ctx_ex2(lg[u] - ms);
            z += e;
            float f[8];
            ctx_unpack<X16>(raw[X16 ? u : 0], v[X16 ? 0 : u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(e, f[j], acc[j]);
        }
        m = mn;
    }
    // merge the four 8-lane groups of the warp (lanes oct, oct + 8, oct + 16, oct + 24 hold the same channels)
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), zo = __shfl_xor_sync(0xffffffffu, z, o);
        const float mn = fmaxf(m, mo);
        const float ms = mn == -INFINITY ? 0.f : mn;
        const float s1 = ctx_ex2(m - ms), s2 = ctx_ex2(mo - ms);
        z = z * s1 + zo * s2;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = acc[j] * s1 + __shfl_xor_sync(0xffffffffu, acc[j], o) * s2;
        m = mn;
    }
    if (lane == 0) { sm_m[warp] = m; sm_z[warp] = z; }
    if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sm_acc[warp][oct * 8 + j] = acc[j];
    }
    __syncthreads();
    if (t < 64) {
        float M = -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) M = fmaxf(M, sm_m[k]);
        const float Ms = M == -INFINITY ? 0.f : M;
        float Z = 0.f, A = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = ctx_ex2(sm_m[k] - Ms);
            Z += sm_z[k] * s;
            A += sm_acc[k][t] * s;
        }
        float* dst = a.partial + part_off + ((size_t)b * nblk + lbx) * CTX_STRIDE;
        if (t == 0) { dst[0] = M; dst[1] = Z; }
        dst[2 + t] = A;
        __threadfence();
    }
    __syncthreads();
    unsigned* cnt = counters + l * a.B + b;
    if (t == 0) {
        const unsigned prev = atomicAdd(cnt, 1u);
        is_last = prev == (unsigned)(nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- finalize (one block per (level, image)): merge the partials in index order -> context[64] -> add = W2 lrelu_0.2(W1 ctx).
    // A chain of dependent L2 round trips that runs after the last pooling block, so it is kept short: the MLP weights are
    // requested first (thread = (row t >> 2, 16 columns)), (m, z) of a partial come with one 8-byte load and stay in registers
    // when there are at most 256 partials, and the channel sums are walked by 8 thread groups with 8 loads in flight.
    float4 wr1[4], wr2[4];
    if (w1) {
        const float4* a1 = reinterpret_cast<const float4*>(w1 + (t >> 2) * 64 + (t & 3) * 16);
        const float4* a2 = reinterpret_cast<const float4*>(w2 + (t >> 2) * 64 + (t & 3) * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) { wr1[j] = __ldg(a1 + j); wr2[j] = __ldg(a2 + j); }
    }
    const float* pp = a.partial + part_off + (size_t)b * nblk * CTX_STRIDE;
    float2 mz0 = make_float2(-INFINITY, 0.f);
    if (t < nblk) mz0 = __ldcg(reinterpret_cast<const float2*>(pp + t * CTX_STRIDE));
    float M = mz0.x;
    for (int k = t + 256; k < nblk; k += 256) M = fmaxf(M, __ldcg(pp + k * CTX_STRIDE));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    if (lane == 0) red[warp] = M;
    __syncthreads();
    M = red[lane & 7];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    __syncthreads();
    float Z = 0.f;
    if (t < nblk) {
        const float e = ctx_ex2(mz0.x - M);
        esc[t] = e;
        Z = mz0.y * e;
    }
    for (int k = t + 256; k < nblk; k += 256) {
        const float2 mz = __ldcg(reinterpret_cast<const float2*>(pp + k * CTX_STRIDE));
        const float e = ctx_ex2(mz.x - M);
        esc[k] = e;
        Z += mz.y * e;
    }
    // fixed-order (deterministic) tree: lanes, then warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(0xffffffffu, Z, o);
    if (lane == 0) red[warp] = Z;
    __syncthreads();
    Z = red[lane & 7];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) Z += __shfl_xor_sync(0xffffffffu, Z, o);
    {
        const int c2 = (t & 31) * 2, q = t >> 5;          // 8 groups (warps), a lane owns 2 channels
        float2 A = make_float2(0.f, 0.f);
        int k = q;
        for (; k + 56 < nblk; k += 64) {
            float2 av[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = __ldcg(reinterpret_cast<const float2*>(pp + (k + 8 * u) * CTX_STRIDE + 2 + c2));
#pragma unroll
            for (int u = 0; u < 8; ++u) { const float e = esc[k + 8 * u]; A.x = fmaf(av[u].x, e, A.x); A.y = fmaf(av[u].y, e, A.y); }
        }
        for (; k < nblk; k += 8) {
            const float2 av = __ldcg(reinterpret_cast<const float2*>(pp + k * CTX_STRIDE + 2 + c2));
            const float e = esc[k];
            A.x = fmaf(av.x, e, A.x); A.y = fmaf(av.y, e, A.y);
        }
        part[q][c2] = A.x; part[q][c2 + 1] = A.y;
    }
    __syncthreads();
    if (t < 64) {
        float sacc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sacc += part[j][t];
        ctx[t] = sacc / Z;
        if (pool_out) {          // training: the pooled context and the soft-max statistics (log2 domain) for the backward pass
            float* po = pool_out + ((size_t)l * a.B + b) * CTX_STRIDE;
            po[t] = sacc / Z;
            if (t == 0) { po[64] = M; po[65] = Z; }
        }
    }
    if (!w1) {                   // pooling only (the MLP stays with autograd)
        if (t == 0) *cnt = 0u;
        return;
    }
    __syncthreads();
    {   // hid[r] = lrelu_0.2(W1[r,:] . ctx): thread = (row t >> 2, columns (t & 3) * 16 ..), two shuffle steps
        const float* cv = ctx + (t & 3) * 16;
        float h = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            h += wr1[j].x * cv[4 * j] + wr1[j].y * cv[4 * j + 1] + wr1[j].z * cv[4 * j + 2] + wr1[j].w * cv[4 * j + 3];
        h += __shfl_xor_sync(0xffffffffu, h, 1);
        h += __shfl_xor_sync(0xffffffffu, h, 2);
        if ((t & 3) == 0) hid[t >> 2] = h >= 0.f ? h : 0.2f * h;
    }
    __syncthreads();
    {   // add[r] = W2[r,:] . hid
        const float* hv = hid + (t & 3) * 16;
        float o2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o2 += wr2[j].x * hv[4 * j] + wr2[j].y * hv[4 * j + 1] + wr2[j].z * hv[4 * j + 2] + wr2[j].w * hv[4 * j + 3];
        o2 += __shfl_xor_sync(0xffffffffu, o2, 1);
        o2 += __shfl_xor_sync(0xffffffffu, o2, 2);
        if ((t & 3) == 0) add[((size_t)l * a.B + b) * 64 + (t >> 2)] = o2;
    }
    if (t == 0) *cnt = 0u;                           // ready for the next launch / graph replay
}

// x: HOST array of nlev device pointers ([B,P_l,ldx] tensors), P: HOST array; partial: sum_l B*ceil(P_l/128)*66 floats;
// add: [nlev][B][64]; counters: nlev*B unsigned ints, zero before the first call (the kernel leaves them zero).
static int ctx_launch(int nlev, const void* const* x, int ldx, const float* wmask, const float* w1, const float* w2, float* partial,
                      float* add, float* pool_out, int* counters, int B, const int* P, int x_bf16, cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !x || !P || !wmask || !partial || !counters || (ldx & 7) || B <= 0)
        return FCVSR_ERR_ARG;
    CtxArgs a;
    a.nlev = nlev; a.ldx = ldx; a.B = B; a.wmask = wmask; a.partial = partial;
    // Pixels per block: a multiple of 256, sized so that the launch is about one wave of three blocks per SM -- a block's fixed
    // costs (launch, fence + counter round trip) are ~3 us, as long as the pooling of 256 pixels -- and so that no level needs
    // more than CTX_MAX_NBLK blocks (the caller's `partial` buffer is sized for 128 pixels per block, an upper bound)
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    int pmax = 0;
    long long ptot = 0;
    for (int l = 0; l < nlev; ++l) {
        if (!x[l] || P[l] <= 0 || ((uintptr_t)x[l] & 15)) return FCVSR_ERR_ARG;
        pmax = P[l] > pmax ? P[l] : pmax;
        ptot += (long long)B * P[l];
    }
    int ppb = CTX_PPB * (int)((ptot + 3LL * num_sms * CTX_PPB - 1) / (3LL * num_sms * CTX_PPB));
    if (ppb > 16 * CTX_PPB) ppb = 16 * CTX_PPB;
    while ((pmax + ppb - 1) / ppb > CTX_MAX_NBLK) ppb += CTX_PPB;
    a.ppb = ppb;
    int blk = 0;
    long long off = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        a.lv[l].x = x[j]; a.lv[l].P = P[j]; a.lv[l].nblk = (P[j] + ppb - 1) / ppb;
        a.lv[l].blk_begin = blk; a.lv[l].part_off = off;
        if (l < nlev) { blk += a.lv[l].nblk; off += (long long)B * ((P[j] + CTX_PIX_PER_BLOCK - 1) / CTX_PIX_PER_BLOCK) * CTX_STRIDE; }
    }
    if (x_bf16) ctx_block_kernel<true><<<dim3(blk, B), 256, 0, st>>>(a, w1, w2, add, reinterpret_cast<unsigned*>(counters), pool_out);
    else ctx_block_kernel<false><<<dim3(blk, B), 256, 0, st>>>(a, w1, w2, add, reinterpret_cast<unsigned*>(counters), pool_out);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_context_block_multi(int nlev, const void* const* x, int ldx, const float* wmask, const float* w1,
                                         const float* w2, float* partial, float* add, int* counters, int B, const int* P, int x_bf16,
                                         cudaStream_t st) {
    if (!w1 || !w2 || !add) return FCVSR_ERR_ARG;
    return ctx_launch(nlev, x, ldx, wmask, w1, w2, partial, add, nullptr, counters, B, P, x_bf16, st);
}

// Training: the soft-max pooling alone.  pool: [nlev][B][66] = context[64], running max (log2 domain), sum -- what
// fcvsr_context_pool_backward_multi needs; the 64 -> 64 -> 64 MLP of the ContextBlock stays with autograd.
extern "C" int fcvsr_context_pool_multi(int nlev, const void* const* x, int ldx, const float* wmask, float* partial, float* pool,
                                        int* counters, int B, const int* P, int x_bf16, cudaStream_t st) {
    if (!pool) return FCVSR_ERR_ARG;
    return ctx_launch(nlev, x, ldx, wmask, nullptr, nullptr, partial, nullptr, pool, counters, B, P, x_bf16, st);
}

extern "C" int fcvsr_context_block(const void* x, int ldx, const float* wmask, const float* w1, const float* w2,
                                   float* partial, float* add, int* counters, int B, int P, int x_bf16, cudaStream_t st) {
    if (!x) return FCVSR_ERR_ARG;
    const void* xs[1] = {x};
    return fcvsr_context_block_multi(1, xs, ldx, wmask, w1, w2, partial, add, counters, B, &P, x_bf16, st);
}

// Backward of the soft-max pooling ctx[c] = sum_p prob_p x[p][c], prob = softmax_p(w . x[p]) (training step; the inference
// path never needs it).  With g = d loss / d ctx and cg = ctx . g:
//     d loss / d x[p][c] = prob_p * (g[c] + w[c] * (x[p] . g - cg)),      d loss / d w[c] = sum_p prob_p * (x[p] . g - cg) * x[p][c]
// One pass over x: a lane owns 8 channels (as in the forward pass), the two dot products of a pixel need 3 shuffle steps per
// four pixels; the weight gradient is accumulated per lane and reduced per block into dwpart[block][64] (summed by the caller).
struct CtxBwdArgs {
    CtxLevel lv[SC_MAX_LEV]; int nlev; int ldx; int ppb; int B;
    const float* wmask; const float* pool; const float* gctx; float* dwpart;
    float* dx[SC_MAX_LEV];
};

__global__ void __launch_bounds__(256) ctx_pool_bwd_kernel(const CtxBwdArgs a) {
    __shared__ __align__(16) float gs[64], wsm[64], cs[64];
    __shared__ float red[8][64];
    __shared__ float cg_s;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, b = blockIdx.y;
    const int sub = lane >> 3, oct = lane & 7;
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const int P = CTX_SEL(P), ppb = a.ppb, ldx = a.ldx, lbx = bx - CTX_SEL(blk_begin);
    const float* x = reinterpret_cast<const float*>(CTX_SEL(x));
    float* dx = l == 0 ? a.dx[0] : (l == 1 ? a.dx[1] : a.dx[2]);
    const float* po = a.pool + ((size_t)l * a.B + b) * CTX_STRIDE;
    if (t < 64) {
        gs[t] = a.gctx[((size_t)l * a.B + b) * 64 + t];
        wsm[t] = a.wmask[t];
        cs[t] = po[t];
    }
    __syncthreads();
    if (warp == 0) {
        float v = gs[lane] * cs[lane] + gs[lane + 32] * cs[lane + 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) cg_s = v;
    }
    __syncthreads();
    const float M = po[64], rZ = 1.f / po[65], cg = cg_s;
    float g[8], w[8], dw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { g[j] = gs[oct * 8 + j]; w[j] = wsm[oct * 8 + j]; dw[j] = 0.f; }
    const int per_warp = ppb >> 3;
    const int p0 = lbx * ppb + warp * per_warp + sub;
    const size_t img = (size_t)b * P * ldx + oct * 8;
    for (int it = 0; it < per_warp; it += 16) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + it + 4 * u;
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
            if (p < P) {
                const float4* src = reinterpret_cast<const float4*>(x + img + (size_t)p * ldx);
                r0 = __ldg(src); r1 = __ldg(src + 1);
            }
            v[u][0] = r0.x; v[u][1] = r0.y; v[u][2] = r0.z; v[u][3] = r0.w;
            v[u][4] = r1.x; v[u][5] = r1.y; v[u][6] = r1.z; v[u][7] = r1.w;
        }
        float lg[4], sg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { a0 = fmaf(v[u][j], w[j], a0); a1 = fmaf(v[u][j], g[j], a1); }
            lg[u] = a0; sg[u] = a1;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                lg[u] += __shfl_xor_sync(0xffffffffu, lg[u], o);
                sg[u] += __shfl_xor_sync(0xffffffffu, sg[u], o);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + it + 4 * u;
            if (p >= P) continue;
            const float prob = ctx_ex2(lg[u] * CTX_LOG2E - M) * rZ;
            const float tt = prob * (sg[u] - cg);
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] = fmaf(tt, w[j], prob * g[j]);
                dw[j] = fmaf(tt, v[u][j], dw[j]);
            }
            float4* dst = reinterpret_cast<float4*>(dx + img + (size_t)p * ldx);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dw[j] += __shfl_xor_sync(0xffffffffu, dw[j], o);
    }
    if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][oct * 8 + j] = dw[j];
    }
    __syncthreads();
    if (t < 64) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[k][t];
        a.dwpart[((size_t)b * gridDim.x + bx) * 64 + t] = sum;
    }
}

// x / dx: HOST arrays of nlev device pointers ([B,P_l,ldx] fp32; dx is WRITTEN); pool: the output of fcvsr_context_pool_multi;
// gctx: d loss / d context [nlev][B][64]; dwpart: B * nblocks * 64 floats, nblocks = fcvsr_context_pool_backward_blocks(...)
// (the caller sums them: d loss / d wmask).
extern "C" int fcvsr_context_pool_backward_blocks(int nlev, const int* P) {
    int blk = 0;
    for (int l = 0; l < nlev && l < SC_MAX_LEV; ++l) blk += (P[l] + CTX_PPB - 1) / CTX_PPB;
    return blk;
}
extern "C" int fcvsr_context_pool_backward_multi(int nlev, const void* const* x, int ldx, const float* wmask, const float* pool,
                                                 const float* gctx, float* const* dx, float* dwpart, int B, const int* P,
                                                 cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !x || !P || !wmask || !pool || !gctx || !dx || !dwpart || (ldx & 7) || B <= 0 || B > 65535)
        return FCVSR_ERR_ARG;
    CtxBwdArgs a;
    a.nlev = nlev; a.ldx = ldx; a.B = B; a.ppb = CTX_PPB; a.wmask = wmask; a.pool = pool; a.gctx = gctx; a.dwpart = dwpart;
    int blk = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        if (!x[j] || !dx[j] || P[j] <= 0 || (((uintptr_t)x[j] | (uintptr_t)dx[j]) & 15)) return FCVSR_ERR_ARG;
        a.lv[l].x = x[j]; a.lv[l].P = P[j]; a.lv[l].nblk = (P[j] + CTX_PPB - 1) / CTX_PPB;
        a.lv[l].blk_begin = blk; a.lv[l].part_off = 0;
        a.dx[l] = dx[j];
        if (l < nlev) blk += a.lv[l].nblk;
    }
    ctx_pool_bwd_kernel<<<dim3(blk, B), 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

// ---- RCB tail (:720-724): r = lrelu_0.2(res + add[b]) + r0, all 64 channels, float4 per thread ----------------------------
// A level with r_pool runs one thread per 2x2 pixel quad x 4 channels and additionally writes the quad mean: the 1x1 `down`
// convolution commutes with the 2x2 average that follows it in the reference (:753-757, Interpolate(0.5) of an even-sized
// map), so it runs on this pooled tensor at a quarter of the pixels.  pool_plain: store the mean as plain fp32 (exact mode)
// instead of the operand type.
struct RcbLevel {
    const void* res; const float* add; const void* r0; float* r; void* r_op; void* r_pool;
    int H, W, P; int blk_begin, bpr;          // first block of the level; blocks per row of pixels (or of 2x2 quads)
};
struct RcbArgs { RcbLevel lv[SC_MAX_LEV]; int nlev; int op16; int pool_plain; };

__device__ __forceinline__ float4 rcb_value(float4 v, float4 a, float4 q) {
    float4 o;
    o.x = v.x + a.x; o.y = v.y + a.y; o.z = v.z + a.z; o.w = v.w + a.w;
    o.x = (o.x >= 0.f ? o.x : 0.2f * o.x) + q.x;
    o.y = (o.y >= 0.f ? o.y : 0.2f * o.y) + q.y;
    o.z = (o.z >= 0.f ? o.z : 0.2f * o.z) + q.z;
    o.w = (o.w >= 0.f ? o.w : 0.2f * o.w) + q.w;
    return o;
}

template <int RES16, int R016>
__global__ void rcb_finish_kernel(const RcbArgs a) {
    // a block is a run of 16 pixels (or 2x2 quads) of one row: block-uniform divisions instead of per-thread 64-bit div / mod
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const RcbLevel& L = a.lv[l];
    const int lb = bx - L.blk_begin;
    const int rowid = lb / L.bpr, xq = (lb - rowid * L.bpr) * 16 + (threadIdx.x >> 4);
    const int c = (threadIdx.x & 15) * 4;
    if (L.r_pool) {
        const int W = L.W, H = L.H;
        const int w2 = W >> 1, h2 = H >> 1;
        if (xq >= w2) return;
        const int b = rowid / h2, qy = rowid - b * h2, qx = xq;             // rowid = b * (H / 2) + qy
        const size_t quad = (size_t)rowid * w2 + qx;
        const size_t p00 = ((size_t)b * H + 2 * qy) * W + 2 * qx;
        const size_t pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
        const float4 ad = *reinterpret_cast<const float4*>(L.add + (size_t)b * 64 + c);
        float4 v[4], q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = load4_any(L.res, pix[k] * 64 + c, RES16);
            q[k] = load4_any(L.r0, pix[k] * 64 + c, R016);
        }
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 o = rcb_value(v[k], ad, q[k]);
            if (L.r_op) store_operand4(L.r_op, pix[k] * 64 + c, o, a.op16);
            if (L.r) *reinterpret_cast<float4*>(L.r + pix[k] * 64 + c) = o;
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
        }
        m.x *= 0.25f; m.y *= 0.25f; m.z *= 0.25f; m.w *= 0.25f;
        if (a.pool_plain) *reinterpret_cast<float4*>(reinterpret_cast<float*>(L.r_pool) + quad * 64 + c) = m;
        else store_operand4(L.r_pool, quad * 64 + c, m, a.op16);
    } else {
        if (xq >= L.P) return;
        const int b = rowid;                                               // one "row" per image: P pixels
        const size_t pix = (size_t)b * L.P + xq;
        const float4 o = rcb_value(load4_any(L.res, pix * 64 + c, RES16), *reinterpret_cast<const float4*>(L.add + (size_t)b * 64 + c),
                                   load4_any(L.r0, pix * 64 + c, R016));
        if (L.r_op) store_operand4(L.r_op, pix * 64 + c, o, a.op16);     // tensor-core operand copy for the 1x1 down/up convs
        if (L.r) *reinterpret_cast<float4*>(L.r + pix * 64 + c) = o;
    }
}

// res / add / r0 / r / r_op / r_pool: HOST arrays of nlev device pointers (r, r_op, r_pool entries may be NULL; a level
// with r_pool needs even H, W); H, W: HOST arrays.  res_bf16: bit 0 = res tensors are bf16, bit 1 = r0 tensors are.
extern "C" int fcvsr_rcb_finish_multi(int nlev, const void* const* res, const float* const* add, const void* const* r0,
                                      float* const* r, void* const* r_op, void* const* r_pool, const int* H, const int* W, int B,
                                      int op16, int pool_plain, int res_bf16, cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !res || !add || !r0 || !H || !W || B <= 0) return FCVSR_ERR_ARG;
    RcbArgs a;
    a.nlev = nlev; a.op16 = op16; a.pool_plain = pool_plain;
    long long t = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        RcbLevel& L = a.lv[l];
        L.res = res[j]; L.add = add[j]; L.r0 = r0[j]; L.r = r ? r[j] : nullptr; L.r_op = r_op ? r_op[j] : nullptr;
        L.r_pool = r_pool ? r_pool[j] : nullptr;
        L.H = H[j]; L.W = W[j]; L.P = H[j] * W[j]; L.blk_begin = (int)t;
        L.bpr = L.r_pool ? ((W[j] >> 1) + 15) / 16 : (L.P + 15) / 16;
        if (l >= nlev) continue;
        if (!L.res || !L.add || !L.r0 || (!L.r && !L.r_op) || L.H <= 0 || L.W <= 0) return FCVSR_ERR_ARG;
        if (L.r_pool && ((L.H | L.W) & 1)) return FCVSR_ERR_ARG;
        t += L.r_pool ? (long long)B * (L.H >> 1) * L.bpr : (long long)B * L.bpr;
        if (t > 0x7fffffffLL) return FCVSR_ERR_UNSUPPORTED;
    }
    const unsigned grid = (unsigned)t;
#define RF_LAUNCH(A, C) rcb_finish_kernel<A, C><<<grid, 256, 0, st>>>(a)
    switch (res_bf16 & 3) {
        case 0: RF_LAUNCH(0, 0); break;
        case 1: RF_LAUNCH(1, 0); break;
        case 2: RF_LAUNCH(0, 1); break;
        default: RF_LAUNCH(1, 1); break;
    }
#undef RF_LAUNCH
    return fcvsr_launch_status();
}

extern "C" int fcvsr_rcb_finish(const void* res, const float* add, const void* r0, float* r, int B, int P,
                                void* r_op, int op16, void* r_pool, int H, int W, int pool_plain, int res_bf16, cudaStream_t st) {
    if (!res || !add || !r0 || (!r && !r_op)) return FCVSR_ERR_ARG;
    if (r_pool) {
        if (H <= 0 || W <= 0 || ((H | W) & 1) || (size_t)H * W != (size_t)P) return FCVSR_ERR_ARG;
    } else {
        H = 1; W = P;            // plain path only needs the pixel count
    }
    const void* ress[1] = {res}; const float* adds[1] = {add}; const void* r0s[1] = {r0};
    float* rs[1] = {r}; void* rops[1] = {r_op}; void* rpools[1] = {r_pool};
    return fcvsr_rcb_finish_multi(1, ress, adds, r0s, rs, rops, rpools, &H, &W, B, op16, pool_plain, res_bf16, st);
}

// Backward of the RCB tail r = lrelu_0.2(res + add[b]) + r0 (training step): gres = g * (res + add >= 0 ? 1 : 0.2), the gradient
// of r0 is g itself, and gadd[level][b][c] += sum over pixels of gres (fp32 atomics of per-block sums; zero it first).  Same
// row-based mapping as the forward kernel: a block is a run of 16 pixels of one image, a thread 4 channels.
struct RcbBwdLevel { const float* res; const float* add; const float* g; float* gres; float* gadd; int P; int blk_begin, bpr; };
struct RcbBwdArgs { RcbBwdLevel lv[SC_MAX_LEV]; int nlev; };

__global__ void __launch_bounds__(256) rcb_finish_bwd_kernel(const RcbBwdArgs a) {
    __shared__ float4 red[256];
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const RcbBwdLevel& L = a.lv[l];
    const int lb = bx - L.blk_begin;
    const int b = lb / L.bpr, xq = (lb - b * L.bpr) * 16 + (threadIdx.x >> 4);
    const int c = (threadIdx.x & 15) * 4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xq < L.P) {
        const size_t i = ((size_t)b * L.P + xq) * 64 + c;
        const float4 r = *reinterpret_cast<const float4*>(L.res + i);
        const float4 ad = *reinterpret_cast<const float4*>(L.add + (size_t)b * 64 + c);
        const float4 g = *reinterpret_cast<const float4*>(L.g + i);
        o.x = g.x * (r.x + ad.x >= 0.f ? 1.f : 0.2f);
        o.y = g.y * (r.y + ad.y >= 0.f ? 1.f : 0.2f);
        o.z = g.z * (r.z + ad.z >= 0.f ? 1.f : 0.2f);
        o.w = g.w * (r.w + ad.w >= 0.f ? 1.f : 0.2f);
        *reinterpret_cast<float4*>(L.gres + i) = o;
    }
    red[threadIdx.x] = o;
    __syncthreads();
    if (threadIdx.x < 16) {            // sum of the block's 16 pixels for this thread's 4 channels
        float4 sum = red[threadIdx.x];
#pragma unroll
        for (int k = 1; k < 16; ++k) {
            const float4 v = red[threadIdx.x + 16 * k];
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        float* d = L.gadd + (size_t)b * 64 + threadIdx.x * 4;
        atomicAdd(d, sum.x); atomicAdd(d + 1, sum.y); atomicAdd(d + 2, sum.z); atomicAdd(d + 3, sum.w);
    }
}

// res / add / g / gres / gadd: HOST arrays of nlev device pointers (64-channel fp32 tensors [B,P_l,64]; add, gadd: [B,64]);
// gres is WRITTEN, gadd ACCUMULATED (zero it first).
extern "C" int fcvsr_rcb_finish_backward_multi(int nlev, const float* const* res, const float* const* add, const float* const* g,
                                               float* const* gres, float* const* gadd, const int* P, int B, cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !res || !add || !g || !gres || !gadd || !P || B <= 0) return FCVSR_ERR_ARG;
    RcbBwdArgs a;
    a.nlev = nlev;
    long long t = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        RcbBwdLevel& L = a.lv[l];
        L.res = res[j]; L.add = add[j]; L.g = g[j]; L.gres = gres[j]; L.gadd = gadd[j]; L.P = P[j];
        L.blk_begin = (int)t; L.bpr = (P[j] + 15) / 16;
        if (l >= nlev) continue;
        if (!L.res || !L.add || !L.g || !L.gres || !L.gadd || L.P <= 0) return FCVSR_ERR_ARG;
        if (((uintptr_t)L.res | (uintptr_t)L.add | (uintptr_t)L.g | (uintptr_t)L.gres) & 15) return FCVSR_ERR_ARG;
        t += (long long)B * L.bpr;
        if (t > 0x7fffffffLL) return FCVSR_ERR_UNSUPPORTED;
    }
    rcb_finish_bwd_kernel<<<(unsigned)t, 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

// ---- BlockRCB cross-level sum (:766-777): x[b,y,x,:] += coef * r + mean2x2(td) + bilinear_x2(tu) (64 channels) -------------
struct MixLevel {
    const float* xin; float* xout; const void* r; const void* td; const void* tu; void* xout_r;
    float coef; int H, W; int blk_begin, bpr;          // first block of the level, blocks per image row (16 threads per pixel)
};
struct MixArgs { MixLevel lv[SC_MAX_LEV]; int nlev; int ldx, ldo, ldr; int round_main, op16, td_pooled; };

template <int R16, int T16, int X16>
__global__ void level_mix_kernel(const MixArgs a) {
    // a block is a run of 16 pixels of ONE image row: level, row and column come from block-uniform divisions (three 64-bit
    // div / mod per thread were a third of the kernel's instructions; ncu: SM throughput 74 % at 61 % of the copy bandwidth)
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const MixLevel& L = a.lv[l];
    const int H = L.H, W = L.W;
    const int lb = bx - L.blk_begin;
    const int rowid = lb / L.bpr, chunk = lb - rowid * L.bpr;      // rowid = b * H + y
    const int x = chunk * 16 + (threadIdx.x >> 4);
    if (x >= W) return;
    const int b = rowid / H, y = rowid - b * H;
    const int c = (threadIdx.x & 15) * 4;
    const size_t pix = (size_t)rowid * W + x;
    const void* td = L.td;
    const void* tu = L.tu;
    const float coef = L.coef;
    float4 o = load4_any(L.xin, pix * a.ldx + c, X16);      // X16: the carried x is the previous block's bf16 operand copy
    constexpr int t16 = T16;            // compile-time: the loads below stay straight-line (a run-time flag put every load
    const float4 rv = load4_any(L.r, pix * 64 + c, R16);    // behind its own branch: 18.7 -> 23.7 us per launch)
    o.x = fmaf(coef, rv.x, o.x); o.y = fmaf(coef, rv.y, o.y); o.z = fmaf(coef, rv.z, o.z); o.w = fmaf(coef, rv.w, o.w);
    if (td && (a.td_pooled & 1)) {   // td is [B,H,W,64]: the down conv already ran on the 2x2 mean (see rcb_finish_kernel)
        const float4 a0 = load4_any(td, pix * 64 + c, t16);
        o.x += a0.x; o.y += a0.y; o.z += a0.z; o.w += a0.w;
    } else if (td) {   // td is [B,2H,2W,64]
        const size_t base = (((size_t)b * 2 * H + 2 * y) * (2 * W) + 2 * x) * 64 + c;
        const float4 a0 = load4_any(td, base, t16);
        const float4 a1 = load4_any(td, base + 64, t16);
        const float4 a2 = load4_any(td, base + (size_t)2 * W * 64, t16);
        const float4 a3 = load4_any(td, base + (size_t)2 * W * 64 + 64, t16);
        o.x += 0.25f * (a0.x + a1.x + a2.x + a3.x);
        o.y += 0.25f * (a0.y + a1.y + a2.y + a3.y);
        o.z += 0.25f * (a0.z + a1.z + a2.z + a3.z);
        o.w += 0.25f * (a0.w + a1.w + a2.w + a3.w);
    }
    if (tu) {   // tu is [B,H/2,W/2,64]; bilinear, align_corners=False, scale 2
        const int hs = H >> 1, ws = W >> 1;
        const float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
        const float ly = sy - y0, lx = sx - x0;
        const size_t tb = (size_t)b * hs * ws * 64 + c;
        const float4 a00 = load4_any(tu, tb + ((size_t)y0 * ws + x0) * 64, t16);
        const float4 a01 = load4_any(tu, tb + ((size_t)y0 * ws + x1) * 64, t16);
        const float4 a10 = load4_any(tu, tb + ((size_t)y1 * ws + x0) * 64, t16);
        const float4 a11 = load4_any(tu, tb + ((size_t)y1 * ws + x1) * 64, t16);
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        o.x += w00 * a00.x + w01 * a01.x + w10 * a10.x + w11 * a11.x;
        o.y += w00 * a00.y + w01 * a01.y + w10 * a10.y + w11 * a11.y;
        o.z += w00 * a00.z + w01 * a01.z + w10 * a10.z + w11 * a11.z;
        o.w += w00 * a00.w + w01 * a01.w + w10 * a10.w + w11 * a11.w;
    }
    if (L.xout_r) store_operand4(L.xout_r, pix * a.ldr + c, o, a.op16);
    if (!L.xout) return;                                                        // only the operand copy is kept (bf16 mode)
    if (a.round_main) store_operand4(L.xout, pix * a.ldo + c, o, a.op16);      // xout itself is an operand-typed tensor
    else *reinterpret_cast<float4*>(L.xout + pix * a.ldo + c) = o;
}

// The bf16 mode's BlockRCB launch (r, td, tu bf16, td already pooled, only the bf16 operand copy written): eight channels per
// thread, 16-byte accesses on every bf16 tensor, a block = 32 pixels of one image row.  With four channels per thread the
// kernel is bound by its instruction count (ncu: SM throughput 74 % at 4.6 TB/s; halving its DRAM bytes moved it by 3 %), and
// the per-pixel index / weight arithmetic is what eight threads per pixel instead of sixteen halve.  Same expressions per
// channel as level_mix_kernel: the results are bit-identical.
__device__ __forceinline__ void lm8_unpack(const uint4 u, float* f) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
template <int X16>
__global__ void __launch_bounds__(256, 4) level_mix8_kernel(const MixArgs a) {
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const MixLevel& L = a.lv[l];
    const int H = L.H, W = L.W;
    const int lb = bx - L.blk_begin;
    const int rowid = lb / L.bpr, chunk = lb - rowid * L.bpr;      // rowid = b * H + y
    const int x = chunk * 32 + (threadIdx.x >> 3);
    if (x >= W) return;
    const int b = rowid / H, y = rowid - b * H;
    const int c = (threadIdx.x & 7) * 8;
    const size_t pix = (size_t)rowid * W + x;
    const unsigned short* td = reinterpret_cast<const unsigned short*>(L.td);
    const unsigned short* tu = reinterpret_cast<const unsigned short*>(L.tu);
    const float coef = L.coef;
    // every load is issued before the first use
    uint4 xr = make_uint4(0u, 0u, 0u, 0u), dv = xr, u00 = xr, u01 = xr, u10 = xr, u11 = xr;
    float4 xf0 = make_float4(0.f, 0.f, 0.f, 0.f), xf1 = xf0;
    if (X16) {
        xr = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(L.xin) + pix * a.ldx + c);
    } else {
        xf0 = *reinterpret_cast<const float4*>(L.xin + pix * a.ldx + c);
        xf1 = *reinterpret_cast<const float4*>(L.xin + pix * a.ldx + c + 4);
    }
    const uint4 rv = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(L.r) + pix * 64 + c);
    if (td) dv = *reinterpret_cast<const uint4*>(td + pix * 64 + c);
    float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
    if (tu) {
        const int hs = H >> 1, ws = W >> 1;
        const float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
        const float ly = sy - y0, lx = sx - x0;
        const unsigned short* tb = tu + (size_t)b * hs * ws * 64 + c;
        u00 = *reinterpret_cast<const uint4*>(tb + ((size_t)y0 * ws + x0) * 64);
        u01 = *reinterpret_cast<const uint4*>(tb + ((size_t)y0 * ws + x1) * 64);
        u10 = *reinterpret_cast<const uint4*>(tb + ((size_t)y1 * ws + x0) * 64);
        u11 = *reinterpret_cast<const uint4*>(tb + ((size_t)y1 * ws + x1) * 64);
        w00 = (1.f - ly) * (1.f - lx); w01 = (1.f - ly) * lx; w10 = ly * (1.f - lx); w11 = ly * lx;
    }
    float o[8], t[8];
    if (X16) {
        lm8_unpack(xr, o);
    } else {
        o[0] = xf0.x; o[1] = xf0.y; o[2] = xf0.z; o[3] = xf0.w; o[4] = xf1.x; o[5] = xf1.y; o[6] = xf1.z; o[7] = xf1.w;
    }
    lm8_unpack(rv, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(coef, t[i], o[i]);
    if (td) {
        lm8_unpack(dv, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += t[i];
    }
    if (tu) {
        float t1[8], t2[8], t3[8];
        lm8_unpack(u00, t); lm8_unpack(u01, t1); lm8_unpack(u10, t2); lm8_unpack(u11, t3);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += w00 * t[i] + w01 * t1[i] + w10 * t2[i] + w11 * t3[i];
    }
    uint4 pk;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(o[1]), "f"(o[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(o[3]), "f"(o[2]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.z) : "f"(o[5]), "f"(o[4]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.w) : "f"(o[7]), "f"(o[6]));
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(L.xout_r) + pix * a.ldr + c) = pk;
}

// xin / xout / r / td / tu / xout_r: HOST arrays of nlev device pointers (td, tu, xout_r entries may be NULL); coef, H, W: HOST
// arrays.  Flags as fcvsr_level_mix.
extern "C" int fcvsr_level_mix_multi(int nlev, const float* const* xin, int ldx, float* const* xout, int ldo, const void* const* r,
                                     const float* coef, const void* const* td, const void* const* tu, int B, const int* H,
                                     const int* W, void* const* xout_r, int ldr, int round_main, int op16, int td_pooled,
                                     cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !xin || !xout || !r || !coef || !H || !W || (ldx & 3) || (ldo & 3) || B <= 0) return FCVSR_ERR_ARG;
    if ((td_pooled & 8) && !xout_r) return FCVSR_ERR_ARG;
    MixArgs a;
    a.nlev = nlev; a.ldx = ldx; a.ldo = ldo; a.ldr = ldr; a.round_main = round_main; a.op16 = op16; a.td_pooled = td_pooled;
    long long t = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        MixLevel& L = a.lv[l];
        L.xin = xin[j]; L.xout = xout[j]; L.r = r[j]; L.td = td ? td[j] : nullptr; L.tu = tu ? tu[j] : nullptr;
        L.xout_r = xout_r ? xout_r[j] : nullptr; L.coef = coef[j]; L.H = H[j]; L.W = W[j];
        L.blk_begin = (int)t; L.bpr = (W[j] + 15) / 16;
        if (l >= nlev) continue;
        if (!L.xin || (!L.xout && !L.xout_r) || !L.r || L.H <= 0 || L.W <= 0 || (L.tu && ((L.H | L.W) & 1)) || (L.xout_r && (ldr & 3))) return FCVSR_ERR_ARG;
        t += (long long)B * L.H * L.bpr;
        if (t > 0x7fffffffLL) return FCVSR_ERR_UNSUPPORTED;
    }
    // bf16 mode's BlockRCB launch: only the operand copy is written, every side tensor is bf16, td is pooled
    bool all8 = op16 && (td_pooled & 7) == 7 && xout_r && !(ldx & 7) && !(ldr & 7);
    for (int l = 0; l < nlev && all8; ++l)
        all8 = !xout[l] && xout_r[l] && !(((uintptr_t)xin[l] | (uintptr_t)r[l] | (uintptr_t)xout_r[l] | (uintptr_t)(td ? td[l] : nullptr) |
                                           (uintptr_t)(tu ? tu[l] : nullptr)) & 15);
    if (all8) {
        long long t8 = 0;
        for (int l = 0; l < SC_MAX_LEV; ++l) {
            a.lv[l].blk_begin = (int)t8; a.lv[l].bpr = (a.lv[l].W + 31) / 32;
            if (l < nlev) t8 += (long long)B * a.lv[l].H * a.lv[l].bpr;
        }
        if (td_pooled & 8) level_mix8_kernel<1><<<(unsigned)t8, 256, 0, st>>>(a);
        else level_mix8_kernel<0><<<(unsigned)t8, 256, 0, st>>>(a);
        return fcvsr_launch_status();
    }
    const unsigned grid = (unsigned)t;
#define LM_LAUNCH(R, T, X) level_mix_kernel<R, T, X><<<grid, 256, 0, st>>>(a)
    switch ((td_pooled >> 1) & 7) {
        case 0: LM_LAUNCH(0, 0, 0); break;
        case 1: LM_LAUNCH(1, 0, 0); break;
        case 2: LM_LAUNCH(0, 1, 0); break;
        case 3: LM_LAUNCH(1, 1, 0); break;
        case 7: LM_LAUNCH(1, 1, 1); break;          // bf16 mode, blocks 2 and 3 of a group: x carried as the bf16 operand copy
        default: return FCVSR_ERR_UNSUPPORTED;      // a bf16 xin only together with bf16 r / td / tu
    }
#undef LM_LAUNCH
    return fcvsr_launch_status();
}

extern "C" int fcvsr_level_mix(const float* xin, int ldx, float* xout, int ldo, const void* r, float coef,
                               const void* td, const void* tu, int B, int H, int W, void* xout_r, int ldr,
                               int round_main, int op16, int td_pooled, cudaStream_t st) {
    const float* xins[1] = {xin}; float* xouts[1] = {xout}; const void* rs[1] = {r};
    const void* tds[1] = {td}; const void* tus[1] = {tu}; void* xors[1] = {xout_r};
    return fcvsr_level_mix_multi(1, xins, ldx, xouts, ldo, rs, &coef, tds, tus, B, &H, &W, xors, ldr, round_main, op16, td_pooled, st);
}

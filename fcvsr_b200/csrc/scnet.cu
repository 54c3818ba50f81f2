// Non-GEMM kernels of the SCNetbk trunk (CVSR_freq.py:657-822).
//
//   ctx_partial / ctx_finalize   ContextBlock (:657-701): softmax-over-HW attention pooling done as an
//                                online-softmax reduction (running max / sum / weighted channel sums
//                                per block, merged in fixed order) followed by the 64->64->64 MLP.
//   rcb_finish                   RCB tail (:720-724):  r = lrelu_0.2(res + add_term) + r0
//   level_mix                    BlockRCB cross-level sum (:766-777):
//                                x += coef*r + avgpool2(td) + bilinear_x2(tu)
//                                (Interpolate(0.5) of an even-sized map == 2x2 mean; Interpolate(2.0) is
//                                bilinear, align_corners=False; td/tu are the 1x1 down/up conv outputs)
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

#define CTX_PIX_PER_BLOCK 128
#define CTX_STRIDE 66          // m, z, acc[64]

#define SC_MAX_LEV 3
// Every helper below can run the three pyramid levels of a BlockRCB (CVSR_freq.py:766-777) in ONE launch: the levels hold 1, 1/4
// and 1/16 of the pixels, so per-level launches of the small ones are pure latency (8 us for a 6-CTA kernel) and used to run on
// side streams.  A launch gets up to SC_MAX_LEV level descriptors; blocks / threads find their level by index range.
struct CtxLevel { const void* x; int P; int nblk; int blk_begin; long long part_off; };    // part_off: floats into `partial`
struct CtxArgs { CtxLevel lv[SC_MAX_LEV]; int nlev; int ldx; int ppb; int B; const float* wmask; float* partial; };

#define CTX_SEL(field) (l == 0 ? a.lv[0].field : (l == 1 ? a.lv[1].field : a.lv[2].field))

// One block = ppb pixels (a multiple of 128), 8 warps, 8 pixels in flight per warp (lane owns 2 channels, so a
// pixel is one coalesced 256-byte load and its logit one 5-step shuffle reduction).  The level descriptor is read field by
// field with compile-time indices: indexing the kernel parameters with a run-time level copies them to local memory.
template <bool X16>
__global__ void __launch_bounds__(256) ctx_partial_kernel(const CtxArgs a) {
    __shared__ float sm_m[8], sm_z[8];
    __shared__ float sm_acc[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y, c = 2 * lane;
    const int bx = blockIdx.x;
    const int l = (a.nlev > 1 && bx >= a.lv[1].blk_begin) ? ((a.nlev > 2 && bx >= a.lv[2].blk_begin) ? 2 : 1) : 0;
    const int P = CTX_SEL(P), ppb = a.ppb, ldx = a.ldx, lbx = bx - CTX_SEL(blk_begin), nblk = CTX_SEL(nblk);
    const long long part_off = CTX_SEL(part_off);
    const float* x = reinterpret_cast<const float*>(CTX_SEL(x));
    const float2 w = *reinterpret_cast<const float2*>(a.wmask + c);
    float m = -INFINITY, z = 0.f;
    float2 acc = make_float2(0.f, 0.f);
    const int p0 = lbx * ppb + warp * (ppb >> 3);        // ppb pixels per block (multiple of 128), 1/8 per warp
    const float* xb = x + (size_t)b * P * ldx + c;
    const unsigned short* xb16 = reinterpret_cast<const unsigned short*>(x) + (size_t)b * P * ldx + c;   // x16: bf16 tensor
    for (int it = 0; it < (ppb >> 6); ++it) {
        float2 v[8];
        float lg[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = p0 + it * 8 + u;
            if (X16) {       // compile-time: the eight loads of an iteration stay back to back
                const uint32_t w2 = p < P ? *reinterpret_cast<const uint32_t*>(xb16 + (size_t)p * ldx) : 0u;
                v[u] = make_float2(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u));
            } else {
                v[u] = p < P ? *reinterpret_cast<const float2*>(xb + (size_t)p * ldx) : make_float2(0.f, 0.f);
            }
            lg[u] = v[u].x * w.x + v[u].y * w.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) lg[u] += __shfl_xor_sync(0xffffffffu, lg[u], o);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (p0 + it * 8 + u >= P) continue;
            const float mn = fmaxf(m, lg[u]);
            const float sc = __expf(m - mn), e = __expf(lg[u] - mn);
            z = z * sc + e;
            acc.x = acc.x * sc + e * v[u].x;
            acc.y = acc.y * sc + e * v[u].y;
            m = mn;
        }
    }
    if (lane == 0) { sm_m[warp] = m; sm_z[warp] = z; }
    sm_acc[warp][c] = acc.x; sm_acc[warp][c + 1] = acc.y;
    __syncthreads();
    if (threadIdx.x < 64) {
        float M = -INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) M = fmaxf(M, sm_m[k]);
        float Z = 0.f, A = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = sm_m[k] == -INFINITY ? 0.f : __expf(sm_m[k] - M);
            Z += sm_z[k] * s;
            A += sm_acc[k][threadIdx.x] * s;
        }
        float* dst = a.partial + part_off + ((size_t)b * nblk + lbx) * CTX_STRIDE;
        if (threadIdx.x == 0) { dst[0] = M; dst[1] = Z; }
        dst[2 + threadIdx.x] = A;
    }
}

// grid nlev * B, 1024 threads: merge the block partials of one (level, image) in fixed order -> context[64] ->
// add = W2 lrelu_0.2(W1 ctx).  The kernel is pure latency (a few CTAs on the whole GPU), so every stage is one round of
// independent loads: warp-shuffle reductions for max / sum, the per-partial scale exp(m_k - M) computed once into shared
// memory, 16 thread groups x 64 channels walking the partial list with 8 loads in flight, and the two 64x64 mat-vecs done
// one warp per output row (coalesced 256-byte row reads + shuffle reduction).
#define CTX_MAX_NBLK 4096
__global__ void __launch_bounds__(1024) ctx_finalize_kernel(const CtxArgs a, const float* __restrict__ w1,
                                                            const float* __restrict__ w2, float* __restrict__ add) {
    __shared__ float red[32];
    __shared__ float esc[CTX_MAX_NBLK];
    __shared__ float part[16][64];
    __shared__ float ctx[64], hid[64];
    const int l = blockIdx.x / a.B, b = blockIdx.x - l * a.B;
    const int nblk = CTX_SEL(nblk);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const float* pp = a.partial + CTX_SEL(part_off) + (size_t)b * nblk * CTX_STRIDE;
    float M = -INFINITY;
    for (int k = t; k < nblk; k += 1024) M = fmaxf(M, pp[k * CTX_STRIDE]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    if (lane == 0) red[warp] = M;
    __syncthreads();
    M = red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    __syncthreads();
    float Z = 0.f;
    for (int k = t; k < nblk; k += 1024) {
        const float e = __expf(pp[k * CTX_STRIDE] - M);
        esc[k] = e;
        Z += pp[k * CTX_STRIDE + 1] * e;
    }
    // fixed-order (deterministic) tree: lanes, then warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(0xffffffffu, Z, o);
    if (lane == 0) red[warp] = Z;
    __syncthreads();
    Z = red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(0xffffffffu, Z, o);
    const int c = t & 63, q = t >> 6;
    float A = 0.f;
    int k = q;
    for (; k + 112 < nblk; k += 128) {
        float av[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) av[u] = pp[(k + 16 * u) * CTX_STRIDE + 2 + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) A += av[u] * esc[k + 16 * u];
    }
    for (; k < nblk; k += 16) A += pp[k * CTX_STRIDE + 2 + c] * esc[k];
    part[q][c] = A;
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += part[j][t];
        ctx[t] = s / Z;
    }
    __syncthreads();
    for (int r = warp; r < 64; r += 32) {            // hid[r] = lrelu_0.2(W1[r,:] . ctx)
        float h = w1[r * 64 + lane] * ctx[lane] + w1[r * 64 + 32 + lane] * ctx[32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        if (lane == 0) hid[r] = h >= 0.f ? h : 0.2f * h;
    }
    __syncthreads();
    for (int r = warp; r < 64; r += 32) {            // add[r] = W2[r,:] . hid
        float o2 = w2[r * 64 + lane] * hid[lane] + w2[r * 64 + 32 + lane] * hid[32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) o2 += __shfl_xor_sync(0xffffffffu, o2, o);
        if (lane == 0) add[(size_t)blockIdx.x * 64 + r] = o2;
    }
}

// x: HOST array of nlev device pointers ([B,P_l,ldx] tensors), P: HOST array; partial: sum_l B*ceil(P_l/128)*66 floats;
// add: [nlev][B][64].
extern "C" int fcvsr_context_block_multi(int nlev, const void* const* x, int ldx, const float* wmask, const float* w1,
                                         const float* w2, float* partial, float* add, int B, const int* P, int x_bf16,
                                         cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !x || !P || !wmask || !w1 || !w2 || !partial || !add || (ldx & 1) || B <= 0) return FCVSR_ERR_ARG;
    CtxArgs a;
    a.nlev = nlev; a.ldx = ldx; a.B = B; a.wmask = wmask; a.partial = partial;
    // 128 pixels per block (as sized by the caller's `partial` buffer) unless that needs more than CTX_MAX_NBLK blocks
    int ppb = CTX_PIX_PER_BLOCK, pmax = 0;
    for (int l = 0; l < nlev; ++l) { if (!x[l] || P[l] <= 0) return FCVSR_ERR_ARG; pmax = P[l] > pmax ? P[l] : pmax; }
    while ((pmax + ppb - 1) / ppb > CTX_MAX_NBLK) ppb += CTX_PIX_PER_BLOCK;
    a.ppb = ppb;
    int blk = 0;
    long long off = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        a.lv[l].x = x[j]; a.lv[l].P = P[j]; a.lv[l].nblk = (P[j] + ppb - 1) / ppb;
        a.lv[l].blk_begin = blk; a.lv[l].part_off = off;
        if (l < nlev) { blk += a.lv[l].nblk; off += (long long)B * ((P[j] + CTX_PIX_PER_BLOCK - 1) / CTX_PIX_PER_BLOCK) * CTX_STRIDE; }
    }
    if (x_bf16) ctx_partial_kernel<true><<<dim3(blk, B), 256, 0, st>>>(a);
    else ctx_partial_kernel<false><<<dim3(blk, B), 256, 0, st>>>(a);
    ctx_finalize_kernel<<<nlev * B, 1024, 0, st>>>(a, w1, w2, add);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_context_block(const void* x, int ldx, const float* wmask, const float* w1, const float* w2,
                                   float* partial, float* add, int B, int P, int x_bf16, cudaStream_t st) {
    if (!x) return FCVSR_ERR_ARG;
    const void* xs[1] = {x};
    return fcvsr_context_block_multi(1, xs, ldx, wmask, w1, w2, partial, add, B, &P, x_bf16, st);
}

// ---- RCB tail (:720-724): r = lrelu_0.2(res + add[b]) + r0, all 64 channels, float4 per thread ----------------------------
// A level with r_pool runs one thread per 2x2 pixel quad x 4 channels and additionally writes the quad mean: the 1x1 `down`
// convolution commutes with the 2x2 average that follows it in the reference (:753-757, Interpolate(0.5) of an even-sized
// map), so it runs on this pooled tensor at a quarter of the pixels.  pool_plain: store the mean as plain fp32 (exact mode)
// instead of the operand type.
struct RcbLevel {
    const void* res; const float* add; const void* r0; float* r; void* r_op; void* r_pool;
    int H, W, P; long long t_begin;          // t_begin: first thread index of the level
};
struct RcbArgs { RcbLevel lv[SC_MAX_LEV]; int nlev; int op16; int pool_plain; long long total; };

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
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= a.total) return;
    const int l = (a.nlev > 1 && gi >= a.lv[1].t_begin) ? ((a.nlev > 2 && gi >= a.lv[2].t_begin) ? 2 : 1) : 0;
    const RcbLevel& L = a.lv[l];
    const size_t i = (size_t)(gi - L.t_begin);
    const int c = (int)(i & 15) * 4;
    if (L.r_pool) {
        const size_t quad = i >> 4;
        const int W = L.W, H = L.H;
        const int w2 = W >> 1, h2 = H >> 1;
        const int qx = (int)(quad % w2), qy = (int)((quad / w2) % h2), b = (int)(quad / ((size_t)w2 * h2));
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
        const size_t pix = i >> 4;
        const int b = (int)(pix / L.P);
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
        L.H = H[j]; L.W = W[j]; L.P = H[j] * W[j]; L.t_begin = t;
        if (l >= nlev) continue;
        if (!L.res || !L.add || !L.r0 || (!L.r && !L.r_op) || L.H <= 0 || L.W <= 0) return FCVSR_ERR_ARG;
        if (L.r_pool && ((L.H | L.W) & 1)) return FCVSR_ERR_ARG;
        t += L.r_pool ? (long long)B * (L.P / 4) * 16 : (long long)B * L.P * 16;
    }
    a.total = t;
    const unsigned grid = (unsigned)((t + 255) / 256);
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

// ---- BlockRCB cross-level sum (:766-777): x[b,y,x,:] += coef * r + mean2x2(td) + bilinear_x2(tu) (64 channels) -------------
struct MixLevel {
    const float* xin; float* xout; const void* r; const void* td; const void* tu; void* xout_r;
    float coef; int H, W; long long t_begin;
};
struct MixArgs { MixLevel lv[SC_MAX_LEV]; int nlev; int ldx, ldo, ldr; int round_main, op16, td_pooled; long long total; };

template <int R16, int T16>
__global__ void level_mix_kernel(const MixArgs a) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= a.total) return;
    const int l = (a.nlev > 1 && gi >= a.lv[1].t_begin) ? ((a.nlev > 2 && gi >= a.lv[2].t_begin) ? 2 : 1) : 0;
    const MixLevel& L = a.lv[l];
    const size_t i = (size_t)(gi - L.t_begin);
    const int H = L.H, W = L.W;
    const int c = (int)(i & 15) * 4;
    const size_t pix = i >> 4;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((size_t)W * H));
    const void* td = L.td;
    const void* tu = L.tu;
    const float coef = L.coef;
    float4 o = *reinterpret_cast<const float4*>(L.xin + pix * a.ldx + c);
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
    if (a.round_main) store_operand4(L.xout, pix * a.ldo + c, o, a.op16);      // xout itself is an operand-typed tensor
    else *reinterpret_cast<float4*>(L.xout + pix * a.ldo + c) = o;
}

// xin / xout / r / td / tu / xout_r: HOST arrays of nlev device pointers (td, tu, xout_r entries may be NULL); coef, H, W: HOST
// arrays.  Flags as fcvsr_level_mix.
extern "C" int fcvsr_level_mix_multi(int nlev, const float* const* xin, int ldx, float* const* xout, int ldo, const void* const* r,
                                     const float* coef, const void* const* td, const void* const* tu, int B, const int* H,
                                     const int* W, void* const* xout_r, int ldr, int round_main, int op16, int td_pooled,
                                     cudaStream_t st) {
    if (nlev < 1 || nlev > SC_MAX_LEV || !xin || !xout || !r || !coef || !H || !W || (ldx & 3) || (ldo & 3) || B <= 0) return FCVSR_ERR_ARG;
    MixArgs a;
    a.nlev = nlev; a.ldx = ldx; a.ldo = ldo; a.ldr = ldr; a.round_main = round_main; a.op16 = op16; a.td_pooled = td_pooled;
    long long t = 0;
    for (int l = 0; l < SC_MAX_LEV; ++l) {
        const int j = l < nlev ? l : 0;
        MixLevel& L = a.lv[l];
        L.xin = xin[j]; L.xout = xout[j]; L.r = r[j]; L.td = td ? td[j] : nullptr; L.tu = tu ? tu[j] : nullptr;
        L.xout_r = xout_r ? xout_r[j] : nullptr; L.coef = coef[j]; L.H = H[j]; L.W = W[j]; L.t_begin = t;
        if (l >= nlev) continue;
        if (!L.xin || !L.xout || !L.r || L.H <= 0 || L.W <= 0 || (L.tu && ((L.H | L.W) & 1)) || (L.xout_r && (ldr & 3))) return FCVSR_ERR_ARG;
        t += (long long)B * L.H * L.W * 16;
    }
    a.total = t;
    const unsigned grid = (unsigned)((t + 255) / 256);
#define LM_LAUNCH(R, T) level_mix_kernel<R, T><<<grid, 256, 0, st>>>(a)
    switch ((td_pooled >> 1) & 3) {
        case 0: LM_LAUNCH(0, 0); break;
        case 1: LM_LAUNCH(1, 0); break;
        case 2: LM_LAUNCH(0, 1); break;
        default: LM_LAUNCH(1, 1); break;
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

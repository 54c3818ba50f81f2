// One IAC iteration (IAC.forward, CVSR_train/arch/CVSR_freq.py:1230-1250; SAC :1253-1276) with the per-pixel taps computed on
// chip: the iteration's 64 -> 192 slice of the kernel predictor's last 1x1 convolution (MGAA.F.1, :1522-1523) is one
// 128 x 192 x 64 bf16 tcgen05 GEMM per tile whose accumulator (TMEM) IS the tap set of the tile, so the 6 x 192-channel
// `Pred_K` tensor of the reference (2.3 KB per pixel written, then read once per direction) never exists in HBM.
//
//   taps(y,x')   = F1_i . kp2(y,x') + b_i                            128 haloed pixels x 192 (tensor core, fp32 accumulate)
//   samp(y',x')  = bilinear(prev, x'+dx(y',x'), y'+dy(y',x'))        zeros outside, align_corners
//   v(y,x')      = sum_t K[t](y,x') * samp(clamp(y+t-1), x')         vertical pass, replicate pad
//   out(y,x)     = lrelu_0.1( sum_t K[t](y,x) * v(y, clamp(x+t-1)) + xin(y,x) )
//
// Tile = 8 rows x 14 output columns; with the one-column halo of the horizontal pass that is 8 x 16 = 128 pixels = the M of one
// MMA, so TMEM lane == haloed pixel.  A thread owns one haloed pixel and 16 channels (warp w: lane quarter w & 3 of TMEM, channel
// group w >> 2) and reads the taps from TMEM once per four channels (fp32, never rounded); the sample tile and the result tile
// live in shared memory pixel-major with a 68-float pitch, so that the 16-byte accesses of 8 consecutive pixels (one wavefront)
// fall into 8 different 4-bank groups.
// Order inside a CTA: operand tiles by cp.async -> sample geometry of both directions -> MMA issued -> per direction: bilinear
// gathers (the MMA runs under the first ones) -> both filter passes in one phase (vertical pass from the sample tile, the
// neighbouring columns of the horizontal pass by warp shuffle: a warp holds two tile rows of 16 columns) -> coalesced output
// stage.  Both directions share the taps (:1526-1527 call IAC with the same Pred_K), so one CTA serves both and kp2 is read once.
// The fp32-contract mode runs the same kernel with TF32 operands (template parameter TF: fp32 kp2 / F1 tiles, kind::tf32 MMAs,
// fp32 prev / next tensors).
#include "tc_common.cuh"

#define IT_TH 8
#define IT_TW 14
#define IT_HX 16
#define IT_HY (IT_TH + 2)
#define IT_HALO (IT_HY * IT_HX)
#define IT_THREADS 512
#define IT_C 64
#define IT_P 68                            // pixel pitch (floats) of the sample / vertical-pass tiles
#define IT_N 192
#define IT_OFF_W 16384                     // F1_i tile (192 rows x 128 B, SW128) after the kp2 tile (128 rows x 128 B)
#define IT_OFF_SAMP (IT_HALO * IT_P * 4)    // two [160][68] fp32 tile buffers; the first one starts life as the two operand tiles
#define IT_OFF_GEO (IT_OFF_SAMP + IT_HALO * IT_P * 4)
#define IT_OFF_BIAS (IT_OFF_GEO + 2 * IT_HALO * 32)
#define IT_OFF_BAR IT_OFF_BIAS
#define IT_SMEM (IT_OFF_BAR + 16 + 1024)   // + alignment slack of the 1024-byte swizzle atom

struct IacTcArgs {
    const void* prev[2]; int ldprev[2];    // fp32 (iteration 0) or bf16 (prev16) NHWC, ld in elements
    const float* xin[2]; int ldxin[2];
    void* next[2]; int ldnext[2];          // bf16 NHWC
    const float* offs; int ldoffs; int offs_ch[2];
    const unsigned short* kp; int ldkp;    // kp2: bf16 [B,H,W,ldkp], 64 channels
    const unsigned short* w;               // this iteration's F1 slice: bf16 [192][64], row n = c4*12 + t*4 + cc  (c = 4 c4 + cc)
    const float* bias;                     // [192], same order
    int B, H, W, prev16;
    int round_tf32;                        // TF variant: outputs are TF32-rounded fp32 (the conv3 operand of the last iteration)
};

__device__ __forceinline__ float4 it_bf16x4(uint2 u) {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void it_cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void it_tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// the 3 x 4 taps of four consecutive channels: columns [c4*12, c4*12 + 12) of the thread's lane
__device__ __forceinline__ void it_tmem_ld12(uint32_t taddr, float* k) {
    uint32_t r[12];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%12];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4,%5,%6,%7}, [%12 + 4];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8,%9,%10,%11}, [%12 + 8];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 12; ++i) k[i] = __uint_as_float(r[i]);
}
template <bool P16> struct ItCorner;
template <> struct ItCorner<true> {
    uint2 r;
    __device__ __forceinline__ void load(const void* base, size_t off) {
        r = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(base) + off));
    }
    __device__ __forceinline__ float4 get() const { return it_bf16x4(r); }
};
template <> struct ItCorner<false> {
    float4 r;
    __device__ __forceinline__ void load(const void* base, size_t off) {
        r = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
    }
    __device__ __forceinline__ float4 get() const { return r; }
};
// N halo pixels (hp0, hp0 + 32, ...) of one half-warp: all 4 N corner gathers in flight, then the blends.
// Corner order 00, 01, 10, 11 with fmaf: the same arithmetic as fcvsr_iac_step.
template <bool P16, int N>
__device__ __forceinline__ void it_gather(const void* pbase, int ldp, const int4* geo_i, const float4* geo_w, float* samp_c0,
                                          int hp0) {
    ItCorner<P16> c[N][4];
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const int4 gi = geo_i[hp0 + n * (IT_THREADS / 16)];
        c[n][0].load(pbase, (size_t)gi.x * ldp);
        c[n][1].load(pbase, (size_t)gi.y * ldp);
        c[n][2].load(pbase, (size_t)gi.z * ldp);
        c[n][3].load(pbase, (size_t)gi.w * ldp);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const float4 w = geo_w[hp0 + n * (IT_THREADS / 16)];
        const float4 a0 = c[n][0].get(), a1 = c[n][1].get(), a2 = c[n][2].get(), a3 = c[n][3].get();
        float4 s;
        s.x = fmaf(w.w, a3.x, fmaf(w.z, a2.x, fmaf(w.y, a1.x, w.x * a0.x)));
        s.y = fmaf(w.w, a3.y, fmaf(w.z, a2.y, fmaf(w.y, a1.y, w.x * a0.y)));
        s.z = fmaf(w.w, a3.z, fmaf(w.z, a2.z, fmaf(w.y, a1.z, w.x * a0.z)));
        s.w = fmaf(w.w, a3.w, fmaf(w.z, a2.w, fmaf(w.y, a1.w, w.x * a0.w)));
        *reinterpret_cast<float4*>(samp_c0 + (hp0 + n * (IT_THREADS / 16)) * IT_P) = s;
    }
}

// TF: the fp32-contract mode -- kp2 and the F1 slice are TF32-rounded fp32 tensors (two 32-channel SWIZZLE_128B chunks each, eight
// kind::tf32 MMAs), prev / next are fp32.  The 80 KB of operand tiles then cover both tile buffers, so the first gathers wait for
// the accumulator instead of running under the MMAs.
template <bool P16, bool TF>
__global__ void __launch_bounds__(IT_THREADS, 2) iac_step_tc_kernel(IacTcArgs a) {
    extern __shared__ uint8_t it_smem_raw[];
    // (pointer + offset, not an integer round trip: the compiler must keep seeing the shared address space)
    uint8_t* sm = it_smem_raw + ((1024u - (smem_u32(it_smem_raw) & 1023u)) & 1023u);
    float* buf0 = reinterpret_cast<float*>(sm);                               // [IT_HALO][IT_P]; first holds the two operand tiles
    float* buf1 = reinterpret_cast<float*>(sm + IT_OFF_SAMP);                 // [IT_HALO][IT_P]
    int4* geo_i = reinterpret_cast<int4*>(sm + IT_OFF_GEO);                   // [2 directions][IT_HALO]
    float4* geo_w = reinterpret_cast<float4*>(sm + IT_OFF_GEO + 2 * IT_HALO * 16);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + IT_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + IT_OFF_BAR + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (a.W + IT_TW - 1) / IT_TW;
    const int ty0 = (blockIdx.x / tiles_x) * IT_TH, tx0 = (blockIdx.x % tiles_x) * IT_TW;
    const int b = blockIdx.y;
    const int H = a.H, W = a.W;
    const size_t img = (size_t)b * H * W;
    // passes 2/3 ownership: TMEM lane = haloed pixel, 16 channels per warp group
    const int px = (warp & 3) * 32 + lane, g = warp >> 2;
    const int row = px >> 4, hx = px & 15;
    const bool owns_out = hx >= 1 && hx <= IT_TW && tx0 + hx - 1 < W && ty0 + row < H;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // operand tiles, 16-byte chunks XOR-swizzled inside their 128-byte row (SWIZZLE_128B, K-major): the kp2 rows of the 128
    // haloed pixels (clamped == replicate padding of both passes) and the 192 weight rows of this iteration
    if (TF) {
        const uint32_t sA = smem_u32(sm), sW = smem_u32(sm + 2 * IT_OFF_W);
        const float* kpf = reinterpret_cast<const float*>(a.kp);
        const float* wf = reinterpret_cast<const float*>(a.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {       // 128 pixels x 16 chunks of 4 floats; K chunk kc = channels [32 kc, 32 kc + 32)
            const int q = tid + j * IT_THREADS, ppx = q >> 4, ch = q & 15, kc = ch >> 3, c8 = ch & 7;
            const int yy = min(ty0 + (ppx >> 4), H - 1), xx = min(max(tx0 - 1 + (ppx & 15), 0), W - 1);
            it_cp16(sA + kc * IT_OFF_W + ppx * 128 + ((c8 ^ (ppx & 7)) << 4), kpf + (img + (size_t)yy * W + xx) * a.ldkp + ch * 4);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {       // 192 rows x 16 chunks
            const int q = tid + j * IT_THREADS, n = q >> 4, ch = q & 15, kc = ch >> 3, c8 = ch & 7;
            it_cp16(sW + kc * (IT_N * 128) + n * 128 + ((c8 ^ (n & 7)) << 4), wf + n * 64 + ch * 4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        const uint32_t sA = smem_u32(sm), sW = smem_u32(sm + IT_OFF_W);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int q = tid + j * IT_THREADS, ppx = q >> 3, ch = q & 7;
            const int yy = min(ty0 + (ppx >> 4), H - 1), xx = min(max(tx0 - 1 + (ppx & 15), 0), W - 1);
            it_cp16(sA + ppx * 128 + ((ch ^ (ppx & 7)) << 4), a.kp + (img + (size_t)yy * W + xx) * a.ldkp + ch * 8);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int q = tid + j * IT_THREADS, n = q >> 3, ch = q & 7;
            it_cp16(sW + n * 128 + ((ch ^ (n & 7)) << 4), a.w + n * 64 + ch * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {   // L2 prefetch of the residual rows (needed last) while the offsets are on their way
        const int d = tid >> 8, ppx = (tid >> 1) & 127, half = tid & 1;
        const int py = ty0 + (ppx >> 4), pxx = tx0 + (ppx & 15) - 1;
        if ((ppx & 15) >= 1 && (ppx & 15) <= IT_TW && pxx < W && py < H) {
            const float* xp = (d ? a.xin[1] : a.xin[0]) + (img + (size_t)py * W + pxx) * (d ? a.ldxin[1] : a.ldxin[0]) + half * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xp));
        }
    }
    // phase 0: sample geometry of the 10 x 16 haloed tile for both directions, one (direction, halo pixel) per thread
    if (tid < 2 * IT_HALO) {
        const int d = tid >= IT_HALO, hp = tid - d * IT_HALO;
        const int hy = hp >> 4, hhx = hp & 15;
        const int yy = min(max(ty0 - 1 + hy, 0), H - 1), xx = min(max(tx0 - 1 + hhx, 0), W - 1);
        {   // the displacements are a few pixels: most gathers land in the tile's own neighbourhood, so start pulling it into L2
            const size_t po = (img + (size_t)yy * W + xx) * (d ? a.ldprev[1] : a.ldprev[0]);
            if (P16) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const unsigned short*>(d ? a.prev[1] : a.prev[0]) + po));
            } else {
                const float* pp = reinterpret_cast<const float*>(d ? a.prev[1] : a.prev[0]) + po;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 32));
            }
        }
        const float2 dl = *reinterpret_cast<const float2*>(a.offs + (img + (size_t)yy * W + xx) * a.ldoffs + (d ? a.offs_ch[1] : a.offs_ch[0]));
        const float sx = (float)xx + dl.x, sy = (float)yy + dl.y;
        int4 gi = make_int4(0, 0, 0, 0);
        float4 gw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sx > -1.f && sx < (float)W && sy > -1.f && sy < (float)H) {        // also rejects NaN / inf offsets
            const float fx0 = floorf(sx), fy0 = floorf(sy);
            const float lx = sx - fx0, ly = sy - fy0;
            const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
            const bool vx0 = x0 >= 0, vx1 = x1 < W, vy0 = y0 >= 0, vy1 = y1 < H;
            const int r0 = (vy0 ? y0 : 0) * W, r1 = (vy1 ? y1 : H - 1) * W, q0 = vx0 ? x0 : 0, q1 = vx1 ? x1 : W - 1;
            gi = make_int4(r0 + q0, r0 + q1, r1 + q0, r1 + q1);
            gw = make_float4((vy0 && vx0) ? (1.f - ly) * (1.f - lx) : 0.f, (vy0 && vx1) ? (1.f - ly) * lx : 0.f,
                             (vy1 && vx0) ? ly * (1.f - lx) : 0.f, (vy1 && vx1) ? ly * lx : 0.f);
        }
        geo_i[tid] = gi;
        geo_w[tid] = gw;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // cp.async / generic writes -> tensor-core (async proxy) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + g * 48;
    {   // the accumulator starts as the bias (every MMA below accumulates), so the tap read-out later is a plain TMEM load
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            uint32_t r[16];
            const uint4* bb = reinterpret_cast<const uint4*>(a.bias + g * 48 + t * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 bv = __ldg(bb + j);
                r[4 * j] = bv.x; r[4 * j + 1] = bv.y; r[4 * j + 2] = bv.z; r[4 * j + 3] = bv.w;
            }
            it_tmem_st16(taddr + t * 16, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one()) {
        if (TF) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(IT_N >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const uint64_t a_d = make_desc(smem_u32(sm + kc * IT_OFF_W));
                const uint64_t b_d = make_desc(smem_u32(sm + 2 * IT_OFF_W + kc * (IT_N * 128)));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_d + 2 * k, b_d + 2 * k, idesc, 1u);
            }
        } else {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(IT_N >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t a_d = make_desc(smem_u32(sm)), b_d = make_desc(smem_u32(sm + IT_OFF_W));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base, a_d + 2 * k, b_d + 2 * k, idesc, 1u);
        }
        umma_commit(bar);
    }
    __syncwarp();

    // output stage ownership: a half-warp per output pixel (128 contiguous bytes of bf16), a lane 4 channels
    const int oslot = tid >> 4, oc0 = (tid & 15) * 4;
    // Both directions use the same two tile buffers: buf1 = warped samples (phase 1 -> 2), buf0 = pass results (phase 2 -> 3;
    // it starts life as the MMA's operand tiles, which are dead once the accumulator barrier has been waited for).  Two
    // barriers per direction: the next direction's gathers overwrite the samples only after every thread has passed the
    // barrier behind phase 2, and its phase 2 overwrites the results only after the barrier behind its own gathers, which
    // every thread reaches after its phase 3 reads.
    float* bufS = buf1;
    float* bufO = buf0;
    if (TF) {       // the operand tiles reach into the sample buffer: nothing may be gathered before the MMAs have read them
        mbar_wait(bar, 0, nullptr, 0);
        tc_fence_after();
    }
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
        // phase 1: warped samples; a half-warp owns five halo pixels, a lane 4 channels
        {
            const int ldp = (dir ? a.ldprev[1] : a.ldprev[0]);
            const void* pbase = P16 ? (const void*)(reinterpret_cast<const unsigned short*>((dir ? a.prev[1] : a.prev[0])) + img * ldp + oc0)
                                    : (const void*)(reinterpret_cast<const float*>((dir ? a.prev[1] : a.prev[0])) + img * ldp + oc0);
            const int4* gi = geo_i + dir * IT_HALO;
            const float4* gw = geo_w + dir * IT_HALO;
            if (P16) {          // 12 + 8 eight-byte gathers in flight
                it_gather<P16, 3>(pbase, ldp, gi, gw, bufS + oc0, oslot);
                it_gather<P16, 2>(pbase, ldp, gi, gw, bufS + oc0, oslot + 3 * (IT_THREADS / 16));
            } else {            // 8 + 8 + 4 sixteen-byte gathers
                it_gather<P16, 2>(pbase, ldp, gi, gw, bufS + oc0, oslot);
                it_gather<P16, 2>(pbase, ldp, gi, gw, bufS + oc0, oslot + 2 * (IT_THREADS / 16));
                it_gather<P16, 1>(pbase, ldp, gi, gw, bufS + oc0, oslot + 4 * (IT_THREADS / 16));
            }
        }
        __syncthreads();
        // residual input of this thread's output-stage pixels (coalesced: 256 contiguous bytes per half-warp), used in phase 3
        float4 xi[4];
        {
            const float* xin = (dir ? a.xin[1] : a.xin[0]) + oc0;
            const int ldx = dir ? a.ldxin[1] : a.ldxin[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = oslot + 32 * i, orow = o / IT_TW, ocol = o - orow * IT_TW;
                const int yy = ty0 + orow, xx = tx0 + ocol;
                if (o < IT_TH * IT_TW && yy < H && xx < W) xi[i] = __ldg(reinterpret_cast<const float4*>(xin + (img + (size_t)yy * W + xx) * ldx));
            }
        }
        if (!TF && dir == 0) mbar_wait(bar, 0, nullptr, 0);
        tc_fence_after();
        // phase 2: both SAC passes with the taps read from TMEM once (fp32, never rounded).  Vertical pass: halo rows row,
        // row + 1, row + 2 of the thread's column from the sample tile (halo row hy = row + t <-> image row y + t - 1).
        // Horizontal pass: the vertical-pass values of the two neighbouring columns are in the neighbouring lanes (a warp holds
        // two tile rows of 16 columns, lane & 15 = column), so they come by shuffle -- no second tile buffer, no barrier between
        // the passes, half the shared-memory traffic of the passes.  tcgen05.ld and the shuffles are warp-collective: every lane
        // runs them, only the owners of an output pixel store.
        {
            const float* sp = bufS + (row * IT_HX + hx) * IT_P + g * 16;
            float* ob = bufO + px * IT_P + g * 16;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float k[12];
                it_tmem_ld12(taddr + j * 12, k);
                const float4 s0 = *reinterpret_cast<const float4*>(sp + 4 * j);
                const float4 s1 = *reinterpret_cast<const float4*>(sp + IT_HX * IT_P + 4 * j);
                const float4 s2 = *reinterpret_cast<const float4*>(sp + 2 * IT_HX * IT_P + 4 * j);
                float4 v;
                v.x = fmaf(k[8], s2.x, fmaf(k[4], s1.x, k[0] * s0.x));
                v.y = fmaf(k[9], s2.y, fmaf(k[5], s1.y, k[1] * s0.y));
                v.z = fmaf(k[10], s2.z, fmaf(k[6], s1.z, k[2] * s0.z));
                v.w = fmaf(k[11], s2.w, fmaf(k[7], s1.w, k[3] * s0.w));
                float4 vl, vr;
                vl.x = __shfl_up_sync(0xffffffffu, v.x, 1, 16); vr.x = __shfl_down_sync(0xffffffffu, v.x, 1, 16);
                vl.y = __shfl_up_sync(0xffffffffu, v.y, 1, 16); vr.y = __shfl_down_sync(0xffffffffu, v.y, 1, 16);
                vl.z = __shfl_up_sync(0xffffffffu, v.z, 1, 16); vr.z = __shfl_down_sync(0xffffffffu, v.z, 1, 16);
                vl.w = __shfl_up_sync(0xffffffffu, v.w, 1, 16); vr.w = __shfl_down_sync(0xffffffffu, v.w, 1, 16);
                if (owns_out) {
                    float4 o;
                    o.x = fmaf(k[8], vr.x, fmaf(k[4], v.x, k[0] * vl.x));
                    o.y = fmaf(k[9], vr.y, fmaf(k[5], v.y, k[1] * vl.y));
                    o.z = fmaf(k[10], vr.z, fmaf(k[6], v.z, k[2] * vl.z));
                    o.w = fmaf(k[11], vr.w, fmaf(k[7], v.w, k[3] * vl.w));
                    *reinterpret_cast<float4*>(ob + 4 * j) = o;
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // phase 3: + residual, LeakyReLU(0.1), bf16, coalesced stores (the thread-per-pixel layout of the passes would touch 32
        // different lines per load / store instruction)
        {
            unsigned short* next = reinterpret_cast<unsigned short*>(dir ? a.next[1] : a.next[0]) + oc0;
            const int ldn = dir ? a.ldnext[1] : a.ldnext[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = oslot + 32 * i, orow = o / IT_TW, ocol = o - orow * IT_TW;
                const int yy = ty0 + orow, xx = tx0 + ocol;
                if (o < IT_TH * IT_TW && yy < H && xx < W) {
                    const float4 r = *reinterpret_cast<const float4*>(bufO + (orow * IT_HX + ocol + 1) * IT_P + oc0);
                    const float p0 = r.x + xi[i].x, p1 = r.y + xi[i].y, p2 = r.z + xi[i].z, p3 = r.w + xi[i].w;
                    if (TF) {
                        float4 v = make_float4(p0 >= 0.f ? p0 : 0.1f * p0, p1 >= 0.f ? p1 : 0.1f * p1, p2 >= 0.f ? p2 : 0.1f * p2,
                                               p3 >= 0.f ? p3 : 0.1f * p3);
                        if (a.round_tf32) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
                        float* nf = reinterpret_cast<float*>(dir ? a.next[1] : a.next[0]) + oc0;
                        *reinterpret_cast<float4*>(nf + (img + (size_t)yy * W + xx) * ldn) = v;
                        continue;
                    }
                    uint32_t lo, hi;
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(p1 >= 0.f ? p1 : 0.1f * p1), "f"(p0 >= 0.f ? p0 : 0.1f * p0));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p3 >= 0.f ? p3 : 0.1f * p3), "f"(p2 >= 0.f ? p2 : 0.1f * p2));
                    *reinterpret_cast<uint2*>(next + (img + (size_t)yy * W + xx) * ldn) = make_uint2(lo, hi);
                }
            }
        }
        // no barrier here: the next direction's gathers write the sample buffer, last read before the barrier above
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

extern "C" int fcvsr_iac_step_tc(const void* prev_f, int ldprev_f, const void* prev_b, int ldprev_b, int prev16,
                                 const float* xin_f, int ldxin_f, const float* xin_b, int ldxin_b, void* next_f, int ldnext_f,
                                 void* next_b, int ldnext_b, const float* offs, int ldoffs, int ch_f, int ch_b, const void* kp,
                                 int ldkp, const void* w, const float* bias, int B, int H, int W, cudaStream_t st) {
    if (!prev_f || !prev_b || !xin_f || !xin_b || !next_f || !next_b || !offs || !kp || !w || !bias) return FCVSR_ERR_ARG;
    if (B <= 0 || H <= 0 || W <= 0) return FCVSR_ERR_ARG;
    const bool tf = (prev16 & 2) != 0;                 // fp32-contract mode: fp32 (TF32-rounded) kp / w, fp32 prev / next
    if (tf && (prev16 & 1)) return FCVSR_ERR_ARG;
    if (((ldprev_f | ldprev_b | ldxin_f | ldxin_b) & 3) || ((ldnext_f | ldnext_b | ldkp) & (tf ? 3 : 7)) || ((ldoffs | ch_f | ch_b) & 1))
        return FCVSR_ERR_ARG;
    if (((uintptr_t)xin_f | (uintptr_t)xin_b | (uintptr_t)next_f | (uintptr_t)next_b | (uintptr_t)kp | (uintptr_t)w) & 15)
        return FCVSR_ERR_ARG;
    if (((uintptr_t)prev_f | (uintptr_t)prev_b) & ((prev16 & 1) ? 7 : 15)) return FCVSR_ERR_ARG;
    if (((uintptr_t)offs & 7) || ((uintptr_t)bias & 3)) return FCVSR_ERR_ARG;
    IacTcArgs a;
    a.prev[0] = prev_f; a.prev[1] = prev_b; a.ldprev[0] = ldprev_f; a.ldprev[1] = ldprev_b;
    a.xin[0] = xin_f; a.xin[1] = xin_b; a.ldxin[0] = ldxin_f; a.ldxin[1] = ldxin_b;
    a.next[0] = next_f; a.next[1] = next_b; a.ldnext[0] = ldnext_f; a.ldnext[1] = ldnext_b;
    a.offs = offs; a.ldoffs = ldoffs; a.offs_ch[0] = ch_f; a.offs_ch[1] = ch_b;
    a.kp = reinterpret_cast<const unsigned short*>(kp); a.ldkp = ldkp;
    a.w = reinterpret_cast<const unsigned short*>(w); a.bias = bias;
    a.B = B; a.H = H; a.W = W; a.prev16 = prev16 & 1; a.round_tf32 = (prev16 & 4) ? 1 : 0;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(iac_step_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IT_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(iac_step_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IT_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(iac_step_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IT_SMEM) != cudaSuccess)
            return FCVSR_ERR_CUDA;
        attr_set = true;
    }
    dim3 grid(((H + IT_TH - 1) / IT_TH) * ((W + IT_TW - 1) / IT_TW), B, 1);
    if (tf) iac_step_tc_kernel<false, true><<<grid, IT_THREADS, IT_SMEM, st>>>(a);
    else if (prev16 & 1) iac_step_tc_kernel<true, false><<<grid, IT_THREADS, IT_SMEM, st>>>(a);
    else iac_step_tc_kernel<false, false><<<grid, IT_THREADS, IT_SMEM, st>>>(a);
    return fcvsr_launch_status();
}

// tcgen05 3x3 convolution, second generation: weights RESIDENT in shared memory, single-copy haloed
// input tile.  Same math / epilogue contract as conv_tc.cu (NHWC, TF32 or bf16 operands, fp32 accumulate in
// TMEM); restricted to the shape class that dominates the FCVSR trunk: k = 3, stride 1, Cout % 64 == 0 and a
// weight slab that fits next to the input ring (see tc2_plan below).
//
// Why (profiles/r1_conv_tc_full_summary.txt, profiles/r1_notes.md): conv_tc.cu re-streams the weight tile for
// every 128-pixel tile and loads three shifted copies of the input window, so a 64->64 conv moves 8.4x its input
// bytes from L2 to shared memory and runs at the L2 fabric limit (~10 TB/s), not at the tensor pipe.  Here:
//   * the [NP x 9*Cin] weight slab of the current NP-column pass (NP = 64 or 128) is TMA-loaded ONCE per CTA
//     (SWIZZLE_128B) and stays resident across all tiles of the pass;
//   * the A operand uses the NO-SWIZZLE K-major canonical layout with SBO = 128 B: row r of a 16-byte
//     K-granule plane sits at plane + 16*r, i.e. rows are affine, so filter tap (ky,kx) of the haloed
//     (8+2) x 16 window is the SAME tile viewed from row ky*16 + kx -- one copy serves all nine taps (1.43x the
//     input bytes instead of 3.75x + weights).  The M = 128 rows of a tile are the 8 x 16 window positions;
//     columns 14, 15 of each row wrap into the next row's halo and are discarded, so a tile yields 8 x 14 output
//     pixels (12.5 % padding work, paid to keep every tap a pure descriptor shift);
//   * that layout cannot be written by TMA with 128-byte rows, so four producer warps fill it with 16-byte
//     cp.async (zero-fill outside the image = conv padding), completion tracked by
//     cp.async.mbarrier.arrive.noinc on the stage's full barrier; the ring is as deep as shared memory allows
//     (3 stages next to a 147 KB slab, 7 next to a 73 KB one);
//   * one thread issues the 36 MMAs of a 128-byte K chunk with constant-add descriptors, one commit per chunk.
#include "tc_common.cuh"
#include <stdlib.h>

#define T2_TH 8
#define T2_TWP 16                      // window columns per tile row (M = 8 x 16)
#define T2_TWV 14                      // valid output columns per tile
#define T2_PLANE 2608                  // bytes per 16-byte-granule plane: 163 rows (bank rotation 12 words)
#define T2_STAGE (8 * T2_PLANE)        // one 128-byte K chunk (32 fp32 / 64 bf16 channels) of a haloed tile
#define T2_MAXSTAGE 8
#define T2_PROD_WARPS 2
#define T2_EPI_WARPS 16
#define T2_MMA_WARP T2_PROD_WARPS
#define T2_EPI0 (T2_PROD_WARPS + 1)     // first epilogue warp
#define T2_THREADS (32 * (T2_PROD_WARPS + 1 + T2_EPI_WARPS))
#define T2_SMEM_MAX (227 * 1024)

#ifdef T2_TRACE      // bring-up only (tools/gpu_conv_trace.py builds its own copy of this file with -DT2_TRACE)
__device__ long long t2_trace[64 * 16];
#define T2_STAMP(n, slot) do { if (blockIdx.x == 0 && (n) < 64) t2_trace[(n) * 16 + (slot)] = clock64(); } while (0)
#define T2_STAMP_NS(n) do { if (blockIdx.x == 0) { unsigned long long ns_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_)); \
        t2_trace[(n) * 16 + 15] = (long long)ns_; t2_trace[(n) * 16 + 14] = clock64(); } } while (0)
extern "C" int fcvsr_debug_conv_trace(long long* host, int n) {
    return cudaMemcpyFromSymbol(host, t2_trace, sizeof(long long) * (n < 1024 ? n : 1024)) == cudaSuccess ? 0 : 1;
}
#else
#define T2_STAMP(n, slot) do {} while (0)
#define T2_STAMP_NS(n) do {} while (0)
#endif

struct ConvTc2Params {
    const void* x; int ldx;            // elements
    EpiArgs e;
    int B, H, W, Cin, Cout, kch, npass, np;   // kch: 128-byte K chunks of Cin; np: columns per pass (64 / 128)
    int nstage;
    int tiles_x, tiles_y, tiles;       // tiles per pass
    const float* slope_ptr;
    int* err;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: 8-row core matrices 128 B apart along M (SBO), 16-byte K granules T2_PLANE apart (LBO)
__device__ __forceinline__ uint64_t make_desc_a(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(T2_PLANE >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

template <bool BF16>
__global__ void __launch_bounds__(T2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const ConvTc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int ESZ = BF16 ? 2 : 4;
    constexpr int KCH = 128 / ESZ;                         // channels per 128-byte chunk
    const uint32_t wchunk = (uint32_t)p.np * 128;          // weight bytes per (tap, chunk): np rows x 128 B
    const uint32_t w_bytes = (uint32_t)(9 * p.kch) * wchunk;
    uint8_t* w_buf = smem;
    uint8_t* a_buf = smem + w_bytes;
    uint64_t* bars = (uint64_t*)(a_buf + p.nstage * T2_STAGE + 64);
    uint64_t* a_full = bars;                       // [MAXSTAGE] 128 cp.async arrivals
    uint64_t* a_empty = bars + T2_MAXSTAGE;        // [MAXSTAGE] 1 (tcgen05.commit)
    uint64_t* tm_full = bars + 2 * T2_MAXSTAGE;    // [2]
    uint64_t* tm_empty = tm_full + 2;              // [2] 4 epilogue warps
    uint64_t* w_full = tm_empty + 2;               // weights of the current pass landed
    uint64_t* w_free = w_full + 1;                 // all MMAs of the pass retired (weights may be overwritten)
    uint32_t* tmem_slot = (uint32_t*)(w_free + 1);
    float* bias_s = (float*)((uint8_t*)bars + 256);          // see conv_tc.cu: per-lane global bias loads starve behind the MMAs
    for (int i = threadIdx.x; i < p.Cout; i += T2_THREADS) bias_s[i] = p.e.bias ? p.e.bias[i] : 0.f;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = p.np == 64 ? 128u : 256u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < T2_MAXSTAGE; ++i) { mbar_init(&a_full[i], 32 * T2_PROD_WARPS); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], T2_EPI_WARPS); }
        mbar_init(w_full, 1);
        mbar_init(w_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T2_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) T2_STAMP_NS(62);

    if (warp < T2_PROD_WARPS) {
        // ===== A producers: thread i owns granule plane (i & 7) of window columns (i >> 3) + 8h =====
        const int i = threadIdx.x;
        constexpr int CPT = 16 / (4 * T2_PROD_WARPS);           // window columns per thread
        const int plane = i & 7;
        int stage = 0; uint32_t phase = 0;
        int tn = 0;
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(p.x);
        const size_t pix_bytes = (size_t)p.ldx * ESZ;
        for (int pass = 0; pass < p.npass; ++pass)
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, r_ = t / p.tiles_x;
                const int ty = r_ % p.tiles_y, b = r_ / p.tiles_y;
                const int y0 = ty * T2_TH - 1;
                const uint8_t* img = xb + (size_t)b * p.H * p.W * pix_bytes;
                for (int kc = 0; kc < p.kch; ++kc) {
                    mbar_wait_warp(&a_empty[stage], phase ^ 1, p.err, 11);
                    if (i == 0 && kc == 0) T2_STAMP(tn, 0);
#pragma unroll
                    for (int h = 0; h < CPT; ++h) {
                        const int col = (i >> 3) + h * (4 * T2_PROD_WARPS);
                        const int xx = tx * T2_TWV - 1 + col;
                        const bool xok = xx >= 0 && xx < p.W;
                        const uint32_t dst0 = smem_u32(a_buf + stage * T2_STAGE) + plane * T2_PLANE + col * 16;
                        const uint8_t* src0 = img + (size_t)xx * pix_bytes + kc * 128 + plane * 16;
#pragma unroll
                        for (int j = 0; j < T2_TH + 2; ++j) {
                            const int yy = y0 + j;
                            const bool ok = xok && yy >= 0 && yy < p.H;
                            const void* src = ok ? (const void*)(src0 + (size_t)yy * p.W * pix_bytes) : p.x;
                            cp_async16(dst0 + j * (T2_TWP * 16), src, ok ? 16u : 0u);
                        }
                    }
                    cp_async_arrive_noinc(&a_full[stage]);
                    if (i == 0 && kc == p.kch - 1) T2_STAMP(tn, 1);
                    if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                }
                ++tn;
            }
    } else if (warp == T2_MMA_WARP) {
        // ===== weight loader + MMA issuer (one lane) =====
        if (elect_one()) {
            constexpr uint32_t FMT = BF16 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.np >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t w_desc0 = make_desc(smem_u32(w_buf));
            const uint32_t wstep = wchunk >> 4;
            int sa = 0; uint32_t pa = 0;
            int acc = 0; uint32_t pacc = 0;
            int tn = 0;
            uint32_t pre_t = 0, pre_a = 0;                  // early polls of the upcoming step's barriers (1 = already complete)
            for (int pass = 0; pass < p.npass; ++pass) {
                if (pass > 0) mbar_wait(w_free, (uint32_t)((pass - 1) & 1), p.err, 12);
                mbar_expect_tx(w_full, w_bytes);
                for (int kc = 0; kc < p.kch; ++kc)
                    for (int tap = 0; tap < 9; ++tap)
                        tma_load_2d(w_buf + (kc * 9 + tap) * wchunk, &map_w, w_full, tap * p.Cin + kc * KCH, pass * p.np);
                mbar_wait(w_full, (uint32_t)(pass & 1), p.err, 13);
                tc_fence_after();
                for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                    if (!pre_t) mbar_wait(&tm_empty[acc], pacc ^ 1, p.err, 14);
                    pre_t = 0;
                    tc_fence_after();
                    T2_STAMP(tn, 4);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.np);
                    for (int kc = 0; kc < p.kch; ++kc) {
                        if (!pre_a) mbar_wait(&a_full[sa], pa, p.err, 15);
                        pre_a = 0;
                        tc_fence_after();
                        T2_STAMP(tn, kc == 0 ? 5 : 6);
                        const uint64_t a_desc0 = make_desc_a(smem_u32(a_buf + sa * T2_STAGE));
                        const uint64_t b_desc0 = w_desc0 + (uint64_t)(kc * 9 * wstep);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint64_t a_d = a_desc0 + (uint64_t)(ky * T2_TWP + kx);          // 16 B per row
                                const uint64_t b_d = b_desc0 + (uint64_t)((ky * 3 + kx) * wstep);
                                if (ky == 2 && kx == 2) {
                                    // last tap: the NEXT step's barriers are polled in the same asm block as its MMAs
                                    // (a shared-memory round trip takes 250-400 clk under MMA load, see tc_common.cuh)
                                    const int nsa = sa + 1 == p.nstage ? 0 : sa + 1;
                                    const uint32_t ok = umma_x4_poll3<BF16>(d_tmem, a_d, (uint64_t)(2 * (T2_PLANE >> 4)), b_d, idesc, 1u,
                                                                            &a_full[nsa], nsa ? pa : pa ^ 1, &tm_empty[acc ^ 1],
                                                                            acc ? pacc : pacc ^ 1, &a_full[nsa], nsa ? pa : pa ^ 1);
                                    pre_a = ok & 1;
                                    if (kc == p.kch - 1) pre_t = ok & 2;
                                } else {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const uint64_t ad = a_d + (uint64_t)(k * 2 * (T2_PLANE >> 4)), bd = b_d + 2 * k;
                                        if (BF16) umma_f16(d_tmem, ad, bd, idesc, (kc | ky | kx | k) ? 1u : 0u);
                                        else umma_tf32(d_tmem, ad, bd, idesc, (kc | ky | kx | k) ? 1u : 0u);
                                    }
                                }
                            }
                        umma_commit(&a_empty[sa]);
                        if (++sa == p.nstage) { sa = 0; pa ^= 1; }
                    }
                    umma_commit(&tm_full[acc]);
                    T2_STAMP(tn, 7);
                    ++tn;
                    if (++acc == 2) { acc = 0; pacc ^= 1; }
                }
                umma_commit(w_free);
            }
        }
    } else {
        // ===== epilogue warps 5..: TMEM lane quarter = warp % 4; the warps of a quarter split the columns =====
        const int q = warp & 3;
        const int eh = (warp - T2_EPI0) >> 2;
        const int nchunk = p.np >> 4, cper = nchunk / (T2_EPI_WARPS / 4);
        const int c_begin = eh * cper, c_end = c_begin + cper;
        const int m = q * 32 + lane;
        const int ly = m >> 4, lx = m & 15;
        EpiArgs e = p.e;
        e.bias_sa = smem_u32(bias_s);
        if (e.act == FCVSR_ACT_PRELU) e.slope = p.slope_ptr[0];
        int acc = 0; uint32_t pacc = 0;
        int tn = 0;
        for (int pass = 0; pass < p.npass; ++pass)
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, r_ = t / p.tiles_x;
                const int ty = r_ % p.tiles_y, b = r_ / p.tiles_y;
                const int y = ty * T2_TH + ly, x = tx * T2_TWV + lx;
                const bool valid = lx < T2_TWV && y < p.H && x < p.W;
                const size_t pix = ((size_t)b * p.H + y) * p.W + x;
                mbar_wait_warp(&tm_full[acc], pacc, p.err, 16);
                tc_fence_after();
                if (warp == T2_EPI0 && lane == 0) T2_STAMP(tn, 8);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.np);
                uint32_t ra[16], rb[16];
                tmem_ld16(taddr + c_begin * 16, ra);
                for (int c = c_begin; c < c_end; c += 2) {      // tcgen05.ld of chunk c+1 in flight while chunk c is stored
                    tmem_ld_wait();
                    if (warp == T2_EPI0 && lane == 0 && c == 0) T2_STAMP(tn, 10);
                    const bool has_b = c + 1 < c_end;
                    if (has_b) tmem_ld16(taddr + (c + 1) * 16, rb);
                    if (valid) epi_chunk16<BF16>(e, ra, pix, pass * p.np + c * 16, b, y, x);
                    if (warp == T2_EPI0 && lane == 0 && c == 0) T2_STAMP(tn, 11);
                    if (has_b) {
                        tmem_ld_wait();
                        if (warp == T2_EPI0 && lane == 0 && c == 0) T2_STAMP(tn, 12);
                        if (c + 2 < c_end) tmem_ld16(taddr + (c + 2) * 16, ra);
                        if (valid) epi_chunk16<BF16>(e, rb, pix, pass * p.np + (c + 1) * 16, b, y, x);
                        if (warp == T2_EPI0 && lane == 0 && c == 0) T2_STAMP(tn, 13);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tm_empty[acc]);
                if (warp == T2_EPI0 && lane == 0) T2_STAMP(tn, 9);
                ++tn;
                if (++acc == 2) { acc = 0; pacc ^= 1; }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) T2_STAMP_NS(63);
    if (warp == T2_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Shared-memory plan: columns per pass and input-ring depth for (Cin, Cout, operand size); 0 stages = does not fit.
static void tc2_plan(int Cin, int Cout, int esz, int* np, int* nstage) {
    const int kch = Cin * esz / 128;
    *np = 0; *nstage = 0;
    const int cands[2] = {128, 64};
    for (int ci = 0; ci < 2; ++ci) {
        const int cand = cands[ci];
        if (Cout % cand) continue;
        const long w_bytes = 9L * kch * cand * 128;
        const long left = T2_SMEM_MAX - 1024 - w_bytes - 1024 - 4L * Cout;
        int ns = left > 0 ? (int)(left / T2_STAGE) : 0;
        if (ns > T2_MAXSTAGE) ns = T2_MAXSTAGE;
        if (ns >= 3) { *np = cand; *nstage = ns; return; }
    }
}

extern "C" int fcvsr_conv3x3_tc_resident(const void* x, int ldx, const void* w, int ldw, const float* bias,
                                         const float* pre, int ldpre, const float* res, int ldres, float* y, int ldy,
                                         int B, int H, int W, int Cin, int Cout, int act, float slope,
                                         const float* slope_ptr, int pixel_shuffle, float* y2, int ldy2, int round_out,
                                         int max_ctas, int op16, cudaStream_t st) {
    if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0) return FCVSR_ERR_ARG;
    const int esz = op16 ? 2 : 4, kel = 128 / esz;
    if (Cin <= 0 || Cin % kel || Cout <= 0 || Cout % 64) return FCVSR_ERR_UNSUPPORTED;
    int np, nstage;
    tc2_plan(Cin, Cout, esz, &np, &nstage);
    if (!nstage) return FCVSR_ERR_UNSUPPORTED;
    if (((ldx * esz) & 15) || ((ldw * esz) & 15) || (((uintptr_t)x | (uintptr_t)w) & 15)) return FCVSR_ERR_UNSUPPORTED;
    const int esz_y = ((op16 && round_out) || round_out == 2) ? 2 : 4, esz_y2 = op16 ? 2 : 4;
    if (((ldy * esz_y) & 15) || ((uintptr_t)y & 15) || (res && ((ldres & 3) || ((uintptr_t)res & 15))) ||
        (pre && ((ldpre & 7) || ((uintptr_t)pre & 31))) || (y2 && (((ldy2 * esz_y2) & 15) || ((uintptr_t)y2 & 15))))
        return FCVSR_ERR_UNSUPPORTED;
    if (op16 && ((round_out && (ldy & 7)) || (y2 && (ldy2 & 7)))) return FCVSR_ERR_UNSUPPORTED;
    if (act == FCVSR_ACT_PRELU && !slope_ptr) return FCVSR_ERR_ARG;
    if (pixel_shuffle && (y2 || ((Cout >> 2) % 16) || round_out == 2)) return FCVSR_ERR_UNSUPPORTED;
    if (round_out == 2 && (((ldy * 2) & 31) || ((uintptr_t)y & 31))) return FCVSR_ERR_UNSUPPORTED;
    static EncodeTiledFn2 enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return FCVSR_ERR_CUDA;
        enc = (EncodeTiledFn2)ptr;
    }
    CUtensorMap map_w;
    {   // weights [Cout][ldw], K-major (row = output channel, 9*Cin used columns); box = one 128-byte chunk x np rows
        cuuint64_t dims[2] = {(cuuint64_t)(9 * Cin), (cuuint64_t)Cout};
        cuuint64_t strides[1] = {(cuuint64_t)ldw * esz};
        cuuint32_t box[2] = {(cuuint32_t)kel, (cuuint32_t)np};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&map_w, op16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box,
                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCVSR_ERR_CUDA;
    }
    ConvTc2Params p;
    p.x = x; p.ldx = ldx;
    p.e.bias = bias; p.e.bias_sa = 0; p.e.pre = pre; p.e.ldpre = ldpre; p.e.res = res; p.e.ldres = ldres; p.e.res2 = nullptr; p.e.ldres2 = 0;
    p.e.y = y; p.e.ldy = ldy; p.e.y2 = y2; p.e.ldy2 = ldy2; p.e.round_out = round_out;
    p.e.act = act; p.e.slope = slope; p.e.ps = pixel_shuffle; p.e.c4 = Cout >> 2; p.e.H = H; p.e.W = W;
    {
        const uintptr_t a = (uintptr_t)y | (uintptr_t)res | (uintptr_t)y2;
        p.e.wide = !(a & 31) && !((ldy * esz_y) & 31) && (!res || !((ldres * 4) & 31)) && (!y2 || !((ldy2 * esz_y2) & 31)) &&
                   (!pixel_shuffle || !(((Cout >> 2) * esz_y) & 31));
    }
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kch = Cin / kel; p.np = np; p.npass = Cout / np; p.nstage = nstage;
    p.tiles_x = (W + T2_TWV - 1) / T2_TWV; p.tiles_y = (H + T2_TH - 1) / T2_TH;
    p.tiles = p.tiles_x * p.tiles_y * B;
    p.slope_ptr = slope_ptr;
    static int* err = nullptr;
    static int num_sms = 0;
    if (!err) {
        if (cudaMalloc(&err, sizeof(int)) != cudaSuccess) return FCVSR_ERR_CUDA;
        cudaMemset(err, 0, sizeof(int));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(conv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_MAX) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_MAX) != cudaSuccess)
            return FCVSR_ERR_CUDA;
    }
    p.err = err;
    const size_t smem = 1024 + (size_t)9 * p.kch * np * 128 + (size_t)nstage * T2_STAGE + 64 + 256 + 4 * (size_t)Cout;
    int grid = p.tiles < num_sms ? p.tiles : num_sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (op16) conv_tc2_kernel<true><<<grid, T2_THREADS, smem, st>>>(map_w, p);
    else conv_tc2_kernel<false><<<grid, T2_THREADS, smem, st>>>(map_w, p);
    return fcvsr_launch_status();
}

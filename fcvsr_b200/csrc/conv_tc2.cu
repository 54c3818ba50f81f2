// tcgen05 3x3 convolution, second generation: weights RESIDENT in shared memory, single-copy haloed
// input tile.  Same math / epilogue contract as conv_tc.cu (NHWC fp32, TF32 operands, fp32 accumulate in
// TMEM); restricted to the shape class that dominates the FCVSR trunk: k = 3, stride 1, Cin in {32, 64}
// per launch (Cin = 128 is run as two K-halves chained through the `pre` addend), Cout % 64 == 0.
//
// Why (profiles/r1_conv_tc_full_summary.txt): conv_tc.cu re-streams the 147 KB weight tile for every
// 128-pixel tile and loads three shifted copies of the input window, so a 64->64 conv moves 124 MB from
// L2 to shared memory for 14.7 MB of input and the tensor pipe idles ~70 % of the time behind the L2
// fabric (5.2 TB/s).  Here:
//   * the [64 x 9*Cin] weight slab of the current 64-column pass is TMA-loaded ONCE per CTA
//     (SWIZZLE_128B, 147 KB for Cin = 64) and stays resident across all tiles of the pass;
//   * the A operand uses the NO-SWIZZLE K-major canonical layout with SBO = 128 B: row r of a
//     16-byte K-granule plane sits at plane + 16*r, i.e. rows are affine, so filter tap (ky,kx) of the
//     haloed (8+2) x 16 window is the SAME tile viewed from row ky*16 + kx -- one copy serves all nine
//     taps (1.43x input bytes instead of 3.75x + weights).  The M = 128 rows of a tile are the 8 x 16
//     window positions; columns 14, 15 of each row wrap into the next row's halo and are discarded, so a
//     tile yields 8 x 14 output pixels (12.5 % padding work, paid to keep every tap a pure pointer shift);
//   * that layout cannot be written by TMA with 128-byte rows, so four producer warps fill it with
//     16-byte cp.async (zero-fill outside the image = conv padding), completion tracked by
//     cp.async.mbarrier.arrive.noinc on the stage's full barrier (3-stage ring);
//   * one thread issues the 36 MMAs of a chunk with constant-add descriptors, one commit per chunk.
#include "tc_common.cuh"
#include <stdlib.h>

#define T2_TH 8
#define T2_TWP 16                      // window columns per tile row (M = 8 x 16)
#define T2_TWV 14                      // valid output columns per tile
#define T2_ROWS ((T2_TH + 2) * T2_TWP) // 160 haloed positions actually loaded
#define T2_PLANE 2608                  // bytes per 16-byte-granule plane: 163 rows (bank rotation 12 words)
#define T2_STAGE (8 * T2_PLANE)        // one 32-channel chunk of a haloed tile
#define T2_NSTAGE 3
#define T2_WCHUNK (64 * 128)           // weight bytes per (tap, chunk): 64 rows x 128 B
#define T2_THREADS 288                 // 4 producer warps, 1 MMA warp, 4 epilogue warps

struct ConvTc2Params {
    const float* x; int ldx;
    const float* bias; const float* pre; int ldpre; const float* res; int ldres;
    float* y; int ldy; float* y2; int ldy2; int round_out;
    int B, H, W, Cin, Cout, kch, npass;
    int tiles_x, tiles_y, tiles;       // tiles per pass
    int act; float slope; const float* slope_ptr; int ps;
    int* err;
    long long* trace;   // bring-up only: per-tile clock64 stamps of CTA 0 (dbg & 16)
    int dbg;        // bring-up only (FCVSR_TC_DBG): 1 no MMA, 2 no A loads, 8 no epilogue math/stores
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

__global__ void __launch_bounds__(T2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const ConvTc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t w_bytes = (uint32_t)(9 * p.kch) * T2_WCHUNK;
    uint8_t* w_buf = smem;
    uint8_t* a_buf = smem + w_bytes;
    uint64_t* bars = (uint64_t*)(a_buf + T2_NSTAGE * T2_STAGE + 64);
    uint64_t* a_full = bars;                   // [NSTAGE] 128 cp.async arrivals
    uint64_t* a_empty = bars + T2_NSTAGE;      // [NSTAGE] 1 (tcgen05.commit)
    uint64_t* tm_full = bars + 2 * T2_NSTAGE;  // [2]
    uint64_t* tm_empty = tm_full + 2;          // [2] 4 epilogue warps
    uint64_t* w_full = tm_empty + 2;           // weights of the current pass landed
    uint64_t* w_free = w_full + 1;             // all MMAs of the pass retired (weights may be overwritten)
    uint32_t* tmem_slot = (uint32_t*)(w_free + 1);
    float* epi_stage = (float*)(bars + 32);          // 4 warps x 32 x 17 floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < T2_NSTAGE; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], 4); }
        mbar_init(w_full, 1);
        mbar_init(w_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== A producers: 128 threads, thread i owns granule plane (i & 7) of window column (i >> 3) =====
        const int i = threadIdx.x;
        const int col = i >> 3, plane = i & 7;
        int stage = 0; uint32_t phase = 0;
        for (int pass = 0; pass < p.npass; ++pass)
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, r_ = t / p.tiles_x;
                const int ty = r_ % p.tiles_y, b = r_ / p.tiles_y;
                const int xx = tx * T2_TWV - 1 + col;
                const bool xok = xx >= 0 && xx < p.W;
                const int y0 = ty * T2_TH - 1;
                const float* img = p.x + (size_t)b * p.H * p.W * p.ldx;
                for (int kc = 0; kc < p.kch; ++kc) {
                    mbar_wait_warp(&a_empty[stage], phase ^ 1, p.err, 11);
                    if ((p.dbg & 16) && blockIdx.x == 0 && i == 0 && pass == 0) p.trace[(t / gridDim.x) * 16 + kc * 2 + 0] = clock64();
                    const uint32_t dst0 = smem_u32(a_buf + stage * T2_STAGE) + plane * T2_PLANE + col * 16;
                    const float* src0 = img + (size_t)xx * p.ldx + kc * 32 + plane * 4;
#pragma unroll
                    for (int j = 0; j < T2_TH + 2; ++j) {
                        if (p.dbg & 2) break;
                        const int yy = y0 + j;
                        const bool ok = xok && yy >= 0 && yy < p.H;
                        const float* src = ok ? src0 + (size_t)yy * p.W * p.ldx : p.x;
                        cp_async16(dst0 + j * (T2_TWP * 16), src, ok ? 16u : 0u);
                    }
                    cp_async_arrive_noinc(&a_full[stage]);
                    if ((p.dbg & 16) && blockIdx.x == 0 && i == 0 && pass == 0) p.trace[(t / gridDim.x) * 16 + kc * 2 + 1] = clock64();
                    if (++stage == T2_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
    } else if (warp == 4) {
        // ===== weight loader + MMA issuer (one lane) =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t w_desc0 = make_desc(smem_u32(w_buf));
            int sa = 0; uint32_t pa = 0;
            int acc = 0; uint32_t pacc = 0;
            for (int pass = 0; pass < p.npass; ++pass) {
                if (pass > 0) mbar_wait(w_free, (uint32_t)((pass - 1) & 1), p.err, 12);
                mbar_expect_tx(w_full, w_bytes);
                for (int tap = 0; tap < 9; ++tap)
                    for (int kc = 0; kc < p.kch; ++kc)
                        tma_load_2d(w_buf + (kc * 9 + tap) * T2_WCHUNK, &map_w, w_full, tap * p.Cin + kc * 32, pass * 64);
                mbar_wait(w_full, (uint32_t)(pass & 1), p.err, 13);
                tc_fence_after();
                for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                    mbar_wait(&tm_empty[acc], pacc ^ 1, p.err, 14);
                    tc_fence_after();
                    const bool tr = (p.dbg & 16) && blockIdx.x == 0 && pass == 0;
                    if (tr) p.trace[(t / gridDim.x) * 16 + 4] = clock64();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
                    for (int kc = 0; kc < p.kch; ++kc) {
                        mbar_wait(&a_full[sa], pa, p.err, 15);
                        tc_fence_after();
                        if (tr) p.trace[(t / gridDim.x) * 16 + 5 + kc] = clock64();
                        const uint64_t a_desc0 = make_desc_a(smem_u32(a_buf + sa * T2_STAGE));
                        const uint64_t b_desc0 = w_desc0 + (uint64_t)((kc * 9 * T2_WCHUNK) >> 4);
                        if (!(p.dbg & 1))
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint64_t a_d = a_desc0 + (uint64_t)(ky * T2_TWP + kx);          // 16 B per row
                                const uint64_t b_d = b_desc0 + (uint64_t)(((ky * 3 + kx) * T2_WCHUNK) >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_tf32(d_tmem, a_d + (uint64_t)(k * 2 * (T2_PLANE >> 4)), b_d + 2 * k, idesc,
                                              (kc | ky | kx | k) ? 1u : 0u);
                            }
                        umma_commit(&a_empty[sa]);
                        if (++sa == T2_NSTAGE) { sa = 0; pa ^= 1; }
                    }
                    umma_commit(&tm_full[acc]);
                    if (tr) p.trace[(t / gridDim.x) * 16 + 7] = clock64();
                    if (++acc == 2) { acc = 0; pacc ^= 1; }
                }
                umma_commit(w_free);
            }
        }
    } else {
        // ===== epilogue warps 5..8: TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int ly = m >> 4, lx = m & 15;
        const float slope = p.act == FCVSR_ACT_PRELU ? p.slope_ptr[0] : p.slope;
        const int c4 = p.Cout >> 2;
        int acc = 0; uint32_t pacc = 0;
        for (int pass = 0; pass < p.npass; ++pass)
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x, r_ = t / p.tiles_x;
                const int ty = r_ % p.tiles_y, b = r_ / p.tiles_y;
                const int y = ty * T2_TH + ly, x = tx * T2_TWV + lx;
                const bool valid = lx < T2_TWV && y < p.H && x < p.W;
                const size_t pix = ((size_t)b * p.H + y) * p.W + x;
                mbar_wait_warp(&tm_full[acc], pacc, p.err, 16);
                tc_fence_after();
                const bool tr = (p.dbg & 16) && blockIdx.x == 0 && pass == 0 && threadIdx.x == 160;
                if (tr) p.trace[(t / gridDim.x) * 16 + 8] = clock64();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64);
                // One tcgen05.ld for the whole 64-column accumulator row, residual loads issued before the
                // TMEM wait, then 16 independent 16-byte stores per tensor: the epilogue of a tile is a handful of
                // long-latency operations in flight at once instead of four serialized load->store rounds.
                uint32_t r[64];
                tmem_ld64(taddr, r);
                const int n0 = pass * 64;
                float4 rs[16], pr[16];
                const bool has_res = valid && p.res, has_pre = valid && p.pre;
                if (has_res) {
                    const float4* rp = reinterpret_cast<const float4*>(p.res + pix * p.ldres + n0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) rs[j] = rp[j];
                }
                if (has_pre) {
                    const float4* rp = reinterpret_cast<const float4*>(p.pre + pix * p.ldpre + n0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) pr[j] = rp[j];
                }
                tmem_ld_wait();
                if (tr) p.trace[(t / gridDim.x) * 16 + 10] = clock64();
                // the accumulator is in registers: release the TMEM buffer to the MMA warp right away
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tm_empty[acc]);
                if (valid && !(p.dbg & 8)) {
                    float v[64];
#pragma unroll
                    for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) v[j] += __ldg(p.bias + n0 + j);
                    }
                    if (has_pre) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { v[4 * j] += pr[j].x; v[4 * j + 1] += pr[j].y; v[4 * j + 2] += pr[j].z; v[4 * j + 3] += pr[j].w; }
                    }
#pragma unroll
                    for (int j = 0; j < 64; ++j) v[j] = fcvsr_act(v[j], p.act, slope);
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { v[4 * j] += rs[j].x; v[4 * j + 1] += rs[j].y; v[4 * j + 2] += rs[j].z; v[4 * j + 3] += rs[j].w; }
                    }
                    if (p.y2) {
                        float4* d2 = reinterpret_cast<float4*>(p.y2 + pix * p.ldy2 + n0);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            d2[j] = make_float4(round_tf32(v[4 * j]), round_tf32(v[4 * j + 1]), round_tf32(v[4 * j + 2]),
                                                round_tf32(v[4 * j + 3]));
                    }
                    if (p.round_out) {
#pragma unroll
                        for (int j = 0; j < 64; ++j) v[j] = round_tf32(v[j]);
                    }
                    if (p.ps) {
#pragma unroll
                        for (int cb = 0; cb < 64; cb += 16) {
                            const int ij = (n0 + cb) / c4, c = (n0 + cb) - ij * c4;
                            const size_t opix = ((size_t)b * 2 * p.H + 2 * y + (ij >> 1)) * (2 * (size_t)p.W) + 2 * x + (ij & 1);
                            float4* dp = reinterpret_cast<float4*>(p.y + opix * p.ldy + c);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                dp[j] = make_float4(v[cb + 4 * j], v[cb + 4 * j + 1], v[cb + 4 * j + 2], v[cb + 4 * j + 3]);
                        }
                    } else {
                        float4* dp = reinterpret_cast<float4*>(p.y + pix * p.ldy + n0);
#pragma unroll
                        for (int j = 0; j < 16; ++j) dp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
                if (tr) p.trace[(t / gridDim.x) * 16 + 9] = clock64();
                if (++acc == 2) { acc = 0; pacc ^= 1; }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
static long long* g_trace = nullptr;
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int fcvsr_conv3x3_tc_resident(const float* x, int ldx, const float* w, int ldw, const float* bias,
                                         const float* pre, int ldpre, const float* res, int ldres, float* y, int ldy,
                                         int B, int H, int W, int Cin, int Cout, int act, float slope,
                                         const float* slope_ptr, int pixel_shuffle, float* y2, int ldy2, int round_out,
                                         int max_ctas, cudaStream_t st) {
    if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0) return FCVSR_ERR_ARG;
    if ((Cin != 32 && Cin != 64) || Cout <= 0 || Cout % 64) return FCVSR_ERR_UNSUPPORTED;
    if ((ldx & 3) || (ldy & 3) || (ldw & 3) || (res && (ldres & 3)) || (pre && (ldpre & 3)) || (y2 && (ldy2 & 3)))
        return FCVSR_ERR_UNSUPPORTED;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)res | (uintptr_t)pre | (uintptr_t)y2) & 15)
        return FCVSR_ERR_UNSUPPORTED;
    if (act == FCVSR_ACT_PRELU && !slope_ptr) return FCVSR_ERR_ARG;
    if (pixel_shuffle && (y2 || ((Cout >> 2) % 16))) return FCVSR_ERR_UNSUPPORTED;
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
    {   // weights [Cout][ldw] fp32, K-major (row = output channel, 9*Cin used columns); box = 32 k x 64 rows
        cuuint64_t dims[2] = {(cuuint64_t)(9 * Cin), (cuuint64_t)Cout};
        cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
        cuuint32_t box[2] = {32, 64};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCVSR_ERR_CUDA;
    }
    ConvTc2Params p;
    p.x = x; p.ldx = ldx; p.bias = bias; p.pre = pre; p.ldpre = ldpre; p.res = res; p.ldres = ldres;
    p.y = y; p.ldy = ldy; p.y2 = y2; p.ldy2 = ldy2; p.round_out = round_out;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kch = Cin / 32; p.npass = Cout / 64;
    p.tiles_x = (W + T2_TWV - 1) / T2_TWV; p.tiles_y = (H + T2_TH - 1) / T2_TH;
    p.tiles = p.tiles_x * p.tiles_y * B;
    p.act = act; p.slope = slope; p.slope_ptr = slope_ptr; p.ps = pixel_shuffle;
    static int* err = nullptr;
    static int num_sms = 0;
    const size_t smem = 1024 + (size_t)9 * 2 * T2_WCHUNK + T2_NSTAGE * T2_STAGE + 64 + 256 + 4 * 32 * EPI_PITCH * 4;
    if (!err) {
        if (cudaMalloc(&err, sizeof(int)) != cudaSuccess) return FCVSR_ERR_CUDA;
        cudaMemset(err, 0, sizeof(int));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return FCVSR_ERR_CUDA;
    }
    p.err = err;
    if (!g_trace) { cudaMalloc(&g_trace, 64 * 16 * sizeof(long long)); cudaMemset(g_trace, 0, 64 * 16 * sizeof(long long)); }
    p.trace = g_trace;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FCVSR_TC_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
    int grid = p.tiles < num_sms ? p.tiles : num_sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    conv_tc2_kernel<<<grid, T2_THREADS, smem, st>>>(map_w, p);
    return fcvsr_launch_status();
}

// bring-up only: copy the clock64 trace of CTA 0 (FCVSR_TC_DBG & 16) to the host
extern "C" int fcvsr_debug_conv_trace(long long* host_out, int n) {
    if (!g_trace || n > 64 * 16) return FCVSR_ERR_ARG;
    return cudaMemcpy(host_out, g_trace, n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? FCVSR_OK : FCVSR_ERR_CUDA;
}

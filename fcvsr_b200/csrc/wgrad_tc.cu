// Weight gradient of a 3x3 / 1x1 stride-1 convolution on the tcgen05 tensor cores (sm_100a only).
//
//   dW[tap][ci][co] += sum over output pixels p of  X[p + tap][ci] * dY[p][co]            (bf16 operands, fp32 accumulation)
//
// is the training-step counterpart of conv_tc.cu (what cuDNN's wgrad computes for the reference's nn.Conv2d layers under
// loss.backward(), CVSR_train/train_LD_freqCVSR_22.py:250).  The contraction runs over PIXELS, so with NHWC tensors both
// operands arrive "MN-major": a TMA box [pixels][64 channels] is 128-byte rows of channels, i.e. rows index K and the
// contiguous direction indexes M (input channels) or N (output channels).  tcgen05.mma takes that layout directly (a_major =
// b_major = MN in the instruction descriptor; canonical SWIZZLE_128B MN-major layout ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in
// 16-byte units: 8-row K groups SBO = 1024 B apart, 64-element MN chunks LBO apart), so no transposed copy of the activations
// is ever made.
//
// Per CTA: one 64-channel chunk of Cin, one 64-channel chunk of Cout, a share of the 8x16-pixel tiles (split-K over pixels).
//   * A stage: the (8+2) x 16 haloed input window as THREE x-shifted copies (exactly conv_tc's halo re-use: tap (ky,kx) is copy
//     kx viewed from row ky*16), TMA zero fill = convolution padding.  B stage: the 8x16 dY tile.
//   * UMMA M must be 128 but a tap offers only 64 input channels, so two taps are stacked along M through LBO: taps (0,kx) and
//     (1,kx) are 16 rows = 2048 B apart inside copy kx, taps (2,0) and (2,1) one copy apart; tap (2,2) is issued with LBO = 0
//     (its upper half duplicates the lower and is discarded).  Five accumulators [128 x 64] in TMEM hold all nine taps; they
//     accumulate over ALL tiles of the CTA and are drained once, with vector reductions (red.global.add.v4.f32) into dW.
//   * 1x1 convolutions are the single-tap case (one accumulator, LBO = 0).
// Roofline: tensor.  Algorithmic FLOPs = 2 * Cin * Cout * k * k per output pixel (as the forward convolution).
#include "tc_common.cuh"

#define WG_TH 8
#define WG_TW 16
#define WG_ROW 128                                        // bytes per operand row (64 bf16 channels)
#define WG_A_COPY ((WG_TH + 2) * WG_TW * WG_ROW)          // 20480
#define WG_A_STAGE (3 * WG_A_COPY)                        // 61440
#define WG_B_STAGE (WG_TH * WG_TW * WG_ROW)               // 16384
#define WG_STAGE (WG_A_STAGE + WG_B_STAGE)                // 77824
#define WG_NSTAGE 2
#define WG_THREADS (32 * 6)                               // TMA producer, MMA issuer, 4 epilogue warps
#define WG_MAXG 5

#define WG_MAX_PROB 3
// Up to three problems (tensor pairs of different spatial size, same channels and batch: the pyramid levels of a BlockRCB
// convolution, whose weight gradients add up) share one launch: the CTAs walk ONE tile list and accumulate into the same dW.
struct WgradTcParams {
    float* dw;                   // [k*k][Cin][Cout] fp32, accumulated
    int B, Cin, Cout, ks;
    int nprob, total_tiles;
    int tiles_x[WG_MAX_PROB], tiles_y[WG_MAX_PROB], tile_begin[WG_MAX_PROB];
    int* err;
};

// MN-major SWIZZLE_128B shared-memory descriptor: 8-row K groups 1024 B apart (SBO), 64-element MN chunks `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_dy1,
                const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_dy2, const WgradTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + WG_NSTAGE * WG_STAGE);
    uint64_t* full = bars;                   // [NSTAGE]
    uint64_t* empty = bars + WG_NSTAGE;      // [NSTAGE]
    uint64_t* done = bars + 2 * WG_NSTAGE;   // [1] all MMAs of the CTA retired
    uint32_t* tmem_slot = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KS = p.ks;
    const int ngroups = KS == 3 ? 5 : 1;
    const int ci0 = blockIdx.z * 64, co0 = blockIdx.y * 64;
    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ntiles_mine = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x, +gridDim.x, ...

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int pr = (p.nprob > 1 && t >= p.tile_begin[1]) ? ((p.nprob > 2 && t >= p.tile_begin[2]) ? 2 : 1) : 0;
                const int txs = pr == 0 ? p.tiles_x[0] : (pr == 1 ? p.tiles_x[1] : p.tiles_x[2]);
                const int tys = pr == 0 ? p.tiles_y[0] : (pr == 1 ? p.tiles_y[1] : p.tiles_y[2]);
                const int tl = t - (pr == 0 ? 0 : (pr == 1 ? p.tile_begin[1] : p.tile_begin[2]));
                const CUtensorMap* mx = pr == 0 ? &map_x : (pr == 1 ? &map_x1 : &map_x2);
                const CUtensorMap* mdy = pr == 0 ? &map_dy : (pr == 1 ? &map_dy1 : &map_dy2);
                const int tx = tl % txs, r = tl / txs;
                const int ty = r % tys, b = r / tys;
                const int y0 = ty * WG_TH, x0 = tx * WG_TW;
                mbar_wait(&empty[stage], phase ^ 1, p.err, 31);
                uint8_t* sa = smem + stage * WG_STAGE;
                if (KS == 3) {
                    mbar_expect_tx(&full[stage], WG_STAGE);
                    for (int c = 0; c < 3; ++c) tma_load_4d(sa + c * WG_A_COPY, mx, &full[stage], ci0, x0 - 1 + c, y0 - 1, b);
                } else {
                    mbar_expect_tx(&full[stage], 2 * WG_B_STAGE);
                    tma_load_4d(sa, mx, &full[stage], ci0, x0, y0, b);
                }
                tma_load_4d(sa + WG_A_STAGE, mdy, &full[stage], co0, x0, y0, b);
                if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // kind::f16, bf16 x bf16 -> fp32, A and B MN-major (bits 15 / 16), N = 64, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                mbar_wait(&full[stage], phase, p.err, 32);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * WG_STAGE);
                const uint32_t sb = sa + WG_A_STAGE;
                for (int g = 0; g < ngroups; ++g) {
                    // group -> (start offset of the lower tap, LBO to the upper tap)
                    uint32_t a_off, lbo;
                    if (KS == 1) { a_off = 0; lbo = 0; }
                    else if (g < 3) { a_off = g * WG_A_COPY; lbo = WG_TW * WG_ROW; }                  // (0,kx) + (1,kx)
                    else if (g == 3) { a_off = 2 * WG_TW * WG_ROW; lbo = WG_A_COPY; }                 // (2,0) + (2,1)
                    else { a_off = 2 * WG_A_COPY + 2 * WG_TW * WG_ROW; lbo = 0; }                     // (2,2) twice
                    const uint64_t a_d0 = make_desc_mn(sa + a_off, lbo);
                    const uint64_t b_d0 = make_desc_mn(sb, 0);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(g * 64);
#pragma unroll
                    for (int k = 0; k < 8; ++k)                    // 16 pixel rows (2048 B) per MMA
                        umma_f16(d_tmem, a_d0 + (uint64_t)(k * 128), b_d0 + (uint64_t)(k * 128), idesc, (it | k) ? 1u : 0u);
                }
                umma_commit(&empty[stage]);
                if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
            }
            umma_commit(done);
        }
    } else {
        // epilogue: after every MMA of the CTA has retired, lane m of the accumulators = (upper / lower tap, input channel)
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        if (ntiles_mine > 0) {
            mbar_wait_warp(done, 0, p.err, 33);
            tc_fence_after();
            const int m = q * 32 + lane;
            const int ci = ci0 + (m & 63);
            const int upper = m >> 6;
            for (int g = 0; g < ngroups; ++g) {
                int tap;
                if (KS == 1) tap = upper ? -1 : 0;
                else if (g < 3) tap = upper ? 3 + g : g;                   // (1,kx) : (0,kx)
                else if (g == 3) tap = upper ? 7 : 6;                      // (2,1) : (2,0)
                else tap = upper ? -1 : 8;                                 // (2,2), duplicate discarded
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c * 16, r);
                    tmem_ld_wait();
                    if (tap >= 0 && ci < p.Cin) {
                        float* dst = p.dw + ((size_t)tap * p.Cin + ci) * p.Cout + co0 + c * 16;
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                                         : "memory");
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*WgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// x[i]: bf16 NHWC [B,H_i,W_i,ldx] (Cin channels), dy[i]: bf16 NHWC [B,H_i,W_i,lddy] (Cout channels), 16-byte aligned, ld % 8 == 0
// (HOST arrays of nprob <= 3 device pointers; H, W HOST arrays); dw: fp32 [k*k][Cin][Cout], ACCUMULATED over every problem
// (zero-fill first; Cout % 4 == 0 and 16-byte aligned for the vector reductions).
// k in {1, 3}, stride 1, padding k/2, Cin % 64 == 0, Cout % 64 == 0; other shapes return FCVSR_ERR_UNSUPPORTED.
extern "C" int fcvsr_conv2d_wgrad_tc_multi(int nprob, const void* const* x, int ldx, const void* const* dy, int lddy, float* dw, int B,
                                           const int* H, const int* W, int Cin, int Cout, int ksize, cudaStream_t st) {
    if (nprob < 1 || nprob > WG_MAX_PROB || !x || !dy || !dw || !H || !W || B <= 0) return FCVSR_ERR_ARG;
    if ((ksize != 1 && ksize != 3) || Cin <= 0 || Cout <= 0 || (Cin & 63) || (Cout & 63)) return FCVSR_ERR_UNSUPPORTED;
    if ((ldx & 7) || (lddy & 7) || ((uintptr_t)dw & 15)) return FCVSR_ERR_UNSUPPORTED;
    static WgEncodeFn enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return FCVSR_ERR_CUDA;
        enc = (WgEncodeFn)ptr;
    }
    CUtensorMap map_x[WG_MAX_PROB], map_dy[WG_MAX_PROB];
    WgradTcParams p;
    p.dw = dw; p.B = B; p.Cin = Cin; p.Cout = Cout; p.ks = ksize; p.nprob = nprob;
    int tiles = 0;
    for (int i = 0; i < WG_MAX_PROB; ++i) {
        const int j = i < nprob ? i : 0;
        if (i >= nprob) { map_x[i] = map_x[0]; map_dy[i] = map_dy[0]; p.tiles_x[i] = p.tiles_x[0]; p.tiles_y[i] = p.tiles_y[0]; p.tile_begin[i] = tiles; continue; }
        if (!x[j] || !dy[j] || H[j] <= 0 || W[j] <= 0) return FCVSR_ERR_ARG;
        if (((uintptr_t)x[j] | (uintptr_t)dy[j]) & 15) return FCVSR_ERR_UNSUPPORTED;
        {
            cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W[j], (cuuint64_t)H[j], (cuuint64_t)B};
            cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)W[j] * ldx * 2, (cuuint64_t)H[j] * W[j] * ldx * 2};
            cuuint32_t box[4] = {64, WG_TW, (cuuint32_t)(ksize == 3 ? WG_TH + 2 : WG_TH), 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (enc(&map_x[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)x[j], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return FCVSR_ERR_CUDA;
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W[j], (cuuint64_t)H[j], (cuuint64_t)B};
            cuuint64_t strides[3] = {(cuuint64_t)lddy * 2, (cuuint64_t)W[j] * lddy * 2, (cuuint64_t)H[j] * W[j] * lddy * 2};
            cuuint32_t box[4] = {64, WG_TW, WG_TH, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (enc(&map_dy[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)dy[j], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return FCVSR_ERR_CUDA;
        }
        p.tiles_x[i] = (W[j] + WG_TW - 1) / WG_TW; p.tiles_y[i] = (H[j] + WG_TH - 1) / WG_TH;
        p.tile_begin[i] = tiles;
        tiles += p.tiles_x[i] * p.tiles_y[i] * B;
    }
    p.total_tiles = tiles;
    static int* err = nullptr;
    static int num_sms = 0;
    const size_t smem = 1024 + (size_t)WG_NSTAGE * WG_STAGE + 256;
    if (!err) {
        if (cudaMalloc(&err, sizeof(int)) != cudaSuccess) return FCVSR_ERR_CUDA;
        cudaMemset(err, 0, sizeof(int));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return FCVSR_ERR_CUDA;
    }
    p.err = err;
    const int slabs = (Cin / 64) * (Cout / 64);
    int gx = num_sms / slabs;                              // split-K over pixel tiles: fill the SMs once
    if (gx < 1) gx = 1;
    if (gx > p.total_tiles) gx = p.total_tiles;
    dim3 grid(gx, Cout / 64, Cin / 64);
    wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(map_x[0], map_dy[0], map_x[1], map_dy[1], map_x[2], map_dy[2], p);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_conv2d_wgrad_tc(const void* x, int ldx, const void* dy, int lddy, float* dw, int B, int H, int W, int Cin,
                                     int Cout, int ksize, cudaStream_t st) {
    if (!x || !dy) return FCVSR_ERR_ARG;
    const void* xs[1] = {x};
    const void* dys[1] = {dy};
    return fcvsr_conv2d_wgrad_tc_multi(1, xs, ldx, dys, lddy, dw, B, &H, &W, Cin, Cout, ksize, st);
}

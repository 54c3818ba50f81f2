// tcgen05 / TMEM implicit-GEMM convolution fed by TMA  (sm_100a only).
//
// Computes, for NHWC tensors, y = act(conv_kxk(x, w) + bias) + res - res2 (k in {1,3}, stride 1, zero padding k/2) as a GEMM
// D[pixels, Cout] = sum_{tap, cin} A_tap[pixels, cin] * W[Cout, tap, cin] on the 5th-generation tensor cores: bf16 operand tensors
// (kind::f16) or TF32-rounded fp32 tensors (kind::tf32), fp32 accumulation in tensor memory.  It replaces every 3x3 / 1x1 nn.Conv2d
// of the reference hot path whose shape fits (CVSR_freq.py:1371-1430 MGAAbk, :705-822 SCNetbk, :2739-2749 tail); the library-call
// equivalent in the reference is cuDNN (conv2d) -- there is no such kernel in the reference to port.
//
// Design (one persistent CTA per SM, 19 warps, warp-specialised: A producer, B producer, MMA issuer, 16 epilogue warps):
//   * M tile = 8 x 16 output pixels (128 = UMMA M), N tile = Cout (<= 128 per pass), K walked as (128-byte channel chunk: 64 bf16
//     or 32 TF32 channels) x (filter tap); one tcgen05.mma covers 32 bytes of K.
//   * A operand, halo re-use: for each channel chunk the producer TMA-loads THREE copies of the (8+2) x 16-pixel input window,
//     shifted by -1/0/+1 pixel in x (TMA zero-fills out-of-image pixels = the conv padding).  Each copy is a [160 rows][128 B]
//     SWIZZLE_128B K-major tile, so the operand of filter tap (ky,kx) is copy[kx] viewed from row ky*16 on: a 2048-byte
//     (1024-aligned) shift of the matrix descriptor.  Nine taps read 3 loads instead of 9.
//   * B operand (weights, [Cout][tap*Cin] K-major): one filter row of taps per ring stage; when every stage of the filter fits
//     the ring (bf16 64 -> 64 3x3: 72 KB) the weights are loaded once per CTA and stay resident.  Several filters can be stacked
//     in one weight matrix with a row offset per problem (fcvsr_conv2d_tc_multi_w).
//   * Accumulators: FOUR N-column tiles in TMEM; the epilogue warps form 1, 2 or 4 sets that drain different accumulator tiles
//     concurrently (tcgen05.ld, lane == pixel), apply bias (staged in shared memory) / activation / residuals in registers and
//     store 256-bit runs per thread (optionally through pixel_shuffle(2), optionally a second operand-typed output); bf16
//     64- and 128-channel outputs of resident-filter convolutions are staged in SWIZZLE_128B shared tiles per set (one per 64
//     channels) and leave with one TMA store per 64 channels and tile.  The activation is one uniform switch per 16-column
//     chunk: the epilogue warps share issue slots and the shared-memory pipe with the TMA / MMA path, so its instruction count
//     shows in the tile period.
//   * Up to four problems (tensors of different spatial size: the pyramid levels of SCNetbk) share one persistent tile list;
//     programmatic dependent launch overlaps the prologue with the previous convolution's drain.
//
// Roofline: tensor for N >= 128; the N = 64 shapes are bound by shared-memory bandwidth (an MMA reads 6 KB of operands for 32 clk of
// tensor time; 312 KB per tile = the measured 2440 clk tile period, profiles/r2_notes.md 5).  Algorithmic FLOPs per output pixel
// = 2 * Cin * Cout * k * k; algorithmic HBM bytes per pixel = esz * (Cin + Cout) (+4*Cout per residual).
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

#define TC_TH 8
#define TC_TW 16
#define TC_KCH 32                      // channels per K chunk in TF32 mode (128 bytes of fp32); 64 in bf16 mode
#define TC_NA 2                        // A ring stages of a 3x3 conv (61 KB each); a 1x1 conv's stage is 16 KB: TC_NA1 stages
#define TC_NA1 7                       // ... in the same TC_NA * TC_A_STAGE_BYTES: its tiles are 4-16 MMAs, so load latency is the limit
#define TC_NA_MAX 8
#define TC_NACC 4                      // TMEM accumulator tiles (n_tile <= 128 columns each)
#define TC_NB_MAX 16                   // B ring: as many stages as fit in TC_B_RING_BYTES, at most 16
#define TC_B_RING_BYTES (96 * 1024)
#define TC_ROW_BYTES 128
#define TC_A_COPY_BYTES ((TC_TH + 2) * TC_TW * TC_ROW_BYTES)      // 20480
#define TC_A_STAGE_BYTES (3 * TC_A_COPY_BYTES)                    // 61440
#define TC_STG_BYTES (TC_TH * TC_TW * TC_ROW_BYTES)               // 16384: one staged output tile
#define TC_SMEM_LIMIT 232448                                      // 227 KB of dynamic shared memory per CTA

#define TC_EPI_WARPS 16                 // epilogue warps: TC_EPI_WARPS / 4 per TMEM lane quarter, each a share of the columns
#define TC_THREADS (32 * (3 + TC_EPI_WARPS))
#define TC_MAX_COUT 2048               // bias staging in shared memory

#ifdef TC_TRACE      // bring-up only (tools/gpu_conv_trace.py builds its own copy of this file with -DTC_TRACE)
__device__ long long tc_trace[64 * 16];
#define TC_STAMP(n, slot) do { if (blockIdx.x == 0 && (n) < 62) tc_trace[(n) * 16 + (slot)] = clock64(); } while (0)
#define TC_STAMP_NS(n) do { if (blockIdx.x == 0) { unsigned long long ns_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_)); \
        tc_trace[(n) * 16 + 15] = (long long)ns_; tc_trace[(n) * 16 + 14] = clock64(); } } while (0)
extern "C" int fcvsr_debug_conv_trace(long long* host, int n) {
    return cudaMemcpyFromSymbol(host, tc_trace, sizeof(long long) * (n < 1024 ? n : 1024)) == cudaSuccess ? 0 : 1;
}
#else
#define TC_STAMP(n, slot) do {} while (0)
#define TC_STAMP_NS(n) do {} while (0)
#endif

#define TC_MAX_PROB 4
// One launch can run the same convolution (same weights, bias, activation, strides) on up to three tensors of different
// spatial size: the three pyramid levels of SCNetbk (BlockRCB applies one body to every level, CVSR_freq.py:766-770).  The
// persistent CTAs walk ONE tile list that spans the levels, so the small levels (1/4 and 1/16 of the pixels) fill the SMs
// next to level 0 instead of running as under-filled launches on side streams.
struct ConvTcProblem {
    const float* res; const float* res2; float* y; float* y2;
    int H, W, tiles_x, tiles_y, tile_begin;    // tile_begin: index of the problem's first tile in the launch's tile list
    int wrow;                                  // first row of the problem's filter in the (stacked) weight matrix / bias vector
};
struct ConvTcParams {
    const float* bias; int ldres; int ldres2;
    int ldy;
    int ldy2; int round_out;             // y2: optional TF32-rounded copy of y (non-shuffled outputs only)
    int B, Cin, Cout, ks;                // Cout = padded (multiple of 16) GEMM N
    int cout_valid;                      // channels actually stored (== Cout, or < 16 for thin heads)
    int n_tile, n_tiles;               // N per pass, number of passes
    int nprob, total_tiles;
    int w_rows, multi_w;                 // rows of the weight matrix (>= Cout: several filters stacked); problems use different filters
    ConvTcProblem prob[TC_MAX_PROB];
    int act; float slope; const float* slope_ptr; int ps;
    int wide;                            // 32-byte aligned tensors: use 256-bit loads / stores in the epilogue
    int epi_sets;                        // 1, 2 or 4 epilogue warp sets (see the epilogue)
    int b_ring_bytes;                    // weight ring size (TC_B_RING_BYTES, or less when the output staging tiles take the space)
    int stage_out;                       // output tile staged in shared memory and written with one TMA store per tile
    int* err;
    int dbg;                             // bring-up only (FCVSR_TC_DBG): 1 no MMA, 2 no A loads, 4 no B loads, 8 no stores
};

struct TileCoord { int nt, tx, ty, b, pr; };
__device__ __forceinline__ TileCoord decode_tile(int t, const ConvTcParams& p) {
    TileCoord c;
    c.pr = (p.nprob > 1 && t >= p.prob[1].tile_begin)
               ? ((p.nprob > 2 && t >= p.prob[2].tile_begin) ? ((p.nprob > 3 && t >= p.prob[3].tile_begin) ? 3 : 2) : 1) : 0;
    const ConvTcProblem& q = p.prob[c.pr];
    t -= q.tile_begin;
    c.nt = t % p.n_tiles; t /= p.n_tiles;
    c.tx = t % q.tiles_x; t /= q.tiles_x;
    c.ty = t % q.tiles_y; c.b = t / q.tiles_y;
    return c;
}

template <int KS, bool BF16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x1,
               const __grid_constant__ CUtensorMap map_x2, const __grid_constant__ CUtensorMap map_x3,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y,
               const __grid_constant__ CUtensorMap map_y1, const __grid_constant__ CUtensorMap map_y2,
               const __grid_constant__ CUtensorMap map_y3, const ConvTcParams p) {
    // Programmatic dependent launch: let the next convolution's CTAs take each SM as soon as this grid's CTA leaves it
    // and run their prologue (barriers, TMEM, bias, first weight stages) while the rest of this grid drains; everything
    // that touches activations sits behind griddepcontrol.wait (a no-op for a normally serialized launch).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;                                        // TC_NA x 61440
    uint8_t* b_buf = smem + TC_NA * TC_A_STAGE_BYTES;             // weight ring, TC_B_RING_BYTES
    // output staging: one [128 pixels][128 B] SWIZZLE_128B tile per epilogue set (bf16 outputs with 64 channels)
    uint8_t* stg_buf = b_buf + p.b_ring_bytes;
    uint64_t* bars = (uint64_t*)(stg_buf + (p.stage_out ? p.epi_sets * (p.n_tile >> 6) * TC_STG_BYTES : 0));
    uint64_t* full_a = bars;                // [TC_NA_MAX]
    uint64_t* empty_a = bars + TC_NA_MAX;   // [TC_NA_MAX]
    uint64_t* full_b = bars + 2 * TC_NA_MAX;            // [TC_NB_MAX]
    uint64_t* empty_b = bars + 2 * TC_NA_MAX + TC_NB_MAX;   // [TC_NB_MAX]
    uint64_t* tm_full = bars + 2 * TC_NA_MAX + 2 * TC_NB_MAX;   // [TC_NACC]
    constexpr int NA = KS == 3 ? TC_NA : TC_NA1;                          // A ring depth
    constexpr uint32_t A_STAGE = KS == 3 ? TC_A_STAGE_BYTES : TC_TH * TC_TW * TC_ROW_BYTES;
    uint64_t* tm_empty = tm_full + TC_NACC;               // [TC_NACC]
    uint32_t* tmem_slot = (uint32_t*)(tm_empty + TC_NACC);
    // bias staged in shared memory: the epilogue reads it with broadcast ld.shared.v4 (4 per 16-column chunk).  Per-lane
    // global loads here cost more than the tile's MMAs: the tensor core's operand fetches saturate the L1/shared pipe
    // and every LDG queues behind them (measured: 64->128 bf16 conv 60 us without bias, 147 us with __ldg bias).
    float* bias_s = (float*)((uint8_t*)bars + 512);
    for (int i = threadIdx.x; i < p.w_rows; i += TC_THREADS)
        bias_s[i] = (p.bias && (p.multi_w || i < p.cout_valid)) ? p.bias[i] : 0.f;
    const uint32_t bias_sa = smem_u32(bias_s);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncopies = KS == 3 ? 3 : 1;
    const int nrows = KS == 3 ? TC_TH + 2 : TC_TH;
    constexpr int KCH = BF16 ? 64 : 32;                 // channels per 128-byte operand row
    const int kchunks = p.Cin / KCH;
    const uint32_t a_copy_bytes = (uint32_t)nrows * TC_TW * TC_ROW_BYTES;
    const uint32_t b_bytes = (uint32_t)p.n_tile * TC_ROW_BYTES;      // multiple of 2048 (n_tile % 16 == 0): stays 1024-aligned
    const uint32_t b_stage_bytes = b_bytes * KS;                    // one filter row of taps per stage
    const int nb_stages = min(TC_NB_MAX, (int)((uint32_t)p.b_ring_bytes / b_stage_bytes));
    // Weights resident: with a single N pass and every (K chunk, filter row) stage fitting in the ring at once, the
    // stages are loaded once per CTA and never released (bf16 64->64 3x3: 72 KB).  ncu showed the L2->SM read path
    // at 97 % of peak with the weights re-streamed per tile; this halves that traffic for the most common shape.
    const bool b_resident = p.n_tiles == 1 && kchunks * KS <= nb_stages && !p.multi_w;
    // TC_NACC accumulator tiles in TMEM (all 512 columns at n_tile = 128): with two, the MMAs of tile t wait for the
    // epilogue of tile t-2, a chain of MMA drain + barrier wake-ups + store latency that idled the tensor pipe ~30 %
    const int nacc = TC_NACC;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(nacc * p.n_tile)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_NA_MAX; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < TC_NB_MAX; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        for (int i = 0; i < TC_NACC; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], TC_EPI_WARPS / p.epi_sets); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_STAMP_NS(62);

    if (warp == 0) {
        // ===== A producer: haloed input windows, 3 x-shifted copies per 32-channel chunk =====
        if (elect_one()) {
            asm volatile("griddepcontrol.wait;" ::: "memory");      // x is the previous kernel's output
            int stage = 0; uint32_t phase = 0;
            int tn = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tn) {
                const TileCoord tc = decode_tile(t, p);
                const int y0 = tc.ty * TC_TH - (KS == 3 ? 1 : 0), x0 = tc.tx * TC_TW - (KS == 3 ? 1 : 0);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&empty_a[stage], phase ^ 1, p.err, 1);
                    if (kc == 0) TC_STAMP(tn, 0);
                    if (p.dbg & 2) { mbar_arrive(&full_a[stage]); if (++stage == NA) { stage = 0; phase ^= 1; } continue; }
                    mbar_expect_tx(&full_a[stage], a_copy_bytes * ncopies);
                    uint8_t* dst = a_buf + stage * A_STAGE;
                    const CUtensorMap* mx = tc.pr == 0 ? &map_x : (tc.pr == 1 ? &map_x1 : (tc.pr == 2 ? &map_x2 : &map_x3));
                    for (int cpy = 0; cpy < ncopies; ++cpy)
                        tma_load_4d(dst + cpy * TC_A_COPY_BYTES, mx, &full_a[stage], kc * KCH, x0 + cpy, y0, tc.b);
                    if (kc == kchunks - 1) TC_STAMP(tn, 1);
                    if (++stage == NA) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== B producer: one stage = the KS taps of one filter row for one 32-channel chunk =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, p);
                if (b_resident && t != (int)blockIdx.x) break;          // loaded with the first tile, kept for all
                for (int kc = 0; kc < kchunks; ++kc)
                    for (int ky = 0; ky < KS; ++ky) {
                        if (!b_resident) mbar_wait(&empty_b[stage], phase ^ 1, p.err, 2);
                        if (p.dbg & 4) { mbar_arrive(&full_b[stage]); if (++stage == nb_stages) { stage = 0; phase ^= 1; } continue; }
                        mbar_expect_tx(&full_b[stage], b_bytes * KS);
#pragma unroll
                        for (int kx = 0; kx < KS; ++kx)
                            tma_load_2d(b_buf + stage * b_stage_bytes + kx * b_bytes, &map_w, &full_b[stage],
                                        (ky * KS + kx) * p.Cin + kc * KCH, tc.nt * p.n_tile + p.prob[tc.pr].wrow);
                        if (++stage == nb_stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 2) {
        // ===== MMA issuer (one elected lane) =====
        // Descriptors are formed once per stage and advanced by constant adds: the uniform-datapath chain
        // of a full make_desc() per instruction costs ~130 clk/MMA, 3x the tensor pipe's 48 clk (N=64).
        if (elect_one()) {
            constexpr uint32_t FMT = BF16 ? 1u : 2u;            // F16F32Format: 1 = BF16, 2 = TF32
            const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t b_step = b_bytes >> 4;
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int acc = 0; uint32_t pacc = 0;
            uint32_t pre_t = 0, pre_a = 0, pre_b = 0;          // early polls of the upcoming step's barriers (1 = already complete)
            int tn = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tn) {
                if (!pre_t) mbar_wait(&tm_empty[acc], pacc ^ 1, p.err, 3);
                pre_t = 0;
                tc_fence_after();
                TC_STAMP(tn, 4);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
                const bool b_ready = b_resident && t != (int)blockIdx.x;   // resident stages were waited for on the first tile
                if (b_resident) { sb = 0; pb = 0; }
                for (int kc = 0; kc < kchunks; ++kc) {
                    if (!pre_a) mbar_wait(&full_a[sa], pa, p.err, 4);
                    pre_a = 0;
                    tc_fence_after();
                    TC_STAMP(tn, kc == 0 ? 5 : 6);
                    const uint64_t a_desc0 = make_desc(smem_u32(a_buf + sa * A_STAGE));
#pragma unroll
                    for (int ky = 0; ky < KS; ++ky) {
                        if (!pre_b && !b_ready) mbar_wait(&full_b[sb], pb, p.err, 5);
                        pre_b = 0;
                        tc_fence_after();
                        const uint64_t b_desc0 = make_desc(smem_u32(b_buf + sb * b_stage_bytes));
                        if (!(p.dbg & 1)) {
#pragma unroll
                            for (int kx = 0; kx < KS; ++kx) {
                                const uint64_t a_d = a_desc0 + (uint64_t)((kx * TC_A_COPY_BYTES + ky * (TC_TW * TC_ROW_BYTES)) >> 4);
                                const uint64_t b_d = b_desc0 + (uint64_t)(kx * b_step);
                                if (kx == KS - 1) {
                                    // last tap of the step: poll the NEXT step's barriers in the same asm block as its MMAs
                                    const int nsb = sb + 1 == nb_stages ? 0 : sb + 1;
                                    const int nsa = sa + 1 == NA ? 0 : sa + 1;
                                    const uint32_t ok = umma_x4_poll3<BF16>(d_tmem, a_d, 2, b_d, idesc, (kc | ky | kx) ? 1u : 0u,
                                                                            &full_b[nsb], nsb ? pb : pb ^ 1, &full_a[nsa], nsa ? pa : pa ^ 1,
                                                                            &tm_empty[acc + 1 == nacc ? 0 : acc + 1], acc + 1 == nacc ? pacc : pacc ^ 1);
                                    pre_b = ok & 1;
                                    if (ky == KS - 1) {
                                        pre_a = ok & 2;
                                        if (kc == kchunks - 1) pre_t = ok & 4;
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (BF16) umma_f16(d_tmem, a_d + 2 * k, b_d + 2 * k, idesc, (kc | ky | kx | k) ? 1u : 0u);
                                        else umma_tf32(d_tmem, a_d + 2 * k, b_d + 2 * k, idesc, (kc | ky | kx | k) ? 1u : 0u);
                                }
                            }
                        }
                        if (!b_resident) umma_commit(&empty_b[sb]);
                        if (++sb == nb_stages) { sb = 0; pb ^= 1; }
                    }
                    umma_commit(&empty_a[sa]);
                    if (++sa == NA) { sa = 0; pa ^= 1; }
                }
                umma_commit(&tm_full[acc]);
                TC_STAMP(tn, 7);
                if (++acc == nacc) { acc = 0; pacc ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps 3..: TMEM lane quarter = warp % 4; the warps of a quarter split the tile's 16-column
        // chunks.  The epilogue is a latency chain (tcgen05.ld -> scattered 32-byte stores that queue behind the
        // tensor core's operand fetches), so more warps in flight shorten it almost linearly. =====
        // p.epi_sets = S in {1, 2, 4}: the 16 warps form S sets; set s drains tiles s, s + S, ... (accumulator tile = tile
        // index mod 4), and inside a set the 4 / S warps of a lane quarter split the tile's 16-column chunks.  A warp's
        // per-tile time is a latency chain (barrier wake-up -> tcgen05.ld -> bias -> scattered stores) that more columns
        // barely lengthen, so several tiles in their epilogue at once raise the drain rate when the epilogue is the limit.
        const int S = p.epi_sets;
        const int q = warp & 3;
        const int g = (warp - 3) >> 2;                 // 0..3 among the warps of this lane quarter
        const int eset = g % S, eh = g / S;            // set, share of the columns inside the set
        const int nchunk = p.n_tile >> 4, wps = (TC_EPI_WARPS / 4) / S, cper = (nchunk + wps - 1) / wps;
        const int c_begin = min(eh * cper, nchunk), c_end = min(c_begin + cper, nchunk);
        const int m = q * 32 + lane;                   // pixel within the tile == TMEM lane
        const int ly = m / TC_TW, lx = m - ly * TC_TW;
        const float slope = p.act == FCVSR_ACT_PRELU ? p.slope_ptr[0] : p.slope;
        const int c4 = p.Cout >> 2;
        int acc = 0; uint32_t pacc = 0;
        // Staged output (p.stage_out): a lane's direct stores touch 32 different lines per instruction, and the LSU retires them
        // at about one line per 3 clk while the tensor core's operand fetches own the shared-memory pipe -- as long as the
        // tile's MMAs.  Instead the set writes the tile into a [128 pixels][128 B] SWIZZLE_128B buffer (conflict-free 16-byte
        // shared stores) and one elected thread hands it to the TMA unit, which writes whole lines and clips at the image border.
        const bool stg = p.stage_out != 0;
        const uint32_t stg_halves = (uint32_t)p.n_tile >> 6;      // 64-channel halves of a staged tile (1 or 2), 16 KB each
        const uint32_t srow = smem_u32(stg_buf + eset * (TC_STG_BYTES * stg_halves)) + (uint32_t)m * TC_ROW_BYTES;
        const uint32_t sxor = (uint32_t)(m & 7);
        const bool stager = eh == 0 && q == 0 && lane == 0;
        const int set_threads = 32 * TC_EPI_WARPS / S;
        asm volatile("griddepcontrol.wait;" ::: "memory");          // res may be, and y may still be read by, earlier kernels
        int tn = eset;
        for (int t = blockIdx.x + tn * gridDim.x; t < p.total_tiles; t += S * gridDim.x, tn += S) {
            acc = tn & (TC_NACC - 1);
            pacc = (uint32_t)(tn >> 2) & 1u;
            const TileCoord tc = decode_tile(t, p);
            const ConvTcProblem& pq = p.prob[tc.pr];
            const int y = tc.ty * TC_TH + ly, x = tc.tx * TC_TW + lx;
            const bool valid = y < pq.H && x < pq.W;
            const size_t pix = ((size_t)tc.b * pq.H + y) * pq.W + x;
            // per-tile copies of the problem's fields: p.prob[tc.pr] is an indexed constant-bank access (LDC with a register
            // index), which the chunk loop below would otherwise repeat four times per 16 columns
            const float* const res_p = valid ? pq.res : nullptr;
            const float* const res2_p = valid ? pq.res2 : nullptr;
            const int wrow = pq.wrow;

            mbar_wait_warp(&tm_full[acc], pacc, p.err, 6);
            tc_fence_after();
            if (warp == 3 && lane == 0) TC_STAMP(tn, 8);
            if (stg) {          // the set's previous tile must have left the staging buffer
                if (stager) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync %0, %1;" ::"r"(1 + eset), "r"(set_threads) : "memory");
            }
            if (warp == 3 && lane == 0) TC_STAMP(tn, 10);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.n_tile);
            // Software-pipelined TMEM reads: the tcgen05.ld of chunk c+1 is in flight while chunk c is
            // post-processed and stored (tcgen05.wait::ld sits right before the data is needed).
            auto process = [&](const uint32_t* r, const int cb) {
#ifdef FCVSR_BRINGUP
                if (p.dbg & 8) return;
#endif
                if (valid && p.cout_valid < p.Cout) {
                    // thin head (Cout in {1,4}): scalar stores of the first cout_valid columns
                    const int n0 = tc.nt * p.n_tile + cb;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + j;
                        if (n < p.cout_valid) {
                            float f = __uint_as_float(r[j]) + bias_s[n];
                            f = fcvsr_act(f, p.act, slope);
                            if (pq.res) f += pq.res[pix * p.ldres + n];
                            if (pq.res2) f -= pq.res2[pix * p.ldres2 + n];
                            pq.y[pix * p.ldy + n] = f;
                        }
                    }
                } else if (valid || stg) {
                    const int n0 = tc.nt * p.n_tile + cb;
                    float v[16];
                    lds_bias16(bias_sa + (wrow + n0) * 4, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r[j]);
                    // one uniform switch per chunk (per element it was 7 of the chunk's ~11 instructions per value)
                    if (p.act == FCVSR_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act != FCVSR_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * slope;
                    }
                    if (res_p) {
                        float rv[16];
                        if (p.wide) {
                            ld_global_v8(res_p + pix * p.ldres + n0, rv);
                            ld_global_v8(res_p + pix * p.ldres + n0 + 8, rv + 8);
                        } else {
                            const float4* rp = reinterpret_cast<const float4*>(res_p + pix * p.ldres + n0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { const float4 t4 = rp[j]; rv[4 * j] = t4.x; rv[4 * j + 1] = t4.y; rv[4 * j + 2] = t4.z; rv[4 * j + 3] = t4.w; }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += rv[j];
                    }
                    if (res2_p) {
                        const float4* rp = reinterpret_cast<const float4*>(res2_p + pix * p.ldres2 + n0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 rv = rp[j];
                            v[4 * j] -= rv.x; v[4 * j + 1] -= rv.y; v[4 * j + 2] -= rv.z; v[4 * j + 3] -= rv.w;
                        }
                    }
                    if (stg) {          // bf16 tile row of this pixel, 16-byte chunks XOR-swizzled by the row (SWIZZLE_128B)
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
                        const uint32_t ch = ((uint32_t)cb & 63u) >> 3;
                        const uint32_t hrow = srow + ((uint32_t)cb >> 6) * TC_STG_BYTES;
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hrow + ((ch ^ sxor) << 4)), "r"(w[0]), "r"(w[1]),
                                     "r"(w[2]), "r"(w[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hrow + (((ch + 1) ^ sxor) << 4)), "r"(w[4]), "r"(w[5]),
                                     "r"(w[6]), "r"(w[7]) : "memory");
                        return;
                    }
                    // operand-typed outputs: TF32-rounded fp32, or bf16 in bf16 mode
                    if (pq.y2) {
                        if (BF16) {
                            if (p.wide) store_bf16x16_v8(reinterpret_cast<__nv_bfloat16*>(pq.y2) + pix * p.ldy2 + n0, v);
                            else store_bf16x16(reinterpret_cast<__nv_bfloat16*>(pq.y2) + pix * p.ldy2 + n0, v);
                        } else if (p.wide) {
                            float vr[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) vr[j] = round_tf32(v[j]);
                            st_global_v8(pq.y2 + pix * p.ldy2 + n0, vr);
                            st_global_v8(pq.y2 + pix * p.ldy2 + n0 + 8, vr + 8);
                        } else {
                            float4* d2 = reinterpret_cast<float4*>(pq.y2 + pix * p.ldy2 + n0);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                d2[j] = make_float4(round_tf32(v[4 * j]), round_tf32(v[4 * j + 1]), round_tf32(v[4 * j + 2]),
                                                    round_tf32(v[4 * j + 3]));
                        }
                    }
                    size_t off;
                    if (p.ps) {
                        const int ij = n0 / c4, c = n0 - ij * c4;
                        const size_t opix = ((size_t)tc.b * 2 * pq.H + 2 * y + (ij >> 1)) * (2 * (size_t)pq.W) + 2 * x + (ij & 1);
                        off = opix * p.ldy + c;
                    } else {
                        off = pix * p.ldy + n0;
                    }
                    if (p.round_out == 2) {          // fp16 output tensor (ld in elements, 32-byte aligned rows)
                        store_f16x16_v8(reinterpret_cast<unsigned short*>(pq.y) + off, v);
                    } else if (BF16 && p.round_out) {
                        if (p.wide) store_bf16x16_v8(reinterpret_cast<__nv_bfloat16*>(pq.y) + off, v);
                        else store_bf16x16(reinterpret_cast<__nv_bfloat16*>(pq.y) + off, v);
                    } else {
                        if (p.round_out) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
                        }
                        if (p.wide) {
                            st_global_v8(pq.y + off, v);
                            st_global_v8(pq.y + off + 8, v + 8);
                        } else {
                            float4* dp = reinterpret_cast<float4*>(pq.y + off);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                }
            };
            {
                uint32_t ra[16], rb[16];
                if (c_begin < c_end) tmem_ld16(taddr + c_begin * 16, ra);
                for (int c = c_begin; c < c_end; c += 2) {
                    tmem_ld_wait();
                    if (warp == 3 && lane == 0 && c == c_begin) TC_STAMP(tn, 11);
                    const bool has_b = c + 1 < c_end;
                    if (has_b) tmem_ld16(taddr + (c + 1) * 16, rb);
                    process(ra, c * 16);
                    if (has_b) {
                        tmem_ld_wait();
                        if (c + 2 < c_end) tmem_ld16(taddr + (c + 2) * 16, ra);
                        process(rb, (c + 1) * 16);
                    }
                }
            }
            if (warp == 3 && lane == 0) TC_STAMP(tn, 12);
            if (stg) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> TMA (async proxy) reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tm_empty[acc]);
            if (stg) {
                asm volatile("bar.sync %0, %1;" ::"r"(1 + eset), "r"(set_threads) : "memory");
                if (stager) {
                    const CUtensorMap* my = tc.pr == 0 ? &map_y : (tc.pr == 1 ? &map_y1 : (tc.pr == 2 ? &map_y2 : &map_y3));
                    for (uint32_t h = 0; h < stg_halves; ++h)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(my), "r"(srow + h * TC_STG_BYTES), "r"(64 * h), "r"(tc.tx * TC_TW), "r"(tc.ty * TC_TH), "r"(tc.b)
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (warp == 3 && lane == 0) TC_STAMP(tn, 9);

        }
        if (stg && stager) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all tiles written before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TC_STAMP_NS(63);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

static int* tc_err_flag() {
    static int* flag = nullptr;
    if (!flag) {
        if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(flag, 0, sizeof(int));
    }
    return flag;
}

// Shared host side: `np` problems (same weights / epilogue, different spatial sizes) in one launch.
static int conv_tc_launch(int np, const void* const* xs, const float* const* ress, const float* const* res2s, float* const* ys,
                          float* const* y2s, const int* Hs, const int* Ws, int ldx, const float* w, const float* bias, int ldres,
                          int ldres2, int ldy, int B, int Cin, int Cout, int ksize, int act, float slope, const float* slope_ptr,
                          int pixel_shuffle, int ldy2, int round_out, int max_ctas, int op16, cudaStream_t st,
                          const int* wrows = nullptr, int w_rows = 0) {
    if (np < 1 || np > TC_MAX_PROB || !w || B <= 0) return FCVSR_ERR_ARG;
    const int kch = op16 ? 64 : 32, esz = op16 ? 2 : 4;
    if ((ksize != 1 && ksize != 3) || Cin % kch || Cin <= 0 || Cout <= 0) return FCVSR_ERR_UNSUPPORTED;
    if (op16 && (ldx & 7)) return FCVSR_ERR_UNSUPPORTED;
    const int cout_valid = Cout;
    const bool thin = Cout < 16;            // thin heads: w is zero-padded to 16 rows by the packer
    if (thin) Cout = 16;
    if (Cout % 16) return FCVSR_ERR_UNSUPPORTED;
    if (ldx & 3) return FCVSR_ERR_UNSUPPORTED;
    if ((uintptr_t)w & 15) return FCVSR_ERR_UNSUPPORTED;
    if (act == FCVSR_ACT_PRELU && !slope_ptr) return FCVSR_ERR_ARG;
    bool any_res = false, any_res2 = false, any_y2 = false;
    uintptr_t align_or = 0;
    for (int i = 0; i < np; ++i) {
        if (!xs[i] || !ys[i] || Hs[i] <= 0 || Ws[i] <= 0) return FCVSR_ERR_ARG;
        const float* res = ress ? ress[i] : nullptr;
        const float* res2 = res2s ? res2s[i] : nullptr;
        float* y2 = y2s ? y2s[i] : nullptr;
        any_res |= res != nullptr; any_res2 |= res2 != nullptr; any_y2 |= y2 != nullptr;
        if ((uintptr_t)xs[i] & 15) return FCVSR_ERR_UNSUPPORTED;
        if (!thin && (((uintptr_t)ys[i] | (uintptr_t)res | (uintptr_t)res2) & 15)) return FCVSR_ERR_UNSUPPORTED;
        if (y2 && ((uintptr_t)y2 & 15)) return FCVSR_ERR_UNSUPPORTED;
        if (round_out == 2 && ((uintptr_t)ys[i] & 31)) return FCVSR_ERR_UNSUPPORTED;
        align_or |= (uintptr_t)ys[i] | (uintptr_t)res | (uintptr_t)y2;
    }
    if (!thin && ((ldy & 3) || (any_res && (ldres & 3)) || (any_res2 && (ldres2 & 3)))) return FCVSR_ERR_UNSUPPORTED;
    if (thin && (pixel_shuffle || any_y2 || round_out)) return FCVSR_ERR_UNSUPPORTED;
    if (any_y2 && (pixel_shuffle || (ldy2 & 3))) return FCVSR_ERR_UNSUPPORTED;
    if (op16 && ((round_out && (ldy & 7)) || (any_y2 && (ldy2 & 7)))) return FCVSR_ERR_UNSUPPORTED;
    if (round_out == 2 && (thin || pixel_shuffle || (ldy & 15))) return FCVSR_ERR_UNSUPPORTED;
    int n_tile = Cout, n_tiles = 1;
    if (Cout > 128) {       // largest N tile <= 128 that is a multiple of 16 and divides Cout
        n_tile = 0;
        for (int cand = 128; cand >= 16; cand -= 16)
            if (Cout % cand == 0) { n_tile = cand; break; }
        if (!n_tile) return FCVSR_ERR_UNSUPPORTED;
        n_tiles = Cout / n_tile;
    }
    if (pixel_shuffle && ((Cout & 3) || ((Cout >> 2) % 16))) return FCVSR_ERR_UNSUPPORTED;
    if (Cout > TC_MAX_COUT) return FCVSR_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return FCVSR_ERR_CUDA;

    CUtensorMap map_x[TC_MAX_PROB], map_w;
    for (int i = 0; i < TC_MAX_PROB; ++i) {
        const int j = i < np ? i : 0;           // unused slots repeat problem 0 (the kernel never reads them)
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Ws[j], (cuuint64_t)Hs[j], (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)ldx * esz, (cuuint64_t)Ws[j] * ldx * esz, (cuuint64_t)Hs[j] * Ws[j] * ldx * esz};
        cuuint32_t box[4] = {(cuuint32_t)kch, TC_TW, (cuuint32_t)(ksize == 3 ? TC_TH + 2 : TC_TH), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (i >= np) { map_x[i] = map_x[0]; continue; }
        if (enc(&map_x[i], op16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)xs[j], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCVSR_ERR_CUDA;
    }
    {
        const int ktot = ksize * ksize * Cin;
        cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)(wrows ? w_rows : Cout)};
        cuuint64_t strides[1] = {(cuuint64_t)ktot * esz};
        cuuint32_t box[2] = {(cuuint32_t)kch, (cuuint32_t)n_tile};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&map_w, op16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCVSR_ERR_CUDA;
    }
    ConvTcParams p;
    p.bias = bias; p.ldres = ldres; p.ldres2 = ldres2; p.ldy = ldy; p.ldy2 = ldy2; p.round_out = round_out;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.ks = ksize; p.cout_valid = cout_valid;
    p.n_tile = n_tile; p.n_tiles = n_tiles;
    p.nprob = np;
    p.w_rows = wrows ? w_rows : Cout; p.multi_w = wrows != nullptr;
    if (wrows) {        // stacked filters: whole 64-channel 16-column-aligned filters, single N pass
        if (thin || n_tiles != 1 || pixel_shuffle || w_rows < Cout || w_rows > TC_MAX_COUT) return FCVSR_ERR_UNSUPPORTED;
        for (int i = 0; i < np; ++i)
            if (wrows[i] < 0 || (wrows[i] & 15) || wrows[i] + Cout > w_rows) return FCVSR_ERR_ARG;
    }
    int tiles = 0;
    for (int i = 0; i < TC_MAX_PROB; ++i) {
        ConvTcProblem& q = p.prob[i];
        const int j = i < np ? i : 0;
        q.res = ress ? ress[j] : nullptr; q.res2 = res2s ? res2s[j] : nullptr; q.y = ys[j]; q.y2 = y2s ? y2s[j] : nullptr;
        q.H = Hs[j]; q.W = Ws[j];
        q.wrow = wrows ? wrows[j] : 0;
        q.tiles_x = (q.W + TC_TW - 1) / TC_TW; q.tiles_y = (q.H + TC_TH - 1) / TC_TH;
        q.tile_begin = tiles;
        if (i < np) tiles += q.tiles_x * q.tiles_y * B * n_tiles;
    }
    p.total_tiles = tiles;
    {   // 1x1: four sets (epilogue-bound, 4-16 MMAs per tile); 3x3 bf16: two sets (64->64 24.8 -> 22.0 us, 64->256 at 360x640
        // 296 -> 272 us = 1.0 PFLOP/s); 3x3 TF32 (twice the MMAs per tile): one set, all warps on the same tile
        int es = 0;
#ifdef FCVSR_BRINGUP
        { static int es_env = -1; if (es_env < 0) { const char* e = getenv("FCVSR_EPI_SETS"); es_env = e ? atoi(e) : 0; } es = es_env; }
#endif
        p.epi_sets = es == 1 || es == 2 || es == 4 ? es : (ksize == 1 ? 4 : (op16 ? 2 : 1));
    }
    p.act = act; p.slope = slope; p.slope_ptr = slope_ptr; p.ps = pixel_shuffle;
    p.err = tc_err_flag();
    // bf16 outputs with 64 channels: the epilogue stages the tile in shared memory and stores it with TMA.  The staging tiles
    // (one per epilogue set) come out of the weight ring, which keeps at least the 72 KB that hold a 64 -> 64 3x3 filter resident.
    const int smem_fixed = 1024 + TC_NA * TC_A_STAGE_BYTES + 512 + 4 * p.w_rows;
    p.stage_out = op16 && round_out == 1 && (n_tile == 64 || n_tile == 128) && n_tiles == 1 && !pixel_shuffle && !any_y2 && !thin &&
                  !(ldy & 7);
    // a 128-channel tile is staged as two 64-channel halves (32 KB per set): two sets of eight warps, each warp one half
    if (p.stage_out && n_tile == 128 && p.epi_sets == 4) p.epi_sets = 2;
    const int stg_total = p.epi_sets * (n_tile >> 6) * TC_STG_BYTES;
    p.b_ring_bytes = TC_B_RING_BYTES;
    if (p.stage_out) {
        p.b_ring_bytes = ((TC_SMEM_LIMIT - smem_fixed) & ~1023) - stg_total;
        if (p.b_ring_bytes > TC_B_RING_BYTES) p.b_ring_bytes = TC_B_RING_BYTES;
        // only where the smaller ring still holds every weight stage at once (64 -> 64 3x3, the 1x1 convolutions): a streamed
        // filter (128 -> 64 3x3) measured slower with three ring stages than with four (40.0 vs 38.6 us at 4 x 180 x 320)
        if (p.b_ring_bytes < (Cin / kch) * ksize * ksize * n_tile * TC_ROW_BYTES) {
            p.stage_out = 0; p.b_ring_bytes = TC_B_RING_BYTES;
            if (ksize == 1 && p.epi_sets == 2 && n_tile == 128) p.epi_sets = 4;
        }
    }
    CUtensorMap map_y[TC_MAX_PROB];
    for (int i = 0; i < TC_MAX_PROB; ++i) {
        if (!p.stage_out) { map_y[i] = map_x[0]; continue; }      // never read
        const int j = i < np ? i : 0;
        if (i >= np) { map_y[i] = map_y[0]; continue; }
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Ws[j], (cuuint64_t)Hs[j], (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)Ws[j] * ldy * 2, (cuuint64_t)Hs[j] * Ws[j] * ldy * 2};
        cuuint32_t box[4] = {64, TC_TW, TC_TH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&map_y[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ys[j], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return FCVSR_ERR_CUDA;
    }
    {   // 256-bit epilogue accesses need 32-byte aligned rows for every tensor the epilogue touches
        const int esz_y = ((op16 && round_out) || round_out == 2) ? 2 : 4, esz_y2 = op16 ? 2 : 4;
        p.wide = !thin && !(align_or & 31) && !((ldy * esz_y) & 31) && (!any_res || !((ldres * 4) & 31)) &&
                 (!any_y2 || !((ldy2 * esz_y2) & 31)) && (!pixel_shuffle || !(((Cout >> 2) * esz_y) & 31));
    }
    p.dbg = 0;
#ifdef FCVSR_BRINGUP      // bring-up switches (tools/gpu_conv_trace.py builds its own copy with -DFCVSR_BRINGUP); not in the product library
    { static int narrow = -1; if (narrow < 0) narrow = getenv("FCVSR_TC_NARROW") ? 1 : 0; if (narrow) p.wide = 0; }
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("FCVSR_TC_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
#endif

    static int num_sms = 0;
    static bool attr_set = false;
    const size_t smem_max = TC_SMEM_LIMIT;
    const size_t smem = (size_t)smem_fixed + p.b_ring_bytes + (p.stage_out ? stg_total : 0);
    if (smem > smem_max) return FCVSR_ERR_UNSUPPORTED;
    if (!attr_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(conv_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess)
            return FCVSR_ERR_CUDA;
        attr_set = true;
    }
    int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;   // leave SMs to kernels running on other streams
    int pdl = 1;
#ifdef FCVSR_BRINGUP
    { static int pdl_env = -1; if (pdl_env < 0) { const char* e = getenv("FCVSR_PDL"); pdl_env = e ? atoi(e) : 1; } pdl = pdl_env; }
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t le;
    if (op16) {
        if (ksize == 3) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<3, true>, map_x[0], map_x[1], map_x[2], map_x[3], map_w, map_y[0], map_y[1], map_y[2], map_y[3], p);
        else le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, true>, map_x[0], map_x[1], map_x[2], map_x[3], map_w, map_y[0], map_y[1], map_y[2], map_y[3], p);
    } else {
        if (ksize == 3) le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<3, false>, map_x[0], map_x[1], map_x[2], map_x[3], map_w, map_y[0], map_y[1], map_y[2], map_y[3], p);
        else le = cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, false>, map_x[0], map_x[1], map_x[2], map_x[3], map_w, map_y[0], map_y[1], map_y[2], map_y[3], p);
    }
    if (le != cudaSuccess) return FCVSR_ERR_CUDA;
    return fcvsr_launch_status();
}

extern "C" int fcvsr_conv2d_tc(const float* x, int ldx, const float* w, const float* bias, const float* res, int ldres,
                               const float* res2, int ldres2, float* y, int ldy, int B, int H, int W, int Cin, int Cout,
                               int ksize, int act, float slope, const float* slope_ptr, int pixel_shuffle,
                               float* y2, int ldy2, int round_out, int max_ctas, int op16, cudaStream_t st) {
    if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0) return FCVSR_ERR_ARG;
    const void* xs[1] = {x};
    const float* ress[1] = {res};
    const float* res2s[1] = {res2};
    float* ys[1] = {y};
    float* y2s[1] = {y2};
    return conv_tc_launch(1, xs, ress, res2s, ys, y2s, &H, &W, ldx, w, bias, ldres, ldres2, ldy, B, Cin, Cout, ksize, act, slope,
                          slope_ptr, pixel_shuffle, ldy2, round_out, max_ctas, op16, st);
}

// The same convolution on up to three tensors of different spatial size in ONE launch (the pyramid levels of SCNetbk,
// CVSR_freq.py:766-770: BlockRCB applies one body to every level).  x / res / y / y2: HOST arrays of nprob device pointers
// (res, y2 may be NULL or hold NULL entries); H, W: HOST arrays; ld*, B, channels, epilogue as fcvsr_conv2d_tc.
extern "C" int fcvsr_conv2d_tc_multi(int nprob, const void* const* x, int ldx, const float* w, const float* bias,
                                     const float* const* res, int ldres, float* const* y, int ldy, const int* H, const int* W,
                                     int B, int Cin, int Cout, int ksize, int act, float slope, const float* slope_ptr,
                                     float* const* y2, int ldy2, int round_out, int op16, cudaStream_t st) {
    if (!x || !y || !H || !W) return FCVSR_ERR_ARG;
    return conv_tc_launch(nprob, x, res, nullptr, y, y2, H, W, ldx, w, bias, ldres, 0, ldy, B, Cin, Cout, ksize, act, slope,
                          slope_ptr, 0, ldy2, round_out, 0, op16, st);
}

// As fcvsr_conv2d_tc_multi with a different filter per problem: `w` / `bias` hold w_rows >= Cout rows (several [Cout][k*k*Cin]
// filters stacked) and problem i uses rows wrow[i] .. wrow[i] + Cout (multiples of 16; HOST array).  Up to four problems: the
// 1x1 `down` and `up` convolutions of a BlockRCB (CVSR_freq.py:753-763) -- two filters on two pyramid levels each -- are one launch.
extern "C" int fcvsr_conv2d_tc_multi_w(int nprob, const void* const* x, int ldx, const float* w, const float* bias, const int* wrow,
                                       int w_rows, float* const* y, int ldy, const int* H, const int* W, int B, int Cin, int Cout,
                                       int ksize, int act, float slope, int round_out, int op16, cudaStream_t st) {
    if (!x || !y || !H || !W || !wrow) return FCVSR_ERR_ARG;
    return conv_tc_launch(nprob, x, nullptr, nullptr, y, nullptr, H, W, ldx, w, bias, 0, 0, ldy, B, Cin, Cout, ksize, act, slope,
                          nullptr, 0, 0, round_out, 0, op16, st, wrow, w_rows);
}

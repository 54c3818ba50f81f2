// Kernels specific to the motion-guided adaptive alignment (MGAAbk, CVSR_freq.py:1365-1547).
//
//   corr_gather      CorrBlock lookup (:1279-1337) -- a pure gather, see oracle.corr_lookup
//   offset_blk_*     ConvBlk(dim=4, index=i) on the 4-channel frequency maps (:344-357, :1494-1496)
//                    for all ACNum iterations and both directions in one launch per stage
//   iac_step         one IAC iteration (:1230-1250): flow_warp (:1188-1227) + SAC (:1253-1276)
//                    + residual + LeakyReLU(0.1), fused in one pass with the warped tile and the
//                    vertical-pass tile held in shared memory
//
// Spectra layout used by this library: complex-interleaved NHWC, [B, H, Wf, G*128] floats where
// group g (x1, x2, x3) occupies floats [g*128, g*128+128) and channel c of the group sits at
// (2c: real, 2c+1: imag).  The reference's xk_f = cat([imag, real]) channel j therefore maps to
// float index  j < 64 ? 2j+1 : 2(j-64)  -- the host applies that permutation to the 1x1 weights.
#include "common.cuh"
#include <cuda_fp16.h>

// ------------------------------------------------------------------------------------------------
// CorrBlock gather.  S: [B, P, ldS] floats (P = H*Wf), groups a_off / b_off (float offsets) hold the
// two spectra.  out: [B, P, ldo], 81 channels.  prod (reference layout [C2][P]) flat index F = p*C2
// + row*2 + col  ->  ref channel F / P, position F % P.
// One thread per (position, i) = 9 consecutive output channels j = 0..8 (same column offset, rows y0-4 .. y0+4).  Only positions
// with x0 <= 5 and y0 <= C2/2 + 3 can hit the 64 x 2 "image" at all: everything else is a plain zero store, without the 64-bit
// index arithmetic (the first version spent 39 us per launch on it for 9 MB of traffic).  out2 (optional): second copy of
// the result with the same layout -- the lookup feeds both offset branches (CVSR_freq.py:1487-1488).
__global__ void corr_gather_kernel(const float* __restrict__ S, int ldS, int a_off, int b_off, void* __restrict__ out,
                                   void* __restrict__ out2, int ldo, int H, int Wf, int C2, float inv_sqrt_c, int total,
                                   int op_mode) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int P = H * Wf;
    const int i = idx % 9;
    const int bp = idx / 9;
    const int p = bp % P, b = bp / P;
    const int y0 = p / Wf, x0 = p - y0 * Wf;
    const int col = x0 + i - 4;
    const int half = C2 / 2;
    float v[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) v[j] = 0.f;
    if (col >= 0 && col < 2 && y0 - 4 < half) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int row = y0 + j - 4;
            if (row >= 0 && row < half) {
                const long long F = (long long)p * C2 + row * 2 + col;
                const int ch = (int)(F / P), p2 = (int)(F - (long long)ch * P);
                const int mi = ch < half ? 2 * ch + 1 : 2 * (ch - half);
                const float* s = S + ((size_t)b * P + p2) * ldS;
                v[j] = s[a_off + mi] * s[b_off + mi] * inv_sqrt_c;
            }
        }
    }
    // op_mode: 0 plain fp32, 1 TF32-rounded fp32, 2 bf16 (the lookup feeds convcorr[0] on the tensor cores)
    const size_t o = ((size_t)b * P + p) * ldo + i * 9;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        if (op_mode == 0) {
            reinterpret_cast<float*>(out)[o + j] = v[j];
            if (out2) reinterpret_cast<float*>(out2)[o + j] = v[j];
        } else {
            store_operand1(out, o + j, v[j], op_mode == 2);
            if (out2) store_operand1(out2, o + j, v[j], op_mode == 2);
        }
    }
}

// Vector-store variant (op_mode bit 2: the caller allows zero-filling the padding channels up to the next multiple of 8 (bf16) /
// 4 (fp32) behind the 81 real ones, and the rows are 16-byte aligned).  The lookup is almost all zeros -- only positions with
// x0 <= 5 and y0 < C2/2 + 4 can hit the 64 x 2 "image" -- so the kernel is a store kernel: one 16-byte store per (position,
// channel group) and tensor instead of nine scalar stores per thread (95 us -> memory speed at 6 x 180 x 161).
template <bool OP16>
__global__ void corr_gather_vec_kernel(const float* __restrict__ S, int ldS, int a_off, int b_off, void* __restrict__ out,
                                       void* __restrict__ out2, int ldo, int H, int Wf, int C2, float inv_sqrt_c, long long total,
                                       int nchunk, int round32) {
    constexpr int CPC = OP16 ? 8 : 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int P = H * Wf;
    const int ck = (int)(idx % nchunk);
    const long long bp = idx / nchunk;
    const int p = (int)(bp % P), b = (int)(bp / P);
    const int y0 = p / Wf, x0 = p - y0 * Wf;
    const int half = C2 / 2;
    float v[CPC];
#pragma unroll
    for (int e = 0; e < CPC; ++e) v[e] = 0.f;
    if (x0 <= 5 && y0 - 4 < half) {
#pragma unroll
        for (int e = 0; e < CPC; ++e) {
            const int k = ck * CPC + e;                 // output channel = i * 9 + j
            if (k >= 81) continue;
            const int i = k / 9, j = k - i * 9;
            const int col = x0 + i - 4, row = y0 + j - 4;
            if (col >= 0 && col < 2 && row >= 0 && row < half) {
                const long long F = (long long)p * C2 + row * 2 + col;
                const int ch = (int)(F / P), p2 = (int)(F - (long long)ch * P);
                const int mi = ch < half ? 2 * ch + 1 : 2 * (ch - half);
                const float* s = S + ((size_t)b * P + p2) * ldS;
                v[e] = s[a_off + mi] * s[b_off + mi] * inv_sqrt_c;
            }
        }
    }
    const size_t o = ((size_t)b * P + p) * ldo + ck * CPC;
    if (OP16) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[e]) : "f"(v[2 * e + 1]), "f"(v[2 * e]));
        const uint4 u = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out) + o) = u;
        if (out2) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out2) + o) = u;
    } else {
        float4 u = make_float4(v[0], v[1], v[2], v[3]);
        if (round32) u = make_float4(round_tf32(u.x), round_tf32(u.y), round_tf32(u.z), round_tf32(u.w));
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = u;
        if (out2) *reinterpret_cast<float4*>(reinterpret_cast<float*>(out2) + o) = u;
    }
}

extern "C" int fcvsr_corr_gather2(const float* S, int ldS, int a_off, int b_off, void* out, void* out2, int ldo, int B, int H,
                                  int Wf, int C2, int op_mode, cudaStream_t st) {
    if (!S || !out || C2 <= 0) return FCVSR_ERR_ARG;
    if (op_mode & 4) {
        const int om = op_mode & 3, esz = om == 2 ? 2 : 4, cpc = om == 2 ? 8 : 4, nchunk = (81 + cpc - 1) / cpc;
        if (om > 2 || ldo < nchunk * cpc || ((ldo * esz) & 15) || ((uintptr_t)out & 15) || ((uintptr_t)out2 & 15)) return FCVSR_ERR_ARG;
        const long long totalv = (long long)B * H * Wf * nchunk;
        const unsigned grid = (unsigned)((totalv + 255) / 256);
        if (om == 2) corr_gather_vec_kernel<true><<<grid, 256, 0, st>>>(S, ldS, a_off, b_off, out, out2, ldo, H, Wf, C2, rsqrtf((float)C2), totalv, nchunk, 0);
        else corr_gather_vec_kernel<false><<<grid, 256, 0, st>>>(S, ldS, a_off, b_off, out, out2, ldo, H, Wf, C2, rsqrtf((float)C2), totalv, nchunk, om == 1);
        return fcvsr_launch_status();
    }
    const long long total = (long long)B * H * Wf * 9;
    if (total > 0x7fffffffLL) return FCVSR_ERR_ARG;
    corr_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S, ldS, a_off, b_off, out, out2, ldo, H, Wf, C2,
                                                                      rsqrtf((float)C2), (int)total, op_mode);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_corr_gather(const float* S, int ldS, int a_off, int b_off, void* out, int ldo, int B, int H,
                                 int Wf, int C2, int op_mode, cudaStream_t st) {
    return fcvsr_corr_gather2(S, ldS, a_off, b_off, out, nullptr, ldo, B, H, Wf, C2, op_mode, st);
}

// ------------------------------------------------------------------------------------------------
// ConvBlk on 4-channel maps.  `off` is [2][B][H*Wf][4] (direction-major: forward, backward).
// Weights of iteration i (kernel size k = 2i+1) live at w + 16 * sum_{t<i} (2t+1)^2, packed
// [tap][ci][co].  blockIdx = (pos chunk, i, b*2+dir).
#define OB_THREADS 128
__device__ __forceinline__ int ob_woff(int i) {   // 16 * sum_{t<i} (2t+1)^2 = 16 * i(2i-1)(2i+1)/3
    return 16 * (i * (2 * i - 1) * (2 * i + 1) / 3);
}

// Tile = 8 rows x 64 columns of one (iteration, direction, image); thread = (row, quad of 4 consecutive columns), all 4 output
// channels.  The haloed input tile ((8 + K - 1) x (64 + K - 1) float4 pixels) is staged in shared memory: gathering the windows
// straight from global memory had neighbouring lanes 64 bytes apart, so every warp load touched 16 lines at 25 % sector use and
// the kernel was L1-bound (ncu: l1tex 82 %, FMA pipe 43 %, 33.9 M sectors for a 5.6 MB input).  In shared memory the pixels of
// a row are de-interleaved by (x mod 4): window element d of quad c lives in array d % 4, slot c + d / 4, so the 16 quads of a
// half-warp read 16 consecutive 16-byte slots -- conflict-free.  Per filter row a thread holds its K + 3 window pixels in
// registers and walks the K taps with the tap's 16 weights loaded once (4 broadcast ld.shared.v4 per 64 FMAs).
#define OB_TH 8
#define OB_TQ 16                                   // quads per tile row
#define OB_TW (4 * OB_TQ)
#define OB_KMAX 11
#define OB_SLOTS (OB_TQ + (OB_KMAX + 2) / 4 + 1)   // 16-byte slots per (row, x mod 4) array
#define OB_ROWF4 (4 * OB_SLOTS)                    // float4 per staged row

template <int K, bool SECOND>
__device__ __forceinline__ void ob_conv_quad(const float4* __restrict__ tile, const float* __restrict__ ws, int r, int c,
                                             float (&acc)[4][4]) {
#pragma unroll 1
    for (int ky = 0; ky < K; ++ky) {
        const float4* rowp = tile + (r + ky) * OB_ROWF4 + c;
        float4 row[K + 3];
#pragma unroll
        for (int j = 0; j < K + 3; ++j) row[j] = rowp[(j & 3) * OB_SLOTS + (j >> 2)];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const float4* wt = reinterpret_cast<const float4*>(ws + (ky * K + kx) * 16);
            const float4 w0 = wt[0], w1 = wt[1], w2 = wt[2], w3 = wt[3];       // rows ci = 0..3, columns co
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = row[j + kx];
                // one FMA per MAC, chained into the accumulator (the sum-then-add form costs a fifth instruction per 4 MACs)
                acc[j][0] = fmaf(v.w, w3.x, fmaf(v.z, w2.x, fmaf(v.y, w1.x, fmaf(v.x, w0.x, acc[j][0]))));
                acc[j][1] = fmaf(v.w, w3.y, fmaf(v.z, w2.y, fmaf(v.y, w1.y, fmaf(v.x, w0.y, acc[j][1]))));
                acc[j][2] = fmaf(v.w, w3.z, fmaf(v.z, w2.z, fmaf(v.y, w1.z, fmaf(v.x, w0.z, acc[j][2]))));
                acc[j][3] = fmaf(v.w, w3.w, fmaf(v.z, w2.w, fmaf(v.y, w1.w, fmaf(v.x, w0.w, acc[j][3]))));
            }
        }
    }
}

template <bool SECOND>
__global__ void __launch_bounds__(OB_THREADS) offset_blk_conv_kernel(const float* __restrict__ in, size_t in_iter_stride,
                                                                    const float* __restrict__ w,
                                                                    const float* __restrict__ prelu,   // [A] (stage 1)
                                                                    float* __restrict__ out, size_t out_iter_stride,
                                                                    float* __restrict__ partial,       // stage 2
                                                                    int H, int Wf, int tiles_x) {
    __shared__ __align__(16) float ws[121 * 16];
    __shared__ __align__(16) float4 tile[(OB_TH + OB_KMAX - 1) * OB_ROWF4];
    __shared__ float red[OB_THREADS / 32][4];
    // the iteration is the slowest grid dimension, largest kernel first (k = 11 costs 121x k = 1): the light blocks fill the tail
    const int it = gridDim.z - 1 - blockIdx.z, k = 2 * it + 1, pad = it;
    const int bz = blockIdx.y, b = bz >> 1, dir = bz & 1;
    const int B = gridDim.y >> 1;
    const int ty0 = (blockIdx.x / tiles_x) * OB_TH, tx0 = (blockIdx.x % tiles_x) * OB_TW;
    const int P = H * Wf;
    const float* wi = w + ob_woff(it);
    for (int t = threadIdx.x; t < k * k * 16; t += blockDim.x) ws[t] = wi[t];
    {   // haloed input tile, zero outside the image (the convolution's zero padding), rows de-interleaved by (x mod 4)
        const float* src = in + (SECOND ? (size_t)it * in_iter_stride : 0) + ((size_t)dir * B + b) * P * 4;
        const int tw = OB_TW + k - 1, th = OB_TH + k - 1;
        for (int e = threadIdx.x; e < th * tw; e += blockDim.x) {
            const int rr = e / tw, pp = e - rr * tw;
            const int yy = ty0 - pad + rr, xx = tx0 - pad + pp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < Wf) v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)yy * Wf + xx) * 4));
            tile[rr * OB_ROWF4 + (pp & 3) * OB_SLOTS + (pp >> 2)] = v;
        }
    }
    __syncthreads();
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
    const int y = ty0 + r, x0 = tx0 + 4 * c;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int co = 0; co < 4; ++co) acc[j][co] = 0.f;
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    if (y < H && x0 < Wf) {
        switch (it) {
            case 0: ob_conv_quad<1, SECOND>(tile, ws, r, c, acc); break;
            case 1: ob_conv_quad<3, SECOND>(tile, ws, r, c, acc); break;
            case 2: ob_conv_quad<5, SECOND>(tile, ws, r, c, acc); break;
            case 3: ob_conv_quad<7, SECOND>(tile, ws, r, c, acc); break;
            case 4: ob_conv_quad<9, SECOND>(tile, ws, r, c, acc); break;
            default: ob_conv_quad<11, SECOND>(tile, ws, r, c, acc); break;
        }
        const float sl = SECOND ? 0.f : prelu[it];
        float* dst = out + (size_t)it * out_iter_stride + (((size_t)dir * B + b) * P + (size_t)y * Wf + x0) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (x0 + j >= Wf) continue;
            if (!SECOND) {
#pragma unroll
                for (int co = 0; co < 4; ++co) acc[j][co] = acc[j][co] >= 0.f ? acc[j][co] : acc[j][co] * sl;
            }
            *reinterpret_cast<float4*>(dst + j * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
#pragma unroll
            for (int co = 0; co < 4; ++co) sum[co] += acc[j][co];
        }
    }
    if (SECOND) {
        // deterministic per-block partial sums for the CALayer mean
#pragma unroll
        for (int co = 0; co < 4; ++co) {
            float s = warp_sum(sum[co]);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][co] = s;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            float s = 0.f;
            for (int wv = 0; wv < OB_THREADS / 32; ++wv) s += red[wv][threadIdx.x];
            // partial[it][b*2+dir][blk][4]
            partial[(((size_t)it * gridDim.y + bz) * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = s;
        }
    }
}

// Stage 3: CALayer(4, r=1) gate from the partial sums, (gate+1) * t2 * sim, packed as the complex
// input of irfft2: complex channel (it*2+dir)*2 + m = (v[m], v[2+m])  (:1497-1498).
__global__ void __launch_bounds__(OB_THREADS) offset_blk_finish_kernel(const float* __restrict__ t2, size_t iter_stride,
                                                                      const float* __restrict__ partial, int nblk,
                                                                      const float* __restrict__ ca_w,   // [A][2][4][4]
                                                                      const float* __restrict__ sim, int ldsim,
                                                                      float* __restrict__ z, int A, int H, int Wf) {
    __shared__ float gate[4];
    const int it = blockIdx.y;
    const int b = blockIdx.z >> 1, dir = blockIdx.z & 1;
    const int P = H * Wf;
    if (threadIdx.x < 32) {
        float m[4];
        const float* pp = partial + ((size_t)it * gridDim.z + blockIdx.z) * nblk * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float s = 0.f;
            for (int t = threadIdx.x; t < nblk; t += 32) s += pp[t * 4 + c];
            m[c] = warp_sum(s) / (float)P;
        }
        if (threadIdx.x < 4) {
            const float* w1 = ca_w + it * 32;
            const float* w2 = w1 + 16;
            float hdn[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) s += w1[o * 4 + c] * m[c];
                hdn[o] = fmaxf(s, 0.f);
            }
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) s += w2[threadIdx.x * 4 + c] * hdn[c];
            gate[threadIdx.x] = 1.f + 1.f / (1.f + __expf(-s));
        }
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int B = gridDim.z >> 1;
    const float4 v = *reinterpret_cast<const float4*>(t2 + (size_t)it * iter_stride + (((size_t)dir * B + b) * P + p) * 4);
    const float4 s = *reinterpret_cast<const float4*>(sim + ((size_t)b * P + p) * ldsim);
    const float o0 = v.x * gate[0] * s.x, o1 = v.y * gate[1] * s.y, o2 = v.z * gate[2] * s.z, o3 = v.w * gate[3] * s.w;
    float* dst = z + (((size_t)b * P + p) * (4 * A) + (it * 2 + dir) * 2) * 2;
    *reinterpret_cast<float4*>(dst) = make_float4(o0, o2, o1, o3);
}

extern "C" int fcvsr_offset_blocks(const float* off, const float* w1, const float* w2, const float* prelu,
                                   const float* ca_w, const float* sim, int ldsim, float* t1, float* t2,
                                   float* partial, float* z, int B, int H, int Wf, int A, cudaStream_t st) {
    if (!off || !w1 || !w2 || !prelu || !ca_w || !sim || !t1 || !t2 || !partial || !z || A < 1 || A > 6 || (ldsim & 3))
        return FCVSR_ERR_ARG;
    const int P = H * Wf;
    const int nblk = (P + OB_THREADS - 1) / OB_THREADS;
    // conv blocks: 8 x 64 tiles, one CALayer partial sum each (`partial` holds max(nblk, nqblk) entries per iteration and image)
    const int tiles_x = (Wf + OB_TW - 1) / OB_TW, nqblk = tiles_x * ((H + OB_TH - 1) / OB_TH);
    dim3 grid(nblk, A, B * 2), gridq(nqblk, B * 2, A);
    const size_t iter_stride = (size_t)B * P * 8;
    offset_blk_conv_kernel<false><<<gridq, OB_THREADS, 0, st>>>(off, 0, w1, prelu, t1, iter_stride, nullptr, H, Wf, tiles_x);
    offset_blk_conv_kernel<true><<<gridq, OB_THREADS, 0, st>>>(t1, iter_stride, w2, nullptr, t2, iter_stride, partial, H, Wf, tiles_x);
    offset_blk_finish_kernel<<<grid, OB_THREADS, 0, st>>>(t2, iter_stride, partial, nqblk, ca_w, sim, ldsim, z, A, H, Wf);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------
// The same 4 -> 4 channel convolution as a stand-alone operator for the training step (the ConvBlk convolutions of
// CVSR_freq.py:344-357 under autograd): forward, data gradient (the forward kernel on flipped / transposed weights) and weight
// gradient.  The generic CUDA-core kernels tile 64 pixels x 64 channels, so a 4 x 4 filter used 1/16 .. 1/256 of their work
// (11 x 11: 387 us forward, 1.5 ms backward at 16 x 64 x 33; 36 such convolutions per training step).
template <int DUMMY>
__global__ void __launch_bounds__(OB_THREADS) conv4_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                                           int H, int Wd, int tiles_x, int k) {
    __shared__ __align__(16) float ws[121 * 16];
    __shared__ __align__(16) float4 tile[(OB_TH + OB_KMAX - 1) * OB_ROWF4];
    const int pad = k >> 1, b = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * OB_TH, tx0 = (blockIdx.x % tiles_x) * OB_TW;
    const size_t P = (size_t)H * Wd;
    for (int t = threadIdx.x; t < k * k * 16; t += blockDim.x) ws[t] = w[t];
    {
        const float* src = x + (size_t)b * P * 4;
        const int tw = OB_TW + k - 1, th = OB_TH + k - 1;
        for (int e = threadIdx.x; e < th * tw; e += blockDim.x) {
            const int rr = e / tw, pp = e - rr * tw;
            const int yy = ty0 - pad + rr, xx = tx0 - pad + pp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < Wd) v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)yy * Wd + xx) * 4));
            tile[rr * OB_ROWF4 + (pp & 3) * OB_SLOTS + (pp >> 2)] = v;
        }
    }
    __syncthreads();
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
    const int yo = ty0 + r, x0 = tx0 + 4 * c;
    if (yo >= H || x0 >= Wd) return;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int co = 0; co < 4; ++co) acc[j][co] = 0.f;
    switch (k) {
        case 1: ob_conv_quad<1, false>(tile, ws, r, c, acc); break;
        case 3: ob_conv_quad<3, false>(tile, ws, r, c, acc); break;
        case 5: ob_conv_quad<5, false>(tile, ws, r, c, acc); break;
        case 7: ob_conv_quad<7, false>(tile, ws, r, c, acc); break;
        case 9: ob_conv_quad<9, false>(tile, ws, r, c, acc); break;
        default: ob_conv_quad<11, false>(tile, ws, r, c, acc); break;
    }
    float* dst = y + ((size_t)b * P + (size_t)yo * Wd + x0) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (x0 + j < Wd) *reinterpret_cast<float4*>(dst + j * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
}

// x, y: [B,H,W,4] fp32; w: [k*k][ci][co] (the layout of fcvsr_conv2d_direct); k odd, <= 11, zero padding k / 2, stride 1.
// The data gradient is the same call with dy as x and w'[tap'][co][ci] = w[k*k - 1 - tap'][ci][co].
extern "C" int fcvsr_conv4x4(const float* x, const float* w, float* y, int B, int H, int W, int ksize, cudaStream_t st) {
    if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0 || !(ksize & 1) || ksize < 1 || ksize > OB_KMAX) return FCVSR_ERR_ARG;
    if (((uintptr_t)x | (uintptr_t)y) & 15) return FCVSR_ERR_ARG;
    const int tiles_x = (W + OB_TW - 1) / OB_TW, tiles = tiles_x * ((H + OB_TH - 1) / OB_TH);
    if (B > 65535) return FCVSR_ERR_UNSUPPORTED;
    conv4_kernel<0><<<dim3(tiles, B), OB_THREADS, 0, st>>>(x, w, y, H, W, tiles_x, ksize);
    return fcvsr_launch_status();
}

// Weight gradient dw[tap][ci][co] += sum_p x[p + tap - pad][ci] * dy[p][co] of the same convolution.  A thread owns ONE filter
// tap (and, when there are fewer than 64 taps, one slice of the tile's rows) and walks the tile's pixels: x comes from the
// haloed tile in shared memory (neighbouring taps = neighbouring pixels), dy is a broadcast read, 16 FMAs per pixel, no
// reduction across threads; one vector of 16 atomics per thread and tile at the end.
__global__ void __launch_bounds__(OB_THREADS) conv4_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 float* __restrict__ dw, int H, int Wd, int tiles_x, int k) {
    constexpr int XP = OB_TW + OB_KMAX - 1;                       // x tile pitch in float4
    __shared__ __align__(16) float4 xt[(OB_TH + OB_KMAX - 1) * XP];
    __shared__ __align__(16) float4 gt[OB_TH * OB_TW];
    const int pad = k >> 1, b = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * OB_TH, tx0 = (blockIdx.x % tiles_x) * OB_TW;
    const size_t P = (size_t)H * Wd;
    const float* xs = x + (size_t)b * P * 4;
    const float* gs = dy + (size_t)b * P * 4;
    const int tw = OB_TW + k - 1, th = OB_TH + k - 1;
    for (int e = threadIdx.x; e < th * tw; e += blockDim.x) {
        const int rr = e / tw, pp = e - rr * tw;
        const int yy = ty0 - pad + rr, xx = tx0 - pad + pp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < Wd) v = __ldg(reinterpret_cast<const float4*>(xs + ((size_t)yy * Wd + xx) * 4));
        xt[rr * XP + pp] = v;
    }
    for (int e = threadIdx.x; e < OB_TH * OB_TW; e += blockDim.x) {
        const int rr = e / OB_TW, pp = e - rr * OB_TW;
        const int yy = ty0 + rr, xx = tx0 + pp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy < H && xx < Wd) v = __ldg(reinterpret_cast<const float4*>(gs + ((size_t)yy * Wd + xx) * 4));
        gt[e] = v;
    }
    __syncthreads();
    const int nt = k * k;
    const int groups = OB_THREADS / nt > OB_TH ? OB_TH : (OB_THREADS / nt < 1 ? 1 : OB_THREADS / nt);   // row slices
    const int tap = threadIdx.x % nt, slice = threadIdx.x / nt;
    if (slice >= groups) return;
    const int ky = tap / k, kx = tap - ky * k;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int r = slice; r < OB_TH; r += groups) {
        const float4* xr = xt + (r + ky) * XP + kx;
        const float4* gr = gt + r * OB_TW;
#pragma unroll 4
        for (int c = 0; c < OB_TW; ++c) {
            const float4 xv = xr[c], gv = gr[c];
            acc[0] = fmaf(xv.x, gv.x, acc[0]); acc[1] = fmaf(xv.x, gv.y, acc[1]); acc[2] = fmaf(xv.x, gv.z, acc[2]); acc[3] = fmaf(xv.x, gv.w, acc[3]);
            acc[4] = fmaf(xv.y, gv.x, acc[4]); acc[5] = fmaf(xv.y, gv.y, acc[5]); acc[6] = fmaf(xv.y, gv.z, acc[6]); acc[7] = fmaf(xv.y, gv.w, acc[7]);
            acc[8] = fmaf(xv.z, gv.x, acc[8]); acc[9] = fmaf(xv.z, gv.y, acc[9]); acc[10] = fmaf(xv.z, gv.z, acc[10]); acc[11] = fmaf(xv.z, gv.w, acc[11]);
            acc[12] = fmaf(xv.w, gv.x, acc[12]); acc[13] = fmaf(xv.w, gv.y, acc[13]); acc[14] = fmaf(xv.w, gv.z, acc[14]); acc[15] = fmaf(xv.w, gv.w, acc[15]);
        }
    }
    float* d = dw + (size_t)tap * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(d + i, acc[i]);
}

// dw: [k*k][4][4] fp32, ACCUMULATED with atomics (zero it first), as fcvsr_conv2d_wgrad.
extern "C" int fcvsr_conv4x4_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int ksize, cudaStream_t st) {
    if (!x || !dy || !dw || B <= 0 || H <= 0 || W <= 0 || !(ksize & 1) || ksize < 1 || ksize > OB_KMAX) return FCVSR_ERR_ARG;
    if (((uintptr_t)x | (uintptr_t)dy) & 15) return FCVSR_ERR_ARG;
    const int tiles_x = (W + OB_TW - 1) / OB_TW, tiles = tiles_x * ((H + OB_TH - 1) / OB_TH);
    if (B > 65535) return FCVSR_ERR_UNSUPPORTED;
    conv4_wgrad_kernel<<<dim3(tiles, B), OB_THREADS, 0, st>>>(x, dy, dw, H, W, tiles_x, ksize);
    return fcvsr_launch_status();
}

// ------------------------------------------------------------------------------------------------
// One IAC iteration for both directions.  Tile = 8 x 16 output pixels x 64 channels; a half-warp owns a
// pixel at a time and a lane owns 4 consecutive channels (16-byte accesses, two pixels per warp instruction).
//   samp(y',x')  = bilinear(prev, x'+dx(y',x'), y'+dy(y',x'))        zeros outside, align_corners
//   v(y,x')      = sum_t K[t](y,x') * samp(clamp(y+t-1), x')         (vertical pass, replicate pad)
//   out(y,x)     = lrelu_0.1( sum_t K[t](y,x) * v(y, clamp(x+t-1)) + xin(y,x) )
// K[t] of this iteration = taps[..., t*64 + c] (host packs F.1's live rows as [iter][t][c]).
// The kernel is a chain of dependent global loads (offsets -> gathers; taps -> passes), so it is organised for
// memory-level parallelism: every thread has 8 gather loads in flight in phase 1, and in phases 2/3 a half-warp
// owns up to five columns of ONE tile row, loads their taps once (kept in registers for both the vertical and
// the horizontal pass -- the reference applies kernel1 twice, :1253-1276) with all loads issued up front.
#define IAC_TH 8
#define IAC_TW 16
#define IAC_C 64
struct IacArgs {
    const float* prev[2]; int ldprev[2];
    const float* xin[2];  int ldxin[2];
    float* next[2];       int ldnext[2];
    const float* offs; int ldoffs; int offs_ch[2];   // channel of dx for each direction
    const float* taps; int ldtaps;                   // already offset to this iteration's 192 channels
    int B, H, W; int round_out;      // 0 fp32, 1 TF32-rounded fp32, 2 bf16 (next[] is then a bf16 tensor)
    int prev16;                      // prev[] are bf16 tensors (ld in elements): the IAC ping-pong of the bf16 mode
};

#define IAC_THREADS 512
#define IAC_HW (IAC_THREADS / 16)                   // half-warps
#define IAC_HALO ((IAC_TH + 2) * (IAC_TW + 2))
#define IAC_NCOL 5                                   // columns of the 18-wide haloed row owned by a half-warp

// one tap (4 channels) as stored in registers: fp16 pairs in tensor-core modes, fp32 in the exact mode
template <bool HALF> struct IacTap;
template <> struct IacTap<true> {
    uint2 r;
    __device__ __forceinline__ void load(const void* taps, size_t idx) {
        r = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(taps) + idx));
    }
    __device__ __forceinline__ float4 get() const {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
};
template <> struct IacTap<false> {
    float4 r;
    __device__ __forceinline__ void load(const void* taps, size_t idx) {
        r = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(taps) + idx));
    }
    __device__ __forceinline__ float4 get() const { return r; }
};

__device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}

template <bool HALF>
__global__ void __launch_bounds__(IAC_THREADS, HALF ? 2 : 1) iac_step_kernel(IacArgs a) {
    extern __shared__ float smem[];
    float* samp = smem;                                             // [(TH+2)*(TW+2)][64]
    float* vbuf = smem + IAC_HALO * IAC_C;                          // [TH*(TW+2)][64]
    // per halo pixel: the four bilinear corners as pixel indices (clamped into the image) and their weights (zero for corners
    // or samples outside it): computed once per pixel in phase 0 instead of by each of the 16 channel lanes in phase 1
    int4* geo_i = reinterpret_cast<int4*>(vbuf + IAC_TH * (IAC_TW + 2) * IAC_C);         // [(TH+2)*(TW+2)]
    float4* geo_w = reinterpret_cast<float4*>(geo_i + IAC_HALO);
    const int hw = threadIdx.x >> 4, c0 = (threadIdx.x & 15) * 4;
    const int tiles_x = (a.W + IAC_TW - 1) / IAC_TW;
    const int ty0 = (blockIdx.x / tiles_x) * IAC_TH, tx0 = (blockIdx.x % tiles_x) * IAC_TW;
    const int b = blockIdx.y, dir = blockIdx.z;
    const float* prev = dir ? a.prev[1] : a.prev[0];
    const int ldp = dir ? a.ldprev[1] : a.ldprev[0];
    const int offs_ch = dir ? a.offs_ch[1] : a.offs_ch[0];
    const int ldxin = dir ? a.ldxin[1] : a.ldxin[0], ldnext = dir ? a.ldnext[1] : a.ldnext[0];
    const int H = a.H, W = a.W;
    const size_t img = (size_t)b * H * W;

    // phases 2/3 ownership: tile row `row`, haloed columns cg, cg+4, ... ; taps requested before anything else
    const int row = hw >> 2, cg = hw & 3;
    const int y = ty0 + row;
    IacTap<HALF> k[IAC_NCOL][3];
#pragma unroll
    for (int j = 0; j < IAC_NCOL; ++j) {
        const int hx = cg + 4 * j;
        if (hx < IAC_TW + 2 && y < H) {
            const int xx = min(max(tx0 - 1 + hx, 0), W - 1);
            const size_t idx = (img + (size_t)y * W + xx) * a.ldtaps + c0;
#pragma unroll
            for (int t = 0; t < 3; ++t) k[j][t].load(a.taps, idx + t * IAC_C);
        }
    }
    // phase 0: sample geometry of the haloed tile (clamped pixel == replicate padding), one halo pixel per thread
    for (int hp = threadIdx.x; hp < IAC_HALO; hp += IAC_THREADS) {
        const int hy = hp / (IAC_TW + 2), hx = hp - hy * (IAC_TW + 2);
        const int yy = min(max(ty0 - 1 + hy, 0), H - 1), xx = min(max(tx0 - 1 + hx, 0), W - 1);
        const float2 d = *reinterpret_cast<const float2*>(a.offs + (img + (size_t)yy * W + xx) * a.ldoffs + offs_ch);
        const float sx = (float)xx + d.x, sy = (float)yy + d.y;
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
        geo_i[hp] = gi;
        geo_w[hp] = gw;
    }
    __syncthreads();
    // phase 1: warped samples, two halo pixels (8 independent 16-byte gathers) in flight per thread
    const float* pbase = prev + img * ldp + c0;
    const unsigned short* pbase16 = reinterpret_cast<const unsigned short*>(prev) + img * ldp + c0;
    const bool p16 = a.prev16 != 0;
    for (int hp = hw; hp < IAC_HALO; hp += 2 * IAC_HW) {
        const int hp2 = min(hp + IAC_HW, IAC_HALO - 1);
        const int4 i0 = geo_i[hp], i1 = geo_i[hp2];
        const float4 w0 = geo_w[hp], w1 = geo_w[hp2];
        float4 a0, a1, a2, a3, b0, b1, b2, b3;
        if (p16) {          // 8-byte gathers of 4 bf16 channels
            const uint2 u0 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i0.x * ldp));
            const uint2 u1 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i0.y * ldp));
            const uint2 u2 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i0.z * ldp));
            const uint2 u3 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i0.w * ldp));
            const uint2 v0 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i1.x * ldp));
            const uint2 v1 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i1.y * ldp));
            const uint2 v2 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i1.z * ldp));
            const uint2 v3 = __ldg(reinterpret_cast<const uint2*>(pbase16 + (size_t)i1.w * ldp));
            a0 = bf16x4_to_float4(u0); a1 = bf16x4_to_float4(u1); a2 = bf16x4_to_float4(u2); a3 = bf16x4_to_float4(u3);
            b0 = bf16x4_to_float4(v0); b1 = bf16x4_to_float4(v1); b2 = bf16x4_to_float4(v2); b3 = bf16x4_to_float4(v3);
        } else {
            a0 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i0.x * ldp));
            a1 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i0.y * ldp));
            a2 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i0.z * ldp));
            a3 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i0.w * ldp));
            b0 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i1.x * ldp));
            b1 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i1.y * ldp));
            b2 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i1.z * ldp));
            b3 = __ldg(reinterpret_cast<const float4*>(pbase + (size_t)i1.w * ldp));
        }
        // same accumulation order as before (corner 00, 01, 10, 11 with fmaf): bit-identical samples
        float4 s0, s1;
        s0.x = fmaf(w0.w, a3.x, fmaf(w0.z, a2.x, fmaf(w0.y, a1.x, w0.x * a0.x)));
        s0.y = fmaf(w0.w, a3.y, fmaf(w0.z, a2.y, fmaf(w0.y, a1.y, w0.x * a0.y)));
        s0.z = fmaf(w0.w, a3.z, fmaf(w0.z, a2.z, fmaf(w0.y, a1.z, w0.x * a0.z)));
        s0.w = fmaf(w0.w, a3.w, fmaf(w0.z, a2.w, fmaf(w0.y, a1.w, w0.x * a0.w)));
        s1.x = fmaf(w1.w, b3.x, fmaf(w1.z, b2.x, fmaf(w1.y, b1.x, w1.x * b0.x)));
        s1.y = fmaf(w1.w, b3.y, fmaf(w1.z, b2.y, fmaf(w1.y, b1.y, w1.x * b0.y)));
        s1.z = fmaf(w1.w, b3.z, fmaf(w1.z, b2.z, fmaf(w1.y, b1.z, w1.x * b0.z)));
        s1.w = fmaf(w1.w, b3.w, fmaf(w1.z, b2.w, fmaf(w1.y, b1.w, w1.x * b0.w)));
        *reinterpret_cast<float4*>(samp + hp * IAC_C + c0) = s0;
        if (hp + IAC_HW < IAC_HALO) *reinterpret_cast<float4*>(samp + hp2 * IAC_C + c0) = s1;
    }
    __syncthreads();
    // phase 2: vertical pass for this half-warp's columns of its row
#pragma unroll
    for (int j = 0; j < IAC_NCOL; ++j) {
        const int hx = cg + 4 * j;
        if (hx < IAC_TW + 2) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < H) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float4 kk = k[j][t].get();
                    const float4 sv = *reinterpret_cast<const float4*>(samp + ((row + t) * (IAC_TW + 2) + hx) * IAC_C + c0);
                    acc.x = fmaf(kk.x, sv.x, acc.x);
                    acc.y = fmaf(kk.y, sv.y, acc.y);
                    acc.z = fmaf(kk.z, sv.z, acc.z);
                    acc.w = fmaf(kk.w, sv.w, acc.w);
                }
            }
            *reinterpret_cast<float4*>(vbuf + (row * (IAC_TW + 2) + hx) * IAC_C + c0) = acc;
        }
    }
    // residual input of the owned output pixels (haloed column hx <-> x = tx0 + hx - 1), in flight across the barrier
    const float* xin = dir ? a.xin[1] : a.xin[0];
    float4 xi[IAC_NCOL];
#pragma unroll
    for (int j = 0; j < IAC_NCOL; ++j) {
        const int hx = cg + 4 * j, x = tx0 + hx - 1;
        if (hx >= 1 && hx <= IAC_TW && x < W && y < H)
            xi[j] = __ldg(reinterpret_cast<const float4*>(xin + (img + (size_t)y * W + x) * ldxin + c0));
    }
    __syncthreads();
    // phase 3: horizontal pass + residual + LeakyReLU(0.1)
    float* next = dir ? a.next[1] : a.next[0];
#pragma unroll
    for (int j = 0; j < IAC_NCOL; ++j) {
        const int hx = cg + 4 * j, x = tx0 + hx - 1;
        if (hx >= 1 && hx <= IAC_TW && x < W && y < H) {
            float4 acc = xi[j];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const float4 kk = k[j][t].get();
                const float4 v = *reinterpret_cast<const float4*>(vbuf + (row * (IAC_TW + 2) + hx - 1 + t) * IAC_C + c0);
                acc.x = fmaf(kk.x, v.x, acc.x);
                acc.y = fmaf(kk.y, v.y, acc.y);
                acc.z = fmaf(kk.z, v.z, acc.z);
                acc.w = fmaf(kk.w, v.w, acc.w);
            }
            acc.x = acc.x >= 0.f ? acc.x : 0.1f * acc.x;
            acc.y = acc.y >= 0.f ? acc.y : 0.1f * acc.y;
            acc.z = acc.z >= 0.f ? acc.z : 0.1f * acc.z;
            acc.w = acc.w >= 0.f ? acc.w : 0.1f * acc.w;
            const size_t o = (img + (size_t)y * W + x) * ldnext + c0;
            if (a.round_out) store_operand4(next, o, acc, a.round_out == 2);
            else *reinterpret_cast<float4*>(next + o) = acc;
        }
    }
}

extern "C" int fcvsr_iac_step(const float* prev_f, int ldprev_f, const float* prev_b, int ldprev_b, const float* xin_f,
                              int ldxin_f, const float* xin_b, int ldxin_b, float* next_f, int ldnext_f, float* next_b,
                              int ldnext_b, const float* offs, int ldoffs, int ch_f, int ch_b, const float* taps,
                              int ldtaps, int taps_half, int B, int H, int W, int round_out, cudaStream_t st) {
    if (!prev_f || !prev_b || !xin_f || !xin_b || !next_f || !next_b || !offs || !taps) return FCVSR_ERR_ARG;
    if (((ldprev_f | ldprev_b | ldxin_f | ldxin_b | ldnext_f | ldnext_b | ldtaps) & 3) || ((ldoffs | ch_f | ch_b) & 1))
        return FCVSR_ERR_ARG;
    if (((uintptr_t)prev_f | (uintptr_t)prev_b | (uintptr_t)xin_f | (uintptr_t)xin_b | (uintptr_t)next_f | (uintptr_t)next_b) & 15)
        return FCVSR_ERR_ARG;
    if ((uintptr_t)taps & (taps_half ? 7 : 15)) return FCVSR_ERR_ARG;
    IacArgs a;
    a.prev[0] = prev_f; a.prev[1] = prev_b; a.ldprev[0] = ldprev_f; a.ldprev[1] = ldprev_b;
    a.xin[0] = xin_f; a.xin[1] = xin_b; a.ldxin[0] = ldxin_f; a.ldxin[1] = ldxin_b;
    a.next[0] = next_f; a.next[1] = next_b; a.ldnext[0] = ldnext_f; a.ldnext[1] = ldnext_b;
    a.offs = offs; a.ldoffs = ldoffs; a.offs_ch[0] = ch_f; a.offs_ch[1] = ch_b;
    a.taps = taps; a.ldtaps = ldtaps; a.B = B; a.H = H; a.W = W; a.round_out = round_out & 3; a.prev16 = (round_out >> 2) & 1;
    const size_t smem = (IAC_HALO + IAC_TH * (IAC_TW + 2)) * IAC_C * sizeof(float) + IAC_HALO * (sizeof(int4) + sizeof(float4));
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(iac_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(iac_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return FCVSR_ERR_CUDA;
        attr_set = true;
    }
    dim3 grid(((H + IAC_TH - 1) / IAC_TH) * ((W + IAC_TW - 1) / IAC_TW), B, 2);
    if (taps_half) iac_step_kernel<true><<<grid, IAC_THREADS, smem, st>>>(a);
    else iac_step_kernel<false><<<grid, IAC_THREADS, smem, st>>>(a);
    return fcvsr_launch_status();
}

// y[pix, 0:Cy] = operand-typed copy of x[pix, 0:C] (TF32-rounded fp32 or bf16), channels C..Cy-1 zero-filled:
// the tensor-core operand copy of a tensor that is also a full-precision residual, padded to the K chunk.
__global__ void round_copy_kernel(const float* __restrict__ x, int ldx, void* __restrict__ y, int ldy, int c4, int cy4,
                                  size_t total4, int op16) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const size_t pix = i / cy4;
    const int c = (int)(i - pix * cy4) * 4;
    const float4 v = c < c4 * 4 ? *reinterpret_cast<const float4*>(x + pix * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    store_operand4(y, pix * ldy + c, v, op16);
}

// Both operand-typed copies of a dense [npix, C] tensor in one pass: y32 = TF32-rounded fp32, y16 = bf16 (either may be NULL).
// The training step needs both of an activation (forward convolution / tcgen05 weight gradient) and of an upstream gradient
// (data-gradient convolution / weight gradient); as separate launches these copies were 2239 launches of ~3 us per step.
__global__ void round_copy_dual_kernel(const float* __restrict__ x, float* __restrict__ y32, void* __restrict__ y16, size_t total4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    if (y32) store_operand4(y32, i * 4, v, 0);
    if (y16) store_operand4(y16, i * 4, v, 1);
}

// The same for up to three tensors in one launch (the pyramid levels of a convolution's input or upstream gradient).
struct RoundDualSeg { const float* x; float* y32; void* y16; unsigned long long begin4; };
struct RoundDualArgs { RoundDualSeg seg[3]; int n; unsigned long long total4; };
__global__ void round_copy_dual_multi_kernel(const RoundDualArgs a) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total4) return;
    const int s = (a.n > 1 && i >= a.seg[1].begin4) ? ((a.n > 2 && i >= a.seg[2].begin4) ? 2 : 1) : 0;
    const float* x = s == 0 ? a.seg[0].x : (s == 1 ? a.seg[1].x : a.seg[2].x);
    float* y32 = s == 0 ? a.seg[0].y32 : (s == 1 ? a.seg[1].y32 : a.seg[2].y32);
    void* y16 = s == 0 ? a.seg[0].y16 : (s == 1 ? a.seg[1].y16 : a.seg[2].y16);
    const size_t j = (size_t)(i - (s == 0 ? 0ull : (s == 1 ? a.seg[1].begin4 : a.seg[2].begin4)));
    const float4 v = reinterpret_cast<const float4*>(x)[j];
    if (y32) store_operand4(y32, j * 4, v, 0);
    if (y16) store_operand4(y16, j * 4, v, 1);
}

// x / y_tf32 / y_bf16: HOST arrays of n <= 3 device pointers (entries of y_tf32 / y_bf16, or the arrays themselves, may be NULL);
// numel: HOST array, each % 4 == 0.
extern "C" int fcvsr_round_copy_dual_multi(int n, const float* const* x, float* const* y_tf32, void* const* y_bf16,
                                           const long long* numel, cudaStream_t st) {
    if (n < 1 || n > 3 || !x || !numel || (!y_tf32 && !y_bf16)) return FCVSR_ERR_ARG;
    RoundDualArgs a;
    a.n = n;
    unsigned long long t = 0;
    for (int i = 0; i < 3; ++i) {
        const int j = i < n ? i : 0;
        a.seg[i].x = x[j]; a.seg[i].y32 = y_tf32 ? y_tf32[j] : nullptr; a.seg[i].y16 = y_bf16 ? y_bf16[j] : nullptr;
        a.seg[i].begin4 = t;
        if (i >= n) continue;
        if (!x[j] || numel[j] <= 0 || (numel[j] & 3) || (!a.seg[i].y32 && !a.seg[i].y16)) return FCVSR_ERR_ARG;
        if (((uintptr_t)x[j] | (uintptr_t)a.seg[i].y32) & 15 || ((uintptr_t)a.seg[i].y16 & 7)) return FCVSR_ERR_ARG;
        t += (unsigned long long)numel[j] / 4;
    }
    a.total4 = t;
    round_copy_dual_multi_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(a);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_round_copy_dual(const float* x, float* y_tf32, void* y_bf16, long long numel, cudaStream_t st) {
    if (!x || (!y_tf32 && !y_bf16) || numel <= 0 || (numel & 3) || ((uintptr_t)x & 15) || ((uintptr_t)y_tf32 & 15) || ((uintptr_t)y_bf16 & 7))
        return FCVSR_ERR_ARG;
    const size_t total4 = (size_t)numel / 4;
    round_copy_dual_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(x, y_tf32, y_bf16, total4);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_round_copy(const float* x, int ldx, void* y, int ldy, int C, int Cy, long long npix, int op16,
                                cudaStream_t st) {
    if (!x || !y || (C & 3) || (Cy & 3) || Cy < C || (ldx & 3) || (ldy & 3)) return FCVSR_ERR_ARG;
    const size_t total4 = (size_t)npix * (Cy / 4);
    round_copy_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(x, ldx, y, ldy, C / 4, Cy / 4, total4, op16);
    return fcvsr_launch_status();
}

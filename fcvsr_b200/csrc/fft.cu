// Batched mixed-radix FFT kernels for NHWC tensors (shared-memory Stockham autosort).
//
// Replaces the torch.fft calls of the reference hot path:
//   torch.fft.rfft2(x, norm='backward')        CVSR_freq.py:1452-1454   (MGAAbk)
//   torch.fft.irfft2(z, s=(H,W))               CVSR_freq.py:1499,1504   (offset maps)
//   fft.fftn / fftshift / *mask / ifftn(.real) CVSR_freq.py:2082-2088   (Split_freq)
//
// A 2-D transform is two passes of 1-D "line" kernels.  Because channels are the contiguous
// dimension (NHWC), a block transforms CB channels of one line at once: every global access is a
// CB*8-byte contiguous segment and every shared-memory access is conflict-free, for the W pass
// and the H pass alike.  Real transforms pack two real channels into one complex FFT
// (z = a + i b), which halves the work and makes the real tensor's (c, c+1) float2 the complex
// sample directly.
//
// Sizes: any N = 2^a 3^b 5^c 7^d 11^e 13^f 17^g (covers 64, 180, 320, 272 = 16*17, 480, 540, 960).
// HBM-bound by design: each pass reads and writes its tensor exactly once.
#include "common.cuh"
#include "fft_reg.cuh"
#include <stdlib.h>

#define FFT_MAX_PASSES 16
#define FFT_THREADS 256

struct FftPlan {
    int n;
    int npass;
    int radix[FFT_MAX_PASSES];
};

static bool make_plan(int n, FftPlan* p) {
    static const int cand[] = {4, 2, 3, 5, 7, 11, 13, 17};
    p->n = n;
    p->npass = 0;
    int rem = n;
    for (int r : cand) {
        while (rem % r == 0) {
            if (p->npass >= FFT_MAX_PASSES) return false;
            p->radix[p->npass++] = r;
            rem /= r;
        }
    }
    // any other prime factor (19, 23, ...): a generic radix pass, O(R) per output element -- torch.fft accepts every length and
    // so does this operator; such lengths are rare (H, W are multiples of 4) and short of the tuned radices only in speed
    for (int r = 19; rem > 1; r += 2) {
        if (r * r > rem) r = rem;
        while (rem % r == 0) {
            if (p->npass >= FFT_MAX_PASSES) return false;
            p->radix[p->npass++] = r;
            rem /= r;
        }
    }
    return rem == 1 && n >= 2;
}

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    unsigned w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    return w;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// One Stockham pass of radix R over a line of length n held as [n][cb] float2 in shared memory.
//   a_k = in[q + s*(p + m*k)],  out[q + s*(R*p + j)] = w_n'^(p*j) * sum_k a_k w_R^(jk)
// with n' = n/s the current sub-transform length, m = n'/R, twiddles taken from the length-n table
// (tw[t] = exp(-+2 pi i t / n), already conjugated for the inverse).
template <int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out, int n, int cb_log2,
                                         int s, const float2* __restrict__ tw, bool inverse) {
    const int nb = n / R;
    const int m = nb / s;
    const int cb = 1 << cb_log2;
    float2 wr[R];
#pragma unroll
    for (int t = 0; t < R; ++t) wr[t] = tw[t * nb];
    const float inv_s = 1.0f / (float)s;        // bf < 2^20: (bf + 0.5) * (1/s) truncates to the exact quotient
    for (int idx = threadIdx.x; idx < (nb << cb_log2); idx += blockDim.x) {
        const int bf = idx >> cb_log2, ch = idx & (cb - 1);
        const int p = (int)(((float)bf + 0.5f) * inv_s), q = bf - p * s;
        float2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = in[((q + s * (p + m * k)) << cb_log2) + ch];
        float2 o[R];
        if (R == 2) {
            o[0] = cadd(v[0], v[1]);
            o[1] = csub(v[0], v[1]);
        } else if (R == 4) {
            float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
            float2 t2 = cadd(v[1], v[3]), t3 = csub(v[1], v[3]);
            // multiply t3 by -i (forward) or +i (inverse)
            float2 t3r = inverse ? make_float2(-t3.y, t3.x) : make_float2(t3.y, -t3.x);
            o[0] = cadd(t0, t2);
            o[2] = csub(t0, t2);
            o[1] = cadd(t1, t3r);
            o[3] = csub(t1, t3r);
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                float2 acc = v[0];
#pragma unroll
                for (int k = 1; k < R; ++k) acc = cadd(acc, cmul(v[k], wr[(j * k) % R]));
                o[j] = acc;
            }
        }
        const int obase = q + s * R * p;
        out[(obase << cb_log2) + ch] = o[0];
#pragma unroll
        for (int j = 1; j < R; ++j) out[((obase + s * j) << cb_log2) + ch] = cmul(o[j], tw[p * j * s]);
    }
}

// The same pass for a run-time radix R (prime factors above 17): one output element per thread, the R inputs read from shared memory.
__device__ __forceinline__ void fft_pass_generic(const float2* __restrict__ in, float2* __restrict__ out, int n, int cb_log2, int s,
                                                 const float2* __restrict__ tw, int R) {
    const int nb = n / R, m = nb / s;
    const int cb = 1 << cb_log2;
    for (int idx = threadIdx.x; idx < (n << cb_log2); idx += blockDim.x) {
        const int e = idx >> cb_log2, ch = idx & (cb - 1);
        const int bf = e / R, j = e - bf * R;
        const int p = bf / s, q = bf - p * s;
        float2 acc = in[((q + s * p) << cb_log2) + ch];
        int jk = 0;
        for (int k = 1; k < R; ++k) {
            jk += j;
            if (jk >= R) jk -= R;
            acc = cadd(acc, cmul(in[((q + s * (p + m * k)) << cb_log2) + ch], tw[jk * nb]));
        }
        out[((q + s * R * p + s * j) << cb_log2) + ch] = j ? cmul(acc, tw[p * j * s]) : acc;
    }
}

// Runs all passes; returns the buffer that holds the result.
__device__ float2* fft_line(float2* a, float2* b, const FftPlan& plan, int cb_log2, const float2* tw, bool inverse) {
    int s = 1;
    for (int i = 0; i < plan.npass; ++i) {
        const int r = plan.radix[i];
        switch (r) {
            case 2: fft_pass<2>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 3: fft_pass<3>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 4: fft_pass<4>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 5: fft_pass<5>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 7: fft_pass<7>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 11: fft_pass<11>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 13: fft_pass<13>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            case 17: fft_pass<17>(a, b, plan.n, cb_log2, s, tw, inverse); break;
            default: fft_pass_generic(a, b, plan.n, cb_log2, s, tw, r); break;
        }
        __syncthreads();
        float2* t = a;
        a = b;
        b = t;
        s *= r;
    }
    return a;
}

__device__ __forceinline__ void load_twiddles(float2* tws, const float2* __restrict__ tw, int n, bool inverse) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float2 t = tw[i];
        tws[i] = inverse ? make_float2(t.x, -t.y) : t;
    }
}

// ---- real -> complex along W --------------------------------------------------------------------
// x real [B,H,W,ldx] (C real channels from x), out complex [B,H,Wf,C].  grid (B*H, C/2/CB).
__global__ void __launch_bounds__(FFT_THREADS) fft_r2c_w_kernel(const float* __restrict__ x, int ldx,
                                                                float2* __restrict__ out, const float2* __restrict__ tw,
                                                                int W, int C, int cb_log2, FftPlan plan) {
    extern __shared__ float2 sm[];
    const int cb = 1 << cb_log2, n = W, wf = W / 2 + 1;
    float2* a = sm;
    float2* b = sm + (size_t)n * cb;
    float2* tws = sm + 2 * (size_t)n * cb;
    const size_t line = blockIdx.x;
    const int c0 = blockIdx.y * cb;          // first complex lane == real channel pair index
    load_twiddles(tws, tw, n, false);
    const float* src = x + line * (size_t)W * ldx + 2 * c0;
    // batches of 8 independent loads per thread: the line must be in flight as a whole, not one element at a time
    for (int idx0 = threadIdx.x; idx0 < (n << cb_log2); idx0 += 8 * FFT_THREADS) {
        float2 t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * FFT_THREADS;
            const int i = idx >> cb_log2, ch = idx & (cb - 1);
            if (idx < (n << cb_log2)) t[u] = *reinterpret_cast<const float2*>(src + (size_t)i * ldx + 2 * ch);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * FFT_THREADS;
            if (idx < (n << cb_log2)) a[idx] = t[u];
        }
    }
    __syncthreads();
    const float2* r = fft_line(a, b, plan, cb_log2, tws, false);
    float2* dst = out + line * (size_t)wf * C + 2 * c0;
    for (int idx = threadIdx.x; idx < (wf << cb_log2); idx += blockDim.x) {
        const int k = idx >> cb_log2, ch = idx & (cb - 1);
        const float2 zk = r[idx];
        const int kn = k == 0 ? 0 : n - k;
        float2 zn = r[(kn << cb_log2) + ch];
        zn.y = -zn.y;
        const float2 fa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
        const float2 d = csub(zk, zn);
        const float2 fb = make_float2(0.5f * d.y, -0.5f * d.x);
        *reinterpret_cast<float4*>(dst + (size_t)k * C + 2 * ch) = make_float4(fa.x, fa.y, fb.x, fb.y);
    }
}

// ---- complex -> complex along H ------------------------------------------------------------------
// in/out complex [B,H,Wf,C]; optional real mask [H*Wf] multiplied at load; grid (B*Wf, C/CB).
__global__ void __launch_bounds__(FFT_THREADS) fft_c2c_h_kernel(const float2* in, float2* out,
                                                                const float2* __restrict__ tw,
                                                                const float* __restrict__ mask, int H, int Wf, int C,
                                                                int cb_log2, int inverse, float scale, int round_out, FftPlan plan, unsigned* out_bf16) {
    extern __shared__ float2 sm[];
    const int cb = 1 << cb_log2, n = H;
    float2* a = sm;
    float2* b = sm + (size_t)n * cb;
    float2* tws = sm + 2 * (size_t)n * cb;
    const int bidx = blockIdx.x / Wf, wf = blockIdx.x - bidx * Wf;
    const int c0 = blockIdx.y * cb;
    load_twiddles(tws, tw, n, inverse != 0);
    const size_t base = ((size_t)bidx * H * Wf + wf) * C + c0;
    if (mask) mask += (size_t)blockIdx.z * H * Wf;                     // replica z: its own mask ...
    out += (size_t)blockIdx.z * gridDim.x * H * C;                     // ... and output ([nrep][B,H,Wf,C])
    for (int idx0 = threadIdx.x; idx0 < (n << cb_log2); idx0 += 8 * FFT_THREADS) {
        float2 t[8];
        float mk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * FFT_THREADS;
            const int i = idx >> cb_log2, ch = idx & (cb - 1);
            mk[u] = 1.f;
            if (idx < (n << cb_log2)) {
                t[u] = in[base + (size_t)i * Wf * C + ch];
                if (mask) mk[u] = mask[i * Wf + wf];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * FFT_THREADS;
            if (idx < (n << cb_log2)) a[idx] = make_float2(t[u].x * mk[u], t[u].y * mk[u]);
        }
    }
    __syncthreads();
    const float2* r = fft_line(a, b, plan, cb_log2, tws, inverse != 0);
    for (int idx = threadIdx.x; idx < (n << cb_log2); idx += blockDim.x) {
        const int i = idx >> cb_log2, ch = idx & (cb - 1);
        float2 v = r[idx];
        v = make_float2(v.x * scale, v.y * scale);
        if (round_out) v = make_float2(round_tf32(v.x), round_tf32(v.y));
        out[base + (size_t)i * Wf * C + ch] = v;
        if (out_bf16) out_bf16[base + (size_t)i * Wf * C + ch] = pack_bf16x2(v.x, v.y);
    }
}

// ---- complex -> real along W (torch c2r semantics: Im of the DC and Nyquist bins is ignored) -----
// in complex [B,H,Wf,C], y real [B,H,W,ldy] (C real channels); grid (B*H, C/2/CB).
__global__ void __launch_bounds__(FFT_THREADS) fft_c2r_w_kernel(const float2* __restrict__ in, float* __restrict__ y,
                                                                int ldy, const float2* __restrict__ tw, int W, int C,
                                                                int cb_log2, float scale, FftPlan plan) {
    extern __shared__ float2 sm[];
    const int cb = 1 << cb_log2, n = W, wf = W / 2 + 1;
    float2* a = sm;
    float2* b = sm + (size_t)n * cb;
    float2* tws = sm + 2 * (size_t)n * cb;
    const size_t line = blockIdx.x;
    const int c0 = blockIdx.y * cb;
    load_twiddles(tws, tw, n, true);
    const float2* src = in + line * (size_t)wf * C + 2 * c0;
    for (int idx0 = threadIdx.x; idx0 < (wf << cb_log2); idx0 += 4 * FFT_THREADS) {
        float4 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * FFT_THREADS;
            const int k = idx >> cb_log2, ch = idx & (cb - 1);
            if (idx < (wf << cb_log2)) t[u] = *reinterpret_cast<const float4*>(src + (size_t)k * C + 2 * ch);   // A = (x,y), B = (z,w)
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * FFT_THREADS;
            if (idx >= (wf << cb_log2)) continue;
            const int k = idx >> cb_log2, ch = idx & (cb - 1);
            float4 ab = t[u];
            if (k == 0 || 2 * k == n) {
                ab.y = 0.f;
                ab.w = 0.f;
            }
            a[idx] = make_float2(ab.x - ab.w, ab.y + ab.z);                 // A + iB
            if (k != 0 && 2 * k != n) a[((n - k) << cb_log2) + ch] = make_float2(ab.x + ab.w, ab.z - ab.y);  // conj(A) + i conj(B)
        }
    }
    __syncthreads();
    const float2* r = fft_line(a, b, plan, cb_log2, tws, true);
    float* dst = y + line * (size_t)W * ldy + 2 * c0;
    for (int idx = threadIdx.x; idx < (n << cb_log2); idx += blockDim.x) {
        const int i = idx >> cb_log2, ch = idx & (cb - 1);
        const float2 v = r[idx];
        *reinterpret_cast<float2*>(dst + (size_t)i * ldy + 2 * ch) = make_float2(v.x * scale, v.y * scale);
    }
}

// =================================================================================================
// Two-phase register-resident kernels (fft_reg.cuh) for lengths N = R1 * R2 with both radices built.
// Same tensor layouts, arguments and semantics as the Stockham kernels above, which remain the path for
// every other length.
// =================================================================================================
// Register budget: the kernels are half issue-bound, half latency-bound (ncu: issue slots ~50 %, occupancy 23 % at 128
// registers).  Capping at 80 (W passes, 3 blocks of <= 256 threads) / 64 (H pass, 4 blocks) registers costs a few spills in
// the radix-20+ butterflies but measured +5-12 % bandwidth.
#define FFT2_THREADS_MAX 256

#define FFT2_CASE_A(R) case R: fftreg::phase_a<R, INV>(load, S, tws, r2, cb_log2); break;
#define FFT2_CASE_B(R) case R: fftreg::phase_b<R, INV>(S, r1, cb_log2, store); break;

template <bool INV>
__global__ void __launch_bounds__(FFT2_THREADS_MAX, 4) fft2_c2c_h_kernel(const float2* in, float2* out, const float2* __restrict__ tw,
                                                                     const float* __restrict__ mask, int H, int Wf, int C,
                                                                     int cb_log2, float scale, int round_out, int r1, int r2, unsigned* out_bf16) {
    extern __shared__ float2 sm[];
    const int cb = 1 << cb_log2, n = H;
    float2* S = sm;
    float2* tws = sm + (size_t)n * cb;
    const int bidx = blockIdx.x / Wf, wf = blockIdx.x - bidx * Wf;
    const int c0 = blockIdx.y * cb;
    load_twiddles(tws, tw, n, INV);
    const size_t base = ((size_t)bidx * H * Wf + wf) * C + c0;
    const size_t rs = (size_t)Wf * C;
    if (mask) mask += (size_t)blockIdx.z * H * Wf + wf;                // replica z: its own mask ...
    out += (size_t)blockIdx.z * gridDim.x * H * C;                     // ... and output ([nrep][B,H,Wf,C])
    __syncthreads();
    {
        auto load = [&](int i, int ch) {
            float2 t = in[base + (size_t)i * rs + ch];
            if (mask) { const float m = __ldg(mask + i * Wf); t.x *= m; t.y *= m; }
            return t;
        };
        switch (r1) { FFT2_FOR_EACH_RADIX(FFT2_CASE_A) }
    }
    __syncthreads();
    {
        auto store = [&](int k, int ch, float2 v) {
            v = make_float2(v.x * scale, v.y * scale);
            if (round_out) v = make_float2(round_tf32(v.x), round_tf32(v.y));
            out[base + (size_t)k * rs + ch] = v;
            if (out_bf16) out_bf16[base + (size_t)k * rs + ch] = pack_bf16x2(v.x, v.y);   // operand copy (one complex = one word)
        };
        switch (r2) { FFT2_FOR_EACH_RADIX(FFT2_CASE_B) }
    }
}

__global__ void __launch_bounds__(FFT2_THREADS_MAX, 3) fft2_r2c_w_kernel(const float* __restrict__ x, int ldx, float2* __restrict__ out,
                                                                     const float2* __restrict__ tw, int W, int C, int cb_log2,
                                                                     int r1, int r2) {
    constexpr bool INV = false;
    extern __shared__ float2 sm[];
    const int cb = 1 << cb_log2, n = W, wf = W / 2 + 1;
    float2* S = sm;
    float2* Z = sm + (size_t)n * cb;
    float2* tws = sm + 2 * (size_t)n * cb;
    const size_t line = blockIdx.x;
    const int c0 = blockIdx.y * cb;          // first complex lane == real channel pair index
    load_twiddles(tws, tw, n, false);
    const float* src = x + line * (size_t)W * ldx + 2 * c0;
    __syncthreads();
    {
        auto load = [&](int i, int ch) { return __ldg(reinterpret_cast<const float2*>(src + (size_t)i * ldx + 2 * ch)); };
        switch (r1) { FFT2_FOR_EACH_RADIX(FFT2_CASE_A) }
    }
    __syncthreads();
    {
        auto store = [&](int k, int ch, float2 v) { Z[(k << cb_log2) + ch] = v; };
        switch (r2) { FFT2_FOR_EACH_RADIX(FFT2_CASE_B) }
    }
    __syncthreads();
    float2* dst = out + line * (size_t)wf * C + 2 * c0;
    for (int idx = threadIdx.x; idx < (wf << cb_log2); idx += blockDim.x) {
        const int k = idx >> cb_log2, ch = idx & (cb - 1);
        const float2 zk = Z[idx];
        const int kn = k == 0 ? 0 : n - k;
        float2 zn = Z[(kn << cb_log2) + ch];
        zn.y = -zn.y;
        const float2 fa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
        const float2 d = csub(zk, zn);
        const float2 fb = make_float2(0.5f * d.y, -0.5f * d.x);
        *reinterpret_cast<float4*>(dst + (size_t)k * C + 2 * ch) = make_float4(fa.x, fa.y, fb.x, fb.y);
    }
}

__global__ void __launch_bounds__(FFT2_THREADS_MAX, 3) fft2_c2r_w_kernel(const float2* __restrict__ in, float* __restrict__ y, int ldy,
                                                                     const float2* __restrict__ tw, int W, int C, int cb_log2,
                                                                     float scale, int r1, int r2) {
    constexpr bool INV = true;
    extern __shared__ float2 sm[];
    const int cb = 1 << cb_log2, n = W, wf = W / 2 + 1;
    float2* S = sm;
    float2* P = sm + (size_t)n * cb;
    float2* tws = sm + 2 * (size_t)n * cb;
    const size_t line = blockIdx.x;
    const int c0 = blockIdx.y * cb;
    load_twiddles(tws, tw, n, true);
    const float2* src = in + line * (size_t)wf * C + 2 * c0;
    // packed spectrum Z[k] = A_k + i B_k, Z[n-k] = conj(A_k) + i conj(B_k) staged once in shared memory (fetching each
    // (A, B) pair from two threads instead measured 30 % slower); torch c2r semantics: Im of DC / Nyquist ignored
    for (int idx0 = threadIdx.x; idx0 < (wf << cb_log2); idx0 += 4 * blockDim.x) {
        float4 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            const int k = idx >> cb_log2, ch = idx & (cb - 1);
            if (idx < (wf << cb_log2)) t[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * C + 2 * ch));   // A = (x,y), B = (z,w)
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            if (idx >= (wf << cb_log2)) continue;
            const int k = idx >> cb_log2, ch = idx & (cb - 1);
            float4 ab = t[u];
            if (k == 0 || 2 * k == n) {
                ab.y = 0.f;
                ab.w = 0.f;
            }
            P[idx] = make_float2(ab.x - ab.w, ab.y + ab.z);
            if (k != 0 && 2 * k != n) P[((n - k) << cb_log2) + ch] = make_float2(ab.x + ab.w, ab.z - ab.y);
        }
    }
    __syncthreads();
    {
        auto load = [&](int i, int ch) { return P[(i << cb_log2) + ch]; };
        switch (r1) { FFT2_FOR_EACH_RADIX(FFT2_CASE_A) }
    }
    __syncthreads();
    float* dst = y + line * (size_t)W * ldy + 2 * c0;
    {
        auto store = [&](int k, int ch, float2 v) {
            *reinterpret_cast<float2*>(dst + (size_t)k * ldy + 2 * ch) = make_float2(v.x * scale, v.y * scale);
        };
        switch (r2) { FFT2_FOR_EACH_RADIX(FFT2_CASE_B) }
    }
}

// N = r1 * r2 with both radices built (r1 >= r2, most balanced pair); false -> use the Stockham kernels
static bool make_plan2(int n, int* r1, int* r2) {
#ifdef FCVSR_BRINGUP
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("FCVSR_FFT_LEGACY"); legacy = e ? atoi(e) : 0; }
    if (legacy) return false;
#endif
#define FFT2_LIST(R) R,
    static const int rad[] = {FFT2_FOR_EACH_RADIX(FFT2_LIST)};
#undef FFT2_LIST
    int best = 0;
    for (int a : rad)
        for (int b : rad)
            if (a * b == n && a >= b && (best == 0 || a < best)) { best = a; *r1 = a; *r2 = b; }
    return best != 0;
}
static int fft2_threads(int r1, int r2, int cbl) {
    int items = (r1 > r2 ? r1 : r2) << cbl;
    int t = (items + 31) & ~31;
    return t > FFT2_THREADS_MAX ? FFT2_THREADS_MAX : t;
}

// ------------------------------------------------------------------------------------------------
static int pick_cb_log2(int n, int lanes, int two_phase = 0) {
    // largest power of two CB <= 16 that divides `lanes` and keeps 2*n*CB*8 + n*8 <= ~200 KB
    int cap = 4;
#ifdef FCVSR_BRINGUP
    { static int cap_env = -1; if (cap_env < 0) { const char* e = getenv("FCVSR_FFT_CBL"); cap_env = e ? atoi(e) : 4; } cap = cap_env; }
#endif
    // two-phase kernels: 8 lanes per block (64-byte segments) keep 3-4 blocks of ~100-register threads resident per SM
    int cbl = two_phase && cap > 3 ? 3 : cap;
    while (cbl > 0 && ((lanes % (1 << cbl)) != 0 || (size_t)(2 * (size_t)n * (1 << cbl) + n) * 8 > 200 * 1024)) --cbl;
    return cbl;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return FCVSR_ERR_CUDA;
    }
    return FCVSR_OK;
}

extern "C" int fcvsr_fft_r2c_w(const float* x, int ldx, float* out, const float* tw, int B, int H, int W, int C,
                               cudaStream_t st) {
    FftPlan plan;
    if (!x || !out || !tw || (C & 1) || (W & 1) || (ldx & 1) || !make_plan(W, &plan)) return FCVSR_ERR_ARG;
    int r1, r2;
    const bool two = make_plan2(W, &r1, &r2);
    const int lanes = C / 2, cbl = pick_cb_log2(W, lanes, two);
    const size_t smem = (2 * (size_t)W * (1 << cbl) + W) * sizeof(float2);
    dim3 grid(B * H, lanes >> cbl);
    if (two) {
        if (set_smem(fft2_r2c_w_kernel, smem)) return FCVSR_ERR_CUDA;
        fft2_r2c_w_kernel<<<grid, fft2_threads(r1, r2, cbl), smem, st>>>(x, ldx, (float2*)out, (const float2*)tw, W, C, cbl, r1, r2);
        return fcvsr_launch_status();
    }
    if (set_smem(fft_r2c_w_kernel, smem)) return FCVSR_ERR_CUDA;
    fft_r2c_w_kernel<<<grid, FFT_THREADS, smem, st>>>(x, ldx, (float2*)out, (const float2*)tw, W, C, cbl, plan);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_fft_c2c_h(const float* in, float* out, const float* tw, const float* mask, int B, int H, int Wf,
                               int C, int inverse, float scale, int round_out, int nrep, void* out_bf16, cudaStream_t st) {
    FftPlan plan;
    if (!in || !out || !tw || nrep < 1 || (nrep > 1 && (in == out || out_bf16)) || !make_plan(H, &plan)) return FCVSR_ERR_ARG;
    int r1, r2;
    const bool two = make_plan2(H, &r1, &r2);
    const int cbl = pick_cb_log2(H, C, two);
    const size_t smem = (2 * (size_t)H * (1 << cbl) + H) * sizeof(float2);
    dim3 grid(B * Wf, C >> cbl, nrep);
    if (two) {
        const size_t smem2 = ((size_t)H * (1 << cbl) + H) * sizeof(float2);
        if (set_smem(fft2_c2c_h_kernel<false>, smem2) || set_smem(fft2_c2c_h_kernel<true>, smem2)) return FCVSR_ERR_CUDA;
        const int nt = fft2_threads(r1, r2, cbl);
        if (inverse)
            fft2_c2c_h_kernel<true><<<grid, nt, smem2, st>>>((const float2*)in, (float2*)out, (const float2*)tw, mask, H, Wf, C, cbl,
                                                             scale, round_out, r1, r2, (unsigned*)out_bf16);
        else
            fft2_c2c_h_kernel<false><<<grid, nt, smem2, st>>>((const float2*)in, (float2*)out, (const float2*)tw, mask, H, Wf, C, cbl,
                                                              scale, round_out, r1, r2, (unsigned*)out_bf16);
        return fcvsr_launch_status();
    }
    if (set_smem(fft_c2c_h_kernel, smem)) return FCVSR_ERR_CUDA;
    fft_c2c_h_kernel<<<grid, FFT_THREADS, smem, st>>>((const float2*)in, (float2*)out, (const float2*)tw, mask, H, Wf,
                                                      C, cbl, inverse, scale, round_out, plan, (unsigned*)out_bf16);
    return fcvsr_launch_status();
}

extern "C" int fcvsr_fft_c2r_w(const float* in, float* y, int ldy, const float* tw, int B, int H, int W, int C,
                               float scale, cudaStream_t st) {
    FftPlan plan;
    if (!in || !y || !tw || (C & 1) || (W & 1) || (ldy & 1) || !make_plan(W, &plan)) return FCVSR_ERR_ARG;
    int r1, r2;
    const bool two = make_plan2(W, &r1, &r2);
    const int lanes = C / 2, cbl = pick_cb_log2(W, lanes, two);
    const size_t smem = (2 * (size_t)W * (1 << cbl) + W) * sizeof(float2);
    dim3 grid(B * H, lanes >> cbl);
    if (two) {
        if (set_smem(fft2_c2r_w_kernel, smem)) return FCVSR_ERR_CUDA;
        fft2_c2r_w_kernel<<<grid, fft2_threads(r1, r2, cbl), smem, st>>>((const float2*)in, y, ldy, (const float2*)tw, W, C, cbl, scale, r1, r2);
        return fcvsr_launch_status();
    }
    if (set_smem(fft_c2r_w_kernel, smem)) return FCVSR_ERR_CUDA;
    fft_c2r_w_kernel<<<grid, FFT_THREADS, smem, st>>>((const float2*)in, y, ldy, (const float2*)tw, W, C, cbl, scale,
                                                      plan);
    return fcvsr_launch_status();
}

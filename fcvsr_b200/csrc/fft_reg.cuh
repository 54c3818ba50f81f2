// Register-resident DFTs of compile-time length (2..32) and the two block-level phases built from them.
//
// A line of length N = R1 * R2 is transformed in two phases with ONE shared-memory exchange:
//   phase A  thread (n2, ch): x[R2*n1 + n2], n1 = 0..R1-1  -> R1-point DFT in registers -> * w_N^(n2*k1) -> S[k1][n2]
//   phase B  thread (k1, ch): S[k1][n2], n2 = 0..R2-1      -> R2-point DFT in registers -> X[k1 + R1*k2]
// All butterfly indices and inner twiddles are compile-time constants (constexpr trigonometry), so the per-element
// cost is ~40 instructions instead of the ~200 of the pass-per-radix Stockham kernels, and a thread has R1 independent
// global loads in flight.  Forward sign exp(-2 pi i jk/R); INV conjugates every twiddle.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace fftreg {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double sin_small(double x) {       // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) { term *= -x2 / ((2.0 * i) * (2.0 * i + 1.0)); sum += term; }
    return sum;
}
constexpr double cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) { term *= -x2 / ((2.0 * i - 1.0) * (2.0 * i)); sum += term; }
    return sum;
}
// cos / sin of 2 pi k / n with an exact octant reduction in integers
constexpr double oct_cos(int o) { return o == 0 ? 1.0 : o == 1 ? 0.70710678118654752440 : o == 2 ? 0.0 : o == 3 ? -0.70710678118654752440 : o == 4 ? -1.0 : o == 5 ? -0.70710678118654752440 : o == 6 ? 0.0 : 0.70710678118654752440; }
constexpr double oct_sin(int o) { return oct_cos((o + 6) & 7); }
constexpr double cos2pi(int k, int n) {
    k = ((k % n) + n) % n;
    const int o = (8 * k) / n, rem = (8 * k) % n;
    const double th = 2.0 * kPi * (double)rem / (8.0 * (double)n);
    return oct_cos(o) * cos_small(th) - oct_sin(o) * sin_small(th);
}
constexpr double sin2pi(int k, int n) {
    k = ((k % n) + n) % n;
    const int o = (8 * k) / n, rem = (8 * k) % n;
    const double th = 2.0 * kPi * (double)rem / (8.0 * (double)n);
    return oct_sin(o) * cos_small(th) + oct_cos(o) * sin_small(th);
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) / +i (inverse)
template <bool INV> __device__ __forceinline__ float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// a * exp(-+ 2 pi i K / N) with the constant folded at compile time
template <int K, int N, bool INV>
__device__ __forceinline__ float2 mul_tw(float2 a) {
    constexpr float c = (float)cos2pi(K, N);
    constexpr float s = (float)(INV ? sin2pi(K, N) : -sin2pi(K, N));
    if constexpr ((K % N) == 0) return a;
    else if constexpr ((4 * K) % N == 0 && ((4 * K) / N) % 4 == 1) return mul_mi<INV>(a);
    else if constexpr ((2 * K) % N == 0 && ((2 * K) / N) % 2 == 1) return make_float2(-a.x, -a.y);
    else return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
}

constexpr int first_factor(int r) { return r % 4 == 0 ? 4 : r % 2 == 0 ? 2 : r % 3 == 0 ? 3 : r % 5 == 0 ? 5 : r % 7 == 0 ? 7 : r; }

template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<1, INV> { __device__ __forceinline__ static void run(float2 (&)[1]) {} };
template <bool INV> struct Dft<2, INV> {
    __device__ __forceinline__ static void run(float2 (&v)[2]) {
        const float2 a = v[0], b = v[1];
        v[0] = caddf(a, b); v[1] = csubf(a, b);
    }
};
template <bool INV> struct Dft<3, INV> {
    __device__ __forceinline__ static void run(float2 (&v)[3]) {
        const float2 t = caddf(v[1], v[2]);
        const float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
        const float2 d = csubf(v[1], v[2]);
        const float2 s = mul_mi<INV>(make_float2(0.86602540378443864676f * d.x, 0.86602540378443864676f * d.y));
        v[0] = caddf(v[0], t); v[1] = caddf(m, s); v[2] = csubf(m, s);
    }
};
template <bool INV> struct Dft<4, INV> {
    __device__ __forceinline__ static void run(float2 (&v)[4]) {
        const float2 t0 = caddf(v[0], v[2]), t1 = csubf(v[0], v[2]);
        const float2 t2 = caddf(v[1], v[3]), t3 = mul_mi<INV>(csubf(v[1], v[3]));
        v[0] = caddf(t0, t2); v[2] = csubf(t0, t2); v[1] = caddf(t1, t3); v[3] = csubf(t1, t3);
    }
};
template <bool INV> struct Dft<5, INV> {
    __device__ __forceinline__ static void run(float2 (&v)[5]) {
        constexpr float c1 = (float)cos2pi(1, 5), c2 = (float)cos2pi(2, 5), s1 = (float)sin2pi(1, 5), s2 = (float)sin2pi(2, 5);
        const float2 t1 = caddf(v[1], v[4]), t2 = caddf(v[2], v[3]), t3 = csubf(v[1], v[4]), t4 = csubf(v[2], v[3]);
        const float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        const float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        const float2 n1 = mul_mi<INV>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        const float2 n2 = mul_mi<INV>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
        v[1] = caddf(m1, n1); v[4] = csubf(m1, n1); v[2] = caddf(m2, n2); v[3] = csubf(m2, n2);
    }
};

// odd prime R: X_j = v0 + sum_k cos(jk) p_k  -+ i sum_k sin(jk) q_k with p_k = v_k + v_{R-k}, q_k = v_k - v_{R-k}
template <int R, bool INV>
__device__ __forceinline__ void dft_prime(float2 (&v)[R]) {
    constexpr int Hh = (R - 1) / 2;
    float2 p[Hh], q[Hh];
    static_for<0, Hh>([&](auto kc) { constexpr int k = kc + 1; p[kc] = caddf(v[k], v[R - k]); q[kc] = csubf(v[k], v[R - k]); });
    float2 x0 = v[0];
    static_for<0, Hh>([&](auto kc) { x0 = caddf(x0, p[kc]); });
    const float2 v0 = v[0];
    static_for<0, Hh>([&](auto jc) {
        constexpr int j = jc + 1;
        float2 m = v0, n = make_float2(0.f, 0.f);
        static_for<0, Hh>([&](auto kc) {
            constexpr int k = kc + 1;
            constexpr float c = (float)cos2pi(j * k, R), s = (float)sin2pi(j * k, R);
            m.x = fmaf(c, p[kc].x, m.x); m.y = fmaf(c, p[kc].y, m.y);
            n.x = fmaf(s, q[kc].x, n.x); n.y = fmaf(s, q[kc].y, n.y);
        });
        const float2 ni = mul_mi<INV>(n);
        v[j] = caddf(m, ni); v[R - j] = csubf(m, ni);
    });
    v[0] = x0;
}

template <int R, bool INV> struct Dft {
    static constexpr int A = first_factor(R), B = R / A;
    __device__ __forceinline__ static void run(float2 (&v)[R]) {
        if constexpr (B == 1) {
            dft_prime<R, INV>(v);
        } else {
            float2 y[R];                                     // y[n2 * A + k1]
            static_for<0, B>([&](auto n2c) {
                constexpr int n2 = n2c;
                float2 t[A];
                static_for<0, A>([&](auto n1c) { t[n1c] = v[B * n1c + n2]; });
                Dft<A, INV>::run(t);
                static_for<0, A>([&](auto k1c) { constexpr int k1 = k1c; y[n2 * A + k1] = mul_tw<n2 * k1, R, INV>(t[k1]); });
            });
            static_for<0, A>([&](auto k1c) {
                constexpr int k1 = k1c;
                float2 t[B];
                static_for<0, B>([&](auto n2c) { t[n2c] = y[n2c * A + k1]; });
                Dft<B, INV>::run(t);
                static_for<0, B>([&](auto k2c) { v[k1 + A * k2c] = t[k2c]; });
            });
        }
    }
};

// ---- block-level phases ---------------------------------------------------------------------------
// load(n, ch) -> float2 sample n of lane ch;  S: [N][cb] float2;  tws: length-N twiddles (already conjugated for INV)
template <int R1, bool INV, class Load>
__device__ __forceinline__ void phase_a(Load load, float2* __restrict__ S, const float2* __restrict__ tws, int r2, int cb_log2) {
    const int items = r2 << cb_log2;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int n2 = it >> cb_log2, ch = it & ((1 << cb_log2) - 1);
        float2 v[R1];
        static_for<0, R1>([&](auto n1) { v[n1] = load(r2 * n1 + n2, ch); });
        Dft<R1, INV>::run(v);
        static_for<0, R1>([&](auto k1c) {
            constexpr int k1 = k1c;
            float2 o = v[k1];
            if (k1) o = cmulf(o, tws[n2 * k1]);
            S[((k1 * r2 + n2) << cb_log2) + ch] = o;
        });
    }
}
// store(k, ch, value): output bin k of lane ch
template <int R2, bool INV, class Store>
__device__ __forceinline__ void phase_b(const float2* __restrict__ S, int r1, int cb_log2, Store store) {
    const int items = r1 << cb_log2;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int k1 = it >> cb_log2, ch = it & ((1 << cb_log2) - 1);
        float2 v[R2];
        static_for<0, R2>([&](auto n2) { v[n2] = S[((k1 * R2 + n2) << cb_log2) + ch]; });
        Dft<R2, INV>::run(v);
        static_for<0, R2>([&](auto k2c) { constexpr int k2 = k2c; store(k1 + r1 * k2, ch, v[k2]); });
    }
}

}  // namespace fftreg

// radices built for the two-phase kernels (sizes 64..960 of the FCVSR configs: 64 = 8*8, 180 = 12*15, 320 = 16*20,
// 272 = 16*17, 480 = 20*24, 540 = 20*27, 960 = 30*32); other lengths use the pass-per-radix Stockham kernels
#define FFT2_FOR_EACH_RADIX(X) X(8) X(12) X(15) X(16) X(17) X(20) X(24) X(27) X(30) X(32)

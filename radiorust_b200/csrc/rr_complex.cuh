// Complex helpers and compile-time radix-R DFT butterflies (registers only).
//
// Part of the B200-native replacement for radiorust's IQ sample chain.  The
// reference does all FFT arithmetic through rustfft (Cargo.toml:19, call sites
// src/blocks/filters.rs:244-252); these butterflies are the register-level
// building block of the hand-written replacement.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace rr {

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

// Complex number in registers.  Layout-compatible with float2 / double2 and
// with num::Complex<Flt> (#[repr(C)], re then im).
template <typename T> struct cx {
    T x, y;
    __host__ __device__ __forceinline__ cx() {}
    __host__ __device__ __forceinline__ cx(T a, T b) : x(a), y(b) {}
};

template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return cx<T>(a.x + b.x, a.y + b.y); }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return cx<T>(a.x - b.x, a.y - b.y); }
template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
    return cx<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename T> __device__ __forceinline__ cx<T> cmulc(cx<T> a, cx<T> b) {
    return cx<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> __device__ __forceinline__ cx<T> conj(cx<T> a) { return cx<T>(a.x, -a.y); }
template <typename T> __device__ __forceinline__ cx<T> csqr(cx<T> a) { return cx<T>(a.x * a.x - a.y * a.y, (a.x + a.x) * a.y); }
template <typename T> __device__ __forceinline__ cx<T> cscale(cx<T> a, T s) { return cx<T>(a.x * s, a.y * s); }

// cos/sin(2*pi*k/32), k = 0..31 (enough for radices up to 32).  Indexed with
// compile-time constants after unrolling, so they fold into immediates.
__host__ __device__ __forceinline__ constexpr double tw32_cos(int k) {
    // switch (not a table) so that no local-memory array is ever materialised
    switch (k & 31) {
        case 0: return 1.0;
        case 1: case 31: return 0.98078528040323044913;
        case 2: case 30: return 0.92387953251128675613;
        case 3: case 29: return 0.83146961230254523708;
        case 4: case 28: return 0.70710678118654752440;
        case 5: case 27: return 0.55557023301960222474;
        case 6: case 26: return 0.38268343236508977173;
        case 7: case 25: return 0.19509032201612826785;
        case 8: case 24: return 0.0;
        case 9: case 23: return -0.19509032201612826785;
        case 10: case 22: return -0.38268343236508977173;
        case 11: case 21: return -0.55557023301960222474;
        case 12: case 20: return -0.70710678118654752440;
        case 13: case 19: return -0.83146961230254523708;
        case 14: case 18: return -0.92387953251128675613;
        case 15: case 17: return -0.98078528040323044913;
        default: return -1.0;  // 16
    }
}
__host__ __device__ __forceinline__ constexpr double tw32_sin(int k) { return tw32_cos((k + 24) & 31); }  // sin(x) = cos(x - pi/2)

// v * exp(DIR * -j * 2*pi * k / R), DIR = +1 forward, -1 inverse.  k and R are
// compile-time after unrolling; trivial rotations cost no multiplies.
template <int R, int DIR, typename T> __device__ __forceinline__ cx<T> rot(cx<T> v, int k) {
    const int k32 = ((k * (32 / R)) & 31);          // position on the 32-point circle
    const int e = (DIR > 0) ? k32 : ((32 - k32) & 31);  // forward: exp(-j..), inverse: exp(+j..)
    // multiply by exp(-j*2*pi*e/32) = cos - j sin
    if (e == 0) return v;
    if (e == 8) return cx<T>(v.y, -v.x);            // * -j
    if (e == 16) return cx<T>(-v.x, -v.y);          // * -1
    if (e == 24) return cx<T>(-v.y, v.x);           // * +j
    const T c = (T)tw32_cos(e), s = (T)tw32_sin(e);
    if (e == 4) { const T h = (T)0.70710678118654752440; return cx<T>((v.x + v.y) * h, (v.y - v.x) * h); }
    if (e == 12) { const T h = (T)0.70710678118654752440; return cx<T>((v.y - v.x) * h, -(v.x + v.y) * h); }
    if (e == 20) { const T h = (T)0.70710678118654752440; return cx<T>(-(v.x + v.y) * h, (v.x - v.y) * h); }
    if (e == 28) { const T h = (T)0.70710678118654752440; return cx<T>((v.x - v.y) * h, (v.x + v.y) * h); }
    // (x + jy)(c - js) = (xc + ys) + j(yc - xs)
    return cx<T>(v.x * c + v.y * s, v.y * c - v.x * s);
}

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, typename F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

constexpr __host__ __device__ int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }
constexpr __host__ __device__ int bitrev_c(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// In-register radix-R DFT (R in {2,4,8,16,32}), natural order in and out:
//   v[k] <- sum_n v[n] * exp(DIR * -j*2*pi*n*k/R)
// Decimation-in-frequency radix-2 stages, fully unrolled; the closing
// bit-reversal is compile-time register renaming.
template <int R, int DIR, typename T> __device__ __forceinline__ void dft_regs(cx<T> (&v)[R]) {
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
            for (int n = 0; n < half; ++n) {
                const cx<T> a = v[base + n], b = v[base + n + half];
                v[base + n] = a + b;
                v[base + n + half] = rot<32, DIR, T>(a - b, n * (16 / half));  // W_{2*half}^n on the 32-circle
            }
        }
    }
    constexpr int bits = ilog2c(R);
    cx<T> t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[i] = v[i];
    static_for<0, R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        v[i] = t[bitrev_c(i, bits)];
    });
}

// w^0 .. w^(R-1) from w with multiplication depth <= log2(R) (keeps the
// rounding growth of generated twiddles at a few ulp).
template <int R, typename T> __device__ __forceinline__ void pow_chain(cx<T> w, cx<T> (&p)[R]) {
    p[0] = cx<T>((T)1, (T)0);
    if (R > 1) p[1] = w;
    static_for<2, R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int hi = 1 << ilog2c(i);  // highest power of two <= i
        constexpr int lo = i - hi;
        if constexpr (lo == 0) p[i] = csqr(p[hi / 2]);
        else p[i] = cmul(p[hi], p[lo]);
    });
}

// ---- vectorised memory helpers --------------------------------------------
template <typename T> __device__ __forceinline__ cx<T> ld_cx(const cx<T>* p) {
    typename vec2<T>::type v = *reinterpret_cast<const typename vec2<T>::type*>(p);
    return cx<T>(v.x, v.y);
}
template <typename T> __device__ __forceinline__ void st_cx(cx<T>* p, cx<T> v) {
    typename vec2<T>::type o;
    o.x = v.x;
    o.y = v.y;
    *reinterpret_cast<typename vec2<T>::type*>(p) = o;
}
// two adjacent complex<float> as one 128-bit access
__device__ __forceinline__ void ld_cx2(const cx<float>* p, cx<float>& a, cx<float>& b) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = cx<float>(v.x, v.y);
    b = cx<float>(v.z, v.w);
}
__device__ __forceinline__ void st_cx2(cx<float>* p, cx<float> a, cx<float> b) {
    *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
}
// streaming (read-once) 128-bit global load of two adjacent complex<float>
__device__ __forceinline__ void ldg_stream_cx2(const cx<float>* p, cx<float>& a, cx<float>& b) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    a = cx<float>(v.x, v.y);
    b = cx<float>(v.z, v.w);
}

}  // namespace rr

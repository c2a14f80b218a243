// Packed complex<f32> arithmetic on Blackwell's two-wide fp32 instructions.
//
// sm_100 adds FADD2 / FMUL2 / FFMA2 (PTX add/mul/fma.rn.f32x2): one issue slot,
// two fp32 lanes taken from an aligned register pair.  A complex<f32> kept as
// (re, im) in such a pair makes a complex add one instruction and a complex
// multiply two -- and the operand modifiers of the SASS forms (half swap
// `.LO_HI`, scalar broadcast `.F32`, per-half negate `.NP/.PN`) absorb the
// (im, re) swap, the (wr, wr) broadcast and the sign pattern of a complex
// product, so none of them costs an instruction.  ptxas folds them when the
// 64-bit operand is assembled with mov.b64 from the two 32-bit halves, which is
// why every helper below packs its operands right at the instruction.
//
// The arithmetic (one rounding per add/mul, fused multiply-add) is the same as
// the scalar fp32 path of rr_complex.cuh; the fp32 pipe does the same number of
// lane operations, the issue slots halve (measured on B200: FFMA 3.94, FFMA2
// 1.99 warp-instructions/clk/SM, both 127 lane-FMA/clk/SM).
//
// Used by the IQ-chain kernels that replace rustfft in the reference's Filter
// (src/blocks/filters.rs:244-252).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rr_complex.cuh"

namespace rr {

struct pc {
    float x, y;
    __device__ __forceinline__ pc() {}
    __device__ __forceinline__ pc(float a, float b) : x(a), y(b) {}
};

__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ pc upk2(unsigned long long v) {
    pc r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ pc operator+(pc a, pc b) { return upk2(f2_add(pk2(a.x, a.y), pk2(b.x, b.y))); }
__device__ __forceinline__ pc operator-(pc a, pc b) { return upk2(f2_sub(pk2(a.x, a.y), pk2(b.x, b.y))); }
// a + j*b and a - j*b (the swap and the signs ride on the operand modifiers)
__device__ __forceinline__ pc add_j(pc a, pc b) { return upk2(f2_add(pk2(a.x, a.y), pk2(-b.y, b.x))); }
__device__ __forceinline__ pc sub_j(pc a, pc b) { return upk2(f2_add(pk2(a.x, a.y), pk2(b.y, -b.x))); }
// a * s, a * s + c (real scalar)
__device__ __forceinline__ pc pscale(pc a, float s) { return upk2(f2_mul(pk2(a.x, a.y), pk2(s, s))); }
__device__ __forceinline__ pc pfma_s(pc a, float s, pc c) { return upk2(f2_fma(pk2(a.x, a.y), pk2(s, s), pk2(c.x, c.y))); }
// a * w
__device__ __forceinline__ pc pcmul(pc a, pc w) {
    const unsigned long long t = f2_mul(pk2(a.x, a.y), pk2(w.x, w.x));
    return upk2(f2_fma(pk2(a.y, a.x), pk2(-w.y, w.y), t));
}
// a * conj(w)
__device__ __forceinline__ pc pcmulc(pc a, pc w) {
    const unsigned long long t = f2_mul(pk2(a.x, a.y), pk2(w.x, w.x));
    return upk2(f2_fma(pk2(a.y, a.x), pk2(w.y, -w.y), t));
}
// a * w + c
__device__ __forceinline__ pc pcfma(pc a, pc w, pc c) {
    const unsigned long long t = f2_fma(pk2(a.x, a.y), pk2(w.x, w.x), pk2(c.x, c.y));
    return upk2(f2_fma(pk2(a.y, a.x), pk2(-w.y, w.y), t));
}
// a * (c - j*s) with compile-time-constant c, s: (x*c + y*s, y*c - x*s)
__device__ __forceinline__ pc prot_cs(pc a, float c, float s) {
    const unsigned long long t = f2_mul(pk2(a.x, a.y), pk2(c, c));
    return upk2(f2_fma(pk2(a.y, a.x), pk2(s, -s), t));
}

// v * exp(DIR * -j*2*pi*k/32); k is a compile-time constant after unrolling
template <int DIR> __device__ __forceinline__ pc prot32(pc v, int k) {
    const int k32 = k & 31;
    const int e = (DIR > 0) ? k32 : ((32 - k32) & 31);
    if (e == 0) return v;
    if (e == 8) return pc(v.y, -v.x);
    if (e == 16) return pc(-v.x, -v.y);
    if (e == 24) return pc(-v.y, v.x);
    return prot_cs(v, (float)tw32_cos(e), (float)tw32_sin(e));
}

// In-register radix-R DFT, natural order in and out (decimation in frequency,
// bit reversal by register renaming): v[k] <- sum_n v[n] exp(DIR * -j*2*pi*n*k/R)
template <int R, int DIR> __device__ __forceinline__ void pdft_regs(pc (&v)[R]) {
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
            for (int n = 0; n < half; ++n) {
                const pc a = v[base + n], b = v[base + n + half];
                const int k = n * (16 / half);  // W_{2*half}^n on the 32-circle
                const int k32 = k & 31;
                const int e = (DIR > 0) ? k32 : ((32 - k32) & 31);
                v[base + n] = a + b;
                if (e == 0) v[base + n + half] = a - b;
                else if (e == 8) {  // (a - b) * -j = (dy, -dx)
                    const pc d = a - b;
                    v[base + n + half] = pc(d.y, -d.x);
                } else if (e == 24) {
                    const pc d = a - b;
                    v[base + n + half] = pc(-d.y, d.x);
                } else if (e == 16) v[base + n + half] = b - a;
                else v[base + n + half] = prot_cs(a - b, (float)tw32_cos(e), (float)tw32_sin(e));
            }
        }
    }
    constexpr int bits = ilog2c(R);
    pc t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[i] = v[i];
    static_for<0, R>([&](auto I) {
        constexpr int i = decltype(I)::value;
        v[i] = t[bitrev_c(i, bits)];
    });
}

}  // namespace rr

// Three-pass shared-memory FFT plan: N = R1*R2*R3 points per (thread group),
// NT threads, every butterfly in registers, two shared-memory exchanges per
// transform.
//
// Forward = decimation in frequency (natural in, digit-reversed out);
// inverse = decimation in time (digit-reversed in, natural out).  Because the
// last forward pass and the first inverse pass touch the same R3 elements in
// the same thread, the H(f) multiply of overlap-save fast convolution
// (reference: src/blocks/filters.rs:244-252, rustfft forward -> multiply ->
// rustfft inverse) is fused between them with no memory round trip; H is
// stored pre-permuted by the host (see FftPlan::hperm_index).
//
// Index algebra (n = input index, k = output bin):
//   n = n1*S1 + n2*S2 + n3,  S1 = R2*R3, S2 = R3
//   k = k1 + R1*k2 + R1*R2*k3, stored at position k1*S1 + k2*S2 + k3
//   pass 1: DFT_R1 over n1, times W_N^(k1*(n2*S2+n3))
//   pass 2: DFT_R2 over n2, times W_(R2*R3)^(k2*n3)
//   pass 3: DFT_R3 over n3
#pragma once
#include "rr_complex.cuh"

namespace rr {

// G > 1 interleaves G independent transforms in shared memory (element p of
// transform g at G*sidx(p) + g; pass `sm + g`): lanes that work on the same
// element of neighbouring transforms then touch consecutive words, which is
// what the strided column passes of the four-step FFT (rr_big_os.cu) need.
template <typename T, int R1_, int R2_, int R3_, int NT_, int G_ = 1> struct FftPlan {
    static constexpr int R1 = R1_, R2 = R2_, R3 = R3_, NT = NT_;
    static constexpr int N = R1 * R2 * R3;
    static constexpr int S1 = R2 * R3, S2 = R3;
    static constexpr int B1 = N / R1 / NT, B2 = N / R2 / NT, B3 = N / R3 / NT;
    static_assert(B1 * R1 * NT == N && B2 * R2 * NT == N && B3 * R3 * NT == N, "thread count must divide every pass");
    static_assert(R3 >= 2 && (R3 & (R3 - 1)) == 0, "R3 must be a power of two");
    // padding: PAD elements after every run of R3.  One element: pass 3's per-thread runs start (R3+1)*sizeof(cx) bytes
    // apart (all of a half warp's 8-byte pieces on distinct banks), and the groups of R3 lanes that pass 2 puts
    // (R2*R3 + R2)*sizeof(cx) bytes apart alternate between the two halves of the 128-byte bank window.  (Two elements --
    // 16-byte aligned runs and vector accesses for f32 -- were measured 4-9 % slower in the overlap-save kernels: the
    // vector stores of pass 1 and the 8-byte accesses of pass 2 then take twice their minimum of wavefronts.)
    static constexpr int PAD = 1;
    static constexpr int LOG_R3 = ilog2c(R3);
    static constexpr int G = G_;
    static constexpr int SMEM_ELEMS = G * (N + PAD * (N / R3));
    static constexpr bool VEC1 = (G == 1) && (sizeof(T) == 4) && (B1 % 2 == 0) && (PAD % 2 == 0);
    static constexpr bool VEC2 = (G == 1) && (sizeof(T) == 4) && (B2 % 2 == 0) && (PAD % 2 == 0);
    static constexpr bool VEC3 = (G == 1) && (sizeof(T) == 4) && (PAD % 2 == 0);

    __host__ __device__ static constexpr int sidx(int p) { return G * (p + PAD * (p >> LOG_R3)); }

    // position (in the permuted H table) of true bin k, so that pass 3's
    // butterfly u reads H for output k3 at hperm[k3*(N/R3) + u], coalesced.
    __host__ __device__ static constexpr int hperm_index(int k) {
        const int k1 = k % R1, k2 = (k / R1) % R2, k3 = k / (R1 * R2);
        return k3 * (N / R3) + (k1 * R2 + k2);
    }
    // position in shared memory (before padding) of true bin k after the
    // forward transform
    __host__ __device__ static constexpr int bin_position(int k) {
        const int k1 = k % R1, k2 = (k / R1) % R2, k3 = k / (R1 * R2);
        return k1 * S1 + k2 * S2 + k3;
    }

    // per-thread loop-invariant twiddle bases
    cx<T> w1[B1];  // W_N^q for this thread's pass-1 columns q
    cx<T> w2[B2];  // W_(R2*R3)^n3 for this thread's pass-2 butterflies

    // twN[e] = exp(-j*2*pi*e/N), e < N  (host-computed in extended precision)
    __device__ __forceinline__ void init(const cx<T>* __restrict__ twN, int tid) {
#pragma unroll
        for (int b = 0; b < B1; ++b) w1[b] = ld_cx(&twN[B1 * tid + b]);
#pragma unroll
        for (int b = 0; b < B2; ++b) w2[b] = ld_cx(&twN[((B2 * tid + b) & (R3 - 1)) * R1]);
    }

    __device__ __forceinline__ static int q_of(int tid, int b) { return B1 * tid + b; }

    // ---- pass 1 forward: data already in registers: v[b][n1] = x[q_b + S1*n1]
    __device__ __forceinline__ void p1_forward(cx<T>* __restrict__ sm, int tid, cx<T> (&v)[B1][R1]) const {
#pragma unroll
        for (int b = 0; b < B1; ++b) {
            dft_regs<R1, +1, T>(v[b]);
            cx<T> p[R1];
            pow_chain<R1, T>(w1[b], p);
#pragma unroll
            for (int k = 1; k < R1; ++k) v[b][k] = cmul(v[b][k], p[k]);
        }
        if constexpr (VEC1) {
#pragma unroll
            for (int b = 0; b < B1; b += 2) {
                const int q = B1 * tid + b;
#pragma unroll
                for (int k = 0; k < R1; ++k) st_cx2(&sm[sidx(k * S1 + q)], v[b][k], v[b + 1][k]);
            }
        } else {
#pragma unroll
            for (int b = 0; b < B1; ++b) {
                const int q = B1 * tid + b;
#pragma unroll
                for (int k = 0; k < R1; ++k) st_cx(&sm[sidx(k * S1 + q)], v[b][k]);
            }
        }
    }

    // ---- pass 1 inverse: v[b][n1] = y[q_b + S1*n1] (natural order) -----------
    __device__ __forceinline__ void p1_inverse(const cx<T>* __restrict__ sm, int tid, cx<T> (&v)[B1][R1]) const {
        if constexpr (VEC1) {
#pragma unroll
            for (int b = 0; b < B1; b += 2) {
                const int q = B1 * tid + b;
#pragma unroll
                for (int k = 0; k < R1; ++k) ld_cx2(&sm[sidx(k * S1 + q)], v[b][k], v[b + 1][k]);
            }
        } else {
#pragma unroll
            for (int b = 0; b < B1; ++b) {
                const int q = B1 * tid + b;
#pragma unroll
                for (int k = 0; k < R1; ++k) v[b][k] = ld_cx(&sm[sidx(k * S1 + q)]);
            }
        }
#pragma unroll
        for (int b = 0; b < B1; ++b) {
            cx<T> p[R1];
            pow_chain<R1, T>(w1[b], p);
#pragma unroll
            for (int k = 1; k < R1; ++k) v[b][k] = cmulc(v[b][k], p[k]);
            dft_regs<R1, -1, T>(v[b]);
        }
    }

    // ---- pass 2, in place ---------------------------------------------------
    template <int DIR> __device__ __forceinline__ void p2(cx<T>* __restrict__ sm, int tid) const {
        cx<T> v[B2][R2];
        int base[B2];
#pragma unroll
        for (int b = 0; b < B2; ++b) {
            const int id = B2 * tid + b;
            base[b] = (id >> LOG_R3) * S1 + (id & (R3 - 1));
        }
        if constexpr (VEC2) {
#pragma unroll
            for (int b = 0; b < B2; b += 2)
#pragma unroll
                for (int n = 0; n < R2; ++n) ld_cx2(&sm[sidx(base[b] + n * S2)], v[b][n], v[b + 1][n]);
        } else {
#pragma unroll
            for (int b = 0; b < B2; ++b)
#pragma unroll
                for (int n = 0; n < R2; ++n) v[b][n] = ld_cx(&sm[sidx(base[b] + n * S2)]);
        }
#pragma unroll
        for (int b = 0; b < B2; ++b) {
            cx<T> p[R2];
            pow_chain<R2, T>(w2[b], p);
            if (DIR > 0) {
                dft_regs<R2, +1, T>(v[b]);
#pragma unroll
                for (int k = 1; k < R2; ++k) v[b][k] = cmul(v[b][k], p[k]);
            } else {
#pragma unroll
                for (int k = 1; k < R2; ++k) v[b][k] = cmulc(v[b][k], p[k]);
                dft_regs<R2, -1, T>(v[b]);
            }
        }
        if constexpr (VEC2) {
#pragma unroll
            for (int b = 0; b < B2; b += 2)
#pragma unroll
                for (int n = 0; n < R2; ++n) st_cx2(&sm[sidx(base[b] + n * S2)], v[b][n], v[b + 1][n]);
        } else {
#pragma unroll
            for (int b = 0; b < B2; ++b)
#pragma unroll
                for (int n = 0; n < R2; ++n) st_cx(&sm[sidx(base[b] + n * S2)], v[b][n]);
        }
    }

    // ---- pass 3 helpers -------------------------------------------------------
    __device__ __forceinline__ static void p3_load(const cx<T>* __restrict__ sm, int u, cx<T> (&v)[R3]) {
        const cx<T>* run = &sm[sidx(u * R3)];
        if constexpr (VEC3) {
#pragma unroll
            for (int n = 0; n < R3; n += 2) ld_cx2(&run[n], v[n], v[n + 1]);
        } else {
#pragma unroll
            for (int n = 0; n < R3; ++n) v[n] = ld_cx(&run[G * n]);
        }
    }
    __device__ __forceinline__ static void p3_store(cx<T>* __restrict__ sm, int u, const cx<T> (&v)[R3]) {
        cx<T>* run = &sm[sidx(u * R3)];
        if constexpr (VEC3) {
#pragma unroll
            for (int n = 0; n < R3; n += 2) st_cx2(&run[n], v[n], v[n + 1]);
        } else {
#pragma unroll
            for (int n = 0; n < R3; ++n) st_cx(&run[G * n], v[n]);
        }
    }

    // forward pass 3 -> multiply by H (permuted layout) -> inverse pass 3
    __device__ __forceinline__ static void p3_fwd_mul_inv(cx<T>* __restrict__ sm, int tid, const cx<T>* __restrict__ hperm) {
#pragma unroll
        for (int b = 0; b < B3; ++b) {
            const int u = B3 * tid + b;
            cx<T> v[R3];
            p3_load(sm, u, v);
            dft_regs<R3, +1, T>(v);
#pragma unroll
            for (int k = 0; k < R3; ++k) v[k] = cmul(v[k], ld_cx(&hperm[k * (N / R3) + u]));
            dft_regs<R3, -1, T>(v);
            p3_store(sm, u, v);
        }
    }
    template <int DIR> __device__ __forceinline__ static void p3(cx<T>* __restrict__ sm, int tid) {
#pragma unroll
        for (int b = 0; b < B3; ++b) {
            const int u = B3 * tid + b;
            cx<T> v[R3];
            p3_load(sm, u, v);
            dft_regs<R3, DIR, T>(v);
            p3_store(sm, u, v);
        }
    }
};

// Plan selection per transform size.  Every pass uses all NT threads.
template <typename T, int N, int G = 1> struct PlanFor;
#define RR_PLAN(TYPE, NN, A, B, C, THREADS) \
    template <int G> struct PlanFor<TYPE, NN, G> { using type = FftPlan<TYPE, A, B, C, THREADS, G>; }
RR_PLAN(float, 64, 4, 4, 4, 16);
RR_PLAN(float, 128, 4, 4, 8, 16);
RR_PLAN(float, 256, 4, 8, 8, 32);
RR_PLAN(float, 512, 8, 8, 8, 64);
RR_PLAN(float, 1024, 8, 8, 16, 64);
RR_PLAN(float, 2048, 8, 16, 16, 128);
RR_PLAN(float, 4096, 16, 16, 16, 128);
RR_PLAN(float, 8192, 16, 32, 16, 256);
RR_PLAN(float, 16384, 16, 32, 32, 512);
RR_PLAN(double, 64, 4, 4, 4, 16);
RR_PLAN(double, 128, 4, 4, 8, 16);
RR_PLAN(double, 256, 4, 8, 8, 32);
RR_PLAN(double, 512, 8, 8, 8, 64);
RR_PLAN(double, 1024, 8, 8, 16, 64);
RR_PLAN(double, 2048, 8, 16, 16, 128);
RR_PLAN(double, 4096, 16, 16, 16, 256);
#undef RR_PLAN

}  // namespace rr

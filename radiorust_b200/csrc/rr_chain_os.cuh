// Fused device-resident chain: FreqShifter -> Filter (-> Downsampler).
//
// One CTA owns one stream and walks its chunks in time order, so the filter
// history (one chunk, held in REGISTERS across iterations), the decimator's
// L-1 sample tail (shared memory) and the NCO phase never leave the SM between
// chunks; HBM sees each input sample once and each output sample once.
//
// Replaces the three Tokio tasks of the reference and their hot loops:
//   NCO/mixer        src/blocks/transform.rs:341-348 (table :321-340)
//   overlap-save     src/blocks/filters.rs:240-253   (2 rustfft calls + H mul)
//   decimating FIR   src/blocks/resampling.rs:103-121
// Per chunk: [prev|cur] (2n) -> FFT -> *H -> IFFT -> first n samples -> FIR at
// the firing instants j_m = ceil(m*P/Q).
#pragma once
#include "rr_fft_plan.cuh"
#include "rr_kernels.h"

namespace rr {

template <typename T> __device__ __forceinline__ void sincos_t(T x, T* s, T* c);
template <> __device__ __forceinline__ void sincos_t<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ void sincos_t<double>(double x, double* s, double* c) { sincos(x, s, c); }

// exp(j * sign * (start + (i/denom)*TAU)) evaluated like transform.rs:335 in T
template <typename T> __device__ __forceinline__ cx<T> nco_phasor(uint32_t i, uint32_t denom, int sign, T start) {
    const T tau = (T)6.283185307179586476925286766559;
    const T frac = (T)i / (T)denom;
    const T ph = start + (sign < 0 ? -(frac * tau) : frac * tau);
    T s, c;
    sincos_t<T>(ph, &s, &c);
    return cx<T>(c, s);
}

__device__ __forceinline__ uint32_t mulmod_u32(uint32_t a, uint32_t b, uint32_t m) {
    return (uint32_t)(((unsigned long long)a * (unsigned long long)b) % m);
}
__device__ __forceinline__ uint32_t addmod_u32(uint32_t a, uint32_t b, uint32_t m) {
    const unsigned long long s = (unsigned long long)a + b;
    return (uint32_t)(s >= m ? s - m : s);
}

template <typename T, int N> struct ChainOsCfg {
    static constexpr int NT = PlanFor<T, N>::type::NT;
    static constexpr int MIN_CTAS = (sizeof(T) == 4 && N <= 8192) ? 2 : 1;
};

template <typename T, int N, int EPI>
__global__ void __launch_bounds__(ChainOsCfg<T, N>::NT, ChainOsCfg<T, N>::MIN_CTAS)
k_chain_os(const ChainOsArgs<T> a) {
    using P = typename PlanFor<T, N>::type;
    constexpr int NT = P::NT, R1 = P::R1, B1 = P::B1, S1 = P::S1, H1 = R1 / 2;
    constexpr int n = N / 2;
    const int tid = threadIdx.x;
    const int s = blockIdx.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* zbuf = sm + P::SMEM_ELEMS;                               // [L-1 | n]   (EPI == 1)
    T* ir_s = reinterpret_cast<T*>(zbuf + (n + (EPI ? a.L - 1 : 0)));  // [L]

    const cx<T>* __restrict__ in = reinterpret_cast<const cx<T>*>(a.in) + (long long)s * a.in_stride;
    cx<T>* __restrict__ out = reinterpret_cast<cx<T>*>(a.out) + (long long)s * a.out_stride;
    const cx<T>* __restrict__ hist_gi = reinterpret_cast<const cx<T>*>(a.hist_in) + (long long)s * a.hist_stride;
    cx<T>* __restrict__ hist_go = reinterpret_cast<cx<T>*>(a.hist_out) + (long long)s * n;
    const cx<T>* __restrict__ hperm = reinterpret_cast<const cx<T>*>(a.hperm);

    P plan;
    plan.init(reinterpret_cast<const cx<T>*>(a.twN), tid);

    // ---- chunk range of this CTA (gridDim.y > 1 only for EPI == 0) ----------
    const int c_first = a.first_is_history ? 1 : 0;
    int cb = c_first, ce = a.n_chunks;
    if (gridDim.y > 1) {
        const int total = a.n_chunks - c_first;
        const int per = (total + (int)gridDim.y - 1) / (int)gridDim.y;
        cb = c_first + (int)blockIdx.y * per;
        ce = min(cb + per, a.n_chunks);
        if (cb >= ce && !(blockIdx.y == 0 && total <= 0)) return;
    }
    const bool last_part = (ce >= a.n_chunks);

    // ---- NCO per-thread constants ------------------------------------------
    const bool has_nco = (a.nco != nullptr);
    uint32_t denom = 1, numer_abs = 0, idx0 = 0, n_mod = 0;
    int sign = 0;
    T start = (T)0;
    cx<T> rot_j((T)1, (T)0);  // phasor advance for +S1 samples
    if (has_nco) {
        const NcoStream ns = a.nco[s];
        denom = ns.denom;
        numer_abs = ns.numer_abs;
        sign = ns.sign;
        start = (T)ns.start_phase;
        idx0 = (uint32_t)(((unsigned long long)ns.idx + (unsigned long long)(a.nco_offset % (long long)denom)) % denom);
        n_mod = (uint32_t)(n % denom);
        const uint32_t step_j = mulmod_u32(numer_abs, (uint32_t)(S1 % denom), denom);
        double sj, cj;
        sincospi(2.0 * (double)step_j / (double)denom, &sj, &cj);
        rot_j = cx<T>((T)cj, (T)(sign < 0 ? -sj : sj));
    }

    // load one chunk (mixing it with the NCO) into registers
    auto load_chunk = [&](int c, cx<T> (&dst)[B1][H1]) {
        const cx<T>* src = in + (long long)c * n;
        if constexpr (P::VEC1) {
#pragma unroll
            for (int b = 0; b < B1; b += 2)
#pragma unroll
                for (int j = 0; j < H1; ++j) ldg_stream_cx2(&src[B1 * tid + b + S1 * j], dst[b][j], dst[b + 1][j]);
        } else {
#pragma unroll
            for (int b = 0; b < B1; ++b)
#pragma unroll
                for (int j = 0; j < H1; ++j) dst[b][j] = ld_cx(&src[B1 * tid + b + S1 * j]);
        }
        if (has_nco) {
            // table index of the chunk's first sample: (idx0 + c*n) mod denom
            const uint32_t kc = addmod_u32(idx0, mulmod_u32((uint32_t)c % denom, n_mod, denom), denom);
#pragma unroll
            for (int b = 0; b < B1; ++b) {
                const uint32_t k0 = addmod_u32(kc, (uint32_t)(B1 * tid + b) % denom, denom);
                cx<T> ph = nco_phasor<T>(mulmod_u32(numer_abs, k0, denom), denom, sign, start);
#pragma unroll
                for (int j = 0; j < H1; ++j) {
                    dst[b][j] = cmul(dst[b][j], ph);
                    if (j + 1 < H1) ph = cmul(ph, rot_j);
                }
            }
        }
    };

    // ---- prologue: history, taps, decimator tail ----------------------------
    cx<T> hist[B1][H1];
    if (cb > 0) {
        load_chunk(cb - 1, hist);  // history-only first chunk, or the chunk before this part
    } else {
#pragma unroll
        for (int b = 0; b < B1; ++b)
#pragma unroll
            for (int j = 0; j < H1; ++j) hist[b][j] = ld_cx(&hist_gi[B1 * tid + b + S1 * j]);
    }
    int L = 0;
    long long zJ = 0;  // z samples consumed by the decimator before the current chunk
    if constexpr (EPI == 1) {
        L = a.L;
        const cx<T>* zt = reinterpret_cast<const cx<T>*>(a.ztail_in) + (long long)s * (L - 1);
        for (int i = tid; i < L - 1; i += NT) zbuf[i] = zt[i];
        for (int i = tid; i < L; i += NT) ir_s[i] = a.ir[i];
        zJ = a.rate.j0;
    }

    for (int c = cb; c < ce; ++c) {
        cx<T> v[B1][R1];
        {
            cx<T> cur[B1][H1];
            load_chunk(c, cur);
#pragma unroll
            for (int b = 0; b < B1; ++b)
#pragma unroll
                for (int j = 0; j < H1; ++j) {
                    v[b][j] = hist[b][j];
                    v[b][j + H1] = cur[b][j];
                    hist[b][j] = cur[b][j];
                }
        }
        plan.p1_forward(sm, tid, v);
        __syncthreads();
        plan.template p2<+1>(sm, tid);
        __syncthreads();
        P::p3_fwd_mul_inv(sm, tid, hperm);
        __syncthreads();
        plan.template p2<-1>(sm, tid);
        __syncthreads();
        plan.p1_inverse(sm, tid, v);  // v[b][j], j < H1: filter output sample B1*tid+b + S1*j

        if constexpr (EPI == 0) {
            if (a.emit) {
                cx<T>* dst = out + (long long)(c - c_first) * n;
                if constexpr (P::VEC1) {
#pragma unroll
                    for (int b = 0; b < B1; b += 2)
#pragma unroll
                        for (int j = 0; j < H1; ++j) st_cx2(&dst[B1 * tid + b + S1 * j], v[b][j], v[b + 1][j]);
                } else {
#pragma unroll
                    for (int b = 0; b < B1; ++b)
#pragma unroll
                        for (int j = 0; j < H1; ++j) st_cx(&dst[B1 * tid + b + S1 * j], v[b][j]);
                }
            }
            __syncthreads();  // pass-1 inverse reads done before the next pass-1 forward writes
        } else {
            cx<T>* zc = zbuf + (L - 1);
            // (L-1) may be odd: per-element stores keep this alignment-safe
#pragma unroll
            for (int b = 0; b < B1; ++b)
#pragma unroll
                for (int j = 0; j < H1; ++j) st_cx(&zc[B1 * tid + b + S1 * j], v[b][j]);
            __syncthreads();
            if (a.emit) {
                // outputs m with firing count j_m in (zJ, zJ + n]
                const long long Pq = a.rate.P, Qq = a.rate.Q;
                const long long m_lo = (zJ * Qq) / Pq + 1;
                const long long m_hi = ((zJ + n) * Qq) / Pq;
                if constexpr (NT >= 32) {
                    const int lane = tid & 31, warp = tid >> 5;
                    for (long long m = m_lo + warp; m <= m_hi; m += NT / 32) {
                        const long long jm = (m * Pq + Qq - 1) / Qq;
                        const cx<T>* win = zbuf + (int)(jm - zJ - 1);  // window = win[0 .. L)
                        T ax = (T)0, ay = (T)0;
                        for (int t = lane; t < L; t += 32) {
                            const cx<T> z = win[t];
                            const T h = ir_s[t];
                            ax = fma(z.x, h, ax);
                            ay = fma(z.y, h, ay);
                        }
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
                            ax += __shfl_xor_sync(0xffffffffu, ax, o);
                            ay += __shfl_xor_sync(0xffffffffu, ay, o);
                        }
                        if (lane == 0) st_cx(&out[m - a.rate.m0 - 1], cx<T>(ax, ay));
                    }
                } else {
                    for (long long m = m_lo + tid; m <= m_hi; m += NT) {
                        const long long jm = (m * Pq + Qq - 1) / Qq;
                        const cx<T>* win = zbuf + (int)(jm - zJ - 1);
                        T ax = (T)0, ay = (T)0;
                        for (int t = 0; t < L; ++t) {
                            ax = fma(win[t].x, ir_s[t], ax);
                            ay = fma(win[t].y, ir_s[t], ay);
                        }
                        st_cx(&out[m - a.rate.m0 - 1], cx<T>(ax, ay));
                    }
                }
            }
            zJ += n;
            __syncthreads();
            for (int i = tid; i < L - 1; i += NT) zbuf[i] = zbuf[n + i];  // keep the last L-1 samples
            // (made visible by the barrier after the next pass-1 forward)
        }
    }

    // ---- epilogue: persist state --------------------------------------------
    if (last_part && a.hist_out != nullptr) {
#pragma unroll
        for (int b = 0; b < B1; ++b)
#pragma unroll
            for (int j = 0; j < H1; ++j) st_cx(&hist_go[B1 * tid + b + S1 * j], hist[b][j]);
    }
    if constexpr (EPI == 1) {
        __syncthreads();
        cx<T>* zt = reinterpret_cast<cx<T>*>(a.ztail_out) + (long long)s * (L - 1);
        for (int i = tid; i < L - 1; i += NT) zt[i] = zbuf[i];
    }
}

template <typename T, int N, int EPI>
cudaError_t launch_chain_os_n(int n_streams, int parts, const ChainOsArgs<T>& a, cudaStream_t st) {
    using P = typename PlanFor<T, N>::type;
    constexpr int n = N / 2;
    size_t smem = sizeof(cx<T>) * P::SMEM_ELEMS;
    if (EPI == 1) smem += sizeof(cx<T>) * (size_t)(n + a.L - 1) + sizeof(T) * (size_t)a.L;
    if (EPI == 1 || parts < 1) parts = 1;
    auto kern = k_chain_os<T, N, EPI>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3((unsigned)n_streams, (unsigned)parts), P::NT, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace rr

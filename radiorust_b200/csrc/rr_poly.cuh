// Polyphase fused chain: FreqShifter -> Filter -> Downsampler in one kernel that
// never forms the full-rate filter output.
//
// The reference runs, per stream:
//   x'[t] = x[t] * nco[t]                                  transform.rs:341-348
//   z[u]  = sum_m h[m] * x'[u - m]      (2 FFTs of 2n)     filters.rs:240-253
//   y_m   = sum_t ir[t] * z[j_m - L + t], j_m = ceil(m*P/Q)   resampling.rs:103-121
// Both filters are FIR, so y_m = sum_tau g[tau] * x'[j_m - 1 - tau] with
// g = h * reverse(ir).  Writing m = Q*I + q gives j_m = I*P + s_q: Q output
// phases, each a decimation by P.  Splitting x' into its P polyphase branches
// x_p[i] = x'[base + i*P + p] turns every phase into
//   y_q[i] = sum_p (G[q][p] (*) x_p)[i],   G[q][p][l] = g[P - 1 + s_q - p + l*P]
// i.e. P short FFTs of K points (instead of 2 FFTs of 2n points per n samples),
// a multiply-accumulate against FFT(G) and Q inverse FFTs of K points per
// V*P input samples, V = K - 1 - Lmax valid outputs per block and phase.
//
// Launch shape: one CTA = G thread groups; group g transforms branch p0 + g of
// the current round, so G lanes read G consecutive input samples; the groups'
// transforms are interleaved in shared memory (FftPlan<.., G>).  Accumulators
// live in registers across the P/G rounds, are reduced over groups through
// shared memory and kept there per block; at the end the groups run the
// inverse transforms of different (block, phase) jobs concurrently.
#pragma once
#include "rr_chain_os.cuh"

namespace rr {

template <typename T, int K, int G> struct PolyCfg {
    using Plan = typename PlanFor<T, K, G>::type;
    static constexpr int THREADS = Plan::NT * G;
    static constexpr int MIN_CTAS = (sizeof(T) == 4 && THREADS <= 512) ? 2 : 1;
};

// exact NCO phasor for push-relative sample offset `off` (may be negative):
// table index k = (idx + off) mod denom, i_k = numer*k mod denom (transform.rs:333-338)
template <typename T>
__device__ __forceinline__ cx<T> nco_phasor_at(long long off, uint32_t idx, uint32_t numer_abs, uint32_t denom, int sign, T start) {
    long long k = ((long long)idx + off) % (long long)denom;
    if (k < 0) k += denom;
    return nco_phasor<T>(mulmod_u32(numer_abs, (uint32_t)k, denom), denom, sign, start);
}

template <typename T, int K, int Q, int G>
__global__ void __launch_bounds__(PolyCfg<T, K, G>::THREADS, PolyCfg<T, K, G>::MIN_CTAS)
k_poly(const PolyArgs<T> a) {
    using P = typename PolyCfg<T, K, G>::Plan;
    constexpr int NT = P::NT, R1 = P::R1, B1 = P::B1, S1 = P::S1, R3 = P::R3, B3 = P::B3;
    constexpr int THREADS = NT * G;
    constexpr int KR = K / R3;
    const int tid = threadIdx.x;
    const int g = tid % G, t = tid / G;
    const int s = blockIdx.y;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* W = reinterpret_cast<cx<T>*>(smem_raw);  // interleaved work area, >= K*G elements
    cx<T>* Ysave = W + P::SMEM_ELEMS;                // [nbpc*Q][K]
    cx<T>* smg = W + g;

    const cx<T>* __restrict__ in = reinterpret_cast<const cx<T>*>(a.in) + (long long)s * a.in_stride;
    const cx<T>* __restrict__ hist = reinterpret_cast<const cx<T>*>(a.hist2) + (long long)s * 2 * a.n;
    const cx<T>* __restrict__ gtab = reinterpret_cast<const cx<T>*>(a.gtab);
    cx<T>* __restrict__ out = reinterpret_cast<cx<T>*>(a.out) + (long long)s * a.out_stride;
    const long long Pd = a.P;
    const long long len = a.len, hist_len = 2 * a.n;

    P plan;
    plan.init(reinterpret_cast<const cx<T>*>(a.twK), t);

    const bool has_nco = (a.nco != nullptr);
    uint32_t denom = 1, numer_abs = 0, idx0 = 0;
    int sign = 0;
    T start = (T)0;
    cx<T> rotG((T)1, (T)0);  // phasor advance for +G samples
    if (has_nco) {
        const NcoStream ns = a.nco[s];
        denom = ns.denom;
        numer_abs = ns.numer_abs;
        sign = ns.sign;
        start = (T)ns.start_phase;
        idx0 = ns.idx;
        const uint32_t step = mulmod_u32(numer_abs, (uint32_t)(G % denom), denom);
        double sj, cj;
        sincospi(2.0 * (double)step / (double)denom, &sj, &cj);
        rotG = cx<T>((T)cj, (T)(sign < 0 ? -sj : sj));
    }

    const int blk0 = blockIdx.x * a.nbpc;
    const int blk1 = min(blk0 + a.nbpc, a.n_blocks);

    for (int blk = blk0; blk < blk1; ++blk) {
        const long long Ibase = a.I_lo + (long long)blk * a.V - a.Lmax;
        const long long boff = Ibase * Pd - a.J0 - Pd;  // push-relative offset of local element (i = 0, p = 0)
        // blocks whose whole input window lies inside the pushed samples skip the range checks
        const bool interior = (boff >= 0) && (boff + (long long)K * Pd <= len);

        cx<T> ph[B1][R1];
        if (has_nco) {
#pragma unroll
            for (int bb = 0; bb < B1; ++bb)
#pragma unroll
                for (int r = 0; r < R1; ++r) {
                    const long long off = boff + (long long)(B1 * t + bb + S1 * r) * Pd + g;
                    ph[bb][r] = nco_phasor_at<T>(off, idx0, numer_abs, denom, sign, start);
                }
        }
        cx<T> acc[Q][B3][R3];
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int b = 0; b < B3; ++b)
#pragma unroll
                for (int k = 0; k < R3; ++k) acc[q][b][k] = cx<T>((T)0, (T)0);

        for (long long p0 = 0; p0 < Pd; p0 += G) {
            const long long p = p0 + g;
            const bool active = p < Pd;
            cx<T> v[B1][R1];
            if (interior) {
#pragma unroll
                for (int bb = 0; bb < B1; ++bb)
#pragma unroll
                    for (int r = 0; r < R1; ++r) {
                        const long long off = boff + (long long)(B1 * t + bb + S1 * r) * Pd + p;
                        v[bb][r] = active ? ld_cx(&in[off]) : cx<T>((T)0, (T)0);
                    }
                if (has_nco) {
#pragma unroll
                    for (int bb = 0; bb < B1; ++bb)
#pragma unroll
                        for (int r = 0; r < R1; ++r) {
                            v[bb][r] = cmul(v[bb][r], ph[bb][r]);
                            ph[bb][r] = cmul(ph[bb][r], rotG);
                        }
                }
            } else {
#pragma unroll
                for (int bb = 0; bb < B1; ++bb)
#pragma unroll
                    for (int r = 0; r < R1; ++r) {
                        const long long off = boff + (long long)(B1 * t + bb + S1 * r) * Pd + p;
                        cx<T> x((T)0, (T)0);
                        if (active) {
                            if (off >= 0) {
                                if (off < len) {
                                    x = ld_cx(&in[off]);
                                    if (has_nco) x = cmul(x, ph[bb][r]);
                                }
                            } else if (off >= -hist_len) {
                                x = ld_cx(&hist[off + hist_len]);  // already mixed
                            }
                        }
                        v[bb][r] = x;
                        if (has_nco) ph[bb][r] = cmul(ph[bb][r], rotG);
                    }
            }
            plan.p1_forward(smg, t, v);
            __syncthreads();
            plan.template p2<+1>(smg, t);
            __syncthreads();
#pragma unroll
            for (int b = 0; b < B3; ++b) {
                const int u = B3 * t + b;
                cx<T> v3[R3];
                P::p3_load(smg, u, v3);
                dft_regs<R3, +1, T>(v3);
                if (active) {
#pragma unroll
                    for (int q = 0; q < Q; ++q) {
                        const cx<T>* gt = gtab + ((long long)q * Pd + p) * K + u;
#pragma unroll
                        for (int k = 0; k < R3; ++k) {
                            const cx<T> w = ld_cx(&gt[k * KR]);
                            acc[q][b][k].x = fma(v3[k].x, w.x, fma(-v3[k].y, w.y, acc[q][b][k].x));
                            acc[q][b][k].y = fma(v3[k].x, w.y, fma(v3[k].y, w.x, acc[q][b][k].y));
                        }
                    }
                }
            }
            __syncthreads();  // pass-3 reads done before the next round's pass-1 writes
        }

        // ---- reduce the G partial spectra of this block, keep them per (block, phase)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
#pragma unroll
            for (int b = 0; b < B3; ++b)
#pragma unroll
                for (int k = 0; k < R3; ++k) st_cx(&W[(k * KR + B3 * t + b) * G + g], acc[q][b][k]);
            __syncthreads();
            cx<T>* ys = Ysave + ((long long)(blk - blk0) * Q + q) * K;
            for (int j = tid; j < K; j += THREADS) {
                cx<T> sum = ld_cx(&W[j * G]);
#pragma unroll
                for (int gg = 1; gg < G; ++gg) sum = sum + ld_cx(&W[j * G + gg]);
                st_cx(&ys[j], sum);
            }
            __syncthreads();
        }
    }

    // ---- inverse transforms: group g takes job j0 + g, job = (block, phase) -----
    const int njobs = (blk1 - blk0) * Q;
    for (int j0 = 0; j0 < njobs; j0 += G) {
        const int job = j0 + g;
        const bool activej = job < njobs;
#pragma unroll
        for (int b = 0; b < B3; ++b) {
            const int u = B3 * t + b;
            cx<T> v3[R3];
#pragma unroll
            for (int k = 0; k < R3; ++k) v3[k] = activej ? ld_cx(&Ysave[(long long)job * K + k * KR + u]) : cx<T>((T)0, (T)0);
            dft_regs<R3, -1, T>(v3);
            P::p3_store(smg, u, v3);
        }
        __syncthreads();
        plan.template p2<-1>(smg, t);
        __syncthreads();
        cx<T> v[B1][R1];
        plan.p1_inverse(smg, t, v);
        if (activej) {
            const int blk = blk0 + job / Q;
            const int q = job % Q;
            const long long Ibase = a.I_lo + (long long)blk * a.V - a.Lmax;
#pragma unroll
            for (int bb = 0; bb < B1; ++bb)
#pragma unroll
                for (int r = 0; r < R1; ++r) {
                    const int i = B1 * t + bb + S1 * r;
                    if (i >= a.Lmax && i < a.Lmax + a.V) {
                        const long long m = (Ibase + i) * a.Q + q;
                        if (m >= a.m_lo && m <= a.m_hi) st_cx(&out[m - a.m0 - 1], v[bb][r]);
                    }
                }
        }
        __syncthreads();
    }
}

template <typename T, int K, int Q, int G>
cudaError_t launch_poly_n(int n_streams, const PolyArgs<T>& a, cudaStream_t st) {
    using C = PolyCfg<T, K, G>;
    using P = typename C::Plan;
    const size_t smem = sizeof(cx<T>) * ((size_t)P::SMEM_ELEMS + (size_t)a.nbpc * Q * K);
    auto kern = k_poly<T, K, Q, G>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nsb = (a.n_blocks + a.nbpc - 1) / a.nbpc;
    kern<<<dim3((unsigned)nsb, (unsigned)n_streams), C::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace rr

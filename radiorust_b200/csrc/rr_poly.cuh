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
// Launch shape: one CTA = G thread groups of NT threads; in every round group g
// transforms branch p0 + g, so G neighbouring lanes read G consecutive input
// samples.  The G transforms are interleaved in shared memory (element e of
// transform g at G*sidx(e) + g: conflict-free for every pass).  Spectral
// accumulators live in registers across the ceil(P/G) rounds, are reduced over
// the groups through shared memory and parked there per block; at the end the
// groups run the inverse transforms of different (block, phase) jobs at once.
//
// NCO: the phasor of sample (i, p) of a block factors into phb[i] (depends on
// the row only: a K-entry shared-memory table per stream) times c_p (one scalar
// per transform, applied to the K/NT spectrum values each thread owns before
// the multiply-accumulate).  c_p is re-derived exactly from the integer phase
// recurrence (transform.rs:333-338) every 16 rounds and rotated in between.
//
// Per round: global loads of the NEXT round are issued right after pass 1 has
// parked the current data in shared memory, so their latency hides behind
// passes 2 and 3; two work buffers alternate so a round needs two barriers.
#pragma once
#include "rr_chain_os.cuh"

namespace rr {

template <typename T, int K, int G> struct PolyCfg {
    using Plan = typename PlanFor<T, K, 1>::type;
    static constexpr int NT = Plan::NT;
    static constexpr int THREADS = NT * G;
    static constexpr int R1 = Plan::R1, R2 = Plan::R2, R3 = Plan::R3;
    static constexpr int B1 = Plan::B1, B2 = Plan::B2, B3 = Plan::B3;
    static constexpr int S1 = Plan::S1, S2 = Plan::S2;
    // one pad element per R3 run keeps every pass of the interleaved layout at the minimum number of
    // shared-memory wavefronts for G = 8 and G = 10 (checked by enumeration)
    static constexpr int PAD = 1, LOG_R3 = Plan::LOG_R3;
    static constexpr int ROW = K + PAD * (K / R3);  // padded elements per transform
    static constexpr int WORK = G * ROW;            // one interleaved work buffer
    // element strides (in complex elements) inside the interleaved layout
    static constexpr int STR1 = G * (S1 + PAD * (S1 / R3));
    static constexpr int STR2 = G * (S2 + PAD);
    static constexpr int STR3 = G;
    __host__ __device__ static constexpr int sidx(int p) { return G * (p + PAD * (p >> LOG_R3)); }
    // two alternating work buffers save one barrier per round; f64 keeps one (shared-memory budget)
    static constexpr int NBUF = (sizeof(T) == 4) ? 2 : 1;
    // one CTA per SM: two would need <= 64 registers per thread, which spills the butterflies
    // (measured: 105 vs 185 GS/s on config 3)
    static constexpr int MIN_CTAS = 1;
    // shared memory: work buffers, 3 K-entry tables, nbpc*Q parked spectra
    __host__ __device__ static constexpr size_t smem_elems(int nbpc, int Q) {
        return (size_t)NBUF * WORK + 3 * (size_t)K + (size_t)nbpc * Q * K;
    }
};

// exact NCO phasor for push-relative sample offset `off` (may be negative):
// table index k = (idx + off) mod denom, i_k = numer*k mod denom (transform.rs:333-338)
template <typename T>
__device__ __forceinline__ cx<T> nco_phasor_at(long long off, uint32_t idx, uint32_t numer_abs, uint32_t denom, int sign, T start) {
    long long k = ((long long)idx + off) % (long long)denom;
    if (k < 0) k += denom;
    return nco_phasor<T>(mulmod_u32(numer_abs, (uint32_t)k, denom), denom, sign, start);
}
// exp(j * sign * 2*pi * (numer*d mod denom) / denom): the phase advance over d samples (start phase excluded)
template <typename T>
__device__ __forceinline__ cx<T> nco_rotation(long long d, uint32_t numer_abs, uint32_t denom, int sign) {
    long long k = d % (long long)denom;
    if (k < 0) k += denom;
    const uint32_t r = mulmod_u32(numer_abs, (uint32_t)k, denom);
    double s, c;
    sincospi(2.0 * (double)r / (double)denom, &s, &c);
    return cx<T>((T)c, (T)(sign < 0 ? -s : s));
}

// (a * b) mod m for a, b < m < 2^31 without a 64-bit division: the quotient estimate from the
// double reciprocal is off by at most one
__device__ __forceinline__ uint32_t mulmod_fast(uint32_t a, uint32_t b, uint32_t m, double inv_m) {
    const unsigned long long prod = (unsigned long long)a * b;
    const unsigned long long q = (unsigned long long)((double)prod * inv_m);
    long long r = (long long)(prod - q * m);
    if (r < 0) r += m;
    else if (r >= (long long)m) r -= m;
    return (uint32_t)r;
}

// two adjacent complex values with the widest load available
__device__ __forceinline__ void ld_pair(const cx<float>* p, cx<float>& a, cx<float>& b) { ld_cx2(p, a, b); }
__device__ __forceinline__ void ld_pair(const cx<double>* p, cx<double>& a, cx<double>& b) {
    a = ld_cx(p);
    b = ld_cx(p + 1);
}

template <typename T, int K, int Q, int G>
__global__ void __launch_bounds__(PolyCfg<T, K, G>::THREADS, PolyCfg<T, K, G>::MIN_CTAS)
k_poly(const PolyArgs<T> a) {
    using C = PolyCfg<T, K, G>;
    constexpr int NT = C::NT, R1 = C::R1, R2 = C::R2, R3 = C::R3, B1 = C::B1, B2 = C::B2, B3 = C::B3;
    constexpr int S1 = C::S1, S2 = C::S2, THREADS = C::THREADS, KR = K / R3;
    constexpr int STR1 = C::STR1, STR2 = C::STR2, STR3 = C::STR3;
    static_assert(R1 % 2 == 0 && R2 % 2 == 0, "table rows are read in pairs");
    const int tid = threadIdx.x;
    const int g = tid % G, t = tid / G;
    const int s = blockIdx.y;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* W0 = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* tw1s = W0 + C::NBUF * C::WORK;  // [NT*B1][R1]  W_K^(k*q)
    cx<T>* tw2s = tw1s + K;                // [NT*B2][R2]  W_(R2*R3)^(k*n3)
    cx<T>* phb = tw2s + K;                 // [NT*B1][R1]  NCO rotation over i*P samples, i = q + S1*k
    cx<T>* Ysave = phb + K;                // [nbpc*Q][K]

    const cx<T>* __restrict__ in = reinterpret_cast<const cx<T>*>(a.in) + (long long)s * a.in_stride;
    const cx<T>* __restrict__ hist_end = reinterpret_cast<const cx<T>*>(a.hist2) + ((long long)s + 1) * 2 * a.n;
    const cx<T>* __restrict__ gtab = reinterpret_cast<const cx<T>*>(a.gtab);
    const cx<T>* __restrict__ twK = reinterpret_cast<const cx<T>*>(a.twK);
    cx<T>* __restrict__ out = reinterpret_cast<cx<T>*>(a.out) + (long long)s * a.out_stride;
    const int Pd = (int)a.P;
    const int NR = (Pd + G - 1) / G;  // rounds per block
    const long long len = a.len, hist_len = 2 * a.n;
    const long long row_stride = (long long)S1 * Pd;  // elements between the rows of one pass-1 butterfly

    // ---- NCO constants ---------------------------------------------------------
    const bool has_nco = (a.nco != nullptr);
    uint32_t denom = 1, numer_abs = 0;
    int sign = 0;
    T start = (T)0;
    double inv_denom = 1.0;
    cx<T> rotG((T)1, (T)0);
    uint32_t kblk = 0, kstep_blk = 0;  // table index of sample (i = 0, p = g) of the current block; its advance per block
    if (has_nco) {
        const NcoStream ns = a.nco[s];
        denom = ns.denom;
        numer_abs = ns.numer_abs;
        sign = ns.sign;
        start = (T)ns.start_phase;
        inv_denom = 1.0 / (double)denom;
        rotG = nco_rotation<T>(G, numer_abs, denom, sign);
        kstep_blk = (uint32_t)(((long long)a.V * Pd) % (long long)denom);
    }

    // ---- per-thread invariant shared-memory offsets (elements, g included) -------
    int off1[B1], off2[B2], off3[B3];
#pragma unroll
    for (int b = 0; b < B1; ++b) off1[b] = C::sidx(B1 * t + b) + g;
#pragma unroll
    for (int b = 0; b < B2; ++b) {
        const int id = B2 * t + b;
        off2[b] = C::sidx((id >> C::LOG_R3) * S1 + (id & (R3 - 1))) + g;
    }
#pragma unroll
    for (int b = 0; b < B3; ++b) off3[b] = C::sidx((B3 * t + b) * R3) + g;

    // ---- tables: twiddles of this thread's butterflies, NCO row phasors ------------
    if (g == 0) {
#pragma unroll
        for (int b = 0; b < B1; ++b) {
            cx<T> p[R1];
            pow_chain<R1, T>(ld_cx(&twK[B1 * t + b]), p);
#pragma unroll
            for (int k = 0; k < R1; ++k) st_cx(&tw1s[(B1 * t + b) * R1 + k], p[k]);
        }
#pragma unroll
        for (int b = 0; b < B2; ++b) {
            cx<T> p[R2];
            pow_chain<R2, T>(ld_cx(&twK[((B2 * t + b) & (R3 - 1)) * R1]), p);
#pragma unroll
            for (int k = 0; k < R2; ++k) st_cx(&tw2s[(B2 * t + b) * R2 + k], p[k]);
        }
    }
    for (int j = tid; j < K; j += THREADS) {
        const int i = (j / R1) + S1 * (j % R1);  // slot (q, k) holds row i = q + S1*k
        st_cx(&phb[j], has_nco ? nco_rotation<T>((long long)i * Pd, numer_abs, denom, sign) : cx<T>((T)1, (T)0));
    }
    __syncthreads();

    const int blk0 = blockIdx.x * a.nbpc;
    const int blk1 = min(blk0 + a.nbpc, a.n_blocks);

    // ---- per-block window: element (i, p) of block blk sits at push offset boff + i*P + p ------
    long long boff = 0;
    bool interior = false;
    int lo = 0, hi = 0, hlo = 0;  // window-relative bounds: [lo, hi) pushed samples, [hlo, lo) history
    const cx<T>* hbase = hist_end;  // history sample of window element rel at hbase[rel]
    auto set_block = [&](int blk) {
        const long long Ibase = a.I_lo + (long long)blk * a.V - a.Lmax;
        boff = Ibase * Pd - a.J0 - Pd;
        const long long span = (long long)K * Pd;
        interior = (boff >= 0) && (boff + span <= len);
        const long long l = boff < 0 ? -boff : 0;
        const long long h = len - boff;
        lo = (int)(l < span ? l : span);
        hi = (int)(h < 0 ? 0 : (h < span ? h : span));
        const long long hl = l - hist_len;
        hlo = (int)(hl < 0 ? 0 : (hl < span ? hl : span));
        hbase = hist_end - l;
    };
    // raw load of round r of the current block for this thread's pass-1 butterflies
    cx<T> v[B1][R1];
    auto load_round = [&](int r) {
        const int p = r * G + g;
        const bool active = p < Pd;
        const cx<T>* bin = in + boff;
        if (interior) {
#pragma unroll
            for (int b = 0; b < B1; ++b) {
                const cx<T>* q = bin + (long long)(B1 * t + b) * Pd + p;
#pragma unroll
                for (int k = 0; k < R1; ++k) {
                    v[b][k] = active ? ld_cx(q) : cx<T>((T)0, (T)0);
                    q += row_stride;
                }
            }
        } else {
#pragma unroll
            for (int b = 0; b < B1; ++b) {
                int rel = (B1 * t + b) * Pd + p;
#pragma unroll
                for (int k = 0; k < R1; ++k) {
                    cx<T> x((T)0, (T)0);
                    if (active) {
                        if (rel >= lo) {
                            if (rel < hi) x = ld_cx(&bin[rel]);
                        } else if (rel >= hlo) {
                            x = ld_cx(&hbase[rel]);  // history: already mixed
                        }
                    }
                    v[b][k] = x;
                    rel += (int)row_stride;
                }
            }
        }
    };

    if (blk0 < blk1) {
        set_block(blk0);
        if (has_nco) {
            long long k0 = ((long long)a.nco[s].idx + boff + g) % (long long)denom;
            if (k0 < 0) k0 += denom;
            kblk = (uint32_t)k0;
        }
        load_round(0);
    }

    for (int blk = blk0; blk < blk1; ++blk) {
        cx<T> acc[Q][B3][R3];
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int b = 0; b < B3; ++b)
#pragma unroll
                for (int k = 0; k < R3; ++k) acc[q][b][k] = cx<T>((T)0, (T)0);

        const bool cur_interior = interior;
        const int cur_lo = lo;
        cx<T> cp((T)1, (T)0);  // NCO phasor of sample (i = 0, p) of this block
        for (int r = 0; r < NR; ++r) {
            const int p = r * G + g;
            cx<T>* W = W0 + (C::NBUF == 2 ? (r & 1) : 0) * C::WORK;
            if (has_nco && (r & 15) == 0) {
                const uint32_t k = addmod_u32(kblk, (uint32_t)(r * G) % denom, denom);
                cp = nco_phasor<T>(mulmod_fast(numer_abs, k, denom, inv_denom), denom, sign, start);
            }

            // ---- NCO row part + pass 1 ------------------------------------------------
            if (has_nco) {
                if (cur_interior) {
#pragma unroll
                    for (int b = 0; b < B1; ++b)
#pragma unroll
                        for (int k = 0; k < R1; k += 2) {
                            cx<T> f0, f1;
                            ld_pair(&phb[(B1 * t + b) * R1 + k], f0, f1);
                            v[b][k] = cmul(v[b][k], f0);
                            v[b][k + 1] = cmul(v[b][k + 1], f1);
                        }
                } else {
                    // history samples are already mixed: cancel the c_p that the spectrum gets later
#pragma unroll
                    for (int b = 0; b < B1; ++b) {
                        int rel = (B1 * t + b) * Pd + p;
#pragma unroll
                        for (int k = 0; k < R1; ++k) {
                            v[b][k] = (rel >= cur_lo) ? cmul(v[b][k], ld_cx(&phb[(B1 * t + b) * R1 + k])) : cmulc(v[b][k], cp);
                            rel += (int)row_stride;
                        }
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < B1; ++b) {
                dft_regs<R1, +1, T>(v[b]);
                cx<T> f[R1];
#pragma unroll
                for (int k = 0; k < R1; k += 2) ld_pair(&tw1s[(B1 * t + b) * R1 + k], f[k], f[k + 1]);
                st_cx(&W[off1[b]], v[b][0]);
#pragma unroll
                for (int k = 1; k < R1; ++k) st_cx(&W[off1[b] + k * STR1], cmul(v[b][k], f[k]));
            }
            // ---- prefetch: next round's samples (of the next block after the last round), this
            //      round's table entries ---------------------------------------------------
            if (r + 1 < NR) {
                load_round(r + 1);
            } else if (blk + 1 < blk1) {
                set_block(blk + 1);
                load_round(0);
            }
            cx<T> gt0[B3][R3];
            {
                const cx<T>* gp = gtab + ((long long)r * K) * G + g;
#pragma unroll
                for (int b = 0; b < B3; ++b)
#pragma unroll
                    for (int k = 0; k < R3; ++k) gt0[b][k] = ld_cx(&gp[(k * KR + B3 * t + b) * G]);
            }
            __syncthreads();
            // ---- pass 2 -------------------------------------------------------------------
#pragma unroll
            for (int b = 0; b < B2; ++b) {
                cx<T> w[R2], f[R2];
#pragma unroll
                for (int k = 0; k < R2; ++k) w[k] = ld_cx(&W[off2[b] + k * STR2]);
#pragma unroll
                for (int k = 0; k < R2; k += 2) ld_pair(&tw2s[(B2 * t + b) * R2 + k], f[k], f[k + 1]);
                dft_regs<R2, +1, T>(w);
                st_cx(&W[off2[b]], w[0]);
#pragma unroll
                for (int k = 1; k < R2; ++k) st_cx(&W[off2[b] + k * STR2], cmul(w[k], f[k]));
            }
            __syncthreads();
            // ---- pass 3 + NCO transform part + multiply-accumulate ---------------------------
#pragma unroll
            for (int b = 0; b < B3; ++b) {
                cx<T> w[R3];
#pragma unroll
                for (int k = 0; k < R3; ++k) w[k] = ld_cx(&W[off3[b] + k * STR3]);
                dft_regs<R3, +1, T>(w);
                if (has_nco) {
#pragma unroll
                    for (int k = 0; k < R3; ++k) w[k] = cmul(w[k], cp);
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) {
#pragma unroll
                    for (int k = 0; k < R3; ++k) {
                        cx<T> h;
                        if (q == 0) h = gt0[b][k];
                        else h = ld_cx(&gtab[(((long long)q * NR + r) * K + k * KR + B3 * t + b) * G + g]);
                        acc[q][b][k].x = fma(w[k].x, h.x, fma(-w[k].y, h.y, acc[q][b][k].x));
                        acc[q][b][k].y = fma(w[k].x, h.y, fma(w[k].y, h.x, acc[q][b][k].y));
                    }
                }
            }
            if (has_nco) cp = cmul(cp, rotG);
            // two buffers: the next round works in the other one, and two barriers separate this
            // round's reads from the round after next
            if (C::NBUF == 1) __syncthreads();
        }
        if (has_nco) kblk = addmod_u32(kblk, kstep_blk, denom);
        __syncthreads();

        // ---- reduce the G partial spectra of this block, park them per (block, phase) ------
#pragma unroll
        for (int q = 0; q < Q; ++q) {
#pragma unroll
            for (int b = 0; b < B3; ++b)
#pragma unroll
                for (int k = 0; k < R3; ++k) st_cx(&W0[(k * KR + B3 * t + b) * G + g], acc[q][b][k]);
            __syncthreads();
            cx<T>* ys = Ysave + ((long long)(blk - blk0) * Q + q) * K;
            for (int j = tid; j < K; j += THREADS) {
                cx<T> sum = ld_cx(&W0[j * G]);
#pragma unroll
                for (int gg = 1; gg < G; ++gg) sum = sum + ld_cx(&W0[j * G + gg]);
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
        cx<T>* W = W0;
#pragma unroll
        for (int b = 0; b < B3; ++b) {
            cx<T> w[R3];
#pragma unroll
            for (int k = 0; k < R3; ++k) w[k] = activej ? ld_cx(&Ysave[(long long)job * K + k * KR + B3 * t + b]) : cx<T>((T)0, (T)0);
            dft_regs<R3, -1, T>(w);
#pragma unroll
            for (int k = 0; k < R3; ++k) st_cx(&W[off3[b] + k * STR3], w[k]);
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < B2; ++b) {
            cx<T> w[R2];
            w[0] = ld_cx(&W[off2[b]]);
#pragma unroll
            for (int k = 1; k < R2; ++k) w[k] = cmulc(ld_cx(&W[off2[b] + k * STR2]), ld_cx(&tw2s[(B2 * t + b) * R2 + k]));
            dft_regs<R2, -1, T>(w);
#pragma unroll
            for (int k = 0; k < R2; ++k) st_cx(&W[off2[b] + k * STR2], w[k]);
        }
        __syncthreads();
        const int blk = blk0 + job / Q;
        const int q = job % Q;
        const long long Ibase = a.I_lo + (long long)blk * a.V - a.Lmax;
#pragma unroll
        for (int b = 0; b < B1; ++b) {
            cx<T> w[R1];
            w[0] = ld_cx(&W[off1[b]]);
#pragma unroll
            for (int k = 1; k < R1; ++k) w[k] = cmulc(ld_cx(&W[off1[b] + k * STR1]), ld_cx(&tw1s[(B1 * t + b) * R1 + k]));
            dft_regs<R1, -1, T>(w);
            if (activej) {
#pragma unroll
                for (int k = 0; k < R1; ++k) {
                    const int i = B1 * t + b + S1 * k;
                    if (i >= a.Lmax && i < a.Lmax + a.V) {
                        const long long m = (Ibase + i) * a.Q + q;
                        if (m >= a.m_lo && m <= a.m_hi) st_cx(&out[m - a.m0 - 1], w[k]);
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int K, int Q, int G>
cudaError_t launch_poly_n(int n_streams, const PolyArgs<T>& a, cudaStream_t st) {
    using C = PolyCfg<T, K, G>;
    const size_t smem = sizeof(cx<T>) * C::smem_elems(a.nbpc, Q);
    auto kern = k_poly<T, K, Q, G>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nsb = (a.n_blocks + a.nbpc - 1) / a.nbpc;
    kern<<<dim3((unsigned)nsb, (unsigned)n_streams), C::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace rr

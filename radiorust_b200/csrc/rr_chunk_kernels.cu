// FmMod (the TX-side phase accumulator) and the chunk reorganisers (Rechunker / Overlapper) on the device, so
// that chains with mismatched chunk lengths and the transmit direction stay device-resident.
// sm_100a.
#include "rr_chain_os.cuh"

namespace rr {

// ---------------------------------------------------------------------------
// FmMod, src/blocks/modulation.rs:46-51:
//     current_phase += sample.re * factor;  current_phase %= TAU;  out = (cos, sin)(current_phase)
// a serial recurrence in Flt whose rounding the reference fixes step by step (product rounded, sum rounded,
// exact fmod), so the phase is carried by ONE thread per stream in exactly that order; everything around it
// (coalesced loads, the products, sincos, coalesced stores) is done by the whole CTA on a shared-memory tile.
// A CTA takes SPC streams; the launcher picks SPC so that the grid covers the SMs about three times and the
// serial chains of the CTAs sharing an SM overlap each other's load/store phases.  The bound is the latency of
// the recurrence (add, two compares, two fused multiply-adds: ~35 cycles per sample and stream with the shared-memory
// traffic), not HBM: 4096 streams run at ~135 GS/s on a B200 (33 % of the 16 B/sample HBM figure), one stream at 54 MS/s.
// ---------------------------------------------------------------------------
template <typename T> struct FmConst;
template <> struct FmConst<float> {
    static __device__ __forceinline__ float tau() { return 6.2831855f; }  // f32::TAU
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }  // never contracted
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float rem(float a, float m) { return fabsf(a) >= m ? fmodf(a, m) : a; }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ void sc(float p, float* s, float* c) { sincosf(p, s, c); }
};
template <> struct FmConst<double> {
    static __device__ __forceinline__ double tau() { return 6.283185307179586476925286766559; }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double rem(double a, double m) { return fabs(a) >= m ? fmod(a, m) : a; }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ void sc(double p, double* s, double* c) { sincos(p, s, c); }
};

constexpr int kFmTileElems = 2048;  // samples of all SPC streams per tile

template <typename T, int SPC>
__global__ void __launch_bounds__(256) k_fmmod(const cx<T>* __restrict__ in, long long in_stride, cx<T>* __restrict__ out,
                                               long long out_stride, long long len, int n_streams, T* __restrict__ phase,
                                               T factor) {
    constexpr int TL = kFmTileElems / SPC;
    constexpr int PITCH = TL + 1;  // odd: the SPC serial threads walk distinct banks
    __shared__ T buf[SPC * PITCH];
    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * SPC;
    const int sl = tid;  // thread sl < SPC carries the recurrence of stream s0 + sl
    const bool serial = sl < SPC && s0 + sl < n_streams;
    T ph = serial ? phase[s0 + sl] : (T)0;
    const T tau = FmConst<T>::tau();
    // products of the next tile travel in registers: their loads are issued before the serial phase of the current
    // tile and land in shared memory after its store phase, so the DRAM latency hides behind the recurrence
    constexpr int PER = kFmTileElems / 256;
    T nxt[PER];
    auto fetch = [&](long long t0) {
        const long long left = len - t0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int idx = tid + 256 * u, s = idx / TL, k = idx - s * TL;
            nxt[u] = (T)0;
            if (s0 + s < n_streams && k < left) nxt[u] = ld_cx(&in[(long long)(s0 + s) * in_stride + t0 + k]).x;
        }
    };
    auto park = [&]() {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int idx = tid + 256 * u, s = idx / TL, k = idx - s * TL;
            buf[s * PITCH + k] = FmConst<T>::mul(nxt[u], factor);
        }
    };
    fetch(0);
    park();
    for (long long t0 = 0; t0 < len; t0 += TL) {
        const int tl = (int)((len - t0) < (long long)TL ? (len - t0) : (long long)TL);
        if (t0 + TL < len) fetch(t0 + TL);
        __syncthreads();
        if (serial) {
            // Critical path per sample: add -> compares -> two fused multiply-adds, no branch, no predicate.  For |a| < 2 TAU the exact
            // fmod is a itself, a - TAU or a + TAU (the subtraction is exact, Sterbenz); a batch of 8 runs on that
            // assumption and is redone with fmod when any of its sums was larger.  Shared-memory traffic is batched.
            T* row = buf + sl * PITCH;
            int k = 0;
            for (; k + 8 <= tl; k += 8) {
                T v[8], r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = row[k + u];
                T q = ph;
                bool big = false;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const T a = FmConst<T>::add(q, v[u]);
                    big |= !(FmConst<T>::abs(a) < tau + tau);  // also catches NaN/inf
                    // a - TAU*[a >= TAU] + TAU*[a <= -TAU] with 0/1 factors in registers (no predicate on the chain);
                    // the products are exact and so is the sum (Sterbenz), whichever way it is contracted
                    const T c1 = a >= tau ? (T)1 : (T)0, c2 = a <= -tau ? (T)1 : (T)0;
                    q = FmConst<T>::fma(c2, tau, FmConst<T>::fma(c1, -tau, a));
                    r[u] = q;
                }
                if (big) {
                    q = ph;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        q = FmConst<T>::rem(FmConst<T>::add(q, v[u]), tau);
                        r[u] = q;
                    }
                }
                ph = q;
#pragma unroll
                for (int u = 0; u < 8; ++u) row[k + u] = r[u];
            }
            for (; k < tl; ++k) {
                ph = FmConst<T>::rem(FmConst<T>::add(ph, row[k]), tau);
                row[k] = ph;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < SPC * TL; idx += 256) {
            const int s = idx / TL, k = idx - s * TL;
            if (s0 + s < n_streams && k < tl) {
                T sn, cs;
                FmConst<T>::sc(buf[s * PITCH + k], &sn, &cs);
                st_cx(&out[(long long)(s0 + s) * out_stride + t0 + k], cx<T>(cs, sn));
            }
        }
        __syncthreads();
        if (t0 + TL < len) park();
    }
    if (serial) phase[s0 + sl] = ph;
}

template <typename T, int SPC>
static cudaError_t fmmod_spc(const void* in, long long in_stride, void* out, long long out_stride, long long len, int S, void* phase,
                             double factor, cudaStream_t st) {
    k_fmmod<T, SPC><<<(S + SPC - 1) / SPC, 256, 0, st>>>(reinterpret_cast<const cx<T>*>(in), in_stride, reinterpret_cast<cx<T>*>(out),
                                                         out_stride, len, S, reinterpret_cast<T*>(phase), (T)factor);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_fmmod(const void* in, long long in_stride, void* out, long long out_stride, long long len, int n_streams, void* phase,
                         double factor, int sm_count, cudaStream_t st) {
    if (len <= 0 || n_streams <= 0) return cudaSuccess;
    int spc = 32;
    while (spc > 1 && (n_streams + spc - 1) / spc < 3 * sm_count) spc >>= 1;  // measured: ~3.5 CTAs per SM is the optimum
    switch (spc) {
        case 32: return fmmod_spc<T, 32>(in, in_stride, out, out_stride, len, n_streams, phase, factor, st);
        case 16: return fmmod_spc<T, 16>(in, in_stride, out, out_stride, len, n_streams, phase, factor, st);
        case 8: return fmmod_spc<T, 8>(in, in_stride, out, out_stride, len, n_streams, phase, factor, st);
        case 4: return fmmod_spc<T, 4>(in, in_stride, out, out_stride, len, n_streams, phase, factor, st);
        case 2: return fmmod_spc<T, 2>(in, in_stride, out, out_stride, len, n_streams, phase, factor, st);
        default: return fmmod_spc<T, 1>(in, in_stride, out, out_stride, len, n_streams, phase, factor, st);
    }
}

// ---------------------------------------------------------------------------
// Overlapper, src/blocks/chunks.rs:203-225: output chunk j = chunks [j, j + k) of the sequence
// [history | pushed], i.e. out[j*span + t] = seq[base + j*hop + t], t < span.  The same gather with hop = 0
// writes the next history.  seq = a (a_len samples per stream) followed by b.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_overlap(const cx<T>* __restrict__ a, long long a_stride, long long a_len,
                                                 const cx<T>* __restrict__ b, long long b_stride, cx<T>* __restrict__ out,
                                                 long long out_stride, long long n_out, long long span, long long hop, long long base,
                                                 int n_streams) {
    // blockIdx.z walks the streams, blockIdx.y the output chunks, blockIdx.x the chunk; 4 loads in flight per thread
    for (long long s = blockIdx.z; s < n_streams; s += gridDim.z) {
        const cx<T>* sa = a + s * a_stride;
        const cx<T>* sb = b + s * b_stride - a_len;
        for (long long j = blockIdx.y; j < n_out; j += gridDim.y) {
            cx<T>* dst = out + s * out_stride + j * span;
            const long long q0 = base + j * hop;
            const long long tile = (long long)blockDim.x * 4;
            for (long long t0 = (long long)blockIdx.x * tile + threadIdx.x; t0 < span; t0 += (long long)gridDim.x * tile) {
                cx<T> v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long t = t0 + (long long)u * blockDim.x, q = q0 + t;
                    if (t < span) v[u] = q < a_len ? ld_cx(&sa[q]) : ld_cx(&sb[q]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long t = t0 + (long long)u * blockDim.x;
                    if (t < span) st_cx(&dst[t], v[u]);
                }
            }
        }
    }
}

template <typename T>
cudaError_t launch_overlap(const void* a, long long a_stride, long long a_len, const void* b, long long b_stride, void* out,
                           long long out_stride, long long n_out, long long span, long long hop, long long base, int n_streams,
                           cudaStream_t st) {
    const long long total = n_out * span;
    if (total <= 0 || n_streams <= 0) return cudaSuccess;
    long long bx = (span + 4095) / 4096;  // four rounds of 1024 samples per CTA
    if (bx > 64) bx = 64;
    const long long by = n_out < 64 ? n_out : 64, bz = n_streams < 16384 ? n_streams : 16384;
    k_overlap<T><<<dim3((unsigned)bx, (unsigned)by, (unsigned)bz), 256, 0, st>>>(reinterpret_cast<const cx<T>*>(a), a_stride, a_len,
                                                                    reinterpret_cast<const cx<T>*>(b), b_stride,
                                                                    reinterpret_cast<cx<T>*>(out), out_stride, n_out, span, hop, base,
                                                                    n_streams);
    return cudaGetLastError();
}

#define RR_INST(T)                                                                                                              \
    template cudaError_t launch_fmmod<T>(const void*, long long, void*, long long, long long, int, void*, double, int, cudaStream_t); \
    template cudaError_t launch_overlap<T>(const void*, long long, long long, const void*, long long, void*, long long, long long,      \
                                           long long, long long, long long, int, cudaStream_t);
RR_INST(float)
RR_INST(double)
#undef RR_INST

}  // namespace rr

// Overlap-save fast convolution for long Filters, 2n = 2^15 .. 2^20 points (BASELINE configs 2 and 4: f32, n = 65536;
// config 5: f64, 2^20 points).  Replaces the two rustfft calls and the H multiply of src/blocks/filters.rs:244-252.
//
// Four-step FFT, N = Na x 1024, n = n1*1024 + n2, k = k1 + Na*k2, as three STREAMING kernels (every sample is loaded
// once and stored once per kernel, in full 64/128-byte segments; every butterfly runs in registers):
//
//   k_long_cols_fwd : W adjacent columns n2 per CTA.  DFT_Na over n1 as two register passes (A x B points, ONE exchange
//                     through shared memory), times W_N^(n2*k1)                                   -> scratch[k1][n2]
//   k_long_rows     : one WARP per row k1: DFT_1024 over n2 as radix-32 x radix-32 with one exchange in the warp's own
//                     strip (no CTA barrier), times H[k1 + Na*k2], inverse the same way, in place -> scratch[k1][n2]
//   k_long_cols_inv : the mirror image of the first kernel, first half of the block only (filters.rs:253) -> out
//
// The column kernels interleave the W columns in shared memory (element (i, column) at i*W + column): lanes of one row
// touch consecutive words, so every access is conflict free whatever the butterfly's stride.  The row kernel's strip has
// an odd pitch (33 elements): both directions of its transposition are conflict free.  Twiddles W^(base*k), k < R, are
// built in registers from one or two table reads with multiplication depth <= 4 (geo_twiddle).  complex<f32> arithmetic
// runs on the packed two-wide instructions (rr_pk.cuh), complex<f64> on rr_complex.cuh's scalar forms.
//
// The scratch streams through HBM (11 point transfers per output sample).  The column kernels run at 71-75 % of the DRAM
// peak, the row kernel at 61 % with the fp32 pipe 66 % busy (two radix-32 passes each way: ~85 flop per point and
// direction); both floors lie within 1.5 x of each other, so no single change moves the path far.  Measured and not kept:
// launch groups whose scratch fits L2 going round side streams (commit 2f092e5), the three phases in one
// persistent kernel with teams of CTAs and team barriers (commit 71727ba), the row kernel's next row prefetched by a bulk
// copy (fewer warps fit; 854 -> 1020 us).  Other sizes keep rr_big_os.cu's kernels.  sm_100a.
#include <cstdlib>

#include "rr_kernels.h"
#include "rr_pk.cuh"

namespace rr {

namespace {

constexpr int kRowLen = 1024;  // Nb

template <typename T> struct CT;
template <> struct CT<float> {
    using C = pc;
    template <int R, int DIR> static __device__ __forceinline__ void dft(C (&v)[R]) { pdft_regs<R, DIR>(v); }
    static __device__ __forceinline__ C mul(C a, C w) { return pcmul(a, w); }
    static __device__ __forceinline__ C mulc(C a, C w) { return pcmulc(a, w); }
    static __device__ __forceinline__ C sqr(C a) { return pcmul(a, a); }
    static __device__ __forceinline__ C one() { return pc(1.f, 0.f); }
    static __device__ __forceinline__ C ld(const cx<float>* p) {
        const float2 v = *reinterpret_cast<const float2*>(p);
        return pc(v.x, v.y);
    }
    static __device__ __forceinline__ C ldg(const cx<float>* p) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p));
        return pc(v.x, v.y);
    }
    static __device__ __forceinline__ C ldcg(const cx<float>* p) {  // L2 only: the scratch is rewritten by other SMs
        const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
        return pc(v.x, v.y);
    }
    static __device__ __forceinline__ void st(cx<float>* p, C v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }
    static __device__ __forceinline__ void st_stream(cx<float>* p, C v) { __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }
};
template <> struct CT<double> {
    using C = cx<double>;
    template <int R, int DIR> static __device__ __forceinline__ void dft(C (&v)[R]) { dft_regs<R, DIR, double>(v); }
    static __device__ __forceinline__ C mul(C a, C w) { return cmul(a, w); }
    static __device__ __forceinline__ C mulc(C a, C w) { return cmulc(a, w); }
    static __device__ __forceinline__ C sqr(C a) { return csqr(a); }
    static __device__ __forceinline__ C one() { return cx<double>(1.0, 0.0); }
    static __device__ __forceinline__ C ld(const cx<double>* p) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        return cx<double>(v.x, v.y);
    }
    static __device__ __forceinline__ C ldg(const cx<double>* p) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(p));
        return cx<double>(v.x, v.y);
    }
    static __device__ __forceinline__ C ldcg(const cx<double>* p) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
        return cx<double>(v.x, v.y);
    }
    static __device__ __forceinline__ void st(cx<double>* p, C v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); }
    static __device__ __forceinline__ void st_stream(cx<double>* p, C v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }
};

// v[k] *= t0 * s^k (CONJ: times the conjugate of that), k = 0 .. R-1.  s^k = (s^4)^(k/4) * s^(k%4): at most twelve
// factors live, multiplication depth <= 4 (a few ulp).  UNIT: t0 == 1.
template <typename T, int R, bool CONJ, bool UNIT> __device__ __forceinline__ void geo_twiddle(typename CT<T>::C (&v)[R], typename CT<T>::C t0, typename CT<T>::C s) {
    using X = CT<T>;
    using C = typename X::C;
    constexpr int NLO = R < 4 ? R : 4, NHI = R / 4 > 0 ? R / 4 : 1;
    C lo[4], hi[8];
    static_assert(NHI <= 8, "radix up to 32");
    const C s2 = X::sqr(s);
    lo[0] = t0;
    lo[1] = UNIT ? s : X::mul(t0, s);
    lo[2] = UNIT ? s2 : X::mul(t0, s2);
    lo[3] = X::mul(lo[1], s2);
    if constexpr (NHI > 1) {
        hi[1] = X::sqr(s2);
        static_for<2, NHI>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int h = 1 << ilog2c(i), l = i - h;
            if constexpr (l == 0) hi[i] = X::sqr(hi[h / 2]);
            else hi[i] = X::mul(hi[h], hi[l]);
        });
    }
    static_for<0, R>([&](auto K) {
        constexpr int k = decltype(K)::value;
        constexpr int kl = k % NLO, kh = k / 4;
        if constexpr (!(UNIT && kl == 0)) v[k] = CONJ ? X::mulc(v[k], lo[kl]) : X::mul(v[k], lo[kl]);
        if constexpr (kh > 0) v[k] = CONJ ? X::mulc(v[k], hi[kh]) : X::mul(v[k], hi[kh]);
    });
}

// rows of padding after each group of B rows: with 64-byte rows the B-row stride of the second pass would put a warp's
// four rows on the same half of the banks
template <typename T, int W> constexpr int pad_rows() { return (W * 2 * (int)sizeof(T) >= 128) ? 0 : 1; }
template <typename T, int A, int B, int W> constexpr size_t cols_smem() { return (size_t)A * (B + pad_rows<T, W>()) * W * 2 * sizeof(T); }
constexpr int kThreads = 128;  // every kernel here: the column tiles walk their A*W / B*W tasks in groups of kThreads / W rows

// ---- one tile of W columns, forward: [prev | cur] (n samples each) -> dst[k1][n2] ---------------------------------
template <typename T, int A, int B, int W>
__device__ __forceinline__ void cols_fwd_tile(cx<T>* __restrict__ sm, const cx<T>* __restrict__ prev, const cx<T>* __restrict__ cur, int tile,
                                              cx<T>* __restrict__ dst_block, const cx<T>* __restrict__ twN, const cx<T>* __restrict__ twA,
                                              const cx<T>* __restrict__ twC) {
    using X = CT<T>;
    using C = typename X::C;
    constexpr int Na = A * B, Nb = kRowLen, NT = kThreads, GR = NT / W, BP = B + pad_rows<T, W>();
    static_assert(GR >= 1 && A % GR == 0 && B % GR == 0, "row groups must divide both passes");
    const int c = threadIdx.x % W, g = threadIdx.x / W;
    const long long n = (long long)Na * Nb / 2;
    const int n2 = tile * W + c;
    // pass 1: rows n1 = a*B + bi of this column, DFT over a -> ka, times W_Na^(bi*ka)
#pragma unroll
    for (int it = 0; it < B / GR; ++it) {
        const int bi = g + it * GR;
        C v[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const long long i = (long long)(a * B + bi) * Nb + n2;
            v[a] = (a < A / 2) ? X::ldg(&prev[i]) : X::ldg(&cur[i - n]);  // n1 < Na/2 <=> a < A/2
        }
        X::template dft<A, +1>(v);
        geo_twiddle<T, A, false, true>(v, X::one(), X::ldg(&twA[bi]));
#pragma unroll
        for (int ka = 0; ka < A; ++ka) X::st(&sm[(ka * BP + bi) * W + c], v[ka]);
    }
    __syncthreads();
    // pass 2: DFT over bi -> kb: bin k1 = ka + A*kb, times the four-step twiddle W_N^(n2*k1)
    // W_N^(n2*k) = W_N^(tile*W*k) (a warp-uniform read of the N-point table) * W_N^(c*k) (twC: [k][c], coalesced): a
    // gathered read of twN[n2*k] costs a wavefront per lane
    cx<T>* dst = dst_block + n2;
    const C tw_step = X::mul(X::ldg(&twN[(long long)tile * W * A]), X::ldg(&twC[A * W + c]));
#pragma unroll
    for (int it = 0; it < A / GR; ++it) {
        const int ka = g + it * GR;
        C u[B];
#pragma unroll
        for (int bi = 0; bi < B; ++bi) u[bi] = X::ld(&sm[(ka * BP + bi) * W + c]);
        X::template dft<B, +1>(u);
        geo_twiddle<T, B, false, false>(u, X::mul(X::ldg(&twN[(long long)tile * W * ka]), X::ldg(&twC[ka * W + c])), tw_step);
#pragma unroll
        for (int kb = 0; kb < B; ++kb) X::st(&dst[(long long)(ka + A * kb) * Nb], u[kb]);
    }
}

// ---- one tile of W columns, inverse: src[k1][n2] -> dst[n1*Nb + n2], n1 < Na/2 ---------------------------------------
template <typename T, int A, int B, int W>
__device__ __forceinline__ void cols_inv_tile(cx<T>* __restrict__ sm, const cx<T>* __restrict__ src_block, int tile, const cx<T>* __restrict__ twN,
                                              const cx<T>* __restrict__ twA, const cx<T>* __restrict__ twC, cx<T>* __restrict__ dst_block) {
    using X = CT<T>;
    using C = typename X::C;
    constexpr int Nb = kRowLen, NT = kThreads, GR = NT / W, BP = B + pad_rows<T, W>();
    const int c = threadIdx.x % W, g = threadIdx.x / W;
    const int n2 = tile * W + c;
    const cx<T>* src = src_block + n2;
    const C tw_step = X::mul(X::ldg(&twN[(long long)tile * W * A]), X::ldg(&twC[A * W + c]));
#pragma unroll
    for (int it = 0; it < A / GR; ++it) {
        const int ka = g + it * GR;
        C u[B];
#pragma unroll
        for (int kb = 0; kb < B; ++kb) u[kb] = X::ldcg(&src[(long long)(ka + A * kb) * Nb]);
        geo_twiddle<T, B, true, false>(u, X::mul(X::ldg(&twN[(long long)tile * W * ka]), X::ldg(&twC[ka * W + c])), tw_step);
        X::template dft<B, -1>(u);
        geo_twiddle<T, B, true, true>(u, X::one(), X::ldg(&twA[ka]));  // conj W_Na^(bi*ka)
#pragma unroll
        for (int bi = 0; bi < B; ++bi) X::st(&sm[(ka * BP + bi) * W + c], u[bi]);
    }
    __syncthreads();
    cx<T>* dst = dst_block + n2;
#pragma unroll
    for (int it = 0; it < B / GR; ++it) {
        const int bi = g + it * GR;
        C v[A];
#pragma unroll
        for (int ka = 0; ka < A; ++ka) v[ka] = X::ld(&sm[(ka * BP + bi) * W + c]);
        X::template dft<A, -1>(v);
#pragma unroll
        for (int a = 0; a < A / 2; ++a)  // n1 = a*B + bi < Na/2: the valid half of overlap-save (filters.rs:253)
            X::st_stream(&dst[(long long)(a * B + bi) * Nb], v[a]);
    }
}

constexpr int kRowWarps = 4;
constexpr int kRowPitch = 33;
template <typename T> constexpr size_t rows_smem() { return (size_t)kRowWarps * 32 * kRowPitch * 2 * sizeof(T); }

// ---- one row of 1024 points by one warp, in place: forward, times H, inverse ----------------------------------------
// n2 = 32*i1 + i2, k2 = p + 32*q; wl = W_1024^lane; sm = the warp's strip of 32 x 33 elements
template <typename T>
__device__ __forceinline__ void rows_row(cx<T>* __restrict__ sm, cx<T>* __restrict__ row, const cx<T>* __restrict__ hrow, typename CT<T>::C wl) {
    using X = CT<T>;
    using C = typename X::C;
    const int lane = threadIdx.x & 31;
    cx<T>* p = row + lane;
    const cx<T>* h = hrow + lane;
    // lane i2: DFT over i1 -> p, times W_1024^(i2*p)
    C v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = X::ldcg(&p[32 * i]);
    X::template dft<32, +1>(v);
    geo_twiddle<T, 32, false, true>(v, X::one(), wl);
#pragma unroll
    for (int i = 0; i < 32; ++i) X::st(&sm[i * kRowPitch + lane], v[i]);
    __syncwarp();
    // lane p: DFT over i2 -> q, times H, inverse DFT over q -> i2, times conj W_1024^(i2*p)
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = X::ld(&sm[lane * kRowPitch + i]);
    X::template dft<32, +1>(v);
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = X::mul(v[q], X::ldg(&h[32 * q]));
    X::template dft<32, -1>(v);
    geo_twiddle<T, 32, true, true>(v, X::one(), wl);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) X::st(&sm[lane * kRowPitch + i], v[i]);
    __syncwarp();
    // lane i2: inverse DFT over p -> i1
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = X::ld(&sm[i * kRowPitch + lane]);
    X::template dft<32, -1>(v);
#pragma unroll
    for (int i = 0; i < 32; ++i) X::st(&p[32 * i], v[i]);
    __syncwarp();
}

// ---- the three kernels -------------------------------------------------------------------------------------------------
template <typename T, int A, int B, int W>
__global__ void __launch_bounds__(kThreads)
k_long_cols_fwd(const cx<T>* __restrict__ in, long long in_stride, const cx<T>* __restrict__ hist, long long hist_stride, int first_chunk, int n_blocks,
                cx<T>* __restrict__ scratch, const cx<T>* __restrict__ twN, const cx<T>* __restrict__ twA, const cx<T>* __restrict__ twC) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int Na = A * B;
    const int w = blockIdx.y;
    const int s = w / n_blocks, b = w % n_blocks;
    const long long n = (long long)Na * kRowLen / 2;
    // window = [chunk ch-1 | chunk ch]; chunk -1 is the stored history (filters.rs:241-243)
    const int ch = first_chunk + b;
    const cx<T>* cur = in + (long long)s * in_stride + (long long)ch * n;
    const cx<T>* prev = (ch > 0) ? cur - n : hist + (long long)s * hist_stride;
    cols_fwd_tile<T, A, B, W>(reinterpret_cast<cx<T>*>(smem_raw), prev, cur, blockIdx.x, scratch + (long long)w * Na * kRowLen, twN, twA, twC);
}

template <typename T, int A, int B, int W>
__global__ void __launch_bounds__(kThreads)
k_long_cols_inv(const cx<T>* __restrict__ scratch, int n_blocks, const cx<T>* __restrict__ twN, const cx<T>* __restrict__ twA,
                const cx<T>* __restrict__ twC, cx<T>* __restrict__ out, long long out_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int Na = A * B;
    const int w = blockIdx.y;
    const int s = w / n_blocks, b = w % n_blocks;
    const long long n = (long long)Na * kRowLen / 2;
    cols_inv_tile<T, A, B, W>(reinterpret_cast<cx<T>*>(smem_raw), scratch + (long long)w * Na * kRowLen, blockIdx.x, twN, twA, twC,
                              out + (long long)s * out_stride + (long long)b * n);
}

template <typename T>
__global__ void __launch_bounds__(32 * kRowWarps)
k_long_rows(cx<T>* __restrict__ scratch, long long n_rows, int Na, const cx<T>* __restrict__ hbig, const cx<T>* __restrict__ twB) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw) + warp * (32 * kRowPitch);
    const typename CT<T>::C wl = CT<T>::ldg(&twB[lane]);
#pragma unroll 1
    for (long long row = (long long)blockIdx.x * kRowWarps + warp; row < n_rows; row += (long long)gridDim.x * kRowWarps)
        rows_row<T>(sm, scratch + row * kRowLen, hbig + (row % Na) * kRowLen, wl);
}

// Na = A x B and the columns per tile: rows of 128 bytes or more where 128 / W row groups divide both passes; Na = 1024
// takes 64-byte rows (68 KB of shared memory per CTA)
template <typename T, int NA> struct ColShape;
#define RR_COLSHAPE(NA_, A_, B_, WF32, WF64)                                                   \
    template <typename T> struct ColShape<T, NA_> {                                           \
        static constexpr int A = A_, B = B_, W = sizeof(T) == 4 ? WF32 : WF64;                \
    }
RR_COLSHAPE(32, 8, 4, 32, 32);
RR_COLSHAPE(64, 8, 8, 16, 16);
RR_COLSHAPE(128, 16, 8, 16, 16);
RR_COLSHAPE(256, 16, 16, 16, 8);
RR_COLSHAPE(512, 32, 16, 16, 8);
RR_COLSHAPE(1024, 32, 32, 8, 4);
#undef RR_COLSHAPE

template <typename T, int NA> cudaError_t launch_cols(bool fwd, int n_streams, const BigOsArgs<T>& a, cudaStream_t st) {
    using S = ColShape<T, NA>;
    constexpr int A = S::A, B = S::B, W = S::W;
    constexpr size_t smem = cols_smem<T, A, B, W>();
    const dim3 grid((unsigned)(kRowLen / W), (unsigned)(n_streams * a.n_blocks));
    cudaError_t e;
    if (fwd) {
        auto k = k_long_cols_fwd<T, A, B, W>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kThreads, smem, st>>>(reinterpret_cast<const cx<T>*>(a.in), a.in_stride, reinterpret_cast<const cx<T>*>(a.hist), a.hist_stride,
                                                        a.first_chunk, a.n_blocks, reinterpret_cast<cx<T>*>(a.scratch), reinterpret_cast<const cx<T>*>(a.twN),
                                                        reinterpret_cast<const cx<T>*>(a.twA), reinterpret_cast<const cx<T>*>(a.twC));
    } else {
        auto k = k_long_cols_inv<T, A, B, W>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kThreads, smem, st>>>(reinterpret_cast<const cx<T>*>(a.scratch), a.n_blocks, reinterpret_cast<const cx<T>*>(a.twN),
                                                        reinterpret_cast<const cx<T>*>(a.twA), reinterpret_cast<const cx<T>*>(a.twC), reinterpret_cast<cx<T>*>(a.out),
                                                        a.out_stride);
    }
    return cudaGetLastError();
}

template <typename T> cudaError_t launch_rows(int Na, int n_streams, const BigOsArgs<T>& a, cudaStream_t st) {
    auto k = k_long_rows<T>;
    constexpr size_t smem = rows_smem<T>();
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long n_rows = (long long)n_streams * a.n_blocks * Na;
    const long long ctas = (n_rows + kRowWarps - 1) / kRowWarps;
    k<<<(unsigned)ctas, 32 * kRowWarps, smem, st>>>(reinterpret_cast<cx<T>*>(a.scratch), n_rows, Na, reinterpret_cast<const cx<T>*>(a.hbig),
                                                     reinterpret_cast<const cx<T>*>(a.twB));
    return cudaGetLastError();
}

template <typename T> cudaError_t cols_dispatch(bool fwd, int Na, int S, const BigOsArgs<T>& a, cudaStream_t st) {
    switch (Na) {
        case 32: return launch_cols<T, 32>(fwd, S, a, st);
        case 64: return launch_cols<T, 64>(fwd, S, a, st);
        case 128: return launch_cols<T, 128>(fwd, S, a, st);
        case 256: return launch_cols<T, 256>(fwd, S, a, st);
        case 512: return launch_cols<T, 512>(fwd, S, a, st);
        case 1024: return launch_cols<T, 1024>(fwd, S, a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

// RR_BIG_OS_LEGACY=1 keeps rr_big_os.cu's kernels at every size (A/B runs)
bool long_os_supported(int n) {
    static const bool legacy = [] {
        const char* e = std::getenv("RR_BIG_OS_LEGACY");
        return e && std::atoi(e) != 0;
    }();
    if (legacy || n < 16384 || (n & (n - 1)) != 0) return false;
    return 2LL * n <= (1LL << 20);
}
void long_os_shape(int n, int* Na, int* Nb) {
    *Na = (int)(2LL * n / kRowLen);
    *Nb = kRowLen;
}
// row k1 of the table holds bins k1 + Na*k2 in natural order of k2 (lane p of the row kernel reads bin p + 32*q)
long long long_os_hperm_index(int n, long long k) {
    const long long Na = 2LL * n / kRowLen;
    return (k % Na) * kRowLen + k / Na;
}

// positions of the column kernels' small twiddle table (BigOsArgs::twC): entry i holds W_N^e, e = long_os_twc_exponent(i)
template <typename T> int long_os_twc_size(int n) {
    int A = 0, W = 0;
    switch ((int)(2LL * n / kRowLen)) {
#define X(NA_) case NA_: A = ColShape<T, NA_>::A; W = ColShape<T, NA_>::W; break;
        X(32) X(64) X(128) X(256) X(512) X(1024)
#undef X
        default: return 0;
    }
    return (A + 1) * W;
}
template <typename T> long long long_os_twc_exponent(int n, int i) {
    int A = 0, W = 0;
    switch ((int)(2LL * n / kRowLen)) {
#define X(NA_) case NA_: A = ColShape<T, NA_>::A; W = ColShape<T, NA_>::W; break;
        X(32) X(64) X(128) X(256) X(512) X(1024)
#undef X
        default: return 0;
    }
    const int k = i / W, c = i % W;  // rows k < A: W_N^(c*k); row A: W_N^(c*A)
    return (long long)c * k;
}
template int long_os_twc_size<float>(int);
template int long_os_twc_size<double>(int);
template long long long_os_twc_exponent<float>(int, int);
template long long long_os_twc_exponent<double>(int, int);

template <typename T> cudaError_t launch_long_os(int n, int n_streams, const BigOsArgs<T>& a, cudaStream_t st) {
    int Na, Nb;
    long_os_shape(n, &Na, &Nb);
    cudaError_t e = cols_dispatch<T>(true, Na, n_streams, a, st);
    if (e != cudaSuccess) return e;
    e = launch_rows<T>(Na, n_streams, a, st);
    if (e != cudaSuccess) return e;
    return cols_dispatch<T>(false, Na, n_streams, a, st);
}
template cudaError_t launch_long_os<float>(int, int, const BigOsArgs<float>&, cudaStream_t);
template cudaError_t launch_long_os<double>(int, int, const BigOsArgs<double>&, cudaStream_t);

}  // namespace rr

// Overlap-save fast convolution for chunks too long for one SM's shared memory
// (Filter with 2n = 32768 .. 2^24 points, e.g. BASELINE config 5: f64,
// 2^20-point FFT).  Replaces the two rustfft calls and the H multiply of
// src/blocks/filters.rs:244-252 with a four-step FFT, N = Na * Nb:
//
//   k_big_cols_fwd : per column n2, DFT_Na over n1 (stride Nb), times W_N^(n2*k1)  -> scratch[k1][n2]
//   k_big_rows     : per row k1, DFT_Nb over n2, times H[k1 + Na*k2], inverse DFT_Nb -> scratch[k1][n2]
//   k_big_cols_inv : per column n2, times conj W_N^(n2*k1), inverse DFT_Na over k1 -> y[n1*Nb + n2], n1 < Na/2
//
// The scratch (N points per overlap-save block) is written and re-read twice;
// the host bounds the number of blocks per launch so that it stays L2 resident.
// Column passes interleave G transforms in shared memory (FftPlan<.., G>) so a
// group of G lanes reads G consecutive samples of a row.  sm_100a.
#include "rr_fft_plan.cuh"
#include "rr_kernels.h"

namespace rr {

// columns a CTA interleaves
template <typename T> constexpr int kColGOf = 8;  // (16 f32 columns = 128-byte rows was measured: slower at 2^17 points, 1024-thread CTAs)

template <typename T, int NA>
__global__ void __launch_bounds__(PlanFor<T, NA, kColGOf<T>>::type::NT* kColGOf<T>)
k_big_cols_fwd(const cx<T>* __restrict__ in, long long in_stride, const cx<T>* __restrict__ hist, long long hist_stride, int first_chunk, int n_blocks,
               int Nb, cx<T>* __restrict__ scratch, const cx<T>* __restrict__ twN, const cx<T>* __restrict__ twA) {
    using P = typename PlanFor<T, NA, kColGOf<T>>::type;
    constexpr int NT = P::NT, R1 = P::R1, B1 = P::B1, S1 = P::S1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kColG = kColGOf<T>;
    const int g = threadIdx.x % kColG, t = threadIdx.x / kColG;
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw) + g;
    const int w = blockIdx.y;
    const int s = w / n_blocks, b = w % n_blocks;
    const long long n = (long long)NA * Nb / 2;
    // window = [chunk c-1 | chunk c]; chunk -1 is the stored history (filters.rs:241-243)
    const int c = first_chunk + b;
    const cx<T>* cur = in + (long long)s * in_stride + (long long)c * n;
    const cx<T>* prev = (c > 0) ? cur - n : hist + (long long)s * hist_stride;
    const int col = blockIdx.x * kColG + g;

    P plan;
    plan.init(twA, t);
    cx<T> v[B1][R1];
#pragma unroll
    for (int bb = 0; bb < B1; ++bb)
#pragma unroll
        for (int r = 0; r < R1; ++r) {
            const long long i = (long long)(B1 * t + bb + S1 * r) * Nb + col;
            v[bb][r] = (r < R1 / 2) ? ld_cx(&prev[i]) : ld_cx(&cur[i - n]);  // n1 < Na/2 <=> first half
        }
    plan.p1_forward(sm, t, v);
    __syncthreads();
    plan.template p2<+1>(sm, t);
    __syncthreads();
    // pass 3 in registers, then straight to the scratch: run u = (k1, k2) holds bins k = k1 + R1*k2 + R1*R2*k3 of its
    // column, each times the four-step twiddle W_N^(k*col) -- one gathered table read per run and a rotation per bin
    // (a gathered read per point cost 32 wavefronts per warp instruction and saturated the LSU pipe)
    constexpr int R2 = P::R2, R3 = P::R3, B3 = P::B3;
    cx<T>* dst = scratch + (long long)w * NA * Nb;
    const cx<T> tw_step = ld_cx(&twN[(long long)(R1 * R2) * col]);  // R1*R2 < Na, col < Nb: inside the table
#pragma unroll
    for (int rb = 0; rb < B3; ++rb) {
        const int u = B3 * t + rb;
        cx<T> y[R3];
        P::p3_load(sm, u, y);
        dft_regs<R3, +1, T>(y);
        const int k0 = u / R2 + R1 * (u % R2);
        cx<T> tw = ld_cx(&twN[(long long)k0 * col]);
#pragma unroll
        for (int k3 = 0; k3 < R3; ++k3) {
            st_cx(&dst[(long long)(k0 + R1 * R2 * k3) * Nb + col], cmul(y[k3], tw));
            tw = cmul(tw, tw_step);
        }
    }
}

template <typename T, int NA>
__global__ void __launch_bounds__(PlanFor<T, NA, kColGOf<T>>::type::NT* kColGOf<T>)
k_big_cols_inv(const cx<T>* __restrict__ scratch, int n_blocks, int Nb, const cx<T>* __restrict__ twN,
               const cx<T>* __restrict__ twA, cx<T>* __restrict__ out, long long out_stride) {
    using P = typename PlanFor<T, NA, kColGOf<T>>::type;
    constexpr int NT = P::NT, R1 = P::R1, B1 = P::B1, S1 = P::S1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kColG = kColGOf<T>;
    const int g = threadIdx.x % kColG, t = threadIdx.x / kColG;
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw) + g;
    const int w = blockIdx.y;
    const int s = w / n_blocks, b = w % n_blocks;
    const long long n = (long long)NA * Nb / 2;
    const int col = blockIdx.x * kColG + g;
    const cx<T>* src = scratch + (long long)w * NA * Nb;

    P plan;
    plan.init(twA, t);
    // the mirror image of the forward kernel's tail: bins straight from the scratch into pass 3's registers
    constexpr int R2 = P::R2, R3 = P::R3, B3 = P::B3;
    const cx<T> tw_step = ld_cx(&twN[(long long)(R1 * R2) * col]);
#pragma unroll
    for (int rb = 0; rb < B3; ++rb) {
        const int u = B3 * t + rb;
        const int k0 = u / R2 + R1 * (u % R2);
        cx<T> tw = ld_cx(&twN[(long long)k0 * col]);
        cx<T> y[R3];
#pragma unroll
        for (int k3 = 0; k3 < R3; ++k3) {
            y[k3] = cmulc(ld_cx(&src[(long long)(k0 + R1 * R2 * k3) * Nb + col]), tw);
            tw = cmul(tw, tw_step);
        }
        dft_regs<R3, -1, T>(y);
        P::p3_store(sm, u, y);
    }
    __syncthreads();
    plan.template p2<-1>(sm, t);
    __syncthreads();
    cx<T> v[B1][R1];
    plan.p1_inverse(sm, t, v);
    cx<T>* dst = out + (long long)s * out_stride + (long long)b * n;
#pragma unroll
    for (int bb = 0; bb < B1; ++bb)
#pragma unroll
        for (int r = 0; r < R1 / 2; ++r)  // n1 < Na/2: the valid half of overlap-save (filters.rs:253)
            st_cx(&dst[(long long)(B1 * t + bb + S1 * r) * Nb + col], v[bb][r]);
}

// ROWS rows per CTA, each by its own group of NT threads.  The groups are independent: they synchronise among
// themselves only (a warp barrier up to 32 threads, a named barrier above), not across the CTA.
template <int NT> __device__ __forceinline__ void row_group_sync(int grp) {
    if constexpr (NT <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(NT) : "memory");
    }
}

template <typename T, int NB, int ROWS>
__global__ void __launch_bounds__(PlanFor<T, NB>::type::NT* ROWS)
k_big_rows(cx<T>* __restrict__ scratch, int Na, const cx<T>* __restrict__ hbig, const cx<T>* __restrict__ twB) {
    using P = typename PlanFor<T, NB>::type;
    constexpr int NT = P::NT, R1 = P::R1, B1 = P::B1, S1 = P::S1;
    static_assert(NT <= 32 || ROWS <= 15, "one named barrier per row group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / NT, t = threadIdx.x % NT;
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)grp * P::SMEM_ELEMS;
    const int k1 = blockIdx.x * ROWS + grp;
    cx<T>* row = scratch + ((long long)blockIdx.y * Na + k1) * NB;
    const cx<T>* hrow = hbig + (long long)k1 * NB;
    // a thread's B1 consecutive points travel as 16-byte pairs where the type allows it
    constexpr bool PAIRS = sizeof(T) == 4 && (B1 % 2) == 0;

    P plan;
    plan.init(twB, t);
    cx<T> v[B1][R1];
    if constexpr (PAIRS) {
#pragma unroll
        for (int bb = 0; bb < B1; bb += 2)
#pragma unroll
            for (int r = 0; r < R1; ++r) ld_cx2(&row[B1 * t + bb + S1 * r], v[bb][r], v[bb + 1][r]);
    } else {
#pragma unroll
        for (int bb = 0; bb < B1; ++bb)
#pragma unroll
            for (int r = 0; r < R1; ++r) v[bb][r] = ld_cx(&row[B1 * t + bb + S1 * r]);
    }
    plan.p1_forward(sm, t, v);
    row_group_sync<NT>(grp);
    plan.template p2<+1>(sm, t);
    row_group_sync<NT>(grp);
    P::p3_fwd_mul_inv(sm, t, hrow);
    row_group_sync<NT>(grp);
    plan.template p2<-1>(sm, t);
    row_group_sync<NT>(grp);
    plan.p1_inverse(sm, t, v);
    if constexpr (PAIRS) {
#pragma unroll
        for (int bb = 0; bb < B1; bb += 2)
#pragma unroll
            for (int r = 0; r < R1; ++r) st_cx2(&row[B1 * t + bb + S1 * r], v[bb][r], v[bb + 1][r]);
    } else {
#pragma unroll
        for (int bb = 0; bb < B1; ++bb)
#pragma unroll
            for (int r = 0; r < R1; ++r) st_cx(&row[B1 * t + bb + S1 * r], v[bb][r]);
    }
}

// ---- shapes -------------------------------------------------------------------
static inline void shape_for(long long N, int* Na, int* Nb) {
    long long a = N / 256;
    if (a < 64) a = 64;
    if (a > 1024) a = 1024;
    *Na = (int)a;
    *Nb = (int)(N / a);
}
template <typename T> constexpr int max_row_plan() { return sizeof(T) == 4 ? 16384 : 4096; }

template <typename T> bool big_os_supported(int n) {
    if (long_os_supported(n)) return true;
    if (n < 2048 || (n & (n - 1)) != 0) return false;
    const long long N = 2LL * n;
    int Na, Nb;
    shape_for(N, &Na, &Nb);
    return Nb >= 64 && Nb <= max_row_plan<T>() && (long long)Na * Nb == N;
}
template <typename T> void big_os_shape(int n, int* Na, int* Nb) {
    if (long_os_supported(n)) return long_os_shape(n, Na, Nb);
    shape_for(2LL * n, Na, Nb);
}

template <typename T, int NB> static long long row_hperm(long long k2) { return PlanFor<T, NB>::type::hperm_index((int)k2); }

#define RR_ROW_SIZES_F32(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)
#define RR_ROW_SIZES_F64(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
#define RR_COL_SIZES(X) X(64) X(128) X(256) X(512) X(1024)

template <> long long big_os_hperm_index<float>(int n, long long k) {
    if (long_os_supported(n)) return long_os_hperm_index(n, k);
    int Na, Nb;
    shape_for(2LL * n, &Na, &Nb);
    const long long k1 = k % Na, k2 = k / Na;
    switch (Nb) {
#define X(NN) case NN: return k1 * Nb + row_hperm<float, NN>(k2);
        RR_ROW_SIZES_F32(X)
#undef X
    }
    return -1;
}
template <> long long big_os_hperm_index<double>(int n, long long k) {
    if (long_os_supported(n)) return long_os_hperm_index(n, k);
    int Na, Nb;
    shape_for(2LL * n, &Na, &Nb);
    const long long k1 = k % Na, k2 = k / Na;
    switch (Nb) {
#define X(NN) case NN: return k1 * Nb + row_hperm<double, NN>(k2);
        RR_ROW_SIZES_F64(X)
#undef X
    }
    return -1;
}

template <typename T, int NA> static cudaError_t launch_cols(bool fwd, int Nb, int n_streams, const BigOsArgs<T>& a, cudaStream_t st) {
    constexpr int kColG = kColGOf<T>;
    using P = typename PlanFor<T, NA, kColGOf<T>>::type;
    const size_t smem = sizeof(cx<T>) * P::SMEM_ELEMS;
    dim3 grid((unsigned)(Nb / kColG), (unsigned)(n_streams * a.n_blocks));
    cudaError_t e;
    if (fwd) {
        auto k = k_big_cols_fwd<T, NA>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, P::NT * kColG, smem, st>>>(reinterpret_cast<const cx<T>*>(a.in), a.in_stride, reinterpret_cast<const cx<T>*>(a.hist),
                                             a.hist_stride, a.first_chunk, a.n_blocks, Nb, reinterpret_cast<cx<T>*>(a.scratch),
                                             reinterpret_cast<const cx<T>*>(a.twN), reinterpret_cast<const cx<T>*>(a.twA));
    } else {
        auto k = k_big_cols_inv<T, NA>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, P::NT * kColG, smem, st>>>(reinterpret_cast<const cx<T>*>(a.scratch), a.n_blocks, Nb,
                                             reinterpret_cast<const cx<T>*>(a.twN), reinterpret_cast<const cx<T>*>(a.twA),
                                             reinterpret_cast<cx<T>*>(a.out), a.out_stride);
    }
    return cudaGetLastError();
}

template <typename T, int NB> static cudaError_t launch_rows(int Na, int n_streams, const BigOsArgs<T>& a, cudaStream_t st) {
    using P = typename PlanFor<T, NB>::type;
    constexpr int ROWS = (P::NT >= 256) ? 1 : 256 / P::NT;
    const size_t smem = sizeof(cx<T>) * P::SMEM_ELEMS * ROWS;
    auto k = k_big_rows<T, NB, ROWS>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)(Na / ROWS), (unsigned)(n_streams * a.n_blocks));
    k<<<grid, P::NT * ROWS, smem, st>>>(reinterpret_cast<cx<T>*>(a.scratch), Na, reinterpret_cast<const cx<T>*>(a.hbig),
                                        reinterpret_cast<const cx<T>*>(a.twB));
    return cudaGetLastError();
}

template <typename T> static cudaError_t cols_dispatch(bool fwd, int Na, int Nb, int S, const BigOsArgs<T>& a, cudaStream_t st) {
    switch (Na) {
#define X(NN) case NN: return launch_cols<T, NN>(fwd, Nb, S, a, st);
        RR_COL_SIZES(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

template <> cudaError_t launch_big_os<float>(int n, int n_streams, const BigOsArgs<float>& a, cudaStream_t st) {
    if (long_os_supported(n)) return launch_long_os<float>(n, n_streams, a, st);
    int Na, Nb;
    shape_for(2LL * n, &Na, &Nb);
    cudaError_t e = cols_dispatch<float>(true, Na, Nb, n_streams, a, st);
    if (e != cudaSuccess) return e;
    switch (Nb) {
#define X(NN) case NN: e = launch_rows<float, NN>(Na, n_streams, a, st); break;
        RR_ROW_SIZES_F32(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return cols_dispatch<float>(false, Na, Nb, n_streams, a, st);
}
template <> cudaError_t launch_big_os<double>(int n, int n_streams, const BigOsArgs<double>& a, cudaStream_t st) {
    if (long_os_supported(n)) return launch_long_os<double>(n, n_streams, a, st);
    int Na, Nb;
    shape_for(2LL * n, &Na, &Nb);
    cudaError_t e = cols_dispatch<double>(true, Na, Nb, n_streams, a, st);
    if (e != cudaSuccess) return e;
    switch (Nb) {
#define X(NN) case NN: e = launch_rows<double, NN>(Na, n_streams, a, st); break;
        RR_ROW_SIZES_F64(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return cols_dispatch<double>(false, Na, Nb, n_streams, a, st);
}

template bool big_os_supported<float>(int);
template bool big_os_supported<double>(int);
template void big_os_shape<float>(int, int*, int*);
template void big_os_shape<double>(int, int*, int*);

}  // namespace rr

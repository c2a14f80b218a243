// Fourier analysis block: window * x -> forward FFT (-> DC to the centre).
//
// Reference: blocks::analysis::Fourier, src/blocks/analysis.rs:60-132 -- per chunk of n samples:
// window_values[i] = W(2(i+0.5)/n - 1) * sqrt(n / sum W^2) (rounded to Flt, :86-103), the chunk is
// multiplied by them (:107-109), rustfft forward (unnormalised, :110-112), rotate_right(n/2) when the
// DC bin is wanted in the centre (:113-115).
//
// Power-of-two n with a three-pass plan (64 .. 16384 f32, .. 4096 f64): one CTA per (chunk, stream),
// the same register/shared-memory passes as the overlap-save kernel.  Any other n up to 4096: a direct
// DFT (the reference accepts every length; its own known-answer test uses n = 3 and n = 4).
#pragma once
#include "rr_fft_plan.cuh"
#include "rr_kernels.h"

namespace rr {

template <typename T, int N>
__global__ void __launch_bounds__(PlanFor<T, N>::type::NT)
k_fourier(const cx<T>* __restrict__ in, long long in_stride, cx<T>* __restrict__ out, long long out_stride,
          const T* __restrict__ window, const cx<T>* __restrict__ twN, int rot) {
    using P = typename PlanFor<T, N>::type;
    constexpr int NT = P::NT, R1 = P::R1, B1 = P::B1, S1 = P::S1;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* sm = reinterpret_cast<cx<T>*>(smem_raw);
    const cx<T>* src = in + (long long)blockIdx.y * in_stride + (long long)blockIdx.x * N;
    cx<T>* dst = out + (long long)blockIdx.y * out_stride + (long long)blockIdx.x * N;
    P plan;
    plan.init(twN, tid);
    cx<T> v[B1][R1];
#pragma unroll
    for (int b = 0; b < B1; ++b)
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            const int i = B1 * tid + b + S1 * j;
            v[b][j] = cscale(ld_cx(&src[i]), window[i]);
        }
    plan.p1_forward(sm, tid, v);
    __syncthreads();
    plan.template p2<+1>(sm, tid);
    __syncthreads();
    P::template p3<+1>(sm, tid);
    __syncthreads();
    for (int k = tid; k < N; k += NT) {
        int o = k + rot;
        if (o >= N) o -= N;
        st_cx(&dst[o], sm[P::sidx(P::bin_position(k))]);
    }
}

template <typename T, int N>
cudaError_t launch_fourier_n(const void* in, long long in_stride, void* out, long long out_stride, int n_chunks, int n_streams,
                             const T* window, const void* twN, int rot, cudaStream_t st) {
    using P = typename PlanFor<T, N>::type;
    const size_t smem = sizeof(cx<T>) * P::SMEM_ELEMS;
    auto kern = k_fourier<T, N>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3((unsigned)n_chunks, (unsigned)n_streams), P::NT, smem, st>>>(reinterpret_cast<const cx<T>*>(in), in_stride,
                                                                             reinterpret_cast<cx<T>*>(out), out_stride, window,
                                                                             reinterpret_cast<const cx<T>*>(twN), rot);
    return cudaGetLastError();
}

}  // namespace rr

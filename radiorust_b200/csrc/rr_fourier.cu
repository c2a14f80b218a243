// Size dispatch of the Fourier block's FFT kernel + the direct DFT for every other chunk length.
#include "rr_fourier.cuh"

namespace rr {

#define RR_FOR_SIZES_F32(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)
#define RR_FOR_SIZES_F64(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

#define X(NN) \
    extern template cudaError_t launch_fourier_n<float, NN>(const void*, long long, void*, long long, int, int, const float*, const void*, int, cudaStream_t);
RR_FOR_SIZES_F32(X)
#undef X
#define X(NN) \
    extern template cudaError_t launch_fourier_n<double, NN>(const void*, long long, void*, long long, int, int, const double*, const void*, int, cudaStream_t);
RR_FOR_SIZES_F64(X)
#undef X

// out[(k + rot) mod n] = sum_j w[j] x[j] exp(-j*2*pi*jk/n); products in T like a rustfft butterfly, the sum in
// double so that the O(n) terms do not cost accuracy
template <typename T>
__global__ void __launch_bounds__(256) k_dft_direct(const cx<T>* __restrict__ in, long long in_stride, cx<T>* __restrict__ out,
                                                    long long out_stride, const T* __restrict__ window, int n, int rot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<double>* xs = reinterpret_cast<cx<double>*>(smem_raw);  // [n] windowed input
    cx<double>* tw = xs + n;                                   // [n] exp(-j*2*pi*m/n)
    const cx<T>* src = in + (long long)blockIdx.y * in_stride + (long long)blockIdx.x * n;
    cx<T>* dst = out + (long long)blockIdx.y * out_stride + (long long)blockIdx.x * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const cx<T> v = cscale(ld_cx(&src[j]), window[j]);
        xs[j] = cx<double>((double)v.x, (double)v.y);
        double s, c;
        sincospi(-2.0 * (double)j / (double)n, &s, &c);
        tw[j] = cx<double>(c, s);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        cx<double> acc(0.0, 0.0);
        int m = 0;
        for (int j = 0; j < n; ++j) {
            acc = acc + cmul(xs[j], tw[m]);
            m += k;
            if (m >= n) m -= n;
        }
        int o = k + rot;
        if (o >= n) o -= n;
        st_cx(&dst[o], cx<T>((T)acc.x, (T)acc.y));
    }
}

template <> bool fourier_fft_supported<float>(int n) {
    switch (n) {
#define X(NN) case NN: return true;
        RR_FOR_SIZES_F32(X)
#undef X
    }
    return false;
}
template <> bool fourier_fft_supported<double>(int n) {
    switch (n) {
#define X(NN) case NN: return true;
        RR_FOR_SIZES_F64(X)
#undef X
    }
    return false;
}

template <typename T>
static cudaError_t direct(int n, const void* in, long long in_stride, void* out, long long out_stride, int n_chunks, int n_streams,
                          const T* window, int rot, cudaStream_t st) {
    if (n < 1 || n > kFourierDirectMax) return cudaErrorInvalidValue;
    const size_t smem = (size_t)n * 2 * sizeof(cx<double>);
    auto kern = k_dft_direct<T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<dim3((unsigned)n_chunks, (unsigned)n_streams), 256, smem, st>>>(reinterpret_cast<const cx<T>*>(in), in_stride,
                                                                           reinterpret_cast<cx<T>*>(out), out_stride, window, n, rot);
    return cudaGetLastError();
}

template <>
cudaError_t launch_fourier<float>(int n, const void* in, long long in_stride, void* out, long long out_stride, int n_chunks, int n_streams,
                                  const float* window, const void* twN, int rot, cudaStream_t st) {
    switch (n) {
#define X(NN) case NN: return launch_fourier_n<float, NN>(in, in_stride, out, out_stride, n_chunks, n_streams, window, twN, rot, st);
        RR_FOR_SIZES_F32(X)
#undef X
    }
    return direct<float>(n, in, in_stride, out, out_stride, n_chunks, n_streams, window, rot, st);
}
template <>
cudaError_t launch_fourier<double>(int n, const void* in, long long in_stride, void* out, long long out_stride, int n_chunks, int n_streams,
                                   const double* window, const void* twN, int rot, cudaStream_t st) {
    switch (n) {
#define X(NN) case NN: return launch_fourier_n<double, NN>(in, in_stride, out, out_stride, n_chunks, n_streams, window, twN, rot, st);
        RR_FOR_SIZES_F64(X)
#undef X
    }
    return direct<double>(n, in, in_stride, out, out_stride, n_chunks, n_streams, window, rot, st);
}

}  // namespace rr

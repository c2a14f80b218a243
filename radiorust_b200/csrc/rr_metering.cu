// metering::bandwidth and metering::rescale_energy (src/metering.rs:32-110) on Fourier-transformed chunks that
// live on the device (the output of the Fourier stage).  sm_100a.
#include "rr_chain_os.cuh"

namespace rr {

// norm_sqr in Flt (re*re + im*im, each operation rounded as in num::Complex::norm_sqr; no contraction)
__device__ __forceinline__ float nsq_flt(const cx<float>& v) { return __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)); }
__device__ __forceinline__ double nsq_flt(const cx<double>& v) { return __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)); }

// ---------------------------------------------------------------------------
// bandwidth (metering.rs:42-84).  One CTA per chunk.  total = sum of bin energies (f64); limit = total*dp/2;
// bins are walked from the band edge (index wrap = (n+1)/2, i.e. the most negative frequency) upwards and, in a
// second pass, from the other edge downwards, until the running energy exceeds the limit; the bin where it
// does counts by the fraction that fits.  The running energy is a block-wide prefix sum in f64 (fixed order:
// deterministic; the reference adds sequentially, the two differ by rounding of the f64 sums only).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_bandwidth(const cx<T>* __restrict__ in, long long in_stride, long long chunk_len,
                                                   long long n_chunks, double double_percentile, double sample_rate,
                                                   double* __restrict__ out) {
    __shared__ double wsum[8];
    __shared__ double s_val[2];
    __shared__ long long s_first;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = chunk_len;
    const cx<T>* src = in + (long long)blockIdx.y * in_stride + (long long)blockIdx.x * n;

    double acc = 0.0;
    for (long long t = tid; t < n; t += 256) acc += (double)nsq_flt(ld_cx(&src[t]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < 8; ++w) total += wsum[w];
        s_val[0] = total * double_percentile / 2.0;  // energy_limit (metering.rs:72)
    }
    __syncthreads();
    const double limit = s_val[0];
    const long long wrap = (n + 1) / 2;  // metering.rs:73

    double used_total = 0.0;
    for (int dir = 0; dir < 2; ++dir) {
        double carry = 0.0;
        double used = (double)n;  // every bin fits (metering.rs:62-63)
        for (long long i0 = 0; i0 < n; i0 += 256) {
            const long long i = i0 + tid;
            double e = 0.0;
            if (i < n) {
                long long idx = wrap + (dir == 0 ? i : n - 1 - i);
                if (idx >= n) idx -= n;
                e = (double)nsq_flt(ld_cx(&src[idx]));
            }
            double incl = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            __syncthreads();  // wsum / s_first of the previous round are no longer read
            if (lane == 31) wsum[warp] = incl;
            if (tid == 0) s_first = -1;
            __syncthreads();
            double before = carry;
            for (int w = 0; w < warp; ++w) before += wsum[w];
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.0;
            const double old_energy = before + excl, new_energy = old_energy + e;
            const bool crosses = i < n && new_energy > limit;  // metering.rs:58
            const unsigned m = __ballot_sync(0xffffffffu, crosses);
            if (m != 0 && lane == __ffs(m) - 1) atomicMin((unsigned long long*)&s_first, (unsigned long long)i);  // -1 = max
            __syncthreads();
            const long long first = s_first;
            if (first >= 0) {
                if (i == first) s_val[1] = (double)i + (limit - old_energy) / (new_energy - old_energy);  // metering.rs:59
                __syncthreads();
                used = s_val[1];
                break;
            }
            double tot = carry;
            for (int w = 0; w < 8; ++w) tot += wsum[w];
            carry = tot;
        }
        used_total += used;
        __syncthreads();
    }
    if (tid == 0) {
        const double bw = ((double)n - used_total) * sample_rate / (double)n;  // metering.rs:78
        out[(long long)blockIdx.y * n_chunks + blockIdx.x] = bw > 0.0 ? bw : 0.0;
    }
}

template <typename T>
cudaError_t launch_bandwidth(const void* in, long long in_stride, long long chunk_len, long long n_chunks, int n_streams,
                             double double_percentile, double sample_rate, double* out, cudaStream_t st) {
    if (n_chunks <= 0 || chunk_len <= 0) return cudaSuccess;
    k_bandwidth<T><<<dim3((unsigned)n_chunks, (unsigned)n_streams), 256, 0, st>>>(reinterpret_cast<const cx<T>*>(in), in_stride, chunk_len,
                                                                                 n_chunks, double_percentile, sample_rate, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// rescale_energy (metering.rs:93-110): `resolution` real numbers per chunk, output o = the energy of the
// input bins under [o, o+1)*n/resolution, edge bins weighted by their overlap.  One thread per output; every
// operation in Flt and in the reference's order (bit-exact).
// ---------------------------------------------------------------------------
template <typename T> struct RsOps;
template <> struct RsOps<float> {
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct RsOps<double> {
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};

template <typename T>
__global__ void __launch_bounds__(256) k_rescale_energy(const cx<T>* __restrict__ in, long long in_stride, long long chunk_len,
                                                        long long n_chunks, long long resolution, T* __restrict__ out) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= resolution) return;
    const long long c = blockIdx.y, s = blockIdx.z, n = chunk_len;
    const cx<T>* src = in + s * in_stride + c * n;
    using O = RsOps<T>;
    const T fn = (T)n, fr = (T)resolution;
    const T left = O::mul(O::div((T)o, fr), fn);
    const T right = O::mul(O::div(O::add((T)o, (T)1), fr), fn);
    long long lf = (long long)floor(left), rc = (long long)ceil(right);
    if (lf > n - 1) lf = n - 1;
    if (rc > n) rc = n;
    T acc = (T)0;
    for (long long idx = lf; idx < rc; ++idx) {
        const T lb = fmax((T)idx, left);
        const T rb = fmin(O::add((T)idx, (T)1), right);
        const T scale = O::add(rb, -lb);
        acc = O::add(acc, O::mul(nsq_flt(ld_cx(&src[idx])), scale));
    }
    out[(s * n_chunks + c) * resolution + o] = acc;
}

template <typename T>
cudaError_t launch_rescale_energy(const void* in, long long in_stride, long long chunk_len, long long n_chunks, int n_streams,
                                  long long resolution, void* out, cudaStream_t st) {
    if (n_chunks <= 0 || resolution <= 0) return cudaSuccess;
    k_rescale_energy<T><<<dim3((unsigned)((resolution + 255) / 256), (unsigned)n_chunks, (unsigned)n_streams), 256, 0, st>>>(
        reinterpret_cast<const cx<T>*>(in), in_stride, chunk_len, n_chunks, resolution, reinterpret_cast<T*>(out));
    return cudaGetLastError();
}

#define RR_INST(T)                                                                                                                  \
    template cudaError_t launch_bandwidth<T>(const void*, long long, long long, long long, int, double, double, double*, cudaStream_t); \
    template cudaError_t launch_rescale_energy<T>(const void*, long long, long long, long long, int, long long, void*, cudaStream_t);
RR_INST(float)
RR_INST(double)
#undef RR_INST

}  // namespace rr

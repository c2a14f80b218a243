// Dispatch of the polyphase fused chain kernel + the history update kernel.
#include "rr_poly.cuh"

namespace rr {

#define RR_POLY_KS(X, T, G) X(T, 256, G) X(T, 512, G) X(T, 1024, G)
#define RR_POLY_ALL(X) RR_POLY_KS(X, float, 8) RR_POLY_KS(X, float, 10) RR_POLY_KS(X, double, 8) RR_POLY_KS(X, double, 10)

#define X(T, K, G)                                                                                              \
    extern template cudaError_t launch_poly_n<T, K, 1, G>(int, const PolyArgs<T>&, cudaStream_t);              \
    extern template cudaError_t launch_poly_n<T, K, 2, G>(int, const PolyArgs<T>&, cudaStream_t);              \
    extern template cudaError_t launch_poly_n<T, K, 3, G>(int, const PolyArgs<T>&, cudaStream_t);              \
    extern template cudaError_t launch_poly_n<T, K, 4, G>(int, const PolyArgs<T>&, cudaStream_t);
RR_POLY_ALL(X)
#undef X

template <typename T> bool poly_supported(int K, int Q, int G) {
    return (K == 256 || K == 512 || K == 1024) && Q >= 1 && Q <= 4 && (G == 8 || G == 10);
}
template bool poly_supported<float>(int, int, int);
template bool poly_supported<double>(int, int, int);

template <typename T> int poly_hperm_index(int K, int k) {
    switch (K) {
        case 256: return PlanFor<T, 256>::type::hperm_index(k);
        case 512: return PlanFor<T, 512>::type::hperm_index(k);
        case 1024: return PlanFor<T, 1024>::type::hperm_index(k);
    }
    return -1;
}
template int poly_hperm_index<float>(int, int);
template int poly_hperm_index<double>(int, int);

template <typename T, int K> static size_t smem_k(int Q, int G, int nbpc) {
    const size_t e = (G == 8) ? PolyCfg<T, K, 8>::smem_elems(nbpc, Q) : PolyCfg<T, K, 10>::smem_elems(nbpc, Q);
    return sizeof(cx<T>) * e;
}
template <typename T> size_t poly_smem_bytes(int K, int Q, int G, int nbpc) {
    switch (K) {
        case 256: return smem_k<T, 256>(Q, G, nbpc);
        case 512: return smem_k<T, 512>(Q, G, nbpc);
        case 1024: return smem_k<T, 1024>(Q, G, nbpc);
    }
    return (size_t)-1;
}
template size_t poly_smem_bytes<float>(int, int, int, int);
template size_t poly_smem_bytes<double>(int, int, int, int);

template <typename T, int K, int G> static cudaError_t poly_q(int Q, int S, const PolyArgs<T>& a, cudaStream_t st) {
    switch (Q) {
        case 1: return launch_poly_n<T, K, 1, G>(S, a, st);
        case 2: return launch_poly_n<T, K, 2, G>(S, a, st);
        case 3: return launch_poly_n<T, K, 3, G>(S, a, st);
        case 4: return launch_poly_n<T, K, 4, G>(S, a, st);
    }
    return cudaErrorInvalidValue;
}
template <typename T, int G> static cudaError_t poly_k(int K, int Q, int S, const PolyArgs<T>& a, cudaStream_t st) {
    switch (K) {
        case 256: return poly_q<T, 256, G>(Q, S, a, st);
        case 512: return poly_q<T, 512, G>(Q, S, a, st);
        case 1024: return poly_q<T, 1024, G>(Q, S, a, st);
    }
    return cudaErrorInvalidValue;
}
template <typename T> cudaError_t launch_poly(int K, int Q, int G, int n_streams, const PolyArgs<T>& a, cudaStream_t st) {
    if (G == 8) return poly_k<T, 8>(K, Q, n_streams, a, st);
    if (G == 10) return poly_k<T, 10>(K, Q, n_streams, a, st);
    return cudaErrorInvalidValue;
}
template cudaError_t launch_poly<float>(int, int, int, int, const PolyArgs<float>&, cudaStream_t);
template cudaError_t launch_poly<double>(int, int, int, int, const PolyArgs<double>&, cudaStream_t);

// ---------------------------------------------------------------------------
// hist2_out[j] = sample (len - 2n + j) of [hist2_in (2n, already mixed) | in (len, mixed here)]
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_hist2_update(const cx<T>* __restrict__ in, long long in_stride, long long len,
                                                      const cx<T>* __restrict__ hist2_in, cx<T>* __restrict__ hist2_out,
                                                      long long n, const NcoStream* __restrict__ nco, long long j_lo, long long j_hi) {
    // four independent loads in flight per thread (coalesced: consecutive threads, consecutive entries), then the stores
    constexpr int UNR = 4;
    const int s = blockIdx.y;
    const cx<T>* src_in = in + (long long)s * in_stride;
    const cx<T>* src_h = hist2_in + (long long)s * 2 * n;
    cx<T>* dst = hist2_out + (long long)s * 2 * n;
    NcoStream ns{};
    if (nco) ns = nco[s];
    const long long j0 = j_lo + (long long)blockIdx.x * (256 * UNR) + threadIdx.x;
    cx<T> v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const long long j = j0 + 256 * u;
        if (j < j_hi) {
            const long long off = len - 2 * n + j;
            v[u] = off >= 0 ? ld_cx(&src_in[off]) : ld_cx(&src_h[off + 2 * n]);
        }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const long long j = j0 + 256 * u;
        if (j < j_hi) {
            const long long off = len - 2 * n + j;
            if (off >= 0 && nco) v[u] = cmul(v[u], nco_phasor_at<T>(off, ns.idx, ns.numer_abs, ns.denom, ns.sign, (T)ns.start_phase));
            st_cx(&dst[j], v[u]);
        }
    }
}
template <typename T>
cudaError_t launch_hist2_update(const void* in, long long in_stride, long long len, const void* hist2_in, void* hist2_out,
                                long long n, const NcoStream* nco, int n_streams, cudaStream_t st, long long j_lo, long long j_hi) {
    if (j_hi < 0 || j_hi > 2 * n) j_hi = 2 * n;
    if (j_lo >= j_hi) return cudaSuccess;
    dim3 grid((unsigned)((j_hi - j_lo + 1023) / 1024), (unsigned)n_streams);
    k_hist2_update<T><<<grid, 256, 0, st>>>(reinterpret_cast<const cx<T>*>(in), in_stride, len,
                                            reinterpret_cast<const cx<T>*>(hist2_in), reinterpret_cast<cx<T>*>(hist2_out), n, nco, j_lo, j_hi);
    return cudaGetLastError();
}
template cudaError_t launch_hist2_update<float>(const void*, long long, long long, const void*, void*, long long, const NcoStream*,
                                                int, cudaStream_t, long long, long long);
template cudaError_t launch_hist2_update<double>(const void*, long long, long long, const void*, void*, long long, const NcoStream*,
                                                 int, cudaStream_t, long long, long long);

}  // namespace rr

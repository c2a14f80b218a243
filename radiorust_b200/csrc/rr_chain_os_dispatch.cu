// Size dispatch for the fused overlap-save chain kernel.
#include "rr_fft_plan.cuh"
#include "rr_kernels.h"

namespace rr {

template <typename T, int N, int EPI>
cudaError_t launch_chain_os_n(int n_streams, int parts, const ChainOsArgs<T>& a, cudaStream_t st);

constexpr size_t kMaxSmem = 227 * 1024;

#define RR_FOR_SIZES_F32(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)
#define RR_FOR_SIZES_F64(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

template <typename T, int N> static size_t smem_needed(int epi, int L) {
    using P = typename PlanFor<T, N>::type;
    size_t s = sizeof(cx<T>) * P::SMEM_ELEMS;
    if (epi) s += sizeof(cx<T>) * (size_t)(N / 2 + L - 1) + sizeof(T) * (size_t)L;
    return s;
}

template <> bool chain_os_supported<float>(int n, int epi, int L) {
    if (epi && (L < 1 || L - 1 > n)) return false;
    switch (2 * n) {
#define X(NN) case NN: return smem_needed<float, NN>(epi, L) <= kMaxSmem;
        RR_FOR_SIZES_F32(X)
#undef X
    }
    return false;
}
template <> bool chain_os_supported<double>(int n, int epi, int L) {
    if (epi && (L < 1 || L - 1 > n)) return false;
    switch (2 * n) {
#define X(NN) case NN: return smem_needed<double, NN>(epi, L) <= kMaxSmem;
        RR_FOR_SIZES_F64(X)
#undef X
    }
    return false;
}

template <> int chain_os_threads<float>(int n) {
    switch (2 * n) {
#define X(NN) case NN: return PlanFor<float, NN>::type::NT;
        RR_FOR_SIZES_F32(X)
#undef X
    }
    return 0;
}
template <> int chain_os_threads<double>(int n) {
    switch (2 * n) {
#define X(NN) case NN: return PlanFor<double, NN>::type::NT;
        RR_FOR_SIZES_F64(X)
#undef X
    }
    return 0;
}

template <> int chain_os_hperm_index<float>(int n, int k) {
    switch (2 * n) {
#define X(NN) case NN: return PlanFor<float, NN>::type::hperm_index(k);
        RR_FOR_SIZES_F32(X)
#undef X
    }
    return -1;
}
template <> int chain_os_hperm_index<double>(int n, int k) {
    switch (2 * n) {
#define X(NN) case NN: return PlanFor<double, NN>::type::hperm_index(k);
        RR_FOR_SIZES_F64(X)
#undef X
    }
    return -1;
}

template <> cudaError_t launch_chain_os<float>(int n, int epi, int n_streams, int parts, const ChainOsArgs<float>& a, cudaStream_t st) {
    switch (2 * n) {
#define X(NN) \
    case NN: return epi ? launch_chain_os_n<float, NN, 1>(n_streams, parts, a, st) : launch_chain_os_n<float, NN, 0>(n_streams, parts, a, st);
        RR_FOR_SIZES_F32(X)
#undef X
    }
    return cudaErrorInvalidValue;
}
template <> cudaError_t launch_chain_os<double>(int n, int epi, int n_streams, int parts, const ChainOsArgs<double>& a, cudaStream_t st) {
    switch (2 * n) {
#define X(NN) \
    case NN: return epi ? launch_chain_os_n<double, NN, 1>(n_streams, parts, a, st) : launch_chain_os_n<double, NN, 0>(n_streams, parts, a, st);
        RR_FOR_SIZES_F64(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

}  // namespace rr

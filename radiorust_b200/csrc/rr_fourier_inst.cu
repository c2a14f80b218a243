// One translation unit per (type, N): compiled with -DRR_T=float -DRR_N=8192.
#include "rr_fourier.cuh"
namespace rr {
template cudaError_t launch_fourier_n<RR_T, RR_N>(const void*, long long, void*, long long, int, int, const RR_T*, const void*, int, cudaStream_t);
}  // namespace rr

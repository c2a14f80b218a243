// One translation unit per (type, N): compiled with -DRR_T=float -DRR_N=8192
// so the heavy fused kernels build in parallel.
#include "rr_chain_os.cuh"
namespace rr {
template cudaError_t launch_chain_os_n<RR_T, RR_N, 0>(int, int, const ChainOsArgs<RR_T>&, cudaStream_t);
template cudaError_t launch_chain_os_n<RR_T, RR_N, 1>(int, int, const ChainOsArgs<RR_T>&, cudaStream_t);
}  // namespace rr

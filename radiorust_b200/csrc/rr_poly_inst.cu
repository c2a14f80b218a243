// One translation unit per (type, K, G): -DRR_T=float -DRR_K=512 -DRR_G=8.
#include "rr_poly.cuh"
namespace rr {
template cudaError_t launch_poly_n<RR_T, RR_K, 1, RR_G>(int, const PolyArgs<RR_T>&, cudaStream_t);
template cudaError_t launch_poly_n<RR_T, RR_K, 2, RR_G>(int, const PolyArgs<RR_T>&, cudaStream_t);
template cudaError_t launch_poly_n<RR_T, RR_K, 3, RR_G>(int, const PolyArgs<RR_T>&, cudaStream_t);
template cudaError_t launch_poly_n<RR_T, RR_K, 4, RR_G>(int, const PolyArgs<RR_T>&, cudaStream_t);
}  // namespace rr

// k_front: rank-reduced front end of the fused chain FreqShifter -> Filter ->
// Downsampler for integer decimation P (rr_design.h: design_rank_tables).
//
// The fused filter of the chain is y_m = sum_l sum_p M[p][l] x'[(m-1-l)P + p]
// (x' = NCO-mixed input, M[p][l] = g[P-1-p + lP], g = Filter taps * reversed
// Downsampler taps).  M = sum_c a_c b_c^T with `RK` real orthonormal a_c to
// below f32 rounding, so
//     u_c[i] = sum_p a_c[p] * x'[iP + p]                 (this kernel)
//     y      = sum_c (b_c * u_c)   at the output rate     (k_poly2 on u, RK branches)
// Every input sample is read once, mixed once (one complex multiply) and feeds
// RK real multiply-adds on the two-wide fp32 pipe; nothing of size P*K is ever
// transformed.  This is the HBM-bound half of the path: 8 B read and 8*RK/P B
// written per input sample.
//
// A warp takes 32 consecutive rows (= 32*P consecutive samples of one stream,
// one TMA box) into its own shared-memory tile; lane = row.  With P/2 odd the
// 16-byte loads of 32 lanes that sit P*8 bytes apart are conflict free.  The
// NCO phasor of sample (i, p) factors into the row phasor (exact per row from
// the integer phase recurrence, transform.rs:333-338, applied to the RK results
// of the row) and exp(j*w*p) (a P-entry table, applied to the sample).
// Rows that reach before the pushed samples (hist2: already mixed) or past them
// are loaded by their lane from global memory instead.
//
// Reference semantics: transform.rs:333-348 (NCO), filters.rs:240-253,
// resampling.rs:103-121.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "rr_kernels.h"
#include "rr_poly.cuh"
#include "rr_pk.cuh"

namespace rr {

namespace {

constexpr int FR_WARPS = 4;
constexpr int FR_ROWS = 32;  // rows per tile = lanes

__device__ __forceinline__ uint32_t f_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RK, bool HAS_NCO>
__global__ void __launch_bounds__(FR_WARPS * 32) k_front(const __grid_constant__ CUtensorMap tmap, const FrontArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.y;
    const int P = a.P;
    const int steps = P / 2;

    extern __shared__ __align__(128) unsigned char smem[];
    const int tile_bytes = FR_ROWS * P * 8;
    const int tile_stride = (tile_bytes + 127) / 128 * 128;
    unsigned char* tile = smem + warp * tile_stride;
    float* coef = reinterpret_cast<float*>(smem + FR_WARPS * tile_stride);  // [steps][2][RK]
    float2* colph = reinterpret_cast<float2*>(coef + steps * 2 * RK);        // [P]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(colph + P);
    const uint32_t bar = f_smem_u32(&bars[warp]);

    uint32_t denom = 1, numer_abs = 0, idx0 = 0;
    int sign = 0;
    float start = 0.f;
    if (HAS_NCO) {
        const NcoStream ns = a.nco[s];
        denom = ns.denom;
        numer_abs = ns.numer_abs;
        sign = ns.sign;
        idx0 = ns.idx;
        start = (float)ns.start_phase;
    }
    for (int e = threadIdx.x; e < steps * 2 * RK; e += FR_WARPS * 32) coef[e] = a.acoef[e];
    for (int p = threadIdx.x; p < P; p += FR_WARPS * 32) {
        cx<float> r(1.f, 0.f);
        if (HAS_NCO) r = nco_rotation<float>(p, numer_abs, denom, sign);
        colph[p] = make_float2(r.x, r.y);
    }
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in) + (long long)s * a.in_stride;
    const float2* __restrict__ hist_end = reinterpret_cast<const float2*>(a.hist2) + ((long long)s + 1) * 2 * a.n;
    float4* __restrict__ u = reinterpret_cast<float4*>(reinterpret_cast<float2*>(a.u) + (long long)s * a.u_stride);
    const long long len = a.len, hist_len = 2 * a.n;
    const uint32_t tile_s = f_smem_u32(tile);
    const uint32_t coef_s = f_smem_u32(coef), col_s = f_smem_u32(colph);

    // rotation of the row phasor from one tile to the next (32 rows)
    pc rot_tile(1.f, 0.f);
    if (HAS_NCO) {
        const cx<float> r = nco_rotation<float>((long long)FR_ROWS * P, numer_abs, denom, sign);
        rot_tile = pc(r.x, r.y);
    }

    const int tile0 = (blockIdx.x * FR_WARPS + warp) * a.tiles_per_warp;
    uint32_t phase = 0;
    pc rowph(1.f, 0.f);
    for (int k = 0; k < a.tiles_per_warp; ++k) {
        const int v0 = (tile0 + k) * FR_ROWS;  // first row (of this push's u rows) of the tile
        if (v0 >= a.n_rows) break;
        const long long pos0 = ((long long)a.row_first + v0) * P - a.J0;  // push offset of the tile's first sample
        const bool interior = pos0 >= 0 && pos0 + (long long)FR_ROWS * P <= len;
        if (interior && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(tile_s),
                         "l"(&tmap), "r"(bar), "r"((int)pos0), "r"(0), "r"(s)
                         : "memory");
        }
        // row phasor: exact at the warp's first tile and every 8th one, rotated in between
        if (HAS_NCO) {
            if ((k & 7) == 0) {
                long long kk = ((long long)idx0 + pos0 + (long long)lane * P) % (long long)denom;
                if (kk < 0) kk += denom;
                const cx<float> c = nco_phasor<float>(mulmod_u32(numer_abs, (uint32_t)kk, denom), denom, sign, start);
                rowph = pc(c.x, c.y);
            } else {
                rowph = pcmul(rowph, rot_tile);
            }
        }
        pc acc[RK];
#pragma unroll
        for (int c = 0; c < RK; ++c) acc[c] = pc(0.f, 0.f);

        if (interior) {
            // wait for the tile
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "FW_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra FD_%=;\n"
                "bra FW_%=;\n"
                "FD_%=:\n"
                "}\n" ::"r"(bar),
                "r"(phase)
                : "memory");
            phase ^= 1;
            const uint32_t row_s = tile_s + lane * (P * 8);
#pragma unroll 5
            for (int st = 0; st < steps; ++st) {
                pc x0, x1;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x1.x), "=f"(x1.y) : "r"(row_s + st * 16));
                if (HAS_NCO) {
                    pc c0, c1;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c0.x), "=f"(c0.y), "=f"(c1.x), "=f"(c1.y) : "r"(col_s + st * 16));
                    x0 = pcmul(x0, c0);
                    x1 = pcmul(x1, c1);
                }
                float cf[2 * RK];
                static_assert((2 * RK) % 4 == 2 || (2 * RK) % 4 == 0, "");
#pragma unroll
                for (int q = 0; q < (2 * RK) / 4; ++q)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(cf[4 * q]), "=f"(cf[4 * q + 1]), "=f"(cf[4 * q + 2]), "=f"(cf[4 * q + 3])
                                 : "r"(coef_s + st * (2 * RK * 4) + q * 16));
                if ((2 * RK) % 4 == 2)
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cf[2 * RK - 2]), "=f"(cf[2 * RK - 1]) : "r"(coef_s + st * (2 * RK * 4) + (2 * RK - 2) * 4));
#pragma unroll
                for (int c = 0; c < RK; ++c) {
                    acc[c] = pfma_s(x0, cf[c], acc[c]);
                    acc[c] = pfma_s(x1, cf[RK + c], acc[c]);
                }
            }
            __syncwarp();  // every lane is done with the tile before the next copy lands in it
        } else {
            // edge rows: straight from global memory; history samples are already mixed, so they only get
            // the conjugate of the row phasor that the row's results receive below
            const long long prow = pos0 + (long long)lane * P;
            const pc hc(rowph.x, -rowph.y);
            for (int p = 0; p < P; ++p) {
                const long long pos = prow + p;
                pc x(0.f, 0.f);
                if (pos >= 0) {
                    if (pos < len) {
                        const float2 q = in[pos];
                        x = pc(q.x, q.y);
                        if (HAS_NCO) {
                            const float2 c = colph[p];
                            x = pcmul(x, pc(c.x, c.y));
                        }
                    }
                } else if (pos >= -hist_len) {
                    const float2 q = hist_end[pos];
                    x = pc(q.x, q.y);
                    if (HAS_NCO) x = pcmul(x, hc);
                }
                const float* cf = coef + (p >> 1) * (2 * RK) + (p & 1) * RK;
#pragma unroll
                for (int c = 0; c < RK; ++c) acc[c] = pfma_s(x, cf[c], acc[c]);
            }
        }
        const int v = v0 + lane;
        if (v < a.n_rows) {
            float4* dst = u + ((long long)v * RK) / 2;
#pragma unroll
            for (int c = 0; c < RK; c += 2) {
                pc y0 = acc[c], y1 = acc[c + 1];
                if (HAS_NCO) {
                    y0 = pcmul(y0, rowph);
                    y1 = pcmul(y1, rowph);
                }
                dst[c / 2] = make_float4(y0.x, y0.y, y1.x, y1.y);
            }
        }
    }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn front_encode_fn() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    });
    return fn;
}

}  // namespace

bool front_supported(int rank_pad, long long P) {
    // lane = row needs P/2 odd (conflict-free 16-byte loads at a pitch of P*8 bytes); one TMA box per tile
    return rank_pad == 10 && P >= 2 && P <= 256 && (P % 4) == 2 && front_encode_fn() != nullptr;
}

cudaError_t launch_front(int rank_pad, int n_streams, const FrontArgs& a0, cudaStream_t st) {
    EncodeFn enc = front_encode_fn();
    if (!enc || rank_pad != 10) return cudaErrorNotSupported;
    FrontArgs a = a0;
    const int tiles = (a.n_rows + FR_ROWS - 1) / FR_ROWS;
    a.tiles_per_warp = 8;
    while (a.tiles_per_warp > 1 && (long long)n_streams * ((tiles + FR_WARPS * a.tiles_per_warp - 1) / (FR_WARPS * a.tiles_per_warp)) < 1184) a.tiles_per_warp /= 2;
    // the stream as overlapping rows of P samples: element (c0, i, s) = in[s*in_stride + c0 + i*P]
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)a.len, (cuuint64_t)FR_ROWS, (cuuint64_t)n_streams};
    const cuuint64_t sstride = n_streams > 1 ? (cuuint64_t)a.in_stride * 8 : (((cuuint64_t)a.len * 8 + 15) / 16) * 16;
    const cuuint64_t strides[2] = {(cuuint64_t)a.P * 8, sstride};
    const cuuint32_t box[3] = {(cuuint32_t)a.P, (cuuint32_t)FR_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(a.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const int tile_stride = (FR_ROWS * a.P * 8 + 127) / 128 * 128;
    const size_t smem = (size_t)FR_WARPS * tile_stride + (size_t)(a.P / 2) * 2 * 10 * 4 + (size_t)a.P * 8 + FR_WARPS * 8 + 16;
    const dim3 grid((unsigned)((tiles + FR_WARPS * a.tiles_per_warp - 1) / (FR_WARPS * a.tiles_per_warp)), (unsigned)n_streams);
    cudaError_t e;
    if (a.nco) {
        auto kern = k_front<10, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, FR_WARPS * 32, smem, st>>>(tm, a);
    } else {
        auto kern = k_front<10, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, FR_WARPS * 32, smem, st>>>(tm, a);
    }
    return cudaGetLastError();
}

}  // namespace rr

// k_front_wide: the rank-reduced front end (rr_front.cu's algebra) for decimation factors k_front's tiles do not take:
// any P (odd, or in the thousands), R = 16 or 32 columns.
//     u_c[i] = sum_p a_c[p] * x'[i*P + p],   x' = the NCO-mixed input,   c < R
// (rr_design.h: design_rank_tables_q; BASELINE config 2: P/Q = 1250/3, rank 21; config 4's de-emphasis Filter +
// Downsampler: 625/3, rank 16).  With P in the thousands this is a tall-and-skinny real x complex matrix product --
// rows i, inner dimension p, R columns -- on the fp32 pipe (not tensor cores: f32 parity at 1e-5 rules out their
// input formats, and the contraction is R wide): a CTA takes TILE rows, walks p in chunks of KC samples staged in shared
// memory (128 contiguous bytes per row and chunk, mixed on the way in), and every thread accumulates TM rows x R/4
// columns in registers with the packed two-wide FFMA2 (one complex sample times one real coefficient each).  Per step
// a thread loads TM samples and R/4 coefficients for TM*R/4 FFMA2, so the fp32 pipe is what binds (R FFMA2 per sample).
//
// NCO: the phasor of sample (i, p) = row phasor (exact from the integer phase recurrence, transform.rs:333-338, applied
// to the row's R results) x exp(j*w*p) (a P-entry table, applied to the sample when it is staged).  Rows that reach
// before the pushed samples read hist2 (already mixed: they get the conjugate of the row phasor instead of the column
// factor), rows that reach past them read zeros (those samples only meet zero entries of the phase matrices).
//
// Reference semantics: transform.rs:333-348 (NCO), filters.rs:240-253, resampling.rs:103-121.  sm_100a.
#include <cuda_runtime.h>

#include <type_traits>

#include "rr_kernels.h"
#include "rr_pk.cuh"
#include "rr_poly.cuh"

namespace rr {

namespace {

constexpr int FW_THREADS = 256;
constexpr int FW_KC = 16;           // samples per chunk
constexpr int FW_PITCH = FW_KC + 2; // tile row pitch in samples (144 B = 9 x 16 B): the eight rows a warp reads at once, two samples
                                    // of a row per 16-byte load, lie on distinct banks

template <int R> constexpr int fw_tm() { return 2; }                        // rows per thread
template <int R> constexpr int fw_tile() { return (FW_THREADS / 4) * fw_tm<R>(); }  // rows per CTA (four column groups)
template <int R> constexpr size_t fw_smem(int P) {
    return (size_t)2 * fw_tile<R>() * FW_PITCH * 8 + (size_t)2 * FW_KC * R * 4 + (size_t)fw_tile<R>() * 8 + (size_t)P * 8;
}

template <int R, bool HAS_NCO>
__global__ void __launch_bounds__(FW_THREADS) k_front_wide(const FrontArgs a) {
    constexpr int TM = fw_tm<R>(), TN = R / 4, TILE = fw_tile<R>(), KC = FW_KC, PITCH = FW_PITCH;
    constexpr int LD = TILE * KC / FW_THREADS;  // samples a thread stages per chunk
    static_assert(TILE * KC % FW_THREADS == 0 && KC * R <= 2 * FW_THREADS, "staging loops");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* As = reinterpret_cast<float2*>(smem_raw);                    // [2][TILE][PITCH]
    float* Bs = reinterpret_cast<float*>(As + 2 * TILE * PITCH);         // [2][KC][R]
    float2* rowph = reinterpret_cast<float2*>(Bs + 2 * KC * R);          // [TILE]
    float2* colph = rowph + TILE;                                        // [P]
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int P = a.P;
    const long long len = a.len, hist_len = 2 * a.n;
    const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in) + (long long)s * a.in_stride;
    const float2* __restrict__ hist_end = reinterpret_cast<const float2*>(a.hist2) + ((long long)s + 1) * 2 * a.n;
    const int row0 = blockIdx.x * TILE;  // first u row of this CTA
    // push offset of sample (row 0 of the tile, p = 0)
    const long long pos_tile = (a.row_first + row0) * (long long)P - a.J0;

    if (HAS_NCO) {
        const NcoStream ns = a.nco[s];
        for (int r = tid; r < TILE; r += FW_THREADS) {
            const cx<float> c = nco_phasor_at<float>(pos_tile + (long long)r * P, ns.idx, ns.numer_abs, ns.denom, ns.sign, (float)ns.start_phase);
            rowph[r] = make_float2(c.x, c.y);
        }
        for (int p = tid; p < P; p += FW_THREADS) {
            const cx<float> c = nco_rotation<float>(p, ns.numer_abs, ns.denom, ns.sign);
            colph[p] = make_float2(c.x, c.y);
        }
    }
    __syncthreads();

    // staging: this thread takes column kk_t of rows r_t + RS*j of every chunk (a half warp per row: 128 contiguous bytes).
    // fetch() only issues the loads -- nothing may depend on them before the chunk in between has been computed --
    // stash() mixes the samples and puts them into the other buffer
    constexpr int RS = FW_THREADS / KC;  // rows between a thread's samples
    const int kk_t = tid % KC, r_t = tid / KC;
    const long long pos_t = pos_tile + (long long)r_t * P + kk_t;  // push offset of (row r_t, column kk_t) of chunk 0
    const long long pos_step = (long long)RS * P;
    float2 pre[LD];
    float preb[2];
    // a tile whose rows all exist and lie inside the pushed samples takes its whole chunks without a bounds check
    const bool tile_inside = pos_tile >= 0 && pos_tile + (long long)TILE * P <= len && row0 + TILE <= a.n_rows;
    // the Filter's next history = the fully mixed samples at push offsets >= hist_from: tiles that reach there write them
    float2* __restrict__ hist_o = a.hist_out ? reinterpret_cast<float2*>(a.hist_out) + (long long)s * a.hist_stride : nullptr;
    const bool tile_hist = hist_o != nullptr && pos_tile + (long long)TILE * P > a.hist_from;
    auto fetch = [&](int k0, auto checked) {
        constexpr bool CHECK = decltype(checked)::value;
        const bool col_ok = !CHECK || k0 + kk_t < P;
        const float2* src0 = in + pos_t + k0;
#pragma unroll
        for (int j = 0; j < LD; ++j) {
            if (CHECK) {
                const long long pos = pos_t + j * pos_step + k0;
                const bool ok = col_ok && (row0 + r_t + RS * j < a.n_rows) && (pos >= 0 ? pos < len : pos >= -hist_len);
                const float2* src = pos >= 0 ? in + pos : hist_end + pos;
                pre[j] = make_float2(0.f, 0.f);
                if (ok) pre[j] = __ldg(src);
            } else {
                pre[j] = __ldg(src0 + j * pos_step);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int e = tid + j * FW_THREADS;
            const int kk = e / R, c = e % R;
            preb[j] = (e < KC * R && (!CHECK || k0 + kk < P)) ? __ldg(a.acoef + (long long)(k0 + kk) * R + c) : 0.f;
        }
    };
    auto stash = [&](int buf, int k0, auto checked) {
        constexpr bool CHECK = decltype(checked)::value;
        float2* A = As + buf * TILE * PITCH;
        pc cp(1.f, 0.f);
        if (HAS_NCO) {
            const float2 c = colph[CHECK ? min(k0 + kk_t, P - 1) : k0 + kk_t];
            cp = pc(c.x, c.y);
        }
#pragma unroll
        for (int j = 0; j < LD; ++j) {
            const int r = r_t + RS * j;
            float2 x = pre[j];
            if (HAS_NCO) {
                // pushed samples get the column factor; history samples are already mixed: they get the conjugate of the
                // row phasor the results are multiplied by
                pc y = pcmul(pc(x.x, x.y), cp);
                if (CHECK && pos_t + j * pos_step + k0 < 0) {
                    const float2 rp = rowph[r];
                    y = pcmulc(pc(x.x, x.y), pc(rp.x, rp.y));
                }
                x = make_float2(y.x, y.y);
            }
            A[r * PITCH + kk_t] = x;
            if (tile_hist) {
                const long long pos = pos_t + j * pos_step + k0;
                if (pos >= a.hist_from && pos >= 0 && pos < len && (!CHECK || (k0 + kk_t < P && row0 + r < a.n_rows))) {
                    float2 h = x;
                    if (HAS_NCO) {
                        const float2 rp = rowph[r];
                        const pc y = pcmul(pc(x.x, x.y), pc(rp.x, rp.y));
                        h = make_float2(y.x, y.y);
                    }
                    hist_o[pos - a.hist_from] = h;
                }
            }
        }
        float* B = Bs + buf * KC * R;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int e = tid + j * FW_THREADS;
            if (e < KC * R) B[e] = preb[j];
        }
    };
    auto fetch_any = [&](int k0) {
        if (tile_inside && k0 + KC <= P) fetch(k0, std::false_type{});
        else fetch(k0, std::true_type{});
    };
    auto stash_any = [&](int buf, int k0) {
        if (tile_inside && k0 + KC <= P) stash(buf, k0, std::false_type{});
        else stash(buf, k0, std::true_type{});
    };

    // thread (column group cg, row group rg): rows rg + (FW_THREADS/4)*m, columns cg*TN .. cg*TN + TN-1
    const int cg = tid & 3, rg = tid >> 2;
    // accumulators, samples and (coefficient, coefficient) pairs as packed 64-bit values: a sample is loaded as one, a
    // coefficient is paired once per step and serves the thread's TM rows
    unsigned long long acc[TM][TN];
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[m][c] = 0ull;

    const int n_chunks = (P + KC - 1) / KC;
    fetch_any(0);
    stash_any(0, 0);
    __syncthreads();
#pragma unroll 1
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < n_chunks) fetch_any((ch + 1) * KC);
        const unsigned long long* A = reinterpret_cast<const unsigned long long*>(As + buf * TILE * PITCH) + rg * PITCH;
        const float* B = Bs + buf * KC * R + cg * TN;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 2) {
            // two samples of a row per 16-byte load
            ulonglong2 xx[TM];
#pragma unroll
            for (int m = 0; m < TM; ++m) xx[m] = *reinterpret_cast<const ulonglong2*>(A + (FW_THREADS / 4) * m * PITCH + kk);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                unsigned long long bp[TN];
#pragma unroll
                for (int c = 0; c < TN; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(B + (kk + h) * R + c);
                    bp[c] = pk2(v.x, v.x);
                    bp[c + 1] = pk2(v.y, v.y);
                    bp[c + 2] = pk2(v.z, v.z);
                    bp[c + 3] = pk2(v.w, v.w);
                }
#pragma unroll
                for (int m = 0; m < TM; ++m) {
                    const unsigned long long x = h ? xx[m].y : xx[m].x;
#pragma unroll
                    for (int c = 0; c < TN; ++c) acc[m][c] = f2_fma(x, bp[c], acc[m][c]);
                }
            }
        }
        if (ch + 1 < n_chunks) stash_any(buf ^ 1, (ch + 1) * KC);
        __syncthreads();
    }

    float2* __restrict__ u = reinterpret_cast<float2*>(a.u) + (long long)s * a.u_stride;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
        const int r = rg + (FW_THREADS / 4) * m;
        const int v = row0 + r;
        if (v >= a.n_rows) continue;
        pc ph(1.f, 0.f);
        if (HAS_NCO) {
            const float2 c = rowph[r];
            ph = pc(c.x, c.y);
        }
        float4* dst = reinterpret_cast<float4*>(u + (long long)v * R + cg * TN);
#pragma unroll
        for (int c = 0; c < TN; c += 2) {
            pc y0 = upk2(acc[m][c]), y1 = upk2(acc[m][c + 1]);
            if (HAS_NCO) {
                y0 = pcmul(y0, ph);
                y1 = pcmul(y1, ph);
            }
            dst[c / 2] = make_float4(y0.x, y0.y, y1.x, y1.y);
        }
    }
}

template <int R> cudaError_t launch_wide_r(int n_streams, const FrontArgs& a, cudaStream_t st) {
    const size_t smem = fw_smem<R>(a.P);
    const dim3 grid((unsigned)((a.n_rows + fw_tile<R>() - 1) / fw_tile<R>()), (unsigned)n_streams);
    cudaError_t e = cudaSuccess;
    auto go = [&](auto kern) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<grid, FW_THREADS, smem, st>>>(a);
    };
    if (a.nco) go(k_front_wide<R, true>);
    else go(k_front_wide<R, false>);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace

// `acoef` of FrontArgs is [P][rank_pad] here (a_c[p] at p*rank_pad + c); kept rows are not taken
bool front_wide_supported(int rank_pad, long long P) {
    if (rank_pad != 16 && rank_pad != 32) return false;
    if (P < 2 || P > 16384) return false;
    return (rank_pad == 16 ? fw_smem<16>((int)P) : fw_smem<32>((int)P)) <= (size_t)200 * 1024;
}
cudaError_t launch_front_wide(int rank_pad, int n_streams, const FrontArgs& a, cudaStream_t st) {
    if (!front_wide_supported(rank_pad, a.P) || a.n_rows < 1 || n_streams > 65535 || a.kept_rows) return cudaErrorNotSupported;
    if (((uintptr_t)a.u % 16) != 0 || (a.u_stride % 2) != 0) return cudaErrorInvalidValue;
    return rank_pad == 16 ? launch_wide_r<16>(n_streams, a, st) : launch_wide_r<32>(n_streams, a, st);
}

}  // namespace rr

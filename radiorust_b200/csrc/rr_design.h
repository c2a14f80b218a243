// Host-side design math (f64), the cold path of the chain: filter response
// (K6), resampler taps (K7), NCO ratio.  Mirrors the reference bit-for-bit in
// operation order where it matters (citations in rr_design.cpp).  No CUDA.
#pragma once
#include <complex>
#include <cstdint>
#include <functional>
#include <vector>

namespace rr {

double bessel_i0(double x);
double kaiser_rel_with_beta(double beta, double x);
double kaiser_null_at_bin_to_beta(double n);
double sinc(double x);
std::complex<double> deemphasis_factor(double tau, double frequency);

// reduced numer/denom of round(denom*f/sr) / round(sr/precision); returns
// false when denom == 0 (Ratio::new would panic)
bool freq_to_ratio(double sample_rate, double precision, double frequency, int64_t* numer, int64_t* denom);

using FreqResp = std::function<std::complex<double>(int64_t bin, double freq)>;
using WindowFn = std::function<double(double x)>;

// In-place power-of-two complex FFT, unnormalised; inverse = conjugate kernel.
void fft_pow2(std::vector<std::complex<double>>& a, bool inverse);
// The same for any length (Bluestein on fft_pow2 when the length is not a power of two).
void fft_any(std::vector<std::complex<double>>& a, bool inverse);

// extended_response (2n bins).  `as_f32`: round the zero-padded impulse
// response to f32 before the last FFT like the reference does for Flt = f32.
// Any n >= 1.
// `taps` (optional) receives the n windowed impulse-response values (already
// rounded to Flt when as_f32): the filter is z[k] = sum_m taps[m] * x[k - m].
bool design_filter_response(const FreqResp& f, const WindowFn& w, double sample_rate, size_t n, bool as_f32,
                            std::vector<std::complex<double>>* out, std::vector<std::complex<double>>* taps = nullptr);

// Polyphase tables of the fused Filter -> Downsampler path (rr_poly.cuh).
// g = h * reverse(ir) (length n + L - 1); for output phase q and input branch p
// the low-rate filter is G[q][p][l] = g[P - 1 + s_q - p + l*P], l = -1 .. Lmax,
// s_q = ceil(q*P/Q); `out` receives FFT_K of each, natural bin order,
// [Q][P][K].  Returns Lmax = floor((n + L - 2) / P).
int design_poly_tables(const std::vector<std::complex<double>>& h, const std::vector<double>& ir, long long P, long long Q, int K,
                       std::vector<std::complex<double>>* out);

// Rank factorisation of the same fused filter for integer decimation (Q == 1):
//   y_m = sum_l sum_p M[p][l] * x'[(m-1-l)*P + p],   M[p][l] = g[P-1-p + l*P],  l = 0 .. Lmax.
// g is a narrow low-pass, heavily oversampled at the input rate, so the P x (Lmax+1) matrix M is
// numerically of low rank: M = sum_c a_c b_c^T with real orthonormal a_c (one-sided Jacobi SVD of
// [Re M | Im M]).  The front end u_c[i] = sum_p a_c[p] * x'[i*P + p] then needs `rank` real
// multiply-adds per input sample and the low-rate part y = sum_c (b_c * u_c) runs on rank (not P)
// branches.  `rank` = the smallest count whose discarded part is below `tol` relative to |M|_F
// (2e-8 for f32: under the rounding of any f32 evaluation of the same sums), 0 if that needs more
// than max_rank.  `a`: [P][rank_pad] real (zero padded to rank_pad columns), `bfft`: [rank_pad][K]
// = FFT_K of b_c (natural bin order).  Returns Lmax.
int design_rank_tables(const std::vector<std::complex<double>>& h, const std::vector<double>& ir, long long P, int K, double tol,
                       int max_rank, int* rank, std::vector<double>* a, std::vector<std::complex<double>>* b, double* discarded);
// The same for in/out = P/Q with Q > 1: output m = Q*I + q fires at j_m = I*P + s_q and is
//   y_m = sum_l sum_p M_q[p][l] * x'[(I-1-l)*P + p],   M_q[p][l] = g[P-1+s_q-p + l*P],  l = -1 .. Lmax.
// The Q matrices are factored TOGETHER, [M_0 | .. | M_{Q-1}] = sum_c a_c [b_{c,0} | .. | b_{c,Q-1}]^T: the front end
// u_c[i] = sum_p a_c[p] * x'[i*P + p] is shared by the phases, y_q = sum_c (b_{c,q} * u_c).  `bfft`: [Q][rank_pad][K],
// tap l = -1 in slot K-1.
int design_rank_tables_q(const std::vector<std::complex<double>>& h, const std::vector<double>& ir, long long P, long long Q, int K, double tol,
                         int max_rank, int* rank, std::vector<double>* a, std::vector<std::complex<double>>* b, double* discarded);

// unit-energy windowed-sinc taps, f64 (cast by the caller)
void design_resampler_taps(size_t ir_len, double ratio, double null_bin, std::vector<double>* out);

// exp(-j*2*pi*e/N), e < N
void make_twiddles(size_t N, std::vector<std::complex<double>>* out);

}  // namespace rr

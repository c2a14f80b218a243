#include "rr_design.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace rr {

// src/math.rs:7-20
double bessel_i0(double x) {
    const double base = x * x / 4.0;
    double addend = 1.0, sum = 1.0;
    for (uint64_t i = 1;; ++i) {
        addend *= base / (double)(i * i);
        const double old = sum;
        sum += addend;
        if (sum == old || !std::isfinite(sum)) break;
    }
    return sum;
}

// src/math.rs:26-28
double kaiser_rel_with_beta(double beta, double x) { return bessel_i0(beta * std::sqrt(1.0 - x * x)); }

// src/math.rs:37-39 (sqrt(n^2 - 1); no pi factor, replicated as is)
double kaiser_null_at_bin_to_beta(double n) { return std::sqrt(n * n - 1.0); }

// src/math.rs:42-49
double sinc(double x) {
    if (x == 0.0) return 1.0;
    const double t = x * M_PI;
    return std::sin(t) / t;
}

// src/blocks/filters.rs:20-27 (Complex::finv = conj / norm_sqr)
std::complex<double> deemphasis_factor(double tau, double frequency) {
    const double re = 1.0, im = tau * (2.0 * M_PI) * frequency;
    const double nrm = re * re + im * im;
    return std::complex<double>(re / nrm, -im / nrm);
}

// src/blocks/transform.rs:298-302 + num Ratio::new
bool freq_to_ratio(double sample_rate, double precision, double frequency, int64_t* numer, int64_t* denom) {
    int64_t d = (int64_t)std::round(sample_rate / precision);  // f64::round: half away from zero
    int64_t n = (int64_t)std::round((double)d * frequency / sample_rate);
    if (d == 0) return false;
    int64_t g = std::gcd(n < 0 ? -n : n, d < 0 ? -d : d);
    if (g == 0) g = 1;
    n /= g;
    d /= g;
    if (d < 0) {
        n = -n;
        d = -d;
    }
    *numer = n;
    *denom = d;
    return true;
}

void make_twiddles(size_t N, std::vector<std::complex<double>>* out) {
    out->resize(N);
    for (size_t e = 0; e < N; ++e) {
        // octant-exact angles through long double keep the table at <= 0.5 ulp(f64)
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)e / (long double)N;
        (*out)[e] = std::complex<double>((double)cosl(a), (double)sinl(a));
    }
}

void fft_pow2(std::vector<std::complex<double>>& a, bool inverse) {
    const size_t n = a.size();
    if (n <= 1) return;
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    std::vector<std::complex<long double>> tw(n / 2);
    for (size_t k = 0; k < n / 2; ++k) {
        const long double ang = (inverse ? 2.0L : -2.0L) * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        tw[k] = std::complex<long double>(cosl(ang), sinl(ang));
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t step = n / len;
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> w((double)tw[k * step].real(), (double)tw[k * step].imag());
                const std::complex<double> u = a[i + k];
                const std::complex<double> v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
        }
    }
}

// Any length: powers of two directly, everything else by Bluestein's chirp-z identity on a power-of-two transform
// (rustfft plans every length too, filters.rs:200,227-228).  The chirp angle is reduced exactly (k*k mod 2n in
// integers) and evaluated in long double, so the result is as accurate as the power-of-two transform itself.
void fft_any(std::vector<std::complex<double>>& a, bool inverse) {
    const size_t n = a.size();
    if (n <= 1) return;
    if ((n & (n - 1)) == 0) {
        fft_pow2(a, inverse);
        return;
    }
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1;
    std::vector<std::complex<double>> chirp(n), A(m, std::complex<double>(0.0, 0.0)), B(m, std::complex<double>(0.0, 0.0));
    const long double pi = 3.14159265358979323846264338327950288L;
    for (size_t k = 0; k < n; ++k) {
        const unsigned long long kk = (unsigned long long)(((unsigned __int128)k * k) % (2ULL * n));
        const long double ang = (inverse ? 1.0L : -1.0L) * pi * (long double)kk / (long double)n;
        chirp[k] = std::complex<double>((double)cosl(ang), (double)sinl(ang));  // exp(-+ j pi k^2 / n)
    }
    for (size_t k = 0; k < n; ++k) A[k] = a[k] * chirp[k];
    B[0] = std::conj(chirp[0]);
    for (size_t k = 1; k < n; ++k) B[k] = B[m - k] = std::conj(chirp[k]);
    fft_pow2(A, false);
    fft_pow2(B, false);
    for (size_t k = 0; k < m; ++k) A[k] *= B[k];
    fft_pow2(A, true);
    const double inv_m = 1.0 / (double)m;
    for (size_t k = 0; k < n; ++k) a[k] = A[k] * inv_m * chirp[k];
}

// src/blocks/filters.rs:184-238
bool design_filter_response(const FreqResp& f, const WindowFn& w, double sample_rate, size_t n, bool as_f32,
                            std::vector<std::complex<double>>* out, std::vector<std::complex<double>>* taps) {
    if (n < 1) return false;
    const double n_flt = (double)n;
    const double scale = 2.0 * n_flt * n_flt;  // :186
    std::vector<std::complex<double>> response(n, std::complex<double>(0.0, 0.0));
    const double freq_step = sample_rate / n_flt;
    const size_t max_bin_abs = (n - 1) / 2;
    for (size_t i = 0; i <= max_bin_abs; ++i) {  // :193-199
        const double freq = (double)i * freq_step;
        response[i] = f((int64_t)i, freq) / scale;
        if (i > 0) response[n - i] = f(-(int64_t)i, -freq) / scale;
    }
    fft_any(response, true);                                                // :200 (unnormalised inverse)
    for (size_t i = 0; i < n / 2; ++i) std::swap(response[i], response[i + n / 2]);  // :201-203
    double energy_pre = 0.0, energy_post = 0.0;
    for (size_t i = 0; i < n; ++i) {  // :204-214
        energy_pre += std::norm(response[i]);
        response[i] *= w(2.0 * ((double)i + 0.5) / n_flt - 1.0);
        energy_post += std::norm(response[i]);
    }
    const double scale2 = std::sqrt(energy_pre / energy_post);  // :216
    for (auto& y : response) y *= scale2;
    out->assign(2 * n, std::complex<double>(0.0, 0.0));  // :220-226
    for (size_t i = 0; i < n; ++i) {
        if (as_f32) (*out)[n + i] = std::complex<double>((double)(float)response[i].real(), (double)(float)response[i].imag());
        else (*out)[n + i] = response[i];
    }
    if (taps) taps->assign(out->begin() + (long)n, out->end());
    fft_any(*out, false);  // :227-238 (the reference runs this one in Flt; f64 here, rounded once by the caller)
    return true;
}

// src/blocks/resampling.rs:84-98 (Downsampler) and :219-233 (Upsampler)
void design_resampler_taps(size_t ir_len, double ratio, double null_bin, std::vector<double>* out) {
    const double ir_len_flt = (double)ir_len;
    const double beta = kaiser_null_at_bin_to_beta(null_bin);
    out->resize(ir_len);
    double energy = 0.0;
    for (size_t i = 0; i < ir_len; ++i) {
        const double x = ((double)i + 0.5) - ir_len_flt / 2.0;
        const double y = sinc(x * ratio) * kaiser_rel_with_beta(beta, x * 2.0 / ir_len_flt);
        (*out)[i] = y;
        energy += y * y;
    }
    const double scale = 1.0 / std::sqrt(energy);
    for (auto& y : *out) y *= scale;
}

int design_poly_tables(const std::vector<std::complex<double>>& h, const std::vector<double>& ir, long long P, long long Q, int K,
                       std::vector<std::complex<double>>* out) {
    const long long n = (long long)h.size(), L = (long long)ir.size();
    const long long Lg = n + L - 1;
    // g[tau] = sum_v ir[L-1-v] * h[tau - v]
    std::vector<std::complex<double>> g((size_t)Lg, std::complex<double>(0.0, 0.0));
    for (long long v = 0; v < L; ++v) {
        const double c = ir[(size_t)(L - 1 - v)];
        if (c == 0.0) continue;
        std::complex<double>* dst = g.data() + v;
        for (long long m = 0; m < n; ++m) dst[m] += c * h[(size_t)m];
    }
    const int Lmax = (int)((Lg - 1) / P);
    out->assign((size_t)(Q * P) * (size_t)K, std::complex<double>(0.0, 0.0));
    std::vector<std::complex<double>> buf((size_t)K);
    for (long long q = 0; q < Q; ++q) {
        const long long sq = (q * P + Q - 1) / Q;
        for (long long p = 0; p < P; ++p) {
            std::fill(buf.begin(), buf.end(), std::complex<double>(0.0, 0.0));
            for (long long l = -1; l <= Lmax; ++l) {
                const long long idx = P - 1 + sq - p + l * P;
                if (idx < 0 || idx >= Lg) continue;
                buf[(size_t)((l + K) % K)] = g[(size_t)idx];
            }
            fft_pow2(buf, false);
            std::copy(buf.begin(), buf.end(), out->begin() + (size_t)(q * P + p) * (size_t)K);
        }
    }
    return Lmax;
}

namespace {
// one-sided Jacobi (Hestenes) on the columns of W (rows x nc, row-major): afterwards the columns are orthogonal,
// W_final = W * V with V (nc x nc) the accumulated rotations
void hestenes_columns(std::vector<double>& W, std::vector<double>& V, int rows, int nc) {
    V.assign((size_t)nc * nc, 0.0);
    for (int p = 0; p < nc; ++p) V[(size_t)p * nc + p] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < nc - 1; ++i)
            for (int j = i + 1; j < nc; ++j) {
                double aii = 0.0, ajj = 0.0, aij = 0.0;
                for (int r = 0; r < rows; ++r) {
                    const double wi = W[(size_t)r * nc + i], wj = W[(size_t)r * nc + j];
                    aii += wi * wi;
                    ajj += wj * wj;
                    aij += wi * wj;
                }
                if (aij == 0.0 || std::fabs(aij) <= 1e-17 * std::sqrt(aii * ajj)) continue;
                off = std::max(off, std::fabs(aij) / std::sqrt(aii * ajj));
                const double zeta = (ajj - aii) / (2.0 * aij);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int r = 0; r < rows; ++r) {
                    const double wi = W[(size_t)r * nc + i], wj = W[(size_t)r * nc + j];
                    W[(size_t)r * nc + i] = cs * wi - sn * wj;
                    W[(size_t)r * nc + j] = sn * wi + cs * wj;
                }
                for (int r = 0; r < nc; ++r) {
                    const double vi = V[(size_t)r * nc + i], vj = V[(size_t)r * nc + j];
                    V[(size_t)r * nc + i] = cs * vi - sn * vj;
                    V[(size_t)r * nc + j] = sn * vi + cs * vj;
                }
            }
        if (off < 1e-15) break;
    }
}

// Row space of W (rows x np, row-major) by pivoted Gram-Schmidt (each new vector re-orthogonalised): W = C * Qb + R with
// Qb (r x np) orthonormal rows and |R|_F^2 = *resid <= stop2.  Cheap when the numerical rank is far below both dimensions
// (the fused filter of a long decimation: P in the thousands, rank in the twenties).
int row_space(const std::vector<double>& W, int rows, int np, int rmax, double stop2, std::vector<double>* Qb, std::vector<double>* C, double* resid) {
    std::vector<double> R = W, rn((size_t)rows, 0.0);
    double res = 0.0;
    for (int i = 0; i < rows; ++i) {
        double t = 0.0;
        for (int p = 0; p < np; ++p) t += R[(size_t)i * np + p] * R[(size_t)i * np + p];
        rn[(size_t)i] = t;
        res += t;
    }
    Qb->assign((size_t)rmax * np, 0.0);
    C->assign((size_t)rows * rmax, 0.0);
    int r = 0;
    std::vector<double> q((size_t)np);
    while (r < rmax && res > stop2) {
        int best = 0;
        for (int i = 1; i < rows; ++i)
            if (rn[(size_t)i] > rn[(size_t)best]) best = i;
        if (!(rn[(size_t)best] > 0.0)) break;
        for (int p = 0; p < np; ++p) q[(size_t)p] = R[(size_t)best * np + p];
        for (int pass = 0; pass < 2; ++pass)
            for (int k = 0; k < r; ++k) {
                double d = 0.0;
                for (int p = 0; p < np; ++p) d += q[(size_t)p] * (*Qb)[(size_t)k * np + p];
                for (int p = 0; p < np; ++p) q[(size_t)p] -= d * (*Qb)[(size_t)k * np + p];
            }
        double nq = 0.0;
        for (int p = 0; p < np; ++p) nq += q[(size_t)p] * q[(size_t)p];
        if (!(nq > 0.0)) break;
        nq = std::sqrt(nq);
        for (int p = 0; p < np; ++p) (*Qb)[(size_t)r * np + p] = q[(size_t)p] / nq;
        res = 0.0;
        for (int i = 0; i < rows; ++i) {
            double d = 0.0;
            const double* qb = Qb->data() + (size_t)r * np;
            double* ri = R.data() + (size_t)i * np;
            for (int p = 0; p < np; ++p) d += ri[p] * qb[p];
            (*C)[(size_t)i * rmax + r] = d;
            double t = 0.0;
            for (int p = 0; p < np; ++p) {
                ri[p] -= d * qb[p];
                t += ri[p] * ri[p];
            }
            rn[(size_t)i] = t;
            res += t;
        }
        ++r;
    }
    *resid = res;
    return r;
}
}  // namespace

int design_rank_tables(const std::vector<std::complex<double>>& h, const std::vector<double>& ir, long long P, int K, double tol,
                       int max_rank, int* rank, std::vector<double>* a, std::vector<std::complex<double>>* b, double* discarded) {
    const long long n = (long long)h.size(), L = (long long)ir.size();
    const long long Lg = n + L - 1;
    std::vector<std::complex<double>> g((size_t)Lg, std::complex<double>(0.0, 0.0));
    for (long long v = 0; v < L; ++v) {
        const double c = ir[(size_t)(L - 1 - v)];
        if (c == 0.0) continue;
        std::complex<double>* dst = g.data() + v;
        for (long long m = 0; m < n; ++m) dst[m] += c * h[(size_t)m];
    }
    const int Lmax = (int)((Lg - 1) / P);
    const int nl = Lmax + 1, np = (int)P;
    *rank = 0;
    if (nl >= K) return Lmax;
    // W = [Re M^T ; Im M^T]: 2*nl rows, P columns (column p = branch p); Hestenes rotations make the
    // columns orthogonal, V accumulates them: W_final = W * V, column c = (b_c re | b_c im), a_c = V[:, c]
    const int rows = 2 * nl;
    std::vector<double> W((size_t)rows * np, 0.0), V;
    for (int p = 0; p < np; ++p)
        for (int l = 0; l < nl; ++l) {
            const long long idx = P - 1 - p + (long long)l * P;
            if (idx < 0 || idx >= Lg) continue;
            W[(size_t)l * np + p] = g[(size_t)idx].real();
            W[(size_t)(nl + l) * np + p] = g[(size_t)idx].imag();
        }
    double total = 0.0;
    for (double w : W) total += w * w;
    if (!(total > 0.0)) return Lmax;
    hestenes_columns(W, V, rows, np);
    std::vector<double> sig2((size_t)np, 0.0);
    std::vector<int> order((size_t)np);
    for (int c = 0; c < np; ++c) {
        order[(size_t)c] = c;
        for (int r = 0; r < rows; ++r) sig2[(size_t)c] += W[(size_t)r * np + c] * W[(size_t)r * np + c];
    }
    std::sort(order.begin(), order.end(), [&](int x, int y) { return sig2[(size_t)x] > sig2[(size_t)y]; });
    // smallest rank whose tail is below tol
    int rk = np;
    double tail = 0.0;
    for (int c = np - 1; c >= 0; --c) {
        const double t2 = tail + sig2[(size_t)order[(size_t)c]];
        if (std::sqrt(t2 / total) > tol) break;
        tail = t2;
        rk = c;
    }
    if (discarded) *discarded = std::sqrt(tail / total);
    if (rk < 1) rk = 1;
    if (rk > max_rank) return Lmax;
    *rank = rk;
    a->assign((size_t)np * (size_t)max_rank, 0.0);
    b->assign((size_t)max_rank * (size_t)K, std::complex<double>(0.0, 0.0));
    std::vector<std::complex<double>> buf((size_t)K);
    for (int c = 0; c < rk; ++c) {
        const int col = order[(size_t)c];
        for (int p = 0; p < np; ++p) (*a)[(size_t)p * max_rank + c] = V[(size_t)p * np + col];
        std::fill(buf.begin(), buf.end(), std::complex<double>(0.0, 0.0));
        for (int l = 0; l < nl; ++l) buf[(size_t)l] = std::complex<double>(W[(size_t)l * np + col], W[(size_t)(nl + l) * np + col]);
        fft_pow2(buf, false);
        std::copy(buf.begin(), buf.end(), b->begin() + (size_t)c * (size_t)K);
    }
    return Lmax;
}

// The Q > 1 form: the Q phase matrices M_q[p][l] = g[P-1+s_q-p + l*P], l = -1 .. Lmax, side by side share one set of
// a_c (the front end does not know the phase); the low-rate part has one b_{c,q} per phase.
int design_rank_tables_q(const std::vector<std::complex<double>>& h, const std::vector<double>& ir, long long P, long long Q, int K, double tol,
                         int max_rank, int* rank, std::vector<double>* a, std::vector<std::complex<double>>* b, double* discarded) {
    const long long n = (long long)h.size(), L = (long long)ir.size();
    const long long Lg = n + L - 1;
    std::vector<std::complex<double>> g((size_t)Lg, std::complex<double>(0.0, 0.0));
    for (long long v = 0; v < L; ++v) {
        const double c = ir[(size_t)(L - 1 - v)];
        if (c == 0.0) continue;
        std::complex<double>* dst = g.data() + v;
        for (long long m = 0; m < n; ++m) dst[m] += c * h[(size_t)m];
    }
    const int Lmax = (int)((Lg - 1) / P);
    const int nl = Lmax + 2, np = (int)P, nq = (int)Q;  // slot l + 1 holds tap l
    *rank = 0;
    if (nl > K) return Lmax;
    // W: per phase q the rows [Re M_q^T ; Im M_q^T] (2*nl rows of P columns)
    const int rows = 2 * nl * nq;
    std::vector<double> W((size_t)rows * np, 0.0);
    for (int p = 0; p < np; ++p)
        for (int q = 0; q < nq; ++q) {
            const long long sq = ((long long)q * P + Q - 1) / Q;
            for (int l = -1; l <= Lmax; ++l) {
                const long long idx = P - 1 + sq - p + (long long)l * P;
                if (idx < 0 || idx >= Lg) continue;
                W[(size_t)(q * 2 * nl + l + 1) * np + p] = g[(size_t)idx].real();
                W[(size_t)(q * 2 * nl + nl + l + 1) * np + p] = g[(size_t)idx].imag();
            }
        }
    double total = 0.0;
    for (double w : W) total += w * w;
    if (!(total > 0.0)) return Lmax;
    // Many branches (P well above the rank): first the row space of W by pivoted Gram-Schmidt, W = C * Qb + R with
    // |R| a tenth of the tolerance, then the small C (rows x r) takes W's place; a_c = Qb^T v_c.  Few branches: W itself.
    std::vector<double> Qb, V;
    int nc = np;
    double resid = 0.0;
    const bool compress = np > 128;
    if (compress) {
        std::vector<double> C;
        const int rmax = std::min(std::min(rows, np), 4 * max_rank + 16);
        const double stop = 0.1 * tol;
        const int r = row_space(W, rows, np, rmax, stop * stop * total, &Qb, &C, &resid);
        if (r < 1) return Lmax;
        std::vector<double> W2((size_t)rows * r);
        for (int i = 0; i < rows; ++i)
            for (int k = 0; k < r; ++k) W2[(size_t)i * r + k] = C[(size_t)i * rmax + k];
        W.swap(W2);
        nc = r;
    }
    // Hestenes rotations make the columns orthogonal, V accumulates them: W_final = W * V,
    // column c = (b_{c,q} re | b_{c,q} im)_q, a_c = V[:, c] (through Qb when compressed)
    hestenes_columns(W, V, rows, nc);
    std::vector<double> sig2((size_t)nc, 0.0);
    std::vector<int> order((size_t)nc);
    for (int c = 0; c < nc; ++c) {
        order[(size_t)c] = c;
        for (int r = 0; r < rows; ++r) sig2[(size_t)c] += W[(size_t)r * nc + c] * W[(size_t)r * nc + c];
    }
    std::sort(order.begin(), order.end(), [&](int x, int y) { return sig2[(size_t)x] > sig2[(size_t)y]; });
    int rk = nc;
    double tail = resid;
    if (std::sqrt(tail / total) > tol) return Lmax;
    for (int c = nc - 1; c >= 0; --c) {
        const double t2 = tail + sig2[(size_t)order[(size_t)c]];
        if (std::sqrt(t2 / total) > tol) break;
        tail = t2;
        rk = c;
    }
    if (discarded) *discarded = std::sqrt(tail / total);
    if (rk < 1) rk = 1;
    if (rk > max_rank) return Lmax;
    *rank = rk;
    a->assign((size_t)np * (size_t)max_rank, 0.0);
    b->assign((size_t)nq * (size_t)max_rank * (size_t)K, std::complex<double>(0.0, 0.0));
    std::vector<std::complex<double>> buf((size_t)K);
    for (int c = 0; c < rk; ++c) {
        const int col = order[(size_t)c];
        if (compress) {
            for (int k = 0; k < nc; ++k) {
                const double v = V[(size_t)k * nc + col];
                for (int p = 0; p < np; ++p) (*a)[(size_t)p * max_rank + c] += v * Qb[(size_t)k * np + p];
            }
        } else {
            for (int p = 0; p < np; ++p) (*a)[(size_t)p * max_rank + c] = V[(size_t)p * np + col];
        }
        for (int q = 0; q < nq; ++q) {
            std::fill(buf.begin(), buf.end(), std::complex<double>(0.0, 0.0));
            for (int l = -1; l <= Lmax; ++l)
                buf[(size_t)((l + K) % K)] =
                    std::complex<double>(W[(size_t)(q * 2 * nl + l + 1) * nc + col], W[(size_t)(q * 2 * nl + nl + l + 1) * nc + col]);
            fft_pow2(buf, false);
            std::copy(buf.begin(), buf.end(), b->begin() + ((size_t)q * max_rank + c) * (size_t)K);
        }
    }
    return Lmax;
}

}  // namespace rr

/* C restatement of the hot loops of radiorust's IQ sample chain.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call it.
 * The product (radiorust_b200/) never links or executes anything under oracle/.
 *
 * It follows the reference loops line by line (citations per function) for the
 * three blocks of the benchmarked chain, complex f32 and f64:
 *   FreqShifter  src/blocks/transform.rs:321-348   (phase table + per-sample multiply)
 *   Filter       src/blocks/filters.rs:240-260     (overlap-save: FFT(2n) * H, IFFT(2n), first n)
 *   Downsampler  src/blocks/resampling.rs:103-133  (ring buffer + L-tap dot product)
 * The cold design paths (H(f), taps) are NOT here: callers pass the tables that
 * oracle/radiorust_oracle.py designed (filters.rs:184-238, resampling.rs:82-101).
 * rustfft (Cargo.toml:19, absent from /root/reference) is replaced by the
 * radix-4 Stockham (autosort) FFT below, compiled in AVX-512 / AVX2 / baseline
 * clones that the loader picks at run time (rustfft picks AVX/SSE butterflies the
 * same way): both are unnormalised DFTs, so results agree to rounding.  Parity status: validated against the numpy oracle in
 * tests/test_oracle_c.py; end-to-end outputs of these blocks are "parity
 * unpinned" upstream (the reference ships no tests for them).
 *
 * Threads: streams are independent; oracle_chain_* runs `n_threads` pthreads over
 * the stream list (the reference itself runs one Tokio task per block).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define CLONES
#endif

#define DEFINE_ORACLE(T, SUF, SIN, COS, TAU_CONST)                                                          \
    typedef struct { T re, im; } cx_##SUF;                                                                  \
                                                                                                            \
    /* radix-4 Stockham autosort FFT (a closing radix-2 pass when log2 n is odd), unnormalised, result in `a`;  */   \
    /* tw[k] = exp(-+ j 2 pi k / n) for k < n, `b` is scratch of n elements; INV selects the sign of j           */   \
    CLONES static void fft_##SUF(cx_##SUF* a, cx_##SUF* b, size_t n, const cx_##SUF* tw, int inv) {                   \
        cx_##SUF *x = a, *y = b;                                                                            \
        size_t m = n, s = 1;                                                                                \
        for (; m >= 4; m >>= 2, s <<= 2) {                                                                  \
            const size_t m1 = m >> 2, tstep = n / m;                                                        \
            for (size_t p = 0; p < m1; ++p) {                                                               \
                const cx_##SUF w1 = tw[p * tstep], w2 = tw[2 * p * tstep], w3 = tw[3 * p * tstep];          \
                const cx_##SUF* xa = x + s * p;                                                             \
                const cx_##SUF* xb = x + s * (p + m1);                                                      \
                const cx_##SUF* xc = x + s * (p + 2 * m1);                                                  \
                const cx_##SUF* xd = x + s * (p + 3 * m1);                                                  \
                cx_##SUF* y0 = y + s * (4 * p);                                                             \
                cx_##SUF* y1 = y0 + s;                                                                      \
                cx_##SUF* y2 = y1 + s;                                                                      \
                cx_##SUF* y3 = y2 + s;                                                                      \
                for (size_t q = 0; q < s; ++q) {                                                            \
                    const T apc_r = xa[q].re + xc[q].re, apc_i = xa[q].im + xc[q].im;                       \
                    const T amc_r = xa[q].re - xc[q].re, amc_i = xa[q].im - xc[q].im;                       \
                    const T bpd_r = xb[q].re + xd[q].re, bpd_i = xb[q].im + xd[q].im;                       \
                    const T bmd_r = xb[q].re - xd[q].re, bmd_i = xb[q].im - xd[q].im;                       \
                    /* j*(b-d) forward: (-im, re); inverse: (im, -re) */                                    \
                    const T jr = inv ? bmd_i : -bmd_i, ji = inv ? -bmd_r : bmd_r;                           \
                    const T t1r = amc_r - jr, t1i = amc_i - ji;                                             \
                    const T t2r = apc_r - bpd_r, t2i = apc_i - bpd_i;                                       \
                    const T t3r = amc_r + jr, t3i = amc_i + ji;                                             \
                    y0[q].re = apc_r + bpd_r; y0[q].im = apc_i + bpd_i;                                     \
                    y1[q].re = t1r * w1.re - t1i * w1.im; y1[q].im = t1r * w1.im + t1i * w1.re;             \
                    y2[q].re = t2r * w2.re - t2i * w2.im; y2[q].im = t2r * w2.im + t2i * w2.re;             \
                    y3[q].re = t3r * w3.re - t3i * w3.im; y3[q].im = t3r * w3.im + t3i * w3.re;             \
                }                                                                                           \
            }                                                                                               \
            cx_##SUF* t = x; x = y; y = t;                                                                  \
        }                                                                                                   \
        if (m == 2) {                                                                                       \
            for (size_t q = 0; q < s; ++q) {                                                                \
                const cx_##SUF u = x[q], v = x[q + s];                                                      \
                y[q].re = u.re + v.re; y[q].im = u.im + v.im;                                               \
                y[q + s].re = u.re - v.re; y[q + s].im = u.im - v.im;                                       \
            }                                                                                               \
            cx_##SUF* t = x; x = y; y = t;                                                                  \
        }                                                                                                   \
        if (x != a) memcpy(a, x, sizeof(cx_##SUF) * n);                                                     \
    }                                                                                                       \
                                                                                                            \
    typedef struct {                                                                                        \
        /* chain parameters (shared, read-only) */                                                          \
        size_t n, n_chunks, L;                                                                              \
        const cx_##SUF* hext; /* 2n */                                                                      \
        const T* ir;          /* L  */                                                                      \
        double in_rate, out_rate;                                                                           \
        int with_nco, with_filter, with_down;                                                               \
        /* per stream */                                                                                    \
        const cx_##SUF* const* x; /* [n_streams] -> n*n_chunks samples */                                   \
        cx_##SUF* const* y;       /* [n_streams] -> output buffers */                                       \
        size_t* n_out;            /* [n_streams] */                                                         \
        const int64_t* numer;     /* reduced ratio per stream (transform.rs:298-302) */                     \
        const int64_t* denom;                                                                               \
        const cx_##SUF* const* phase_tabs; /* optional: prebuilt phase tables (the reference builds its table  */ \
                                           /* once per retune, transform.rs:321-340, outside the steady state) */ \
        size_t s_begin, s_end;                                                                              \
        int status;                                                                                         \
    } job_##SUF;                                                                                            \
                                                                                                            \
    /* FreqShifter's phase table of `denom` entries (transform.rs:321-340) */                               \
    void oracle_phase_table_##SUF(int64_t numer, int64_t denom, T* out) {                                   \
        cx_##SUF* phase = (cx_##SUF*)out;                                                                   \
        int64_t i = 0;                                                                                      \
        for (int64_t k = 0; k < denom; ++k) {                                                               \
            const T ph = (T)0 + (T)i / (T)denom * (T)TAU_CONST; /* transform.rs:335 */                      \
            phase[k].re = COS(ph); phase[k].im = SIN(ph);                                                   \
            i = (i + numer) % denom; /* C '%' truncates like Rust's */                                      \
        }                                                                                                   \
    }                                                                                                       \
                                                                                                            \
    CLONES static void fir_##SUF(const cx_##SUF* ring, size_t ring_pos, size_t L, const T* ir, cx_##SUF* out) { \
        /* resampling.rs:112-120: L-tap dot product over the ring, oldest sample first */                    \
        T sr = 0, si = 0;                                                                                   \
        size_t q = 0;                                                                                       \
        for (size_t i = ring_pos; i < L; ++i, ++q) { sr += ring[i].re * ir[q]; si += ring[i].im * ir[q]; } \
        for (size_t i = 0; i < ring_pos; ++i, ++q) { sr += ring[i].re * ir[q]; si += ring[i].im * ir[q]; } \
        out->re = sr; out->im = si;                                                                         \
    }                                                                                                       \
                                                                                                            \
    static void* worker_##SUF(void* arg) {                                                                  \
        job_##SUF* jb = (job_##SUF*)arg;                                                                    \
        const size_t n = jb->n, N = 2 * n, L = jb->L;                                                       \
        cx_##SUF* buf = (cx_##SUF*)malloc(sizeof(cx_##SUF) * N);                                            \
        cx_##SUF* scr = (cx_##SUF*)malloc(sizeof(cx_##SUF) * N);                                            \
        cx_##SUF* twf = (cx_##SUF*)malloc(sizeof(cx_##SUF) * (N + 1));                                      \
        cx_##SUF* twi = (cx_##SUF*)malloc(sizeof(cx_##SUF) * (N + 1));                                      \
        cx_##SUF* mixed = (cx_##SUF*)malloc(sizeof(cx_##SUF) * n);                                          \
        cx_##SUF* prev = (cx_##SUF*)malloc(sizeof(cx_##SUF) * n);                                           \
        cx_##SUF* ring = (cx_##SUF*)calloc(L ? L : 1, sizeof(cx_##SUF));                                    \
        if (!buf || !scr || !twf || !twi || !mixed || !prev || !ring) { jb->status = -1; return NULL; }     \
        for (size_t k = 0; k < N; ++k) {                                                                \
            const double a = -2.0 * M_PI * (double)k / (double)N;                                           \
            twf[k].re = (T)cos(a); twf[k].im = (T)sin(a);                                                   \
            twi[k].re = (T)cos(a); twi[k].im = (T)(-sin(a));                                                \
        }                                                                                                   \
        for (size_t s = jb->s_begin; s < jb->s_end; ++s) {                                                  \
            const cx_##SUF* x = jb->x[s];                                                                   \
            cx_##SUF* y = jb->y[s];                                                                         \
            size_t produced = 0;                                                                            \
            /* FreqShifter: phase table of `denom` entries (transform.rs:321-340) */                        \
            const cx_##SUF* phase = NULL;                                                                   \
            cx_##SUF* phase_own = NULL;                                                                     \
            size_t plen = 0, pidx = 0;                                                                      \
            if (jb->with_nco) {                                                                             \
                plen = (size_t)jb->denom[s];                                                                \
                if (jb->phase_tabs) {                                                                       \
                    phase = jb->phase_tabs[s];                                                              \
                } else {                                                                                    \
                    phase_own = (cx_##SUF*)malloc(sizeof(cx_##SUF) * plen);                                 \
                    if (!phase_own) { jb->status = -1; break; }                                             \
                    oracle_phase_table_##SUF(jb->numer[s], jb->denom[s], (T*)phase_own);                    \
                    phase = phase_own;                                                                      \
                }                                                                                           \
            }                                                                                               \
            int have_prev = 0;                                                                              \
            size_t ring_pos = 0;                                                                            \
            double pos = 0.0;                                                                               \
            memset(ring, 0, sizeof(cx_##SUF) * (L ? L : 1));                                                \
            for (size_t c = 0; c < jb->n_chunks; ++c) {                                                     \
                const cx_##SUF* in = x + c * n;                                                             \
                /* transform.rs:341-348 */                                                                  \
                if (jb->with_nco) {                                                                         \
                    for (size_t t = 0; t < n; ++t) {                                                        \
                        const cx_##SUF p = phase[pidx];                                                     \
                        mixed[t].re = in[t].re * p.re - in[t].im * p.im;                                    \
                        mixed[t].im = in[t].re * p.im + in[t].im * p.re;                                    \
                        if (++pidx == plen) pidx = 0;                                                       \
                    }                                                                                       \
                    in = mixed;                                                                             \
                }                                                                                           \
                const cx_##SUF* z = in;                                                                     \
                size_t zn = n;                                                                              \
                if (jb->with_filter) {                                                                      \
                    if (!have_prev) { /* filters.rs:240,260: first chunk only primes the history */         \
                        memcpy(prev, in, sizeof(cx_##SUF) * n);                                             \
                        have_prev = 1;                                                                      \
                        continue;                                                                           \
                    }                                                                                       \
                    memcpy(buf, prev, sizeof(cx_##SUF) * n);       /* filters.rs:241-243 */                 \
                    memcpy(buf + n, in, sizeof(cx_##SUF) * n);                                              \
                    memcpy(prev, in, sizeof(cx_##SUF) * n);        /* filters.rs:260 */                     \
                    fft_##SUF(buf, scr, N, twf, 0);                /* filters.rs:244-246 */                 \
                    for (size_t k = 0; k < N; ++k) {               /* filters.rs:247-249 */                 \
                        const cx_##SUF h = jb->hext[k], v = buf[k];                                         \
                        buf[k].re = v.re * h.re - v.im * h.im;                                              \
                        buf[k].im = v.re * h.im + v.im * h.re;                                              \
                    }                                                                                       \
                    fft_##SUF(buf, scr, N, twi, 1);                /* filters.rs:250-252 */                 \
                    z = buf;                                       /* truncate(n), filters.rs:253 */        \
                }                                                                                           \
                if (jb->with_down) {                                                                        \
                    /* resampling.rs:103-121 */                                                             \
                    for (size_t t = 0; t < zn; ++t) {                                                       \
                        ring[ring_pos] = z[t];                                                              \
                        if (++ring_pos == L) ring_pos = 0;                                                  \
                        pos += jb->out_rate;                                                                \
                        if (pos >= jb->in_rate) {                                                           \
                            pos -= jb->in_rate;                                                             \
                            fir_##SUF(ring, ring_pos, L, jb->ir, &y[produced]);                             \
                            ++produced;                                                                     \
                        }                                                                                   \
                    }                                                                                       \
                } else {                                                                                    \
                    memcpy(y + produced, z, sizeof(cx_##SUF) * zn);                                         \
                    produced += zn;                                                                         \
                }                                                                                           \
            }                                                                                               \
            jb->n_out[s] = produced;                                                                        \
            free(phase_own);                                                                                \
        }                                                                                                   \
        free(buf); free(scr); free(twf); free(twi); free(mixed); free(prev); free(ring);                               \
        return NULL;                                                                                        \
    }                                                                                                       \
                                                                                                            \
    /* Runs [FreqShifter ->] [Filter ->] [Downsampler] over n_streams independent streams, each fed   */   \
    /* n_chunks chunks of n samples from a fresh start (Downsampler output framing is not modelled:  */   \
    /* samples are returned concatenated).  Returns 0 on success.                                    */   \
    int oracle_chain_##SUF(size_t n_streams, size_t n, size_t n_chunks, const T* const* x, T* const* y, size_t* n_out, \
                           int with_nco, const int64_t* numer, const int64_t* denom, int with_filter, const T* hext,    \
                           int with_down, const T* ir, size_t L, double in_rate, double out_rate, int n_threads,        \
                           const T* const* phase_tabs) {                                                                \
        if (n_threads < 1) n_threads = 1;                                                                   \
        if ((size_t)n_threads > n_streams) n_threads = (int)(n_streams ? n_streams : 1);                    \
        if (with_filter && (n < 2 || (n & (n - 1)))) return -2;                                             \
        job_##SUF* jobs = (job_##SUF*)calloc((size_t)n_threads, sizeof(job_##SUF));                         \
        pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));                           \
        if (!jobs || !th) return -1;                                                                        \
        for (int i = 0; i < n_threads; ++i) {                                                               \
            job_##SUF* jb = &jobs[i];                                                                       \
            jb->n = n; jb->n_chunks = n_chunks; jb->L = with_down ? L : 0;                                  \
            jb->hext = (const cx_##SUF*)hext; jb->ir = ir;                                                  \
            jb->in_rate = in_rate; jb->out_rate = out_rate;                                                 \
            jb->with_nco = with_nco; jb->with_filter = with_filter; jb->with_down = with_down;              \
            jb->x = (const cx_##SUF* const*)x; jb->y = (cx_##SUF* const*)y; jb->n_out = n_out;              \
            jb->numer = numer; jb->denom = denom; jb->phase_tabs = (const cx_##SUF* const*)phase_tabs;      \
            jb->s_begin = n_streams * (size_t)i / (size_t)n_threads;                                        \
            jb->s_end = n_streams * (size_t)(i + 1) / (size_t)n_threads;                                    \
            jb->status = 0;                                                                                 \
            if (i + 1 < n_threads) pthread_create(&th[i], NULL, worker_##SUF, jb);                          \
        }                                                                                                   \
        worker_##SUF(&jobs[n_threads - 1]);                                                                 \
        int status = jobs[n_threads - 1].status;                                                            \
        for (int i = 0; i + 1 < n_threads; ++i) { pthread_join(th[i], NULL); if (jobs[i].status) status = jobs[i].status; } \
        free(jobs); free(th);                                                                               \
        return status;                                                                                      \
    }

DEFINE_ORACLE(float, f32, sinf, cosf, 6.283185307179586476925286766559f)
DEFINE_ORACLE(double, f64, sin, cos, 6.283185307179586476925286766559)

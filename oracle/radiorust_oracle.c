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
 * rustfft (Cargo.toml:19, absent from /root/reference) is replaced by the plain
 * iterative radix-2 FFT below: both are unnormalised DFTs, so results agree to
 * rounding.  Parity status: validated against the numpy oracle in
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

#define DEFINE_ORACLE(T, SUF, SIN, COS, TAU_CONST)                                                          \
    typedef struct { T re, im; } cx_##SUF;                                                                  \
                                                                                                            \
    /* in-place radix-2 DIT FFT, unnormalised; tw[k] = exp(-+ j 2 pi k / n), k < n/2 */                      \
    static void fft_##SUF(cx_##SUF* a, size_t n, const cx_##SUF* tw) {                                      \
        for (size_t i = 1, j = 0; i < n; ++i) {                                                             \
            size_t bit = n >> 1;                                                                            \
            for (; j & bit; bit >>= 1) j ^= bit;                                                            \
            j ^= bit;                                                                                       \
            if (i < j) { cx_##SUF t = a[i]; a[i] = a[j]; a[j] = t; }                                        \
        }                                                                                                   \
        for (size_t len = 2; len <= n; len <<= 1) {                                                         \
            const size_t half = len >> 1, step = n / len;                                                   \
            for (size_t i = 0; i < n; i += len) {                                                           \
                for (size_t k = 0; k < half; ++k) {                                                         \
                    const cx_##SUF w = tw[k * step];                                                        \
                    const cx_##SUF u = a[i + k], x = a[i + k + half];                                       \
                    const T vr = x.re * w.re - x.im * w.im, vi = x.re * w.im + x.im * w.re;                 \
                    a[i + k].re = u.re + vr; a[i + k].im = u.im + vi;                                       \
                    a[i + k + half].re = u.re - vr; a[i + k + half].im = u.im - vi;                         \
                }                                                                                           \
            }                                                                                               \
        }                                                                                                   \
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
        size_t s_begin, s_end;                                                                              \
        int status;                                                                                         \
    } job_##SUF;                                                                                            \
                                                                                                            \
    static void* worker_##SUF(void* arg) {                                                                  \
        job_##SUF* jb = (job_##SUF*)arg;                                                                    \
        const size_t n = jb->n, N = 2 * n, L = jb->L;                                                       \
        cx_##SUF* buf = (cx_##SUF*)malloc(sizeof(cx_##SUF) * N);                                            \
        cx_##SUF* twf = (cx_##SUF*)malloc(sizeof(cx_##SUF) * (N / 2 + 1));                                  \
        cx_##SUF* twi = (cx_##SUF*)malloc(sizeof(cx_##SUF) * (N / 2 + 1));                                  \
        cx_##SUF* mixed = (cx_##SUF*)malloc(sizeof(cx_##SUF) * n);                                          \
        cx_##SUF* prev = (cx_##SUF*)malloc(sizeof(cx_##SUF) * n);                                           \
        cx_##SUF* ring = (cx_##SUF*)calloc(L ? L : 1, sizeof(cx_##SUF));                                    \
        if (!buf || !twf || !twi || !mixed || !prev || !ring) { jb->status = -1; return NULL; }             \
        for (size_t k = 0; k < N / 2; ++k) {                                                                \
            const double a = -2.0 * M_PI * (double)k / (double)N;                                           \
            twf[k].re = (T)cos(a); twf[k].im = (T)sin(a);                                                   \
            twi[k].re = (T)cos(a); twi[k].im = (T)(-sin(a));                                                \
        }                                                                                                   \
        for (size_t s = jb->s_begin; s < jb->s_end; ++s) {                                                  \
            const cx_##SUF* x = jb->x[s];                                                                   \
            cx_##SUF* y = jb->y[s];                                                                         \
            size_t produced = 0;                                                                            \
            /* FreqShifter: phase table of `denom` entries (transform.rs:321-340) */                        \
            cx_##SUF* phase = NULL;                                                                         \
            size_t plen = 0, pidx = 0;                                                                      \
            if (jb->with_nco) {                                                                             \
                const int64_t numer = jb->numer[s], denom = jb->denom[s];                                   \
                plen = (size_t)denom;                                                                       \
                phase = (cx_##SUF*)malloc(sizeof(cx_##SUF) * plen);                                         \
                if (!phase) { jb->status = -1; break; }                                                     \
                int64_t i = 0;                                                                              \
                for (size_t k = 0; k < plen; ++k) {                                                         \
                    const T ph = (T)0 + (T)i / (T)denom * (T)TAU_CONST; /* transform.rs:335 */              \
                    phase[k].re = COS(ph); phase[k].im = SIN(ph);                                           \
                    i = (i + numer) % denom; /* C '%' truncates like Rust's */                              \
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
                    fft_##SUF(buf, N, twf);                        /* filters.rs:244-246 */                 \
                    for (size_t k = 0; k < N; ++k) {               /* filters.rs:247-249 */                 \
                        const cx_##SUF h = jb->hext[k], v = buf[k];                                         \
                        buf[k].re = v.re * h.re - v.im * h.im;                                              \
                        buf[k].im = v.re * h.im + v.im * h.re;                                              \
                    }                                                                                       \
                    fft_##SUF(buf, N, twi);                        /* filters.rs:250-252 */                 \
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
                            T sr = 0, si = 0;                                                               \
                            size_t q = 0;                                                                   \
                            for (size_t i = ring_pos; i < L; ++i, ++q) { sr += ring[i].re * jb->ir[q]; si += ring[i].im * jb->ir[q]; } \
                            for (size_t i = 0; i < ring_pos; ++i, ++q) { sr += ring[i].re * jb->ir[q]; si += ring[i].im * jb->ir[q]; } \
                            y[produced].re = sr; y[produced].im = si;                                       \
                            ++produced;                                                                     \
                        }                                                                                   \
                    }                                                                                       \
                } else {                                                                                    \
                    memcpy(y + produced, z, sizeof(cx_##SUF) * zn);                                         \
                    produced += zn;                                                                         \
                }                                                                                           \
            }                                                                                               \
            jb->n_out[s] = produced;                                                                        \
            free(phase);                                                                                    \
        }                                                                                                   \
        free(buf); free(twf); free(twi); free(mixed); free(prev); free(ring);                               \
        return NULL;                                                                                        \
    }                                                                                                       \
                                                                                                            \
    /* Runs [FreqShifter ->] [Filter ->] [Downsampler] over n_streams independent streams, each fed   */   \
    /* n_chunks chunks of n samples from a fresh start (Downsampler output framing is not modelled:  */   \
    /* samples are returned concatenated).  Returns 0 on success.                                    */   \
    int oracle_chain_##SUF(size_t n_streams, size_t n, size_t n_chunks, const T* const* x, T* const* y, size_t* n_out, \
                           int with_nco, const int64_t* numer, const int64_t* denom, int with_filter, const T* hext,    \
                           int with_down, const T* ir, size_t L, double in_rate, double out_rate, int n_threads) {      \
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
            jb->numer = numer; jb->denom = denom;                                                           \
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

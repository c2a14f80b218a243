// Host side of the C ABI (include/radiorust_b200.h): contexts, pinned pool,
// design-math exports and the chain object that sequences the sm_100a kernels.
//
// A chain replays, for `n_streams` streams in lock step, what the reference's
// per-block Tokio tasks do for one stream: every
// rr_chain_push is "n_chunks Signal::Samples messages per stream", every
// rr_chain_event is one in-band Signal::Event.  All per-block bookkeeping that
// the reference keeps in task locals (Filter's has-history flag
// filters.rs:164, Downsampler's pos/ring resampling.rs:65-67, FreqShifter's
// phase_idx transform.rs:308, FmDemod's previous sample modulation.rs:103) is
// host scalars here plus device-resident sample state; the sample arithmetic
// itself only ever runs in the CUDA kernels -- there is no CPU fallback.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/radiorust_b200.h"
#include "rr_design.h"
#include "rr_kernels.h"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return RR_ERR_CUDA;
}
#define RR_CUDA(expr)                                     \
    do {                                                  \
        cudaError_t e__ = (expr);                         \
        if (e__ != cudaSuccess) return fail_cuda(e__, #expr); \
    } while (0)
#define RR_TRY(expr)              \
    do {                          \
        int r__ = (expr);         \
        if (r__ != RR_OK) return r__; \
    } while (0)
// kernel launch wrapper: counts launches (rr_kernel_launch_count)
#define RR_LAUNCH(n_kernels, expr)                        \
    do {                                                  \
        cudaError_t e__ = (expr);                         \
        if (e__ != cudaSuccess) return fail_cuda(e__, #expr); \
        g_launches.fetch_add((n_kernels));                \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    // grow-only; contents are NOT preserved
    int ensure(size_t need, bool zero = false) {
        if (need <= bytes && p) {
            return RR_OK;
        }
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        if (need == 0) need = 16;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) {
            p = nullptr;
            g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
            return e == cudaErrorMemoryAllocation ? RR_ERR_NOMEM : RR_ERR_CUDA;
        }
        bytes = need;
        if (zero) {
            e = cudaMemset(p, 0, need);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemset");
        }
        return RR_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct Shape {
    size_t chunk_len = 0, n_chunks = 0;
    double rate = 0.0;
    size_t len() const { return chunk_len * n_chunks; }
};

struct View {
    const void* p = nullptr;
    long long stride = 0;
    Shape sh;
};

// ---- host mirror of one stream's NCO (transform.rs:307-309) -----------------
struct NcoHost {
    int64_t numer = 0, denom = 1;
    uint64_t idx = 0;
    double start_phase = 0.0;  // value representable in Flt
    bool has_table = false;    // phase_vec non-empty
};

// small, copyable host state per stage: everything needed to know the output
// shape of a push without touching the device (dry run == real run)
struct StageHost {
    // FILTER
    double f_sr = NAN;
    size_t f_n = 0;
    bool f_has_hist = false;
    bool f_dirty = true;
    // RESAMPLERS
    double r_in_rate = NAN;
    bool r_valid = false;
    int r_L = 0;
    long long P = 1, Q = 1, j0 = 0, m0 = 0;
    size_t r_pending = 0;
    // FMDEMOD
    bool fm_has_prev = false;
    // FREQSHIFT
    double n_sr = NAN;
};

struct StageAct {
    bool active = false;       // the stage received samples in this push
    bool redesign = false;     // filter response / resampler taps must be rebuilt
    bool first_is_history = false;
    size_t n_new = 0;          // resampler: samples produced by this push
    size_t pending_before = 0;
    long long j0 = 0, m0 = 0;  // resampler counters before the push
    bool nco_recalc = false;
    Shape out;
};

inline long long floordiv128(long long a, long long b, long long c) { return (long long)(((__int128)a * b) / c); }
inline long long ceildiv128(long long a, long long b, long long c) { return (long long)(((__int128)a * b + (c - 1)) / c); }

bool integer_valued(double x) { return std::isfinite(x) && x >= 0.0 && x < 9.0e15 && std::floor(x) == x; }

}  // namespace

struct rr_ctx {
    int device = 0;
    int sm_count = 0;
};

namespace {

struct Stage {
    rr_stage_desc d{};
    StageHost h;
    // FREQSHIFT
    std::vector<double> shift;
    std::vector<uint8_t> shift_dirty;
    bool any_shift_dirty = true;
    std::vector<NcoHost> nco_h;
    std::vector<rr::NcoStream> nco_stage;  // upload staging
    DevBuf nco_d;
    // FILTER
    DevBuf hperm, tw, hist[2];
    int hist_cur = 0;
    // big overlap-save tables
    DevBuf big_h, big_twA, big_twB, big_scratch;
    // RESAMPLERS
    DevBuf ir, tail[2], obuf[2];
    int tail_cur = 0, obuf_cur = 0;
    size_t obuf_cap = 0;  // samples per stream in obuf
    std::vector<double> ir_host;
    // FMDEMOD
    DevBuf fm_prev, fm_last;
    // generic output buffer
    DevBuf out;
    size_t out_cap = 0;
};

}  // namespace

struct rr_chain {
    rr_ctx* ctx = nullptr;
    int dtype = RR_C32;
    size_t esz = 8;  // bytes per complex sample
    int S = 1;
    std::vector<Stage> st;
    cudaStream_t stream = nullptr;
    DevBuf host_in, host_out;  // device staging of rr_chain_push
    std::string plan;
    // optional CUDA-event timing of the dominant kernel of a push (bench.py's roofline)
    bool timing = false;
    std::vector<cudaEvent_t> evs;  // pairs (start, stop), one per timed launch since rr_chain_set_timing
    size_t ev_used = 0;
    std::string timed_kernel;
    int timing_begin() {
        if (!timing) return 0;
        if (ev_used + 2 > evs.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return 0;
            evs.push_back(a);
            evs.push_back(b);
        }
        cudaEventRecord(evs[ev_used], stream);
        return 1;
    }
    void timing_end(const char* name) {
        cudaEventRecord(evs[ev_used + 1], stream);
        ev_used += 2;
        timed_kernel = name;
    }
};

#define RR_TIMED_LAUNCH(c, name, n_kernels, expr)                        \
    do {                                                                 \
        const int t__ = (c)->timing_begin();                             \
        RR_LAUNCH(n_kernels, expr);                                      \
        if (t__) (c)->timing_end(name);                                  \
    } while (0)

namespace {

// ---------------------------------------------------------------------------
// shape logic shared by the dry run (capacity check, rr_chain_max_output) and
// the real run
// ---------------------------------------------------------------------------
int advance_stage(const rr_stage_desc& d, StageHost& h, const Shape& in, StageAct* act) {
    StageAct a;
    a.out = in;
    if (in.n_chunks == 0 || in.chunk_len == 0) {
        a.out.n_chunks = 0;
        *act = a;
        return RR_OK;
    }
    a.active = true;
    switch (d.kind) {
        case RR_STAGE_FREQSHIFT:
            a.nco_recalc = !(h.n_sr == in.rate);
            h.n_sr = in.rate;
            break;
        case RR_STAGE_GAIN:
            break;
        case RR_STAGE_FMDEMOD:
            a.first_is_history = !h.fm_has_prev;
            h.fm_has_prev = true;
            break;
        case RR_STAGE_FILTER: {
            // filters.rs:179-187: redesign on new params / sample rate / chunk length, and drop history
            a.redesign = h.f_dirty || !(h.f_sr == in.rate) || h.f_n != in.chunk_len;
            if (a.redesign) {
                h.f_dirty = false;
                h.f_sr = in.rate;
                h.f_n = in.chunk_len;
                h.f_has_hist = false;
            }
            a.first_is_history = !h.f_has_hist;  // filters.rs:240,260
            a.out.n_chunks = in.n_chunks - (a.first_is_history ? 1 : 0);
            h.f_has_hist = true;
            break;
        }
        case RR_STAGE_DOWNSAMPLE:
        case RR_STAGE_UPSAMPLE: {
            const bool down = d.kind == RR_STAGE_DOWNSAMPLE;
            if (!h.r_valid || !(h.r_in_rate == in.rate)) {  // resampling.rs:75 / :205
                const double in_rate = in.rate, out_rate = d.output_rate;
                if (!(in_rate >= 0.0)) return fail(RR_ERR_INVALID, "input sample rate must be positive");
                if (down && !(in_rate >= out_rate))
                    return fail(RR_ERR_INVALID, "input sample rate must be greater than or equal to output sample rate");
                if (!down && !(in_rate <= out_rate))
                    return fail(RR_ERR_INVALID, "input sample rate must be smaller than or equal to output sample rate");
                if (!down && !(d.bandwidth < in_rate))
                    return fail(RR_ERR_INVALID, "bandwidth must be smaller than input sample rate");
                const double margin = down ? (out_rate - d.bandwidth) / 2.0 : (in_rate - d.bandwidth) / 2.0;
                const double lf = std::ceil((down ? in_rate : out_rate) / margin * d.quality);
                if (!(lf > 0.0) || lf > 1.0e8) return fail(RR_ERR_INVALID, "resampler impulse response length out of range");
                if (!integer_valued(in_rate) || !integer_valued(out_rate) || in_rate == 0.0 || out_rate == 0.0)
                    return fail(RR_ERR_UNSUPPORTED, "device resamplers need integer-valued, non-zero sample rates");
                long long Pn = (long long)in_rate, Qn = (long long)out_rate;
                const long long g = std::gcd(Pn, Qn);
                h.P = Pn / g;
                h.Q = Qn / g;
                h.j0 = 0;
                h.m0 = 0;
                h.r_L = (int)lf;
                h.r_in_rate = in_rate;
                h.r_valid = true;
                a.redesign = true;
            }
            const long long len = (long long)in.len();
            a.j0 = h.j0;
            a.m0 = h.m0;
            long long total_after;
            if (down) total_after = floordiv128(h.j0 + len, h.Q, h.P);  // outputs m with ceil(m*P/Q) <= J
            else total_after = ceildiv128(h.j0 + len, h.Q, h.P);        // q_p = ceil(p*Q/P), resampling.rs:249-266
            a.n_new = (size_t)(total_after - h.m0);
            const long long j1 = (h.j0 + len) % h.P;
            h.j0 = j1;
            h.m0 = down ? floordiv128(j1, h.Q, h.P) : ceildiv128(j1, h.Q, h.P);
            a.pending_before = h.r_pending;
            const size_t ocl = d.output_chunk_len ? (size_t)d.output_chunk_len : 1;
            const size_t total = h.r_pending + a.n_new;
            a.out.chunk_len = ocl;
            a.out.n_chunks = total / ocl;
            a.out.rate = d.output_rate;
            h.r_pending = total - a.out.n_chunks * ocl;
            break;
        }
        default:
            return fail(RR_ERR_INVALID, "unknown stage kind");
    }
    *act = a;
    return RR_OK;
}

template <typename T> struct TypeTag { using type = T; };
#define RR_DISPATCH(chain, T, ...)                     \
    do {                                               \
        if ((chain)->dtype == RR_C32) {                \
            using T = float;                           \
            __VA_ARGS__;                               \
        } else {                                       \
            using T = double;                          \
            __VA_ARGS__;                               \
        }                                              \
    } while (0)

rr::WindowFn make_window(int kind, double beta, rr_window_fn fn, void* user) {
    switch (kind) {
        case RR_WINDOW_RECTANGULAR:
            return [](double) { return 1.0; };
        case RR_WINDOW_CUSTOM:
            return [fn, user](double x) { return fn ? fn(user, x) : 1.0; };
        default:
            return [beta](double x) { return rr::kaiser_rel_with_beta(beta, x); };
    }
}

// upload a vector<complex<double>> as complex<T>
template <typename T> int upload_complex(DevBuf& buf, const std::vector<std::complex<double>>& v, cudaStream_t st) {
    RR_TRY(buf.ensure(v.size() * 2 * sizeof(T)));
    std::vector<T> tmp(v.size() * 2);
    for (size_t i = 0; i < v.size(); ++i) {
        tmp[2 * i] = (T)v[i].real();
        tmp[2 * i + 1] = (T)v[i].imag();
    }
    RR_CUDA(cudaMemcpyAsync(buf.p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaStreamSynchronize(st));  // tmp is pageable and dies here
    return RR_OK;
}
template <typename T> int upload_real(DevBuf& buf, const std::vector<double>& v, cudaStream_t st) {
    RR_TRY(buf.ensure(v.size() * sizeof(T)));
    std::vector<T> tmp(v.size());
    for (size_t i = 0; i < v.size(); ++i) tmp[i] = (T)v[i];
    RR_CUDA(cudaMemcpyAsync(buf.p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaStreamSynchronize(st));
    return RR_OK;
}

// ---- FreqShifter retune (transform.rs:318-340) ------------------------------
template <typename T> int nco_refresh(rr_chain* c, Stage& s, double sample_rate, bool rate_changed) {
    const int S = c->S;
    if (s.nco_h.empty()) s.nco_h.resize(S);
    s.nco_stage.resize(S);
    RR_TRY(s.nco_d.ensure(sizeof(rr::NcoStream) * (size_t)S));
    const bool all = rate_changed;
    if (!all && !s.any_shift_dirty) return RR_OK;
    const T tau = (T)6.283185307179586476925286766559;
    for (int i = 0; i < S; ++i) {
        NcoHost& h = s.nco_h[i];
        if (all || s.shift_dirty[i]) {
            // start_phase = arg(phase_vec[phase_idx]) in Flt, 0 when no table exists yet
            T start = (T)0;
            if (h.has_table) {
                const uint64_t prod = ((uint64_t)(h.numer < 0 ? -h.numer : h.numer) * (h.idx % (uint64_t)h.denom)) % (uint64_t)h.denom;
                const T iv = h.numer < 0 ? -(T)prod : (T)prod;
                const T ph = (T)h.start_phase + iv / (T)h.denom * tau;
                const T im = std::sin(ph), re = std::cos(ph);
                start = std::atan2(im, re);
            }
            int64_t numer = 0, denom = 1;
            if (!rr::freq_to_ratio(sample_rate, s.d.precision, s.shift[i], &numer, &denom))
                return fail(RR_ERR_INVALID, "FreqShifter: sample_rate / precision rounds to zero (Ratio::new panics)");
            if (denom <= 0 || denom >= (int64_t)1 << 31)
                return fail(RR_ERR_UNSUPPORTED, "FreqShifter: phase table length (sample_rate / precision) must be below 2^31");
            h.numer = numer;
            h.denom = denom;
            h.idx = 0;
            h.start_phase = (double)start;
            h.has_table = true;
            s.shift_dirty[i] = 0;
        }
        rr::NcoStream& d = s.nco_stage[i];
        const int64_t na = h.numer < 0 ? -h.numer : h.numer;
        d.numer_abs = (uint32_t)(na % h.denom);
        d.denom = (uint32_t)h.denom;
        d.idx = (uint32_t)(h.idx % (uint64_t)h.denom);
        d.sign = h.numer < 0 ? -1 : (h.numer > 0 ? 1 : 0);
        d.start_phase = h.start_phase;
    }
    s.any_shift_dirty = false;
    RR_CUDA(cudaMemcpyAsync(s.nco_d.p, s.nco_stage.data(), sizeof(rr::NcoStream) * (size_t)S, cudaMemcpyHostToDevice, c->stream));
    RR_CUDA(cudaStreamSynchronize(c->stream));
    return RR_OK;
}
void nco_host_advance(Stage& s, long long len) {
    for (auto& h : s.nco_h) h.idx = (h.idx + (uint64_t)len) % (uint64_t)h.denom;
}

// ---- Filter design + upload (filters.rs:184-238) ----------------------------
template <typename T> int filter_redesign(rr_chain* c, Stage& s, double sample_rate, size_t n) {
    if (n < 2 || (n & (n - 1)) != 0)
        return fail(RR_ERR_UNSUPPORTED, "Filter: the device path needs a power-of-two chunk length >= 2");
    const rr_stage_desc& d = s.d;
    if (!d.freq_resp) return fail(RR_ERR_INVALID, "Filter: freq_resp callback is null");
    rr_freq_resp_fn fn = d.freq_resp;
    void* user = d.freq_resp_user;
    rr::FreqResp f = [fn, user](int64_t bin, double freq) {
        double re = 0.0, im = 0.0;
        fn(user, bin, freq, &re, &im);
        return std::complex<double>(re, im);
    };
    rr::WindowFn w = make_window(d.window_kind, d.window_beta, d.window_fn, d.window_user);
    std::vector<std::complex<double>> H;
    if (!rr::design_filter_response(f, w, sample_rate, n, sizeof(T) == 4, &H))
        return fail(RR_ERR_UNSUPPORTED, "Filter: design failed");
    const size_t N = 2 * n;
    if (rr::chain_os_supported<T>((int)n, 0, 0)) {
        std::vector<std::complex<double>> hp(N), tw;
        for (size_t k = 0; k < N; ++k) hp[(size_t)rr::chain_os_hperm_index<T>((int)n, (int)k)] = H[k];
        RR_TRY(upload_complex<T>(s.hperm, hp, c->stream));
        rr::make_twiddles(N, &tw);
        RR_TRY(upload_complex<T>(s.tw, tw, c->stream));
    } else if (rr::big_os_supported<T>((int)n)) {
        int Na = 0, Nb = 0;
        rr::big_os_shape<T>((int)n, &Na, &Nb);
        std::vector<std::complex<double>> hp(N), tw;
        for (size_t k = 0; k < N; ++k) hp[(size_t)rr::big_os_hperm_index<T>((int)n, (long long)k)] = H[k];
        RR_TRY(upload_complex<T>(s.big_h, hp, c->stream));
        rr::make_twiddles(N, &tw);
        RR_TRY(upload_complex<T>(s.tw, tw, c->stream));
        rr::make_twiddles((size_t)Na, &tw);
        RR_TRY(upload_complex<T>(s.big_twA, tw, c->stream));
        rr::make_twiddles((size_t)Nb, &tw);
        RR_TRY(upload_complex<T>(s.big_twB, tw, c->stream));
    } else {
        return fail(RR_ERR_UNSUPPORTED, "Filter: chunk length not supported by the device path");
    }
    const size_t hb = (size_t)c->S * n * 2 * sizeof(T);
    RR_TRY(s.hist[0].ensure(hb));
    RR_TRY(s.hist[1].ensure(hb));
    return RR_OK;
}

// ---- resampler taps (resampling.rs:82-101, :216-236) -------------------------
template <typename T> int resampler_redesign(rr_chain* c, Stage& s, double in_rate) {
    const rr_stage_desc& d = s.d;
    const bool down = d.kind == RR_STAGE_DOWNSAMPLE;
    const int L = s.h.r_L;
    const double margin = down ? (d.output_rate - d.bandwidth) / 2.0 : (in_rate - d.bandwidth) / 2.0;
    const double ratio = down ? d.output_rate / in_rate : in_rate / d.output_rate;
    const double null_bin = (double)L * margin / (down ? in_rate : d.output_rate);
    rr::design_resampler_taps((size_t)L, ratio, null_bin, &s.ir_host);
    RR_TRY(upload_real<T>(s.ir, s.ir_host, c->stream));
    // ring buffer / accumulators restart at zero (resampling.rs:99-101, :234-236)
    const size_t state_len = down ? (size_t)(L > 1 ? L - 1 : 1) : (size_t)L;
    const size_t tb = (size_t)c->S * state_len * 2 * sizeof(T);
    for (int k = 0; k < 2; ++k) {
        RR_TRY(s.tail[k].ensure(tb));
        RR_CUDA(cudaMemsetAsync(s.tail[k].p, 0, tb, c->stream));
    }
    s.tail_cur = 0;
    return RR_OK;
}

// make room for `need` samples per stream in both output staging buffers of a
// resampler, preserving the `keep` pending samples of the current one
template <typename T> int obuf_reserve(rr_chain* c, Stage& s, size_t need, size_t keep) {
    if (need <= s.obuf_cap && s.obuf[0].p && s.obuf[1].p) return RR_OK;
    size_t cap = need + need / 2 + 16;
    const size_t bytes = (size_t)c->S * cap * 2 * sizeof(T);
    DevBuf nb0, nb1;
    RR_TRY(nb0.ensure(bytes));
    RR_TRY(nb1.ensure(bytes));
    if (keep && s.obuf[s.obuf_cur].p) {
        RR_LAUNCH(1, rr::launch_copy2d<T>(s.obuf[s.obuf_cur].p, (long long)s.obuf_cap, nb0.p, (long long)cap, (long long)keep, c->S,
                                          c->stream));
        RR_CUDA(cudaStreamSynchronize(c->stream));
    }
    s.obuf[0].release();
    s.obuf[1].release();
    s.obuf[0] = nb0;
    s.obuf[1] = nb1;
    s.obuf_cur = 0;
    s.obuf_cap = cap;
    return RR_OK;
}

template <typename T> int out_reserve(rr_chain* c, Stage& s, size_t need) {
    if (need <= s.out_cap && s.out.p) return RR_OK;
    RR_TRY(s.out.ensure((size_t)c->S * need * 2 * sizeof(T)));
    s.out_cap = need;
    return RR_OK;
}

struct Dest {
    void* p = nullptr;
    long long stride = 0;
};

// chooses where a stage writes: the caller's buffer when it is the last stage
template <typename T> int pick_dest(rr_chain* c, Stage& s, bool last, void* user_out, long long user_stride, size_t len, Dest* d) {
    if (last) {
        d->p = user_out;
        d->stride = user_stride;
        return RR_OK;
    }
    RR_TRY(out_reserve<T>(c, s, len));
    d->p = s.out.p;
    d->stride = (long long)s.out_cap;
    return RR_OK;
}

// after a resampler wrote n_new samples behind `pending_before` ones in its
// staging buffer: hand whole chunks downstream, keep the remainder
template <typename T>
int resampler_emit(rr_chain* c, Stage& s, const StageAct& a, bool last, void* user_out, long long user_stride, View* next) {
    const size_t emit = a.out.len();
    const size_t total = a.pending_before + a.n_new;
    DevBuf& cur = s.obuf[s.obuf_cur];
    DevBuf& other = s.obuf[s.obuf_cur ^ 1];
    const size_t rem = total - emit;
    if (rem > 0 && emit > 0) {
        const char* src = (const char*)cur.p + emit * 2 * sizeof(T);
        RR_LAUNCH(1, rr::launch_copy2d<T>(src, (long long)s.obuf_cap, other.p, (long long)s.obuf_cap, (long long)rem, c->S, c->stream));
    }
    next->sh = a.out;
    if (last) {
        if (emit > 0)
            RR_LAUNCH(1, rr::launch_copy2d<T>(cur.p, (long long)s.obuf_cap, user_out, user_stride, (long long)emit, c->S, c->stream));
        next->p = user_out;
        next->stride = user_stride;
    } else {
        next->p = cur.p;
        next->stride = (long long)s.obuf_cap;
    }
    if (emit > 0) s.obuf_cur ^= 1;  // the remainder (possibly empty) now lives in `other`
    return RR_OK;
}

template <typename T>
int run_push(rr_chain* c, double sample_rate, size_t chunk_len, size_t n_chunks, const void* dev_in, size_t in_stride, void* dev_out,
             size_t out_capacity, size_t out_stride, size_t* out_count, double* out_rate) {
    const int S = c->S;
    const int ns = (int)c->st.size();
    cudaStream_t st = c->stream;
    Shape in;
    in.chunk_len = chunk_len;
    in.n_chunks = n_chunks;
    in.rate = sample_rate;
    if (chunk_len * n_chunks > in_stride && S > 1) return fail(RR_ERR_INVALID, "in_stride smaller than the pushed samples");

    // ---- dry run: output shape + capacity before anything is touched -------
    {
        Shape sh = in;
        for (int i = 0; i < ns; ++i) {
            StageHost h = c->st[i].h;
            StageAct a;
            RR_TRY(advance_stage(c->st[i].d, h, sh, &a));
            sh = a.out;
        }
        if (sh.len() > out_capacity) return fail(RR_ERR_CAPACITY, "output buffer too small for this push");
        if (sh.len() > out_stride && S > 1) return fail(RR_ERR_INVALID, "out_stride smaller than the produced samples");
    }

    std::string plan;
    View cur;
    cur.p = dev_in;
    cur.stride = (long long)in_stride;
    cur.sh = in;
    Stage* pending_nco = nullptr;  // FreqShifter deferred into the next Filter's load stage

    for (int i = 0; i < ns; ++i) {
        Stage& s = c->st[i];
        const bool last = (i == ns - 1);
        StageAct a;
        RR_TRY(advance_stage(s.d, s.h, cur.sh, &a));
        if (!a.active) {
            cur.sh = a.out;
            continue;
        }
        const long long len = (long long)cur.sh.len();
        if (!plan.empty() && plan.back() != '+') plan += "|";
        switch (s.d.kind) {
            case RR_STAGE_FREQSHIFT: {
                RR_TRY(nco_refresh<T>(c, s, cur.sh.rate, a.nco_recalc));
                const bool fuse = !last && c->st[i + 1].d.kind == RR_STAGE_FILTER && cur.sh.chunk_len >= 32 &&
                                  (cur.sh.chunk_len & (cur.sh.chunk_len - 1)) == 0 &&
                                  rr::chain_os_supported<T>((int)cur.sh.chunk_len, 0, 0);
                if (fuse) {
                    pending_nco = &s;
                    plan += "nco+";
                } else {
                    Dest d;
                    RR_TRY(pick_dest<T>(c, s, last, dev_out, (long long)out_stride, (size_t)len, &d));
                    RR_LAUNCH(2, rr::launch_freqshift<T>(cur.p, cur.stride, d.p, d.stride, len, S, (rr::NcoStream*)s.nco_d.p, st));
                    nco_host_advance(s, len);
                    cur.p = d.p;
                    cur.stride = d.stride;
                    plan += "freqshift";
                }
                break;
            }
            case RR_STAGE_GAIN: {
                Dest d;
                RR_TRY(pick_dest<T>(c, s, last, dev_out, (long long)out_stride, (size_t)len, &d));
                RR_LAUNCH(1, rr::launch_gain<T>(cur.p, cur.stride, d.p, d.stride, len, S, s.d.gain, st));
                cur.p = d.p;
                cur.stride = d.stride;
                plan += "gain";
                break;
            }
            case RR_STAGE_FMDEMOD: {
                Dest d;
                RR_TRY(pick_dest<T>(c, s, last, dev_out, (long long)out_stride, (size_t)len, &d));
                RR_TRY(s.fm_prev.ensure((size_t)S * 2 * sizeof(T), true));
                RR_TRY(s.fm_last.ensure((size_t)S * 2 * sizeof(T), true));
                const double factor = cur.sh.rate / s.d.deviation / 6.283185307179586476925286766559;  // modulation.rs:116
                RR_LAUNCH(2, rr::launch_fmdemod<T>(cur.p, cur.stride, d.p, d.stride, len, S, s.fm_prev.p, s.fm_last.p,
                                                   a.first_is_history ? 0 : 1, factor, st));
                cur.p = d.p;
                cur.stride = d.stride;
                plan += "fmdemod";
                break;
            }
            case RR_STAGE_FILTER: {
                const size_t n = cur.sh.chunk_len;
                if (a.redesign) RR_TRY(filter_redesign<T>(c, s, cur.sh.rate, n));
                const bool small = rr::chain_os_supported<T>((int)n, 0, 0);
                Stage* nco = pending_nco;
                pending_nco = nullptr;
                if (small) {
                    // try to fold the following Downsampler into the epilogue
                    bool fuse_down = false;
                    StageAct da;
                    Stage* ds = nullptr;
                    if (!last && c->st[i + 1].d.kind == RR_STAGE_DOWNSAMPLE && a.out.n_chunks > 0) {
                        ds = &c->st[i + 1];
                        StageHost probe = ds->h;
                        StageAct pa;
                        int r = advance_stage(ds->d, probe, a.out, &pa);
                        if (r == RR_OK && rr::chain_os_supported<T>((int)n, 1, probe.r_L)) fuse_down = true;
                    }
                    rr::ChainOsArgs<T> k{};
                    k.in = cur.p;
                    k.in_stride = cur.stride;
                    k.n_chunks = (int)cur.sh.n_chunks;
                    k.first_is_history = a.first_is_history ? 1 : 0;
                    k.emit = 1;
                    k.hist_in = s.hist[s.hist_cur].p;
                    k.hist_out = s.hist[s.hist_cur ^ 1].p;
                    k.hperm = s.hperm.p;
                    k.twN = s.tw.p;
                    k.nco = nco ? (const rr::NcoStream*)nco->nco_d.p : nullptr;
                    k.nco_offset = 0;
                    if (fuse_down) {
                        RR_TRY(advance_stage(ds->d, ds->h, a.out, &da));
                        if (da.redesign) RR_TRY(resampler_redesign<T>(c, *ds, a.out.rate));
                        RR_TRY(obuf_reserve<T>(c, *ds, da.pending_before + da.n_new, da.pending_before));
                        k.out = (char*)ds->obuf[ds->obuf_cur].p + da.pending_before * 2 * sizeof(T);
                        k.out_stride = (long long)ds->obuf_cap;
                        k.ir = (const T*)ds->ir.p;
                        k.L = ds->h.r_L;
                        k.ztail_in = ds->tail[ds->tail_cur].p;
                        k.ztail_out = ds->tail[ds->tail_cur ^ 1].p;
                        k.rate.P = ds->h.P;
                        k.rate.Q = ds->h.Q;
                        k.rate.j0 = da.j0;
                        k.rate.m0 = da.m0;
                        RR_TIMED_LAUNCH(c, "k_chain_os<epi=down>", 1, rr::launch_chain_os<T>((int)n, 1, S, 1, k, st));
                        ds->tail_cur ^= 1;
                        s.hist_cur ^= 1;
                        plan += "fused_os[filter+down]";
                        const bool ds_last = (i + 1 == ns - 1);
                        RR_TRY(resampler_emit<T>(c, *ds, da, ds_last, dev_out, (long long)out_stride, &cur));
                        ++i;  // the Downsampler stage is done
                    } else {
                        Dest d;
                        RR_TRY(pick_dest<T>(c, s, last, dev_out, (long long)out_stride, a.out.len(), &d));
                        k.out = d.p;
                        k.out_stride = d.stride;
                        int parts = 1;
                        const int total = (int)a.out.n_chunks;
                        if (total > 1) {
                            const int want = (2 * c->ctx->sm_count + S - 1) / S;
                            parts = want < total ? want : total;
                            if (parts < 1) parts = 1;
                        }
                        RR_TIMED_LAUNCH(c, "k_chain_os<epi=none>", 1, rr::launch_chain_os<T>((int)n, 0, S, parts, k, st));
                        s.hist_cur ^= 1;
                        cur.p = d.p;
                        cur.stride = d.stride;
                        cur.sh = a.out;
                        plan += "fused_os[filter]";
                    }
                    if (nco) {
                        RR_LAUNCH(1, rr::launch_nco_advance((rr::NcoStream*)nco->nco_d.p, S, len, st));
                        nco_host_advance(*nco, len);
                    }
                } else {
                    // large chunk: four-step FFT through L2-resident scratch
                    if (nco) return fail(RR_ERR_UNSUPPORTED, "internal: NCO deferred into an unfused filter");
                    if (!rr::big_os_supported<T>((int)n)) return fail(RR_ERR_UNSUPPORTED, "Filter: chunk length not supported");
                    const size_t N = 2 * n;
                    const int n_blocks = (int)a.out.n_chunks;
                    Dest d;
                    RR_TRY(pick_dest<T>(c, s, last, dev_out, (long long)out_stride, a.out.len(), &d));
                    if (n_blocks > 0) {
                        // bound the scratch so it stays L2 resident: process blocks in groups
                        size_t max_blocks = ((size_t)48 << 20) / (N * 2 * sizeof(T));
                        if (max_blocks < 1) max_blocks = 1;
                        size_t per_launch = max_blocks / (size_t)S;
                        if (per_launch < 1) per_launch = 1;
                        if (per_launch > (size_t)n_blocks) per_launch = (size_t)n_blocks;
                        RR_TRY(s.big_scratch.ensure((size_t)S * per_launch * N * 2 * sizeof(T)));
                        const int c_first = a.first_is_history ? 1 : 0;
                        for (size_t b0 = 0; b0 < (size_t)n_blocks; b0 += per_launch) {
                            const size_t nb = std::min(per_launch, (size_t)n_blocks - b0);
                            rr::BigOsArgs<T> k{};
                            k.in = cur.p;
                            k.in_stride = cur.stride;
                            k.hist = s.hist[s.hist_cur].p;
                            k.first_chunk = c_first + (int)b0;
                            k.n_blocks = (int)nb;
                            k.scratch = s.big_scratch.p;
                            k.hbig = s.big_h.p;
                            k.twN = s.tw.p;
                            k.twA = s.big_twA.p;
                            k.twB = s.big_twB.p;
                            k.out = (char*)d.p + b0 * n * 2 * sizeof(T);
                            k.out_stride = d.stride;
                            RR_LAUNCH(3, rr::launch_big_os<T>((int)n, S, k, st));
                        }
                    }
                    // new history = last pushed chunk (filters.rs:260)
                    RR_LAUNCH(1, rr::launch_copy2d<T>((const char*)cur.p + (size_t)(len - (long long)n) * 2 * sizeof(T), cur.stride,
                                                      s.hist[s.hist_cur ^ 1].p, (long long)n, (long long)n, S, st));
                    s.hist_cur ^= 1;
                    cur.p = d.p;
                    cur.stride = d.stride;
                    cur.sh = a.out;
                    plan += "big_os";
                }
                break;
            }
            case RR_STAGE_DOWNSAMPLE:
            case RR_STAGE_UPSAMPLE: {
                const bool down = s.d.kind == RR_STAGE_DOWNSAMPLE;
                if (a.redesign) RR_TRY(resampler_redesign<T>(c, s, cur.sh.rate));
                RR_TRY(obuf_reserve<T>(c, s, a.pending_before + a.n_new, a.pending_before));
                void* o = (char*)s.obuf[s.obuf_cur].p + a.pending_before * 2 * sizeof(T);
                rr::RateState rs;
                rs.P = s.h.P;
                rs.Q = s.h.Q;
                rs.j0 = a.j0;
                rs.m0 = a.m0;
                if (down) {
                    RR_LAUNCH(2, rr::launch_downsample<T>(cur.p, cur.stride, len, s.tail[s.tail_cur].p, s.tail[s.tail_cur ^ 1].p,
                                                          (const T*)s.ir.p, s.h.r_L, rs, (long long)a.n_new, o, (long long)s.obuf_cap, S, st));
                    plan += "downsample";
                } else {
                    RR_LAUNCH(1, rr::launch_upsample<T>(cur.p, cur.stride, len, s.tail[s.tail_cur].p, s.tail[s.tail_cur ^ 1].p,
                                                        (const T*)s.ir.p, s.h.r_L, rs, (long long)a.n_new, o, (long long)s.obuf_cap, S, st));
                    plan += "upsample";
                }
                s.tail_cur ^= 1;
                RR_TRY(resampler_emit<T>(c, s, a, last, dev_out, (long long)out_stride, &cur));
                break;
            }
            default:
                return fail(RR_ERR_INVALID, "unknown stage kind");
        }
        if (s.d.kind != RR_STAGE_FILTER && s.d.kind != RR_STAGE_DOWNSAMPLE && s.d.kind != RR_STAGE_UPSAMPLE) cur.sh = a.out;
    }
    // a chain whose last active stage did not write into the caller's buffer
    // (inactive tail stages, or an empty chain): copy the samples through
    if (cur.sh.len() > 0 && cur.p != dev_out) {
        RR_LAUNCH(1, rr::launch_copy2d<T>(cur.p, cur.stride, dev_out, (long long)out_stride, (long long)cur.sh.len(), S, st));
    }
    if (out_count) *out_count = cur.sh.len();
    if (out_rate) *out_rate = cur.sh.rate;
    c->plan = plan;
    return RR_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* rr_last_error(void) { return g_err.c_str(); }

int rr_version(int* major, int* minor) {
    if (major) *major = RR_VERSION_MAJOR;
    if (minor) *minor = RR_VERSION_MINOR;
    return RR_OK;
}

uint64_t rr_kernel_launch_count(void) { return g_launches.load(); }

int rr_ctx_create(int device, rr_ctx** out) {
    if (!out) return fail(RR_ERR_INVALID, "rr_ctx_create: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(RR_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(RR_ERR_INVALID, "rr_ctx_create: device index out of range");
    RR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(RR_ERR_UNSUPPORTED, "this library holds sm_100a code only; device compute capability is below 10.0");
    rr_ctx* c = new rr_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return RR_OK;
}
int rr_ctx_destroy(rr_ctx* ctx) {
    delete ctx;
    return RR_OK;
}
int rr_ctx_device(const rr_ctx* ctx) { return ctx ? ctx->device : -1; }

int rr_pinned_alloc(rr_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return fail(RR_ERR_INVALID, "rr_pinned_alloc: null argument");
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return RR_OK;
}
int rr_pinned_free(rr_ctx* ctx, void* p) {
    if (!ctx) return fail(RR_ERR_INVALID, "rr_pinned_free: null context");
    if (p) RR_CUDA(cudaFreeHost(p));
    return RR_OK;
}
int rr_host_register(rr_ctx* ctx, void* p, size_t bytes) {
    if (!ctx || !p) return fail(RR_ERR_INVALID, "rr_host_register: null argument");
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return RR_OK;
}
int rr_host_unregister(rr_ctx* ctx, void* p) {
    if (!ctx || !p) return fail(RR_ERR_INVALID, "rr_host_unregister: null argument");
    RR_CUDA(cudaHostUnregister(p));
    return RR_OK;
}
int rr_device_alloc(rr_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return fail(RR_ERR_INVALID, "rr_device_alloc: null argument");
    RR_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return fail(RR_ERR_NOMEM, "rr_device_alloc: out of device memory");
    }
    RR_CUDA(e);
    return RR_OK;
}
int rr_device_free(rr_ctx* ctx, void* p) {
    if (!ctx) return fail(RR_ERR_INVALID, "rr_device_free: null context");
    RR_CUDA(cudaSetDevice(ctx->device));
    if (p) RR_CUDA(cudaFree(p));
    return RR_OK;
}
int rr_memcpy_h2d(rr_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (!ctx) return fail(RR_ERR_INVALID, "rr_memcpy_h2d: null context");
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return RR_OK;
}
int rr_memcpy_d2h(rr_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (!ctx) return fail(RR_ERR_INVALID, "rr_memcpy_d2h: null context");
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return RR_OK;
}

// ---- design math ------------------------------------------------------------
double rr_bessel_i0(double x) { return rr::bessel_i0(x); }
double rr_sinc(double x) { return rr::sinc(x); }
double rr_kaiser_rel_with_beta(double beta, double x) { return rr::kaiser_rel_with_beta(beta, x); }
double rr_kaiser_null_at_bin_to_beta(double n) { return rr::kaiser_null_at_bin_to_beta(n); }
void rr_deemphasis_factor(double tau, double frequency, double* out_re, double* out_im) {
    const std::complex<double> v = rr::deemphasis_factor(tau, frequency);
    if (out_re) *out_re = v.real();
    if (out_im) *out_im = v.imag();
}
int rr_freq_to_ratio(double sample_rate, double precision, double frequency, int64_t* numer, int64_t* denom) {
    int64_t n = 0, d = 1;
    if (!rr::freq_to_ratio(sample_rate, precision, frequency, &n, &d))
        return fail(RR_ERR_INVALID, "denominator == 0 (Ratio::new panics in the reference)");
    if (numer) *numer = n;
    if (denom) *denom = d;
    return RR_OK;
}
int rr_design_filter_response(rr_freq_resp_fn f, void* f_user, int32_t window_kind, double window_beta, rr_window_fn w, void* w_user,
                              double sample_rate, size_t n, int32_t dtype, double* out_2n_complex) {
    if (!f || !out_2n_complex) return fail(RR_ERR_INVALID, "rr_design_filter_response: null argument");
    rr::FreqResp fr = [f, f_user](int64_t bin, double freq) {
        double re = 0.0, im = 0.0;
        f(f_user, bin, freq, &re, &im);
        return std::complex<double>(re, im);
    };
    std::vector<std::complex<double>> H;
    if (!rr::design_filter_response(fr, make_window(window_kind, window_beta, w, w_user), sample_rate, n, dtype == RR_C32, &H))
        return fail(RR_ERR_UNSUPPORTED, "rr_design_filter_response: n must be a power of two >= 2");
    for (size_t i = 0; i < H.size(); ++i) {
        out_2n_complex[2 * i] = H[i].real();
        out_2n_complex[2 * i + 1] = H[i].imag();
    }
    return RR_OK;
}
static int design_taps(bool down, double input_rate, double output_rate, double bandwidth, double quality, size_t* ir_len, double* ir) {
    if (!ir_len) return fail(RR_ERR_INVALID, "ir_len is null");
    if (!(output_rate >= 0.0) || !(bandwidth >= 0.0) || !(input_rate >= 0.0)) return fail(RR_ERR_INVALID, "rates must be positive");
    if (down && !(bandwidth < output_rate)) return fail(RR_ERR_INVALID, "bandwidth must be smaller than output sample rate");
    if (down && !(input_rate >= output_rate)) return fail(RR_ERR_INVALID, "input sample rate must be >= output sample rate");
    if (!down && !(input_rate <= output_rate)) return fail(RR_ERR_INVALID, "input sample rate must be <= output sample rate");
    if (!down && !(bandwidth < input_rate)) return fail(RR_ERR_INVALID, "bandwidth must be smaller than input sample rate");
    const double margin = down ? (output_rate - bandwidth) / 2.0 : (input_rate - bandwidth) / 2.0;
    const double lf = std::ceil((down ? input_rate : output_rate) / margin * quality);
    if (!(lf > 0.0) || lf > 1.0e8) return fail(RR_ERR_INVALID, "impulse response length out of range");
    const size_t L = (size_t)lf;
    const size_t cap = *ir_len;
    *ir_len = L;
    if (!ir) return RR_OK;
    if (cap < L) return fail(RR_ERR_CAPACITY, "ir buffer too small");
    std::vector<double> taps;
    const double ratio = down ? output_rate / input_rate : input_rate / output_rate;
    const double null_bin = (double)L * margin / (down ? input_rate : output_rate);
    rr::design_resampler_taps(L, ratio, null_bin, &taps);
    std::memcpy(ir, taps.data(), L * sizeof(double));
    return RR_OK;
}
int rr_design_downsampler_taps(double input_rate, double output_rate, double bandwidth, double quality, size_t* ir_len, double* ir) {
    return design_taps(true, input_rate, output_rate, bandwidth, quality, ir_len, ir);
}
int rr_design_upsampler_taps(double input_rate, double output_rate, double bandwidth, double quality, size_t* ir_len, double* ir) {
    return design_taps(false, input_rate, output_rate, bandwidth, quality, ir_len, ir);
}

// ---- chain ---------------------------------------------------------------------
int rr_chain_create(rr_ctx* ctx, const rr_chain_desc* desc, rr_chain** out) {
    if (!ctx || !desc || !out) return fail(RR_ERR_INVALID, "rr_chain_create: null argument");
    *out = nullptr;
    if (desc->dtype != RR_C32 && desc->dtype != RR_C64) return fail(RR_ERR_INVALID, "rr_chain_create: bad dtype");
    if (desc->n_streams < 1) return fail(RR_ERR_INVALID, "rr_chain_create: n_streams must be >= 1");
    if (desc->n_stages < 0 || (desc->n_stages > 0 && !desc->stages)) return fail(RR_ERR_INVALID, "rr_chain_create: bad stage list");
    for (int i = 0; i < desc->n_stages; ++i) {
        const rr_stage_desc& d = desc->stages[i];
        switch (d.kind) {
            case RR_STAGE_FREQSHIFT:
                if (!(d.precision > 0.0)) return fail(RR_ERR_INVALID, "FreqShifter: precision must be positive");
                break;
            case RR_STAGE_FILTER:
                if (!d.freq_resp) return fail(RR_ERR_INVALID, "Filter: freq_resp callback is null");
                break;
            case RR_STAGE_DOWNSAMPLE:  // resampling.rs:51-56
                if (!(d.output_rate >= 0.0)) return fail(RR_ERR_INVALID, "output sample rate must be positive");
                if (!(d.bandwidth >= 0.0)) return fail(RR_ERR_INVALID, "bandwidth must be positive");
                if (!(d.bandwidth < d.output_rate)) return fail(RR_ERR_INVALID, "bandwidth must be smaller than output sample rate");
                if (!(d.quality >= 1.0)) return fail(RR_ERR_INVALID, "quality must be >= 1.0");
                break;
            case RR_STAGE_UPSAMPLE:  // resampling.rs:186-188
                if (!(d.output_rate >= 0.0)) return fail(RR_ERR_INVALID, "output sample rate must be positive");
                if (!(d.bandwidth >= 0.0)) return fail(RR_ERR_INVALID, "bandwidth must be positive");
                if (!(d.quality >= 1.0)) return fail(RR_ERR_INVALID, "quality must be >= 1.0");
                break;
            case RR_STAGE_FMDEMOD:
            case RR_STAGE_GAIN:
                break;
            default:
                return fail(RR_ERR_INVALID, "rr_chain_create: unknown stage kind");
        }
    }
    RR_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<rr_chain> c(new rr_chain());
    c->ctx = ctx;
    c->dtype = desc->dtype;
    c->esz = desc->dtype == RR_C32 ? 8 : 16;
    c->S = desc->n_streams;
    c->st.resize((size_t)desc->n_stages);
    for (int i = 0; i < desc->n_stages; ++i) {
        Stage& s = c->st[(size_t)i];
        s.d = desc->stages[i];
        if (s.d.kind == RR_STAGE_FREQSHIFT) {
            s.shift.assign((size_t)c->S, s.d.shift);
            s.shift_dirty.assign((size_t)c->S, 1);
            s.any_shift_dirty = true;
        }
    }
    RR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c.release();
    return RR_OK;
}

int rr_chain_destroy(rr_chain* c) {
    if (!c) return RR_OK;
    cudaSetDevice(c->ctx->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& s : c->st) {
        DevBuf* bufs[] = {&s.nco_d, &s.hperm, &s.tw, &s.hist[0], &s.hist[1], &s.big_h, &s.big_twA, &s.big_twB, &s.big_scratch,
                          &s.ir, &s.tail[0], &s.tail[1], &s.obuf[0], &s.obuf[1], &s.fm_prev, &s.fm_last, &s.out};
        for (DevBuf* b : bufs) b->release();
    }
    c->host_in.release();
    c->host_out.release();
    for (cudaEvent_t e : c->evs) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RR_OK;
}

static int stage_of(rr_chain* c, int stage, int kind, Stage** out) {
    if (!c) return fail(RR_ERR_INVALID, "null chain");
    if (stage < 0 || stage >= (int)c->st.size()) return fail(RR_ERR_INVALID, "stage index out of range");
    if (c->st[(size_t)stage].d.kind != kind) return fail(RR_ERR_INVALID, "stage is of a different kind");
    *out = &c->st[(size_t)stage];
    return RR_OK;
}

int rr_chain_set_shift(rr_chain* c, int stage, int stream, double shift_hz) {
    Stage* s = nullptr;
    RR_TRY(stage_of(c, stage, RR_STAGE_FREQSHIFT, &s));
    if (stream >= c->S) return fail(RR_ERR_INVALID, "stream index out of range");
    if (stream < 0) {
        for (int i = 0; i < c->S; ++i) {
            s->shift[(size_t)i] = shift_hz;
            s->shift_dirty[(size_t)i] = 1;
        }
    } else {
        s->shift[(size_t)stream] = shift_hz;
        s->shift_dirty[(size_t)stream] = 1;
    }
    s->any_shift_dirty = true;
    return RR_OK;
}
int rr_chain_set_shifts(rr_chain* c, int stage, const double* shifts_hz, int n) {
    Stage* s = nullptr;
    RR_TRY(stage_of(c, stage, RR_STAGE_FREQSHIFT, &s));
    if (!shifts_hz || n != c->S) return fail(RR_ERR_INVALID, "rr_chain_set_shifts: need one shift per stream");
    for (int i = 0; i < n; ++i) {
        s->shift[(size_t)i] = shifts_hz[i];
        s->shift_dirty[(size_t)i] = 1;
    }
    s->any_shift_dirty = true;
    return RR_OK;
}
int rr_chain_get_shift(rr_chain* c, int stage, int stream, double* shift_hz) {
    Stage* s = nullptr;
    RR_TRY(stage_of(c, stage, RR_STAGE_FREQSHIFT, &s));
    if (stream < 0 || stream >= c->S || !shift_hz) return fail(RR_ERR_INVALID, "bad stream index");
    *shift_hz = s->shift[(size_t)stream];
    return RR_OK;
}

int rr_chain_update_filter(rr_chain* c, int stage, rr_freq_resp_fn f, void* f_user, int32_t window_kind, double window_beta,
                           rr_window_fn w, void* w_user, int keep_window) {
    Stage* s = nullptr;
    RR_TRY(stage_of(c, stage, RR_STAGE_FILTER, &s));
    if (!f) return fail(RR_ERR_INVALID, "rr_chain_update_filter: freq_resp is null");
    s->d.freq_resp = f;
    s->d.freq_resp_user = f_user;
    if (!keep_window) {
        s->d.window_kind = window_kind;
        s->d.window_beta = window_beta;
        s->d.window_fn = w;
        s->d.window_user = w_user;
    }
    s->h.f_dirty = true;
    return RR_OK;
}
int rr_chain_set_deviation(rr_chain* c, int stage, double deviation) {
    Stage* s = nullptr;
    RR_TRY(stage_of(c, stage, RR_STAGE_FMDEMOD, &s));
    s->d.deviation = deviation;
    return RR_OK;
}
int rr_chain_set_gain(rr_chain* c, int stage, double gain) {
    Stage* s = nullptr;
    RR_TRY(stage_of(c, stage, RR_STAGE_GAIN, &s));
    s->d.gain = gain;
    return RR_OK;
}

int rr_chain_event(rr_chain* c, int is_interrupt) {
    if (!c) return fail(RR_ERR_INVALID, "null chain");
    if (!is_interrupt) return RR_OK;  // plain events are forwarded untouched by every block
    for (auto& s : c->st) {
        if (s.d.kind == RR_STAGE_FILTER) s.h.f_has_hist = false;   // filters.rs:262-267
        if (s.d.kind == RR_STAGE_FMDEMOD) s.h.fm_has_prev = false;  // modulation.rs:133-138
    }
    return RR_OK;
}

size_t rr_chain_max_output(rr_chain* c, double sample_rate, size_t chunk_len, size_t n_chunks) {
    if (!c) return 0;
    Shape sh;
    sh.chunk_len = chunk_len;
    sh.n_chunks = n_chunks;
    sh.rate = sample_rate;
    for (auto& s : c->st) {
        StageHost h = s.h;
        StageAct a;
        if (advance_stage(s.d, h, sh, &a) != RR_OK) return 0;
        sh = a.out;
    }
    return sh.len();
}

int rr_chain_push_device(rr_chain* c, double sample_rate, size_t chunk_len, size_t n_chunks, const void* dev_in, size_t in_stride,
                         void* dev_out, size_t out_capacity, size_t out_stride, size_t* out_count, double* out_sample_rate) {
    if (!c) return fail(RR_ERR_INVALID, "null chain");
    if (out_count) *out_count = 0;
    if (n_chunks == 0 || chunk_len == 0) return RR_OK;
    if (!dev_in) return fail(RR_ERR_INVALID, "rr_chain_push_device: dev_in is null");
    if (chunk_len > (size_t)1 << 30 || n_chunks > (size_t)1 << 30) return fail(RR_ERR_INVALID, "push too large");
    RR_CUDA(cudaSetDevice(c->ctx->device));
    int r;
    if (c->dtype == RR_C32)
        r = run_push<float>(c, sample_rate, chunk_len, n_chunks, dev_in, in_stride, dev_out, out_capacity, out_stride, out_count,
                            out_sample_rate);
    else
        r = run_push<double>(c, sample_rate, chunk_len, n_chunks, dev_in, in_stride, dev_out, out_capacity, out_stride, out_count,
                             out_sample_rate);
    return r;
}

int rr_chain_push(rr_chain* c, double sample_rate, size_t chunk_len, size_t n_chunks, const void* host_in, size_t in_stride,
                  void* host_out, size_t out_capacity, size_t out_stride, size_t* out_count, double* out_sample_rate) {
    if (!c) return fail(RR_ERR_INVALID, "null chain");
    if (out_count) *out_count = 0;
    if (n_chunks == 0 || chunk_len == 0) return RR_OK;
    if (!host_in) return fail(RR_ERR_INVALID, "rr_chain_push: host_in is null");
    RR_CUDA(cudaSetDevice(c->ctx->device));
    const size_t len = chunk_len * n_chunks;
    const size_t S = (size_t)c->S;
    if (S > 1 && in_stride < len) return fail(RR_ERR_INVALID, "in_stride smaller than the pushed samples");
    const size_t max_out = rr_chain_max_output(c, sample_rate, chunk_len, n_chunks);
    if (max_out > out_capacity) return fail(RR_ERR_CAPACITY, "output buffer too small for this push");
    RR_TRY(c->host_in.ensure(S * len * c->esz));
    RR_TRY(c->host_out.ensure(S * (max_out ? max_out : 1) * c->esz));
    RR_CUDA(cudaMemcpy2DAsync(c->host_in.p, len * c->esz, host_in, in_stride * c->esz, len * c->esz, S, cudaMemcpyHostToDevice,
                              c->stream));
    size_t produced = 0;
    RR_TRY(rr_chain_push_device(c, sample_rate, chunk_len, n_chunks, c->host_in.p, len, c->host_out.p, max_out, max_out ? max_out : 1,
                                &produced, out_sample_rate));
    if (produced > 0) {
        if (!host_out) return fail(RR_ERR_INVALID, "rr_chain_push: host_out is null");
        RR_CUDA(cudaMemcpy2DAsync(host_out, out_stride * c->esz, c->host_out.p, (max_out ? max_out : 1) * c->esz, produced * c->esz, S,
                                  cudaMemcpyDeviceToHost, c->stream));
    }
    if (out_count) *out_count = produced;
    return RR_OK;
}

int rr_chain_sync(rr_chain* c) {
    if (!c) return fail(RR_ERR_INVALID, "null chain");
    RR_CUDA(cudaSetDevice(c->ctx->device));
    RR_CUDA(cudaStreamSynchronize(c->stream));
    return RR_OK;
}
int rr_chain_set_timing(rr_chain* c, int enable) {
    if (!c) return fail(RR_ERR_INVALID, "null chain");
    RR_CUDA(cudaSetDevice(c->ctx->device));
    c->timing = enable != 0;
    c->ev_used = 0;
    return RR_OK;
}
int rr_chain_kernel_time(rr_chain* c, double* total_ms, int* n_launches, const char** kernel_name) {
    if (!c || !total_ms) return fail(RR_ERR_INVALID, "null argument");
    RR_CUDA(cudaSetDevice(c->ctx->device));
    double sum = 0.0;
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float ms = 0.f;
        RR_CUDA(cudaEventSynchronize(c->evs[i + 1]));
        RR_CUDA(cudaEventElapsedTime(&ms, c->evs[i], c->evs[i + 1]));
        sum += ms;
    }
    *total_ms = sum;
    if (n_launches) *n_launches = (int)(c->ev_used / 2);
    if (kernel_name) *kernel_name = c->timed_kernel.c_str();
    return RR_OK;
}
void* rr_chain_cuda_stream(rr_chain* c) { return c ? (void*)c->stream : nullptr; }
const char* rr_chain_plan(rr_chain* c) { return c ? c->plan.c_str() : ""; }

}  // extern "C"

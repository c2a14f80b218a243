/* Plain-C consumer of include/radiorust_b200.h: proves that the header is valid C99 (what a Rust
 * `extern "C"` block, cgo or any other FFI sees), links the shared library and exercises the entry
 * points that need no GPU.  With a device present (argv[1] == "gpu") it also replays the call
 * sequence of a radiorust block task -- recv -> push -> send -- for FreqShifter -> Filter ->
 * Downsampler on one stream.  Built and run by tests/test_c_harness.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "radiorust_b200.h"

static void lowpass(void* user, int64_t bin, double freq_hz, double* re, double* im) {
    (void)bin;
    *re = fabs(freq_hz) <= *(const double*)user ? 1.0 : 0.0;
    *im = 0.0;
}

static int fail(const char* what) {
    fprintf(stderr, "FAIL %s: %s\n", what, rr_last_error());
    return 1;
}

int main(int argc, char** argv) {
    int major = -1, minor = -1;
    if (rr_version(&major, &minor) != RR_OK || major != RR_VERSION_MAJOR || minor != RR_VERSION_MINOR) return fail("rr_version");
    /* src/math.rs:57-85 */
    if (fabs(rr_bessel_i0(0.5) - 1.06348337074132) > 1e-10) return fail("rr_bessel_i0");
    if (fabs(rr_sinc(0.4) - 0.756826728640657) > 1e-10) return fail("rr_sinc");
    int64_t numer = 0, denom = 0;
    if (rr_freq_to_ratio(1024000.0, 1.0, 100000.0, &numer, &denom) != RR_OK || numer != 25 || denom != 256) return fail("rr_freq_to_ratio");
    size_t L = 0;
    if (rr_design_downsampler_taps(2400000.0, 48000.0, 6000.0, 3.0, &L, NULL) != RR_OK || L != 343) return fail("rr_design_downsampler_taps");
    double cutoff = 3000.0, disc = 0.0, err = 0.0;
    int rank = 0;
    if (rr_design_fused_rank(lowpass, &cutoff, RR_WINDOW_KAISER, sqrt(3.0), NULL, NULL, 2400000.0, 4096, 48000.0, 6000.0, 3.0, 2.0e-8, 10,
                             &rank, &disc, &err) != RR_OK || rank != 10 || !(err <= 2.5e-8))
        return fail("rr_design_fused_rank");
    rr_ctx* ctx = NULL;
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) {
        /* no device expected: creation must fail loudly, there is no CPU fallback */
        if (rr_ctx_create(0, &ctx) == RR_OK) {
            rr_ctx_destroy(ctx);
            printf("abi_check ok (a device is present)\n");
            return 0;
        }
        if (strstr(rr_last_error(), "no CPU fallback") == NULL) return fail("rr_ctx_create message");
        printf("abi_check ok (host part)\n");
        return 0;
    }

    /* ---- the block task's call sequence on the device ---- */
    if (rr_ctx_create(0, &ctx) != RR_OK) return fail("rr_ctx_create");
    rr_stage_desc st[3];
    memset(st, 0, sizeof st);
    st[0].kind = RR_STAGE_FREQSHIFT;
    st[0].precision = 1.0;
    st[0].shift = -577000.0;
    st[1].kind = RR_STAGE_FILTER;
    st[1].freq_resp = lowpass;
    st[1].freq_resp_user = &cutoff;
    st[1].window_kind = RR_WINDOW_KAISER;
    st[1].window_beta = sqrt(3.0);
    st[2].kind = RR_STAGE_DOWNSAMPLE;
    st[2].output_chunk_len = 64;
    st[2].output_rate = 48000.0;
    st[2].bandwidth = 6000.0;
    st[2].quality = 3.0;
    rr_chain_desc desc;
    memset(&desc, 0, sizeof desc);
    desc.dtype = RR_C32;
    desc.n_streams = 1;
    desc.n_stages = 3;
    desc.stages = st;
    rr_chain* chain = NULL;
    if (rr_chain_create(ctx, &desc, &chain) != RR_OK) return fail("rr_chain_create");
    const size_t n = 4096, chunks = 12;
    float* in = NULL;
    float* out = NULL;
    const size_t cap = rr_chain_max_output(chain, 2400000.0, n, chunks);
    /* chain edges: chunks come from the pinned pool (ChunkBufPool::get_with_capacity, bufferpool.rs:210-222) and go
     * back through the recycler when their last owner drops them (bufferpool.rs:82-90) */
    rr_pool* pool = NULL;
    if (rr_pool_create(ctx, &pool) != RR_OK) return fail("rr_pool_create");
    {
        void* a = NULL;
        void* b = NULL;
        size_t ca = 0, cb = 0;
        uint64_t n_alloc = 0, n_reuse = 0, n_idle = 0, n_live = 0;
        if (rr_pool_get(pool, 1000, &a, &ca) != RR_OK || ca < 1000) return fail("rr_pool_get");
        if (rr_pool_put(pool, a) != RR_OK) return fail("rr_pool_put");
        if (rr_pool_put(pool, a) == RR_OK) return fail("rr_pool_put accepted a buffer that is not on loan");
        if (rr_pool_get(pool, 500, &b, &cb) != RR_OK || b != a || cb != ca) return fail("rr_pool_get did not recycle");
        if (rr_pool_stats(pool, &n_alloc, &n_reuse, &n_idle, &n_live) != RR_OK || n_alloc != 1 || n_reuse != 1 || n_idle != 0 || n_live != 1)
            return fail("rr_pool_stats");
        if (rr_pool_put(pool, b) != RR_OK) return fail("rr_pool_put (2)");
    }
    if (rr_pool_get(pool, n * chunks * 8, (void**)&in, NULL) != RR_OK || rr_pool_get(pool, (cap + 1) * 8, (void**)&out, NULL) != RR_OK)
        return fail("rr_pool_get (chunks)");
    /* a tone 1 kHz above the shift: after the chain it is a 1 kHz tone at 48 kS/s */
    for (size_t i = 0; i < n * chunks; ++i) {
        const double ph = 2.0 * 3.14159265358979323846 * (577000.0 + 1000.0) * (double)i / 2400000.0;
        in[2 * i] = (float)cos(ph);
        in[2 * i + 1] = (float)sin(ph);
    }
    size_t total = 0;
    double rate = 0.0;
    for (size_t c = 0; c < chunks; c += 4) { /* Signal::Samples messages, four chunks at a time */
        size_t cnt = 0;
        if (rr_chain_push(chain, 2400000.0, n, 4, in + c * n * 2, 4 * n, out + total * 2, cap - total, cap - total, &cnt, &rate) != RR_OK)
            return fail("rr_chain_push");
        if (rr_chain_sync(chain) != RR_OK) return fail("rr_chain_sync");
        total += cnt;
    }
    if (rate != 48000.0 || total == 0 || total % 64 != 0) return fail("output framing");
    /* steady state: |y| constant, phase advancing by 2*pi*1000/48000 per sample */
    double worst = 0.0;
    for (size_t i = total - 200; i + 1 < total; ++i) {
        const double a = atan2(out[2 * i + 1], out[2 * i]), b = atan2(out[2 * i + 3], out[2 * i + 2]);
        double d = b - a - 2.0 * 3.14159265358979323846 * 1000.0 / 48000.0;
        while (d > 3.14159265358979323846) d -= 2.0 * 3.14159265358979323846;
        while (d < -3.14159265358979323846) d += 2.0 * 3.14159265358979323846;
        if (fabs(d) > worst) worst = fabs(d);
    }
    if (worst > 1e-3) {
        fprintf(stderr, "FAIL tone phase step off by %g rad\n", worst);
        return 1;
    }
    if (rr_chain_event(chain, 1) != RR_OK) return fail("rr_chain_event");
    if (rr_chain_samples_lost_count(chain) != 0) return fail("no Rechunker/Overlapper: no SamplesLost");
    {
        /* Rechunker(256) -> FmMod: 100 samples stay in the partial chunk; an event drops them (one SamplesLost);
         * 256 zeros then leave as one chunk of (cos 0, sin 0) */
        rr_stage_desc tx[2];
        memset(tx, 0, sizeof tx);
        tx[0].kind = RR_STAGE_RECHUNK;
        tx[0].output_chunk_len = 256;
        tx[1].kind = RR_STAGE_FMMOD;
        tx[1].deviation = 5000.0;
        rr_chain_desc d2;
        memset(&d2, 0, sizeof d2);
        d2.dtype = RR_C32;
        d2.n_streams = 1;
        d2.n_stages = 2;
        d2.stages = tx;
        rr_chain* c2 = NULL;
        if (rr_chain_create(ctx, &d2, &c2) != RR_OK) return fail("rr_chain_create (tx)");
        memset(in, 0, 256 * 8);
        size_t cnt = 7;
        if (rr_chain_push(c2, 48000.0, 100, 1, in, 100, out, cap, cap, &cnt, &rate) != RR_OK || rr_chain_sync(c2) != RR_OK || cnt != 0)
            return fail("Rechunker holds a partial chunk");
        if (rr_chain_event(c2, 0) != RR_OK || rr_chain_samples_lost_count(c2) != 1) return fail("SamplesLost count");
        if (rr_chain_set_output_chunk_len(c2, 0, 0) == RR_OK) return fail("chunk length 0 accepted");
        if (rr_chain_push(c2, 48000.0, 256, 1, in, 256, out, cap, cap, &cnt, &rate) != RR_OK || rr_chain_sync(c2) != RR_OK || cnt != 256)
            return fail("Rechunker emits a whole chunk");
        for (size_t i = 0; i < 256; ++i)
            if (out[2 * i] != 1.0f || out[2 * i + 1] != 0.0f) return fail("FmMod of silence");
        rr_chain_destroy(c2);
    }
    printf("abi_check ok: %zu output samples at %.0f S/s, plan %s, %llu kernel launches\n", total, rate, rr_chain_plan(chain),
           (unsigned long long)rr_kernel_launch_count());
    if (rr_pool_put(pool, in) != RR_OK || rr_pool_put(pool, out) != RR_OK) return fail("rr_pool_put (chunks)");
    rr_pool_destroy(pool);
    rr_chain_destroy(chain);
    rr_ctx_destroy(ctx);
    return 0;
}

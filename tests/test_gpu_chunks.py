"""GPU tests of the device-side FmMod, Rechunker and Overlapper (SURVEY.md 8f rank 4;
src/blocks/modulation.rs:13-80, src/blocks/chunks.rs:42-248) through the C ABI, against the oracle.

Rechunker / Overlapper move samples without arithmetic: outputs must be bit-equal.  FmMod is a serial
phase recurrence whose rounding the reference fixes step by step; the CUDA kernel keeps that order, so its
phase agrees with the oracle's to the last bits of sincos (checked with an absolute bound far below what a
re-associated prefix sum would leave), relative L2 <= 1e-5 (f32) / 1e-12 (f64).
"""
import math

import numpy as np
import pytest

from oracle import radiorust_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import radiorust_b200 as rr

    c = rr.Context(0)
    yield c
    c.close()


# ---------------------------------------------------------------------------------------------------
# FmMod
# ---------------------------------------------------------------------------------------------------
def fmmod_vectorised(flt, deviation, sr, x, phase0):
    """modulation.rs:44-51 for many streams at once: the loop runs over samples, numpy over streams, every
    operation rounded to Flt in the reference's order (checked against oracle.FmMod below)."""
    R = orc.real_dtype(flt)
    factor = R(deviation / sr * orc.TAU)
    tau = R(orc.TAU)
    cur = phase0.astype(R).copy()
    re = np.ascontiguousarray(x.real.astype(R))
    ph = np.empty(re.shape, dtype=R)
    for i in range(re.shape[1]):
        cur = (cur + (re[:, i] * factor).astype(R)).astype(R)
        cur = np.fmod(cur, tau).astype(R)
        ph[:, i] = cur
    y = np.empty(re.shape, dtype=orc.complex_dtype(flt))
    y.real = np.cos(ph)
    y.imag = np.sin(ph)
    return y, cur


@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_fmmod_vectorised_restatement_is_the_oracle(flt):
    x = np.stack([orc.synth_noise(900 + s, 700, flt) * 3.0 for s in range(3)])
    got, _ = fmmod_vectorised(flt, 75000.0, 240000.0, x, np.zeros(3))
    for s in range(3):
        want = orc.FmMod(flt, 75000.0).process(orc.Samples(240000.0, x[s]))[0].chunk
        assert np.array_equal(got[s], want)


@pytest.mark.parametrize("flt,S,lens", [
    ("f32", 1, [5000, 1, 2047, 2049]),       # one stream per CTA, tiles of 2048
    ("f32", 3, [700, 4100]),
    ("f32", 40, [300, 5000]),
    ("f32", 1300, [1030, 7]),                 # 2 streams per CTA
    ("f32", 2500, [520]),                     # 4
    ("f32", 5000, [300]),                     # 8
    ("f32", 10000, [200, 57]),                # 16
    ("f32", 20000, [130]),                    # 32, tiles of 64
    ("f64", 2, [3000, 2100]),
    ("f64", 777, [600]),
    ("f64", 20011, [70, 65]),
])
def test_fmmod_matches_oracle(ctx, flt, S, lens):
    import radiorust_b200 as rr

    sr, dev = 240000.0, 75000.0
    ch = rr.Chain(ctx, [rr.FmMod(dev)], flt, n_streams=S)
    rng = np.random.default_rng(31 + S)
    phase = np.zeros(S)
    C = orc.complex_dtype(flt)
    for k, n in enumerate(lens):
        x = (rng.standard_normal((S, n)) + 1j * rng.standard_normal((S, n))).astype(C)
        if k == 0:
            x[:, ::50] *= 40.0  # steps of several turns: the general fmod branch
        got, rate = ch.push(sr, x, n)
        assert rate == sr and "fmmod" in ch.plan
        want, phase = fmmod_vectorised(flt, dev, sr, x, phase)
        assert orc.rel_l2(got, want) <= (1e-5 if flt == "f32" else 1e-12)
        # same serial rounding: only sincos differs (a re-associated sum would be off by ~1e-4 in f32 here)
        assert np.max(np.abs(got - want)) <= (2e-6 if flt == "f32" else 1e-14)
    ch.close()


def test_fmmod_deviation_rate_and_events(ctx):
    """set_deviation (modulation.rs:76-79) and a new sample rate change the factor of the next chunk
    (:41-44); events pass through and do NOT reset the phase (:58-60)."""
    import radiorust_b200 as rr

    S = 5
    ch = rr.Chain(ctx, [rr.FmMod(5000.0)], "f32", n_streams=S)
    blocks = [orc.FmMod("f32", 5000.0) for _ in range(S)]
    rng = np.random.default_rng(5)

    def step(sr, n):
        x = (rng.standard_normal((S, n)) + 1j * rng.standard_normal((S, n))).astype(np.complex64)
        got, rate = ch.push(sr, x, n)
        assert rate == sr
        for s in range(S):
            want = blocks[s].process(orc.Samples(sr, x[s]))[0].chunk
            assert np.max(np.abs(got[s] - want)) <= 2e-6

    step(48000.0, 400)
    ch.set_deviation(0, 12000.0)
    for b in blocks:
        b.deviation = 12000.0
    step(48000.0, 300)
    ch.event(True)
    step(48000.0, 200)  # phase continues
    step(96000.0, 333)
    assert ch.samples_lost_count() == 0
    ch.close()


def test_fmmod_fmdemod_round_trip_full_size(ctx):
    """Size-independent property at a BASELINE-sized batch (4096 streams x 4096 samples): FmDemod(FmMod(x)) gives
    back Re x, since the discriminator's arg(x[t] conj x[t-1]) * sr/(dev TAU) undoes phase += re * dev/sr*TAU
    as long as |re * factor| < pi."""
    import radiorust_b200 as rr

    S, n, sr, dev = 4096, 4096, 240000.0, 20000.0
    ch = rr.Chain(ctx, [rr.FmMod(dev), rr.FmDemod(dev)], "f32", n_streams=S)
    rng = np.random.default_rng(77)
    x = rng.uniform(-1.0, 1.0, (S, n)).astype(np.float32).astype(np.complex64)
    y, rate = ch.push(sr, x, n)
    assert y.shape == (S, n) and rate == sr
    # first output repeats "the previous output" (0, modulation.rs:119-124)
    assert np.all(y[:, 0] == 0)
    err = np.abs(y[:, 1:].real - x[:, 1:].real)
    assert err.max() <= 2e-5 and np.all(y.imag == 0)
    ch.close()


# ---------------------------------------------------------------------------------------------------
# Rechunker / Overlapper against the oracle blocks, message by message
# ---------------------------------------------------------------------------------------------------
def oracle_step(blocks, msg):
    """One message through per-stream oracle chains: ([per-stream concatenated samples], chunk lengths, rates,
    number of SamplesLost seen at the chain's end)."""
    per_stream, lens, rates, lost = [], None, None, 0
    for s, chain in enumerate(blocks):
        m = msg if isinstance(msg, orc.Event) else orc.Samples(msg.sample_rate, msg.chunk[s])
        outs = chain.push(m)
        smp = [o for o in outs if isinstance(o, orc.Samples)]
        per_stream.append(np.concatenate([o.chunk for o in smp]) if smp else np.zeros(0, dtype=np.complex64))
        if s == 0:
            lens = [len(o.chunk) for o in smp]
            rates = [o.sample_rate for o in smp]
            lost = sum(1 for o in outs if isinstance(o, orc.Event) and o.name == "SamplesLost")
    return np.stack(per_stream), lens, rates, lost


def run_script(ctx, gpu_stages, make_oracle, script, flt="f32", S=3, exact=True):
    """script: list of ("push", rate, chunk_len, n_chunks) | ("event", interrupt) | ("ocl", stage, len)."""
    import radiorust_b200 as rr

    ch = rr.Chain(ctx, gpu_stages, flt, n_streams=S)
    blocks = [orc.Chain(make_oracle()) for _ in range(S)]
    rng = np.random.default_rng(123)
    C = orc.complex_dtype(flt)
    plans = []
    for step in script:
        lost_before = ch.samples_lost_count()
        if step[0] == "push":
            _, rate, n, c = step
            x = (rng.standard_normal((S, n * c)) + 1j * rng.standard_normal((S, n * c))).astype(C)
            got, got_rate = ch.push(rate, x, n)
            want = None
            lens, rates, lost = [], [], 0
            parts = []
            for k in range(c):
                w, l, r, lo = oracle_step(blocks, orc.Samples(rate, x[:, k * n:(k + 1) * n]))
                parts.append(w)
                lens += l
                rates += r
                lost += lo
            want = np.concatenate(parts, axis=1)
            assert got.shape == want.shape, (step, got.shape, want.shape)
            if exact:
                assert np.array_equal(got, want), step
            elif want.size:
                assert orc.rel_l2(got, want) <= (1e-5 if flt == "f32" else 1e-12), step
            if lens:
                assert got_rate == rates[0] and all(r == rates[0] for r in rates)
            assert ch.samples_lost_count() - lost_before == lost, step
            plans.append(ch.plan)
        elif step[0] == "event":
            ev = orc.Event("test", bool(step[1]))
            _, _, _, lost = oracle_step(blocks, ev)
            ch.event(bool(step[1]))
            assert ch.samples_lost_count() - lost_before == lost, step
        elif step[0] == "ocl":
            ch.set_output_chunk_len(step[1], step[2])
            for b in blocks:
                b.blocks[step[1]].set_output_chunk_len(step[2])
    ch.close()
    return plans


def test_rechunker_reference_test(ctx):
    """chunks.rs:250-271 (`test_rechunker`): 4096-sample chunks into Rechunker(1024) come out 1024 long."""
    import radiorust_b200 as rr

    ch = rr.Chain(ctx, [rr.Rechunker(1024)], "f32", n_streams=1)
    assert ch.max_output(1.0, 4096, 1) == 4096
    y, rate = ch.push(1.0, np.zeros((1, 4096), dtype=np.complex64), 4096)
    assert y.shape == (1, 4096) and rate == 1.0
    # the next stage sees chunks of 1024: a Filter behind it designs for n = 1024 and holds one chunk back
    ch2 = rr.Chain(ctx, [rr.Rechunker(1024), rr.Filter.new(orc.lowpass(0.2))], "f32", n_streams=1)
    y2, _ = ch2.push(1.0, np.zeros((1, 4096), dtype=np.complex64), 4096)
    assert y2.shape == (1, 3072)
    ch.close()
    ch2.close()


@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_rechunker_alone(ctx, flt):
    script = [("push", 48000.0, 100, 1), ("push", 48000.0, 30, 1), ("push", 48000.0, 7, 3), ("push", 48000.0, 1000, 1),
              ("push", 48000.0, 64, 1), ("push", 48000.0, 36, 1), ("push", 48000.0, 128, 2), ("push", 48000.0, 5, 1),
              ("event", False),  # partial chunk dropped, SamplesLost
              ("push", 48000.0, 128, 1), ("event", True),  # nothing pending: no SamplesLost
              ("push", 48000.0, 50, 1), ("push", 44100.0, 100, 1),  # rate change drops the 50 (SamplesLost)
              ("push", 44100.0, 200, 1)]
    plans = run_script(ctx, [__import__("radiorust_b200").Rechunker(128)], lambda: [orc.Rechunker(128)], script, flt)
    assert any("rechunk(hold)" in p for p in plans) and any(p == "rechunk" for p in plans)


def test_rechunker_set_output_chunk_len(ctx):
    """chunks.rs:101-112,171-175: a shorter length with a longer partial chunk pending splits the partial chunk."""
    import radiorust_b200 as rr

    script = [("push", 8000.0, 90, 1), ("ocl", 0, 25), ("push", 8000.0, 3, 1), ("push", 8000.0, 40, 1), ("ocl", 0, 200),
              ("push", 8000.0, 150, 1), ("push", 8000.0, 150, 1), ("ocl", 0, 7), ("push", 8000.0, 1, 1), ("push", 8000.0, 1, 1)]
    run_script(ctx, [rr.Rechunker(100)], lambda: [orc.Rechunker(100)], script)


def test_rechunker_view_feeds_next_stage(ctx):
    """With nothing pending the next stage reads the pushed buffer itself under the new chunk length."""
    import radiorust_b200 as rr

    script = [("push", 48000.0, 4096, 2), ("push", 48000.0, 1000, 1), ("push", 48000.0, 24, 1), ("push", 48000.0, 3072, 1)]
    plans = run_script(ctx, [rr.Rechunker(1024), rr.GainControl(0.5)], lambda: [orc.Rechunker(1024), orc.GainControl("f32", 0.5)], script)
    assert plans[0].startswith("rechunk(view)") and plans[1] == "rechunk(hold)" and plans[2].startswith("rechunk|")
    assert plans[3].startswith("rechunk(view)")


def test_rechunker_in_front_of_the_chain(ctx):
    """The use the reference documents (chunks.rs:35-41): Rechunker gives the Filter its chunk length.  A sample-rate
    change with a partial chunk pending sends SamplesLost, which resets the Filter behind it (it redesigns anyway)."""
    import radiorust_b200 as rr

    lp = orc.lowpass(3000.0)
    gpu = [rr.Rechunker(256), rr.FreqShifter(shift=1000.0), rr.Filter.new(lp), rr.Downsampler(16, 12000.0, 3000.0)]
    mk = lambda: [orc.Rechunker(256), orc.FreqShifter("f32", 1.0, 1000.0), orc.Filter.new("f32", lp), orc.Downsampler("f32", 16, 12000.0, 3000.0)]
    script = [("push", 48000.0, 1000, 1), ("push", 48000.0, 1000, 3), ("push", 48000.0, 24, 1), ("push", 48000.0, 100, 1),
              ("event", True),  # Rechunker drops 100 samples; the Filter starts a new segment
              ("push", 48000.0, 2048, 1), ("push", 48000.0, 10, 1), ("push", 96000.0, 2048, 2), ("push", 96000.0, 512, 4)]
    run_script(ctx, gpu, mk, script, exact=False)


@pytest.mark.parametrize("k", [1, 2, 3, 5])
@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_overlapper(ctx, k, flt):
    import radiorust_b200 as rr

    script = [("push", 48000.0, 64, 1), ("push", 48000.0, 64, 1), ("push", 48000.0, 64, 4), ("push", 48000.0, 64, 1),
              ("event", False),  # any event: history cleared, SamplesLost (chunks.rs:226-233)
              ("push", 48000.0, 100, 7),  # a new chunk length is fine with an empty history
              ("event", True), ("push", 1e6 / 3.0, 9, 2), ("push", 1e6 / 3.0, 9, 11)]
    plans = run_script(ctx, [rr.Overlapper(k)], lambda: [orc.Overlapper(k)], script, flt)
    assert all(p == "overlap" for p in plans)


def test_overlapper_feeds_fourier(ctx):
    """Overlapper(2) -> Fourier: the usual spectrum-display chain (overlapping analysis windows)."""
    import radiorust_b200 as rr

    beta = orc.kaiser_null_at_bin_to_beta(2.0)
    gpu = [rr.Overlapper(2), rr.Fourier(("kaiser", beta), True)]
    mk = lambda: [orc.Overlapper(2), orc.Fourier("f32", orc.Kaiser(beta), True)]
    run_script(ctx, gpu, mk, [("push", 48000.0, 512, 1), ("push", 48000.0, 512, 3), ("push", 48000.0, 512, 1)], exact=False)


def test_overlapper_mixed_history_is_refused(ctx):
    import radiorust_b200 as rr

    ch = rr.Chain(ctx, [rr.Overlapper(3)], "f32", n_streams=2)
    x = np.ones((2, 64), dtype=np.complex64)
    y, _ = ch.push(48000.0, x, 64)
    assert y.shape == (2, 0)
    with pytest.raises(rr.RadiorustError) as e:
        ch.push(48000.0, x[:, :32], 32)
    assert e.value.code == -3
    with pytest.raises(rr.RadiorustError):
        ch.push(44100.0, x, 64)
    # the refused pushes left the history alone: two more chunks complete the first output
    y, _ = ch.push(48000.0, np.concatenate([2 * x, 3 * x], axis=1), 64)
    assert y.shape == (2, 192) and np.array_equal(y[0], np.repeat([1, 2, 3], 64).astype(np.complex64))
    ch.close()


def test_create_rejects_bad_parameters(ctx):
    """chunks.rs:58,195: assert!(output_chunk_len > 0), assert!(chunk_count > 0)."""
    import radiorust_b200 as rr

    with pytest.raises(rr.RadiorustError, match="chunk length must be positive"):
        rr.Chain(ctx, [rr.Rechunker(0)])
    with pytest.raises(rr.RadiorustError, match="chunk count must be positive"):
        rr.Chain(ctx, [rr.Overlapper(0)])
    ch = rr.Chain(ctx, [rr.Rechunker(8)])
    with pytest.raises(rr.RadiorustError, match="chunk length must be positive"):
        ch.set_output_chunk_len(0, 0)
    ch.close()

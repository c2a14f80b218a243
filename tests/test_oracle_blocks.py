"""Oracle self-consistency: the vectorised oracle blocks against the literal
per-sample transcriptions of the reference loops and against independent
closed forms evaluated in f64.  CPU only."""
import math

import numpy as np
import pytest

from oracle import radiorust_oracle as o


def _run(block, sr, x, n):
    outs = []
    for k in range(0, len(x), n):
        for m in block.process(o.Samples(sr, x[k : k + n])):
            if isinstance(m, o.Samples):
                outs.append(m.chunk)
    return np.concatenate(outs) if outs else np.zeros(0, x.dtype)


@pytest.mark.parametrize("flt", ["f32", "f64"])
@pytest.mark.parametrize("shift", [100.0, -37.0, 0.0, 499.0])
def test_freqshifter_matches_literal_table(flt, shift):
    sr, n = 1000.0, 64
    x = o.synth_noise(1, 5 * n, flt)
    fs = o.FreqShifter(flt, 1.0, shift)
    y = _run(fs, sr, x, n)
    tab = o.literal_freqshift_table(flt, sr, 1.0, shift)
    idx = np.arange(len(x)) % len(tab)
    want = (x * tab[idx]).astype(x.dtype)
    assert np.array_equal(y, want)


def test_freqshifter_closed_form_and_retune():
    sr, n = 1024000.0, 4096
    x = o.synth_noise(2, 4 * n, "f64")
    fs = o.FreqShifter("f64", 1.0, 123457.0)
    y0 = _run(fs, sr, x[: 2 * n], n)
    t = np.arange(2 * n)
    want = x[: 2 * n] * np.exp(2j * np.pi * ((123457 * t) % 1024000) / 1024000.0)
    assert o.rel_l2(y0, want) < 1e-13
    # retune keeps the phase continuous (transform.rs:322-327)
    fs.set_shift(-5000.0)
    y1 = _run(fs, sr, x[2 * n :], n)
    ph0 = 2 * np.pi * ((123457 * (2 * n)) % 1024000) / 1024000.0
    t1 = np.arange(2 * n)
    want1 = x[2 * n :] * np.exp(1j * (ph0 + 2 * np.pi * (-5000.0) * t1 / sr))
    assert o.rel_l2(y1, want1) < 1e-12


def test_freq_to_ratio():
    assert o.freq_to_ratio(1024000.0, 1.0, 100000.0) == (25, 256)
    assert o.freq_to_ratio(1024000.0, 1.0, 123457.0) == (123457, 1024000)
    assert o.freq_to_ratio(48000.0, 1.0, 0.0) == (0, 1)
    assert o.freq_to_ratio(48000.0, 1.0, -700.0) == (-7, 480)


@pytest.mark.parametrize("flt,tol", [("f32", 3e-6), ("f64", 1e-13)])
def test_filter_is_linear_convolution(flt, tol):
    """Overlap-save == linear convolution with the designed impulse response
    (SURVEY.md 8a-3), including the one-chunk start-up delay."""
    sr, n = 48000.0, 256
    x = o.synth_noise(3, 6 * n, flt)
    f = o.Filter.new(flt, o.lowpass(3000.0))
    y = _run(f, sr, x, n)
    assert len(y) == 5 * n  # first chunk is history only (filters.rs:79-81)
    H = o.design_filter_response(o.lowpass(3000.0), o.Kaiser.with_null_at_bin(2.0), sr, n, "f64")
    h = np.fft.ifft(H)[n:]  # ext = [0]*n || h, unnormalised inverse folded into the scale
    h = h * (2 * n) / (2 * n)
    xx = x.astype(np.complex128)
    full = np.convolve(xx, h * (2 * n))  # undo 1/N of np.fft.ifft
    # z[i] = sum_m h[m] x[i + n - m]  ->  full[i + n]
    want = full[n : n + 5 * n]
    assert o.rel_l2(y, want) < tol


def test_filter_dc_gain_and_redesign_drops_history():
    sr, n = 48000.0, 128
    f = o.Filter.new("f64", o.lowpass(6000.0))
    ones = np.ones(n, dtype=np.complex128)
    assert f.process(o.Samples(sr, ones)) == []
    y = f.process(o.Samples(sr, ones))[0].chunk
    assert np.allclose(y, 1.0, atol=2e-2)  # unity pass-band gain
    f.update(o.lowpass(3000.0))
    assert f.process(o.Samples(sr, ones)) == []  # filters.rs:187
    assert len(f.process(o.Samples(sr, ones))) == 1
    # interrupt event clears history too (filters.rs:262-267)
    ev = f.process(o.DISCONNECTION)
    assert ev == [o.DISCONNECTION]
    assert f.process(o.Samples(sr, ones)) == []
    # chunk length change -> redesign
    assert f.process(o.Samples(sr, np.ones(64, dtype=np.complex128))) == []


@pytest.mark.parametrize("flt", ["f32", "f64"])
@pytest.mark.parametrize("rates", [(1024000.0, 48000.0, 6000.0), (48000.0, 48000.0 / 5, 2000.0), (1000.0, 333.0, 100.0)])
def test_downsampler_matches_literal(flt, rates):
    in_rate, out_rate, bw = rates
    n = 512
    x = o.synth_noise(4, 3 * n, flt)
    d = o.Downsampler(flt, 7, out_rate, bw)
    y = _run(d, in_rate, x, n)
    lit = o.literal_downsample(x, d.ir, in_rate, out_rate)
    k = (len(lit) // 7) * 7
    assert len(y) == k
    assert np.array_equal(y, lit[:k])


def test_downsampler_closed_form_and_taps():
    """j_m = ceil(m*in/out); unit-energy taps; L as SURVEY.md 8a-5."""
    in_rate, out_rate = 1024000.0, 48000.0
    d = o.Downsampler("f64", 1, out_rate, 6000.0)
    x = o.synth_noise(5, 4096, "f64")
    y = _run(d, in_rate, x, 1024)
    L = len(d.ir)
    assert L == 147
    assert abs(float(np.sum(d.ir.astype(np.float64) ** 2)) - 1.0) < 1e-12
    xp = np.concatenate([np.zeros(L, dtype=x.dtype), x])
    for m in (1, 2, 3, 4, 50, len(y)):
        j = math.ceil(m * in_rate / out_rate)
        want = np.sum(d.ir * xp[L + j - L : L + j])
        assert abs(y[m - 1] - want) < 1e-12
    assert len(y) == 192
    assert o.Downsampler("f32", 1, 48000.0, 6000.0)._design(2400000.0) is None
    dd = o.Downsampler("f32", 1, 48000.0, 6000.0)
    dd._design(2400000.0)
    assert len(dd.ir) == 343
    dd = o.Downsampler("f32", 1, 48000.0, 40000.0)
    dd._design(10e6)
    assert len(dd.ir) == 7500
    dd = o.Downsampler("f32", 1, 48000.0, 6000.0)
    dd._design(20e6)
    assert len(dd.ir) == 2858


@pytest.mark.parametrize("flt", ["f32", "f64"])
@pytest.mark.parametrize("rates", [(48000.0, 2400000.0, 20000.0), (1000.0, 3300.0, 400.0), (48000.0, 48000.0, 10000.0)])
def test_upsampler_matches_literal(flt, rates):
    in_rate, out_rate, bw = rates
    n = 40
    x = o.synth_noise(6, 3 * n, flt)
    u = o.Upsampler(flt, 1, out_rate, bw)
    y = _run(u, in_rate, x, n)
    lit = o.literal_upsample(x, u.ir, in_rate, out_rate)
    assert np.array_equal(y, lit)


def test_upsampler_closed_form():
    in_rate, out_rate = 48000.0, 2400000.0
    u = o.Upsampler("f64", 1, out_rate, 20000.0)
    x = o.synth_noise(7, 64, "f64")
    y = _run(u, in_rate, x, 16)
    L = len(u.ir)
    assert L == 515
    assert len(y) == 64 * 50
    z = np.zeros(64 * 50 + L, dtype=np.complex128)
    for p in range(64):
        q = math.ceil(p * out_rate / in_rate)
        z[q : q + L] += x[p] * u.ir
    assert o.rel_l2(y, z[: len(y)]) < 1e-14


@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_fmdemod_matches_literal_and_interrupt(flt):
    sr = 384000.0
    x = o.synth_noise(8, 300, flt)
    d = o.FmDemod(flt, 75000.0)
    y = _run(d, sr, x, 100)
    # numpy's vectorised and scalar f32 atan2/complex-mul paths differ by <= 1 ulp
    assert np.allclose(y, o.literal_fmdemod(x, sr, 75000.0), rtol=1e-6 if flt == "f32" else 1e-14, atol=0)
    assert y[0] == 0 and np.all(y.imag == 0)
    # interrupt: previous_sample = None, the previous *output* is repeated
    d.process(o.DISCONNECTION)
    y2 = d.process(o.Samples(sr, x[:10]))[0].chunk
    assert y2[0] == y[-1]
    assert np.allclose(y2[1:], o.literal_fmdemod(x[:10], sr, 75000.0)[1:], rtol=1e-6, atol=0)


def test_fm_roundtrip():
    """FmMod -> FmDemod recovers the (band-limited) message."""
    sr, dev = 240000.0, 75000.0
    t = np.arange(4000)
    msg = 0.5 * np.sin(2 * np.pi * 1000.0 * t / sr)
    m = o.FmMod("f64", dev)
    iq = m.process(o.Samples(sr, msg.astype(np.complex128)))[0].chunk
    d = o.FmDemod("f64", dev)
    y = d.process(o.Samples(sr, iq))[0].chunk
    assert np.max(np.abs(y.real[1:] - msg[1:])) < 1e-9


def test_chain_c1_shapes():
    """Config C1 wiring: 4096-sample chunks -> exactly 192 outputs/chunk after
    the Filter's one-chunk delay."""
    sr, n = 1024000.0, 4096
    x = o.synth_noise(20260000 + 100000, 4 * n, "f32")
    ch = o.Chain([o.FreqShifter("f32", 1.0, 100000.0), o.Filter.new("f32", o.lowpass(3000.0)), o.Downsampler("f32", 192, 48000.0, 6000.0)])
    y = ch.run(sr, x, n)
    assert len(y) == 3 * 192
    # events pass through in-band and in order
    msgs = ch.push(o.Event("x", False))
    assert len(msgs) == 1 and isinstance(msgs[0], o.Event)


# ---------------------------------------------------------------------------------------------------
# chunks.rs: Rechunker / Overlapper
# ---------------------------------------------------------------------------------------------------
def test_rechunker_reference_test():
    """chunks.rs:250-271 (`test_rechunker`): a 4096-sample chunk into Rechunker(1024) -> the first message out is a
    1024-sample chunk (and so are the other three)."""
    r = o.Rechunker(1024)
    outs = r.process(o.Samples(1.0, np.zeros(4096, dtype=np.uint8)))
    assert [len(m.chunk) for m in outs] == [1024] * 4 and all(m.sample_rate == 1.0 for m in outs)


def test_rechunker_is_a_fifo_and_reports_losses():
    rng = np.random.default_rng(3)
    x = rng.integers(0, 1 << 30, 5000)
    r = o.Rechunker(128)
    got, pos = [], 0
    for n in [100, 30, 7, 7, 7, 1000, 64, 36, 256, 5, 1285, 1103]:
        for m in r.process(o.Samples(8000.0, x[pos:pos + n])):
            assert len(m.chunk) == 128
            got.append(m.chunk)
        pos += n
    assert pos == 3900 and np.array_equal(np.concatenate(got), x[: 128 * (3900 // 128)])
    assert r.patchwork is not None and len(r.patchwork[1]) == 3900 % 128
    # an event drops the partial chunk and is preceded by SamplesLost (chunks.rs:84-91) ...
    ev = o.Event("retune", False)
    outs = r.process(ev)
    assert outs == [o.SAMPLES_LOST, ev] and o.SAMPLES_LOST.is_interrupt() and r.patchwork is None
    assert r.process(ev) == [ev]  # ... only when there is one
    # so does a sample-rate change (chunks.rs:71-79)
    r.process(o.Samples(8000.0, x[:50]))
    outs = r.process(o.Samples(16000.0, x[:200]))
    assert outs[0] is o.SAMPLES_LOST and np.array_equal(outs[1].chunk, x[:128]) and outs[1].sample_rate == 16000.0
    # a shorter length splits a longer partial chunk at the next input (chunks.rs:101-112)
    r.set_output_chunk_len(20)
    outs = r.process(o.Samples(16000.0, x[200:203]))
    assert [len(m.chunk) for m in outs] == [20, 20, 20] and np.array_equal(np.concatenate([m.chunk for m in outs]), x[128:188])


def test_overlapper():
    ov = o.Overlapper(3)
    chunks = [np.arange(4) + 10 * i for i in range(6)]
    assert ov.process(o.Samples(2.0, chunks[0])) == [] and ov.process(o.Samples(2.0, chunks[1])) == []
    for i in range(2, 5):
        (m,) = ov.process(o.Samples(2.0, chunks[i]))
        assert np.array_equal(m.chunk, np.concatenate(chunks[i - 2:i + 1])) and m.sample_rate == 2.0
    ev = o.Event("x", False)
    assert ov.process(ev) == [o.SAMPLES_LOST, ev] and ov.history == []  # chunks.rs:226-233, even for plain events
    assert ov.process(o.Samples(2.0, chunks[5])) == []
    # mixed rates: the length-weighted mean (chunks.rs:207-214)
    ov = o.Overlapper(2)
    ov.process(o.Samples(1000.0, np.zeros(10)))
    (m,) = ov.process(o.Samples(4000.0, np.zeros(30)))
    assert m.sample_rate == (1000.0 * 10 + 4000.0 * 30) / 40 and len(m.chunk) == 40
    (m,) = o.Overlapper(1).process(o.Samples(5.0, chunks[0]))
    assert np.array_equal(m.chunk, chunks[0])

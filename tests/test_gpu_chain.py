"""GPU parity tests: the CUDA chain (through the C ABI) against the numpy oracle.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-5 for f32 and
<= 1e-12 for f64, over all output samples.
"""
import math

import numpy as np
import pytest

from oracle import radiorust_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f64": 1e-12}


@pytest.fixture(scope="module")
def ctx():
    import radiorust_b200 as rr

    c = rr.Context(0)
    yield c
    c.close()


def noise(seed, n, flt):
    return orc.synth_noise(seed, n, flt)


def oracle_blocks(stages, flt):
    import radiorust_b200 as rr

    out = []
    for s in stages:
        if isinstance(s, rr.FreqShifter):
            out.append(orc.FreqShifter(flt, s.precision, s.shift))
        elif isinstance(s, rr.Filter):
            w = s.window
            if callable(w):
                win = orc.CustomWindow(w)
            elif w[0] == "kaiser":
                win = orc.Kaiser(w[1])
            else:
                win = orc.Rectangular()
            out.append(orc.Filter(flt, s.freq_resp, win))
        elif isinstance(s, rr.Downsampler):
            out.append(orc.Downsampler(flt, s.output_chunk_len, s.output_rate, s.bandwidth, s.quality))
        elif isinstance(s, rr.Upsampler):
            out.append(orc.Upsampler(flt, s.output_chunk_len, s.output_rate, s.bandwidth, s.quality))
        elif isinstance(s, rr.FmDemod):
            out.append(orc.FmDemod(flt, s.deviation))
        elif isinstance(s, rr.GainControl):
            out.append(orc.GainControl(flt, s.gain))
    return out


def run_both(ctx, stages, flt, sr, x, chunk_len, pushes=None, n_streams=1):
    """x: [S, total]. pushes: list of chunk counts per push (default: one push)."""
    import radiorust_b200 as rr

    x = np.atleast_2d(x)
    total_chunks = x.shape[1] // chunk_len
    if pushes is None:
        pushes = [total_chunks]
    assert sum(pushes) == total_chunks
    ch = rr.Chain(ctx, stages, flt, n_streams=x.shape[0])
    got = []
    pos = 0
    for k in pushes:
        seg = np.ascontiguousarray(x[:, pos * chunk_len : (pos + k) * chunk_len])
        y, rate = ch.push(sr, seg, chunk_len)
        got.append(y.copy())
        pos += k
    got = np.concatenate(got, axis=1)
    plan = ch.plan
    ch.close()
    want = []
    for s in range(x.shape[0]):
        oc = orc.Chain(oracle_blocks(stages, flt))
        want.append(oc.run(sr, x[s], chunk_len))
    want = np.stack(want)
    return got, want, plan


def check(got, want, flt, scale=1.0):
    assert got.shape == want.shape, (got.shape, want.shape)
    assert want.size > 0
    err = orc.rel_l2(got, want)
    assert err <= TOL[flt] * scale, err
    return err


# ---------------------------------------------------------------------------
# BASELINE config 1: FreqShifter -> Filter(3 kHz low-pass) -> Downsampler, n = 4096
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("shift", [123457.0, 100000.0, -250000.0, 0.0])
def test_config1_chain_f32(ctx, shift):
    import radiorust_b200 as rr

    sr, n = 1_024_000.0, 4096
    x = noise(20260000 + 1 * 100000, 24 * n, "f32")
    stages = [rr.FreqShifter(shift), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(192, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f32", sr, x, n, pushes=[1, 2, 5, 16])
    assert "poly" in plan or "fused_os" in plan
    check(got, want, "f32")


def test_config1_chain_f64(ctx):
    import radiorust_b200 as rr

    sr, n = 1_024_000.0, 4096
    x = noise(20260000 + 1 * 100000 + 7, 12 * n, "f64")
    stages = [rr.FreqShifter(123457.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(192, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f64", sr, x, n, pushes=[3, 9])
    check(got, want, "f64")


@pytest.mark.parametrize("flt,n", [("f32", 32), ("f32", 64), ("f32", 256), ("f32", 1024), ("f32", 8192), ("f64", 32), ("f64", 512), ("f64", 2048)])
def test_filter_only_sizes(ctx, flt, n):
    import radiorust_b200 as rr

    sr = 48000.0
    x = noise(11 + n, 7 * n, flt)

    def resp(b, f):
        return complex(1.0 / (1.0 + abs(f) / 3000.0), 0.1 * math.copysign(1.0, f) if f else 0.0)

    got, want, plan = run_both(ctx, [rr.Filter.new(resp)], flt, sr, x, n, pushes=[1, 1, 5])
    assert got.shape[1] == 6 * n  # one-chunk start-up delay (filters.rs:79-81)
    check(got, want, flt)


def test_filter_windows(ctx):
    import radiorust_b200 as rr

    sr, n = 48000.0, 512
    x = noise(5, 5 * n, "f32")
    for st in (
        rr.Filter.new_rectangular(lambda b, f: orc.deemphasis_factor(50e-6, f) if (b != 0 and 20 <= abs(f) <= 16000) else 0j),
        rr.Filter.with_window(orc.lowpass(5000.0), ("kaiser", 6.0)),
        rr.Filter.with_window(orc.lowpass(5000.0), lambda v: 1.0 - 0.5 * v * v),
    ):
        got, want, _ = run_both(ctx, [st], "f32", sr, x, n)
        check(got, want, "f32")


def test_multi_stream_independent_shifts(ctx):
    import radiorust_b200 as rr

    sr, n, S = 2_400_000.0, 4096, 5
    x = np.stack([noise(20260000 + 3 * 100000 + s, 6 * n, "f32") for s in range(S)])
    shifts = [(s * 577) % 2_400_000 - 1_200_000 for s in range(S)]
    stages = [rr.FreqShifter(0.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)]
    ch = rr.Chain(ctx, stages, "f32", n_streams=S)
    ch.set_shifts(0, shifts)
    got, _ = ch.push(sr, x, n)
    got = got.copy()
    ch.close()
    for s in range(S):
        oc = orc.Chain([orc.FreqShifter("f32", 1.0, shifts[s]), orc.Filter.new("f32", orc.lowpass(3000.0)), orc.Downsampler("f32", 64, 48000.0, 6000.0)])
        want = oc.run(sr, x[s], n)
        check(got[s : s + 1], want[None, :], "f32")


@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_standalone_stages(ctx, flt):
    """Blocks on their own (unfused kernels): FreqShifter, Downsampler, Upsampler, FmDemod, GainControl."""
    import radiorust_b200 as rr

    sr = 48000.0
    x = noise(3, 6000, flt)
    got, want, _ = run_both(ctx, [rr.FreqShifter(1234.0, 0.5)], flt, sr, x, 1000, pushes=[1, 2, 3])
    check(got, want, flt)
    got, want, _ = run_both(ctx, [rr.GainControl(0.25)], flt, sr, x, 1000)
    assert np.array_equal(got, want)  # exact products (transform.rs:397-416)
    got, want, _ = run_both(ctx, [rr.Downsampler(50, 8000.0, 3000.0)], flt, sr, x, 1000, pushes=[1, 2, 3])
    check(got, want, flt)
    got, want, _ = run_both(ctx, [rr.Downsampler(7, 44100.0, 15000.0)], flt, sr, x, 600, pushes=[4, 6])
    check(got, want, flt)
    got, want, _ = run_both(ctx, [rr.Upsampler(128, 240000.0, 20000.0)], flt, sr, x[:3000], 500, pushes=[1, 2, 3])
    check(got, want, flt)
    got, want, _ = run_both(ctx, [rr.FmDemod(5000.0)], flt, sr, x, 1000, pushes=[2, 4])
    check(got, want, flt)


def test_events_and_retune(ctx):
    """Interrupt events drop Filter history / FmDemod previous sample; retunes are phase continuous."""
    import radiorust_b200 as rr

    sr, n = 1_024_000.0, 1024
    # an FM station 100 kHz below the tuner's centre (a discriminator fed with plain noise measures input rounding
    # only); deviation 10 kHz + 3 kHz audio = +-13 kHz, so the retune (2 kHz off centre) and the narrower Filter keep it whole
    x = orc.synth_fm_station(99, 12 * n, sr, 10000.0, 3000.0, -30.0, "f64")
    x = (x * np.exp(-2j * np.pi * 100000.0 / sr * np.arange(12 * n))).astype(np.complex64)
    stages = [rr.FreqShifter(100000.0), rr.Filter.new(orc.lowpass(20000.0)), rr.FmDemod(30000.0), rr.Downsampler(48, 48000.0, 12000.0)]
    ch = rr.Chain(ctx, stages, "f32")
    ob = oracle_blocks(stages, "f32")
    oc = orc.Chain(ob)
    got, want = [], []

    def feed(lo, hi):
        y, _ = ch.push(sr, x[lo * n : hi * n], n)
        got.append(y[0].copy())
        for k in range(lo, hi):
            for m in oc.push(orc.Samples(sr, x[k * n : (k + 1) * n])):
                if isinstance(m, orc.Samples):
                    want.append(m.chunk)

    feed(0, 3)
    ch.event(True)
    oc.push(orc.DISCONNECTION)
    feed(3, 6)
    ch.set_shift(0, 98000.0)
    ob[0].set_shift(98000.0)
    feed(6, 9)
    ch.update_filter(1, orc.lowpass(16000.0))
    ob[1].update(orc.lowpass(16000.0))
    ch.set_deviation(2, 15000.0)
    ob[2].set_deviation(15000.0)
    feed(9, 12)
    ch.close()
    g, w = np.concatenate(got), np.concatenate(want)
    check(g[None, :], w[None, :], "f32")


def test_big_overlap_save_f32(ctx):
    """Filter with N = 131072 (config 2's chunk length) through the four-step FFT."""
    import radiorust_b200 as rr

    sr, n = 20_000_000.0, 65536
    x = noise(20260000 + 2 * 100000, 4 * n, "f32")
    got, want, plan = run_both(ctx, [rr.Filter.new(orc.lowpass(3000.0))], "f32", sr, x, n, pushes=[1, 3])
    assert "big_os" in plan
    check(got, want, "f32")


def test_big_overlap_save_f64(ctx):
    """f64 long-kernel case (config 5 shape at reduced size: N = 2^17)."""
    import radiorust_b200 as rr

    sr, n = 2_400_000.0, 65536
    x = noise(20260000 + 5 * 100000, 3 * n, "f64")
    got, want, plan = run_both(ctx, [rr.Filter.new(orc.lowpass(20000.0))], "f64", sr, x, n)
    assert "big_os" in plan
    check(got, want, "f64")


@pytest.mark.parametrize("flt", ["f32", "f64"])
@pytest.mark.parametrize("n", [16384, 32768, 131072, 262144, 524288])
def test_long_filter_every_shape(ctx, flt, n):
    """Every column shape of the streaming four-step kernels (rr_long_os.cu: 2n = Na x 1024, Na = 32 ... 1024; n = 65536
    is covered above), two streams, pushes of one and of two chunks."""
    import radiorust_b200 as rr

    sr = 2_400_000.0
    x = np.stack([noise(31337 + n + s, 3 * n, flt) for s in range(2)])
    got, want, plan = run_both(ctx, [rr.Filter.new(orc.lowpass(20000.0))], flt, sr, x, n, pushes=[1, 2])
    assert "big_os" in plan
    assert got.shape[1] == 2 * n
    check(got, want, flt)


def test_config2_chain(ctx):
    """20 MS/s stream, n = 65536: FreqShifter -> Filter -> Downsampler(48 kS/s), L = 2858."""
    import radiorust_b200 as rr

    sr, n = 20_000_000.0, 65536
    x = noise(20260000 + 2 * 100000 + 1, 5 * n, "f32")
    stages = [rr.FreqShifter(1_234_567.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(128, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f32", sr, x, n, pushes=[2, 3])
    check(got, want, "f32")


def test_errors(ctx):
    import radiorust_b200 as rr

    with pytest.raises(rr.RadiorustError):
        rr.Chain(ctx, [rr.Downsampler(16, 48000.0, 60000.0)], "f32")  # bandwidth >= output rate (resampling.rs:53-56)
    ch = rr.Chain(ctx, [rr.Downsampler(16, 48000.0, 6000.0)], "f32")
    with pytest.raises(rr.RadiorustError):  # input rate below output rate (resampling.rs:78-81)
        ch.push(8000.0, noise(1, 256, "f32"), 256)
    ch.close()


# ---------------------------------------------------------------------------
# polyphase fast path (rr_poly.cuh): same results as the stateful overlap-save path
# ---------------------------------------------------------------------------
def test_poly_path_is_used_and_matches(ctx):
    import radiorust_b200 as rr

    sr, n = 2_400_000.0, 4096
    x = noise(20260000 + 3 * 100000 + 17, 40 * n, "f32")
    stages = [rr.FreqShifter(-577000.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f32", sr, x, n, pushes=[1, 1, 1, 7, 30])
    assert "poly" in plan, plan
    check(got, want, "f32")
    # and against the stateful path of the library itself
    ch = rr.Chain(ctx, stages, "f32")
    ch.set_fast_path(False)
    y, _ = ch.push(sr, x, n)
    assert "poly" not in ch.plan
    ch.close()
    assert orc.rel_l2(got, y) < 2e-6


def test_poly_then_event_then_retune(ctx):
    """The decimator tail must be rebuilt from hist2 after polyphase pushes (interrupts, redesigns, path switches)."""
    import radiorust_b200 as rr

    sr, n = 1_024_000.0, 4096
    x = noise(424242, 40 * n, "f32")
    stages = [rr.FreqShifter(100000.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(100, 48000.0, 6000.0)]
    ch = rr.Chain(ctx, stages, "f32")
    ob = oracle_blocks(stages, "f32")
    oc = orc.Chain(ob)
    got, want, plans = [], [], []

    def feed(lo, hi):
        y, _ = ch.push(sr, x[lo * n : hi * n], n)
        plans.append(ch.plan)
        got.append(y[0].copy())
        for k in range(lo, hi):
            for m in oc.push(orc.Samples(sr, x[k * n : (k + 1) * n])):
                if isinstance(m, orc.Samples):
                    want.append(m.chunk)

    feed(0, 8)           # stateful (2 chunks) + poly
    feed(8, 14)          # poly only
    ch.event(True)       # interrupt: Filter drops history, Downsampler keeps its ring
    oc.push(orc.DISCONNECTION)
    feed(14, 20)
    ch.update_filter(1, orc.lowpass(5000.0))  # redesign after poly pushes
    ob[1].update(orc.lowpass(5000.0))
    feed(20, 26)
    ch.set_shift(0, -200000.0)
    ob[0].set_shift(-200000.0)
    feed(26, 32)
    ch.set_fast_path(False)  # path switch: tail rebuilt, stateful continues
    feed(32, 36)
    ch.set_fast_path(True)
    feed(36, 40)
    ch.close()
    assert any("poly" in p for p in plans)
    g, w = np.concatenate(got), np.concatenate(want)
    check(g[None, :], w[None, :], "f32")


@pytest.mark.parametrize("flt,n,sr,out_rate,bw,ocl", [
    ("f32", 4096, 1_024_000.0, 48000.0, 6000.0, 192),     # C1: P/Q = 64/3
    ("f32", 4096, 2_400_000.0, 48000.0, 6000.0, 2048),    # C3: P = 50
    ("f32", 2048, 960_000.0, 48000.0, 20000.0, 17),       # P = 20, odd output chunking
    ("f32", 1024, 250_000.0, 100000.0, 30000.0, 64),      # P/Q = 5/2
    ("f64", 2048, 1_024_000.0, 48000.0, 6000.0, 96),      # f64, 64/3
    ("f64", 4096, 2_400_000.0, 48000.0, 6000.0, 50),      # f64, P = 50
])
def test_poly_rates(ctx, flt, n, sr, out_rate, bw, ocl):
    import radiorust_b200 as rr

    x = noise(77 + n, 26 * n, flt)
    stages = [rr.FreqShifter(sr / 7.0), rr.Filter.new(orc.lowpass(bw / 2)), rr.Downsampler(ocl, out_rate, bw)]
    got, want, plan = run_both(ctx, stages, flt, sr, x, n, pushes=[3, 10, 13])
    assert "poly" in plan, plan
    check(got, want, flt)


@pytest.mark.parametrize("sr,out_rate,ocl,pushes", [
    (1_024_000.0, 48000.0, 192, [3, 10, 1, 2, 13]),   # C1: P/Q = 64/3
    (1_200_000.0, 32000.0, 100, [2, 9, 15]),          # 75/2 ... odd P: k_front_wide
    (1_120_000.0, 64000.0, 64, [4, 1, 1, 12]),        # 35/2 ... odd P as well
    (1_536_000.0, 64000.0, 64, [4, 1, 1, 12]),        # 24/1 is Q = 1; 1 536 000 / 64 000 = 24
    (2_048_000.0, 96000.0, 128, [3, 8, 1, 14]),       # 64/3 at another rate
])
def test_front_end_shared_by_output_phases(ctx, sr, out_rate, ocl, pushes):
    """in/out = P/Q with Q > 1: the steady state runs the rank-reduced front end (sixteen or thirty-two columns shared by
    the Q phases; k_front for even P, k_front_wide for the others) and k_poly on u; several streams with their own shifts,
    pushes of one chunk and of many."""
    import radiorust_b200 as rr

    n, S = 4096, 3
    x = np.stack([noise(9100 + s, sum(pushes) * n, "f32") for s in range(S)])
    stages = [rr.FreqShifter(0.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(ocl, out_rate, 6000.0)]
    ch = rr.Chain(ctx, stages, "f32", n_streams=S)
    shifts = [sr / 7.0, -123456.0, 0.0]
    ch.set_shifts(0, shifts)
    got, plans, pos = [], [], 0
    for k in pushes:
        y, _ = ch.push(sr, np.ascontiguousarray(x[:, pos * n : (pos + k) * n]), n)
        got.append(y.copy())
        plans.append(ch.plan)
        pos += k
    ch.close()
    got = np.concatenate(got, axis=1)
    g = math.gcd(int(sr), int(out_rate))
    P, Q = int(sr) // g, int(out_rate) // g
    if Q > 1:
        assert any("front+poly[" in p for p in plans), plans
    for s in range(S):
        oc = orc.Chain([orc.FreqShifter("f32", 1.0, shifts[s]), orc.Filter.new("f32", orc.lowpass(3000.0)),
                        orc.Downsampler("f32", ocl, out_rate, 6000.0)])
        want = oc.run(sr, x[s], n)
        assert want.shape == got[s].shape, (want.shape, got[s].shape)
        assert orc.rel_l2(got[s], want) <= TOL["f32"]


@pytest.mark.parametrize("flt", ["f32", "f64"])
@pytest.mark.parametrize("sr,out_rate,bw,n", [
    (48000.0, 2_400_000.0, 20000.0, 1024),   # C5: factor 50, 11 taps per phase (k_upsample_phase)
    (48000.0, 384000.0, 20000.0, 512),       # factor 8
    (12000.0, 3_072_000.0, 5000.0, 256),     # factor 256
    (16000.0, 48000.0, 7000.0, 2048),        # factor 3: the tiled kernel
    (8000.0, 2_400_000.0, 3000.0, 128),      # factor 300: the tiled kernel
])
def test_upsampler_integer_factors(ctx, flt, sr, out_rate, bw, n):
    """Upsampler at integer interpolation factors, pushes of one, three and two chunks (the carried partial sums cross
    the push boundaries), two streams."""
    import radiorust_b200 as rr

    ocl = 4096
    x = np.stack([noise(5150 + int(out_rate / sr) + s, 6 * n, flt) for s in range(2)])
    got, want, plan = run_both(ctx, [rr.Upsampler(ocl, out_rate, bw)], flt, sr, x, n, pushes=[1, 3, 2])
    assert "upsample" in plan
    check(got, want, flt)


def test_wide_front_end_without_nco(ctx):
    """in/out = 1250/3 without a FreqShifter in front: k_front_wide<32> with no NCO (rank 21), two streams."""
    import radiorust_b200 as rr

    sr, n = 20_000_000.0, 65536
    x = np.stack([noise(777 + s, 6 * n, "f32") for s in range(2)])
    stages = [rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(128, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f32", sr, x, n, pushes=[3, 1, 2])
    assert "front+poly[" in plan, plan
    check(got, want, "f32")


def test_poly_filter_down_without_nco_multi_stream(ctx):
    import radiorust_b200 as rr

    sr, n, S = 2_400_000.0, 4096, 3
    x = np.stack([noise(900 + s, 30 * n, "f32") for s in range(S)])
    stages = [rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(128, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f32", sr, x, n, pushes=[12, 18])
    assert "poly" in plan
    check(got, want, "f32")


def test_config2_chain_poly(ctx):
    """Config 2 shape: n = 65536 chunks at 20 MS/s, P/Q = 1250/3; big_os start-up then polyphase."""
    import radiorust_b200 as rr

    sr, n = 20_000_000.0, 65536
    x = noise(20260000 + 2 * 100000 + 3, 6 * n, "f32")
    stages = [rr.FreqShifter(1_234_567.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(128, 48000.0, 6000.0)]
    got, want, plan = run_both(ctx, stages, "f32", sr, x, n, pushes=[2, 4])
    assert "poly" in plan, plan
    check(got, want, "f32")

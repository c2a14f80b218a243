"""GPU parity at the shapes BASELINE.json names, at the stated tolerance (relative L2 1e-5 f32 / 1e-12 f64, no
scale factors), through the C ABI against the numpy oracle.

  C1  FreqShifter -> Filter(3 kHz) -> Downsampler, 1.024 MS/s -> 48 kS/s, n = 4096, f32 AND f64
  C2  one 20 MS/s stream, n = 65536 (N = 2^17), Downsampler L = 2858, P/Q = 1250/3
  C4  wide-band FM: 10 MS/s, n = 65536, Filter(|f| <= 100 kHz) -> FmDemod(75 k) -> Filter(de-emphasis 50 us,
      rectangular, 20 Hz..16 kHz, bin != 0) -> Downsampler(4096, 48 k, 40 k) (L = 7500, P/Q = 625/3), stations batched;
      input = FM-modulated band-limited noise + AWGN at -30 dB, chain wiring as
      examples/relm_app/simple_receiver.rs:28-52
  C5  Filter with N = 2^20 (n = 2^19) in f64; Upsampler 48 kS/s -> 2.4 MS/s (L = 515) in f64
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

from oracle import radiorust_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = {"f32": 1e-5, "f64": 1e-12}


@pytest.fixture(scope="module")
def ctx():
    import radiorust_b200 as rr

    c = rr.Context(0)
    yield c
    c.close()


def _push_all(ch, sr, x, n, pushes):
    parts, pos, plans = [], 0, []
    for k in pushes:
        y, rate = ch.push(sr, np.ascontiguousarray(x[:, pos * n : (pos + k) * n]), n)
        parts.append(y.copy())
        plans.append(ch.plan)
        pos += k
    return np.concatenate(parts, axis=1), plans


@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_c1_at_n4096(ctx, flt):
    import radiorust_b200 as rr

    sr, n, k = 1_024_000.0, 4096, 24
    x = orc.synth_noise(20260000 + 100000, k * n, flt)
    stages = [rr.FreqShifter(123457.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(192, 48000.0, 6000.0)]
    ch = rr.Chain(ctx, stages, flt)
    got, plans = _push_all(ch, sr, x[None, :], n, [1, 2, 5, 16])
    ch.close()
    want = orc.Chain([orc.FreqShifter(flt, 1.0, 123457.0), orc.Filter.new(flt, orc.lowpass(3000.0)),
                      orc.Downsampler(flt, 192, 48000.0, 6000.0)]).run(sr, x, n)
    assert got.shape[1] == len(want) == (k - 1) * 192  # exactly 192 outputs per chunk after the start-up chunk
    assert orc.rel_l2(got[0], want) <= TOL[flt]


def test_c4_wideband_fm_full_shape(ctx):
    """256 stations are 256 rows of the same chain; 4 of them x 16 chunks here (the Downsampler's first whole
    4096-sample output chunk needs 14 chunks behind the two Filters' start-up chunks)."""
    import radiorust_b200 as rr

    sr, n, k, S = 10_000_000.0, 65536, 16, 4
    x = np.stack([orc.synth_fm_station(20260000 + 400000 + s, k * n, sr, 75000.0, 15000.0, -30.0, "f32") for s in range(S)])
    deemph = make_golden.deemph_resp(50e-6)
    stages = [rr.Filter.new(orc.lowpass(100000.0)), rr.FmDemod(75000.0), rr.Filter.new_rectangular(deemph),
              rr.Downsampler(4096, 48000.0, 40000.0)]
    ch = rr.Chain(ctx, stages, "f32", n_streams=S)
    got, plans = _push_all(ch, sr, x, n, [1, 3, 4, 8])
    ch.close()
    assert got.shape[1] == 4096, (got.shape, plans)
    for s in range(S):
        want = orc.Chain([orc.Filter.new("f32", orc.lowpass(100000.0)), orc.FmDemod("f32", 75000.0),
                          orc.Filter.new_rectangular("f32", deemph), orc.Downsampler("f32", 4096, 48000.0, 40000.0)]).run(sr, x[s], n)
        assert len(want) == 4096
        err = orc.rel_l2(got[s], want)
        assert err <= TOL["f32"], (s, err, plans)
    # the demodulated audio is a real signal (FmDemod writes im = 0, the de-emphasis response is Hermitian)
    assert np.linalg.norm(got.imag) <= 1e-4 * np.linalg.norm(got.real)


def test_c5_filter_two_to_the_twenty_f64(ctx):
    import radiorust_b200 as rr

    sr, n, k = 2_400_000.0, 1 << 19, 3
    x = orc.synth_noise(20260000 + 500000, k * n, "f64")
    ch = rr.Chain(ctx, [rr.Filter.new(orc.lowpass(20000.0))], "f64")
    got, plans = _push_all(ch, sr, x[None, :], n, [1, 2])
    ch.close()
    assert "big_os" in plans[-1] or "cluster" in plans[-1], plans
    want = orc.Chain([orc.Filter.new("f64", orc.lowpass(20000.0))]).run(sr, x, n)
    assert got.shape[1] == len(want) == 2 * n
    assert orc.rel_l2(got[0], want) <= TOL["f64"]


def test_c5_upsampler_f64(ctx):
    import radiorust_b200 as rr

    sr, n, k, S = 48000.0, 1024, 4, 2
    x = np.stack([orc.synth_noise(20260000 + 500001 + s, k * n, "f64") for s in range(S)])
    ch = rr.Chain(ctx, [rr.Upsampler(4096, 2_400_000.0, 20000.0)], "f64", n_streams=S)
    got, _ = _push_all(ch, sr, x, n, [1, 3])
    ch.close()
    for s in range(S):
        want = orc.Chain([orc.Upsampler("f64", 4096, 2_400_000.0, 20000.0)]).run(sr, x[s], n)
        assert got.shape[1] == len(want) == 50 * k * n
        assert orc.rel_l2(got[s], want) <= TOL["f64"]


def test_c2_single_stream_20msps(ctx):
    import radiorust_b200 as rr

    sr, n, k = 20_000_000.0, 65536, 8
    x = orc.synth_noise(20260000 + 200000, k * n, "f32")
    stages = [rr.FreqShifter(1_234_567.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(128, 48000.0, 6000.0)]
    ch = rr.Chain(ctx, stages, "f32")
    got, plans = _push_all(ch, sr, x[None, :], n, [1, 2, 5])
    ch.close()
    want = orc.Chain([orc.FreqShifter("f32", 1.0, 1_234_567.0), orc.Filter.new("f32", orc.lowpass(3000.0)),
                      orc.Downsampler("f32", 128, 48000.0, 6000.0)]).run(sr, x, n)
    assert got.shape[1] == len(want) > 0
    assert orc.rel_l2(got[0], want) <= TOL["f32"]

"""The two input restrictions of round 1, lifted: Filter at any chunk length (rustfft plans every length,
filters.rs:200,227-228) and resamplers at rates that are not integer valued (the reference carries `pos` in f64,
resampling.rs:67,109-111,197,247-266)."""
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


def _both(ctx, stages, ostages, flt, sr, x, n, pushes):
    import radiorust_b200 as rr

    ch = rr.Chain(ctx, stages, flt)
    parts, plans, pos = [], [], 0
    for k in pushes:
        y, _ = ch.push(sr, x[pos * n:(pos + k) * n], n)
        parts.append(y[0].copy())
        plans.append(ch.plan)
        pos += k
    ch.close()
    got = np.concatenate(parts)
    want = orc.Chain(ostages).run(sr, x[: sum(pushes) * n], n)
    assert got.shape == want.shape, (got.shape, want.shape, plans)
    assert want.size > 0
    err = orc.rel_l2(got, want)
    assert err <= TOL[flt], (err, plans)
    return plans


@pytest.mark.parametrize("flt,n", [("f32", 1000), ("f32", 4800), ("f32", 33), ("f32", 100), ("f32", 3), ("f64", 1000), ("f64", 4800), ("f32", 20000)])
def test_filter_any_chunk_length(ctx, flt, n):
    import radiorust_b200 as rr

    sr = 48000.0
    x = orc.synth_noise(700 + n, 9 * n, flt)

    def resp(b, f):
        return complex(1.0 / (1.0 + abs(f) / 3000.0), 0.05 if f > 0 else (-0.05 if f < 0 else 0.0))

    plans = _both(ctx, [rr.Filter.new(resp)], [orc.Filter.new(flt, resp)], flt, sr, x, n, [1, 1, 3, 4])
    assert any("padded_os" in p for p in plans), plans


def test_chain_with_chunk_length_1000(ctx):
    """FreqShifter -> Filter -> Downsampler at n = 1000 (no power of two anywhere), events included."""
    import radiorust_b200 as rr

    sr, n = 1_000_000.0, 1000
    x = orc.synth_noise(4242, 40 * n, "f32")
    stages = [rr.FreqShifter(123456.0), rr.Filter.new(orc.lowpass(20000.0)), rr.Downsampler(50, 50000.0, 40000.0)]
    ch = rr.Chain(ctx, stages, "f32")
    ob = [orc.FreqShifter("f32", 1.0, 123456.0), orc.Filter.new("f32", orc.lowpass(20000.0)), orc.Downsampler("f32", 50, 50000.0, 40000.0)]
    oc = orc.Chain(ob)
    got, want = [], []

    def feed(lo, hi):
        y, _ = ch.push(sr, x[lo * n:hi * n], n)
        got.append(y[0].copy())
        for k in range(lo, hi):
            for m in oc.push(orc.Samples(sr, x[k * n:(k + 1) * n])):
                if isinstance(m, orc.Samples):
                    want.append(m.chunk)

    feed(0, 3)
    feed(3, 4)
    feed(4, 20)
    ch.event(True)
    oc.push(orc.DISCONNECTION)
    feed(20, 30)
    ch.set_shift(0, -50000.0)
    ob[0].set_shift(-50000.0)
    feed(30, 40)
    ch.close()
    g, w = np.concatenate(got), np.concatenate(want)
    assert g.shape == w.shape
    assert orc.rel_l2(g, w) <= 1e-5


@pytest.mark.parametrize("flt", ["f32", "f64"])
def test_resamplers_at_rates_that_are_not_integers(ctx, flt):
    import radiorust_b200 as rr

    # Downsampler 1 000 000.5 S/s -> 48 kS/s
    sr, n = 1_000_000.5, 4096
    x = orc.synth_noise(31337, 12 * n, flt)
    plans = _both(ctx, [rr.Downsampler(64, 48000.0, 20000.0)], [orc.Downsampler(flt, 64, 48000.0, 20000.0)], flt, sr, x, n, [1, 2, 9])
    assert any("indexed" in p for p in plans), plans
    # behind a Filter (the fused kernels need a rational ratio: this takes the Filter's scratch output)
    plans = _both(ctx, [rr.Filter.new(orc.lowpass(15000.0)), rr.Downsampler(32, 44100.25, 20000.0)],
                  [orc.Filter.new(flt, orc.lowpass(15000.0)), orc.Downsampler(flt, 32, 44100.25, 20000.0)], flt, sr, x, n, [1, 2, 9])
    assert any("indexed" in p for p in plans), plans
    # Upsampler 48 000.25 S/s -> 2 400 000.5 S/s
    x = orc.synth_noise(271828, 6 * 512, flt)
    plans = _both(ctx, [rr.Upsampler(1000, 2_400_000.5, 20000.0)], [orc.Upsampler(flt, 1000, 2_400_000.5, 20000.0)], flt, 48000.25, x, 512,
                  [1, 2, 3])
    assert any("indexed" in p for p in plans), plans

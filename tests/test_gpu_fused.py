"""k_fused (front end + low-rate part in one persistent kernel, u in shared memory) against the oracle and against
the two-kernel path (k_front + k_poly2) it replaces in steady state."""
import os

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


def _run(ctx, stages, x, sr, n, pushes, shifts, env):
    import radiorust_b200 as rr

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ch = rr.Chain(ctx, stages, "f32", n_streams=x.shape[0])
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if shifts is not None:
        ch.set_shifts(0, shifts)
    parts, plans, pos = [], [], 0
    for k in pushes:
        y, _ = ch.push(sr, np.ascontiguousarray(x[:, pos * n:(pos + k) * n]), n)
        parts.append(y.copy())
        plans.append(ch.plan)
        pos += k
    ch.close()
    return np.concatenate(parts, axis=1), plans


@pytest.mark.parametrize("S,pushes", [(6, [3, 12, 12, 11, 14]), (3, [2, 25, 11]), (149, [3, 11, 11]),
                                      (5, [3, 1, 1, 1, 1, 2, 1, 1, 3, 1, 1, 1]),   # one chunk per message: fewer new rows than the filter reaches back
                                      (4, [2, 1, 8, 1, 8])])
def test_fused_matches_oracle_and_two_kernel_path(ctx, S, pushes):
    import radiorust_b200 as rr

    sr, n = 2_400_000.0, 4096
    total = sum(pushes)
    x = np.stack([orc.synth_noise(8800 + s, total * n, "f32") for s in range(S)])
    shifts = [float((s * 577) % 2_400_000 - 1_200_000) for s in range(S)]
    stages = [rr.FreqShifter(0.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)]
    got, plans = _run(ctx, stages, x, sr, n, pushes, shifts, {"RR_FUSED_MIN_STREAMS": "1"})
    assert any("fused[" in p for p in plans), plans
    ref, plans2 = _run(ctx, stages, x, sr, n, pushes, shifts, {"RR_DISABLE_FUSED": "1"})
    assert not any("fused[" in p for p in plans2)
    assert got.shape == ref.shape
    assert orc.rel_l2(got, ref) <= 2e-6
    for s in range(0, S, max(1, S // 4)):
        want = orc.Chain([orc.FreqShifter("f32", 1.0, shifts[s]), orc.Filter.new("f32", orc.lowpass(3000.0)),
                          orc.Downsampler("f32", 64, 48000.0, 6000.0)]).run(sr, x[s], n)
        assert want.shape == got[s].shape
        assert orc.rel_l2(got[s], want) <= 1e-5, s


def test_fused_without_nco_and_with_a_stage_behind(ctx):
    """No FreqShifter in front (HAS_NCO = false), and a GainControl behind the Downsampler (outputs go through the
    Downsampler's staging buffer instead of straight to the caller)."""
    import radiorust_b200 as rr

    sr, n, S = 2_400_000.0, 4096, 4
    pushes = [3, 13, 12]
    x = np.stack([orc.synth_noise(9900 + s, sum(pushes) * n, "f32") for s in range(S)])
    stages = [rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(100, 48000.0, 6000.0), rr.GainControl(0.5)]
    got, plans = _run(ctx, stages, x, sr, n, pushes, None, {"RR_FUSED_MIN_STREAMS": "1"})
    assert any("fused[" in p for p in plans), plans
    for s in range(S):
        want = orc.Chain([orc.Filter.new("f32", orc.lowpass(3000.0)), orc.Downsampler("f32", 100, 48000.0, 6000.0),
                          orc.GainControl("f32", 0.5)]).run(sr, x[s], n)
        assert want.shape == got[s].shape
        assert orc.rel_l2(got[s], want) <= 1e-5


def test_fused_then_event_then_short_pushes(ctx):
    """Path switches: fused pushes, an interrupt (Filter history dropped), short pushes (two-kernel path), fused again."""
    import radiorust_b200 as rr

    sr, n, S = 2_400_000.0, 4096, 3
    x = np.stack([orc.synth_noise(7700 + s, 60 * n, "f32") for s in range(S)])
    stages = [rr.FreqShifter(12345.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)]
    os.environ["RR_FUSED_MIN_STREAMS"] = "1"
    try:
        ch = rr.Chain(ctx, stages, "f32", n_streams=S)
    finally:
        del os.environ["RR_FUSED_MIN_STREAMS"]
    ocs = [orc.Chain([orc.FreqShifter("f32", 1.0, 12345.0), orc.Filter.new("f32", orc.lowpass(3000.0)),
                      orc.Downsampler("f32", 64, 48000.0, 6000.0)]) for _ in range(S)]
    got, want, plans = [], [[] for _ in range(S)], []

    def feed(lo, hi):
        y, _ = ch.push(sr, np.ascontiguousarray(x[:, lo * n:hi * n]), n)
        got.append(y.copy())
        plans.append(ch.plan)
        for s in range(S):
            for k in range(lo, hi):
                for m in ocs[s].push(orc.Samples(sr, x[s, k * n:(k + 1) * n])):
                    if isinstance(m, orc.Samples):
                        want[s].append(m.chunk)

    feed(0, 3)
    feed(3, 15)
    feed(15, 27)
    ch.event(True)
    for oc in ocs:
        oc.push(orc.DISCONNECTION)
    feed(27, 30)
    feed(30, 31)
    feed(31, 43)
    feed(43, 55)
    feed(55, 60)
    ch.close()
    assert sum("fused[" in p for p in plans) >= 2, plans
    g = np.concatenate(got, axis=1)
    for s in range(S):
        w = np.concatenate(want[s])
        assert w.shape == g[s].shape
        assert orc.rel_l2(g[s], w) <= 1e-5


@pytest.mark.parametrize("sr,n,out_rate,bw,ocl", [
    (960_000.0, 2048, 48000.0, 20000.0, 17),      # P = 20
    (1_920_000.0, 4096, 48000.0, 6000.0, 64),     # P = 40 (an even number of 16-byte units per half row: padded slots)
    (2_400_000.0, 8192, 48000.0, 6000.0, 64),     # P = 50, longer Filter: Lmax = 170
    (2_400_000.0, 2048, 48000.0, 20000.0, 100),   # P = 50, wider Downsampler (rank permitting)
    (3_072_000.0, 4096, 48000.0, 6000.0, 64),     # P = 64: no instantiation, k_front + k_poly2
    (4_800_000.0, 4096, 48000.0, 6000.0, 64),     # P = 100
])
def test_fused_other_decimation_factors_and_filter_lengths(ctx, sr, n, out_rate, bw, ocl):
    import radiorust_b200 as rr

    S = 3
    pushes = [3, 14, 13]
    x = np.stack([orc.synth_noise(6100 + s, sum(pushes) * n, "f32") for s in range(S)])
    shifts = [sr / 7.0, -sr / 5.0, 12345.0]
    stages = [rr.FreqShifter(0.0), rr.Filter.new(orc.lowpass(bw / 2)), rr.Downsampler(ocl, out_rate, bw)]
    got, plans = _run(ctx, stages, x, sr, n, pushes, shifts, {"RR_FUSED_MIN_STREAMS": "1"})
    for s in range(S):
        want = orc.Chain([orc.FreqShifter("f32", 1.0, shifts[s]), orc.Filter.new("f32", orc.lowpass(bw / 2)),
                          orc.Downsampler("f32", ocl, out_rate, bw)]).run(sr, x[s], n)
        assert want.shape == got[s].shape
        assert orc.rel_l2(got[s], want) <= 1e-5, (s, plans)
    # (a filter that does not factor to rank 10 keeps the polyphase kernels on all branches: still correct, reported here)
    print(sr, n, plans[-1])

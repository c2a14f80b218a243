"""Small pushes through every hand-synchronised kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python profiles/sanitizer_workload.py

k_chain_os (start-up chunks), k_front + k_poly2 (C3 shape), k_poly2 on all branches, k_poly (Q = 3, f32 and f64),
k_big_* (f32 N = 32768, f64 N = 8192), k_upsample, k_fmdemod, k_fmmod.  Results are checked against the oracle so a
sanitizer-clean run is also a correct one.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import radiorust_b200 as rr  # noqa: E402
from oracle import radiorust_oracle as orc  # noqa: E402


def run(ctx, stages, ostages, flt, sr, n, k, S, pushes, tol):
    x = np.stack([orc.synth_noise(31 + s, k * n, flt) for s in range(S)])
    ch = rr.Chain(ctx, stages, flt, n_streams=S)
    parts, pos, plans = [], 0, []
    for p in pushes:
        y, _ = ch.push(sr, np.ascontiguousarray(x[:, pos * n:(pos + p) * n]), n)
        parts.append(y.copy())
        plans.append(ch.plan)
        pos += p
    ch.close()
    got = np.concatenate(parts, axis=1)
    worst = 0.0
    for s in range(S):
        want = orc.Chain(ostages()).run(sr, x[s], n)
        assert want.shape == got[s].shape, (want.shape, got[s].shape)
        worst = max(worst, orc.rel_l2(got[s], want))
    assert worst <= tol, (worst, plans)
    print(f"ok {plans[-1]:40s} rel_l2 {worst:.2e}", flush=True)


def main():
    which = sys.argv[1:] or ["front", "poly2", "poly", "poly64", "big32", "big64", "misc"]
    ctx = rr.Context(0)
    lp = orc.lowpass
    if "front" in which:
        run(ctx, [rr.FreqShifter(-577000.0), rr.Filter.new(lp(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)],
            lambda: [orc.FreqShifter("f32", 1.0, -577000.0), orc.Filter.new("f32", lp(3000.0)), orc.Downsampler("f32", 64, 48000.0, 6000.0)],
            "f32", 2_400_000.0, 4096, 10, 3, [3, 1, 6], 1e-5)
    if "poly2" in which:
        os.environ["RR_DISABLE_FRONT"] = "1"
        run(ctx, [rr.FreqShifter(1000.0), rr.Filter.new(lp(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)],
            lambda: [orc.FreqShifter("f32", 1.0, 1000.0), orc.Filter.new("f32", lp(3000.0)), orc.Downsampler("f32", 64, 48000.0, 6000.0)],
            "f32", 2_400_000.0, 4096, 8, 2, [3, 5], 1e-5)
        del os.environ["RR_DISABLE_FRONT"]
    if "poly" in which:
        run(ctx, [rr.FreqShifter(123457.0), rr.Filter.new(lp(3000.0)), rr.Downsampler(192, 48000.0, 6000.0)],
            lambda: [orc.FreqShifter("f32", 1.0, 123457.0), orc.Filter.new("f32", lp(3000.0)), orc.Downsampler("f32", 192, 48000.0, 6000.0)],
            "f32", 1_024_000.0, 4096, 8, 2, [3, 5], 1e-5)
    if "poly64" in which:
        run(ctx, [rr.FreqShifter(123457.0), rr.Filter.new(lp(3000.0)), rr.Downsampler(96, 48000.0, 6000.0)],
            lambda: [orc.FreqShifter("f64", 1.0, 123457.0), orc.Filter.new("f64", lp(3000.0)), orc.Downsampler("f64", 96, 48000.0, 6000.0)],
            "f64", 1_024_000.0, 2048, 8, 1, [3, 5], 1e-12)
    if "big32" in which:
        run(ctx, [rr.Filter.new(lp(100000.0))], lambda: [orc.Filter.new("f32", lp(100000.0))], "f32", 10e6, 16384, 3, 2, [1, 2], 1e-5)
    if "big64" in which:
        run(ctx, [rr.Filter.new(lp(20000.0))], lambda: [orc.Filter.new("f64", lp(20000.0))], "f64", 2.4e6, 4096, 3, 1, [1, 2], 1e-12)
    if "misc" in which:
        run(ctx, [rr.Upsampler(512, 240000.0, 20000.0)], lambda: [orc.Upsampler("f32", 512, 240000.0, 20000.0)], "f32", 48000.0, 256, 4, 2,
            [1, 3], 1e-5)
        run(ctx, [rr.FmMod(5000.0), rr.FmDemod(5000.0)], lambda: [orc.FmMod("f32", 5000.0), orc.FmDemod("f32", 5000.0)], "f32", 48000.0,
            1000, 3, 2, [1, 2], 1e-4)
    ctx.close()
    print("sanitizer workload ok")


if __name__ == "__main__":
    main()

"""CUDA chain (through the C ABI) against the committed golden fixtures, plus edge cases."""
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


def rr_stages(case):
    import radiorust_b200 as rr

    out = []
    for b in case["blocks"]:
        k = b[0]
        if k == "freqshift":
            out.append(rr.FreqShifter(b[2], b[1]))
        elif k == "filter_lowpass":
            out.append(rr.Filter.new(orc.lowpass(b[1])))
        elif k == "filter_deemph":
            out.append(rr.Filter.new_rectangular(make_golden.deemph_resp(b[1])))
        elif k == "downsample":
            out.append(rr.Downsampler(b[1], b[2], b[3], b[4]))
        elif k == "upsample":
            out.append(rr.Upsampler(b[1], b[2], b[3], b[4]))
        elif k == "fmdemod":
            out.append(rr.FmDemod(b[1]))
    return out


@pytest.mark.parametrize("name", sorted(make_golden.cases()))
@pytest.mark.parametrize("split", ["one_push", "chunk_by_chunk"])
def test_against_golden(ctx, name, split):
    import radiorust_b200 as rr

    case = make_golden.cases()[name]
    want = np.load(os.path.join(HERE, "golden", name + ".npy"))
    n, k = case["chunk_len"], case["n_chunks"]
    x = make_golden.case_input(case)
    ch = rr.Chain(ctx, rr_stages(case), case["flt"])
    parts = []
    pushes = [k] if split == "one_push" else [1] * k
    pos = 0
    for c in pushes:
        y, _ = ch.push(case["sample_rate"], x[pos * n : (pos + c) * n], n)
        parts.append(y[0].copy())
        pos += c
    ch.close()
    got = np.concatenate(parts)
    assert got.shape == want.shape
    assert orc.rel_l2(got, want) <= TOL[case["flt"]]


def test_empty_and_ragged_pushes(ctx):
    import radiorust_b200 as rr

    ch = rr.Chain(ctx, [rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)], "f32")
    sr = 1_024_000.0
    # a single chunk only primes the filter: no output, no error
    y, _ = ch.push(sr, orc.synth_noise(1, 1024, "f32"), 1024)
    assert y.shape[1] == 0
    # chunk length change redesigns the filter and drops history (filters.rs:179-187)
    y, _ = ch.push(sr, orc.synth_noise(2, 4 * 2048, "f32"), 2048)
    oc = orc.Chain([orc.Filter.new("f32", orc.lowpass(3000.0)), orc.Downsampler("f32", 64, 48000.0, 6000.0)])
    oc.run(sr, orc.synth_noise(1, 1024, "f32"), 1024)
    want = oc.run(sr, orc.synth_noise(2, 4 * 2048, "f32"), 2048)
    assert y.shape[1] == len(want)
    assert orc.rel_l2(y[0], want) <= 1e-5
    ch.close()


def test_output_capacity_error_leaves_state_untouched(ctx):
    import ctypes as C

    import radiorust_b200 as rr
    from radiorust_b200 import _ffi

    ch = rr.Chain(ctx, [rr.GainControl(2.0)], "f32")
    x = orc.synth_noise(3, 256, "f32")
    out = np.zeros(16, dtype=np.complex64)
    cnt, rate = C.c_size_t(), C.c_double()
    rc = _ffi.load().rr_chain_push(ch._h, 48000.0, 256, 1, x.ctypes.data, 256, out.ctypes.data, 16, 16, C.byref(cnt), C.byref(rate))
    assert rc == _ffi.RR_ERR_CAPACITY
    y, _ = ch.push(48000.0, x, 256)
    assert np.array_equal(y[0], x * np.float32(2.0))
    ch.close()


def test_sample_rate_change_redesigns_everything(ctx):
    import radiorust_b200 as rr

    stages = [rr.FreqShifter(50000.0), rr.Filter.new(orc.lowpass(4000.0)), rr.Downsampler(32, 48000.0, 8000.0)]
    ch = rr.Chain(ctx, stages, "f32")
    oc = orc.Chain([orc.FreqShifter("f32", 1.0, 50000.0), orc.Filter.new("f32", orc.lowpass(4000.0)), orc.Downsampler("f32", 32, 48000.0, 8000.0)])
    n = 2048
    got, want = [], []
    for i, sr in enumerate([960_000.0, 960_000.0, 480_000.0, 480_000.0]):
        x = orc.synth_noise(50 + i, 6 * n, "f32")
        y, _ = ch.push(sr, x, n)
        got.append(y[0].copy())
        want.append(oc.run(sr, x, n))
    ch.close()
    g, w = np.concatenate(got), np.concatenate(want)
    assert g.shape == w.shape
    assert orc.rel_l2(g, w) <= 1e-5


def test_device_resident_push_matches_host_push(ctx):
    import torch

    import radiorust_b200 as rr

    sr, n, S, k = 2_400_000.0, 4096, 4, 12
    stages = [rr.FreqShifter(0.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)]
    x = np.stack([orc.synth_noise(700 + s, k * n, "f32") for s in range(S)])
    shifts = [1000.0 * s for s in range(S)]
    a = rr.Chain(ctx, stages, "f32", n_streams=S)
    a.set_shifts(0, shifts)
    ya, _ = a.push(sr, x, n)
    b = rr.Chain(ctx, stages, "f32", n_streams=S)
    b.set_shifts(0, shifts)
    xd = torch.from_numpy(x.view(np.float32).reshape(S, k * n, 2)).cuda()
    cap = b.max_output(sr, n, k)
    yd = torch.zeros((S, cap, 2), dtype=torch.float32, device="cuda")
    cnt, rate = b.push_device(sr, n, k, xd.data_ptr(), k * n, yd.data_ptr(), cap, cap)
    b.sync()
    yb = yd.cpu().numpy().reshape(S, cap * 2).view(np.complex64)[:, :cnt]
    assert cnt == ya.shape[1] and rate == 48000.0
    assert np.array_equal(ya, yb)
    a.close()
    b.close()

"""GPU tests of the Fourier analysis block (src/blocks/analysis.rs:15-132).

Unlike the filter chain this block has a reference test with known answers
(analysis.rs:140-209, `test_fourier`): it is replayed here through the C ABI on the GPU, so parity for
this block is pinned to the reference itself, not only to the oracle.
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


def gpu_fourier(ctx, x, n, flt, window=("rectangular",), center_dc=False):
    import radiorust_b200 as rr

    ch = rr.Chain(ctx, [rr.Fourier(window, center_dc)], flt, n_streams=np.atleast_2d(x).shape[0])
    y, rate = ch.push(48000.0, np.atleast_2d(x), n)
    assert rate == 48000.0 and "fourier" in ch.plan
    ch.close()
    return y


def test_reference_known_answers(ctx):
    """analysis.rs:140-209: [1,1,1] -> [3,0,0] (DC in the centre: [0,3,0]); [1,1.5,1,0.5] -> [4,-j,0,+j]
    (centre: [0,+j,4,-j]); f64, |error| <= 1e-10 as in assert_approx."""
    a = np.array([1.0, 1.0, 1.0], dtype=np.complex128)
    b = np.array([1.0, 1.5, 1.0, 0.5], dtype=np.complex128)
    for flt, tol in (("f64", 1e-10), ("f32", 1e-6)):
        xa, xb = a.astype(orc.complex_dtype(flt)), b.astype(orc.complex_dtype(flt))
        assert np.allclose(gpu_fourier(ctx, xa, 3, flt)[0], [3.0, 0.0, 0.0], atol=tol)
        assert np.allclose(gpu_fourier(ctx, xa, 3, flt, center_dc=True)[0], [0.0, 3.0, 0.0], atol=tol)
        assert np.allclose(gpu_fourier(ctx, xb, 4, flt)[0], [4.0, -1.0j, 0.0, 1.0j], atol=tol)
        assert np.allclose(gpu_fourier(ctx, xb, 4, flt, center_dc=True)[0], [0.0, 1.0j, 4.0, -1.0j], atol=tol)


@pytest.mark.parametrize("flt,n", [("f32", 64), ("f32", 1024), ("f32", 4096), ("f32", 16384), ("f64", 64), ("f64", 4096),
                                   ("f32", 5), ("f32", 100), ("f32", 1000), ("f64", 48), ("f64", 2000), ("f32", 32)])
@pytest.mark.parametrize("center_dc", [False, True])
def test_matches_oracle(ctx, flt, n, center_dc):
    S, chunks = 3, 4
    x = np.stack([orc.synth_noise(60 + s + n, chunks * n, flt) for s in range(S)])
    beta = orc.kaiser_null_at_bin_to_beta(3.0)
    got = gpu_fourier(ctx, x, n, flt, window=("kaiser", beta), center_dc=center_dc)
    blk = orc.Fourier(flt, orc.Kaiser(beta), center_dc)
    for s in range(S):
        want = np.concatenate([blk.process(orc.Samples(48000.0, x[s, c * n:(c + 1) * n]))[0].chunk for c in range(chunks)])
        assert orc.rel_l2(got[s], want) <= (1e-5 if flt == "f32" else 1e-12)


def test_parseval_and_custom_window(ctx):
    """Unit-mean-power window scale (analysis.rs:97): for white noise sum|Y|^2 ~= n * sum|x|^2."""
    n = 2048
    x = orc.synth_noise(7, 16 * n, "f32")
    y = gpu_fourier(ctx, x, n, "f32", window=lambda t: 1.0 - 0.5 * t * t)[0]
    ratio = float(np.sum(np.abs(y) ** 2) / (n * np.sum(np.abs(x) ** 2)))
    assert 0.9 < ratio < 1.1
    blk = orc.Fourier("f32", orc.CustomWindow(lambda t: 1.0 - 0.5 * t * t), False)
    want = np.concatenate([blk.process(orc.Samples(48000.0, x[c * n:(c + 1) * n]))[0].chunk for c in range(16)])
    assert orc.rel_l2(y, want) <= 1e-5


def test_in_a_chain_and_unsupported_length(ctx):
    """FreqShifter -> Fourier stays on the device between the blocks; lengths outside the plans above 4096 are refused."""
    import radiorust_b200 as rr

    n = 1024
    x = orc.synth_noise(11, 8 * n, "f32")
    ch = rr.Chain(ctx, [rr.FreqShifter(1000.0), rr.Fourier(center_dc=True)], "f32")
    y, _ = ch.push(48000.0, x, n)
    ch.close()
    oc = orc.Chain([orc.FreqShifter("f32", 1.0, 1000.0), orc.Fourier("f32", None, True)])
    want = oc.run(48000.0, x, n)
    assert orc.rel_l2(y[0], want) <= 1e-5
    ch = rr.Chain(ctx, [rr.Fourier()], "f32")
    with pytest.raises(rr.RadiorustError):
        ch.push(48000.0, orc.synth_noise(1, 2 * 5000, "f32"), 5000)
    ch.close()


def test_metering_level(ctx):
    """metering::level (metering.rs:21-30), incl. its doc-test vector: level([0, -0.5j, 1]) = 1.25/3."""
    import radiorust_b200 as rr

    v = np.array([0.0, -0.5j, 1.0], dtype=np.complex64)
    assert rr.level(ctx, v, 3)[0, 0] == pytest.approx(1.25 / 3.0, rel=1e-12)
    assert abs(rr.level(ctx, v, 3)[0, 0] - 0.41666667) < 0.001  # the reference's own assertion
    for dt, flt in ((np.complex64, "f32"), (np.complex128, "f64")):
        x = np.stack([orc.synth_noise(5 + s, 7 * 1000, flt) for s in range(3)])
        got = rr.level(ctx, x, 1000)
        assert got.shape == (3, 7)
        for s in range(3):
            for c in range(7):
                ch = x[s, c * 1000:(c + 1) * 1000]
                nsq = (ch.real * ch.real + ch.imag * ch.imag).astype(np.float64)  # norm_sqr in Flt, summed in f64
                assert got[s, c] == pytest.approx(float(np.sum(nsq)) / 1000.0, rel=1e-6 if flt == "f32" else 1e-12)

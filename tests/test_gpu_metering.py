"""GPU tests of metering::bandwidth / rescale_energy (src/metering.rs:32-110; SURVEY.md 8f, fed by the Fourier
block) through the C ABI.  The reference HAS tests for these (metering.rs:113-262): they are replayed on the
GPU, so parity is pinned to the reference's known answers, and random spectra are checked against the oracle."""
import numpy as np
import pytest

from oracle import radiorust_oracle as orc
from test_oracle_known_answers import METERING_BANDWIDTH_CASES, METERING_RESCALE_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import radiorust_b200 as rr

    c = rr.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("bins,want", METERING_BANDWIDTH_CASES)
@pytest.mark.parametrize("cdt", [np.complex128, np.complex64])
def test_bandwidth_reference_tests(ctx, bins, want, cdt):
    import radiorust_b200 as rr

    b = np.array(bins, dtype=cdt)
    got = rr.bandwidth(ctx, 0.01, 48000.0, b, len(b))
    assert got.shape == (1, 1)
    assert abs(got[0, 0] - want) <= (1e-10 if cdt == np.complex128 else 1e-6) * max(1.0, abs(want))  # assert_approx / f32 inputs


@pytest.mark.parametrize("inp,res,want", METERING_RESCALE_CASES)
def test_rescale_energy_reference_tests(ctx, inp, res, want):
    import radiorust_b200 as rr

    got = rr.rescale_energy(ctx, res, np.array(inp, dtype=np.complex128), len(inp))
    assert got.shape == (1, 1, res) and np.allclose(got[0, 0], want, rtol=0, atol=1e-10)


@pytest.mark.parametrize("flt,n", [("f32", 1), ("f32", 2), ("f32", 7), ("f32", 255), ("f32", 256), ("f32", 257), ("f32", 1000), ("f32", 4096),
                                   ("f64", 3), ("f64", 512), ("f64", 1537)])
def test_bandwidth_matches_oracle(ctx, flt, n):
    """Band-limited and white spectra, several chunks and streams per call; the running energy is a parallel f64
    prefix sum on the GPU and a sequential one in the reference: they agree to f64 rounding."""
    import radiorust_b200 as rr

    S, chunks, sr = 3, 5, 48000.0
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((S, chunks * n)) + 1j * rng.standard_normal((S, chunks * n))).astype(orc.complex_dtype(flt))
    k = np.arange(n)
    shape = np.exp(-0.5 * (np.minimum(k, n - k) / max(1.0, n / 16.0)) ** 2)  # energy around DC (index 0, wrapping)
    x[1] *= np.tile(shape, chunks)
    x[2, : 2 * n] = 0  # silent chunks
    for dp in (0.01, 0.5, 1.0, 0.0):
        got = rr.bandwidth(ctx, dp, sr, x, n)
        for s in range(S):
            for c in range(chunks):
                want = orc.bandwidth(dp, sr, x[s, c * n:(c + 1) * n])
                assert abs(got[s, c] - want) <= 1e-9 * sr, (dp, s, c, got[s, c], want)


@pytest.mark.parametrize("flt,n,res", [("f32", 1024, 1024), ("f32", 1024, 300), ("f32", 1000, 1920), ("f32", 5, 1), ("f32", 1, 7),
                                       ("f64", 4096, 800), ("f64", 333, 1000), ("f32", 16384, 777)])
def test_rescale_energy_matches_oracle_bit_for_bit(ctx, flt, n, res):
    import radiorust_b200 as rr

    S, chunks = 2, 3
    x = np.stack([orc.synth_noise(4000 + s + n, chunks * n, flt) for s in range(S)])
    got = rr.rescale_energy(ctx, res, x, n)
    assert got.shape == (S, chunks, res) and got.dtype == orc.real_dtype(flt)
    for s in range(S):
        for c in range(chunks):
            want = orc.rescale_energy(res, x[s, c * n:(c + 1) * n])
            assert np.array_equal(got[s, c], want)


def test_fourier_feeds_bandwidth(ctx):
    """The chain the reference documents (metering.rs:34-41): Fourier block -> bandwidth.  A tone pair 6 kHz apart
    measures ~6 kHz; white noise measures ~(1 - percentile) of the sample rate."""
    import radiorust_b200 as rr

    n, sr = 4096, 48000.0
    t = np.arange(n) / sr
    tones = (np.exp(2j * np.pi * 3000.0 * t) + np.exp(-2j * np.pi * 3000.0 * t)).astype(np.complex64)
    noise = orc.synth_noise(1, n, "f32")
    beta = orc.kaiser_null_at_bin_to_beta(4.0)
    ch = rr.Chain(ctx, [rr.Fourier(("kaiser", beta))], "f32", n_streams=2)
    spec, _ = ch.push(sr, np.stack([tones, noise]), n)
    ch.close()
    bw = rr.bandwidth(ctx, 0.01, sr, spec, n)
    assert abs(bw[0, 0] - 6000.0) < 100.0 and abs(bw[1, 0] - 0.99 * sr) < 0.01 * sr
    blk = orc.Fourier("f32", orc.Kaiser(beta))
    for s, sig in enumerate((tones, noise)):
        want = orc.bandwidth(0.01, sr, blk.process(orc.Samples(sr, sig))[0].chunk)
        assert abs(bw[s, 0] - want) <= 1e-3 * sr  # spectra agree to 1e-5 relative L2, the edge bins carry little energy


def test_argument_errors(ctx):
    import radiorust_b200 as rr

    with pytest.raises(rr.RadiorustError):
        rr.bandwidth(ctx, 0.01, 48000.0, np.zeros((1, 0), dtype=np.complex64), 0)
    with pytest.raises(TypeError):
        rr.bandwidth(ctx, 0.01, 48000.0, np.zeros(8), 8)

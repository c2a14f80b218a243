"""The oracle's C restatement (oracle/radiorust_oracle.c) against the numpy oracle.

Both restate the same reference loops; they differ only in the FFT stand-in for rustfft
(radix-2 here, pocketfft there) and libm's sinf/cosf vs numpy's, so f32 agrees to ~1e-6
and f64 to ~1e-13 relative L2.
"""
import numpy as np
import pytest

from oracle import oracle_c
from oracle import radiorust_oracle as orc


@pytest.mark.parametrize("flt,tol", [("f32", 3e-6), ("f64", 1e-12)])
def test_c_chain_matches_numpy_oracle(flt, tol):
    sr, n = 1_024_000.0, 1024
    S = 3
    x = np.stack([orc.synth_noise(100 + s, 10 * n, flt) for s in range(S)])
    shifts = [123457.0, -100000.0, 0.0]
    got = oracle_c.chain(x, flt, sr, n, shifts=shifts, freq_resp=orc.lowpass(8000.0), down=(48000.0, 12000.0, 3.0), n_threads=2)
    for s in range(S):
        oc = orc.Chain([orc.FreqShifter(flt, 1.0, shifts[s]), orc.Filter.new(flt, orc.lowpass(8000.0)),
                        orc.Downsampler(flt, 1, 48000.0, 12000.0)])
        want = oc.run(sr, x[s], n)
        assert len(got[s]) == len(want) > 0
        assert orc.rel_l2(got[s], want) <= tol


def test_c_single_blocks():
    sr, n = 48000.0, 256
    x = orc.synth_noise(7, 8 * n, "f32")
    got = oracle_c.chain(x, "f32", sr, n, shifts=[1234.0])[0]
    want = orc.Chain([orc.FreqShifter("f32", 1.0, 1234.0)]).run(sr, x, n)
    assert orc.rel_l2(got, want) <= 1e-6
    got = oracle_c.chain(x, "f32", sr, n, freq_resp=orc.lowpass(5000.0))[0]
    want = orc.Chain([orc.Filter.new("f32", orc.lowpass(5000.0))]).run(sr, x, n)
    assert len(got) == len(want) == 7 * n  # one-chunk start-up delay (filters.rs:79-81)
    assert orc.rel_l2(got, want) <= 3e-6
    got = oracle_c.chain(x, "f32", sr, n, down=(8000.0, 3000.0, 3.0))[0]
    want = orc.Chain([orc.Downsampler("f32", 1, 8000.0, 3000.0)]).run(sr, x, n)
    assert len(got) == len(want)
    assert orc.rel_l2(got, want) <= 1e-6

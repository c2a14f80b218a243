"""Pins the oracle against every known answer the reference's own tests hold
for this path: math.rs:57-85, analysis.rs:140-209,
transform.rs:397-416."""
import math

import numpy as np
import pytest

from oracle import radiorust_oracle as o


def assert_approx(a, b):
    """lib.rs:51-59: |a-b| <= 1e-10 or |ln(a/b)| <= 1e-10."""
    ok = abs(a - b) <= 1e-10
    if not ok and a != 0 and b != 0 and (a / b) > 0:
        ok = abs(math.log(a / b)) <= 1e-10
    assert ok, f"{a} and {b} are not approximately equal"


def test_bessel_I0_known_answers():  # math.rs:57-69
    assert o.bessel_I0(0.0) == 1.0
    assert o.bessel_I0(-math.inf) == math.inf
    assert o.bessel_I0(math.inf) == math.inf
    assert math.isnan(o.bessel_I0(math.nan))
    assert_approx(o.bessel_I0(0.5), 1.06348337074132)
    assert_approx(o.bessel_I0(-0.5), 1.06348337074132)
    assert_approx(o.bessel_I0(1.23), 1.41552757215846)
    assert_approx(o.bessel_I0(15.8), 736184.938479417)
    assert_approx(o.bessel_I0(456.0), 2.04094157812291e196)
    assert o.bessel_I0(1000.0) == math.inf
    assert o.bessel_I0(-1000.0) == math.inf


def test_bessel_I0_vec_matches_scalar():
    xs = np.array([0.0, 0.5, -0.5, 1.23, 15.8, 456.0, 1e-9, 3.3, 1000.0])
    v = o.bessel_I0_vec(xs)
    for x, y in zip(xs, v):
        assert y == o.bessel_I0(x)


def test_sinc_known_answers():  # math.rs:71-85
    assert o.sinc(0.0) == 1.0
    assert_approx(o.sinc(0.4), 0.756826728640657)
    assert_approx(o.sinc(-0.4), 0.756826728640657)
    for z in (1.0, -1.0, 2.0, 17.0, 2345.0, -2345.0):
        assert_approx(o.sinc(z), 0.0)
    assert_approx(o.sinc(2.6), 0.11643488132933186)
    assert_approx(o.sinc(-2.6), 0.11643488132933186)
    assert_approx(o.sinc(5.8), -0.03225825116512552)
    assert_approx(o.sinc(-5.8), -0.03225825116512552)
    v = o.sinc(np.array([0.0, 0.4, 2.6, 5.8]))
    assert v[0] == 1.0 and abs(v[1] - 0.756826728640657) < 1e-12


def test_kaiser_beta():  # math.rs:37-39, windowing.rs:46-50
    assert o.Kaiser.with_null_at_bin(2.0).beta == math.sqrt(3.0)
    assert o.kaiser_alpha_to_beta(2.0) == 2.0 * math.pi


def test_fourier_known_answers():  # analysis.rs:140-209
    f1 = o.Fourier("f64")
    f2 = o.Fourier("f64", center_dc=True)
    x = np.array([1, 1, 1], dtype=np.complex128)
    y1 = f1.process(o.Samples(48000.0, x))[0].chunk
    y2 = f2.process(o.Samples(48000.0, x))[0].chunk
    for got, want in zip(y1, [3, 0, 0]):
        assert_approx(got.real, want), assert_approx(got.imag, 0.0)
    for got, want in zip(y2, [0, 3, 0]):
        assert_approx(got.real, want), assert_approx(got.imag, 0.0)
    x = np.array([1, 1.5, 1, 0.5], dtype=np.complex128)
    y1 = f1.process(o.Samples(48000.0, x))[0].chunk
    y2 = f2.process(o.Samples(48000.0, x))[0].chunk
    for got, want in zip(y1, [4, -1j, 0, 1j]):
        assert_approx(got.real, complex(want).real), assert_approx(got.imag, complex(want).imag)
    for got, want in zip(y2, [0, 1j, 4, -1j]):
        assert_approx(got.real, complex(want).real), assert_approx(got.imag, complex(want).imag)


def test_gain_control_known_answers():  # transform.rs:397-416
    g = o.GainControl("f32", 0.25)
    x = np.array([32.0 - 1.0j, 15.0 - 2.0j], dtype=np.complex64)
    y = g.process(o.Samples(48000.0, x))[0].chunk
    assert y[0].real == 8.0 and y[0].imag == -0.25
    assert y[1].real == 3.75 and y[1].imag == -0.5


def test_deemphasis_factor():  # filters.rs:20-27
    z = o.deemphasis_factor(50e-6, 1000.0)
    want = 1.0 / complex(1.0, 50e-6 * 2 * math.pi * 1000.0)
    assert abs(z - want) < 1e-15
    assert o.deemphasis_factor(50e-6, 0.0) == 1.0


# ---------------------------------------------------------------------------------------------------
# metering.rs:113-262 -- the reference's own tests of level / bandwidth / rescale_energy
# ---------------------------------------------------------------------------------------------------
METERING_BANDWIDTH_CASES = [
    # (bins, expected hertz) at double_percentile 0.01, 48 kHz
    ([0j, 0j], 0.0),                                                                   # test_bandwidth_silence
    ([1, 1, 1, 1, 1, 1, -1, complex(math.sqrt(0.5), -math.sqrt(0.5))], 0.99 * 48000.0),  # test_bandwidth_spreadspectrum
    ([complex(7.4, -2.1)] * 3, 0.99 * 48000.0),                                        # test_bandwidth_spreadspectrum_odd
    ([0, 0, 0, 0, 0, 0, 2.1, 0], 0.99 * 48000.0 / 8.0),                                # test_bandwidth_carrier
    ([1.5, 0, 0, 0, 0, 0, 1.5, 0], 2.98 * 48000.0 / 8.0),                              # test_bandwidth_two_carriers
]
METERING_RESCALE_CASES = [
    ([0j, complex(2, 1), -0.5], 3, [0.0, 5.0, 0.25]),                                  # test_rescale_energy_same_size
    ([1, 2, 3, 4], 3, [2.3333333333333, 8.6666666666667, 19.0]),                       # test_rescale_energy_smaller
    ([1, 2, 3], 4, [0.75, 2.25, 4.25, 6.75]),                                          # test_rescale_energy_larger
]


def test_metering_level_reference_tests():
    h = 1.0 / math.sqrt(2.0)
    osc = np.array([1, complex(h, h), 1j, complex(-h, h), -1, complex(-h, -h), -1j, complex(h, -h)], dtype=np.complex128)
    assert abs(10.0 * math.log10(o.level(osc))) <= 1e-10                   # test_level_complex_osc
    assert abs(o.level(np.array([0, -0.5j, 1], dtype=np.complex128)) - 1.25 / 3.0) <= 1e-15  # doc test, metering.rs:9-19


@pytest.mark.parametrize("bins,want", METERING_BANDWIDTH_CASES)
def test_metering_bandwidth_reference_tests(bins, want):
    got = o.bandwidth(0.01, 48000.0, np.array(bins, dtype=np.complex128))
    assert abs(got - want) <= 1e-10 * max(1.0, abs(want))


@pytest.mark.parametrize("inp,res,want", METERING_RESCALE_CASES)
def test_metering_rescale_energy_reference_tests(inp, res, want):
    got = o.rescale_energy(res, np.array(inp, dtype=np.complex128))
    assert got.shape == (res,) and np.allclose(got, want, rtol=0, atol=1e-10)

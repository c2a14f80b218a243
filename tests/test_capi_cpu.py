"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares,
fails loudly without a device, and its host-side design math equals the oracle's."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from oracle import radiorust_oracle as orc
from radiorust_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_ffi.LIB_PATH):
        from radiorust_b200 import build

        build.build()
    return _ffi.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "radiorust_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", src)) - {"rr_freq_resp_fn", "rr_window_fn"})


def test_library_exports_every_header_symbol(lib):
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/radiorust_b200.h but not exported"
        assert n in _ffi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_ffi.SIGNATURES) == names


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.rr_ctx_create(0, C.byref(h))
    assert rc == _ffi.RR_ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.rr_last_error()


def test_design_math_known_answers(lib):
    """src/math.rs:57-85 through the C ABI."""
    for x, want in [(0.5, 1.06348337074132), (1.23, 1.41552757215846), (15.8, 736184.938479417), (456.0, 2.04094157812291e196)]:
        assert math.isclose(lib.rr_bessel_i0(x), want, rel_tol=1e-10)
    assert math.isinf(lib.rr_bessel_i0(1e6)) and math.isnan(lib.rr_bessel_i0(float("nan")))
    for x, want in [(0.4, 0.756826728640657), (2.6, 0.11643488132933186), (5.8, -0.03225825116512552), (0.0, 1.0)]:
        assert math.isclose(lib.rr_sinc(x), want, rel_tol=1e-10, abs_tol=1e-15)
    assert abs(lib.rr_sinc(3.0)) < 1e-15
    assert lib.rr_kaiser_null_at_bin_to_beta(2.0) == math.sqrt(3.0)  # no pi factor (math.rs:37-39)
    re_, im_ = C.c_double(), C.c_double()
    lib.rr_deemphasis_factor(50e-6, 1000.0, C.byref(re_), C.byref(im_))
    assert complex(re_.value, im_.value) == pytest.approx(orc.deemphasis_factor(50e-6, 1000.0), rel=1e-15)


@pytest.mark.parametrize("sr,prec,f", [(1_024_000.0, 1.0, 123457.0), (1_024_000.0, 1.0, 100000.0), (2.4e6, 1.0, -577.0),
                                       (48000.0, 0.5, 1234.0), (48000.0, 1.0, 0.0), (20e6, 1.0, 1_234_567.0)])
def test_freq_to_ratio(lib, sr, prec, f):
    n, d = C.c_int64(), C.c_int64()
    assert lib.rr_freq_to_ratio(sr, prec, f, C.byref(n), C.byref(d)) == 0
    assert (n.value, d.value) == orc.freq_to_ratio(sr, prec, f)


@pytest.mark.parametrize("n,flt", [(64, "f64"), (4096, "f64"), (4096, "f32"), (1024, "f32")])
def test_filter_design_matches_oracle(lib, n, flt):
    sr = 1_024_000.0
    resp = orc.lowpass(3000.0)
    cb = _ffi.FREQ_RESP_FN(lambda u, b, f, re, im: (re.__setitem__(0, resp(b, f).real), im.__setitem__(0, resp(b, f).imag)) and None)
    out = np.zeros(4 * n, dtype=np.float64)
    rc = lib.rr_design_filter_response(cb, None, _ffi.RR_WINDOW_KAISER, math.sqrt(3.0), _ffi.WINDOW_FN(), None, sr, n,
                                       _ffi.RR_C32 if flt == "f32" else _ffi.RR_C64, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    got = out[0::2] + 1j * out[1::2]
    want = orc.design_filter_response(resp, orc.Kaiser.with_null_at_bin(2.0), sr, n, flt).astype(np.complex128)
    # f64: both are f64 designs; f32: the reference runs the last FFT in f32, the library in f64
    assert orc.rel_l2(got, want) <= (1e-12 if flt == "f64" else 5e-7)


@pytest.mark.parametrize("down,in_rate,out_rate,bw,q,L", [(True, 1_024_000.0, 48000.0, 6000.0, 3.0, 147), (True, 2.4e6, 48000.0, 6000.0, 3.0, 343),
                                                        (True, 10e6, 48000.0, 40000.0, 3.0, 7500), (True, 20e6, 48000.0, 6000.0, 3.0, 2858),
                                                        (False, 48000.0, 2.4e6, 20000.0, 3.0, 515)])
def test_resampler_taps_match_oracle(lib, down, in_rate, out_rate, bw, q, L):
    fn = lib.rr_design_downsampler_taps if down else lib.rr_design_upsampler_taps
    n = C.c_size_t(0)
    assert fn(in_rate, out_rate, bw, q, C.byref(n), None) == 0
    assert n.value == L
    buf = np.zeros(L, dtype=np.float64)
    n = C.c_size_t(L)
    assert fn(in_rate, out_rate, bw, q, C.byref(n), buf.ctypes.data_as(C.POINTER(C.c_double))) == 0
    blk = (orc.Downsampler if down else orc.Upsampler)("f64", 1, out_rate, bw, q)
    blk._design(in_rate)
    assert np.allclose(buf, blk.ir, rtol=1e-12, atol=1e-18)
    assert math.isclose(float(np.sum(buf * buf)), 1.0, rel_tol=1e-12)  # unit energy (resampling.rs:96-98)


def test_design_argument_errors(lib):
    n = C.c_size_t(0)
    assert lib.rr_design_downsampler_taps(8000.0, 48000.0, 6000.0, 3.0, C.byref(n), None) == _ffi.RR_ERR_INVALID
    assert lib.rr_design_downsampler_taps(96000.0, 48000.0, 60000.0, 3.0, C.byref(n), None) == _ffi.RR_ERR_INVALID
    d = C.c_int64()
    assert lib.rr_freq_to_ratio(1.0, 10.0, 0.1, C.byref(d), C.byref(d)) == _ffi.RR_ERR_INVALID  # denominator rounds to 0
    assert lib.rr_last_error()


# ---------------------------------------------------------------------------
# rank-reduced form of Filter -> Downsampler (k_front + k_poly2): the factorisation of the fused filter
# must reproduce the full polyphase tables to below f32 rounding, with the rank the kernels are built for
# ---------------------------------------------------------------------------
def _fused_rank(lib, resp, sr, n, out_rate, bw, tol, max_rank, window=("kaiser", math.sqrt(3.0))):
    cb = _ffi.FREQ_RESP_FN(lambda u, b, f, re, im: (re.__setitem__(0, complex(resp(b, f)).real), im.__setitem__(0, complex(resp(b, f)).imag)) and None)
    rank, disc, err = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
    kind = _ffi.RR_WINDOW_KAISER if window[0] == "kaiser" else _ffi.RR_WINDOW_RECTANGULAR
    rc = lib.rr_design_fused_rank(cb, None, kind, window[1], _ffi.WINDOW_FN(), None, sr, n, out_rate, bw, 3.0, tol, max_rank,
                                  C.byref(rank), C.byref(disc), C.byref(err))
    return rc, rank.value, disc.value, err.value


@pytest.mark.parametrize("sr,n,cut,out_rate,bw", [
    (2_400_000.0, 4096, 3000.0, 48000.0, 6000.0),     # C3 (the benchmarked configuration): P = 50
    (1_440_000.0, 4096, 3000.0, 48000.0, 6000.0),     # P = 30
    (480_000.0, 1024, 8000.0, 48000.0, 20000.0),      # P = 10
    (3_360_000.0, 4096, 3000.0, 48000.0, 6000.0),     # P = 70
    (960_000.0, 2048, 10000.0, 48000.0, 20000.0),     # P = 20
    (3_072_000.0, 4096, 3000.0, 48000.0, 6000.0),     # P = 64
    (4_800_000.0, 4096, 3000.0, 48000.0, 6000.0),     # P = 100
])
def test_fused_filter_rank_is_low_and_exact(lib, sr, n, cut, out_rate, bw):
    rc, rank, disc, err = _fused_rank(lib, orc.lowpass(cut), sr, n, out_rate, bw, 2.0e-8, 10)
    assert rc == 0
    assert 1 <= rank <= 10, rank
    assert disc <= 2.0e-8
    assert 0.0 <= err <= 2.5e-8, err  # sum_c a_c b_c^T == the full [P][K] tables to below f32 rounding


def test_fused_filter_rank_numpy_cross_check(lib):
    """The same rank out of numpy's SVD of M[p][l] = g[P-1-p+l*P] built from the oracle's designs."""
    sr, n, P = 2_400_000.0, 4096, 50
    H = orc.design_filter_response(orc.lowpass(3000.0), orc.Kaiser.with_null_at_bin(2.0), sr, n, "f32").astype(np.complex128)
    h = np.fft.ifft(H)[n:]
    ds = orc.Downsampler("f32", 1, 48000.0, 6000.0, 3.0)
    ds._design(sr)
    g = np.convolve(h, np.asarray(ds.ir, dtype=np.float64)[::-1])
    lmax = (len(g) - 1) // P
    M = np.zeros((P, lmax + 1), dtype=np.complex128)
    for p in range(P):
        idx = P - 1 - p + P * np.arange(lmax + 1)
        ok = idx < len(g)
        M[p, ok] = g[idx[ok]]
    sv = np.linalg.svd(M, compute_uv=False)
    tail = np.sqrt(np.cumsum((sv ** 2)[::-1])[::-1] / np.sum(sv ** 2))
    want = int(np.argmax(tail <= 2.0e-8))
    rc, rank, disc, err = _fused_rank(lib, orc.lowpass(3000.0), sr, n, 48000.0, 6000.0, 2.0e-8, 16)
    assert rc == 0 and rank == want, (rank, want)
    assert disc == pytest.approx(float(tail[want]), rel=0.05)


def test_fused_filter_rank_gives_up_on_wideband_filters(lib):
    """An all-pass Filter leaves only the Downsampler's taps: more branches stay independent and the front
    end must not be chosen with too small a rank (rank = 0 -> the chain falls back to k_poly2 / k_poly)."""
    rc, rank, disc, err = _fused_rank(lib, lambda b, f: 1.0, 2_400_000.0, 4096, 48000.0, 6000.0, 2.0e-8, 4)
    assert rc == 0 and rank == 0
    rc, rank, disc, err = _fused_rank(lib, lambda b, f: 1.0, 2_400_000.0, 4096, 240000.0, 100000.0, 2.0e-8, 64)
    assert rc == 0 and (rank == 0 or err <= 2.5e-8)
    # contract violations
    assert _fused_rank(lib, orc.lowpass(3000.0), 1_000_000.0, 4096, 48000.0, 6000.0, 2.0e-8, 10)[0] == _ffi.RR_ERR_UNSUPPORTED  # 125/6


@pytest.mark.parametrize("sr,n,cut,out_rate,bw", [
    (1_024_000.0, 4096, 3000.0, 48000.0, 6000.0),     # C1: in/out = 64/3
    (1_200_000.0, 4096, 3000.0, 32000.0, 6000.0),     # 75/2
    (1_120_000.0, 4096, 3000.0, 64000.0, 6000.0),     # 35/2
])
def test_fused_filter_rank_with_output_phases(lib, sr, n, cut, out_rate, bw):
    """in/out = P/Q, Q > 1: the Q phase matrices factored together (one front end for all phases) reproduce the full
    [Q][P][K] polyphase tables to below f32 rounding within the sixteen columns the front end is built for."""
    rc, rank, disc, err = _fused_rank(lib, orc.lowpass(cut), sr, n, out_rate, bw, 2.0e-8, 16)
    assert rc == 0
    assert 1 <= rank <= 16, rank
    assert disc <= 2.0e-8
    assert 0.0 <= err <= 2.5e-8, err
    assert _fused_rank(lib, orc.lowpass(3000.0), 2_400_000.0, 4000, 48000.0, 6000.0, 2.0e-8, 10)[0] == _ffi.RR_ERR_INVALID


@pytest.mark.parametrize("n", [3, 33, 1000, 4800])
def test_filter_design_at_any_chunk_length(lib, n):
    """filters.rs:200,227-228 plans every length: the host design (Bluestein on the power-of-two transform) against the oracle."""
    f = orc.lowpass(5000.0)

    def cb(_u, b, fr, re, im):
        v = f(int(b), float(fr))
        re[0], im[0] = v.real, v.imag

    out = np.zeros(4 * n)
    rc = lib.rr_design_filter_response(_ffi.FREQ_RESP_FN(cb), None, _ffi.RR_WINDOW_KAISER, math.sqrt(3.0), _ffi.WINDOW_FN(), None, 48000.0, n,
                                       _ffi.RR_C64, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    want = orc.design_filter_response(f, orc.Kaiser.with_null_at_bin(2.0), 48000.0, n, "f64")
    assert np.linalg.norm(out.view(np.complex128) - want) <= 1e-13 * np.linalg.norm(want)


def test_fused_filter_rank_many_branches(lib):
    """BASELINE config 2 (in/out = 1250/3, n = 65536): with P in the thousands the factorisation first takes the row space
    by pivoted Gram-Schmidt (seconds instead of minutes) and must still reproduce the full tables to below f32 rounding."""
    import time

    t0 = time.time()
    rc, rank, disc, err = _fused_rank(lib, orc.lowpass(3000.0), 20_000_000.0, 65536, 48000.0, 6000.0, 2.0e-8, 32)
    assert rc == 0
    assert 17 <= rank <= 32, rank
    assert disc <= 2.0e-8
    assert 0.0 <= err <= 2.5e-8, err
    assert time.time() - t0 < 60.0

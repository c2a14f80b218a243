"""CPU oracle: a literal numpy restatement of radiorust's IQ sample-chain blocks.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``radiorust_b200``) never
imports, links or executes anything under ``oracle/`` and fails loudly when its
CUDA library is missing.

Parity status
-------------
* Design math (``bessel_I0``, ``sinc``, Kaiser beta) is PINNED by the
  reference's own known answers in ``src/math.rs:57-85`` (see
  ``tests/test_oracle_known_answers.py``).
* The FFT convention (unnormalised forward/inverse, sign) is PINNED by
  ``src/blocks/analysis.rs:140-209`` (``test_fourier``).
* ``GainControl`` products are pinned by ``src/blocks/transform.rs:397-416``.
* End-to-end outputs of ``FreqShifter``, ``Filter``, ``Downsampler``,
  ``Upsampler`` and ``FmDemod``: **parity unpinned** -- the reference ships no
  tests or golden vectors for them (``filters.rs:378``, ``resampling.rs:282``,
  ``modulation.rs:160`` are empty test modules) and the reference itself (Rust,
  needs cargo + the un-vendored crates rustfft 6.x / num 0.4 / tokio 1.21)
  cannot be built in this image.  For those blocks this file follows the
  reference loops line by line (citations on every function) and
  ``scipy.fft`` (pocketfft, native c64 / c128) stands in for rustfft: both are
  plain unnormalised DFTs, so they agree to rounding (~1e-7 f32, ~1e-16 f64).

Every class keeps the streaming state of the reference task (history chunk,
ring buffer, phase index, previous sample) so multi-chunk behaviour -- the
one-chunk start-up delay of ``Filter``, retunes, interrupt events -- is
reproduced, not just the steady state.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import scipy.fft

TAU = 2.0 * math.pi

# --------------------------------------------------------------------------
# numbers.rs:23-61 -- the two precisions of the path
# --------------------------------------------------------------------------


def real_dtype(flt: str):
    """``Flt`` in {f32, f64} (numbers.rs:23-42)."""
    return {"f32": np.float32, "f64": np.float64}[flt]


def complex_dtype(flt: str):
    return {"f32": np.complex64, "f64": np.complex128}[flt]


# --------------------------------------------------------------------------
# math.rs:7-49
# --------------------------------------------------------------------------


def bessel_I0(x: float) -> float:
    """Modified Bessel function I0, series summed until the f64 sum stops
    changing (math.rs:7-20)."""
    x = float(x)
    base = x * x / 4.0
    addend = 1.0
    s = 1.0
    i = 1
    while True:
        addend *= base / float(i * i)
        old = s
        s += addend
        if s == old or not math.isfinite(s):
            break
        i += 1
    return s


def bessel_I0_vec(x: np.ndarray) -> np.ndarray:
    """Element-wise ``bessel_I0`` with the same per-element stopping rule."""
    x = np.asarray(x, dtype=np.float64)
    base = x * x / 4.0
    addend = np.ones_like(x)
    s = np.ones_like(x)
    active = np.ones(x.shape, dtype=bool)
    i = 1
    with np.errstate(over="ignore", invalid="ignore"):
        while active.any():
            addend = np.where(active, addend * (base / float(i * i)), addend)
            old = s
            s = np.where(active, s + addend, s)
            active &= ~((s == old) | ~np.isfinite(s))
            i += 1
    return s


def kaiser_rel_with_beta(beta: float, x):
    """math.rs:26-28."""
    if np.ndim(x) == 0:
        return bessel_I0(beta * math.sqrt(1.0 - x * x))
    x = np.asarray(x, dtype=np.float64)
    return bessel_I0_vec(beta * np.sqrt(1.0 - x * x))


def kaiser_alpha_to_beta(alpha: float) -> float:
    """math.rs:31-33."""
    return alpha * math.pi


def kaiser_null_at_bin_to_beta(n: float) -> float:
    """math.rs:37-39 -- note: sqrt(n^2 - 1), *no* pi factor (replicated, not
    'fixed')."""
    return math.sqrt(n * n - 1.0)


def sinc(x):
    """Normalised sinc (math.rs:42-49)."""
    if np.ndim(x) == 0:
        if x == 0:
            return 1.0
        t = x * math.pi
        return math.sin(t) / t
    x = np.asarray(x, dtype=np.float64)
    t = x * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        y = np.sin(t) / t
    return np.where(x == 0, 1.0, y)


# --------------------------------------------------------------------------
# windowing.rs:6-67
# --------------------------------------------------------------------------


class Window:
    def relative_value_at(self, x):  # windowing.rs:9
        raise NotImplementedError


class Rectangular(Window):
    """windowing.rs:14-20."""

    def relative_value_at(self, x):
        return 1.0 if np.ndim(x) == 0 else np.ones(np.shape(x))


class Kaiser(Window):
    """windowing.rs:24-51."""

    def __init__(self, beta: float):
        self.beta = float(beta)

    @classmethod
    def with_beta(cls, beta):
        return cls(beta)

    @classmethod
    def with_alpha(cls, alpha):
        return cls(kaiser_alpha_to_beta(alpha))

    @classmethod
    def with_null_at_bin(cls, n):
        return cls(kaiser_null_at_bin_to_beta(n))

    def relative_value_at(self, x):
        return kaiser_rel_with_beta(self.beta, x)


class CustomWindow(Window):
    """windowing.rs:58-67."""

    def __init__(self, f: Callable[[float], float]):
        self.f = f

    def relative_value_at(self, x):
        if np.ndim(x) == 0:
            return self.f(x)
        return np.array([self.f(float(v)) for v in np.ravel(x)]).reshape(np.shape(x))


# --------------------------------------------------------------------------
# signal.rs:19-46,170-183 -- messages
# --------------------------------------------------------------------------


@dataclass
class Event:
    """signal.rs:19-31.  ``Disconnection`` (signal.rs:37-46) is an interrupt."""

    name: str = "event"
    interrupt: bool = False

    def is_interrupt(self) -> bool:
        return self.interrupt


DISCONNECTION = Event("Disconnection", True)


@dataclass
class Samples:
    """``Signal::Samples`` (signal.rs:172-179)."""

    sample_rate: float
    chunk: np.ndarray


# --------------------------------------------------------------------------
# FFT stand-in for rustfft (unnormalised both ways; filters.rs:200,227-252)
# --------------------------------------------------------------------------


def fft_forward(x: np.ndarray) -> np.ndarray:
    return scipy.fft.fft(x, norm="backward")


def fft_inverse_unnormalised(x: np.ndarray) -> np.ndarray:
    return scipy.fft.ifft(x, norm="forward")


# --------------------------------------------------------------------------
# filters.rs:20-27
# --------------------------------------------------------------------------


def deemphasis_factor(tau: float, frequency: float) -> complex:
    """``1 / (1 + j*tau*2*pi*f)`` (filters.rs:20-27; ``finv`` of num)."""
    z = complex(1.0, tau * TAU * frequency)
    n = z.real * z.real + z.imag * z.imag  # Complex::finv = conj / norm_sqr
    return complex(z.real / n, -z.imag / n)


# --------------------------------------------------------------------------
# transform.rs:297-362 -- FreqShifter
# --------------------------------------------------------------------------


def _rust_round(x: float) -> int:
    """f64::round -- half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def freq_to_ratio(sample_rate: float, precision: float, frequency: float):
    """transform.rs:298-302 (+ ``Ratio::new`` gcd reduction, denom > 0)."""
    denom = _rust_round(sample_rate / precision)
    numer = _rust_round(float(denom) * frequency / sample_rate)
    if denom == 0:
        raise ZeroDivisionError("denominator == 0")  # Ratio::new panics
    g = math.gcd(numer, denom)
    numer //= g
    denom //= g
    if denom < 0:
        numer, denom = -numer, -denom
    return numer, denom


class FreqShifter:
    """NCO + mixer (transform.rs:266-391).

    The reference materialises ``phase_vec`` (up to ``sample_rate/precision``
    entries).  Each entry is a pure function of its index, so this oracle
    evaluates the entries it needs on the fly with the same ``Flt`` arithmetic:
    ``phi_k = start_phase + Flt(i_k)/Flt(denom)*TAU``, ``i_0 = 0``,
    ``i_{k+1} = (i_k + numer) % denom`` (Rust ``%``: sign of the dividend).
    """

    def __init__(self, flt: str = "f32", precision: float = 1.0, shift: float = 0.0):
        self.flt = flt
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self.precision = float(precision)
        self._shift = float(shift)
        self._shift_changed = False
        self.prev_sample_rate: Optional[float] = None
        self.have_table = False
        self.numer = 0
        self.denom = 1
        self.start_phase = self.R(0)
        self.phase_idx = 0

    # transform.rs:376-390
    def shift(self) -> float:
        return self._shift

    def set_shift(self, shift: float):
        self._shift = float(shift)
        self._shift_changed = True

    def _i_of_k(self, k: np.ndarray) -> np.ndarray:
        a = abs(self.numer)
        sgn = -1 if self.numer < 0 else 1
        # (|numer|*k) mod denom without int64 overflow: both < 2^31 in practice
        return sgn * ((a * k.astype(np.int64)) % self.denom)

    def _phase_entries(self, k: np.ndarray) -> np.ndarray:
        R = self.R
        i = self._i_of_k(k)
        with np.errstate(over="ignore"):
            ph = self.start_phase + i.astype(R) / R(self.denom) * R(TAU)  # transform.rs:335
        ph = ph.astype(R)
        out = np.empty(k.shape, dtype=self.C)
        out.real = np.cos(ph)
        out.imag = np.sin(ph)
        return out

    def process(self, sig):
        if isinstance(sig, Event):
            return [sig]  # transform.rs:357-359
        sr, x = sig.sample_rate, np.asarray(sig.chunk, dtype=self.C)
        recalc = self._shift_changed or self.prev_sample_rate != sr  # :318-319
        self.prev_sample_rate = sr
        if recalc:
            if self.have_table:  # :322-325  phase_vec[phase_idx].arg()
                e = self._phase_entries(np.array([self.phase_idx]))[0]
                self.start_phase = self.R(np.arctan2(self.R(e.imag), self.R(e.real)))
            else:
                self.start_phase = self.R(0)
            self.phase_idx = 0
            self._shift_changed = False
            self.numer, self.denom = freq_to_ratio(sr, self.precision, self._shift)
            self.have_table = True
        n = len(x)
        k = (self.phase_idx + np.arange(n, dtype=np.int64)) % self.denom
        y = (x * self._phase_entries(k)).astype(self.C)  # :341-348
        self.phase_idx = int((self.phase_idx + n) % self.denom)
        return [Samples(sr, y)]


# --------------------------------------------------------------------------
# transform.rs:29-92 -- GainControl (a "next" row; trivially exact)
# --------------------------------------------------------------------------


class GainControl:
    def __init__(self, flt: str = "f32", gain: float = 1.0):
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self.gain = float(gain)

    def process(self, sig):
        if isinstance(sig, Event):
            return [sig]
        x = np.asarray(sig.chunk, dtype=self.C)
        return [Samples(sig.sample_rate, (x * self.R(self.gain)).astype(self.C))]


# --------------------------------------------------------------------------
# filters.rs:153-277 -- Filter
# --------------------------------------------------------------------------


def design_filter_response(
    freq_resp: Callable[[int, float], complex],
    window: Window,
    sample_rate: float,
    n: int,
    flt: str,
) -> np.ndarray:
    """``extended_response`` (2n bins, ``Flt``) -- filters.rs:184-238."""
    C = complex_dtype(flt)
    n_flt = float(n)
    scale = 2.0 * n_flt * n_flt  # :186
    response = np.zeros(n, dtype=np.complex128)
    freq_step = sample_rate / n_flt
    max_bin_abs = (n - 1) // 2
    for i in range(max_bin_abs + 1):  # :193-199
        freq = float(i) * freq_step
        response[i] = complex(freq_resp(i, freq)) / scale
        if i > 0:
            response[n - i] = complex(freq_resp(-i, -freq)) / scale
    response = fft_inverse_unnormalised(response)  # :200
    half = n // 2
    for_swap = response.copy()  # :201-203
    response[:half] = for_swap[half : 2 * half]
    response[half : 2 * half] = for_swap[:half]
    idx = np.arange(n, dtype=np.float64)
    w = np.asarray(window.relative_value_at(2.0 * (idx + 0.5) / n_flt - 1.0), dtype=np.float64)
    norm_pre = response.real * response.real + response.imag * response.imag
    energy_pre = float(np.cumsum(norm_pre)[-1]) if n else 0.0  # sequential sum, :204-214
    response = response * w
    norm_post = response.real * response.real + response.imag * response.imag
    energy_post = float(np.cumsum(norm_post)[-1]) if n else 0.0
    scale2 = math.sqrt(energy_pre / energy_post)  # :216
    response = response * scale2
    ext = np.zeros(2 * n, dtype=C)  # :220-226
    ext[n:] = response.astype(C)
    return fft_forward(ext).astype(C)  # :227-238 (Flt FFT)


class Filter:
    """Fast-convolution FIR with one chunk of history (filters.rs:110-298)."""

    def __init__(self, flt: str, freq_resp, window: Optional[Window] = None):
        self.flt = flt
        self.C = complex_dtype(flt)
        self.freq_resp = freq_resp
        self.window = window if window is not None else Kaiser.with_null_at_bin(2.0)  # :132
        self._params_changed = False
        self.prev_sample_rate = None
        self.prev_n = None
        self.previous_chunk = None
        self.extended_response = None

    @classmethod
    def new(cls, flt, freq_resp):
        return cls(flt, freq_resp)

    @classmethod
    def new_rectangular(cls, flt, freq_resp):  # :138-143
        return cls(flt, freq_resp, Rectangular())

    @classmethod
    def with_window(cls, flt, freq_resp, window):
        return cls(flt, freq_resp, window)

    def update(self, freq_resp):  # :279-286
        self.freq_resp = freq_resp
        self._params_changed = True

    def update_with_window(self, freq_resp, window):  # :288-297
        self.freq_resp = freq_resp
        self.window = window
        self._params_changed = True

    def process(self, sig):
        if isinstance(sig, Event):
            if sig.is_interrupt():  # :262-265
                self.previous_chunk = None
            return [sig]
        sr = sig.sample_rate
        x = np.asarray(sig.chunk, dtype=self.C)
        n = len(x)
        recalc = self._params_changed or sr != self.prev_sample_rate or n != self.prev_n  # :179-181
        self.prev_sample_rate, self.prev_n = sr, n
        if recalc:
            self._params_changed = False
            self.previous_chunk = None  # :187
            self.extended_response = design_filter_response(self.freq_resp, self.window, sr, n, self.flt)
        out = []
        if self.previous_chunk is not None:  # :240-258
            buf = np.concatenate([self.previous_chunk, x]).astype(self.C)
            buf = fft_forward(buf).astype(self.C)
            buf = (buf * self.extended_response).astype(self.C)
            buf = fft_inverse_unnormalised(buf).astype(self.C)
            out.append(Samples(sr, buf[:n].copy()))
        self.previous_chunk = x  # :260
        return out


# --------------------------------------------------------------------------
# resampling.rs:45-145 -- Downsampler
# --------------------------------------------------------------------------


def design_resampler_taps(ir_len: int, ratio: float, null_bin: float, flt: str) -> np.ndarray:
    """Shared by Down/Upsampler: ``ir[i] = sinc(x_i*ratio) * Kaiser(x_i*2/L)``
    with unit energy, cast to ``Flt`` (resampling.rs:84-98, :219-233)."""
    R = real_dtype(flt)
    ir_len_flt = float(ir_len)
    window = Kaiser.with_null_at_bin(null_bin)
    i = np.arange(ir_len, dtype=np.float64)
    x = (i + 0.5) - ir_len_flt / 2.0
    y = sinc(x * ratio) * window.relative_value_at(x * 2.0 / ir_len_flt)
    energy = float(np.cumsum(y * y)[-1])
    scale = 1.0 / math.sqrt(energy)
    return (y * scale).astype(R)


class Downsampler:
    """resampling.rs:14-146.  Streaming state: ``ir``, ring (kept here as the
    last ``L`` input samples in time order), ``pos`` (f64) and the partially
    filled output chunk."""

    def __init__(self, flt: str, output_chunk_len: int, output_rate: float, bandwidth: float, quality: float = 3.0):
        assert output_rate >= 0.0, "output sample rate must be positive"  # :51-56
        assert bandwidth >= 0.0, "bandwidth must be positive"
        assert bandwidth < output_rate, "bandwidth must be smaller than output sample rate"
        self.flt = flt
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self.output_chunk_len = int(output_chunk_len)
        self.output_rate = float(output_rate)
        self.bandwidth = float(bandwidth)
        self.quality = float(quality)
        self.margin = (self.output_rate - self.bandwidth) / 2.0  # :62
        self.prev_input_rate = None
        self.ir = None
        self.hist = None  # last L samples, oldest first
        self.pos = 0.0
        self.pending: List[np.ndarray] = []
        self.pending_len = 0

    def _design(self, input_rate: float):
        assert input_rate >= 0.0, "input sample rate must be positive"  # :77-83
        assert input_rate >= self.output_rate, "input sample rate must be greater than or equal to output sample rate"
        ir_len = int(math.ceil(input_rate / self.margin * self.quality))
        assert ir_len > 0
        self.ir = design_resampler_taps(
            ir_len, self.output_rate / input_rate, float(ir_len) * self.margin / input_rate, self.flt
        )
        self.hist = np.zeros(ir_len, dtype=self.C)
        self.pos = 0.0

    def _emit(self, y: np.ndarray) -> List[Samples]:
        """Re-frame into ``output_chunk_len`` chunks (resampling.rs:121-131)."""
        out = []
        if len(y) == 0:
            return out
        self.pending.append(y)
        self.pending_len += len(y)
        if self.pending_len >= self.output_chunk_len:
            cat = np.concatenate(self.pending)
            k = 0
            while len(cat) - k >= self.output_chunk_len:
                out.append(Samples(self.output_rate, cat[k : k + self.output_chunk_len].copy()))
                k += self.output_chunk_len
            rest = cat[k:]
            self.pending = [rest] if len(rest) else []
            self.pending_len = len(rest)
        return out

    def process(self, sig):
        if isinstance(sig, Event):
            return [sig]  # :135-137 (state survives)
        input_rate = sig.sample_rate
        x = np.asarray(sig.chunk, dtype=self.C)
        if input_rate != self.prev_input_rate:  # :75-102
            self.prev_input_rate = input_rate
            self._design(input_rate)
        L = len(self.ir)
        n = len(x)
        # literal f64 ``pos`` recurrence (:109-111) -> firing sample indices
        fires = []
        pos = self.pos
        orate = self.output_rate
        for j in range(n):
            pos += orate
            if pos >= input_rate:
                pos -= input_rate
                fires.append(j)
        self.pos = pos
        ext = np.concatenate([self.hist, x])  # ext[L + j] = x[j]
        if fires:
            f = np.asarray(fires, dtype=np.int64)
            base = f + 1  # window = ext[base : base + L]  (oldest .. newest = x[j])
            acc = np.zeros(len(f), dtype=self.C)
            for t in range(L):  # sequential Flt accumulation, oldest first (:112-120)
                acc = (acc + ext[base + t] * self.ir[t]).astype(self.C)
            y = acc
        else:
            y = np.zeros(0, dtype=self.C)
        self.hist = ext[-L:].copy()
        return self._emit(y)


# --------------------------------------------------------------------------
# resampling.rs:180-279 -- Upsampler
# --------------------------------------------------------------------------


class Upsampler:
    """Scatter-add form kept literal: each input adds ``x*ir`` into the ring
    starting at the next-to-be-emitted cell, then cells are emitted while
    ``pos < output_rate`` (resampling.rs:238-267)."""

    def __init__(self, flt: str, output_chunk_len: int, output_rate: float, bandwidth: float, quality: float = 3.0):
        assert output_rate >= 0.0, "output sample rate must be positive"  # :186-187
        assert bandwidth >= 0.0, "bandwidth must be positive"
        self.flt = flt
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self.output_chunk_len = int(output_chunk_len)
        self.output_rate = float(output_rate)
        self.bandwidth = float(bandwidth)
        self.quality = float(quality)
        self.prev_input_rate = None
        self.ir = None
        self.acc = None  # linearised ring: acc[0] is the next cell to be emitted
        self.pos = 0.0
        self.pending: List[np.ndarray] = []
        self.pending_len = 0

    def _design(self, input_rate: float):
        assert input_rate >= 0.0, "input sample rate must be positive"  # :207-215
        assert input_rate <= self.output_rate, "input sample rate must be smaller than or equal to output sample rate"
        assert self.bandwidth < input_rate, "bandwidth must be smaller than input sample rate"
        margin = (input_rate - self.bandwidth) / 2.0
        ir_len = int(math.ceil(self.output_rate / margin * self.quality))
        assert ir_len > 0
        self.ir = design_resampler_taps(
            ir_len, input_rate / self.output_rate, float(ir_len) * margin / self.output_rate, self.flt
        )
        self.acc = np.zeros(ir_len, dtype=self.C)
        self.pos = 0.0

    _emit = Downsampler._emit

    def process(self, sig):
        if isinstance(sig, Event):
            return [sig]  # :269-271
        input_rate = sig.sample_rate
        x = np.asarray(sig.chunk, dtype=self.C)
        if input_rate != self.prev_input_rate:
            self.prev_input_rate = input_rate
            self._design(input_rate)
        L = len(self.ir)
        orate = self.output_rate
        # worst-case output count for this chunk
        cap = int(math.ceil(len(x) * orate / input_rate)) + 2
        buf = np.zeros(L + cap, dtype=self.C)
        buf[:L] = self.acc
        q = 0
        pos = self.pos
        irC = self.ir
        for p in range(len(x)):
            buf[q : q + L] = (buf[q : q + L] + x[p] * irC).astype(self.C)  # :239-246
            while pos < orate:  # :247-265
                q += 1
                pos += input_rate
            pos -= orate  # :266
        self.pos = pos
        y = buf[:q].copy()
        self.acc = buf[q : q + L].copy()
        return self._emit(y)


# --------------------------------------------------------------------------
# modulation.rs:13-158 -- FmMod (used to synthesise inputs) and FmDemod
# --------------------------------------------------------------------------


class FmMod:
    """modulation.rs:27-70 (sequential phase accumulator, literal)."""

    def __init__(self, flt: str, deviation: float):
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self.deviation = float(deviation)
        self.current_phase = self.R(0)

    def process(self, sig):
        if isinstance(sig, Event):
            return [sig]
        R = self.R
        sr = sig.sample_rate
        x = np.asarray(sig.chunk, dtype=self.C)
        factor = R(self.deviation / sr * TAU)
        ph = np.empty(len(x), dtype=R)
        cur = self.current_phase
        tau = R(TAU)
        re = x.real.astype(R)
        for i in range(len(x)):
            cur = R(cur + R(re[i] * factor))
            cur = R(np.fmod(cur, tau))  # Rust % on floats = fmod
            ph[i] = cur
        self.current_phase = cur
        y = np.empty(len(x), dtype=self.C)
        y.real = np.cos(ph)
        y.imag = np.sin(ph)
        return [Samples(sr, y)]


class FmDemod:
    """modulation.rs:97-148."""

    def __init__(self, flt: str, deviation: float):
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self._deviation = float(deviation)
        self.previous_sample = None
        self.output_sample = self.C(0)

    def deviation(self):
        return self._deviation

    def set_deviation(self, d: float):
        self._deviation = float(d)
        return self

    def process(self, sig):
        if isinstance(sig, Event):
            if sig.is_interrupt():  # :133-136
                self.previous_sample = None
            return [sig]
        R, C = self.R, self.C
        sr = sig.sample_rate
        x = np.asarray(sig.chunk, dtype=C)
        n = len(x)
        factor = R(sr / self._deviation / TAU)  # :116
        y = np.empty(n, dtype=C)
        if n == 0:
            return [Samples(sr, y)]
        prev = np.empty(n, dtype=C)
        prev[1:] = x[:-1]
        start = 0
        if self.previous_sample is None:
            y[0] = self.output_sample  # :119-124 (repeat the previous output value)
            start = 1
        else:
            prev[0] = self.previous_sample
        if n > start:
            prod = (x[start:] * np.conj(prev[start:])).astype(C)
            ang = np.arctan2(prod.imag.astype(R), prod.real.astype(R)).astype(R)
            y[start:] = (ang * factor).astype(R)  # Complex::from(real): im = 0
        self.output_sample = C(y[-1])
        self.previous_sample = C(x[-1])
        return [Samples(sr, y)]


# --------------------------------------------------------------------------
# blocks/analysis.rs:60-132 -- Fourier (anchors the FFT convention)
# --------------------------------------------------------------------------


class Fourier:
    def __init__(self, flt: str = "f64", window: Optional[Window] = None, center_dc: bool = False):
        self.R = real_dtype(flt)
        self.C = complex_dtype(flt)
        self.window = window if window is not None else Rectangular()
        self.center_dc = center_dc

    def process(self, sig):
        if isinstance(sig, Event):
            return [sig]
        x = np.asarray(sig.chunk, dtype=self.C)
        n = len(x)
        idx = np.arange(n, dtype=np.float64)
        w = np.asarray(self.window.relative_value_at(2.0 * (idx + 0.5) / float(n) - 1.0), dtype=np.float64)
        energy = float(np.cumsum(w * w)[-1])
        scale = math.sqrt(float(n) / energy)  # analysis.rs:97
        wv = (w * scale).astype(self.R)
        y = fft_forward((x * wv).astype(self.C)).astype(self.C)
        if self.center_dc:
            y = np.roll(y, n // 2)  # rotate_right(n/2)
        return [Samples(sig.sample_rate, y)]


# --------------------------------------------------------------------------
# chunks.rs -- Rechunker and Overlapper (chunk reorganisers), SamplesLost
# --------------------------------------------------------------------------

SAMPLES_LOST = Event("SamplesLost", True)  # chunks.rs:17-28: an interrupt event


class Rechunker:
    """chunks.rs:56-176: arbitrary input chunk lengths -> chunks of ``output_chunk_len``.

    ``patchwork`` is the partially filled output chunk (None = the reference's
    ``patchwork_opt == None``); it is dropped, and ``SamplesLost`` sent, when an
    event arrives (:84-91) or the sample rate changes (:71-79).
    """

    def __init__(self, output_chunk_len: int):
        assert output_chunk_len > 0, "chunk length must be positive"  # :58
        self.output_chunk_len = int(output_chunk_len)
        self.patchwork = None  # (sample_rate, ndarray)

    def set_output_chunk_len(self, n: int):  # :171-175
        assert n > 0, "chunk length must be positive"
        self.output_chunk_len = int(n)

    def process(self, sig):
        out = []
        if isinstance(sig, Event):
            if self.patchwork is not None:
                out.append(SAMPLES_LOST)
                self.patchwork = None
            out.append(sig)
            return out
        sr = sig.sample_rate
        chunk = np.asarray(sig.chunk)
        if self.patchwork is not None and self.patchwork[0] != sr:
            self.patchwork = None
            out.append(SAMPLES_LOST)
        ocl = self.output_chunk_len
        while True:  # one turn of the task loop per pass while an input chunk is in hand (:66-164)
            if self.patchwork is not None:
                p_sr, pw = self.patchwork
                if len(pw) > ocl:  # only after set_output_chunk_len shrank the length (:101-112)
                    while len(pw) > ocl:
                        out.append(Samples(p_sr, pw[:ocl].copy()))
                        pw = pw[ocl:]
                missing = ocl - len(pw)
                if len(chunk) < missing:  # :114-117
                    self.patchwork = (p_sr, np.concatenate([pw, chunk]))
                    return out
                if len(chunk) == missing:  # :118-126
                    out.append(Samples(p_sr, np.concatenate([pw, chunk])))
                    self.patchwork = None
                    return out
                out.append(Samples(p_sr, np.concatenate([pw, chunk[:missing]])))  # :127-136
                chunk = chunk[missing:]
                self.patchwork = None
            else:
                while len(chunk) > ocl:  # :141-147
                    out.append(Samples(sr, chunk[:ocl].copy()))
                    chunk = chunk[ocl:]
                if len(chunk) == ocl:  # :148-155
                    out.append(Samples(sr, chunk.copy()))
                    return out
                self.patchwork = (sr, chunk[:0].copy())  # :156-159, filled on the next turn


class Overlapper:
    """chunks.rs:188-242: output = the last ``chunk_count`` chunks concatenated, once that many have arrived;
    sample rate = length-weighted mean (:207-214).  ANY event clears the history and is preceded by
    ``SamplesLost`` (:226-233)."""

    def __init__(self, chunk_count: int):
        assert chunk_count > 0, "chunk count must be positive"  # :195
        self.chunk_count = int(chunk_count)
        self.history = []

    def process(self, sig):
        if isinstance(sig, Event):
            self.history = []
            return [SAMPLES_LOST, sig]
        self.history.append((sig.sample_rate, np.asarray(sig.chunk)))
        if len(self.history) < self.chunk_count:
            return []
        count = 0
        acc = 0.0
        for sr, ch in self.history:
            count += len(ch)
            acc += sr * float(len(ch))
        res = Samples(acc / float(count), np.concatenate([ch for _, ch in self.history]))
        self.history.pop(0)
        return [res]


# --------------------------------------------------------------------------
# metering.rs -- level, bandwidth, rescale_energy (literal; pinned by the reference's tests, metering.rs:113-262)
# --------------------------------------------------------------------------


def _norm_sqr(x: np.ndarray) -> np.ndarray:
    """num::Complex::norm_sqr in the sample's own precision: re*re + im*im."""
    R = x.real.dtype.type
    return (x.real * x.real).astype(R) + (x.imag * x.imag).astype(R)


def level(chunk: np.ndarray) -> float:
    """metering.rs:21-30: mean square norm, accumulated in f64 in order."""
    acc = 0.0
    for v in _norm_sqr(np.asarray(chunk)):
        acc += float(v)
    return acc / float(len(chunk))


def bandwidth(double_percentile: float, sample_rate: float, bins: np.ndarray) -> float:
    """metering.rs:42-84."""
    e = [float(v) for v in _norm_sqr(np.asarray(bins))]
    n = len(e)

    def discount(energy_limit, idcs):  # :49-66
        old_energy = 0.0
        used_bins = 0.0
        for idx in idcs:
            new_energy = old_energy + e[idx]
            if new_energy > energy_limit:
                used_bins += (energy_limit - old_energy) / (new_energy - old_energy)
                break
            used_bins += 1.0
            old_energy = new_energy
        return used_bins

    total_energy = 0.0
    for v in e:
        total_energy += v
    energy_limit = total_energy * double_percentile / 2.0
    wrap_idx = (n + 1) // 2
    idcs = list(range(wrap_idx, n)) + list(range(0, wrap_idx))
    used = discount(energy_limit, idcs) + discount(energy_limit, idcs[::-1])
    bw = (float(n) - used) * sample_rate / float(n)
    return bw if bw > 0.0 else 0.0


def rescale_energy(resolution: int, inp: np.ndarray) -> np.ndarray:
    """metering.rs:93-110: every operation in Flt, in the reference's order."""
    inp = np.asarray(inp)
    R = inp.real.dtype.type
    n = len(inp)
    assert n > 0
    nsq = _norm_sqr(inp)
    out = np.zeros(resolution, dtype=R)
    for o in range(resolution):
        left = R(R(R(o) / R(resolution)) * R(n))
        right = R(R(R(R(o) + R(1)) / R(resolution)) * R(n))
        left_floor = min(int(np.floor(left)), n - 1)
        right_ceil = min(int(np.ceil(right)), n)
        acc = R(0)
        for idx in range(left_floor, right_ceil):
            left_bounded = max(R(idx), left)
            right_bounded = min(R(R(idx) + R(1)), right)
            scale = R(right_bounded - left_bounded)
            acc = R(acc + R(nsq[idx] * scale))
        out[o] = acc
    return out


# --------------------------------------------------------------------------
# A chain = blocks connected with feed_into (flow.rs:233-267); messages are
# delivered in order, events in-band (signal.rs:170-183).
# --------------------------------------------------------------------------


class Chain:
    def __init__(self, blocks: Sequence):
        self.blocks = list(blocks)

    def push(self, sig) -> list:
        msgs = [sig]
        for b in self.blocks:
            nxt = []
            for m in msgs:
                nxt.extend(b.process(m))
            msgs = nxt
        return msgs

    def run(self, sample_rate: float, x: np.ndarray, chunk_len: int) -> np.ndarray:
        """Feed ``x`` in ``chunk_len`` pieces; return the concatenated output
        samples (events dropped)."""
        outs = []
        for k in range(0, len(x) - chunk_len + 1, chunk_len):
            for m in self.push(Samples(sample_rate, x[k : k + chunk_len])):
                if isinstance(m, Samples):
                    outs.append(m.chunk)
        if not outs:
            return np.zeros(0, dtype=x.dtype)
        return np.concatenate(outs)


# --------------------------------------------------------------------------
# Literal per-sample transcriptions (slow; for small cross-checks only)
# --------------------------------------------------------------------------


def literal_downsample(x, ir, input_rate, output_rate):
    """resampling.rs:103-121 transcribed sample by sample (ring + pos)."""
    C = x.dtype.type
    L = len(ir)
    ring = np.zeros(L, dtype=x.dtype)
    rp = 0
    pos = 0.0
    out = []
    for s in x:
        ring[rp] = s
        rp += 1
        if rp == L:
            rp = 0
        pos += output_rate
        if pos >= input_rate:
            pos -= input_rate
            acc = C(0)
            t = 0
            for i in range(rp, L):
                acc = C(acc + C(ring[i] * ir[t]))
                t += 1
            for i in range(0, rp):
                acc = C(acc + C(ring[i] * ir[t]))
                t += 1
            out.append(acc)
    return np.asarray(out, dtype=x.dtype)


def literal_upsample(x, ir, input_rate, output_rate):
    """resampling.rs:238-267 transcribed sample by sample."""
    C = x.dtype.type
    L = len(ir)
    ring = np.zeros(L, dtype=x.dtype)
    rp = 0
    pos = 0.0
    out = []
    for s in x:
        t = 0
        for i in range(rp, L):
            ring[i] = C(ring[i] + C(s * ir[t]))
            t += 1
        for i in range(0, rp):
            ring[i] = C(ring[i] + C(s * ir[t]))
            t += 1
        while pos < output_rate:
            out.append(ring[rp])
            ring[rp] = 0
            rp += 1
            if rp >= L:
                rp = 0
            pos += input_rate
        pos -= output_rate
    return np.asarray(out, dtype=x.dtype)


def literal_freqshift_table(flt, sample_rate, precision, shift, start_phase=0.0):
    """transform.rs:326-339: the full ``phase_vec`` built with the integer
    recurrence (small ``denom`` only)."""
    R, C = real_dtype(flt), complex_dtype(flt)
    numer, denom = freq_to_ratio(sample_rate, precision, shift)
    tab = np.empty(denom, dtype=C)
    i = 0
    for k in range(denom):
        ph = R(R(start_phase) + R(R(i) / R(denom)) * R(TAU))
        tab[k] = C(complex(np.cos(ph), np.sin(ph)))
        i = i + numer
        i = int(math.fmod(i, denom))  # Rust %: sign of dividend
    return tab


def literal_fmdemod(x, sample_rate, deviation):
    """modulation.rs:116-126 sample by sample."""
    R = x.real.dtype.type
    C = x.dtype.type
    factor = R(sample_rate / deviation / TAU)
    prev = None
    o = C(0)
    out = []
    for s in x:
        if prev is not None:
            p = C(s * np.conj(prev))
            o = C(R(np.arctan2(R(p.imag), R(p.real))) * factor)
        out.append(o)
        prev = s
    return np.asarray(out, dtype=x.dtype)


# --------------------------------------------------------------------------
# Synthetic inputs of SURVEY.md 8(d) (seeded; shared by tests, bench, smoke)
# --------------------------------------------------------------------------


def synth_noise(seed: int, n: int, flt: str = "f32") -> np.ndarray:
    """i.i.d. N(0,1) re/im, ``seed = 20260000 + config*100000 + stream``."""
    rng = np.random.default_rng(seed)
    C = complex_dtype(flt)
    R = real_dtype(flt)
    x = np.empty(n, dtype=C)
    x.real = rng.standard_normal(n).astype(R)
    x.imag = rng.standard_normal(n).astype(R)
    return x


def synth_fm_station(seed: int, n: int, sample_rate: float, deviation: float = 75000.0, audio_bw: float = 15000.0,
                     noise_db: float = -30.0, flt: str = "f32") -> np.ndarray:
    """Config 4's input: a unit carrier FM-modulated (``FmMod``, modulation.rs:44-51) with band-limited
    noise ("audio": white noise low-passed to ``audio_bw``, peak-normalised to +-1 like a full-scale programme) plus
    white Gaussian noise ``noise_db`` below the carrier.  The modulator's phase recurrence
    ``phase = (phase + re*factor) % TAU`` is evaluated here in closed form in f64 (a cumulative sum) -- this is an
    INPUT generator, the sample-by-sample ``FmMod`` restatement above stays the parity oracle of that block."""
    rng = np.random.default_rng(seed)
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.fft.rfftfreq(n, 1.0 / sample_rate)
    spec[f > audio_bw] = 0.0
    spec[0] = 0.0
    audio = np.fft.irfft(spec, n)
    audio /= np.max(np.abs(audio))
    factor = deviation / sample_rate * TAU  # modulation.rs:44
    phase = np.fmod(np.cumsum(audio * factor), TAU)
    sigma = 10.0 ** (noise_db / 20.0) / math.sqrt(2.0)
    x = np.exp(1j * phase) + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(complex_dtype(flt))


def lowpass(cutoff_hz: float):
    """The configs' ``|f| <= cutoff -> 1 else 0`` response."""

    def f(_bin: int, freq: float) -> complex:
        return 1.0 + 0.0j if abs(freq) <= cutoff_hz else 0.0j

    return f


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """Parity metric of SURVEY.md 8(d): ``||a-b||_2 / ||b||_2``."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = float(np.linalg.norm(b.astype(np.complex128)))
    num = float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)))
    return num / den if den > 0 else num

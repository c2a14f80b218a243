"""ctypes face of oracle/radiorust_oracle.c (TEST INFRASTRUCTURE: the CPU checker and the
CPU baseline of bench.py; never imported by the product package)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_c
from . import radiorust_oracle as orc

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build_c.SO if os.path.exists(build_c.SO) else build_c.build()
        _lib = C.CDLL(so)
        for name in ("oracle_chain_f32", "oracle_chain_f64"):
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                           C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int,
                           C.c_void_p]
        for name in ("oracle_phase_table_f32", "oracle_phase_table_f64"):
            fn = getattr(_lib, name)
            fn.restype = None
            fn.argtypes = [C.c_int64, C.c_int64, C.c_void_p]
    return _lib


def chain(x: np.ndarray, flt: str, sample_rate: float, chunk_len: int, shifts=None, precision: float = 1.0,
          freq_resp=None, window=None, down=None, n_threads: int = 1, timing: dict | None = None,
          prebuilt_phase_tables: bool = False) -> list:
    """Runs [FreqShifter ->] [Filter ->] [Downsampler] over the rows of ``x`` from a fresh start.

    ``shifts``: per-stream shift in hertz or None; ``freq_resp``: Filter closure or None;
    ``down``: (output_rate, bandwidth, quality) or None.  Returns one output array per stream
    (Downsampler output framing is not applied: all produced samples).  ``prebuilt_phase_tables``: FreqShifter's
    phase tables (transform.rs:321-340, built once per retune by the reference) are built before the timed call
    (``timing["table_seconds"]``) instead of inside it.
    """
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=orc.complex_dtype(flt))
    S, total = x.shape
    n_chunks = total // chunk_len
    R = orc.real_dtype(flt)
    numer = denom = None
    if shifts is not None:
        rat = [orc.freq_to_ratio(sample_rate, precision, float(s)) for s in shifts]
        numer = np.array([r[0] for r in rat], dtype=np.int64)
        denom = np.array([r[1] for r in rat], dtype=np.int64)
    hext = None
    if freq_resp is not None:
        win = window if window is not None else orc.Kaiser.with_null_at_bin(2.0)
        hext = np.ascontiguousarray(orc.design_filter_response(freq_resp, win, sample_rate, chunk_len, flt))
    ir = None
    L = 0
    in_rate = out_rate = 0.0
    if down is not None:
        out_rate, bw, quality = down
        d = orc.Downsampler(flt, 1, out_rate, bw, quality)
        d._design(sample_rate)
        ir = np.ascontiguousarray(d.ir, dtype=R)
        L = len(ir)
        in_rate = sample_rate
    cap = total + 8
    ys = [np.zeros(cap, dtype=x.dtype) for _ in range(S)]
    xp = (C.c_void_p * S)(*[x[s].ctypes.data for s in range(S)])
    yp = (C.c_void_p * S)(*[y.ctypes.data for y in ys])
    n_out = (C.c_size_t * S)()
    fn = lib().oracle_chain_f32 if flt == "f32" else lib().oracle_chain_f64
    import time

    tabs_p = None
    if prebuilt_phase_tables and numer is not None:
        tfn = lib().oracle_phase_table_f32 if flt == "f32" else lib().oracle_phase_table_f64
        tt = time.perf_counter()
        tabs = [np.empty(int(d), dtype=x.dtype) for d in denom]
        for s in range(S):
            tfn(int(numer[s]), int(denom[s]), tabs[s].ctypes.data)
        tabs_p = (C.c_void_p * S)(*[t.ctypes.data for t in tabs])
        if timing is not None:
            timing["table_seconds"] = time.perf_counter() - tt

    t0 = time.perf_counter()
    rc = fn(S, chunk_len, n_chunks, xp, yp, n_out, 1 if shifts is not None else 0,
            numer.ctypes.data if numer is not None else None, denom.ctypes.data if denom is not None else None,
            1 if hext is not None else 0, hext.ctypes.data if hext is not None else None,
            1 if down is not None else 0, ir.ctypes.data if ir is not None else None, L, in_rate, out_rate, n_threads, tabs_p)
    if timing is not None:
        timing["seconds"] = time.perf_counter() - t0  # the C hot loops only (design excluded)
    if rc != 0:
        raise RuntimeError(f"oracle_chain failed: {rc}")
    return [ys[s][: n_out[s]] for s in range(S)]

"""Generates the golden fixtures in this directory with the numpy oracle.

    python tests/golden/make_golden.py

The reference itself (Rust; needs cargo + rustfft, neither present) cannot be run here, and it
ships no golden vectors for these blocks, so these are ORACLE outputs on seeded
inputs: they freeze the oracle (any later change to it shows up as a diff) and give the C-ABI
tests inputs/outputs that do not depend on the oracle code at test time.  Inputs are
regenerated from the recorded seeds (`orc.synth_noise`), only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import radiorust_oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    """name -> dict(flt, sample_rate, chunk_len, n_chunks, seed, blocks=[(kind, params...)])"""
    return {
        # BASELINE config 1 at reduced length
        "c1_chain_f32": dict(flt="f32", sample_rate=1_024_000.0, chunk_len=4096, n_chunks=12, seed=20260000 + 100000,
                             blocks=[("freqshift", 1.0, 123457.0), ("filter_lowpass", 3000.0), ("downsample", 192, 48000.0, 6000.0, 3.0)]),
        # config 3 shape (P = 50)
        "c3_chain_f32": dict(flt="f32", sample_rate=2_400_000.0, chunk_len=4096, n_chunks=12, seed=20260000 + 300000 + 5,
                             blocks=[("freqshift", 1.0, float((5 * 577) % 2_400_000 - 1_200_000)), ("filter_lowpass", 3000.0),
                                     ("downsample", 64, 48000.0, 6000.0, 3.0)]),
        # wide-band FM receive chain of config 4 at reduced rate/length
        # (input: FM-modulated band-limited noise + AWGN at -30 dB, SURVEY.md 8d -- a discriminator fed with
        # filtered noise sits on its phase-wrap discontinuities and measures nothing but input rounding)
        "c4_fm_f32": dict(flt="f32", sample_rate=1_000_000.0, chunk_len=4096, n_chunks=10, seed=20260000 + 400000, input="fm",
                          blocks=[("filter_lowpass", 100000.0), ("fmdemod", 75000.0), ("filter_deemph", 50e-6),
                                  ("downsample", 128, 50000.0, 40000.0, 3.0)]),
        # f64 precision case: filter + upsampler (config 5 at reduced size)
        "c5_filter_f64": dict(flt="f64", sample_rate=2_400_000.0, chunk_len=4096, n_chunks=6, seed=20260000 + 500000,
                              blocks=[("filter_lowpass", 20000.0)]),
        "c5_upsample_f64": dict(flt="f64", sample_rate=48000.0, chunk_len=256, n_chunks=3, seed=20260000 + 500001,
                                blocks=[("upsample", 1000, 2_400_000.0, 20000.0, 3.0)]),
    }


def deemph_resp(tau):
    def f(bin_, freq):  # examples/relm_app/simple_receiver.rs:43-49
        if bin_ != 0 and 20.0 <= abs(freq) <= 16000.0:
            return orc.deemphasis_factor(tau, freq)
        return 0j

    return f


def oracle_blocks(case):
    flt = case["flt"]
    out = []
    for b in case["blocks"]:
        k = b[0]
        if k == "freqshift":
            out.append(orc.FreqShifter(flt, b[1], b[2]))
        elif k == "filter_lowpass":
            out.append(orc.Filter.new(flt, orc.lowpass(b[1])))
        elif k == "filter_deemph":
            out.append(orc.Filter.new_rectangular(flt, deemph_resp(b[1])))
        elif k == "downsample":
            out.append(orc.Downsampler(flt, b[1], b[2], b[3], b[4]))
        elif k == "upsample":
            out.append(orc.Upsampler(flt, b[1], b[2], b[3], b[4]))
        elif k == "fmdemod":
            out.append(orc.FmDemod(flt, b[1]))
    return out


def case_input(case):
    total = case["chunk_len"] * case["n_chunks"]
    if case.get("input") == "fm":
        return orc.synth_fm_station(case["seed"], total, case["sample_rate"], 75000.0, 15000.0, -30.0, case["flt"])
    return orc.synth_noise(case["seed"], total, case["flt"])


def run_case(case):
    x = case_input(case)
    return orc.Chain(oracle_blocks(case)).run(case["sample_rate"], x, case["chunk_len"])


def emit_inputs(out_dir):
    """Inputs, oracle outputs and the case list in the plain binary form rust/radiorust-b200/tests/emit_golden.rs
    reads: with cargo, that test runs REAL radiorust on these inputs and compares (pins the oracle)."""
    os.makedirs(out_dir, exist_ok=True)
    lines = ["# name flt sample_rate chunk_len n_chunks blocks"]
    for name, case in cases().items():
        x = case_input(case)
        x.tofile(os.path.join(out_dir, name + ".input.bin"))
        run_case(case).tofile(os.path.join(out_dir, name + ".oracle.bin"))
        blocks = ";".join(":".join(repr(float(v)) if isinstance(v, float) else str(v) for v in b) for b in case["blocks"])
        lines.append(f"{name} {case['flt']} {case['sample_rate']!r} {case['chunk_len']} {case['n_chunks']} {blocks}")
        if case.get("input") == "fm":
            lines[-1] += ""
    with open(os.path.join(out_dir, "cases.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out_dir)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--emit-inputs":
        emit_inputs(sys.argv[2])
        sys.exit(0)
    for name, case in cases().items():
        y = run_case(case)
        np.save(os.path.join(HERE, name + ".npy"), y)
        print(name, y.dtype, y.shape)

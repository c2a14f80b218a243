"""The Rust side exists as files and stays in step with the C header (no rustc in this image: these are textual
checks; the crates are compiled by a maintainer with cargo, see rust/README.md)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "rust")


def header_functions():
    src = open(os.path.join(ROOT, "include", "radiorust_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", src)) - {"rr_freq_resp_fn", "rr_window_fn"})


def test_sys_bindings_are_generated_from_the_header():
    r = subprocess.run([sys.executable, os.path.join(RUST, "gen_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_header_symbol_is_bound_in_lib_rs():
    lib = open(os.path.join(RUST, "radiorust-b200-sys", "src", "lib.rs")).read()
    names = header_functions()
    assert len(names) >= 50
    for n in names:
        assert re.search(r"pub fn %s\(" % n, lib), f"{n} is declared in the header but not bound in radiorust-b200-sys"
    bound = set(re.findall(r"pub fn (rr_[a-z0-9_]+)\(", lib))
    assert bound == set(names)
    # struct layouts: same field order as the header
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "radiorust_b200.h")).read(), flags=re.S)
    for struct in ("rr_stage_desc", "rr_chain_desc"):
        body = re.search(r"typedef struct %s \{(.*?)\}" % struct, hdr, flags=re.S).group(1)
        c_fields = [re.match(r".*?([A-Za-z_][A-Za-z0-9_]*)$", " ".join(f.split())).group(1) for f in body.split(";") if f.strip()]
        r_body = re.search(r"pub struct %s \{(.*?)\}" % struct, lib, flags=re.S).group(1)
        r_fields = re.findall(r"pub (\w+):", r_body)
        assert c_fields == r_fields, struct


def test_build_rs_compiles_the_same_translation_units_as_build_py():
    from radiorust_b200 import build as b

    listed = [l.split() for l in open(os.path.join(RUST, "radiorust-b200-sys", "translation_units.txt")) if l.strip() and not l.startswith("#")]
    want = [[obj, src] + list(extra) for obj, src, extra in b.translation_units()]
    assert listed == want
    build_rs = open(os.path.join(RUST, "radiorust-b200-sys", "build.rs")).read()
    for flag in b.NVCC_FLAGS:
        assert f'"{flag}"' in build_rs, flag
    assert "arch=compute_100a,code=sm_100a" in build_rs
    for src in {u[1] for u in listed}:
        assert os.path.exists(os.path.join(RUST, "radiorust-b200-sys", "csrc", src))  # the csrc symlink resolves


def test_wrapper_crate_has_the_reference_block_api():
    blocks = open(os.path.join(RUST, "radiorust-b200", "src", "blocks.rs")).read()
    for ty in ("GpuFreqShifter", "GpuFilter", "GpuDownsampler", "GpuUpsampler", "GpuFmDemod", "GpuChain"):
        assert re.search(r"pub struct %s<Flt>" % ty, blocks) or re.search(r"resampler_block!\(%s," % ty, blocks), ty
        assert re.search(r"impl_block_trait! \{ <Flt> Consumer<Signal<Complex<Flt>>> for (%s|\$name)<Flt> \}" % ty, blocks), ty
    # constructors / setters of the reference (transform.rs:282-297,376-390; filters.rs:128-152,279-297; modulation.rs:97,150-157)
    for fn in ("with_shift", "with_precision", "with_precision_and_shift", "precision", "shift", "set_shift", "update_shift",
               "new_rectangular", "with_window", "update", "update_with_window", "with_quality", "deviation", "set_deviation"):
        assert re.search(r"pub fn %s[<(]" % fn, blocks), fn
    pool = open(os.path.join(RUST, "radiorust-b200", "src", "pool.rs")).read()
    for ty in ("PinnedChunkBufPool", "PinnedChunkBuf", "PinnedChunk"):
        assert "pub struct %s<T>" % ty in pool
    for fn in ("get_with_capacity", "finalize", "separate_beginning", "discard_beginning"):  # bufferpool.rs:60-79,141,210
        assert "pub fn %s" % fn in pool
    assert "rr_pool_get" in pool and "rr_pool_put" in pool
    golden = open(os.path.join(RUST, "radiorust-b200", "tests", "emit_golden.rs")).read()
    assert "radiorust::blocks::filters::{deemphasis_factor, Filter}" in golden and "oracle_matches_radiorust" in golden


def test_emit_inputs_for_the_rust_golden_test(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden

    make_golden.emit_inputs(str(tmp_path))
    lines = [l for l in open(tmp_path / "cases.txt") if not l.startswith("#")]
    assert len(lines) == len(make_golden.cases())
    for name, case in make_golden.cases().items():
        size = os.path.getsize(tmp_path / f"{name}.input.bin")
        assert size == case["chunk_len"] * case["n_chunks"] * (8 if case["flt"] == "f32" else 16)
        assert os.path.getsize(tmp_path / f"{name}.oracle.bin") > 0

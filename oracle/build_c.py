"""Builds the oracle's C restatement (test infrastructure) into oracle/_build/.

`python -m oracle.build_c`.  gcc only; no dependency on the product library.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "libradiorust_oracle.so")
SRC = os.path.join(HERE, "radiorust_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    gcc = shutil.which("gcc") or "gcc"
    # -ffp-contract=off: the reference (Rust) never fuses a*b+c; -march=native is left out so the
    # library built here also runs on the GPU box's host CPU
    cmd = [gcc, "-O3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", SO, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force=True))

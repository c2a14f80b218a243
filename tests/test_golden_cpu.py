"""The numpy oracle reproduces the committed golden fixtures (tests/golden/*.npy)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

from oracle import radiorust_oracle as orc  # noqa: E402


@pytest.mark.parametrize("name", sorted(make_golden.cases()))
def test_oracle_reproduces_golden(name):
    case = make_golden.cases()[name]
    want = np.load(os.path.join(HERE, "golden", name + ".npy"))
    got = make_golden.run_case(case)
    assert got.shape == want.shape and got.dtype == want.dtype
    # same code, same seeds: bit-identical unless numpy/scipy changed their rounding
    tol = 1e-6 if case["flt"] == "f32" else 1e-13
    assert orc.rel_l2(got, want) <= tol

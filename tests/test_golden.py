"""Committed regression vectors (tests/golden/oracle_case_*.npz, made by make_golden.py from the oracle):
the oracle must still reproduce them bit-for-bit on CPU, and the CUDA path must hit them on the GPU."""
import os
import sys

import numpy as np
import pytest

import harness

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    gold = np.load(os.path.join(G, "oracle_case_%s.npz" % name))
    ref = harness.run_oracle(make_golden.build(name))
    assert np.array_equal(ref["sel_count"], gold["sel_count"]) and np.array_equal(ref["sel_hash"], gold["sel_hash"])
    # same binary, same libm on this image: bit-exact; allow 1e-13 for a different glibc on another box
    assert harness.rel_diff(ref["o"], gold["o"]) < 1e-13
    assert np.max(np.abs(ref["tb"] - gold["tb"])) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_gpu_hits_golden(name):
    gold = np.load(os.path.join(G, "oracle_case_%s.npz" % name))
    gpu = harness.run_gpu(make_golden.build(name))
    assert np.array_equal(gpu["sel_count"], gold["sel_count"]) and np.array_equal(gpu["sel_hash"], gold["sel_hash"])
    assert harness.rel_diff(gpu["o"], gold["o"]) < 1e-9          # north_star: 1e-9 relative
    assert np.max(np.abs(gpu["tb"] - gold["tb"])) < 1e-5          # north_star: 1e-5 K
    assert np.max(np.abs(gpu["tmr"] - gold["tmr"])) < 1e-5
    scale = np.abs(gold["o"])
    assert np.max(np.abs(gpu["o_by_mol"].sum(axis=1) - gold["o_by_mol_sum"]) / scale) < 1e-9
    assert np.max(np.abs(gpu["oc"].sum(axis=1) - gold["oc_sum"]) / scale) < 1e-9

"""CPU: the C-ABI library loads and exports every symbol include/snag_b200.h declares; host-side glue
(metrics, shard partition, argument validation). No compute calls are made without a GPU."""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import oracle
from snag_b200 import _lib, evaluate
from tests.conftest import ROOT, load_golden


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "snag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snag_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/snag_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes signature table out of sync with the header"


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.snag_version() == 1
    assert b"sm_100" in lib.snag_error_string(-5)
    assert lib.snag_error_string(0) == b"ok"
    with pytest.raises(_lib.SnagError):
        _lib.check(-2, "unit-test")


def test_no_torch_types_in_the_abi():
    """The boundary is plain pointers and sizes: the shared object must not link libtorch / libc10."""
    out = os.popen(f"ldd {_lib.LIB_PATH}").read()
    # library names only: the load addresses ldd prints are random hex and can spell "c10"
    names = [line.split()[0] for line in out.splitlines() if line.strip()]
    assert names, "ldd printed nothing"
    assert not any("torch" in nm or "c10" in nm for nm in names), names


def test_cpu_tensors_are_rejected_loudly():
    from snag_b200 import ops
    with pytest.raises(_lib.SnagError):
        ops.prep_bf16(torch.zeros(4, 64), None, True)


def test_sim_plan_needs_no_gpu_for_validation():
    lib = _lib.load()
    tpc, nl = ctypes.c_int32(), ctypes.c_int32()
    assert lib.snag_sim_plan(0, 10, 64, ctypes.byref(tpc), ctypes.byref(nl)) == -2
    assert lib.snag_sim_plan(10, 10, 65, ctypes.byref(tpc), ctypes.byref(nl)) == -2


@pytest.mark.parametrize("name", ["eval_n384_d96_k10", "eval_ties_dyadic_k4", "eval_n700_d320_k10"])
def test_metrics_glue_matches_reference(name):
    fx = load_golden(name)
    for side in ("l2r", "r2l"):
        m = evaluate.metrics_from_ranks(fx[f"rank_{side}"])
        np.testing.assert_array_equal(m.acc, fx[f"acc_{side}"])
        assert m.mr == float(fx[f"mr_{side}"])
        assert m.mrr == float(fx[f"mrr_{side}"])          # same additions in the same order -> same bits
        o = oracle.metrics(fx[f"rank_{side}"])
        assert o["mrr"] == m.mrr and o["mr"] == m.mr


@pytest.mark.parametrize("n,world", [(10500, 1), (10500, 2), (10500, 8), (1000000, 8), (300, 8), (257, 2), (5, 4)])
def test_shard_bounds_partition(n, world):
    covered = []
    for r in range(world):
        c0, c1 = evaluate.shard_bounds(n, world, r)
        assert 0 <= c0 <= c1 <= n
        if c1 > c0 and c1 != n:
            assert (c1 - c0) % 256 == 0
        covered.extend(range(c0, c1) if n <= 20000 else [c0, c1])
    if n <= 20000:
        assert covered == list(range(n))
    else:
        assert covered[0] == 0 and covered[-1] == n
        assert all(covered[2 * i + 1] == covered[2 * i + 2] for i in range(world - 1))

"""CPU-side checks of the product boundary (no GPU needed): the C-ABI library loads and exports every
symbol include/arks_b200.h declares, fails loudly without a device, and its host-side exact decision
table for headOrTail agrees with the oracle's float/erf expression."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    lib = os.path.join(ROOT, "arcs_b200", "lib", "libarks_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "arcs_b200", "csrc")])
    return lib


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_build())
    header = open(os.path.join(ROOT, "include", "arks_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(arks_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 24
    for name in sorted(declared):
        assert hasattr(lib, name), "libarks_b200.so does not export " + name
    import arcs_b200.api as api
    assert declared == set(api.SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import arcs_b200
    with pytest.raises(arcs_b200.ArksError) as e:
        arcs_b200.ArksIndex(30, 1000)
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("min_reads,r", [(5, 0.05), (1, 0.05), (3, 0.01), (2, 0.2), (5, 0.5), (10, 0.001)])
def test_head_tail_table_matches_reference_expression(min_reads, r):
    import arcs_b200
    n = 600
    table = arcs_b200.head_tail_table(min_reads, r, n)
    r32 = np.float32(r)
    for s in range(n):
        lo = (s + 1) // 2
        for mx in range(lo, s + 1):
            valid, _ = O.head_or_tail(mx, s - mx, min_reads, r32)
            assert valid == (mx >= table[s]), (s, mx, int(table[s]))


def test_product_does_not_touch_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "arcs_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle/" not in txt and "oracle_lib" not in txt and "arks_oracle" not in txt, os.path.join(base, f)

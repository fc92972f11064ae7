"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/pcseq_b200.h declares (no compute calls without a GPU), and the product refuses CPU tensors."""
import os
import re

import pytest
import torch

from pcseqlearning_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pcseq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _header_symbols()
    assert declared, "no declarations found in include/pcseq_b200.h"
    for name in declared:
        assert hasattr(L, name), f"libpcseq_b200.so does not export {name}"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table is out of sync with the header"
    assert L.pcs_version() >= 100


def test_no_cpu_fallback():
    pts = torch.zeros(8, 4)
    with pytest.raises(_lib.PcsError):
        ops.radius_graph(pts, pts, 0.5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pcseqlearning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in txt

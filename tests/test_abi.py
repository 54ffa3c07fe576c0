"""CPU: the C-ABI library loads without a GPU and exports every symbol include/b200_consensus.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "b200_consensus.h").read_text()
DECLARED = sorted(set(re.findall(r"B200_API[^;(]*?\b(b200_\w+)\s*\(", HEADER)))


def test_header_declares_something():
    assert len(DECLARED) >= 15


def test_library_loads_and_exports_every_declared_symbol():
    from ethereum_consensus_b200 import _lib
    lib = _lib.load()
    missing = [s for s in DECLARED if not hasattr(lib, s)]
    assert not missing, missing
    # every ctypes prototype the package binds must be declared in the header too
    assert set(_lib._PROTOS) <= set(DECLARED), sorted(set(_lib._PROTOS) - set(DECLARED))


def test_no_cpu_fallback_without_gpu():
    """Without a device the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ethereum_consensus_b200 import _lib, ssz
    with pytest.raises(_lib.EngineError):
        ssz.merkleize(bytes(64))


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import, link or dlopen anything under oracle/."""
    pat = re.compile(r"(^|\s)(import|from)\s+oracle\b|liboracle|oracle/|#include\s+\".*oracle")
    pkg = ROOT / "ethereum_consensus_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        for line in p.read_text().splitlines():
            assert not pat.search(line), (p, line)

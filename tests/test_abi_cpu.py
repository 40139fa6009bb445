"""CPU: the C-ABI library loads, exports every symbol include/llmgw_b200.h declares, and refuses
to run without a device (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

from llmapigateway_b200 import _abi, _native

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "llmgw_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgw_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not _native.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(str(_native.LIB_PATH))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(_native.EXPORTS) <= set(names)


def test_struct_sizes_match_header_layout():
    assert C.sizeof(_abi.Val) == 16
    assert C.sizeof(_abi.UsageRec) == 8 * 16 + 4 + 2 * _abi.STR_CAP + 4       # 8-byte aligned
    assert C.sizeof(_abi.SegResult) == 12
    assert C.sizeof(_abi.RowEvent) == 8 + C.sizeof(_abi.UsageRec)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import llmapigateway_b200 as L
    with pytest.raises(L.EngineError, match="no usable CUDA device"):
        L.Engine()


def test_product_never_imports_oracle():
    for p in (ROOT / "llmapigateway_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p

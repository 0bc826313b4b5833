"""The C-ABI library loads on a CPU-only box and exports every symbol include/lcd_b200.h declares;
the ctypes table in rtabmap_b200/capi.py covers exactly the same set; and the product fails
loudly (no CPU fallback) when no CUDA device exists."""
import re
from pathlib import Path

import pytest

import rtabmap_b200
from rtabmap_b200 import capi

HEADER = (Path(__file__).resolve().parent.parent / "include" / "lcd_b200.h").read_text()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(lcd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "lcd_create" in syms and "lcd_localize_batch" in syms and len(syms) >= 30


def test_library_exports_every_declared_symbol():
    lib = rtabmap_b200.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"liblcd_b200.so does not export {s}"
    assert lib.lcd_abi_version() == 2
    assert lib.lcd_build_arch() == b"sm_100a"


def test_ctypes_table_matches_header():
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_product_never_imports_the_oracle():
    pkg = Path(rtabmap_b200.__file__).parent
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        src = f.read_text()
        assert "oracle" not in src.replace("oracle/", "").lower() or "import oracle" not in src
        assert "from oracle" not in src and "import oracle" not in src and "liboracle" not in src


def test_engine_fails_loudly_without_gpu(has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(rtabmap_b200.LcdError):
        rtabmap_b200.Engine()

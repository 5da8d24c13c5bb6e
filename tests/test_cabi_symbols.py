"""CPU: the C-ABI library builds, loads and exports every symbol include/stcat_b200.h declares
(no compute calls without a GPU)."""
import os
import re

from stcat_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "stcat_b200.h")).read()
    return sorted(set(re.findall(r"STCAT_API\s+[\w\s\*]+?\b(stcat_\w+)\s*\(", src)))


def test_header_and_binding_table_agree():
    assert declared_symbols() == sorted(cabi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from stcat_b200.build import build

    build()
    lib = cabi.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.stcat_abi_version() == cabi.ABI_VERSION
    assert f"#define STCAT_ABI_VERSION {cabi.ABI_VERSION}" in open(os.path.join(ROOT, "include", "stcat_b200.h")).read()


def test_product_path_raises_without_library(tmp_path, monkeypatch):
    import pytest

    monkeypatch.setattr(cabi, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU"):
        cabi.load_library(str(tmp_path / "missing.so"))

"""The C-ABI library loads on a CPU-only box and exports exactly what include/gatk.h declares."""
import os
import re

from pygat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gatk.h")).read()
    return sorted(set(re.findall(r"GATK_API[^;(]*?\b(gatk_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 20
    assert sorted(_lib.PROTOTYPES) == names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.gatk_version() == 100


def test_queries_work_without_gpu():
    assert _lib.query("gatk_gemm_workspace_bytes", 1, 0, 100, 512, 1000) >= 0
    assert _lib.query("gatk_da_workspace_floats", 8, 64) > 0
    assert _lib.query("gatk_scan_workspace_bytes", 1000) >= 0  # CUB sizes its scratch per device: 0 without one
    assert _lib.query("gatk_transpose_workspace_bytes", 1000, 1000, 5000) >= 3 * 5000 * 4


def test_header_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "gatk.h")).read()
    assert text.count("layers.py:") >= 10 and "models.py:" in text

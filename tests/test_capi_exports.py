"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads,
exports every symbol include/afb200.h declares, and refuses to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from arcanefem_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "afb200.h")).read()
    return sorted(set(re.findall(r"AFB_API\s+[\w\s\*]+?\b(afb_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        capi.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    capi.lib().afb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in capi.lib().afb_version()


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.AfbError):
        capi.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "arcanefem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("no cpu fallback", ""), os.path.join(dirpath, f)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), f


def test_reference_option_names_map_to_variants():
    """The reference's matrix-format option names (modules/testlab/Fem.axl:42-95, modules/poisson/Fem.axl:31) stay valid
    (host logic only: no device needed)."""
    A = capi
    assert A.options_from_name("csr-gpu") == (A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC, A.SPARSITY_FROM_CELLS)
    assert A.options_from_name("nwcsr") == (A.FORMAT_CSR, A.VARIANT_TILED_GATHER, A.SPARSITY_FROM_CONNECTIVITY)
    assert A.options_from_name("coo-sorting-gpu")[0] == A.FORMAT_COO
    assert A.options_from_name("bsr") == (A.FORMAT_BSR, A.VARIANT_CELLWISE_ATOMIC, A.SPARSITY_FROM_CELLS)
    assert A.options_from_name("AF-BSR") == A.options_from_name("bsr-atomic-free") == (A.FORMAT_BSR, A.VARIANT_TILED_GATHER, A.SPARSITY_FROM_CONNECTIVITY)
    for name in ("legacy", "DOK", "coo", "coo-sorting", "csr", "coo-gpu", "blcsr"):
        A.options_from_name(name)
    with pytest.raises(A.AfbError):
        A.options_from_name("ellpack")

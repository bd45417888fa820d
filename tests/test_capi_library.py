"""CPU-side checks of the boundary: the nvcc-built library loads and exports every symbol
include/laps_b200.h declares, the ctypes mirror matches the header, and the product path fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge
from laps_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    return capi.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "laps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(laps_[a-z0-9_]+)\s*\(", text)) - {"laps_barrier_fn"})


def test_header_and_binding_list_the_same_symbols():
    assert header_functions() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), name


def test_params_struct_matches_header_field_order():
    text = open(os.path.join(ROOT, "include", "laps_b200.h")).read()
    body = re.search(r"typedef struct laps_params \{(.*?)\} laps_params;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, rest = decl.split(None, 1)
        names += [n.strip() for n in rest.split(",")]
    assert names == [f[0] for f in capi.LapsParams._fields_]


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = capi.make_params(nx=16, ny=16, nz=16)
    h = C.c_void_p()
    rc = lib.laps_create(C.byref(p), C.byref(h))
    assert rc != 0 and not h
    assert b"no CUDA device" in lib.laps_last_error(None)


def test_create_rejects_bad_arguments(lib):
    h = C.c_void_p()
    for n in (24, 36, 144, 100, 4096):     # below 48 with an odd factor, 9 * 2^k, 25 * 4, too long
        p = capi.make_params(nx=n, ny=16, nz=16)
        assert lib.laps_create(C.byref(p), C.byref(h)) != 0
        assert b"3 * 2^k or 5 * 2^k" in lib.laps_last_error(None)
    p = capi.make_params(nx=16, ny=16, nz=16)
    p.abi_version = 99
    assert lib.laps_create(C.byref(p), C.byref(h)) != 0
    assert b"abi_version" in lib.laps_last_error(None)


def test_package_has_no_reference_to_the_oracle():
    pkg = os.path.join(ROOT, "laps_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_fortran_interface_module_binds_every_symbol_of_the_header():
    """integration/laps_gpu.f90 (the ISO_C_BINDING module a LAPS driver would use) has a bind(C) interface for every
    entry point include/laps_b200.h declares."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    f90 = open(os.path.join(root, "integration", "laps_gpu.f90")).read()
    bound = set(re.findall(r"bind\(C,\s*name='(\w+)'\)", f90))
    assert set(capi.SYMBOLS) <= bound, sorted(set(capi.SYMBOLS) - bound)

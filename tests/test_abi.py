"""The drop-in boundary on a machine WITHOUT a GPU: libmetheor_b200.so loads, exports every function include/metheor_b200.h
declares (and the ctypes mirror lists exactly those), answers the calls that need no device, and refuses — loudly, with an
error code and a message — everything that would need one: there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

from metheor_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # prose in comments mentions functions too
    return set(re.findall(r"\b(mth_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_function():
    L = _lib.lib()
    declared = _declared("metheor_b200.h")
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/metheor_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_struct_mirrors_match_the_header_field_order():
    text = open(os.path.join(ROOT, "include", "metheor_b200.h")).read()
    for cname, mirror in (("mth_batch_compact", _lib.BatchCompact), ("mth_batch", _lib.Batch), ("mth_tag_batch", _lib.TagBatch),
                          ("mth_tag_result", _lib.TagResult)):
        body = re.findall(r"typedef struct(?: %s)? \{((?:(?!typedef struct).)*?)\} %s;" % (cname, cname), text, re.S)[-1]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                for part in decl.split(","):
                    fields.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[^\]]*\])?$", part.strip())[0])
        assert fields == [f[0] for f in mirror._fields_], cname


def test_calls_that_need_no_device():
    L = _lib.lib()
    assert b"sm_100a" in L.mth_version()
    p = _lib.Params()
    L.mth_params_default(C.byref(p))
    assert p.abi_version == 2 and p.pdr.min_depth == 10 and p.pdr.min_cpgs == 4 and p.lpmd.max_distance == 16
    a = L.mth_reservoir_draw(7, 0, 1234, 100)
    assert 1 <= a <= 100 and a == L.mth_reservoir_draw(7, 0, 1234, 100)


def test_no_cpu_fallback_without_a_device():
    L = _lib.lib()
    if L.mth_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    p = _lib.Params()
    L.mth_params_default(C.byref(p))
    p.measures = 1
    ref = (C.c_int64 * 1)(1000)
    assert L.mth_ctx_create(C.byref(ctx), 0, C.byref(p), 1, ref) == -2  # MTH_ERR_CUDA
    assert b"no CPU fallback" in L.mth_last_error(None)
    g = C.c_void_p()
    assert L.mth_genome_create(C.byref(g), 0, 1, ref) == -2
    assert b"no CPU fallback" in L.mth_genome_last_error(None)

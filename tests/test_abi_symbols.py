"""The C-ABI libraries load and export every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"JB_API[^;(]*?\b(jbh?_\w+)\s*\(", txt)))


def test_headers_declare_entry_points():
    assert "jb_decode_batch_create" in declared("jpegb200.h")
    assert "jbh_parse" in declared("jpegb200_host.h")


@pytest.mark.parametrize("header,lib", [("jpegb200.h", "libjpegb200.so"), ("jpegb200_host.h", "libjpegb200_host.so")])
def test_library_exports_every_declared_symbol(header, lib):
    path = os.path.join(ROOT, "jpeglibrary_b200", "lib", lib)
    assert os.path.exists(path), "run `python __graft_entry__.py` (build) first"
    dll = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    for name in declared(header):
        assert hasattr(dll, name), f"{lib} does not export {name}"


def test_python_binding_covers_header():
    import jpeglibrary_b200 as J
    assert set(declared("jpegb200.h")) == set(J._native.CUDA_SYMBOLS)
    assert set(declared("jpegb200_host.h")) == set(J._native.HOST_SYMBOLS)


def test_no_cpu_fallback_without_device():
    import jpeglibrary_b200 as J
    if J._native.cuda.jb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(J.CudaRuntimeError):
        J.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jpeglibrary_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_ffi" not in txt and "jpeg_oracle" not in txt and "libjpeg_oracle" not in txt, f

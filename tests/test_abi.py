"""The C-ABI shared library loads without a GPU and exports every symbol include/ucfp_cuda.h declares; the
product path has no CPU fallback and fails loudly without a device."""
import ctypes
import os
import re
import subprocess

import pytest

from ucfp_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "ucfp_cuda.h")).read()


def declared_symbols():
    return sorted(set(re.findall(r"UCFP_API\s+[\w\s\*]+?\b(ucfp_\w+)\s*\(", HEADER)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ucfp_cuda.h but not exported"
    assert sorted(_ffi.PROTOTYPES) == names, "ucfp_b200/_ffi.py and the header drifted apart"
    assert _ffi.lib().ucfp_abi_version() == 2


def test_header_is_plain_c():
    """The boundary must be bindable from Rust/cgo/JNI: compile the header as C11, no C++ or CUDA types."""
    src = '#include "ucfp_cuda.h"\nint main(void){ return sizeof(ucfp_image_hashes) == 408 && sizeof(ucfp_image_desc) == 24 ? 0 : 1; }\n'
    exe = "/tmp/ucfp_hdr_check"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe],
                   input=src.encode(), check=True)
    assert subprocess.call([exe]) == 0


def test_plain_c_client_compiles_and_links_against_the_abi():
    """tests/c/abi_client.c -- a C11 program written against nothing but include/ucfp_cuda.h -- builds with -Wall -Werror
    and links against the .so (it RUNS on the GPU box: tests/test_host_layer_gpu.py)."""
    exe = "/tmp/ucfp_abi_client"
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_client.c"),
                    "-o", exe, "-L", libdir, "-lucfp_cuda", "-Wl,-rpath," + libdir], check=True)
    assert os.path.exists(exe)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ucfp_b200 import Context, UcfpError
    with pytest.raises(UcfpError) as e:
        Context(0)
    assert e.value.code == _ffi.E_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ucfp_b200/ or include/ may reference it."""
    bad = []
    for base in ("ucfp_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(import|from)\s+oracle\b|#include\s+\"[^\"]*oracle/", text, re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_exact_arithmetic_kernels_contain_no_fma():
    """docs/HASH_SPEC.md fixes separately rounded mul and add.  ptxas contracts mul.f32x2 + add.f32x2 into FFMA2
    even with -fmad=false, so the build is checked at the SASS level: no FFMA in the image-hash object."""
    obj = os.path.join(ROOT, "ucfp_b200", "csrc", "_obj", "image.o")
    if not os.path.exists(obj):
        from ucfp_b200 import build
        build.build(force=True)
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "FMUL" in sass and "FADD" in sass
    assert not re.search(r"\bFFMA2?\b", sass), "fused multiply-add found in image.o"

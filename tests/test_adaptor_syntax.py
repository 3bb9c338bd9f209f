"""The header-only ov_core adaptor (include/plviwo_ov_adaptor.hpp) is compiled downstream; here it is syntax-checked
against stand-in headers that declare the reference members it touches (tests/stubs/README.md)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adaptor_compiles_against_reference_shaped_headers():
    cmd = ["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "stubs"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "stubs", "adaptor_check.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_header_is_plain_c():
    """include/plviwo_fe.h must be consumable from C (cgo / JNI / ctypes style bindings)."""
    src = '#include "plviwo_fe.h"\nint main(void) { FeConfig c; plviwo_fe_default_config(&c); return (int)sizeof(FeLineRow) * 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

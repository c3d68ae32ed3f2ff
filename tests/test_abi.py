"""The C-ABI library loads and exports every symbol include/neurovit_b200.h declares, and the ctypes
signatures in neurovit_b200/_lib.py mirror the header (argument count and scalar/pointer kinds)."""
import ctypes
import os
import re

from neurovit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "neurovit_b200.h")


def _parse_header():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(nv_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        kinds = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds.append("p")
                elif a.startswith("int64_t"):
                    kinds.append("l")
                elif a.startswith("float"):
                    kinds.append("f")
                elif a.startswith("double"):
                    kinds.append("d")
                elif a.startswith("int"):
                    kinds.append("i")
                else:
                    raise AssertionError(f"unparsed argument {a!r} in {name}")
        protos[name] = kinds
    return protos


def test_header_declares_expected_entry_points():
    protos = _parse_header()
    assert "nv_last_error" in protos
    for name in _lib.SIGNATURES:
        assert name in protos, f"{name} bound in _lib.py but not declared in the header"
    for name in protos:
        assert name == "nv_last_error" or name in _lib.SIGNATURES, f"{name} declared but not bound"


def test_ctypes_signatures_match_header():
    kind = {ctypes.c_int: "i", ctypes.c_int64: "l", ctypes.c_float: "f", ctypes.c_double: "d", ctypes.c_void_p: "p"}
    protos = _parse_header()
    for name, argtypes in _lib.SIGNATURES.items():
        got = [kind[a] for a in argtypes]
        assert got == protos[name], f"{name}: ctypes {''.join(got)} != header {''.join(protos[name])}"


def test_library_loads_and_exports_all_symbols(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _parse_header():
        assert hasattr(lib, name), f"symbol {name} missing from {built_lib}"
    assert lib.nv_version() == 1
    lib.nv_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.nv_last_error(), bytes)

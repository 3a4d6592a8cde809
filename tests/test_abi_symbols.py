"""The drop-in boundary on a CPU-only box: libcrb200.so loads and exports every function include/crb.h
declares (no compute calls), the ctypes twin binds all of them, and the (uncompiled) Rust `extern "C"`
block of rust/cellular_raza-b200/src/ffi.rs names only declared functions with the declared arity."""
import ctypes as C
import re
from pathlib import Path

import pytest

from cellular_raza_b200 import _abi, build

ROOT = Path(__file__).resolve().parents[1]


def _header_functions():
    text = (ROOT / "include" / "crb.h").read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(crb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


@pytest.fixture(scope="module")
def lib():
    build.build()
    return C.CDLL(str(build.LIB))


def test_library_exports_every_declared_function(lib):
    fns = _header_functions()
    assert len(fns) >= 45, sorted(fns)
    missing = [n for n in fns if not hasattr(lib, n)]
    assert not missing, f"declared in include/crb.h but not exported: {missing}"
    assert lib.crb_abi_version() == 7


def test_ctypes_twin_binds_every_declared_function():
    bound = {row[0] for row in _abi._SIGNATURES} if hasattr(_abi, "_SIGNATURES") else None
    if bound is None:  # fall back to the source text of the binding table
        bound = set(re.findall(r'\("(crb_[a-z0-9_]+)"', (ROOT / "cellular_raza_b200" / "_abi.py").read_text()))
    fns = _header_functions()
    assert not (set(fns) - bound), f"not bound by _abi.py: {sorted(set(fns) - bound)}"
    assert not (bound - set(fns)), f"bound but not declared: {sorted(bound - set(fns))}"


def test_rust_extern_block_matches_header():
    text = (ROOT / "rust" / "cellular_raza-b200" / "src" / "ffi.rs").read_text()
    text = re.sub(r"//.*", "", text)
    fns = _header_functions()
    rust = {}
    for m in re.finditer(r"pub fn (crb_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", text, flags=re.S):
        args = m.group(2).strip()
        rust[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    assert len(rust) >= 35
    for name, arity in rust.items():
        assert name in fns, f"{name} is not declared in include/crb.h"
        assert arity == fns[name], f"{name}: rust binds {arity} arguments, crb.h declares {fns[name]}"
    # crb_config field order of the header == field order of the #[repr(C)] struct
    h = (ROOT / "include" / "crb.h").read_text()
    hs = h[h.index("typedef struct crb_config {"):h.index("} crb_config;")]
    hs = re.sub(r"/\*.*?\*/", " ", hs, flags=re.S)
    hfields = re.findall(r"\b(?:uint32_t|uint64_t|double)\s+([a-z_0-9]+)", hs)
    rs = text[text.index("pub struct crb_config {"):]
    rs = rs[:rs.index("}")]
    rfields = re.findall(r"pub ([a-z_0-9]+):", rs)
    assert hfields == rfields

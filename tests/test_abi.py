"""include/t3d.h <-> libt3d.so <-> the ctypes table: every declared symbol is exported and bound, and vice versa.
No compute calls (this runs without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "t3d.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(t3d_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_what_ctypes_binds():
    from tomography_3d_reconstructor_b200 import _lib
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from tomography_3d_reconstructor_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (t3d_[a-z0-9_]+)", out))
    for sym in declared_symbols():
        assert sym in exported, sym
        assert getattr(lib, sym) is not None
    assert exported - set(declared_symbols()) <= {"t3d_set_error"}
    assert lib.t3d_version() >= 100
    assert lib.t3d_words_per_row(1) == 4 and lib.t3d_words_per_row(128) == 4 and lib.t3d_words_per_row(129) == 8
    assert lib.t3d_words_per_row(1026) == 36


def test_argument_counts_match_header():
    from tomography_3d_reconstructor_b200 import _lib
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_res, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, txt, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing in the package may import, include, load or execute it."""
    pkg = os.path.join(ROOT, "tomography_3d_reconstructor_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include\s*[<\"].*oracle|libmc_ref|cpu_ref|oracle/_ref", re.M)
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), os.path.join(dirpath, f)

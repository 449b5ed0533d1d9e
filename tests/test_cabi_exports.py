"""CPU: the C-ABI library builds, loads, and exports every symbol that include/lc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lc_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from libcontinual_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lc_b200.h but not exported"
    lib.lc_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.lc_version()


def test_ctypes_table_matches_header():
    from libcontinual_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_product_path_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from libcontinual_b200 import _lib
    import libcontinual_b200.model as M
    with pytest.raises(_lib.LcError):
        M.cifar_resnet32()


def test_product_does_not_import_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "libcontinual_b200")):
        for f in fs:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(dp, f)).read(), flags=re.M):
                bad.append(f)
    assert not bad, bad

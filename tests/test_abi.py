"""The C-ABI library builds, loads and exports every symbol include/lcao_b200.h declares (no compute)."""
import ctypes
import os
import re

from lcaonet_b200 import _lib
from lcaonet_b200.csrc.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "lcao_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcao_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build())
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    lib.lcao_version.restype = ctypes.c_int
    assert lib.lcao_version() >= 100


def test_python_signatures_cover_the_header():
    declared = set(_declared()) - {"lcao_version", "lcao_last_error", "lcao_launch_count", "lcao_linear_bwd_scratch",
                                   "lcao_pair_contract_bwd_scratch"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    text = open(os.path.join(ROOT, "include", "lcao_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(argtypes), name


def test_basis_spec_layout_matches_header():
    # 4 int32 + 2 double + 3*18 int32 + 18 double + 18*8 double
    assert ctypes.sizeof(_lib.BasisSpec) == 16 + 16 + 3 * 18 * 4 + 18 * 8 + 18 * 8 * 8


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.lcao_twobody_fwd(None, 3, None, 10, 128, 3, 0, None, None)
    assert rc == -1 and b"null buffer" in lib.lcao_last_error()
    rc = lib.lcao_threebody_fwd(1, 3, 1, 1, 1, 8, 1, 1, 1, 1, 1, 4, 4, 130, 3, 1, None)
    assert rc == -1 and b"C % 4" in lib.lcao_last_error()

"""CPU: the C-ABI library loads, exports every symbol include/gim_b200.h declares, and the ctypes prototypes match the header."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from optimalstrategiesagainstgenerativeattacks_b200 import _cabi

HEADER = os.path.join(ROOT, "include", "gim_b200.h")


def parse_header():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(gim_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(2), m.group(3)
        codes = ""
        for a in [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]:
            if "*" in a or "gim_stream_t" in a:
                codes += "p"
            elif "long long" in a:
                codes += "l"
            elif re.match(r"(const\s+)?float\b", a):
                codes += "f"
            elif re.match(r"(const\s+)?int\b", a):
                codes += "i"
            else:
                raise AssertionError("unparsed parameter %r of %s" % (a, name))
        decls[name] = codes
    return decls


def test_header_matches_ctypes_prototypes():
    decls = parse_header()
    assert len(decls) >= 40
    for name, codes in _cabi.PROTOTYPES.items():
        assert decls.get(name) == codes, (name, decls.get(name), codes)
    assert set(decls) == set(_cabi.PROTOTYPES) | set(_cabi.OTHER_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_cabi.LIB_PATH)
    for name in parse_header():
        assert hasattr(L, name), name
    L.gim_version.restype = ctypes.c_int
    assert L.gim_version() >= 100


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing somewhere else."""
    import torch
    with pytest.raises(RuntimeError):
        _cabi.ptr(torch.zeros(4))

"""The C-ABI library builds, loads on a CPU-only box, exports every symbol the public header declares,
and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from matchmaker_b200 import _lib


def _header_symbols():
    syms = []
    inc = os.path.join(ROOT, "include")
    for f in sorted(os.listdir(inc)):
        if f.endswith(".h"):
            txt = open(os.path.join(inc, f)).read()
            syms += re.findall(r"MMB200_API\s+[\w\s\*]+?\b(mmb200_\w+)\s*\(", txt)
    return sorted(set(syms))


def test_library_is_built():
    assert os.path.isfile(_lib.LIB_PATH), "run `python -m matchmaker_b200.build` (or __graft_entry__.build())"


def test_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 6
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES out of sync with include/matchmaker_b200.h"


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.mmb200_version() == 100
    assert isinstance(_lib.last_error(), str)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu():
    from matchmaker_b200 import interaction
    q = torch.zeros(1, 4, 64, dtype=torch.float16)
    with pytest.raises(_lib.MatchmakerB200Error):
        interaction.maxsim(q, q)  # CPU tensors are rejected: there is no CPU fallback
    lib = _lib.load()
    rc = lib.mmb200_device_info(-1, None, None, None)
    assert rc == _lib.ERR_CUDA and "CUDA" in _lib.last_error()

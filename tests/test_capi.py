"""CPU tests of the boundary: the library loads, exports every declared symbol, and refuses to
compute without a GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from spring_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spring_b200.h")).read()
    return sorted(set(re.findall(r"\b(spring_b200_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    lib = capi.load()
    syms = declared_symbols()
    assert set(syms) == set(capi.EXPORTS), "capi.EXPORTS out of sync with include/spring_b200.h"
    for s in syms:
        assert hasattr(lib, s), f"libspring_b200.so does not export {s}"


def test_version_and_struct_sizes():
    lib = capi.load()
    assert b"sm_100a" in lib.spring_b200_version()
    assert ctypes.sizeof(capi.CP) == 64          # compression_params, util.h:30-51
    assert capi.CP.num_reads.offset == 28 and capi.CP.max_readlen.offset == 40 and capi.CP.num_thr.offset == 56


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly():
    with pytest.raises(capi.SpringB200Error) as e:
        capi.Context(0)
    assert e.value.code == -2

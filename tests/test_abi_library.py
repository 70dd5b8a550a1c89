"""The C-ABI product library: it loads, exports every symbol include/demcmc_b200.h declares, and
refuses to work without a CUDA device (no CPU fallback).  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

import common
from common import D

HEADER = os.path.join(common.ROOT, "include", "demcmc_b200.h")


def _build():
    import __graft_entry__ as g
    if not os.path.exists(common.CUDA_LIB):
        g.build()


def test_library_exports_every_declared_symbol():
    _build()
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(demcmc_[a-z0-9_]+)\s*\(", src))
    assert declared == set(D._ffi.SYMBOLS), declared ^ set(D._ffi.SYMBOLS)
    L = ctypes.CDLL(common.CUDA_LIB)
    for s in declared:
        assert hasattr(L, s), s
    L.demcmc_backend_name.restype = ctypes.c_char_p
    assert L.demcmc_backend_name() == b"cuda-sm100a"
    assert L.demcmc_abi_version() == D._ffi.ABI_VERSION


def test_library_is_sm100a_only():
    _build()
    out = subprocess.run(["cuobjdump", "-lelf", common.CUDA_LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    D._ffi.use_library(common.CUDA_LIB)
    try:
        with pytest.raises(D._ffi.DemcmcError) as e:
            D.Handle(4, 6, 2, [-1, 0], [1, 1])
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)
        with pytest.raises(D._ffi.DemcmcError, match="no CPU fallback"):
            D.op_project([1.0, 2.0], [3.0, 4.0])
    finally:
        D._ffi._lib = None   # later tests bind what they need


def test_product_package_never_references_the_oracle_or_the_test_double():
    pkg = os.path.join(common.ROOT, "differentialevolutionmcmc.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), (f, "mentions the oracle")
                assert "libdemcmc_emu" not in text, f

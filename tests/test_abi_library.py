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


def test_persistent_kernel_keeps_its_register_budget():
    """k_chunk_persist raises the register budget of its DMMA warps with setmaxnreg (216) because a tile's B
    fragments live in registers; ptxas honours that only for some code shapes and otherwise
    quietly allocates the region within the launch's 128 registers, reloading B fragments from local
    memory inside the DMMA loop.  The SASS must show registers above 200 in use, the tensor-path and TMA
    instructions the design rests on, and no spill traffic next to a DMMA."""
    _build()
    sass = subprocess.run(["cuobjdump", "-sass", common.CUDA_LIB], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(sass) if "Function :" in l]
    fn = [i for i in starts if "k_chunk_persistILb0" in sass[i]]
    assert fn, "k_chunk_persist<false> not in the library"
    end = min([i for i in starts if i > fn[0]] + [len(sass)])
    seg = [l for l in sass[fn[0]:end] if "/*" in l and not l.strip().startswith("/* 0x")]
    regs = max(int(m) for l in seg for m in re.findall(r"\bR(\d+)\b", l))
    assert regs >= 200, regs
    text = "\n".join(seg)
    for mnemonic in ("DMMA.8x8x4", "UBLKCP", "SYNCS", "USETMAXREG", "CCTL.IVALL", "ATOMG.E.ADD.64", "REDG.E.ADD.S32.STRONG.GPU"):
        assert mnemonic in text, mnemonic
    spill = [i for i, l in enumerate(seg) if re.search(r"\b(LDL|STL)\b", l)]
    near = [i for i in spill if any("DMMA" in seg[j] for j in range(max(0, i - 6), min(len(seg), i + 7)))]
    assert not near, f"{len(near)} local-memory accesses within 6 instructions of a DMMA"

"""Reader / writer of the on-disk replay tape julia/record_tape.jl produces (format demcmc_tape_v1): one raw
little-endian .bin per array + manifest.txt with C-order shapes.  `load_tape` returns what Handle.replay and the
parity checks need; `save_tape` writes the same format from an oracle run (used to test the reader here, where no
Julia exists to produce a real one)."""
from __future__ import annotations

import os

import numpy as np

_JL = {"Float64": "<f8", "Int32": "<i4", "UInt8": "u1", "Int64": "<i8"}
TAPE_FIELDS = ("mig_u", "mig_n", "mig_groups", "mig_pick_u", "kind", "idx", "idx_row", "gamma1", "gamma2", "u_acc", "noise", "keep")


def load_tape(path):
    meta, arrays = {}, {}
    with open(os.path.join(path, "manifest.txt")) as f:
        lines = [ln.split() for ln in f if ln.strip()]
    if lines[0] != ["format", "demcmc_tape_v1"]:
        raise ValueError(f"{path}: not a demcmc_tape_v1 manifest")
    for ln in lines[1:]:
        if ln[0] == "array":
            name, dt, shape = ln[1], _JL[ln[2]], tuple(int(v) for v in ln[3:])
            a = np.fromfile(os.path.join(path, name + ".bin"), dtype=dt)
            arrays[name] = a.reshape(shape) if a.size else a
        else:
            for k, v in zip(ln[::2], ln[1::2]):
                meta[k] = float(v) if "." in v or "e" in v.lower() else int(v)
    tape = {k: arrays[k] for k in TAPE_FIELDS if k in arrays}
    if meta.get("kappa", 1.0) == 1.0:
        tape["keep"] = None
    if not (tape.get("idx_row") is not None and (tape["idx_row"] >= 0).any()):
        tape["idx_row"] = None
    return meta, tape, arrays


def save_tape(path, meta, arrays):
    os.makedirs(path, exist_ok=True)
    inv = {np.dtype(v).str.lstrip("<|"): k for k, v in _JL.items()}
    with open(os.path.join(path, "manifest.txt"), "w") as f:
        f.write("format demcmc_tape_v1\n")
        f.write(" ".join(f"{k} {v}" for k, v in meta.items()) + "\n")
        for name, a in arrays.items():
            if a is None:
                continue
            a = np.ascontiguousarray(a)
            f.write(f"array {name} {inv[a.dtype.str.lstrip('<|')]} {' '.join(str(n) for n in a.shape)}\n")
            a.tofile(os.path.join(path, name + ".bin"))

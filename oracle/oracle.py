"""ctypes binding of the CPU oracle (oracle/demcmc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the product package.  See oracle/demcmc_oracle.h for the parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = {"gaussian": 0, "mvnormal": 1, "binomial": 2, "lnr": 3, "lba": 4, "hier_normal": 5, "rastrigin": 6, "mvnormal_full": 7}
UPDATES = {"mh": 0, "maximize": 1, "minimize": 2}
FITNESS = {"posterior": 0, "fun": 1}
PRIORS = {"flat": 0, "normal": 1, "halfcauchy": 2, "uniform": 3, "beta": 4, "normal_ref": 5}
PROPOSALS = {"random_gamma": 0, "fixed_gamma": 1, "variable_gamma": 2}
KIND_DE, KIND_SNOOKER, KIND_MUTATION = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class _Prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ref", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


class _Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d", C.c_int32), ("n_obs", C.c_int64), ("n_dim", C.c_int32),
                ("n_per", C.c_int32), ("x", _dp), ("choice", _ip), ("sigma", _dp),
                ("lba_floor", C.c_double), ("prior", C.POINTER(_Prior)), ("cov", _dp)]


class _Config(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("Np", C.c_int32), ("d", C.c_int32), ("burnin", C.c_int32),
                ("n_initial", C.c_int32), ("alpha", C.c_double), ("beta", C.c_double), ("eps", C.c_double),
                ("sigma", C.c_double), ("kappa", C.c_double), ("theta_snooker", C.c_double),
                ("proposal", C.c_int32), ("n_blocks", C.c_int32), ("blocks", _bp), ("lo", _dp), ("hi", _dp),
                ("base_snapshot", C.c_int32), ("n_threads", C.c_int32), ("seed", C.c_uint64),
                ("resample", C.c_int32), ("update", C.c_int32), ("fitness", C.c_int32), ("reserved", C.c_int32),
                ("block_on", _bp), ("n_block_on", C.c_int64)]


_TAPE_FIELDS = [("mig_u", "f8"), ("mig_n", "i4"), ("mig_groups", "i4"), ("mig_pick_u", "f8"), ("mig_slots", "i4"),
                ("mut_u", "f8"), ("kind", "u1"), ("idx", "i4"), ("u_snk", "f8"), ("u_base", "f8"),
                ("gamma1", "f8"), ("gamma2", "f8"), ("u_acc", "f8"), ("noise", "f8"), ("keep", "u1"), ("idx_row", "i4")]
_TRACE_FIELDS = [("prop_theta", "f8"), ("prop_weight", "f8"), ("log_adj", "f8"), ("accepted", "u1"),
                 ("state_theta", "f8"), ("state_weight", "f8"), ("state_id", "i4"),
                 ("pre_theta", "f8"), ("pre_weight", "f8"), ("pre_id", "i4")]
_CT = {"f8": _dp, "i4": _ip, "u1": _bp}


class _Tape(C.Structure):
    _fields_ = [(n, _CT[t]) for n, t in _TAPE_FIELDS]


class _Trace(C.Structure):
    _fields_ = [(n, _CT[t]) for n, t in _TRACE_FIELDS]


def build(force: bool = False) -> str:
    """Compile oracle/libdemcmc_oracle.so with gcc (Makefile in this directory)."""
    so = os.path.join(_HERE, "libdemcmc_oracle.so")
    src = os.path.join(_HERE, "demcmc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_run.restype = C.c_int
        L.orc_loglike.restype = C.c_double
        L.orc_prior_loglike.restype = C.c_double
        L.orc_posterior.restype = C.c_double
        L.orc_adjust_loglike.restype = C.c_double
        L.orc_accept.restype = C.c_int
        L.orc_accept.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_select_base.restype = C.c_int
        L.orc_select_base.argtypes = [_dp, C.c_int, C.c_double]
        L.orc_select_particle.restype = C.c_int
        L.orc_select_particle.argtypes = [_dp, C.c_int, C.c_double, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def _ptr(a, ct):
    return a.ctypes.data_as(ct) if a is not None else ct()


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Model:
    """Keeps the numpy buffers alive behind an orc_model."""

    def __init__(self, kind, d, prior, x=None, choice=None, n_dim=0, n_per=0, sigma=None, lba_floor=1e-10, cov=None):
        self.kind, self.d = kind, int(d)
        self.x = _f8(x if x is not None else [])
        self.choice = None if choice is None else np.ascontiguousarray(choice, dtype=np.int32)
        self.sigma = None if sigma is None else _f8(sigma)
        assert len(prior) == d, (len(prior), d)
        self._prior = (_Prior * d)()
        for k, p in enumerate(prior):
            name = p[0]
            a = float(p[1]) if len(p) > 1 else 0.0
            b = float(p[2]) if len(p) > 2 else 0.0
            ref = int(p[3]) if len(p) > 3 else 0
            self._prior[k] = _Prior(PRIORS[name], ref, a, b)
        self.cov = None if cov is None else _f8(cov)
        if kind in ("mvnormal", "mvnormal_full"):
            n_obs = self.x.shape[0]
            n_dim = self.x.shape[1]
        elif kind == "hier_normal":
            n_dim, n_per = self.x.shape
            n_obs = n_dim * n_per
        elif kind == "binomial":
            n_obs = 1
        else:
            n_obs = self.x.shape[0]
        self.c = _Model(KINDS[kind], self.d, n_obs, int(n_dim), int(n_per), _ptr(self.x, _dp),
                        _ptr(self.choice, _ip), _ptr(self.sigma, _dp), float(lba_floor), self._prior, _ptr(self.cov, _dp))


class Config:
    def __init__(self, n_groups, Np, d, lo, hi, burnin=1000, n_initial=0, alpha=0.1, beta=0.1, eps=0.001,
                 sigma=0.05, kappa=1.0, theta_snooker=0.0, proposal="random_gamma", blocks=None,
                 base_snapshot=0, n_threads=1, seed=0, resample=False, update="mh", fitness="posterior", blocking_schedule=None):
        self.lo, self.hi = _f8(lo), _f8(hi)
        assert self.lo.shape == (d,) and self.hi.shape == (d,)
        self.blocks = None if blocks is None else np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, d)
        nb = 0 if self.blocks is None else self.blocks.shape[0]
        if n_groups == 1:
            alpha = 0.0  # structs.jl:102-105
        self.c = _Config(n_groups, Np, d, burnin, n_initial, alpha, beta, eps, sigma, kappa, theta_snooker,
                         PROPOSALS[proposal], nb, _ptr(self.blocks, _bp), _ptr(self.lo, _dp), _ptr(self.hi, _dp),
                         int(base_snapshot), int(n_threads), int(seed), int(bool(resample)), UPDATES[update], FITNESS[fitness], 0)
        # blocking_on(de) per iteration (main.jl:137,162); None = on in every iteration (when there are blocks)
        self.blocking_schedule = None if blocking_schedule is None else np.ascontiguousarray(blocking_schedule, dtype=np.uint8)
        if self.blocking_schedule is not None:
            self.c.block_on = _ptr(self.blocking_schedule, _bp)
            self.c.n_block_on = self.blocking_schedule.size
        self.resample = bool(resample)
        self.n_groups, self.Np, self.d, self.n_initial = n_groups, Np, d, n_initial
        self.B = max(1, nb)
        self.P = n_groups * Np


def tape_shapes(cfg: Config, n_iter: int):
    G, P, d, S = cfg.n_groups, cfg.P, cfg.d, n_iter * cfg.B
    return {"mig_u": (n_iter,), "mig_n": (n_iter,), "mig_groups": (n_iter, G), "mig_pick_u": (n_iter, G),
            "mig_slots": (n_iter, G), "mut_u": (S, G), "kind": (S, P), "idx": (S, P, 3), "u_snk": (S, P),
            "u_base": (S, P), "gamma1": (S, P), "gamma2": (S, P), "u_acc": (S, P), "noise": (S, P, d),
            "keep": (S, P, d), "idx_row": (S, P, 3)}


def trace_shapes(cfg: Config, n_iter: int):
    P, d, S = cfg.P, cfg.d, n_iter * cfg.B
    return {"prop_theta": (S, P, d), "prop_weight": (S, P), "log_adj": (S, P), "accepted": (S, P),
            "state_theta": (n_iter, P, d), "state_weight": (n_iter, P), "state_id": (n_iter, P),
            "pre_theta": (n_iter, P, d), "pre_weight": (n_iter, P), "pre_id": (n_iter, P)}


def _alloc(fields, shapes):
    return {n: np.zeros(shapes[n], dtype=t) for n, t in fields}


def _pack(struct_cls, fields, arrays):
    s = struct_cls()
    for n, t in fields:
        a = None if arrays is None else arrays.get(n)
        if a is not None:
            assert a.flags["C_CONTIGUOUS"] and a.dtype == np.dtype(t), n
        setattr(s, n, _ptr(a, _CT[t]))
    return s


def run(cfg: Config, model: Model, theta0, n_iter, tape_in=None, record=True, trace=True, history=True, init_rows=None):
    """Run the oracle.  Returns a dict with samples/accept/lp (reference layout), final state,
    and optionally the recorded tape and the per-sweep trace."""
    L = lib()
    P, d = cfg.P, cfg.d
    theta0 = _f8(theta0).reshape(P, d)
    n_rows = n_iter + cfg.n_initial
    out = {}
    if history:
        # Julia Array{T,3}(n_rows, d, P): row fastest => numpy shape (P, d, n_rows) C-order
        out["samples"] = np.zeros((P, d, n_rows))
        if cfg.n_initial > 0:
            # initialize_samples (utilities.jl:29-41): init_rows[i][id][k] are the caller's prior draws
            out["samples"][:, :, :cfg.n_initial] = _f8(init_rows).reshape(cfg.n_initial, P, d).transpose(1, 2, 0)
        out["accept"] = np.zeros((P, n_rows), dtype=np.uint8)
        out["lp"] = np.zeros((P, n_rows))
    out["final_id"] = np.zeros(P, dtype=np.int32)
    out["final_theta"] = np.zeros((P, d))
    out["final_weight"] = np.zeros(P)
    tape_out = _alloc(_TAPE_FIELDS, tape_shapes(cfg, n_iter)) if record else None
    if tape_out is not None and cfg.c.kappa == 1.0:
        tape_out["keep"] = None
    if tape_out is not None and not cfg.resample:
        tape_out["idx_row"] = None
    tr = _alloc(_TRACE_FIELDS, trace_shapes(cfg, n_iter)) if trace else None
    tin = _pack(_Tape, _TAPE_FIELDS, tape_in) if tape_in is not None else None
    tout = _pack(_Tape, _TAPE_FIELDS, tape_out) if record else None
    trs = _pack(_Trace, _TRACE_FIELDS, tr) if trace else None
    rc = L.orc_run(C.byref(cfg.c), C.byref(model.c), _ptr(theta0, _dp), C.c_int64(n_iter),
                   C.byref(tin) if tin is not None else None, C.byref(tout) if tout is not None else None,
                   C.byref(trs) if trs is not None else None,
                   _ptr(out.get("samples"), _dp), _ptr(out.get("accept"), _bp), _ptr(out.get("lp"), _dp),
                   _ptr(out["final_id"], _ip), _ptr(out["final_theta"], _dp), _ptr(out["final_weight"], _dp))
    if rc != 0:
        raise RuntimeError(f"orc_run failed: {rc}")
    out["tape"] = tape_out
    out["trace"] = tr
    return out


def set_plain_sums(on: bool):
    """bench.py's CPU arm only: uncompensated sums in the MVN likelihood (the speed of a straightforward CPU code)"""
    lib().orc_set_plain_sums(int(bool(on)))


def loglike(model: Model, theta):
    th = _f8(theta)
    return lib().orc_loglike(C.byref(model.c), _ptr(th, _dp))


def prior_loglike(model: Model, theta):
    th = _f8(theta)
    return lib().orc_prior_loglike(C.byref(model.c), _ptr(th, _dp))


def posterior(cfg: Config, model: Model, theta):
    th = _f8(theta)
    return lib().orc_posterior(C.byref(cfg.c), C.byref(model.c), _ptr(th, _dp))


def project(p1, p2):
    p1, p2 = _f8(p1), _f8(p2)
    out = np.zeros_like(p1)
    lib().orc_project(_ptr(p1, _dp), _ptr(p2, _dp), C.c_int(p1.size), _ptr(out, _dp))
    return out


def adjust_loglike(pt, prop, pz):
    pt, prop, pz = _f8(pt), _f8(prop), _f8(pz)
    return lib().orc_adjust_loglike(_ptr(pt, _dp), _ptr(prop, _dp), _ptr(pz, _dp), C.c_int(pt.size))


def reset(prop, pt, mask):
    prop, pt = _f8(prop).copy(), _f8(pt)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    lib().orc_reset(_ptr(prop, _dp), _ptr(pt, _dp), _ptr(mask, _bp), C.c_int(prop.size))
    return prop


def de_proposal(pt, pm, pn, pb, g1, g2, b):
    pt, pm, pn, b = _f8(pt), _f8(pm), _f8(pn), _f8(b)
    pbp = _ptr(_f8(pb), _dp) if pb is not None else _dp()
    out = np.zeros_like(pt)
    lib().orc_de_proposal(_ptr(pt, _dp), _ptr(pm, _dp), _ptr(pn, _dp), pbp, C.c_double(g1), C.c_double(g2),
                          _ptr(b, _dp), C.c_int(pt.size), _ptr(out, _dp))
    return out


def snooker_proposal(pt, pz, pm, pn, g, b):
    pt, pz, pm, pn, b = _f8(pt), _f8(pz), _f8(pm), _f8(pn), _f8(b)
    out = np.zeros_like(pt)
    lib().orc_snooker_proposal(_ptr(pt, _dp), _ptr(pz, _dp), _ptr(pm, _dp), _ptr(pn, _dp), C.c_double(g),
                               _ptr(b, _dp), C.c_int(pt.size), _ptr(out, _dp))
    return out


def accept(w_prop, w_cur, log_adj, u):
    return bool(lib().orc_accept(w_prop, w_cur, log_adj, u))


def select_base(w, u):
    w = _f8(w)
    return lib().orc_select_base(_ptr(w, _dp), w.size, u)


def select_particle(w, u):
    w = _f8(w)
    drew = C.c_int(0)
    i = lib().orc_select_particle(_ptr(w, _dp), w.size, u, C.byref(drew))
    return i, bool(drew.value)


def shift(tags, groups, slots, Np):
    tags = np.ascontiguousarray(tags, dtype=np.int32).copy()
    groups = np.ascontiguousarray(groups, dtype=np.int32)
    slots = np.ascontiguousarray(slots, dtype=np.int32)
    lib().orc_shift(_ptr(tags, _ip), _ptr(groups, _ip), _ptr(slots, _ip), C.c_int(groups.size), C.c_int(Np))
    return tags


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(c, k, o)
    return [int(v) for v in o]


def uniform2(seed, stream, sweep, unit, k):
    u = (C.c_double * 2)()
    lib().orc_uniform2(C.c_uint64(seed), C.c_uint32(stream), C.c_uint32(sweep), C.c_uint32(unit), C.c_uint32(k), u)
    return float(u[0]), float(u[1])

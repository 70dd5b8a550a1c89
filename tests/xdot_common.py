"""Pins the arithmetic of the streamed likelihood kernels (k_xdot, k_chunk_persist) -- shared by the GPU
tests, the CPU tests of the host-only test double and the mutation-test subprocess.

The MVN / hierarchical log-likelihood is evaluated as  sum x'^2 - 2 B + n sum m'^2  with the data centred
on `center`; B = sum_i sum_k x'_ik m'_pk is the ONLY thing the O(N d) kernels compute.  With the default
centre (the column means) B is analytically zero and no likelihood value can tell whether its operands
were right (What's weak 1), so these checks centre the data on a GIVEN vector
(demcmc_model.center) and compare B itself with an extended-precision reference:
    B_p = sum_k (sum_i x'_ik) m'_pk            (exact algebra; the column sums in longdouble)
to within 2^-40 of the Cauchy-Schwarz bound sum_i |x'_i| |m'_p| (the kernel rounds each two-row chain to a
fixed-point grid 2^-44..2^-50 below that bound, de_math.h: xd_scale)."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demcmc_b200 as D  # noqa: E402

TOL = 2.0 ** -40


def make(kind, n, k, rng, offset=0.4):
    """data, prior table, bounds, a sampler of parameter vectors and the centre override for one shape:
    kind mvnormal: n observations x k dimensions; hier_normal: k subjects x n observations each."""
    if kind == "mvnormal":
        mu = rng.normal(size=k)
        x = rng.normal(mu, 1.0, size=(n, k))
        col_mean = x.mean(axis=0)
        d = k + 1
        prior = [("normal", 0, 1)] * k + [("halfcauchy", 0, 1)]
        lo, hi = [-np.inf] * k + [0.0], [np.inf] * d

        def draw(P):
            return np.column_stack([rng.normal(mu, 0.3, size=(P, k)), rng.uniform(0.7, 1.4, P)])
    else:
        b0 = rng.normal(0, 1, k)
        x = rng.normal(1.0 + b0[:, None], 0.5, size=(k, n))
        col_mean = x.mean(axis=1)
        d = k + 3
        prior = [("normal", 1, 1), ("halfcauchy", 0, 1)] + [("normal_ref", 0, 0, 1)] * k + [("halfcauchy", 0, 1)]
        lo, hi = [-np.inf, 0.0] + [-np.inf] * k + [0.0], [np.inf] * d

        def draw(P):
            return np.column_stack([rng.normal(1, 0.2, P), rng.uniform(0.6, 1.5, P), rng.normal(b0, 0.3, size=(P, k)), rng.uniform(0.4, 0.9, P)])
    center = col_mean + offset * (1.0 + rng.uniform(size=k))       # deliberately NOT the column mean
    return dict(kind=kind, x=x, d=d, k=k, prior=prior, lo=lo, hi=hi, draw=draw, center=center)


def means_of(case, theta):
    th = np.asarray(theta, dtype=np.longdouble).reshape(-1, case["d"])
    k = case["k"]
    return th[:, :k] if case["kind"] == "mvnormal" else th[:, :1] + th[:, 2:2 + k]


def reference(case, theta, center=None):
    """(B_ref, bound) per parameter vector, longdouble."""
    c = np.asarray(case["center"] if center is None else center, dtype=np.longdouble)
    x = np.asarray(case["x"], dtype=np.longdouble)
    xc = x - c if case["kind"] == "mvnormal" else (x - c[:, None]).T          # [obs][dim]
    colsum = xc.sum(axis=0)
    rownorm = np.sqrt((xc * xc).sum(axis=1)).sum()
    mp = means_of(case, theta) - c
    B = mp @ colsum
    bound = rownorm * np.sqrt((mp * mp).sum(axis=1))
    return B, bound


def handle(case, G, Np, center="given", **kw):
    h = D.Handle(G, Np, case["d"], case["lo"], case["hi"], **kw)
    h.set_model(case["kind"], case["prior"], x=case["x"], center=case["center"] if center == "given" else None)
    return h


def worst(B, ref, bound):
    """max |B - ref| / bound"""
    return float(np.max(np.abs(np.asarray(B, dtype=np.longdouble) - ref) / bound))


def eval_error(kind, n, k, P, seed=0):
    """error of demcmc_eval_xdot (k_stage_means + k_xdot) on P parameter vectors, in units of the bound"""
    rng = np.random.default_rng(seed)
    case = make(kind, n, k, rng)
    th = case["draw"](P)
    with handle(case, 1, max(P, 3)) as h:
        B = h.eval_xdot(th)
        ll, _ = h.eval(th)
    ref, bound = reference(case, th)
    return worst(B, ref, bound), ll, case, th


def run_error(kind, n, k, G, Np, n_iter, seed=0, **kw):
    """error of the cross term of every proposal of a native run (k_propose staging + k_xdot, or the persistent
    chunk kernel), in units of the bound, and the level sizes the run went through"""
    rng = np.random.default_rng(seed)
    case = make(kind, n, k, rng)
    th0 = case["draw"](G * Np)
    with handle(case, G, Np, trace=True, seed=seed + 1, burnin=2, **kw) as h:
        h.set_state(th0)
        h.run(n_iter)
        tr = h.trace()
        B = h.trace_xdot()
        ctr = h.counters()
    prop = tr["prop_theta"].reshape(-1, case["d"])
    ref, bound = reference(case, prop)
    fin = np.isfinite(np.asarray(ref, dtype=float)) & (np.asarray(bound, dtype=float) > 0)
    return worst(B.reshape(-1)[fin], ref[fin], bound[fin]), ctr, int(fin.sum())


if __name__ == "__main__":      # the mutation-test subprocess: exit 0 = parity holds, 3 = parity violated
    import sys
    D._ffi.use_library(sys.argv[1])
    e1 = eval_error("mvnormal", 130, 50, 33)[0]
    e2 = run_error("mvnormal", 130, 50, 2, 24, 3, theta_snooker=0.2)[0]
    print(f"eval {e1:.3e} run {e2:.3e} tol {TOL:.3e}")
    sys.exit(0 if max(e1, e2) <= TOL else 3)

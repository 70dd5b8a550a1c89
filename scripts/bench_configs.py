#!/usr/bin/env python
"""Throughput of the population step on synthetic data of EVERY BASELINE.json config shape, one
GPU (bench.py measures the headline config, configs[1], under the driver's contract; this script
is the companion table for DESIGN.md / profiles/).  One JSON line per config:

  C1 Gaussian_Example   2 parameters, 50 obs, 4 groups x 6 particles (the reference's CPU case)
  C2 MVN d=50           1e5 obs, 4 x 256, crossover + snooker
  C3 LBA                5 parameters, 1e5 trials, 4 x 256 (group sizes assumed)
  C4 hierarchical       1000 subjects x 50 obs, blocks, 16 x 512
  C5 MVN d=100          1e5 obs, 8 groups x 4096 per GPU (the per-GPU shard of 64 x 4096 on 8 GPUs)

Usage: python scripts/bench_configs.py [c1 c2 ...] [--iters N]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import demcmc_b200 as D  # noqa: E402

INF = np.inf


def lba_sim_vec(rng, n, nu=(3.0, 2.0), A=0.8, k=0.2, tau=0.3):
    nu = np.asarray(nu)
    v = rng.normal(nu, 1.0, size=(n, nu.size))
    bad = ~(v > 0).any(axis=1)
    while bad.any():
        v[bad] = rng.normal(nu, 1.0, size=(int(bad.sum()), nu.size))
        bad = ~(v > 0).any(axis=1)
    a = rng.uniform(0, A, size=(n, nu.size))
    t = np.where(v > 0, (A + k - a) / np.where(v > 0, v, 1.0), np.inf)
    return (np.argmin(t, axis=1) + 1).astype(np.int32), tau + t.min(axis=1)


def config(name):
    if name == "c1":
        rng = np.random.default_rng(50514)
        x = rng.normal(0, 1, 50)
        return dict(kind="gaussian", G=4, Np=6, d=2, prior=[("normal", 0, 1), ("halfcauchy", 0, 1)], lo=[-INF, 0], hi=[INF, INF],
                    data=dict(x=x), theta0=lambda r, P: np.column_stack([r.normal(size=P), np.abs(r.standard_cauchy(P)) + 0.1]),
                    kw=dict(burnin=1000), iters=2000, flops=4.0 * 50, what="Gaussian_Example: 2-parameter Normal, 50 obs, 4 x 6")
    if name in ("c2", "c5"):
        dm, G, Np = (50, 4, 256) if name == "c2" else (100, 8, 4096)
        rng = np.random.default_rng(50514)
        mu = rng.normal(size=dm)
        x = rng.normal(mu, 1.0, size=(100_000, dm))
        return dict(kind="mvnormal", G=G, Np=Np, d=dm + 1, prior=[("normal", 0, 1)] * dm + [("halfcauchy", 0, 1)], lo=[-INF] * dm + [0.0],
                    hi=[INF] * (dm + 1), data=dict(x=x),
                    theta0=lambda r, P: np.column_stack([r.normal(size=(P, dm)), np.abs(r.standard_cauchy(P)) + 0.1]),
                    kw=dict(burnin=0, theta_snooker=0.1), iters=100 if name == "c2" else 12, flops=2.0 * 100_000 * dm,
                    what=f"isotropic MVN d={dm}, 1e5 obs, {G} x {Np}, crossover + snooker 0.1" + (" (the per-GPU shard of configs[4])" if name == "c5" else ""))
    if name == "c3":
        rng = np.random.default_rng(88484)
        choice, rt = lba_sim_vec(rng, 100_000)
        mn = float(rt.min())
        prior = [("normal", 1, 5), ("normal", 1, 5), ("normal", 0.8, 0.2), ("normal", 0.2, 0.1), ("uniform", 0, mn)]
        return dict(kind="lba", G=4, Np=256, d=5, prior=prior, lo=[0, 0, 0, 0, 0], hi=[INF, INF, INF, INF, mn], data=dict(x=rt, choice=choice, n_dim=2),
                    theta0=lambda r, P: np.column_stack([np.abs(r.normal(1, 5, P)), np.abs(r.normal(1, 5, P)), np.abs(r.normal(0.8, 0.2, P)),
                                                         np.abs(r.normal(0.2, 0.1, P)), r.uniform(0, mn, P)]),
                    kw=dict(burnin=0), iters=60, flops=None, what="LBA (Run_LBA): 5 parameters, 1e5 trials, 4 x 256")
    if name == "c4":
        rng = np.random.default_rng(9528)
        S, n = 1000, 50
        b0 = rng.normal(0, 1, S)
        y = rng.normal(1.0 + b0[:, None], 0.5, size=(S, n))
        prior = [("normal", 1, 1), ("halfcauchy", 0, 1)] + [("normal_ref", 0, 0, 1)] * S + [("halfcauchy", 0, 1)]
        blocks = np.array([[1, 1] + [0] * S + [1], [0, 0] + [1] * S + [0]], dtype=np.uint8)

        def th0(r, P):
            sb = np.abs(r.standard_cauchy(P)) + 0.2
            return np.column_stack([r.normal(1, 1, P), sb, r.normal(0, 1, (P, S)) * sb[:, None], np.abs(r.standard_cauchy(P)) + 0.2])
        return dict(kind="hier_normal", G=16, Np=512, d=S + 3, prior=prior, lo=[-INF, 0] + [-INF] * S + [0], hi=[INF] * (S + 3), data=dict(x=y),
                    theta0=th0, kw=dict(burnin=0, blocks=blocks), iters=30, flops=2.0 * S * n,
                    what="hierarchical normal: 1000 subjects x 50 obs, 2 parameter blocks (blocking_on), 16 x 512")
    raise KeyError(name)


# fp64 instructions (DFMA + DMUL + DADD, thread level) the LBA kernel executes per (trial x particle) density, from the
# committed ncu capture profiles/r02_v8_k_ll_pointwise_lba_ncu_full_summary.csv: 742.9 M warp instructions for a level of
# 344 particles x 1e5 trials, 48.6 % of them fp64 => 336 per density.  The fp64 CUDA-core pipe retires peak_dfma / 2
# of them per second (one DFMA = 2 flop), which is the denominator of the instruction roofline below.
LBA_FP64_INST_PER_DENSITY = 336.0


def run(name, iters=None, peaks=None, hbm_gbs=None):
    c = config(name)
    n_iter = iters or c["iters"]
    P = c["G"] * c["Np"]
    h = D.Handle(c["G"], c["Np"], c["d"], c["lo"], c["hi"], seed=20261017, **c["kw"])
    h.set_model(c["kind"], c["prior"], **c["data"])
    h.set_state(c["theta0"](np.random.default_rng(1), P))
    h.run(max(3, n_iter // 10))                       # warm-up
    c0 = h.counters()
    t0 = time.perf_counter()
    h.run(n_iter)
    wall = time.perf_counter() - t0
    c1 = h.counters()
    updates = c1["particle_updates"] - c0["particle_updates"]
    ms = c1["device_ms"]
    acc = h.accept()[:, -n_iter:].mean()
    # the likelihood kernel by itself (MVN / hierarchical): a second pass with CUDA events around every
    # launch of the dominant kernel (k_xdot per level, or k_chunk_persist per chunk)
    ll = None
    if c["flops"]:
        h.set_timing(0, True)
        h.run(n_iter)
        c2 = h.counters()
        if c2["loglike_ms"] > 0:
            dfma, dmma = peaks if peaks else D.fp64_peaks(0)
            tf = c["flops"] * (c2["particle_updates"] - c1["particle_updates"]) / (c2["loglike_ms"] * 1e-3) / 1e12
            persistent = c2["persistent_chunks"] > c1["persistent_chunks"]
            ll = {"kernel": "k_chunk_persist (whole chunk: proposals and accepts included)" if persistent else ("k_xdot" if c["kind"] in ("mvnormal", "hier_normal", "mvnormal_full") else "k_ll_pointwise"),
                  "tflops_event_bracketed": tf, "of_measured_dmma_peak": tf / max(dfma, dmma), "peak_dmma_tflops": dmma,
                  "share_of_step": c2["loglike_ms"] / c2["device_ms"]}
    h.close()
    line = {"config": name, "workload": c["what"], "particle_updates_per_s": updates / (ms * 1e-3), "ms_per_iteration": ms / n_iter,
            "iterations": n_iter, "particles": P, "sweeps_per_iteration": h.B, "levels_per_sweep": (c1["levels"] - c0["levels"]) / (n_iter * h.B),
            "kernel_launches": c1["kernel_launches"] - c0["kernel_launches"], "accept_rate": float(acc), "wall_s": wall,
            "timing": "CUDA events on the library's stream around the whole call (demcmc_counters.device_ms), data resident, no L2 flush"}
    if c["flops"]:
        line["likelihood_tflops_whole_step"] = c["flops"] * updates / (ms * 1e-3) / 1e12
        line["likelihood_kernel"] = ll
    ups = line["particle_updates_per_s"]
    # roofline of each shape: what bounds it, achieved / peak
    if name == "c1":
        line["roofline"] = {"bound": "latency", "note": "24 particles x 50 observations: every level of a chunk in one single-CTA launch (k_chunk_small); during burn-in a chunk is one sweep (select_base reads the sweep-start weights) and the step is the host's per-chunk work; nothing to saturate"}
    elif name == "c3":
        dens = ups * 100_000
        line["roofline"] = {"bound": "fp64 pipe (transcendental)", "trial_densities_per_s": dens, "unit": "fp64 instructions/s",
                            "fp64_instructions_per_density": LBA_FP64_INST_PER_DENSITY}
        if peaks:
            peak_i = peaks[0] * 1e12 / 2.0
            line["roofline"].update(achieved=dens * LBA_FP64_INST_PER_DENSITY, peak=peak_i, frac=dens * LBA_FP64_INST_PER_DENSITY / peak_i,
                                    note="instruction roofline of the fp64 CUDA-core pipe (DFMA-loop peak / 2 instructions per second); ncu of the same kernel: sm__pipe_fp64_cycles_active 75.8 %")
    elif name == "c4" and hbm_gbs:
        gbs = ups * 64.0 * c["d"] / 1e9              # propose (4 reads + 1 write) + accept (1 read + 2 writes) of d doubles
        line["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                            "note": "algorithmic bytes = 64 d per particle update (SURVEY 8d) over the whole step; the likelihood (k_xdot) share is in likelihood_kernel"}
    elif name == "c5" and ll:
        line["roofline"] = {"bound": "tensor", "achieved": ll["tflops_event_bracketed"], "peak": ll["peak_dmma_tflops"], "unit": "TFLOP/s",
                            "frac": ll["of_measured_dmma_peak"], "whole_step_frac": line["likelihood_tflops_whole_step"] / ll["peak_dmma_tflops"]}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--iters", type=int, default=None)
    a = ap.parse_args()
    D._ffi.use_library(D._ffi.DEFAULT_LIB)
    for name in a.configs:
        print(json.dumps(run(name, a.iters)), flush=True)


if __name__ == "__main__":
    main()

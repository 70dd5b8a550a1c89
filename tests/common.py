"""Shared fixtures: the six registered models as (oracle Model, handle.set_model kwargs) pairs on
seeded synthetic data, plus matching sampler configurations."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import demcmc_b200 as D  # noqa: E402
from oracle import oracle as O  # noqa: E402

EMU_LIB = os.path.join(ROOT, "tests", "emu", "libdemcmc_emu.so")
CUDA_LIB = D._ffi.DEFAULT_LIB
INF = np.inf


def use_emu():
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"], check=True)
    D._ffi.use_library(EMU_LIB)


def use_cuda():
    D._ffi.use_library(CUDA_LIB)
    assert D._ffi.lib().demcmc_backend_name() == b"cuda-sm100a"


def lba_sim(rng, n, nu=(3.0, 2.0), A=0.8, k=0.2, tau=0.3):
    """Standard LBA generator: start ~ U(0,A), drift ~ N(nu,1) redrawn until one is positive."""
    nu = np.asarray(nu)
    choice = np.zeros(n, dtype=np.int32)
    rt = np.zeros(n)
    b = A + k
    for i in range(n):
        while True:
            v = rng.normal(nu, 1.0)
            if (v > 0).any():
                break
        a = rng.uniform(0, A, size=nu.size)
        t = np.where(v > 0, (b - a) / np.where(v > 0, v, 1.0), np.inf)
        choice[i] = int(np.argmin(t)) + 1
        rt[i] = tau + t.min()
    return choice, rt


def lnr_sim(rng, n, nu=(-2.0, -2.0, -3.0, -3.0), tau=0.5):
    nu = np.asarray(nu)
    x = np.exp(rng.normal(nu, 1.0, size=(n, nu.size)))
    return (np.argmin(x, axis=1) + 1).astype(np.int32), tau + x.min(axis=1)


class Case:
    """One model + sampler setup usable with both the oracle and a Handle."""

    def __init__(self, name, kind, d, prior, lo, hi, sample_prior, data):
        self.name, self.kind, self.d, self.prior = name, kind, d, prior
        self.lo, self.hi = np.asarray(lo, float), np.asarray(hi, float)
        self.sample_prior = sample_prior
        self.data = data      # kwargs: x, choice, sigma...

    def oracle_model(self):
        return O.Model(self.kind, self.d, self.prior, **self.data)

    def oracle_config(self, G, Np, **kw):
        return O.Config(G, Np, self.d, self.lo, self.hi, **kw)

    def handle(self, G, Np, **kw):
        h = D.Handle(G, Np, self.d, self.lo, self.hi, **kw)
        h.set_model(self.kind, self.prior, **self.data)
        return h

    def theta0(self, rng, P):
        return np.array([self.sample_prior(rng) for _ in range(P)])


def halfcauchy(rng):
    return abs(rng.standard_cauchy())


def make_case(name, rng, n_obs=None, n_subjects=9, n_dim=None):
    if name == "gaussian":
        n = n_obs or 50
        x = rng.normal(0.0, 1.0, n)
        return Case(name, "gaussian", 2, [("normal", 0, 1), ("halfcauchy", 0, 1)], [-INF, 0], [INF, INF],
                    lambda r: [r.normal(), halfcauchy(r)], dict(x=x))
    if name == "mvnormal":
        n = n_obs or 200
        dm = n_dim or 7
        mu = rng.normal(size=dm)
        x = rng.normal(mu, 1.0, size=(n, dm))
        return Case(name, "mvnormal", dm + 1, [("normal", 0, 1)] * dm + [("halfcauchy", 0, 1)], [-INF] * dm + [0],
                    [INF] * (dm + 1), lambda r: list(r.normal(size=dm)) + [halfcauchy(r) + 0.3], dict(x=x))
    if name == "mvnormal_full":  # MvNormal(mu, sigma^2 Sigma) with a known, strongly correlated covariance
        n = n_obs or 150
        dm = 6
        A = rng.normal(size=(dm, dm))
        cov = A @ A.T + 0.5 * np.eye(dm)
        mu = rng.normal(size=dm)
        x = rng.multivariate_normal(mu, 1.3 ** 2 * cov, size=n)
        return Case(name, "mvnormal_full", dm + 1, [("normal", 0, 2)] * dm + [("halfcauchy", 0, 1)], [-INF] * dm + [0],
                    [INF] * (dm + 1), lambda r: list(r.normal(mu, 0.3)) + [halfcauchy(r) + 0.5], dict(x=x, cov=cov))
    if name == "rastrigin":      # test/optimization_tests.jl:8-23: x in [-5, 5]^2, no data, no prior
        return Case(name, "rastrigin", 2, [("flat",), ("flat",)], [-5.0, -5.0], [5.0, 5.0], lambda r: list(r.uniform(-5, 5, 2)), dict())
    if name == "binomial":
        return Case(name, "binomial", 1, [("beta", 1, 1)], [0], [1], lambda r: [r.uniform()], dict(x=np.array([10.0, 4.0])))
    if name == "lnr":
        n = n_obs or 100
        choice, rt = lnr_sim(rng, n)
        mn = rt.min()
        return Case(name, "lnr", 5, [("normal", 0, 3)] * 4 + [("uniform", 0, mn)], [-INF] * 4 + [0], [INF] * 4 + [mn],
                    lambda r: list(r.normal(0, 3, 4)) + [r.uniform(0, mn)], dict(x=rt, choice=choice, n_dim=4))
    if name == "lba":
        n = n_obs or 100
        choice, rt = lba_sim(rng, n)
        mn = rt.min()
        prior = [("normal", 1, 5), ("normal", 1, 5), ("normal", 0.8, 0.2), ("normal", 0.2, 0.1), ("uniform", 0, mn)]
        return Case(name, "lba", 5, prior, [0, 0, 0, 0, 0], [INF, INF, INF, INF, mn],
                    lambda r: [abs(r.normal(1, 5)), abs(r.normal(1, 5)), abs(r.normal(0.8, 0.2)), abs(r.normal(0.2, 0.1)), r.uniform(0, mn)],
                    dict(x=rt, choice=choice, n_dim=2))
    if name == "hier_normal":
        S, n = n_subjects, (n_obs or 12)
        b0 = rng.normal(0, 1, S)
        y = rng.normal(1.0 + b0[:, None], 0.5, size=(S, n))
        prior = [("normal", 1, 1), ("halfcauchy", 0, 1)] + [("normal_ref", 0, 0, 1)] * S + [("halfcauchy", 0, 1)]
        lo = [-INF, 0] + [-INF] * S + [0]
        hi = [INF] * (S + 3)

        def sp(r):
            sb = halfcauchy(r) + 0.2
            return [r.normal(1, 1), sb] + list(r.normal(0, sb, S)) + [halfcauchy(r) + 0.2]
        return Case(name, "hier_normal", S + 3, prior, lo, hi, sp, dict(x=y))
    raise KeyError(name)


ALL_MODELS = ["gaussian", "mvnormal", "binomial", "lnr", "lba", "hier_normal"]


def hier_blocks(S):
    return np.array([[1, 1] + [0] * S + [1], [0, 0] + [1] * S + [0]], dtype=np.uint8)


def compare_run(case, G, Np, n_iter, mode, seed=5, rng_seed=0, rtol=1e-12, **kw):
    """Runs the oracle and the bound library on the same inputs and returns the comparison.
    mode = "replay": the library consumes the oracle's tape (reference semantics);
    mode = "native": both draw from Philox with the same seed (select_base on sweep-start weights)."""
    rng = np.random.default_rng(rng_seed)
    theta0 = case.theta0(rng, G * Np)
    okw = dict(kw)
    n0 = kw.get("n_initial", 0)
    # initialize_samples (utilities.jl:35-39): n_initial prior draws per particle id; the chain starts from row 1
    init_rows = np.stack([case.theta0(rng, G * Np) for _ in range(n0)]) if n0 else None
    cfg = case.oracle_config(G, Np, seed=seed, base_snapshot=1 if mode == "native" else 0, **okw)
    r = O.run(cfg, case.oracle_model(), theta0, n_iter, init_rows=init_rows)
    h = case.handle(G, Np, seed=seed, trace=True, **kw)
    try:
        if n0:
            h.set_history(init_rows)
            h.set_state(None)
        else:
            h.set_state(theta0)
        if mode == "native":
            h.run(n_iter)
        else:
            h.replay(r["tape"], n_iter)
        out = dict(samples=h.samples(), accept=h.accept(), lp=h.lp(), trace=h.trace(), state=h.get_state(),
                   mig=h.migration_slots(), counters=h.counters())
    finally:
        h.close()
    return r, out


def forced_run(case, G, Np, n_iter, mode, seed=5, rng_seed=0, **kw):
    """Teacher-forced comparison (SURVEY.md 7.3 hard part 2): the oracle runs the whole chain; the
    library runs ONE iteration at a time, each started from the oracle's state after the previous
    iteration, so rounding differences (different libm, different reduction order) cannot be
    amplified by the population dynamics (theta' = theta_t + gamma (theta_m - theta_n) has a
    positive Lyapunov exponent).  Returns the same structure as compare_run."""
    rng = np.random.default_rng(rng_seed)
    theta0 = case.theta0(rng, G * Np)
    n0 = kw.get("n_initial", 0)
    init_rows = np.stack([case.theta0(rng, G * Np) for _ in range(n0)]) if n0 else None
    cfg = case.oracle_config(G, Np, seed=seed, base_snapshot=1 if mode == "native" else 0, **kw)
    r = O.run(cfg, case.oracle_model(), theta0, n_iter, init_rows=init_rows)
    h = case.handle(G, Np, seed=seed, trace=True, **kw)
    traces, migs = [], []
    try:
        for it in range(n_iter):
            if it == 0 and n0:
                h.set_history(init_rows)
                h.set_state(None)
            elif it == 0:
                h.set_state(theta0)
            else:
                h.set_state(r["trace"]["state_theta"][it - 1], r["trace"]["state_id"][it - 1])
            if mode == "native":
                h.run(1)
            else:
                B = h.B
                one = {}
                for k, v in r["tape"].items():
                    if v is None:
                        continue
                    one[k] = v[it:it + 1] if k.startswith("mig_") else v[it * B:(it + 1) * B]
                h.replay(one, 1)
            traces.append(h.trace())
            migs.append(h.migration_slots())
        out = dict(samples=h.samples(), accept=h.accept(), lp=h.lp(), state=h.get_state(), counters=h.counters(),
                   trace={k: np.concatenate([t[k] for t in traces]) for k in traces[0]}, mig=np.concatenate(migs))
    finally:
        h.close()
    return r, out


def rel_err(a, b):
    """max |a-b| / max(1,|b|) over finite entries; non-finite entries must agree exactly."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    same_nonfinite = np.array_equal(np.isnan(a[~fin]), np.isnan(b[~fin])) and np.array_equal(a[~fin][~np.isnan(a[~fin])], b[~fin][~np.isnan(b[~fin])])
    if not same_nonfinite:
        return np.inf
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(1.0, np.abs(b[fin]))))


def mvn_resample_check(n_iter, burnin, sd_atol, seed=505514):
    """The assertions of test/multivariate_normal_tests.jl:62-69 on the bound library."""
    rng = np.random.default_rng(seed)
    n_mu, n_d = 30, 100
    data = rng.normal(0.0, 1.0, size=(n_d, n_mu))
    model = D.DEModel(sample_prior=lambda: [rng.normal(0, 1, n_mu), abs(rng.standard_cauchy())],
                      prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                      loglike=D.GPULoglike("mvnormal", data), names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), sample=D.resample, burnin=burnin,
              n_initial=(n_mu + 1) * 4, Np=3, n_groups=1, θsnooker=0.1, seed=7)
    chains = D.sample(model, de, D.MCMCThreads(), n_iter)
    assert len(chains) == n_iter - burnin
    means, sds = chains.mean()[:n_mu], chains.std()[:n_mu]
    assert np.all(np.abs(sds - 0.1) < sd_atol), sds
    assert np.all(np.abs(means) < 0.3)
    assert abs(means.std(ddof=1) - 0.1) < 0.02
    assert np.corrcoef(data.mean(axis=0), means)[0, 1] > 0.98


def lnr_posterior_check():
    """test/lognormal_race_tests.jl restated at its own size on the bound library: LNR(nu = [-2,-2,-3,-3], sigma = 1, tau = 0.5),
    100 trials, DE(burnin = 2000, Np = 24, n_groups = 4), 5000 iterations through sample(..., MCMCThreads(), ...); the
    assertions are the reference's (rhat within 0.05 of 1, posterior means and sds within 5 % of an independent sampler's).  The
    independent sampler is NUTS there; here it is the random-walk Metropolis run of tests/golden/make_lnr_posterior.py
    (scipy densities, no code of this repo), whose moments and Monte-Carlo errors are in tests/golden/lnr_posterior.json."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lnr_posterior.json")) as f:
        gold = json.load(f)
    assert max(gold["split_rhat"]) < 1.01 and np.all(np.array(gold["mcse_mean"]) < 0.01 * np.array(gold["sd"]) * 5)
    choice, rt, min_rt = np.array(gold["choice"], np.int32), np.array(gold["rt"]), gold["min_rt"]
    rng = np.random.default_rng(9918)
    model = D.DEModel(sample_prior=lambda: [rng.normal(0, 3, 4), rng.uniform(0, min_rt)],
                      prior_loglike=D.GPUPrior(D.Normal(0, 3), D.Uniform(0.0, min_rt)),
                      loglike=D.GPULoglike("lnr", choice=choice, rt=rt), names=("ν", "τ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, min_rt)), burnin=2000, Np=24, n_groups=4, seed=3)
    chains = D.sample(model, de, D.MCMCThreads(), 5000)
    assert len(chains) == 3000 and chains.value.shape[2] == 96
    mean, sd, rhat = chains.mean()[:5], chains.std()[:5], chains.rhat()[:5]
    assert np.all(np.abs(rhat - 1.0) < 0.05), rhat                                     # lognormal_race_tests.jl:64
    assert np.allclose(mean, gold["mean"], rtol=0.05), (mean, gold["mean"])            # :65
    assert np.allclose(sd, gold["sd"], rtol=0.05), (sd, gold["sd"])                    # :66


def blocking_posterior_check(seed=58122):
    """test/blocking_tests.jl restated at its own size on the bound library: Normal(mu, sigma), 1000 observations,
    blocks [[true, false], [false, true]] with blocking_on = x -> true, DE(burnin = 1000, Np = 6), 2000 iterations, through
    both sample(model, de, n_iter) and sample(model, de, MCMCThreads(), n_iter); the reference's assertions (:59-62, :69-72)."""
    rng = np.random.default_rng(seed)
    data = rng.normal(0.0, 1.0, 1000)
    model = D.DEModel(sample_prior=lambda: [rng.normal(0, 10), abs(rng.standard_cauchy())],
                      prior_loglike=D.GPUPrior(D.Normal(0, 10), D.HalfCauchy(0, 1)),
                      loglike=D.GPULoglike("gaussian", data), names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), burnin=1000, Np=6, seed=seed,
              blocking_on=lambda de_: True, blocks=[[True, False], [False, True]])
    for args in ((2000,), (D.MCMCThreads(), 2000)):
        chains = D.sample(model, de, *args)
        assert len(chains) == 1000
        mean, rhat = chains.mean()[:2], chains.rhat()[:2]
        assert abs(mean[0] - 0.0) < 0.1 and abs(mean[1] - 1.0) < 0.1, mean
        assert np.all(np.abs(rhat - 1.0) < 0.01), rhat


def optimize_checks():
    """test/optimization_tests.jl restated on the bound library: Rastrigin minimum and Gaussian MLE."""
    # Rastrigin has a lattice of local minima and a greedy 6-particle population settles in one basin:
    # which one depends on the draws (the reference pins ITS outcome with Random.seed!(78454111)).
    # Here: every run must end in a local minimum, and the global one (0 within the reference's 1e-8)
    # must be found by several of ten seeds.
    vals = []
    for seed in range(1, 11):
        rng = np.random.default_rng(78454111 + seed)
        model = D.DEModel(sample_prior=lambda: [rng.uniform(-5, 5, 2)], loglike=D.GPULoglike("rastrigin"), names=("x",))
        de = D.DE(sample_prior=model.sample_prior, bounds=((-5.0, 5.0),), Np=6, n_groups=1, update_particle=D.minimize,
                  evaluate_fitness=D.evaluate_fun, seed=seed)
        particles = D.optimize(model, de, 10_000)
        parms, val = D.get_optimal(de, model, particles)
        x = parms["x"]
        grad = 2 * x + 20 * np.pi * np.sin(2 * np.pi * x)
        assert np.all(np.abs(grad) < 1e-2) and np.all(np.abs(x) <= 5.0), (seed, x, val)
        vals.append(val)
    assert sum(abs(v) < 1e-8 for v in vals) >= 2, vals
    rng = np.random.default_rng(50514)
    data = rng.normal(0, 1, 100)
    model = D.DEModel(sample_prior=lambda: [rng.normal(0, 1), abs(rng.standard_cauchy())], loglike=D.GPULoglike("gaussian", data), names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.1, np.inf)), burnin=1000, Np=6, n_groups=1,
              update_particle=D.maximize, evaluate_fitness=D.evaluate_fun, seed=4)
    particles = D.optimize(model, de, D.MCMCThreads(), 10_000)
    parms, LL = D.get_optimal(de, model, particles)
    assert abs(parms["μ"] - data.mean()) < 1e-4 and abs(parms["σ"] - data.std()) < 1e-4
    assert len(particles) == 6 and sorted(p.id for p in particles) == list(range(1, 7))

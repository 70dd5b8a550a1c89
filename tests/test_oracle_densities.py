"""Pins the oracle's closed-form densities against scipy / mpmath, since the
third-party Julia packages that define them are not in the reference tree."""
import numpy as np
import pytest
from scipy import integrate, stats

from common import O, make_case


def test_gaussian():
    rng = np.random.default_rng(0)
    c = make_case("gaussian", rng)
    m = c.oracle_model()
    for th in ([0.3, 1.2], [-2.0, 0.05], [10.0, 30.0]):
        ref = stats.norm(th[0], th[1]).logpdf(c.data["x"]).sum()
        assert np.isclose(O.loglike(m, th), ref, rtol=1e-13)
        assert np.isclose(O.prior_loglike(m, th), stats.norm(0, 1).logpdf(th[0]) + stats.halfcauchy.logpdf(th[1]), rtol=1e-13)


def test_mvnormal():
    rng = np.random.default_rng(1)
    c = make_case("mvnormal", rng)
    m = c.oracle_model()
    x = c.data["x"]
    for _ in range(3):
        th = np.append(rng.normal(size=7), rng.uniform(0.3, 2.0))
        ref = stats.multivariate_normal(th[:7], th[7] ** 2 * np.eye(7)).logpdf(x).sum()
        assert np.isclose(O.loglike(m, th), ref, rtol=1e-13)


def test_binomial_and_beta_prior():
    c = make_case("binomial", np.random.default_rng(2))
    m = c.oracle_model()
    for p in (0.05, 0.4, 0.93):
        assert np.isclose(O.loglike(m, [p]), stats.binom(10, p).logpmf(4), rtol=1e-13)
        assert O.prior_loglike(m, [p]) == 0.0
    assert O.prior_loglike(m, [1.5]) == -np.inf
    m2 = O.Model("binomial", 1, [("beta", 2.5, 4.0)], x=np.array([10.0, 4.0]))
    assert np.isclose(O.prior_loglike(m2, [0.3]), stats.beta(2.5, 4.0).logpdf(0.3), rtol=1e-13)


def test_lnr():
    rng = np.random.default_rng(3)
    c = make_case("lnr", rng)
    m = c.oracle_model()
    rt, ch = c.data["x"], c.data["choice"]
    th = np.array([-1.5, -2.2, -2.9, -3.1, 0.5 * rt.min()])
    ref = 0.0
    for t, w in zip(rt, ch):
        for r in range(4):
            dist = stats.lognorm(s=1.0, scale=np.exp(th[r]))
            ref += dist.logpdf(t - th[4]) if r == w - 1 else dist.logsf(t - th[4])
    assert np.isclose(O.loglike(m, th), ref, rtol=1e-12)
    # deep tail of the survivor: log(1-Phi(z)) stays accurate where 1-cdf underflows
    m1 = O.Model("lnr", 3, [("flat",)] * 3, x=np.array([np.exp(45.0)]), choice=np.array([1], dtype=np.int32), n_dim=2)
    import mpmath as mp
    z = mp.mpf(45.0) - 2
    ref = float(mp.log(mp.erfc(z / mp.sqrt(2)) / 2)) + float(stats.norm.logpdf(45.0 - 1.0) - 45.0)
    assert np.isclose(O.loglike(m1, [1.0, 2.0, 0.0]), ref, rtol=1e-12)


def _lba_pdf(c, t, nu, A, k, tau):
    b, dt = A + k, t - tau
    den = 1.0
    for r, v in enumerate(nu):
        n1, n2 = (b - A - dt * v) / dt, (b - dt * v) / dt
        if r == c:
            den *= max(0.0, (-v * stats.norm.cdf(n1) + stats.norm.pdf(n1) + v * stats.norm.cdf(n2) - stats.norm.pdf(n2)) / A)
        else:
            F = 1 + ((b - A - dt * v) / A) * stats.norm.cdf(n1) - ((b - dt * v) / A) * stats.norm.cdf(n2) \
                + (dt / A) * stats.norm.pdf(n1) - (dt / A) * stats.norm.pdf(n2)
            den *= 1 - max(0.0, F)
    return den / (1 - np.prod(stats.norm.cdf(-np.asarray(nu))))


def test_lba():
    rng = np.random.default_rng(4)
    c = make_case("lba", rng)
    m = c.oracle_model()
    rt, ch = c.data["x"], c.data["choice"]
    th = [2.7, 1.8, 0.75, 0.25, 0.5 * rt.min()]
    ref = sum(np.log(max(_lba_pdf(w - 1, t, th[:2], th[2], th[3], th[4]), 1e-10)) for t, w in zip(rt, ch))
    assert np.isclose(O.loglike(m, th), ref, rtol=1e-12)
    # the defective densities of the two accumulators integrate to one
    edges = [0.3, 1.0, 3.0, 10.0, 100.0, 1e4, 1e6]
    mass = sum(integrate.quad(lambda t: _lba_pdf(r, t, th[:2], th[2], th[3], 0.3), a, b, limit=200)[0]
               for r in range(2) for a, b in zip(edges[:-1], edges[1:]))
    assert abs(mass - 1.0) < 1e-5
    # rt < tau hits the floor (SequentialSamplingModels) or -inf when the floor is disabled
    m0 = O.Model("lba", 5, c.prior, x=np.array([0.2]), choice=np.array([1], dtype=np.int32), n_dim=2, lba_floor=0.0)
    assert O.loglike(m0, [2.7, 1.8, 0.75, 0.25, 0.3]) == -np.inf
    mf = O.Model("lba", 5, c.prior, x=np.array([0.2]), choice=np.array([1], dtype=np.int32), n_dim=2)
    assert np.isclose(O.loglike(mf, [2.7, 1.8, 0.75, 0.25, 0.3]), np.log(1e-10))


def test_hier_normal():
    rng = np.random.default_rng(5)
    c = make_case("hier_normal", rng)
    m = c.oracle_model()
    y = c.data["x"]
    S = y.shape[0]
    th = np.concatenate([[0.8, 1.3], rng.normal(0, 1, S), [0.6]])
    ref = sum(stats.norm(th[0] + th[2 + s], th[-1]).logpdf(y[s]).sum() for s in range(S))
    assert np.isclose(O.loglike(m, th), ref, rtol=1e-13)
    pref = stats.norm(1, 1).logpdf(th[0]) + stats.halfcauchy.logpdf(th[1]) + stats.norm(0, th[1]).logpdf(th[2:2 + S]).sum() \
        + stats.halfcauchy.logpdf(th[-1])
    assert np.isclose(O.prior_loglike(m, th), pref, rtol=1e-13)


def test_posterior_bounds():
    c = make_case("gaussian", np.random.default_rng(6))
    cfg = c.oracle_config(4, 6)
    m = c.oracle_model()
    assert O.posterior(cfg, m, [0.1, -0.5]) == -np.inf      # sigma below its bound
    assert O.posterior(cfg, m, [np.nan, 1.0]) == -np.inf    # NaN fails in_bounds (utilities.jl:70)
    assert np.isfinite(O.posterior(cfg, m, [0.1, 0.0 + 1e-9]))
    assert np.isclose(O.posterior(cfg, m, [0.1, 1.1]), O.prior_loglike(m, [0.1, 1.1]) + O.loglike(m, [0.1, 1.1]))


@pytest.mark.parametrize("kind,args,dist", [
    ("normal", (1.5, 0.7), stats.norm(1.5, 0.7)),
    ("uniform", (0.0, 0.37), stats.uniform(0.0, 0.37)),
    ("halfcauchy", (0.0, 2.0), stats.halfcauchy(0, 2.0)),
])
def test_priors(kind, args, dist):
    m = O.Model("binomial", 1, [(kind,) + args], x=np.array([10.0, 4.0]))
    for x in (0.01, 0.2, 0.36):
        assert np.isclose(O.prior_loglike(m, [x]), dist.logpdf(x), rtol=1e-13)

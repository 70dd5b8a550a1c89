"""GPU parity tests of the ARITHMETIC of the streamed likelihood kernels k_xdot and k_chunk_persist
(Examples/Multivariate_Guassian_Example.jl:31-33, Examples/Hierarchical_Example.jl:36-44): the cross term
on operands whose answer is not zero, against an extended-precision reference, for every packing shape
(half k-step, several dimension splits, ragged observation tiles, 1-4 octets with a padded last octet) and
both launch paths -- plus mutation tests that must FAIL.  See tests/xdot_common.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

import common  # noqa: F401  (path setup)
import demcmc_b200 as D
import xdot_common as X
from oracle import oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cuda")]


# d = 50: half step, nj = 13 | d = 100: 2 splits of 50 | d = 52: 13 full steps | d = 7: short split, nj = 2 |
# d = 3: one k-step | d = 110: 3 splits of 37 (nj = 10, 3 padded dims) | hierarchical 1000 subjects x 50: 20 splits,
# ONE ragged observation tile | P: 1..4 octets, padded last octet
@pytest.mark.parametrize("kind,n,k", [("mvnormal", 130, 50), ("mvnormal", 65, 100), ("mvnormal", 63, 52), ("mvnormal", 1, 50),
                                      ("mvnormal", 64, 7), ("mvnormal", 200, 3), ("mvnormal", 700, 110), ("mvnormal", 5000, 50),
                                      ("hier_normal", 50, 1000), ("hier_normal", 12, 9), ("hier_normal", 130, 50), ("hier_normal", 65, 102)])
@pytest.mark.parametrize("P", [1, 8, 9, 17, 25, 32, 33, 77])
def test_cross_term_of_arbitrary_vectors(kind, n, k, P):
    err, ll, case, th = X.eval_error(kind, n, k, P, seed=n + k + P)
    assert err <= X.TOL, err
    # and the log-likelihood built on it against the oracle's direct form sum (x - m)^2
    m = O.Model(kind, case["d"], case["prior"], x=case["x"])
    ref = np.array([O.loglike(m, t) for t in th])
    assert common.rel_err(ll, ref) <= 1e-11


@pytest.mark.parametrize("persist", ["1", "0"])
@pytest.mark.parametrize("kind,n,k,G,Np", [("mvnormal", 130, 50, 2, 24), ("mvnormal", 65, 100, 4, 9), ("mvnormal", 63, 52, 2, 40),
                                           ("mvnormal", 300, 7, 3, 13), ("hier_normal", 50, 1000, 2, 12), ("hier_normal", 130, 50, 4, 33),
                                           ("mvnormal", 130, 50, 1, 64)])
def test_cross_term_of_every_proposal_of_a_run(kind, n, k, G, Np, persist, monkeypatch):
    """both launch paths: the persistent chunk kernel (>= 2 groups) and the level-by-level k_xdot launches"""
    monkeypatch.setenv("DEMCMC_PERSIST", persist)
    kw = dict(theta_snooker=0.2 if kind == "mvnormal" else 0.0, alpha=0.3)
    if kind == "hier_normal":
        kw["blocks"] = common.hier_blocks(k)
    err, ctr, n_checked = X.run_error(kind, n, k, G, Np, 6, seed=k + Np, **kw)
    assert n_checked > 0.9 * ctr["particle_updates"]
    assert err <= X.TOL, err
    if persist == "1" and G >= 2:
        assert ctr["persistent_chunks"] > 0
    else:
        assert ctr["persistent_chunks"] == 0


def test_cross_term_full_size_both_paths(monkeypatch):
    """configs[1] at its own size (d = 50, 1e5 observations, 4 x 256): every proposal of 4 iterations"""
    for persist in ("1", "0"):
        monkeypatch.setenv("DEMCMC_PERSIST", persist)
        err, ctr, n_checked = X.run_error("mvnormal", 100_000, 50, 4, 256, 4, seed=50514, theta_snooker=0.1)
        assert n_checked == 4096 and err <= X.TOL, (persist, err)
        assert (ctr["persistent_chunks"] > 0) == (persist == "1")


def test_default_centre_makes_the_cross_term_vanish():
    """what the product does by default, said in a test: data centred on their column means => B == 0 up to the
    rounding of the column means (|B| <= 2^-40 of the bound), so the likelihood rests on O(d) sufficient statistics
    and the stream is kept only because the metric counts "loglike evals incl." """
    rng = np.random.default_rng(3)
    case = X.make("mvnormal", 5000, 50, rng)
    th = case["draw"](64)
    with X.handle(case, 1, 64, center="mean") as h:
        B = h.eval_xdot(th)
        ll = h.eval(th)[0]
        h.set_sufficient_stat(True)
        assert np.array_equal(h.eval(th)[0], ll) or common.rel_err(h.eval(th)[0], ll) <= 1e-15
    _, bound = X.reference(case, th, center=case["x"].mean(axis=0))
    assert np.max(np.abs(B) / np.asarray(bound, dtype=float)) <= X.TOL


def test_sufficient_stat_mode_is_the_same_chain():
    """demcmc_set_sufficient_stat: skipping the stream changes no accept decision and no draw beyond 1e-12"""
    rng = np.random.default_rng(5)
    case = X.make("mvnormal", 2000, 50, rng)
    th0 = case["draw"](4 * 32)
    outs = []
    for on in (False, True):
        with X.handle(case, 4, 32, center="mean", seed=9, burnin=3, theta_snooker=0.1) as h:
            h.set_sufficient_stat(on)
            h.set_state(th0)
            h.run(1)                        # one iteration: later ones amplify last-bit differences of the weights
            outs.append((h.accept(), h.samples(), h.lp(), h.counters()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert common.rel_err(outs[1][1], outs[0][1]) <= 1e-12 and common.rel_err(outs[1][2], outs[0][2]) <= 1e-12
    with X.handle(case, 4, 32) as h:       # an explicit centre has a non-zero cross term: refused
        with pytest.raises(D._ffi.DemcmcError):
            h.set_sufficient_stat(True)


@pytest.mark.parametrize("corrupt", ["1", "2", "3"])
def test_mutations_are_caught(corrupt):
    """DEMCMC_TEST_CORRUPT breaks the B-fragment index (1), the half-step pack (2) or drops the last k-step (3):
    the cross-term parity check must fail -- and must pass without it."""
    script = os.path.join(os.path.dirname(__file__), "xdot_common.py")
    env = dict(os.environ)
    env.pop("DEMCMC_TEST_CORRUPT", None)
    ok = subprocess.run([sys.executable, script, common.CUDA_LIB], env=env, capture_output=True, text=True)
    assert ok.returncode == 0, ok.stdout + ok.stderr
    env["DEMCMC_TEST_CORRUPT"] = corrupt
    bad = subprocess.run([sys.executable, script, common.CUDA_LIB], env=env, capture_output=True, text=True)
    assert bad.returncode == 3, bad.stdout + bad.stderr

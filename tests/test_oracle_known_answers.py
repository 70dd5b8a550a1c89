"""Pins the oracle on every portable known-answer vector the reference's tests hold for this path
(SURVEY.md section 4 / 8c): projection, reset!, particle algebra, the cyclic-shift property of
migration -- and on Philox4x32-10's published test vectors."""
import json
import os

import numpy as np

from common import O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


def test_projection():  # test/utility_tests.jl:71-93
    c = G["projection"]
    exp = np.array(c["expected_num"]) / c["expected_den"]
    assert np.allclose(O.project(c["p1"], c["p2"]), exp, rtol=1e-15)


def test_reset_vector_and_matrix():  # test/utility_tests.jl:42-69
    for key in ("reset_vector", "reset_matrix"):
        c = G[key]
        assert np.array_equal(O.reset(c["p1"], c["p2"], c["mask"]), np.array(c["expected"]))


def test_particle_algebra():  # test/utility_tests.jl:161-199
    for c in G["algebra"]["cases"]:
        out = O.de_proposal(c["pt"], c["pm"], c["pn"], None, c["g"], 0.0, c["b"])
        assert np.allclose(out, c["expected"], rtol=1e-15), c["expr"]
    # p + Uniform(-0.1, 0.1): within 0.2 and different (utility_tests.jl:194-198)
    rng = np.random.default_rng(29542)
    b = rng.uniform(-0.1, 0.1, 2)
    out = O.de_proposal([1.0, 2.0], [0, 0], [0, 0], None, 0.0, 0.0, b)
    assert np.allclose(out, [1.0, 2.0], atol=0.2) and not np.array_equal(out, [1.0, 2.0])


def test_random_gamma_association():
    """((Pt + g1*(Pm-Pn)) + g2*(Pb-Pt)) + b, folded left (crossover.jl:168)."""
    rng = np.random.default_rng(3)
    pt, pm, pn, pb, b = rng.normal(size=(5, 11))
    g1, g2 = 0.77, 0.61
    exp = ((pt + (pm - pn) * g1) + (pb - pt) * g2) + b
    assert np.array_equal(O.de_proposal(pt, pm, pn, pb, g1, g2, b), exp)


def test_snooker_matches_formula():  # crossover.jl:239-273
    rng = np.random.default_rng(4)
    pt, pz, pm, pn, b = rng.normal(size=(5, 6))
    g = 1.7
    pd = pt - pz
    proj = lambda p: pd * (np.dot(p, pd) / np.dot(pd, pd))
    exp = (pt + (proj(pm) - proj(pn)) * g) + b
    out = O.snooker_proposal(pt, pz, pm, pn, g, b)
    assert np.allclose(out, exp, rtol=1e-14)
    adj = O.adjust_loglike(pt, out, pz)
    assert np.isclose(adj, 5 * (np.log(np.linalg.norm(out - pz)) - np.log(np.linalg.norm(pt - pz))), rtol=1e-12)


def test_migration_cyclic_shift_property():  # test/utility_tests.jl:95-154
    Np, Gn = 4, 4
    tags = np.arange(Gn * Np, dtype=np.int32)
    groups = np.array([2, 0, 3], dtype=np.int32)
    slots = np.array([1, 3, 0], dtype=np.int32)
    out = O.shift(tags, groups, slots, Np)
    n = len(groups)
    for i in range(n):
        prev = (i - 1) % n
        assert out[groups[i] * Np + slots[i]] == tags[groups[prev] * Np + slots[prev]]
    untouched = np.setdiff1d(np.arange(Gn * Np), groups * Np + slots)
    assert np.array_equal(out[untouched], tags[untouched])


def test_accept_rule():  # utilities.jl:55-58
    assert O.accept(-1.0, -2.0, 0.0, 0.999999)          # p = 1
    assert O.accept(-3.0, -2.0, 0.0, np.exp(-1.0))       # u <= p inclusive
    assert not O.accept(-3.0, -2.0, 0.0, np.exp(-1.0) * (1 + 1e-15))
    assert not O.accept(-np.inf, -2.0, 0.0, 1e-300)      # out of bounds => p = 0
    assert O.accept(-np.inf, -2.0, 0.0, 0.0)             # ... unless u == 0.0
    assert not O.accept(-np.inf, -np.inf, 0.0, 0.0)      # NaN => reject
    assert not O.accept(1.0, 2.0, np.nan, 0.0)


def test_select_base_and_particle_quirks():  # SURVEY hard part 4
    # underflow: exp.(w) == 0 => NaN => raw log-weights as sampling weights: slot 1 w.p. (n-1)/n, else slot n
    w = np.array([-5000.0, -5100.0, -4900.0, -5050.0])
    picks = [O.select_base(w, u) for u in np.linspace(0.001, 0.999, 999)]
    assert set(picks) == {0, 3}
    assert abs(np.mean(np.array(picks) == 0) - 0.75) < 0.02
    # select_particle overflows to NaN => findmin(w), no draw
    i, drew = O.select_particle(w, 0.3)
    assert (i, drew) == (1, False)
    # well-scaled weights: softmax sampling / inverse softmax
    w = np.log(np.array([0.1, 0.2, 0.3, 0.4]))
    assert [O.select_base(w, u) for u in (0.05, 0.15, 0.45, 0.95)] == [0, 1, 2, 3]
    i, drew = O.select_particle(np.array([0.0, 0.0, 0.0]), 0.5)
    assert (i, drew) == (1, True)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [O.uniform2(9, 5, s, 3, 0) for s in range(2000)]
    u = np.array(u).ravel()
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.02

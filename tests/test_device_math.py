"""The device-side exp / erfcx of the LBA kernel (csrc/de_math.h: de_exp_nonpos, de_erfcx_nonneg) take their polynomial
coefficients from two constant tables.  This test parses the tables out of the header and evaluates the same arithmetic
(same reduction, same Horner order, fused multiply-adds emulated in extended precision) against mpmath."""
import os
import re

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "differentialevolutionmcmc.jl_b200", "csrc", "de_math.h")).read()


def table(name):
    body = re.search(name + r"\[\d+\] = \{(.*?)\};", SRC, flags=re.S).group(1)
    return np.array([eval(v.replace("/", "/")) for v in body.replace("\n", " ").split(",") if v.strip()], dtype=np.float64)


def fma(a, b, c):
    return (np.asarray(a, np.longdouble) * np.asarray(b, np.longdouble) + np.asarray(c, np.longdouble)).astype(np.float64)


def test_erfcx_table():
    c = table("DE_ERFCX_C")
    assert c.size == 27
    x = np.concatenate([np.linspace(0, 1, 300), np.linspace(1, 12, 500), np.logspace(1, 8, 200), [0.0, 1e-300, 26.5]])
    a, b = x + 3.75, fma(2.0, x, 1.0)
    r = 1.0 / (a * b)
    t = (x - 3.75) * b * r
    p = np.full_like(x, c[26])
    for j in range(25, -1, -1):
        p = fma(p, t, c[j])
    got = p * (a * r)
    mp.mp.dps = 40
    ref = np.array([float(mp.exp(mp.mpf(v) ** 2) * mp.erfc(mp.mpf(v))) for v in x])
    assert np.max(np.abs(got - ref) / ref) < 4 * 2.2e-16


def test_exp_table():
    c = table("DE_EXP_C")
    assert c.size == 14 and c[0] == 1.0 and c[13] == 1.0 / 6227020800.0
    x = -np.concatenate([np.linspace(0, 1, 300), np.linspace(1, 50, 400), np.linspace(50, 707.9, 300), [0.0, 1e-300, 0.34657, 0.34658]])
    kd = (fma(x, 1.4426950408889634074, 6755399441055744.0) + -6755399441055744.0)
    r = fma(kd, -6.93147180369123816490e-01, x)
    r = fma(kd, -1.90821492927058770002e-10, r)
    p = np.full_like(x, c[13])
    for j in range(12, -1, -1):
        p = fma(p, r, c[j])
    got = np.ldexp(p, kd.astype(int))
    mp.mp.dps = 40
    ref = np.array([float(mp.exp(mp.mpf(v))) for v in x])
    assert np.max(np.abs(got - ref) / ref) < 3 * 2.2e-16
    assert np.max(np.abs(r)) <= 0.34658

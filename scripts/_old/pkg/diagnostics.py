"""Convergence diagnostics used where the reference calls MCMCChains.describe (rhat, ess):
rank-normalised split-R-hat and bulk ESS (Vehtari, Gelman, Simpson, Carpenter, Bürkner 2021), the
estimators MCMCDiagnosticTools implements.  x has shape (n_draws, n_chains)."""
from __future__ import annotations

import numpy as np
from scipy import stats


def _split(x):
    n = x.shape[0] // 2
    return np.concatenate([x[:n], x[x.shape[0] - n:]], axis=1)


def _rank_normalise(x):
    r = stats.rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return stats.norm.ppf((r - 0.375) / (x.size + 0.25))


def _rhat(x):
    n, m = x.shape
    cm = x.mean(axis=0)
    cv = x.var(axis=0, ddof=1)
    B = n * cm.var(ddof=1)
    W = cv.mean()
    if W == 0:
        return np.nan
    return float(np.sqrt(((n - 1) / n * W + B / n) / W))


def split_rhat(x):
    x = np.asarray(x, dtype=np.float64)
    if x.shape[0] < 4:
        return np.nan
    s = _split(x)
    z = _rank_normalise(s)
    zf = _rank_normalise(np.abs(s - np.median(s)))
    return max(_rhat(z), _rhat(zf))


def _autocov(x):
    n = x.shape[0]
    m = 1 << (2 * n - 1).bit_length()
    xc = x - x.mean(axis=0)
    f = np.fft.rfft(xc, n=m, axis=0)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=0)[:n]
    return ac / n


def _ess(x):
    n, m = x.shape
    if n < 4:
        return np.nan
    acov = _autocov(x)
    cv = acov[0] * n / (n - 1.0)
    W = cv.mean()
    var_plus = W * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=0).var(ddof=1)
    if not var_plus > 0:
        return np.nan
    rho = 1.0 - (W - acov.mean(axis=1)) / var_plus
    rho[0] = 1.0
    # Geyer's initial positive + monotone sequence on pair sums
    tau = -1.0
    prev = np.inf
    t = 0
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        prev = pair
        tau += 2.0 * pair
        t += 2
    tau = max(tau, 1.0 / np.log10(max(n * m, 10)))
    return float(n * m / tau)


def bulk_ess(x):
    x = np.asarray(x, dtype=np.float64)
    return _ess(_rank_normalise(_split(x)))

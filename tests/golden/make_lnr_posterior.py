"""Golden posterior moments for the reference's lognormal-race test (test/lognormal_race_tests.jl), whose check is a
comparison with NUTS (Turing) -- not available here.  Stand-in: an independent sampler of the SAME posterior, written
with scipy densities only (no code of this repo, no oracle): adaptive random-walk Metropolis, 8 chains x 150 000
iterations after 30 000 of adaptation, started from dispersed points.  Writes lnr_posterior.json (data, posterior mean /
sd per parameter, the Monte-Carlo standard error of each from the between-chain spread, split R-hat of the stand-in).
    python tests/golden/make_lnr_posterior.py          (about a minute on one core)"""
import json, os
import numpy as np
from scipy import special

rng = np.random.default_rng(9918)                              # the reference script's seed number (not its stream)
nu_true, tau_true, n = np.array([-2.0, -2.0, -3.0, -3.0]), 0.5, 100
x = np.exp(rng.normal(nu_true, 1.0, size=(n, 4)))
choice = np.argmin(x, axis=1)                                   # 0-based winner
rt = tau_true + x.min(axis=1)
min_rt = float(rt.min())
onehot = np.zeros((n, 4), bool); onehot[np.arange(n), choice] = True


LOG2PI = float(np.log(2 * np.pi))


def logpost(th):
    """th[C][5] -> [C]: sum(logpdf(LNR(nu, sigma = 1, tau), data)) + Normal(0, 3) priors on nu + Uniform(0, min_rt) on tau;
    the winner's LogNormal log-density and the losers' log survival function (scipy.special.log_ndtr)"""
    nu, tau = th[:, :4], th[:, 4]
    ok = (tau >= 0.0) & (tau <= min_rt)
    t = rt[None, :] - np.where(ok, tau, 0.0)[:, None]                     # [C][n] (> 0 wherever ok: tau <= min_rt; == 0 only at tau == min_rt)
    ok &= (t > 0).all(axis=1)
    lt = np.log(np.where(t > 0, t, 1.0))
    z = lt[:, :, None] - nu[:, None, :]                                   # [C][n][4]
    ll = np.where(onehot[None], -0.5 * z * z - 0.5 * LOG2PI - lt[:, :, None], special.log_ndtr(-z)).sum(axis=(1, 2))
    prior = (-0.5 * (nu / 3.0) ** 2 - 0.5 * LOG2PI - np.log(3.0)).sum(axis=1) - np.log(min_rt)
    return np.where(ok, ll + prior, -np.inf)


def rwm(seed, n_chains=8, n_adapt=30_000, n_keep=150_000):
    """n_chains independent chains advanced together (vectorised over the chain axis; every chain has its own proposal scale)"""
    r = np.random.default_rng(seed)
    C = n_chains
    th = np.concatenate([nu_true + r.normal(0, 0.5, (C, 4)), min_rt * r.uniform(0.2, 0.9, (C, 1))], axis=1)
    lp = logpost(th)
    scale = 2.38 ** 2 / 5
    L = np.stack([np.linalg.cholesky(np.diag([0.02, 0.02, 0.05, 0.05, 1e-4]) * scale)] * C)
    hist = np.empty((n_adapt, C, 5))
    out = np.empty((C, n_keep, 5))
    for i in range(n_adapt + n_keep):
        prop = th + np.einsum("cij,cj->ci", L, r.normal(size=(C, 5)))
        lpp = logpost(prop)
        acc = np.log(r.uniform(size=C)) < lpp - lp
        th = np.where(acc[:, None], prop, th); lp = np.where(acc, lpp, lp)
        if i < n_adapt:
            hist[i] = th
            if i >= 2000 and i % 2000 == 0:                               # adapt each chain's proposal to its empirical covariance
                for c in range(C):
                    L[c] = np.linalg.cholesky(np.cov(hist[i // 2:i + 1, c].T) * scale + 1e-10 * np.eye(5))
        else:
            out[:, i - n_adapt] = th
    return out


if __name__ == "__main__":
    chains = rwm(100)                                             # [8][n_keep][5]
    m_c, s_c = chains.mean(axis=1), chains.std(axis=1, ddof=1)
    half = chains.shape[1] // 2
    sp = np.concatenate([chains[:, :half], chains[:, half:2 * half]], axis=0)
    W = sp.var(axis=1, ddof=1).mean(axis=0); B = half * sp.mean(axis=1).var(axis=0, ddof=1)
    rhat = np.sqrt(((half - 1) / half * W + B / half) / W)
    pooled = chains.reshape(-1, 5)
    out = dict(source="tests/golden/make_lnr_posterior.py: adaptive random-walk Metropolis with scipy densities, 8 x 150000 draws",
               reference_test="test/lognormal_race_tests.jl (LNR nu = [-2,-2,-3,-3], sigma = 1, tau = 0.5, 100 trials; compared with NUTS there)",
               choice=(choice + 1).tolist(), rt=rt.tolist(), min_rt=min_rt,
               mean=pooled.mean(axis=0).tolist(), sd=pooled.std(axis=0, ddof=1).tolist(),
               mcse_mean=(m_c.std(axis=0, ddof=1) / np.sqrt(8)).tolist(), mcse_sd=(s_c.std(axis=0, ddof=1) / np.sqrt(8)).tolist(),
               split_rhat=rhat.tolist())
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lnr_posterior.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: out[k] for k in ("mean", "sd", "mcse_mean", "mcse_sd", "split_rhat")})

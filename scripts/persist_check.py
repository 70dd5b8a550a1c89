"""Persistent chunk kernel against the level-by-level launches: same chains, bit for bit.
Run once per setting (the switch is read once per process):
  DEMCMC_PERSIST=0 python scripts/persist_check.py ; DEMCMC_PERSIST=1 [DEMCMC_LANES=2] python scripts/persist_check.py"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import demcmc_b200 as D

D._ffi.use_library(D._ffi.DEFAULT_LIB)
def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]

def mvn(n, dm, G, Np, n_iter, **kw):
    rng = np.random.default_rng(5)
    mu = rng.normal(size=dm)
    x = rng.normal(mu, 1.0, size=(n, dm))
    prior = [("normal", 0, 1)] * dm + [("halfcauchy", 0, 1)]
    lo = [-np.inf] * dm + [0.0]; hi = [np.inf] * (dm + 1)
    theta0 = np.column_stack([rng.normal(size=(G * Np, dm)), np.abs(rng.standard_cauchy(G * Np)) + 0.3])
    with D.Handle(G, Np, dm + 1, lo, hi, seed=3, **kw) as h:
        h.set_model("mvnormal", prior, x=x)
        h.set_state(theta0)
        t0 = time.perf_counter(); h.run(n_iter); t = time.perf_counter() - t0
        c = h.counters()
        return sha(h.samples(), h.lp(), h.accept()), c["kernel_launches"], c["persistent_chunks"], round(c["device_ms"], 3), round(t * 1e3, 1)

def hier(S, per, G, Np, n_iter):
    rng = np.random.default_rng(9)
    b = rng.normal(0, 1, size=S); y = rng.normal(1.0 + b[:, None], 0.5, size=(S, per))
    d = S + 3
    prior = [("normal", 1, 1), ("halfcauchy", 0, 1)] + [("normal_ref", 0, 1)] * S + [("halfcauchy", 0, 1)]
    lo = [-np.inf, 0.0] + [-np.inf] * S + [0.0]; hi = [np.inf] * d
    blocks = np.zeros((2, d), dtype=np.uint8); blocks[0, [0, 1, d - 1]] = 1; blocks[1, 2:d - 1] = 1
    theta0 = np.column_stack([rng.normal(1, 1, G * Np), np.abs(rng.standard_cauchy(G * Np)) + 0.2, rng.normal(size=(G * Np, S)), np.abs(rng.standard_cauchy(G * Np)) + 0.2])
    return None

print("persist =", os.environ.get("DEMCMC_PERSIST", "1"), "lanes =", os.environ.get("DEMCMC_LANES", "1"))
print("mvn small   ", mvn(20000, 10, 4, 32, 30, theta_snooker=0.1, alpha=0.2), flush=True)
print("mvn d50     ", mvn(20000, 50, 4, 64, 40, theta_snooker=0.1, burnin=10), flush=True)
print("mvn d100    ", mvn(10000, 100, 2, 128, 20, theta_snooker=0.1), flush=True)
print("mvn C2 shape", mvn(100000, 50, 4, 256, 100, theta_snooker=0.1), flush=True)

"""Soak: a long native run of configs[1]'s shape through the persistent chunk kernel and through the
level-by-level launches must give the same chain, bit for bit (a lost update, a stale operand or a
mis-ordered accept in the counter protocol would change the hash).  DEMCMC_PERSIST is read per launch."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import demcmc_b200 as D

D._ffi.use_library(D._ffi.DEFAULT_LIB)
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.default_rng(5)
n, dm, G, Np = 100_000, 50, 4, 256
x = rng.normal(rng.normal(size=dm), 1.0, size=(n, dm))
prior = [("normal", 0, 1)] * dm + [("halfcauchy", 0, 1)]
lo = [-np.inf] * dm + [0.0]; hi = [np.inf] * (dm + 1)
theta0 = np.column_stack([rng.normal(size=(G * Np, dm)), np.abs(rng.standard_cauchy(G * Np)) + 0.3])
out = {}
for persist in ("1", "0"):
    os.environ["DEMCMC_PERSIST"] = persist
    with D.Handle(G, Np, dm + 1, lo, hi, seed=3, burnin=0, theta_snooker=0.1) as h:
        h.set_model("mvnormal", prior, x=x)
        h.set_state(theta0)
        t0 = time.perf_counter(); h.run(n_iter); dt = time.perf_counter() - t0
        c = h.counters()
        th, w, ids, acc = h.history_by_slot(n_iter - 50, 50)
        state = h.get_state()
        hs = hashlib.sha256()
        for a in (th, w, ids, acc) + tuple(state):
            hs.update(np.ascontiguousarray(a).tobytes())
        out[persist] = hs.hexdigest()[:20]
        print(f"persist={persist}: {n_iter} iterations, {c['persistent_chunks']} persistent chunks, {c['kernel_launches']} launches, "
              f"{G * Np * n_iter / dt:.0f} updates/s, accept rate (last 50) {acc.mean():.4f}, hash {out[persist]}", flush=True)
assert out["0"] == out["1"], "the persistent kernel and the level-by-level path disagree"
print("identical")

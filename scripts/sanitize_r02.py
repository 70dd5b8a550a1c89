"""Round-2 paths for compute-sanitizer (memcheck): multi-device handle with cross-device migration through the mailboxes,
thinning, device diagnostics, the shared by-id outputs, the full-covariance kernel, the cross term with a given centre."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
import xdot_common as X
n_dev = min(2, D._ffi.lib().demcmc_device_count())
rng = np.random.default_rng(5)
n, dm, G, Np = 900, 50, 4, 16
x = rng.normal(rng.normal(size=dm), 1.0, size=(n, dm))
prior = [("normal", 0, 1)] * dm + [("halfcauchy", 0, 1)]
lo = [-np.inf] * dm + [0.0]; hi = [np.inf] * (dm + 1)
theta0 = np.column_stack([rng.normal(size=(G * Np, dm)), np.abs(rng.standard_cauchy(G * Np)) + 0.3])
with D.Handle(G, Np, dm + 1, lo, hi, seed=3, burnin=2, theta_snooker=0.1, alpha=0.6, store_every=2, devices=list(range(n_dev)) if n_dev > 1 else None) as h:
    h.set_model("mvnormal", prior, x=x)
    h.set_state(theta0)
    h.run(24)
    c = h.counters()
    s, ch, (rh, es), mom = h.samples(), h.chains(), h.diagnostics(), h.moments()
    print("devices", n_dev, "cross-device migrations", c["cross_migrations"], "mailbox", c["mailbox_events"], "persistent chunks", c["persistent_chunks"],
          "rows", h.n_rows, "rhat", float(np.nanmax(rh)), "finite", bool(np.isfinite(s).all() and np.isfinite(ch).all()))
A = rng.normal(size=(6, 6)); cov = A @ A.T + 0.5 * np.eye(6)
xf = rng.multivariate_normal(np.zeros(6), cov, size=300)
with D.Handle(2, 8, 7, [-np.inf] * 6 + [0.0], [np.inf] * 7, seed=1, theta_snooker=0.2) as h:
    h.set_model("mvnormal_full", [("normal", 0, 2)] * 6 + [("halfcauchy", 0, 1)], x=xf, cov=cov)
    h.set_state(np.column_stack([rng.normal(size=(16, 6)), np.ones(16)]))
    h.run(8)
    print("mvnormal_full accept rate", h.accept().mean())
print("cross term, given centre: eval", X.eval_error("mvnormal", 130, 50, 33)[0], "run", X.run_error("hier_normal", 50, 120, 2, 12, 4, blocks=np.array([[1, 1] + [0] * 120 + [1], [0, 0] + [1] * 120 + [0]], dtype=np.uint8))[0])

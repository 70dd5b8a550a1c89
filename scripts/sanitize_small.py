"""Small persistent-kernel run for compute-sanitizer (memcheck / racecheck on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
rng = np.random.default_rng(5)
n, dm, G, Np = 1500, 50, 4, 24
x = rng.normal(rng.normal(size=dm), 1.0, size=(n, dm))
prior = [("normal", 0, 1)] * dm + [("halfcauchy", 0, 1)]
lo = [-np.inf] * dm + [0.0]; hi = [np.inf] * (dm + 1)
theta0 = np.column_stack([rng.normal(size=(G * Np, dm)), np.abs(rng.standard_cauchy(G * Np)) + 0.3])
with D.Handle(G, Np, dm + 1, lo, hi, seed=3, burnin=0, theta_snooker=0.1, alpha=0.3) as h:
    h.set_model("mvnormal", prior, x=x)
    h.set_state(theta0)
    h.run(6)
    c = h.counters()
    print("persistent chunks", c["persistent_chunks"], "launches", c["kernel_launches"], "accept rate", h.accept().mean())

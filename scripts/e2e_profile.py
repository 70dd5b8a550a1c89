"""Times the pieces of the public sample() call on the GPU box (diagnostic): python scripts/e2e_profile.py [n_iter]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import demcmc_b200 as D
from demcmc_b200.api import build_handle, _flatten
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 500
x, prior, lo, hi, theta0 = bench.workload(4)
rng = np.random.default_rng(7)
model = D.DEModel(sample_prior=lambda: [rng.normal(size=50), abs(rng.standard_cauchy())], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                  loglike=D.GPULoglike("mvnormal", x), names=("μ", "σ"))
de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=4, Np=256, burnin=0, θsnooker=0.1, seed=11)
for rep in range(3):
    t = [time.perf_counter()]
    h, shapes, d = build_handle(model, de); t.append(time.perf_counter())
    th0 = np.array([_flatten(model.sample_prior()) for _ in range(1024)]); t.append(time.perf_counter())
    h.set_state(th0); t.append(time.perf_counter())
    h.run(n_iter); t.append(time.perf_counter())
    dev_ms = h.counters()["device_ms"]
    ch = h.chains(0, n_iter); t.append(time.perf_counter())
    h.close(); t.append(time.perf_counter())
    names = ["build_handle(create+set_model)", "sample_prior x1024", "set_state", f"run({n_iter})", "chains()", "close"]
    print(rep, {n: round((b - a) * 1e3, 2) for n, a, b in zip(names, t[:-1], t[1:])}, "device_ms", round(dev_ms, 2), flush=True)
